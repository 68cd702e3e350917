"""Shared setup of the example scripts - the `%run ./00_setup` cell of every reference notebook (P1/00, P2/00).

Environment knobs (all optional):  B200DDL_HOME (workspace root), B200DDL_USER, WORKSHOP_IMAGES (dataset size),
WORKSHOP_IMG (image side, default 224), WORKSHOP_SMALL=1 (tiny settings for CPU smoke runs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import b200ddl  # noqa: E402
from b200ddl import Session  # noqa: E402

SMALL = os.environ.get("WORKSHOP_SMALL", "0") == "1"
IMG_HEIGHT = IMG_WIDTH = int(os.environ.get("WORKSHOP_IMG", "32" if SMALL else "224"))
IMG_CHANNELS = 3
N_IMAGES = int(os.environ.get("WORKSHOP_IMAGES", "96" if SMALL else "2048"))

session = Session()
user, my_name, database_name = session.user, session.my_name, session.database_name
DATABRICKS_HOST, DATABRICKS_TOKEN = session.DATABRICKS_HOST, session.DATABRICKS_TOKEN
catalog = session.catalog


def have_gpu() -> bool:
    import torch

    return torch.cuda.is_available()


def decode_mode() -> str:
    """`WORKSHOP_DECODE=gpu`: JPEG rows are decoded on the GPU (nvJPEG + our resize kernel) instead of by CPU workers - the
    path that scales with the number of GPUs per host (examples/part1/04_monitoring_and_optimization.md)."""
    mode = os.environ.get("WORKSHOP_DECODE", "cpu")
    return mode if (mode == "gpu" and have_gpu()) else "cpu"


def default_arch() -> str:
    return "resnet50" if (have_gpu() and IMG_HEIGHT % 32 == 0 and not SMALL) else "mobilenetv2"
