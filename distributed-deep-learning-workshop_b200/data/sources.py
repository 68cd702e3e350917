"""Data sources: the binary-file reader (reference P1/01:61-66) and a synthetic flower-photos generator (there is no
network and no /databricks-datasets here)."""
from __future__ import annotations

import fnmatch
import io
import os
from typing import Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa

from .table import Table

FLOWER_CLASSES = ["daisy", "dandelion", "roses", "sunflowers", "tulips"]


def read_binary_files(path: str, glob: str = "*.jpg", recursive: bool = True, num_partitions: int = 4) -> Table:
    """`spark.read.format('binaryFile').option('pathGlobFilter', glob).option('recursiveFileLookup','true').load(path)`
    -> Table(path, modificationTime, length, content)."""
    paths = []
    if recursive:
        for root, _, files in os.walk(path):
            for f in files:
                if fnmatch.fnmatch(f, glob):
                    paths.append(os.path.join(root, f))
    else:
        paths = [os.path.join(path, f) for f in os.listdir(path) if fnmatch.fnmatch(f, glob)]
    paths.sort()
    mt, ln, content = [], [], []
    for p in paths:
        st = os.stat(p)
        with open(p, "rb") as fh:
            b = fh.read()
        mt.append(int(st.st_mtime * 1000))
        ln.append(len(b))
        content.append(b)
    t = pa.table({"path": pa.array(["file:" + p for p in paths], pa.string()),
                  "modificationTime": pa.array(mt, pa.timestamp("ms")), "length": pa.array(ln, pa.int64()),
                  "content": pa.array(content, pa.binary())})
    return Table(t, num_partitions)


def synthetic_images(n: int, classes: Sequence[str] = FLOWER_CLASSES, size: Tuple[int, int] = (224, 224),
                     jpeg: bool = True, seed: int = 0, root: str = "dbfs:/databricks-datasets/flower_photos",
                     num_partitions: int = 4, quality: int = 85) -> Table:
    """JPEG-shaped synthetic dataset with the tf_flowers directory layout `<root>/<label>/<id>.jpg`.

    Each class has its own colour/texture signature so that a classifier can actually learn (loss goes down,
    accuracy goes up in the examples); `jpeg=False` stores raw uint8 HxWx3 bytes (no decode cost)."""
    rng = np.random.default_rng(seed)
    h, w = size
    k = len(classes)
    base = rng.uniform(40, 215, size=(k, 3))
    freq = rng.uniform(1.0, 6.0, size=(k, 2))
    paths, mt, ln, content = [], [], [], []
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for i in range(n):
        c = int(rng.integers(0, k))
        pattern = np.sin(yy * freq[c, 0] * 2 * np.pi / h + rng.uniform(0, 6.28)) * \
            np.cos(xx * freq[c, 1] * 2 * np.pi / w + rng.uniform(0, 6.28))
        img = base[c][None, None, :] + 35.0 * pattern[..., None] + rng.normal(0, 12.0, size=(h, w, 3))
        img = np.clip(img, 0, 255).astype(np.uint8)
        if jpeg:
            from PIL import Image

            buf = io.BytesIO()
            Image.fromarray(img).save(buf, format="JPEG", quality=quality)
            b = buf.getvalue()
        else:
            b = img.tobytes()
        paths.append(f"{root}/{classes[c]}/{i:08d}.jpg")
        mt.append(1_600_000_000_000 + i)
        ln.append(len(b))
        content.append(b)
    t = pa.table({"path": pa.array(paths, pa.string()), "modificationTime": pa.array(mt, pa.timestamp("ms")),
                  "length": pa.array(ln, pa.int64()), "content": pa.array(content, pa.binary())})
    return Table(t, num_partitions)
