"""Part 2 / 00_setup  (reference: `Part 2 - Distributed Tuning & Inference/00_setup.py`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import *  # noqa: F401,F403

if __name__ == "__main__":
    print("user          :", user)
    print("my_name       :", my_name)
    print("database_name :", database_name)
    print("tracking uri  :", DATABRICKS_HOST)
