"""Data lake / ETL layer: binary-file reader -> bronze -> silver -> train/val tables."""
from .table import Table, Row, Column, col, pandas_udf, PandasUDF
from .catalog import Catalog, default_catalog, set_default_catalog
from .sources import read_binary_files, synthetic_images, FLOWER_CLASSES
from .scan import ScanTable, Fragment, synthetic_scan, scan_parquet_files

__all__ = ["Table", "Row", "Column", "col", "pandas_udf", "PandasUDF", "Catalog", "default_catalog",
           "set_default_catalog", "read_binary_files", "synthetic_images", "FLOWER_CLASSES", "ScanTable", "Fragment",
           "synthetic_scan", "scan_parquet_files"]
