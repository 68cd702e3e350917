"""ETL layer (reference 01_data_prep.py): reader -> bronze -> silver -> split -> label index -> tables."""
import numpy as np
import pandas as pd
import pytest

from b200ddl import Session
from b200ddl.data import FLOWER_CLASSES, Table, col, pandas_udf, read_binary_files, synthetic_images


@pytest.fixture()
def session(tmp_path):
    return Session(user="first.last@example.com", root=str(tmp_path))


def test_session_namespace(session):
    assert session.my_name == "first_last"                                     # reference P1/00:6
    assert session.database_name == "distributed_dl_workshop_first_last"       # reference P1/00:9
    assert session.tracking_uri.endswith("mlruns")


def test_binary_reader_and_schema(tmp_path):
    from PIL import Image

    for lab in ("daisy", "roses"):
        d = tmp_path / "flowers" / lab
        d.mkdir(parents=True)
        for i in range(3):
            Image.fromarray(np.random.randint(0, 255, (8, 8, 3), dtype=np.uint8)).save(d / f"{i}.jpg")
        (d / "notes.txt").write_text("x")
    t = read_binary_files(str(tmp_path / "flowers"), glob="*.jpg", recursive=True)
    assert t.columns == ["path", "modificationTime", "length", "content"]        # reference P1/01:50-53
    assert t.count() == 6
    rows = t.collect()
    assert all(r.length == len(r.content) for r in rows)


def test_full_etl_pipeline(session):
    cat = session.catalog
    raw = synthetic_images(60, size=(16, 16), jpeg=True, seed=1)
    df = raw.sample(fraction=0.5, seed=3)                                        # reference P1/01:65
    assert 10 < df.count() < 50
    cat.sql(f"DROP DATABASE IF EXISTS {session.database_name} CASCADE")           # reference P1/01:84
    cat.sql(f"CREATE DATABASE {session.database_name}")                           # reference P1/01:87
    cat.conf.set("spark.sql.parquet.compression.codec", "uncompressed")           # reference P1/01:92
    df.write.format("delta").mode("overwrite").saveAsTable(f"{session.database_name}.bronze")
    bronze = cat.table(f"{session.database_name}.bronze")
    assert bronze.count() == df.count()

    @pandas_udf("string")
    def get_label_udf(path: pd.Series) -> pd.Series:                              # reference P1/01:125-127
        return path.map(lambda p: p.split("/")[-2])

    silver = bronze.withColumn("label", get_label_udf(col("path")))
    silver.write.format("delta").mode("overwrite").saveAsTable(f"{session.database_name}.silver")
    train, val = silver.randomSplit([0.9, 0.1], seed=42)                          # reference P1/01:162
    assert train.count() + val.count() == silver.count()
    labels = sorted(r.label for r in train.select("label").distinct().collect())  # reference P1/01:179
    label_to_idx = {l: i for i, l in enumerate(labels)}
    assert set(labels) <= set(FLOWER_CLASSES)

    @pandas_udf("int")
    def get_label_idx_udf(labels_series: pd.Series) -> pd.Series:                 # reference P1/01:187-189
        return labels_series.map(lambda x: label_to_idx[x])

    train = train.withColumn("label_idx", get_label_idx_udf(col("label")))
    train.write.format("delta").mode("overwrite").saveAsTable(f"{session.database_name}.silver_train")
    back = cat.table(f"{session.database_name}.silver_train").select(["content", "label_idx"])
    assert back.columns == ["content", "label_idx"]
    assert back.to_pandas()["label_idx"].between(0, len(labels) - 1).all()
    # overwrite creates a new version and keeps one set of files
    train.limit(3).write.mode("overwrite").saveAsTable(f"{session.database_name}.silver_train")
    assert cat.table(f"{session.database_name}.silver_train").count() == 3
    assert [c["version"] for c in cat.table_history(f"{session.database_name}.silver_train")] == [0, 1]
    with pytest.raises(FileExistsError):
        train.write.saveAsTable(f"{session.database_name}.silver_train")


def test_split_is_deterministic_and_disjoint():
    t = Table.from_pydict({"id": list(range(1000))})
    a1, b1 = t.random_split([0.9, 0.1], seed=42)
    a2, b2 = t.random_split([0.9, 0.1], seed=42)
    ia, ib = set(a1.to_pandas()["id"]), set(b1.to_pandas()["id"])
    assert ia == set(a2.to_pandas()["id"]) and ib == set(b2.to_pandas()["id"])
    assert not (ia & ib) and len(ia | ib) == 1000
    assert 60 < len(ib) < 140


def test_repartition_and_limit_and_udf_partitions():
    t = Table.from_pydict({"x": list(range(10))}).repartition(3)
    assert [p.num_rows for p in t.partitions()] == [3, 4, 3]
    assert t.limit(4).count() == 4

    @pandas_udf("long")
    def double(x):
        return x * 2

    out = t.with_column("y", double("x"), parallelism=3)
    assert out.to_pandas()["y"].tolist() == [2 * i for i in range(10)]
