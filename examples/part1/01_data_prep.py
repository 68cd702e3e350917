"""Part 1 / 01_data_prep  (reference: 01_data_prep.py) - images -> bronze -> silver -> train/val tables.

There is no /databricks-datasets here, so the "flower_photos" directory is synthesised (JPEG files in five class
folders) and then read back with the binary-file reader exactly like the reference does."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import *  # noqa: F401,F403
import pandas as pd
from b200ddl.data import col, pandas_udf, read_binary_files, synthetic_images

# -- materialise a flower_photos-like directory tree (stand-in for the mounted dataset)
photos_dir = os.path.join(session.root, "databricks-datasets", "flower_photos")
if not os.path.isdir(photos_dir):
    for r in synthetic_images(N_IMAGES, size=(IMG_HEIGHT, IMG_WIDTH), jpeg=True, seed=0, root=photos_dir).collect():
        os.makedirs(os.path.dirname(r.path), exist_ok=True)
        with open(r.path, "wb") as f:
            f.write(r.content)

# -- binaryFile reader + 50 % sample (reference :61-66)
df = read_binary_files(photos_dir, glob="*.jpg", recursive=True).sample(fraction=0.5, seed=0)
df.display(5)
print("images:", df.count())

# -- database lifecycle + bronze (reference :84-104)
catalog.sql(f"DROP DATABASE IF EXISTS {database_name} CASCADE")
catalog.sql(f"CREATE DATABASE {database_name}")
catalog.conf.set("spark.sql.parquet.compression.codec", "uncompressed")
df.write.format("delta").mode("overwrite").saveAsTable(f"{database_name}.bronze")
bronze_df = catalog.table(f"{database_name}.bronze")


# -- label from the parent directory -> silver (reference :125-136)
@pandas_udf("string")
def get_label_udf(paths: pd.Series) -> pd.Series:
    return paths.map(lambda path: path.split("/")[-2])


silver_df = bronze_df.withColumn("label", get_label_udf(col("path")))
silver_df.write.format("delta").mode("overwrite").saveAsTable(f"{database_name}.silver")

# -- split + label index (reference :152-197)
train_df, val_df = silver_df.randomSplit([0.9, 0.1], seed=42)
labels = sorted(r.label for r in train_df.select("label").distinct().collect())
label_to_idx = {label: index for index, label in enumerate(labels)}
print(label_to_idx)


@pandas_udf("int")
def get_label_idx_udf(labels_series: pd.Series) -> pd.Series:
    return labels_series.map(lambda x: label_to_idx[x])


train_df = train_df.withColumn("label_idx", get_label_idx_udf(col("label")))
val_df = val_df.withColumn("label_idx", get_label_idx_udf(col("label")))

# -- write the training tables (reference :213-222)
train_df.write.format("delta").mode("overwrite").saveAsTable(f"{database_name}.silver_train")
val_df.write.format("delta").mode("overwrite").saveAsTable(f"{database_name}.silver_val")
print("silver_train:", catalog.table(f"{database_name}.silver_train").count(),
      "silver_val:", catalog.table(f"{database_name}.silver_val").count())
