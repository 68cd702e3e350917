"""Part 2 / 03_pyfunc_distributed_inference  (reference: 03_pyfunc_distributed_inference.py).

One-function training pipeline -> pyfunc artefact (model + image params) -> single-node predict on 10 rows ->
sharded batch inference (`spark_udf`) on 1000 rows, one scoring process per GPU."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import *  # noqa: F401,F403
from dataclasses import dataclass
import numpy as np
import pandas as pd
from b200ddl import optim, pyfunc, tracking
from b200ddl.loader import make_converter
from b200ddl.models import CLASSES, build_model, decode_batch, decode_image
from b200ddl.train import EarlyStopping, Trainer

BATCH_SIZE = 16 if SMALL else 128     # reference :64
EPOCHS = 1 if SMALL else 2            # reference :403
ARCH = default_arch()


@dataclass
class DataCfg:                        # reference :85-94
    train_tbl_name: str
    validation_tbl_name: str
    petastorm_cache_dir: str = session.cache_dir


class FlowerPyFunc(pyfunc.PythonModel):                                    # reference :157-234
    def load_context(self, context):
        with open(context.artifacts["img_params_dict_path"]) as f:
            d = json.load(f)
        self.img_height, self.img_width = d["img_height"], d["img_width"]
        # the serving batch may differ from the training batch (static-shape engine): WORKSHOP_INFER_BATCH, default = training's
        self.model = tracking.keras.load_model(context.artifacts["keras_model_path"],
                                               batch_size=int(os.environ.get("WORKSHOP_INFER_BATCH", "0")) or None)

    def preprocess(self, img_bytes):
        # ONE preprocessing for training and serving (the reference resizes with PIL and skips normalisation here,
        # SURVEY.md Q6); normalisation runs on the device inside the model's first kernel.
        return decode_image(img_bytes, (self.img_height, self.img_width))

    def predict(self, context, model_input: pd.Series) -> np.ndarray:
        # whole-column decode: JPEG/PNG rows on a thread pool; raw fixed-size payloads as one zero-copy view of the
        # column buffer (per-row `self.preprocess` above is what it does row by row)
        arr = decode_batch(model_input, (self.img_height, self.img_width))
        logits = self.model.predict(arr, batch_size=BATCH_SIZE)
        return np.take(CLASSES, np.argmax(logits, axis=1))


def train_model_petastorm_data_ingest(data_cfg, compile_kwargs, fit_kwargs):    # reference :253-377
    tracking.autolog()
    with tracking.start_run(run_name="pyfunc_model_petastorm") as mlflow_run:
        tracking.log_dict({"img_height": IMG_HEIGHT, "img_width": IMG_WIDTH}, "img_params_dict.json")
        batch_size = fit_kwargs.pop("batch_size")
        tracking.log_param("BATCH_SIZE", batch_size)
        train_df = catalog.table(data_cfg.train_tbl_name).select(["content", "label_idx"]).repartition(2)
        val_df = catalog.table(data_cfg.validation_tbl_name).select(["content", "label_idx"]).repartition(2)
        conv_train = make_converter(train_df, data_cfg.petastorm_cache_dir)
        conv_val = make_converter(val_df, data_cfg.petastorm_cache_dir)
        model = build_model(IMG_HEIGHT, IMG_WIDTH, 3, len(CLASSES), arch=ARCH, batch_size=batch_size)
        trainer = Trainer(model).compile(**compile_kwargs)
        with conv_train.make_dataset(batch_size=batch_size, image_size=(IMG_HEIGHT, IMG_WIDTH)) as train_ds, \
             conv_val.make_dataset(batch_size=batch_size, image_size=(IMG_HEIGHT, IMG_WIDTH)) as val_ds:
            history = trainer.fit(train_ds, steps_per_epoch=max(1, len(conv_train) // batch_size),
                                  validation_data=val_ds, validation_steps=max(1, len(conv_val) // batch_size),
                                  **fit_kwargs)
            pyfunc.log_model("pyfunc_model", python_model=FlowerPyFunc(),
                             artifacts={"img_params_dict_path": f"runs:/{mlflow_run.info.run_id}/img_params_dict.json",
                                        "keras_model_path": f"runs:/{mlflow_run.info.run_id}/model"})   # :354-363
            metrics = trainer.evaluate(val_ds, steps=max(1, len(conv_val) // batch_size))
            tracking.log_metrics({"val_" + n: v for n, v in zip(trainer.metrics_names, metrics)})       # :369-371
        conv_train.delete()
        conv_val.delete()
    tracking.autolog(disable=True)
    return mlflow_run, trainer, history


tracking.set_experiment(f"/Users/{user}/distributed_dl_workshop")
data_cfg = DataCfg(f"{database_name}.silver_train", f"{database_name}.silver_val")
compile_kwargs = {"optimizer": optim.Adam(learning_rate=0.001), "loss": "sparse_categorical_crossentropy",
                  "metrics": ["accuracy"]}                                                              # :392-395
callbacks = [EarlyStopping(monitor="val_loss", min_delta=1e-2, patience=3)]                             # :397-401 (used)
fit_kwargs = {"batch_size": BATCH_SIZE, "epochs": EPOCHS, "verbose": 1, "callbacks": callbacks}        # :402-404
mlflow_run, trainer, history = train_model_petastorm_data_ingest(data_cfg, compile_kwargs, fit_kwargs)

# -- single-node inference on 10 rows (reference :440-450)
inference_df = catalog.table(f"{database_name}.silver")
model_uri = f"runs:/{mlflow_run.info.run_id}/pyfunc_model"
loaded_model = pyfunc.load_model(model_uri)
sample_pdf = inference_df.limit(10).toPandas()
print(loaded_model.predict(sample_pdf["content"]))

# -- distributed inference on 1000 rows (reference :466-476)
classify_udf = pyfunc.spark_udf(None, model_uri, result_type="string")
pred_df = inference_df.limit(1000).withColumn("prediction", classify_udf("content")) \
    .select("path", "content", "label", "prediction")
pred_df.display(10)
pdf = pred_df.toPandas()
print(f"scored {len(pdf)} rows with {classify_udf.stats['workers']} worker(s): "
      f"{classify_udf.stats['rows_per_sec']:.0f} rows/s, accuracy {(pdf.label == pdf.prediction).mean():.3f}")

# -- the same UDF at scale (BASELINE.json config 4): WORKSHOP_INFER_IMAGES=1000000 scores a lazily generated table
#    whose fragments are produced and read inside the per-GPU workers; only predictions return to the driver
N_INFER = int(os.environ.get("WORKSHOP_INFER_IMAGES", "0"))
if N_INFER > 0:
    from b200ddl.data import synthetic_scan
    # WORKSHOP_INFER_FRAG_ROWS: rows per generated fragment (a comma-separated list scores the table once per entry)
    #    default: 8192-row fragments for big tables (fewer fragment boundaries: 391 k vs 358 k img/s on 8 GPUs), 4096 otherwise
    default_rows = "8192" if N_INFER >= 500_000 else "4096"
    for frag_rows in [int(v) for v in os.environ.get("WORKSHOP_INFER_FRAG_ROWS", default_rows).split(",")]:
        big = synthetic_scan(N_INFER, size=(IMG_HEIGHT, IMG_WIDTH), num_classes=len(CLASSES), rows_per_fragment=frag_rows)
        scored = big.withColumn("prediction", classify_udf("content")).select("path", "label", "prediction")
        st = dict(classify_udf.stats)
        head = scored.limit(5).toPandas()
        print(head.to_string(index=False))
        print("INFERENCE_STATS " + json.dumps({
            "rows": st["rows"], "workers": st["workers"], "fragments": st["fragments"], "rows_per_fragment": frag_rows,
            "seconds": st["seconds"], "images_per_sec": st["rows_per_sec"], "startup_seconds": st.get("startup_seconds"),
            "serving_batch": int(os.environ.get("WORKSHOP_INFER_BATCH", "0")) or BATCH_SIZE,
            "per_worker": st["per_worker"], "api": "pyfunc.spark_udf over data.synthetic_scan",
            "image": f"{IMG_HEIGHT}x{IMG_WIDTH}x3 uint8", "arch": ARCH, "batch": BATCH_SIZE}))
classify_udf.close()
