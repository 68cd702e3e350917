"""Part 2 / 01_hyperopt_single_machine_model  (reference: 01_hyperopt_single_machine_model.py).

Parallel single-device trials: `fmin(..., trials=ParallelTrials(parallelism=4))` (SparkTrials), every trial a nested
child run, then best-run selection -> registry -> Production -> reload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import *  # noqa: F401,F403
import numpy as np
import torch
from b200ddl import optim, tracking
from b200ddl.hpo import STATUS_OK, ParallelTrials, fmin, hp, space_eval, tpe
from b200ddl.models import build_model, decode_image
from b200ddl.train import Trainer

NUM_CLASSES = 5   # reference :50
BATCH_SIZE = 32   # reference :52
EPOCHS = 1 if SMALL else 3
NUM_EVALS = int(os.environ.get("NUM_EVALS", "4" if SMALL else "20"))   # reference :226

train_pdf = catalog.table(f"{database_name}.silver_train").select(["content", "label_idx"]).toPandas()  # :76-77
val_pdf = catalog.table(f"{database_name}.silver_val").select(["content", "label_idx"]).toPandas()


def to_arrays(pdf):
    x = torch.from_numpy(np.stack([decode_image(c, (IMG_HEIGHT, IMG_WIDTH)) for c in pdf["content"]]))
    return x, torch.from_numpy(pdf["label_idx"].to_numpy(dtype=np.int64))


TRAIN, VAL = to_arrays(train_pdf), to_arrays(val_pdf)


def batches(arrs, bs):
    x, y = arrs
    bs = max(1, min(bs, len(y)))
    n = len(y) // bs * bs

    class DS:
        def __len__(self):
            return max(1, n // bs)

        def __iter__(self):
            for i in range(0, n, bs):
                yield x[i:i + bs], y[i:i + bs]

    return DS()


def objective_function(params):                                           # reference :133-181
    train_ds, val_ds = batches(TRAIN, BATCH_SIZE), batches(VAL, min(BATCH_SIZE, len(VAL[1])))
    optimizer = optim.get(params["optimizer"])(learning_rate=params["learning_rate"])    # getattr(optimizers, name)
    model = build_model(IMG_HEIGHT, IMG_WIDTH, 3, NUM_CLASSES, dropout=params["dropout"], arch="mobilenetv2")
    trainer = Trainer(model).compile(optimizer=optimizer, loss="sparse_categorical_crossentropy", metrics=["accuracy"])
    trainer.fit(train_ds, steps_per_epoch=len(train_ds), epochs=EPOCHS, verbose=0)
    _, accuracy = trainer.evaluate(val_ds)
    tracking.log_metric("accuracy", accuracy)
    tracking.keras.log_model(trainer, "model")
    return {"loss": -accuracy, "status": STATUS_OK}                         # minimise -accuracy (:181)


search_space = {"optimizer": hp.choice("optimizer", ["Adadelta", "Adam"]),   # reference :194-198
                "learning_rate": hp.loguniform("learning_rate", -5, 0),
                "dropout": hp.uniform("dropout", 0.1, 0.9)}

tracking.set_experiment(f"/Users/{user}/distributed_dl_workshop")          # reference :221
with tracking.start_run(run_name="hyperopt_tuning") as parent_run:
    # SparkTrials(parallelism=4), :226 - on a GPU box every slot is a worker PROCESS pinned to GPU slot % device_count
    trials = ParallelTrials(parallelism=4, executor=os.environ.get("HPO_EXECUTOR", "auto"))
    best_hyperparam = fmin(fn=objective_function, space=search_space, algo=tpe.suggest, trials=trials,
                           max_evals=NUM_EVALS)
    tracking.log_params({"best_" + k: v for k, v in best_hyperparam.items()})
print("best (choice = index, Q7):", best_hyperparam, "->", space_eval(search_space, best_hyperparam))

hyperopt_runs = tracking.search_runs(filter_string=f'tags.mlflow.parentRunId = "{parent_run.info.run_id}"',
                                     order_by=["metrics.accuracy DESC"])    # reference :254-258
best_run_id = hyperopt_runs.iloc[0]["run_id"]
registry_model_name = my_name + "_flower_classifier"                       # reference :279
model_version = tracking.register_model(f"runs:/{best_run_id}/model", registry_model_name)
tracking.MlflowClient().transition_model_version_stage(registry_model_name, model_version.version, stage="Production")
model = tracking.keras.load_model(f"models:/{registry_model_name}/production")   # reference :298
model.summary()
