"""Part 2 / 02_hyperopt_distributed_model  (reference: 02_hyperopt_distributed_model.py).

Sequential trials on the driver (default `Trials`), each trial launching a distributed training job through
Runner(np=HVD_NUM_PROCESSES); rank 0 logs a nested child run; checkpoints per epoch on rank 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import *  # noqa: F401,F403
from b200ddl import optim, tracking
from b200ddl import parallel as hvd
from b200ddl.hpo import STATUS_OK, fmin, hp, space_eval, tpe
from b200ddl.loader import make_converter
from b200ddl.models import build_model
from b200ddl.parallel import Runner
from b200ddl.train import ModelCheckpoint, ReduceLROnPlateau, Trainer

NUM_CLASSES = 5                                                            # reference :60
EPOCHS = 1 if SMALL else 3                                                 # reference :63
HVD_NUM_PROCESSES = int(os.environ.get("HVD_NP", "2"))                     # reference :70
MAX_EVALS = int(os.environ.get("MAX_EVALS", "2" if SMALL else "4"))        # reference :357
ARCH = default_arch()
DECODE = decode_mode()   # WORKSHOP_DECODE=gpu: nvJPEG + our resize kernel instead of CPU decode workers
checkpoint_dir = os.path.join(session.checkpoint_root, str(time.time()))   # reference :66-67
os.makedirs(checkpoint_dir, exist_ok=True)

train_df = catalog.table(f"{database_name}.silver_train").select(["content", "label_idx"])
val_df = catalog.table(f"{database_name}.silver_val").select(["content", "label_idx"])
converter_train = make_converter(train_df, session.cache_dir)              # reference :105-112
converter_val = make_converter(val_df, session.cache_dir)
train_size, val_size = len(converter_train), len(converter_val)
tracking.set_experiment(f"/Users/{user}/distributed_dl_workshop")
tracking_uri = tracking.get_tracking_uri()


def train_and_evaluate_hvd(learning_rate=0.001, dropout=0.5, batch_size=32, checkpoint_dir=None):   # reference :161-262
    hvd.init()
    tracking.set_tracking_uri(tracking_uri)
    model = build_model(IMG_HEIGHT, IMG_WIDTH, 3, NUM_CLASSES, dropout=dropout, arch=ARCH, batch_size=batch_size)
    optimizer = hvd.DistributedOptimizer(optim.Adam(learning_rate=learning_rate * hvd.size()))
    param_str = f"learning_rate_{learning_rate:.5g}_dropout_{dropout:.3g}_batch_size_{batch_size}"
    callbacks = [hvd.callbacks.BroadcastGlobalVariablesCallback(0), hvd.callbacks.MetricAverageCallback(),
                 hvd.callbacks.LearningRateWarmupCallback(initial_lr=learning_rate * hvd.size(), warmup_epochs=5),
                 ReduceLROnPlateau(monitor="val_loss", patience=10)]
    if hvd.rank() == 0 and checkpoint_dir:                                                   # :206-211
        callbacks.append(ModelCheckpoint(os.path.join(checkpoint_dir, param_str, "checkpoint-{epoch}.ckpt"),
                                         save_weights_only=True))
    trainer = Trainer(model).compile(optimizer=optimizer, loss="sparse_categorical_crossentropy", metrics=["accuracy"])
    with converter_train.make_dataset(batch_size=batch_size, cur_shard=hvd.rank(), shard_count=hvd.size(),
                                      image_size=(IMG_HEIGHT, IMG_WIDTH), decode=DECODE) as train_ds, \
         converter_val.make_dataset(batch_size=batch_size, cur_shard=hvd.rank(), shard_count=hvd.size(),
                                    image_size=(IMG_HEIGHT, IMG_WIDTH), decode=DECODE) as val_ds:
        steps_per_epoch = max(1, train_size // (batch_size * hvd.size()))
        validation_steps = max(1, val_size // (batch_size * hvd.size()))
        hist = trainer.fit(train_ds, steps_per_epoch=steps_per_epoch, epochs=EPOCHS, verbose=1,
                           validation_data=val_ds, validation_steps=validation_steps, callbacks=callbacks)
    val_loss, val_accuracy = hist.history["val_loss"][-1], hist.history["val_accuracy"][-1]
    if hvd.rank() == 0:                                                                      # :241-260
        with tracking.start_run(run_id=MLFLOW_PARENT_RUN_ID):
            with tracking.start_run(run_name=param_str, nested=True):
                tracking.log_params({"epochs": EPOCHS, "batch_size": batch_size, "learning_rate": learning_rate,
                                     "dropout": dropout, "checkpoint_dir": checkpoint_dir})
                tracking.log_metrics({"val_loss": val_loss, "val_accuracy": val_accuracy})
                tracking.keras.log_model(trainer, "model")
    return val_loss, val_accuracy


# One gang of rank processes for ALL trials (HVD_PERSISTENT=0 restores a fresh HorovodRunner per trial, as in the reference):
# CUDA contexts, loaded extensions, the NCCL group and the symmetric flag buffers are created once; a trial costs shipping
# the pickled function + building its model.  `hr.last_timing` has the per-trial bootstrap breakdown.
PERSISTENT = os.environ.get("HVD_PERSISTENT", "1") == "1"
hr = Runner(np=HVD_NUM_PROCESSES, driver_log_verbosity=os.environ.get("HVD_LOGS", "all"), persistent=PERSISTENT)
trial_timings = []


def objective_function(params):                                                              # reference :294-309
    t0 = time.time()
    loss, acc = hr.run(train_and_evaluate_hvd, learning_rate=params["learning_rate"], dropout=params["dropout"],
                       batch_size=params["batch_size"], checkpoint_dir=checkpoint_dir)
    trial_timings.append({"trial_s": time.time() - t0, **hr.last_timing})
    return {"loss": loss, "status": STATUS_OK}                                               # minimise val_loss


batch_choices = [8, 16] if SMALL else [32, 64, 128]
search_space = {"learning_rate": hp.loguniform("learning_rate", -5, 0),                      # reference :322-326
                "dropout": hp.uniform("dropout", 0.1, 0.9),
                "batch_size": hp.choice("batch_size", batch_choices)}

with tracking.start_run(run_name="hyperopt_horovod_tuning") as parent_run:                   # reference :349-352
    MLFLOW_PARENT_RUN_ID = parent_run.info.run_id
    # default Trials => trials run one after another on the driver, so each may launch a distributed job (:342-344)
    best_hyperparam = fmin(fn=objective_function, space=search_space, algo=tpe.suggest, max_evals=MAX_EVALS)
    tracking.log_params({"best_" + k: v for k, v in best_hyperparam.items()})
hr.close()
import json
print("HPO_TIMING " + json.dumps({"np": HVD_NUM_PROCESSES, "trials": len(trial_timings), "persistent_ranks": PERSISTENT,
                                  "wall_s": sum(t["trial_s"] for t in trial_timings), "per_trial": trial_timings,
                                  "arch": ARCH, "epochs": EPOCHS}))
print("best:", best_hyperparam, "->", space_eval(search_space, best_hyperparam))
print("checkpoints:", session.fs.ls(checkpoint_dir))                                        # reference :380

runs = tracking.search_runs(filter_string=f'tags.mlflow.parentRunId = "{parent_run.info.run_id}"',
                            order_by=["metrics.val_accuracy DESC"])                          # :394-399 (Q3 fixed)
best_run_id = runs.iloc[0]["run_id"]
registry_model_name = my_name + "_flower_classifier"
mv = tracking.register_model(f"runs:/{best_run_id}/model", registry_model_name)
tracking.MlflowClient().transition_model_version_stage(registry_model_name, mv.version, stage="Production")
converter_train.delete()
converter_val.delete()
