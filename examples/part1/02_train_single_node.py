"""Part 1 / 02_model_training_single_node  (reference: 02_model_training_single_node.py).

Whole table pulled into driver memory (`toPandas`), decoded once, then `compile` / `fit` on one device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from common import *  # noqa: F401,F403
import numpy as np
import torch
from b200ddl import optim, tracking
from b200ddl.models import build_model, decode_image
from b200ddl.train import Trainer

BATCH_SIZE = 32   # reference :45
EPOCHS = 3        # reference :46

cols_to_keep = ["content", "label_idx"]
train_tbl = catalog.table(f"{database_name}.silver_train").select(cols_to_keep)
val_tbl = catalog.table(f"{database_name}.silver_val").select(cols_to_keep)
num_classes = train_tbl.select("label_idx").distinct().count()
print(f"train: {train_tbl.count()}  val: {val_tbl.count()}  classes: {num_classes}")


def in_memory_dataset(tbl, batch_size):
    """`toPandas()` + `from_tensor_slices(...).map(preprocess).batch(B)` (reference :97-139)."""
    pdf = tbl.toPandas()
    images = torch.from_numpy(np.stack([decode_image(c, (IMG_HEIGHT, IMG_WIDTH)) for c in pdf["content"]]))
    labels = torch.from_numpy(pdf["label_idx"].to_numpy(dtype=np.int64))
    batch_size = min(batch_size, len(labels))  # tiny validation splits still give one batch
    n = len(labels) // batch_size * batch_size

    class DS:
        def __len__(self):
            return n // batch_size

        def __iter__(self):
            for i in range(0, n, batch_size):
                yield images[i:i + batch_size], labels[i:i + batch_size]

    return DS()


train_ds = in_memory_dataset(train_tbl, BATCH_SIZE)
val_ds = in_memory_dataset(val_tbl, BATCH_SIZE)

tracking.set_experiment(f"/Users/{user}/distributed_dl_workshop")
tracking.autolog()                                                            # reference :195
model = build_model(IMG_HEIGHT, IMG_WIDTH, IMG_CHANNELS, num_classes, arch=default_arch(), batch_size=BATCH_SIZE)
trainer = Trainer(model)
trainer.compile(optimizer=optim.Adam(learning_rate=0.001), loss="sparse_categorical_crossentropy", metrics=["accuracy"])
steps_per_epoch = len(train_ds)          # num_rows // batch_size (the reference divides twice)
validation_steps = max(1, len(val_ds))
with tracking.start_run(run_name="single_node"):
    history = trainer.fit(train_ds, steps_per_epoch=steps_per_epoch, epochs=EPOCHS, verbose=1,
                          validation_data=val_ds, validation_steps=validation_steps)
print({k: [round(x, 4) for x in v] for k, v in history.history.items()})
