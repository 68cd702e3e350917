"""Small-shape launch of every hand-written kernel, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from b200ddl import optim
from b200ddl.models.resnet_engine import EngineTrainStep, ResNet50Engine

eng = ResNet50Engine(batch=2, num_classes=5, image_size=64, zero_init_residual=False)
step = EngineTrainStep(eng, optim.SGD(0.01, momentum=0.9), use_graph=False, warmup_steps=0)
x = torch.randint(0, 256, (2, 64, 64, 3), device="cuda", dtype=torch.uint8)
y = torch.randint(0, 5, (2,), device="cuda")
step.load(x, y)
step._captured = True
step.optimizer.begin_step()
step._launch()
torch.cuda.synchronize()
print("loss", eng.loss_and_acc())
from b200ddl.ops import conv as _C

plans = [pl for pl in _C._plans if hasattr(pl, "halo")]
print(f"conv plans exercised: {len(plans)} total, {sum(pl.resident_filter for pl in plans)} resident-filter, "
      f"{sum(pl.halo for pl in plans)} halo-mode")
