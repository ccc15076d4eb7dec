"""A few eager training steps in the bf16 mode (tcgen05 forward + tcgen05 backward), B = 256: the command the launch list and the
ncu capture of the backward kernel are taken from.  Usage: python profiles/train_step_once.py [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmhc_diffusion_model_b200.synthetic import random_params, synthetic_batch
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
model = Model(16, 22, 1000); model.load_state_dict(random_params(seed=0), strict=True); model = model.to(dev)
dm = DiffusionModelOptimizer(1000, model, 1e-3)
tb = {k: v.to(dev) for k, v in synthetic_batch(256, 9, 60, P_pad=80, seed=5000).items()}
model.precision = "bf16"
for _ in range(steps):
    dm.optimize(dict(tb), None)
torch.cuda.synchronize()
dm.check_nan()
print("ok")
