import sys, os, torch
sys.path.insert(0, "/root/repo")
from pmhc_diffusion_model_b200.synthetic import random_params, synthetic_batch
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer
dev = torch.device("cuda:0")
model = Model(16, 22, 1000); model.load_state_dict(random_params(seed=0), strict=True); model = model.to(dev)
dm = DiffusionModelOptimizer(1000, model, 1e-3)
tb = {k: v.to(dev) for k, v in synthetic_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 256, 9, 60, P_pad=80, seed=5000).items()}
model.precision, model.backward_precision = "bf16", "fp16"
for _ in range(2):
    dm.optimize(dict(tb), None)
torch.cuda.synchronize()
