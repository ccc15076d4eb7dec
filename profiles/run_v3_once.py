"""Profiling target: a few denoiser forwards of one mode on the bench shape with the shipped weights (ncu -k regex:egnn_pair3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.synthetic import synthetic_batch

mode = sys.argv[1] if len(sys.argv) > 1 else "tc32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
params = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "shipped_params.pt"), map_location="cpu")
model = Model(16, 22, 100)
model.load_state_dict(params, strict=True)
model = model.to(dev)
model.precision = mode
batch = {k: v.to(dev) for k, v in synthetic_batch(B, 9, 60, P_pad=80, seed=1).items()}
with torch.no_grad():
    for it in range(reps):
        out = model(dict(batch), 50)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(out["torsions"]).all()))
