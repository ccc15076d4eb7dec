#!/usr/bin/env python
"""Short driver for ncu captures: a few denoiser forwards (and optionally training steps) at the bench shapes.

    python profiles/run_hotpath.py [forward|train] [B] [fp32|bf16]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pmhc_diffusion_model_b200 import synthetic as orc
from pmhc_diffusion_model_b200.diffusion.model import Model  # noqa: E402
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "forward"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
precision = sys.argv[3] if len(sys.argv) > 3 else "fp32"
dev = torch.device("cuda:0")
model = Model(16, 22, 100)
model.load_state_dict(orc.random_params(seed=0), strict=True)
model = model.to(dev)
model.precision = precision
batch = {k: v.to(dev) for k, v in orc.synthetic_batch(B, 9, 60, P_pad=80, seed=1).items()}
if mode == "forward":
    with torch.no_grad():
        for t in (90, 50, 10):
            out = model(batch, t)
else:
    dm = DiffusionModelOptimizer(100, model, 1e-3)
    for t in (90, 50, 10):
        dm.optimize(dict(batch), None, t=t)
torch.cuda.synchronize()
print("done", mode, B)
