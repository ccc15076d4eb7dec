"""Development aid (GPU): sampled complexes/s of every forward mode on the bench workload (1 000 synthetic 9-mers, pocket 60 of 80,
T = 100) with the SHIPPED weights, and the device time of the fused pair kernels (pmhc_profile_*, a separate pass).

    python profiles/time_modes.py [mode ...]
"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pmhc_diffusion_model_b200 import _lib
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer
from pmhc_diffusion_model_b200.synthetic import synthetic_batch

dev = torch.device("cuda:0")
lib = _lib.load()
modes = sys.argv[1:] or ["tc32", "fp16", "bf16"]
B, T = int(os.environ.get("B", 1000)), 100
P_pad, Pn = int(os.environ.get("P_PAD", 80)), int(os.environ.get("PN", 60))
params = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "shipped_params.pt"), map_location="cpu")
model = Model(16, 22, T)
model.load_state_dict(params, strict=True)
model = model.to(dev)
batch = {k: v.to(dev) for k, v in synthetic_batch(B, 9, Pn, P_pad=P_pad, seed=1).items()}
dm = DiffusionModelOptimizer(T, model, 0.0)
dm.sample_seed = 7
for mode in modes:
    model.precision = mode
    steps = 2 if mode != "fp32" else 1
    for it in range(steps + 1):
        noise = dm.gen_noise([B, 16], dev)
        inp = dict(batch)
        inp["frames"] = noise["frames"].to_tensor_7()
        inp["torsions"] = noise["torsions"]
        if it == 1:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        out = dm.sample(inp)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    lib.pmhc_profile_enable(1)
    out = dm.sample(inp)
    torch.cuda.synchronize()
    ms = (ctypes.c_double * 2)()
    n = (ctypes.c_int64 * 2)()
    lib.pmhc_profile_read(ms, n)
    lib.pmhc_profile_enable(0)
    finite = bool(torch.isfinite(out["frames"].to_tensor_7()).all())
    print(f"{mode:5s} B={B} P={P_pad}: {B / dt:9.0f} complexes/s ({dt * 1e3:.2f} ms per trajectory); pair kernels {ms[0] / max(n[0], 1) * 1e3:.1f} us x {n[0]} "
          f"= {ms[0]:.2f} ms; finite {finite}", flush=True)
