#!/usr/bin/env python
"""Launch duration of the pair kernels vs the number of complexes (development aid): separates the fixed cost of a
launch (prologue, tail) from the per-round cost.   python profiles/sweep_b.py [B ...]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmhc_diffusion_model_b200 import synthetic as orc
from pmhc_diffusion_model_b200 import _lib
from pmhc_diffusion_model_b200.diffusion.model import Model

Bs = [int(x) for x in sys.argv[1:]] or [2, 148, 296, 592, 888, 1000, 1184]
dev = torch.device("cuda:0")
lib = _lib.load()
model = Model(16, 22, 100)
model.load_state_dict(orc.random_params(seed=0), strict=True)
model = model.to(dev)
model.precision = "bf16"
for B in Bs:
    batch = {k: v.to(dev) for k, v in orc.synthetic_batch(B, 9, 60, P_pad=80, seed=1).items()}
    with torch.no_grad():
        for t in (90, 50, 30):
            model(batch, t)
        torch.cuda.synchronize()
        lib.pmhc_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(20):
            model(batch, 10 + t)
        e1.record()
        torch.cuda.synchronize()
        ms = (ctypes.c_double * 2)(); n = (ctypes.c_int64 * 2)()
        lib.pmhc_profile_read(ms, n)
        lib.pmhc_profile_enable(0)
    print(f"B={B:5d}  pair kernel avg {1e3 * ms[0] / max(1, n[0]):7.1f} us over {n[0]} launches; whole forward {1e3 * e0.elapsed_time(e1) / 20:7.1f} us")
