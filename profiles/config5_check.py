#!/usr/bin/env python
"""Sampling throughput on BASELINE.json configs[4]-like shapes (mixed peptide lengths 8-15, class-II sized pockets padded to
400), for the record: python profiles/config5_check.py [B] [T]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmhc_diffusion_model_b200 import synthetic
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
model = Model(16, 22, T)
model.load_state_dict(synthetic.random_params(0), strict=True)
model = model.to(dev)
dm = DiffusionModelOptimizer(T, model, 0.0)
for name, L, Pn, P in (("class I  (9-mer, pocket 60 / 80)", 9, 60, 80), ("class I  (8-15-mer, pocket 40-80 / 80)", (8, 15), (40, 80), 80),
                       ("full groove (9-mer, pocket 180 / 192)", 9, 180, 192), ("class II (8-15-mer, pocket 100-400 / 400)", (8, 15), (100, 400), 400)):
    batch = {k: v.to(dev) for k, v in synthetic.synthetic_batch(B, L, Pn, P_pad=P, seed=1).items()}
    noise = dm.gen_noise([B, 16], dev)
    batch["frames"], batch["torsions"] = noise["frames"].to_tensor_7(), noise["torsions"]
    for precision in ("bf16", "fp32"):
        model.precision = precision
        if precision == "fp32" and P > 80:
            continue
        dm.sample(dict(batch))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = dm.sample(dict(batch))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ok = bool(torch.isfinite(out["frames"].to_tensor_7()).all())
        print(f"{name:45s} {precision}: {B / dt:9.0f} complexes/s ({dt * 1e3:7.1f} ms per {B}-complex trajectory, T = {T}) finite={ok}")
