"""Development aid (GPU): eager vs CUDA-graph replay of the sampling trajectory and of the training step (shipped weights)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer
from pmhc_diffusion_model_b200.synthetic import synthetic_batch

dev = torch.device("cuda:0")
params = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "shipped_params.pt"), map_location="cpu")


def timed(fn, warm, steps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


for mode in sys.argv[1:] or ["tc32", "bf16"]:
    model = Model(16, 22, 100)
    model.load_state_dict(params, strict=True)
    model = model.to(dev)
    model.precision = mode
    dm = DiffusionModelOptimizer(100, model, 0.0)
    dm.sample_seed = 7
    for B in (1000, 64):
        batch = {k: v.to(dev) for k, v in synthetic_batch(B, 9, 60, P_pad=80, seed=1).items()}
        n = dm.gen_noise([B, 16], dev)
        batch["frames"], batch["torsions"] = n["frames"].to_tensor_7(), n["torsions"]
        for g in (False, True):
            dt = timed(lambda: dm.sample(dict(batch), graph=g), 2, 5)
            print(f"sample {mode} B={B} graph={g}: {dt * 1e3:.2f} ms per trajectory, {B / dt:.0f} complexes/s", flush=True)
    tmodel = Model(16, 22, 1000)
    tmodel.load_state_dict(params, strict=True)
    tmodel = tmodel.to(dev)
    tmodel.precision = mode
    for B in (256, 64):
        tb = {k: v.to(dev) for k, v in synthetic_batch(B, 9, 60, P_pad=80, seed=2).items()}
        for g in (False, True):
            tdm = DiffusionModelOptimizer(1000, tmodel, 1e-4)
            tdm.use_graph = g
            dt = timed(lambda: tdm.optimize(dict(tb), None), 5, 40)
            tdm.check_nan()
            print(f"train {mode} B={B} graph={g}: {dt * 1e3:.3f} ms per step, {B / dt:.0f} complexes/s", flush=True)
