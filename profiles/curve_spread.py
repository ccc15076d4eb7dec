"""Spread of the loss-curve deviation from an fp32 run (worst 25-step window mean over 150 steps, B = 64) per training mode, six runs each:
what sets the gates of tests/test_gpu_parity.py::test_tensor_core_training_modes_track_the_fp32_loss_curve.
Usage: python profiles/curve_spread.py"""
import sys, torch, random
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import egnn_oracle as orc
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer as DMO
dev = torch.device("cuda:0")
T, B, steps, lr = 1000, 64, 150, 1e-3
batch = orc.synthetic_batch(B, (8, 12), (40, 60), P_pad=80, seed=404)
params = orc.random_params(seed=51)
gb = {k: v.to(dev) for k, v in batch.items()}
rng = random.Random(9)
ts = [rng.randint(0, T - 1) for _ in range(steps)]
def run(fwd, bwd):
    model = Model(16, 22, T); model.load_state_dict(params, strict=True); model = model.to(dev)
    model.precision, model.backward_precision = fwd, bwd
    dm = DMO(T, model, lr); dm.use_graph = True
    losses = []
    for k, t in enumerate(ts):
        dm.optimize(dict(gb), None, t=t, noise_key=1000 + k)
        losses.append(dm.last_losses["total loss"].mean().clone())
    return torch.stack(losses).cpu().view(-1, 25).mean(dim=1)
ref = run("fp32", None)
for fwd, bwd in (("fp32", None), ("bf16", "fp16"), ("bf16", "bf16"), ("bf16", "fp32"), ("tc32", "fp16")):
    devs = []
    for _ in range(6):
        c = run(fwd, bwd)
        devs.append(float(((c - ref).abs() / ref).max()))
    print(fwd, bwd, " ".join("%.4f" % d for d in devs))
