"""Run-to-run reproducibility of the tcgen05 backward: twelve identical training steps (lr 0) per shape, largest difference between the
flat gradients.  The only unordered sums are shared-memory atomics on per-node accumulators, so the differences must stay at fp32
rounding (measured <= 3e-6 of the largest entry); anything larger would be a race.  Usage: python profiles/repeat_check.py"""
import sys, torch
sys.path.insert(0, "/root/repo")
from pmhc_diffusion_model_b200.synthetic import random_params, synthetic_batch
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer
dev = torch.device("cuda:0")
for (B, L, Pn, Pp) in ((256, 9, 60, 80), (300, (8, 15), (30, 180), 192), (37, (2, 16), (150, 400), 400)):
    model = Model(16, 22, 1000); model.load_state_dict(random_params(seed=0), strict=True); model = model.to(dev)
    model.precision, model.backward_precision = "bf16", "fp16"
    dm = DiffusionModelOptimizer(1000, model, 0.0)
    tb = {k: v.to(dev) for k, v in synthetic_batch(B, L, Pn, P_pad=Pp, seed=5000).items()}
    gs = []
    cap = {}
    dm.grad_hook = lambda g: cap.__setitem__("g", g.clone())
    for it in range(12):
        dm.optimize(dict(tb), None, t=321, noise_key=77)
        torch.cuda.synchronize()
        gs.append(cap["g"])
    ref = gs[0]
    dev_max = max(float((g - ref).abs().max()) for g in gs[1:])
    print(B, Pp, "max |g| %.3e  max run-to-run difference %.3e  relative %.2e  finite %s" % (float(ref.abs().max()), dev_max, dev_max / float(ref.abs().max()), bool(torch.isfinite(ref).all())))
