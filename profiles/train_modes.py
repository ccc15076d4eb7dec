"""Training-step time by arithmetic mode (B = 256 per GPU, 9-mer, pocket 60 / 80): fp32 FFMA forward + backward, tensor-core
forward + FFMA backward, tensor-core forward + TF32 tensor-core backward.  Prints step time and the backward kernels' share
(CUDA events around the two egnn_layer_backward launches).  Usage: python profiles/train_modes.py [B] [pocket_n] [P_pad]"""
import ctypes, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmhc_diffusion_model_b200 import _lib
from pmhc_diffusion_model_b200.synthetic import random_params, synthetic_batch
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Pn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
P_pad = int(sys.argv[3]) if len(sys.argv) > 3 else 80
dev = torch.device("cuda:0")
lib = _lib.load()
model = Model(16, 22, 1000)
model.load_state_dict(random_params(seed=0), strict=True)
model = model.to(dev)
dm = DiffusionModelOptimizer(1000, model, 1e-3)
tb = {k: v.to(dev) for k, v in synthetic_batch(B, 9, Pn, P_pad=P_pad, seed=5000).items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for fwd, bwd in (("fp32", "fp32"), ("bf16", "bf16"), ("bf16", "fp16"), ("tc32", "fp16")):
    model.precision, model.backward_precision = fwd, bwd
    for _ in range(3):
        dm.optimize(dict(tb), None)
    torch.cuda.synchronize()
    lib.pmhc_profile_enable(1)
    n = 20
    e0.record()
    for _ in range(n):
        flush.zero_()
        dm.optimize(dict(tb), None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    prof_ms = (ctypes.c_double * 2)()
    prof_n = (ctypes.c_int64 * 2)()
    lib.pmhc_profile_read(prof_ms, prof_n)
    lib.pmhc_profile_enable(0)
    dm.check_nan()
    print(f"forward {fwd} backward {bwd}: {ms:.3f} ms/step = {B / ms * 1e3:,.0f} complexes/s; backward kernels "
          f"{prof_ms[1] / max(prof_n[1], 1) * 1e3:.0f} us per launch x {prof_n[1] // n} per step, forward kernels "
          f"{prof_ms[0] / max(prof_n[0], 1) * 1e3:.0f} us x {prof_n[0] // n}", flush=True)
