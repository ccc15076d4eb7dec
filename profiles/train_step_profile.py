import sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from pmhc_diffusion_model_b200.synthetic import random_params, synthetic_batch
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer
dev = torch.device("cuda:0")
model = Model(16, 22, 1000); model.load_state_dict(random_params(seed=0), strict=True); model = model.to(dev)
model.precision = "bf16"
dm = DiffusionModelOptimizer(1000, model, 1e-3)
tb = {k: v.to(dev) for k, v in synthetic_batch(256, 9, 60, P_pad=80, seed=5000).items()}
for _ in range(3): dm.optimize(dict(tb), None)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(20): dm.optimize(dict(tb), None)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per step {(t1 - t0) / 20 * 1e3:.3f} ms, with final sync {(t2 - t0) / 20 * 1e3:.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    dm.optimize(dict(tb), None)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_stack_n=4).table(sort_by="self_cuda_time_total", row_limit=40, max_name_column_width=40, max_src_column_width=90))
