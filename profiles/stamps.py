"""Per-phase clock64 stamps of the tensor-core pair kernel (CTA 0, engine 0) — development aid."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmhc_diffusion_model_b200 import synthetic as orc
from pmhc_diffusion_model_b200 import _lib
from pmhc_diffusion_model_b200.diffusion.model import Model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device("cuda:0")
lib = _lib.load()
model = Model(16, 22, 100)
model.load_state_dict(orc.random_params(seed=0), strict=True)
model = model.to(dev)
model.precision = "bf16"
batch = {k: v.to(dev) for k, v in orc.synthetic_batch(B, 9, 60, P_pad=80, seed=1).items()}
buf = torch.zeros(256, dtype=torch.int64, device=dev)
with torch.no_grad():
    for t in (90, 50):
        model(batch, t)
    lib.pmhc_debug_set_stamps.argtypes = [ctypes.c_void_p]
    lib.pmhc_debug_set_stamps(ctypes.c_void_p(buf.data_ptr()))
    model(batch, 10)
    torch.cuda.synchronize()
    lib.pmhc_debug_set_stamps(ctypes.c_void_p(0))
st = buf.cpu().tolist()
names = ["req+extras", "waitM", "sums+ep1+req", "decode", "stage_ldst", "stage_fence", "waitX", "ep2att+req", "waitY", "ep2rot+waitX+req", "waitZ", "ep2tor+waitX+req",
         "ep2trn+waitZ+req", "waitM", "ep3"]
NS = len(names) + 1
for layer in (0, 1):
    s = st[128 * layer:128 * layer + 128]
    n = s[127]
    s = s[:n]
    print(f"layer {layer + 1}: {n} stamps, total {s[-1] - s[0]} cycles")
    print("  kernel start -> first complex:", s[1] - s[0], " setup:", s[2] - s[1])
    # tiles: groups of 15 stamps starting at index 3
    i = 3
    tile = 0
    while i + NS - 1 < n and tile < 5:
        d = [s[i + k + 1] - s[i + k] for k in range(NS - 1)]
        print(f"  tile {tile}: total {s[i + NS - 1] - s[i]:6d} | " + " ".join(f"{nm}={v}" for nm, v in zip(names, d)))
        i += NS
        tile += 1
    print("  remaining stamps deltas:", [s[k + 1] - s[k] for k in range(i - 1, min(n - 1, i + 12))])
