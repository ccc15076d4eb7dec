"""Development aid (GPU): per-phase clock64 timeline of one pair row (CTA 0, engine 0, lane 0 of both thread groups) of the
third-generation pair kernels."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pmhc_diffusion_model_b200 import _lib
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.synthetic import synthetic_batch

mode = sys.argv[1] if len(sys.argv) > 1 else "tc32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda:0")
lib = _lib.load()
params = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "shipped_params.pt"), map_location="cpu")
model = Model(16, 22, 100)
model.load_state_dict(params, strict=True)
model = model.to(dev)
model.precision = mode
batch = {k: v.to(dev) for k, v in synthetic_batch(B, 9, 60, P_pad=80, seed=1).items()}
buf = torch.zeros(1024, dtype=torch.int64, device=dev)
with torch.no_grad():
    model(dict(batch), 50)
    lib.pmhc_debug_set_stamps3.argtypes = [ctypes.c_void_p]
    lib.pmhc_debug_set_stamps3(buf.data_ptr())
    model(dict(batch), 50)
    torch.cuda.synchronize()
    lib.pmhc_debug_set_stamps3(None)
st = buf.cpu().tolist()
names = {1: "complex begin", 2: "setup done", 3: "complex end", 10: "tile begin", 11: "staged", 12: "requested", 13: "H1 ready", 14: "A: att dot + req | B: tor converted + req",
         15: "A: rot converted + req | B: trn ready", 16: "D3 ready", 17: "B: trn dot done", 18: "outputs written", 21: "setup: bulk issued", 22: "setup: cp.async issued", 23: "setup: cp.async landed", 24: "setup: bulk landed", 25: "setup: torsion term", 26: "setup: lists", 27: "setup: barrier", 30: "stage: begin", 39: "next pair decoded", 40: "merge: weights", 41: "merge: barrier", 42: "merge: sums", 43: "merge: maxima", 31: "stage: loads issued", 32: "stage: tile written", 33: "stage: extras", 19: "engine barrier passed", 20: "softmax merged", 50: "A: TRN wait passed", 51: "A: row max + logit stored"}
for layer in range(2):
    for grp in range(2):
        blk = st[layer * 512 + grp * 256: layer * 512 + grp * 256 + 256]
        n = blk[255]
        print(f"== layer {layer + 1} group {'AB'[grp]}: {n} stamps")
        prev = None
        t0 = None
        for v in blk[:min(n, 110)]:
            clk, tag = v >> 8, v & 255
            if t0 is None:
                t0 = clk
            d = 0 if prev is None else clk - prev
            print(f"   {clk - t0:8d} (+{d:6d})  {names.get(tag, tag)}")
            prev = clk
