"""Development aid (GPU): at the bench size on the shipped weights, how far are the tc32 and fp32 (FFMA) outputs from the float64
evaluation of the same formulas (oracle) on the complexes where the two modes differ most?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import egnn_oracle as orc
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.synthetic import synthetic_batch
from tests.helpers import load_case, rel_err

dev = torch.device("cuda:0")
params = load_case("fwd_shipped_p80.pt")["params"]
batch = synthetic_batch(1000, 9, 60, P_pad=80, seed=1)
model = Model(16, 22, 100); model.load_state_dict(params, strict=True); model = model.to(dev)
gb = {k: v.to(dev) for k, v in batch.items()}
outs = {}
with torch.no_grad():
    for mode in ("fp32", "tc32"):
        model.precision = mode
        o = model(dict(gb), 37)
        outs[mode] = (o["frames"].to_tensor_7().cpu(), o["torsions"].cpu())
m = batch["mask"]
mf = m[:, :, None, None].float()
d = ((outs["tc32"][1] - outs["fp32"][1]).abs() * mf).flatten(1).amax(1)
worst = torch.topk(d, 8).indices
print("worst tc32-fp32 torsion differences:", d[worst])
sub = {k: v[worst] for k, v in batch.items()}
p64, b64 = orc.to_float64(params, orc.batch_to_frames(sub))
with torch.no_grad():
    o64 = orc.model_forward(p64, b64, 37, 100)
    o32 = orc.model_forward(params, orc.batch_to_frames(sub), 37, 100)
t64 = o64["torsions"].float()
mm = sub["mask"]
for mode in ("fp32", "tc32"):
    e = (outs[mode][1][worst] - t64).abs() * mf[worst]
    print(mode, "vs float64: per-complex max torsion error", e.flatten(1).amax(1))
print("oracle fp32 (torch CPU) vs float64:", ((o32["torsions"] - t64).abs() * mf[worst]).flatten(1).amax(1))
