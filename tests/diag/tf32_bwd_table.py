"""Per-tensor error table of the TF32 tensor-core backward against the oracle's autograd and against the FFMA backward."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import egnn_oracle as orc
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer as DMO
from pmhc_diffusion_model_b200.rigid import Rigid, Rotation
DEV = torch.device("cuda:0")
B, L, Pn, P_pad, seed = 5, (2, 16), (3, 50), 50, 41
if len(sys.argv) > 1:
    B, L, Pn, P_pad, seed = 7, (8, 15), (40, 180), 192, 77
g = torch.Generator().manual_seed(seed)
batch = orc.synthetic_batch(B, L, Pn, P_pad=P_pad, seed=seed)
params = orc.random_params(seed=9)
p_ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
true = orc.gen_noise([B, 16], g)
pred = orc.model_forward(p_ref, orc.batch_to_frames(batch), 30, 100)
orc.get_loss(true, pred, batch["mask"], batch["torsions_mask"])["total loss"].mean().backward()
true_g = {"frames": Rigid(Rotation(quats=true["frames"]["quats"].to(DEV), normalize_quats=False), true["frames"]["trans"].to(DEV)),
          "torsions": true["torsions"].to(DEV)}
grads = {}
for mode in ("fp32", "bf16"):
    model = Model(16, 22, 100); model.load_state_dict(params, strict=True); model = model.to(DEV)
    model.backward_precision = mode
    gb = {k: v.to(DEV) for k, v in batch.items()}
    out = model(gb, 30)
    DMO.get_loss(true_g, out, gb["mask"], gb["torsions_mask"])["total loss"].mean().backward()
    grads[mode] = {k: (None if p.grad is None else p.grad.cpu()) for k, p in model.named_parameters()}
for k, ref in p_ref.items():
    if grads["bf16"][k] is None:
        continue
    r = ref.grad
    e32 = float((grads["fp32"][k] - r).abs().max()); etc = float((grads["bf16"][k] - r).abs().max())
    print(f"{k:34s} scale {float(r.abs().max()):.3e} l2 {float(r.norm()):.3e}  fp32 err {e32:.2e}  tf32 err {etc:.2e}  rel {etc / max(float(r.abs().max()), 1e-30):.2e}")
