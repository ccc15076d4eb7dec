"""Flat gradient of the tensor-core backward modes against the fp32 FFMA backward on the reference's SHIPPED checkpoint (attention logits
up to 2.5e3).  Usage: python profiles/shipped_grad_check.py"""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import egnn_oracle as orc
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer as DMO
from pmhc_diffusion_model_b200.rigid import Rigid, Rotation
dev = torch.device("cuda:0")
params = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "shipped_params.pt"))
g = torch.Generator().manual_seed(5)
B = 24
batch = orc.synthetic_batch(B, (8, 13), (40, 60), P_pad=80, seed=77)
true = orc.gen_noise([B, 16], g)
true_g = {"frames": Rigid(Rotation(quats=true["frames"]["quats"].to(dev), normalize_quats=False), true["frames"]["trans"].to(dev)), "torsions": true["torsions"].to(dev)}
grads = {}
for fwd, bwd in (("fp32", "fp32"), ("fp32", "fp16"), ("fp32", "bf16"), ("tc32", "fp16")):
    model = Model(16, 22, 100); model.load_state_dict(params, strict=True); model = model.to(dev)
    model.precision, model.backward_precision = fwd, bwd
    gb = {k: v.to(dev) for k, v in batch.items()}
    out = model(gb, 30)
    DMO.get_loss(true_g, out, gb["mask"], gb["torsions_mask"])["total loss"].mean().backward()
    grads[(fwd, bwd)] = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]).double().cpu()
ref = grads[("fp32", "fp32")]
for k, v in grads.items():
    print(k, "finite", bool(torch.isfinite(v).all()), "cos %.7f" % float(v @ ref / (v.norm() * ref.norm())), "rel L2 %.2e" % float((v - ref).norm() / ref.norm()), "max|g| %.3e" % float(ref.abs().max()))
