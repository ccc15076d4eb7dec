"""Per-tensor gradient error of a backward arithmetic mode against the fp32 FFMA backward and the CPU oracle's autograd
(development tool; the gates live in tests/test_gpu_parity.py).  Usage: python profiles/bwd_modes_check.py [mode] [case]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import egnn_oracle as orc                                   # noqa: E402
from pmhc_diffusion_model_b200.diffusion.model import Model             # noqa: E402
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer as DMO   # noqa: E402
from pmhc_diffusion_model_b200.rigid import Rigid, Rotation             # noqa: E402

DEV = torch.device("cuda:0")
CASES = {"small": (3, (9, 9), (60, 60), 80, 5), "ragged": (5, (2, 16), (3, 50), 50, 41), "groove": (7, (8, 15), (40, 180), 192, 77),
         "big": (2, (8, 15), (150, 400), 400, 13), "one": (1, (9, 9), (60, 60), 80, 5), "tiny": (6, (1, 3), (1, 5), 8, 21), "nopocket": (4, (2, 16), (0, 1), 8, 22)}


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "fp16"
    B, L, Pn, P_pad, seed = CASES[sys.argv[2] if len(sys.argv) > 2 else "small"]
    g = torch.Generator().manual_seed(seed)
    batch = orc.synthetic_batch(B, L, Pn, P_pad=P_pad, seed=seed)
    batch["pocket_features"][:, ::5] += torch.rand(B, batch["pocket_features"][:, ::5].shape[1], 22, generator=g)
    params = orc.random_params(seed=9)
    p_ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    true = orc.gen_noise([B, 16], g)
    pred = orc.model_forward(p_ref, orc.batch_to_frames(batch), 30, 100)
    orc.get_loss(true, pred, batch["mask"], batch["torsions_mask"])["total loss"].mean().backward()
    true_g = {"frames": Rigid(Rotation(quats=true["frames"]["quats"].to(DEV), normalize_quats=False), true["frames"]["trans"].to(DEV)),
              "torsions": true["torsions"].to(DEV)}
    grads = {}
    for m in ("fp32", mode):
        model = Model(16, 22, 100)
        model.load_state_dict(params, strict=True)
        model = model.to(DEV)
        model.precision = "fp32"
        model.backward_precision = m
        gb = {k: v.to(DEV) for k, v in batch.items()}
        out = model(gb, 30)
        DMO.get_loss(true_g, out, gb["mask"], gb["torsions_mask"])["total loss"].mean().backward()
        torch.cuda.synchronize()
        grads[m] = {k: (None if p.grad is None else p.grad.cpu()) for k, p in model.named_parameters()}
    print(f"{'tensor':42s} {'max|ref|':>10s} {'fp32 err':>10s} {mode + ' err':>10s}   (errors relative to the tensor's largest entry)")
    fg, fr = [], []
    for k, ref in p_ref.items():
        if grads[mode][k] is None:
            continue
        sc = float(ref.grad.abs().max()) + 1e-30
        e32 = float((grads["fp32"][k] - ref.grad).abs().max()) / sc
        em = float((grads[mode][k] - ref.grad).abs().max()) / sc
        flag = " <<<" if em > 2e-2 else ""
        print(f"{k:42s} {sc:10.3e} {e32:10.2e} {em:10.2e}{flag}")
        fg.append(grads[mode][k].flatten()); fr.append(ref.grad.flatten())
    fg, fr = torch.cat(fg).double(), torch.cat(fr).double()
    print("cosine", float(fg @ fr / (fg.norm() * fr.norm())), "rel L2", float((fg - fr).norm() / fr.norm()))


if __name__ == "__main__":
    main()
