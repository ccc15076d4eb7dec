"""Per-component error of the bf16 tensor-core forward against the CPU oracle (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import egnn_oracle as orc
from pmhc_diffusion_model_b200.diffusion.model import Model

dev = torch.device("cuda:0")
cases = [(4, (8, 12), (50, 70), 80, 61), (5, (1, 16), (0, 40), 40, 62), (2, (8, 15), (300, 400), 400, 63), (150, 9, 60, 80, 64)]
for B, L, Pn, P_pad, seed in cases:
    batch = orc.synthetic_batch(B, L, Pn, P_pad=P_pad, seed=seed)
    params = orc.random_params(seed=seed)
    model = Model(16, 22, 100)
    model.load_state_dict(params, strict=True)
    model = model.to(dev)
    gb = {k: v.to(dev) for k, v in batch.items()}
    with torch.no_grad():
        ref = orc.model_forward(params, orc.batch_to_frames(batch), 42, 100)
        rf = orc.frames_to_tensor7(ref["frames"])
        for prec in ("fp32", "bf16"):
            model.precision = prec
            out = model(dict(gb), 42)
            f = out["frames"].to_tensor_7().cpu()
            t = out["torsions"].cpu()
            m = batch["mask"]
            sel = m & ((m.sum(-1, keepdim=True) - 1 + batch["pocket_mask"].sum(-1, keepdim=True)) > 0)
            dq = (f[..., :4] - rf[..., :4])[sel].abs()
            dx = (f[..., 4:] - rf[..., 4:])[sel].abs()
            dt = (t - ref["torsions"])[sel].abs()
            upd = (rf[..., 4:] - batch["frames"][..., 4:])[sel].abs().max()
            print(f"case B={B} P={P_pad} {prec}: quat {dq.max():.3e} trans {dx.max():.3e} (|x| max {rf[..., 4:][sel].abs().max():.1f}, "
                  f"|update| max {upd:.2f}) tors {dt.max():.3e}  mean trans err {dx.mean():.3e}")
            if prec == "bf16":
                w = dx.max(-1).values
                k = int(w.argmax())
                idx = sel.nonzero()[k]
                print("   worst row", idx.tolist(), "err", dx[k].tolist(), "ref", rf[idx[0], idx[1], 4:].tolist())
