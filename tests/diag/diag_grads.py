import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from tests.helpers import load_case
from pmhc_diffusion_model_b200.diffusion.model import Model
from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer
from pmhc_diffusion_model_b200.rigid import Rigid, Rotation
DEV="cuda:0"
case = load_case(sys.argv[1] if len(sys.argv)>1 else "train_shipped_p80.pt")
model = Model(16,22,case["T"]); model.load_state_dict(case["params"]); model=model.to(DEV)
dm = DiffusionModelOptimizer(case["T"], model, 1e-3)
dm.quat_sign_ref = case["zt_quats"].to(DEV)
gb = {k:(v.to(DEV) if isinstance(v,torch.Tensor) else v) for k,v in case["batch"].items()}
gb["frames"] = Rigid.from_tensor_7(gb["frames"])
n=case["noise"]
noise={"frames": Rigid(Rotation(quats=n["q"].to(DEV), normalize_quats=False), n["x"].to(DEV)), "torsions": n["tors"].to(DEV)}
zt = dm.add_noise(gb, noise, case["t"])
pred = model(zt, case["t"])
losses = dm.get_loss(noise, pred, gb["mask"], gb["torsions_mask"])
losses["total loss"].mean().backward()
for k,p in model.named_parameters():
    g=case["grads"][k]
    if g is None: continue
    d=(p.grad.cpu()-g).abs()
    print(f"{k:32s} max|g| {float(g.abs().max()):.3e} maxerr {float(d.max()):.3e}")
    if float(d.max()) > 1e-4*max(1,float(g.abs().max())):
        flat=d.flatten(); top=flat.topk(min(8,flat.numel())).indices
        for ix in top:
            idx=tuple(int(x) for x in torch.unravel_index(ix, d.shape))
            print("     ", idx, "ours %.6e ref %.6e"%(float(p.grad.cpu()[idx]), float(g[idx])))
