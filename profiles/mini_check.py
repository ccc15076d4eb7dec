import sys, os, torch
sys.path.insert(0, '/root/repo')
from pmhc_diffusion_model_b200 import synthetic as orc
from pmhc_diffusion_model_b200.diffusion.model import Model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
model = Model(16, 22, 100); model.load_state_dict(orc.random_params(seed=0), strict=True); model = model.to(dev)
batch = {k: v.to(dev) for k, v in orc.synthetic_batch(B, (8, 12), (40, 60), P_pad=80, seed=1).items()}
with torch.no_grad():
    model.precision = "fp32"; ref = model(dict(batch), 37)
    model.precision = "bf16"; out = model(dict(batch), 37)
    torch.cuda.synchronize()
m = batch["mask"]
print("B", B, "max diff frames", (out["frames"].to_tensor_7() - ref["frames"].to_tensor_7())[m].abs().max().item(),
      "tors", (out["torsions"] - ref["torsions"])[m].abs().max().item())
