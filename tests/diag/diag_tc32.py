"""Development aid (GPU): error of every forward mode on the reference fixtures, with the layer-1 intermediates of the
fixture (frames, torsions, features after the first EGNN layer) compared too, so a wrong kernel can be localised.

    python tests/diag/diag_tc32.py [mode ...]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oracle import egnn_oracle as orc
from pmhc_diffusion_model_b200 import _lib
from pmhc_diffusion_model_b200.diffusion.model import Model
from tests.helpers import load_case, rel_err

dev = torch.device("cuda:0")
lib = _lib.load()
modes = sys.argv[1:] or ["fp32", "tc32", "fp16", "bf16"]


def run_saved(model, gb, t, T, mode):
    """pmhc_model_forward_ex with a `saved` buffer: returns outputs and the layer-1 intermediates."""
    desc, keep = _lib.make_batch(gb["frames"], gb["torsions"], gb["features"], gb["mask"], gb["pocket_frames"], gb["pocket_features"],
                                 gb["pocket_mask"])
    B, P = desc.B, desc.P
    flat = model._flat_params()
    out_f = torch.empty(B, 16, 7, device=dev)
    out_t = torch.empty(B, 16, 7, 2, device=dev)
    saved = torch.zeros(lib.pmhc_saved_floats(B, P), device=dev)
    nbytes = lib.pmhc_workspace_bytes(B, P)
    ws = _lib.workspace(dev, nbytes)
    _lib.check(lib.pmhc_model_forward_ex(flat.data_ptr(), ctypes.byref(desc), float(t) / T, out_f.data_ptr(), out_t.data_ptr(),
                                         saved.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr(dev), _lib.PRECISIONS[mode]), "forward")
    torch.cuda.synchronize()
    BN = B * 16
    o = 2 * BN * 16
    frames1 = saved[o:o + BN * 7].view(B, 16, 7); o += BN * 7
    tors1 = saved[o:o + BN * 14].view(B, 16, 7, 2); o += BN * 14
    feat1 = saved[o:o + BN * 64].view(B, 16, 64); o += BN * 64
    return out_f.cpu(), out_t.cpu(), frames1.cpu(), tors1.cpu(), feat1.cpu()


for name in ["fwd_random_p96.pt", "fwd_shipped_p80.pt", "fwd_shipped_p192.pt"]:
    case = load_case(name)
    m = case["batch"]["mask"].bool()
    model = Model(16, 22, case["T"])
    model.load_state_dict(case["params"], strict=True)
    model = model.to(dev)
    gb = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in case["batch"].items()}
    p64, b64 = orc.to_float64(case["params"], orc.batch_to_frames(case["batch"]))
    with torch.no_grad():
        o64 = orc.model_forward(p64, b64, case["t"], case["T"])
    floor = max(rel_err(orc.frames_to_tensor7(o64["frames"]).float()[m], case["out_frames"][m]),
                rel_err(o64["torsions"].float()[m], case["out_torsions"][m]))
    print(f"== {name}: gate max(1e-4, 2 x {floor:.2e})", flush=True)
    for mode in modes:
        try:
            of, ot, f1, t1, h1 = run_saved(model, gb, case["t"], case["T"], mode)
        except Exception as e:  # noqa: BLE001
            print(f"  {mode}: FAILED {e}")
            continue
        line = f"  {mode:5s} out: frames {rel_err(of[m], case['out_frames'][m]):.2e} torsions {rel_err(ot[m], case['out_torsions'][m]):.2e}"
        if "l1_frames" in case:
            l1f = case["l1_frames"]
            q_err = min(rel_err(f1[m][:, :4], l1f[m][:, :4]), rel_err(-f1[m][:, :4], l1f[m][:, :4]))
            line += (f" | layer 1: quats {q_err:.2e} trans {rel_err(f1[m][:, 4:], l1f[m][:, 4:]):.2e} torsions {rel_err(t1[m], case['l1_torsions'][m]):.2e}"
                     f" features {rel_err(h1[m], torch.relu(case['l1_features'])[m]):.2e}")
        finite = bool(torch.isfinite(of).all() and torch.isfinite(ot).all())
        print(line + ("" if finite else "  NON-FINITE OUTPUT"), flush=True)
        # inference path (no saved buffer) must agree bitwise with the training-mode forward
        model.precision = mode
        with torch.no_grad():
            out = model(dict(gb), case["t"])
        same = torch.equal(out["frames"].to_tensor_7().cpu()[m], of[m]) and torch.equal(out["torsions"].cpu()[m], ot[m])
        print(f"        inference == training-mode forward: {same}", flush=True)

# edge shapes against the oracle (ragged, empty pockets, big pockets, many complexes)
for B, L, Pn, P_pad, seed in [(5, (1, 16), (0, 40), 40, 21), (3, 16, 80, 80, 22), (2, (8, 15), (300, 400), 400, 23), (64, (8, 15), (50, 80), 80, 24),
                              (700, 9, 60, 80, 25)]:
    batch = orc.synthetic_batch(B, L, Pn, P_pad=P_pad, seed=seed)
    params = orc.random_params(seed=seed)
    model = Model(16, 22, 100)
    model.load_state_dict(params, strict=True)
    model = model.to(dev)
    gb = {k: v.to(dev) for k, v in batch.items()}
    with torch.no_grad():
        ref = orc.model_forward(params, orc.batch_to_frames(batch), 42, 100)
    mm = batch["mask"].bool()
    sel = mm & ((mm.sum(-1, keepdim=True) - 1 + batch["pocket_mask"].sum(-1, keepdim=True)) > 0)
    for mode in modes:
        model.precision = mode
        try:
            with torch.no_grad():
                out = model(dict(gb), 42)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"edge B={B} P={P_pad} {mode}: FAILED {e}")
            continue
        f = out["frames"].to_tensor_7().cpu()
        t = out["torsions"].cpu()
        print(f"edge B={B} P={P_pad} {mode:5s}: frames {rel_err(f[sel], orc.frames_to_tensor7(ref['frames'])[sel]):.2e} "
              f"torsions {rel_err(t[sel], ref['torsions'][sel]):.2e} finite {bool(torch.isfinite(f).all() and torch.isfinite(t).all())}", flush=True)
print("diag done")
