"""Development aid (CPU): emulate the operand roundings of the tensor-core pair pipeline on the reference fixtures.

Every contraction is evaluated in float64 with its operands rounded the way a given kernel mode would feed them to
tcgen05 (bf16, a 2-term hi/lo bf16 split ~16 bits, a 3-term split ~24 bits, or exact), so the error of a mode on the
SHIPPED weights can be sized before the kernel exists.  Pipeline = the folded form: heads contract m1 directly
(W_h W2), the message itself is never formed per pair.

    python tests/diag/emulate_split.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F

from oracle import egnn_oracle as orc
from tests.helpers import load_case, rel_err


HALF = [torch.bfloat16]


def bf(x):
    return x.to(torch.float32).to(HALF[0]).to(torch.float64)


def split(x, terms):
    if terms == 0:
        return x.to(torch.float32).to(torch.float64)   # fp32 operand
    out = torch.zeros_like(x)
    r = x.clone()
    for _ in range(terms):
        h = bf(r)
        out = out + h
        r = r - h
    return out


def mm(a, w, ta, tw, cross=True):
    """a [..., K] . w [N, K]^T with operand splits of ta / tw terms.  With 2-term splits the lo.lo product is dropped
    (3 MMAs), with 3-term splits products below 2^-24 are dropped (6 MMAs)."""
    if ta == 0 and tw == 0:
        return a @ w.T
    parts_a, parts_w = [], []
    r = a.clone()
    for _ in range(max(ta, 1)):
        h = bf(r); parts_a.append(h); r = r - h
    r = w.clone()
    for _ in range(max(tw, 1)):
        h = bf(r); parts_w.append(h); r = r - h
    out = 0
    order = max(len(parts_a), len(parts_w))
    for ia, pa in enumerate(parts_a):
        for iw, pw in enumerate(parts_w):
            if ia + iw < order:
                out = out + pa @ pw.T
    return out


def layer(p, pre, q, x, tors, h, mask, ph, pq, px, pmask, cfg, layer1):
    """Factorised + folded EGNN layer in float64 with operand roundings given by cfg."""
    B, N, H = h.shape
    P = ph.shape[1]
    K = N + P
    g = lambda k: p[pre + "." + k].double()
    W1, b1 = g("message_mlp.0.weight"), g("message_mlp.0.bias")
    W2, b2 = g("message_mlp.2.weight"), g("message_mlp.2.bias")
    Hh = ph.shape[-1]
    # node projections (fp32-exact FFMA or split MMA: cfg['proj'])
    Ai = mm(h, W1[:, :H], cfg["proj"], cfg["proj"]) + b1
    Aj_pep = mm(h, W1[:, H:2 * H], cfg["proj"], cfg["proj"])
    Aj_poc = mm(ph, W1[:, H:2 * H], cfg["projp"], cfg["projp"])
    Aj = torch.cat((Aj_pep, Aj_poc), 1)                          # [B,K,64]
    if cfg.get("aj_bf16"):
        Aj = bf(Aj); Ai = bf(Ai)
    We = W1[:, 2 * H:]                                           # [64,31]
    idx = torch.arange(N)
    rel = (N - 1) + (idx[:, None] - idx[None, :])                # [N,N]
    E = torch.zeros(N, K, 64, dtype=torch.float64)
    E[:, :N] = We.T[rel]
    m1 = F.relu(Ai[:, :, None, :] + Aj[:, None, :, :] + E[None])  # [B,N,K,64]
    if cfg.get("m1_bf16"):
        m1 = bf(m1)
    if os.environ.get("STATS"):
        print(f"    {pre}: max m1 {float(m1.max()):.3g}  max|Ai| {float(Ai.abs().max()):.3g} max|Aj| {float(Aj.abs().max()):.3g}")
    # geometry
    qn = torch.cat((q, pq), 1)
    xn = torch.cat((x, px), 1)
    qi = q[:, :, None, :]
    d2 = ((x[:, :, None, :] - xn[:, None]) ** 2).sum(-1)
    qd = ((qi * qn[:, None]).sum(-1)) ** 2
    qinv = orc.quat_inv(qn)[:, None].expand(B, N, K, 4)
    qnb = qn[:, None].expand(B, N, K, 4)
    lq = orc.quat_mul(qinv, orc.quat_mul(qi.expand(B, N, K, 4), qnb))
    not_self = ~torch.eye(N, dtype=torch.bool)
    pair_mask = torch.cat((mask[:, :, None] & mask[:, None, :] & not_self[None], mask[:, :, None] & pmask[:, None, :]), -1)

    def head(name, extra_in, extra_terms):
        W0, b0 = g(name + ".0.weight"), g(name + ".0.bias")
        Wm = W0[:, :64]
        Wf = (Wm @ W2).float().double()           # folded, built in fp32-ish
        bias = (Wm @ b2 + b0)
        tag = name.split("_")[0][:3]
        hid = mm(m1, Wf, cfg.get("m1_" + tag, cfg["m1"]), cfg.get("w_" + tag, cfg["w"])) + bias
        if extra_in is not None:
            hid = hid + mm(extra_in, W0[:, 64:], extra_terms, extra_terms)
        if os.environ.get("STATS"):
            print(f"    {pre}.{name}: max|hid| {float(hid.abs().max()):.3g} max|Wf| {float(Wf.abs().max()):.3g} max|contraction| {float((hid - bias).abs().max()):.3g}")
        return F.relu(hid)

    hatt = head("attention_mlp", torch.stack((-d2, qd), -1), cfg["geo"])
    logits = (hatt.float().double() @ g("attention_mlp.2.weight").T).squeeze(-1) + g("attention_mlp.2.bias")   # fp32 FFMA
    w = torch.softmax(logits - (~pair_mask) * 1e9, -1)
    hrot = head("rotation_mlp", lq, cfg["geo_small"])
    dl = torch.sigmoid(mm(hrot, g("rotation_mlp.2.weight"), cfg.get("h2_rot", cfg["h2"]), cfg.get("w2_rot", cfg["w2"])) + g("rotation_mlp.2.bias"))
    dg = orc.quat_mul(qnb, orc.quat_mul(dl, qinv))
    delta = (dg * w[..., None]).sum(-2)
    has = pair_mask.sum(-1) > 0
    delta = torch.where(has[..., None], delta, delta.new_tensor([1.0, 0, 0, 0]))
    delta = F.normalize(delta, dim=-1)
    nq = orc.quat_mul(delta, q)
    nq = nq / nq.norm(dim=-1, keepdim=True)
    ft = tors.reshape(B, N, 14)
    htor = head("torsion_mlp", ft[:, :, None, :].expand(B, N, K, 14), 0)    # T_t per node in fp32, exact add
    da = mm(htor, g("torsion_mlp.2.weight"), cfg["h2"], cfg["w2"]) + g("torsion_mlp.2.bias")
    da = (da * w[..., None]).sum(-2)
    ntors = orc.sin_cos_mul(orc.angle_to_sin_cos(da), tors)
    htrn = head("translation_mlp", None, 0)
    sc = mm(htrn, g("translation_mlp.2.weight"), cfg["h2t"], cfg["w2t"]) + g("translation_mlp.2.bias")
    nx = x + (sc * (x[:, :, None, :] - xn[:, None]) * w[..., None]).sum(-2)
    o = None
    if layer1:
        S = m1.sum(-2)                                                          # unmasked (T3)
        Msum = mm(S, W2, cfg["node"], cfg["node"]) + K * b2
        f0 = F.relu(mm(torch.cat((h, Msum), -1), g("feature_mlp.0.weight"), cfg["node"], cfg["node"]) + g("feature_mlp.0.bias"))
        o = mm(f0, g("feature_mlp.2.weight"), cfg["node"], cfg["node"]) + g("feature_mlp.2.bias")
    return nq, nx, ntors, o, logits


def forward(case, cfg):
    p = case["params"]
    b = case["batch"]
    t, T = case["t"], case["T"]
    d = lambda v: v.double()
    feats = d(b["features"])
    B, N = b["mask"].shape
    h = torch.cat((feats, torch.full((B, N, 1), t / T, dtype=torch.float64)), -1)
    pf = d(b["pocket_features"])
    ph = torch.cat((pf, pf.new_zeros(B, pf.shape[1], 1)), -1)
    q, x = d(b["frames"][..., :4]), d(b["frames"][..., 4:])
    pq, px = d(b["pocket_frames"][..., :4]), d(b["pocket_frames"][..., 4:])
    mask, pmask = b["mask"].bool(), b["pocket_mask"].bool()
    q1, x1, t1, o1, lg1 = layer(p, "gnn1", q, x, d(b["torsions"]), h, mask, ph, pq, px, pmask, cfg, True)
    i1 = F.relu(o1)
    pi = F.pad(ph, (0, 64 - ph.shape[-1]))
    cfg2 = dict(cfg)
    cfg2["proj"] = cfg["node"]
    q2, x2, t2, _, lg2 = layer(p, "gnn2", q1, x1, t1, i1, mask, pi, pq, px, pmask, cfg2, False)
    return torch.cat((q2, x2), -1), t2, (lg1, lg2)


EXACT = dict(proj=0, projp=0, m1=0, w=0, geo=0, geo_small=0, h2=0, w2=0, h2t=0, w2t=0, node=0)


def variants():
    v = {}
    if os.environ.get("FP16"):
        HALF[0] = torch.float16
        v["fp16 split2 everywhere (geo exact, trn 2nd fp32)"] = dict(proj=0, projp=0, m1=2, w=2, geo=0, geo_small=0, h2=2, w2=2, h2t=0, w2t=0, node=2)
        v["fp16 split2, 2nd layers single fp16"] = dict(proj=0, projp=0, m1=2, w=2, geo=0, geo_small=0, h2=1, w2=1, h2t=0, w2t=0, node=2)
        v["fp16 single"] = dict(proj=0, projp=0, m1=1, w=1, geo=0, geo_small=0, h2=1, w2=1, h2t=1, w2t=1, node=1)
        return v
    if os.environ.get("ABLATE"):
        S3 = dict(proj=0, projp=0, m1=3, w=3, geo=3, geo_small=3, h2=3, w2=3, h2t=0, w2t=0, node=3)
        v["split3 (trn 2nd fp32)"] = dict(S3)
        for k, val in [("m1_att", 2), ("w_att", 2), ("m1_rot", 2), ("w_rot", 2), ("m1_tor", 2), ("w_tor", 2), ("m1_tra", 2), ("w_tra", 2),
                       ("h2", 2), ("w2", 2), ("h2_rot", 2), ("node", 2), ("geo_small", 2)]:
            c = dict(S3); c[k] = val
            v[f"split3 but {k}={val}"] = c
        c = dict(S3); c.update(m1_rot=2, w_rot=2, m1_tor=2, w_tor=2, m1_tra=2, w_tra=2, geo_small=2); v["att split3; other 1st split2; 2nd split3"] = c
        c = dict(c); c.update(h2=2, w2=2); v["att split3; rest split2"] = c
        c = dict(S3); c.update(m1=2, w=2, geo_small=2, m1_att=3, w_att=3, h2=2, w2=3); v["att split3; rest split2 but w2=3"] = c
        c = dict(S3); c.update(m1=2, w=2, geo_small=2, w_att=3, h2=2, w2=3); v["m1 split2 all; w_att 3, w2 3"] = c
        return v
    v["exact(fp64 of factorised form)"] = dict(EXACT)
    v["bf16 all (m1 bf16 arithmetic)"] = dict(proj=0, projp=0, m1=1, w=1, geo=2, geo_small=1, h2=1, w2=1, h2t=1, w2t=1, node=1, m1_bf16=True, aj_bf16=True)
    v["bf16 MMAs, fp32 m1 inputs"] = dict(proj=0, projp=0, m1=1, w=1, geo=3, geo_small=1, h2=1, w2=1, h2t=1, w2t=1, node=2)
    v["split2 everywhere"] = dict(proj=0, projp=0, m1=2, w=2, geo=3, geo_small=2, h2=2, w2=2, h2t=2, w2t=2, node=2)
    v["split2 heads-1st, bf16 2nd layers"] = dict(proj=0, projp=0, m1=2, w=2, geo=3, geo_small=2, h2=1, w2=1, h2t=1, w2t=1, node=2)
    v["split2 1st, 2nd: rot/tor split2, trn fp32"] = dict(proj=0, projp=0, m1=2, w=2, geo=3, geo_small=2, h2=2, w2=2, h2t=0, w2t=0, node=2)
    v["split3 1st, split2 2nd, trn fp32"] = dict(proj=0, projp=0, m1=3, w=3, geo=3, geo_small=3, h2=2, w2=2, h2t=0, w2t=0, node=3)
    v["split3 everywhere"] = dict(proj=0, projp=0, m1=3, w=3, geo=3, geo_small=3, h2=3, w2=3, h2t=3, w2t=3, node=3)
    v["split2, node split3"] = dict(proj=0, projp=0, m1=2, w=2, geo=3, geo_small=2, h2=2, w2=2, h2t=0, w2t=0, node=3)
    v["split2 m1 / split3 w"] = dict(proj=0, projp=0, m1=2, w=3, geo=3, geo_small=2, h2=2, w2=2, h2t=0, w2t=0, node=3)
    return v


if __name__ == "__main__":
    torch.set_num_threads(8)
    names = sys.argv[1:] or ["fwd_shipped_p80.pt", "fwd_shipped_p192.pt", "fwd_random_p96.pt"]
    for name in names:
        case = load_case(name)
        m = case["batch"]["mask"].bool()
        p64, b64 = orc.to_float64(case["params"], orc.batch_to_frames(case["batch"]))
        with torch.no_grad():
            o64 = orc.model_forward(p64, b64, case["t"], case["T"])
        f64 = orc.frames_to_tensor7(o64["frames"])
        floor = max(rel_err(f64.float()[m], case["out_frames"][m]), rel_err(o64["torsions"].float()[m], case["out_torsions"][m]))
        print(f"== {name}: fp32 noise floor of the reference {floor:.2e}; gate max(1e-4, 2 x floor) = {max(1e-4, 2 * floor):.2e}")
        for label, cfg in variants().items():
            with torch.no_grad():
                fr, to, lg = forward(case, cfg)
            ef = rel_err(fr.float()[m], case["out_frames"][m])
            et = rel_err(to.float()[m], case["out_torsions"][m])
            ef64 = rel_err(fr[m], f64[m])
            et64 = rel_err(to[m], o64["torsions"][m])
            print(f"  {label:46s} vs reference: frames {ef:.2e} torsions {et:.2e} | vs float64: frames {ef64:.2e} torsions {et64:.2e}")
