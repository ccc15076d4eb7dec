"""Development aid (GPU): is the tc32 pair kernel bound by issue slots or by latency?  One engine per CTA on B complexes against
two engines per CTA on 2 B complexes: equal times = the engines do not slow each other (latency-bound, more tiles in flight
would pay); twice the time = they share a saturated resource."""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from pmhc_diffusion_model_b200 import _lib
    from pmhc_diffusion_model_b200.diffusion.model import Model
    from pmhc_diffusion_model_b200.synthetic import synthetic_batch
    B = int(sys.argv[1])
    dev = torch.device("cuda:0")
    lib = _lib.load()
    params = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "shipped_params.pt"), map_location="cpu")
    model = Model(16, 22, 100); model.load_state_dict(params, strict=True); model = model.to(dev); model.precision = "tc32"
    batch = {k: v.to(dev) for k, v in synthetic_batch(B, 9, 60, P_pad=80, seed=1).items()}
    with torch.no_grad():
        for _ in range(5): model(dict(batch), 50)
        lib.pmhc_profile_enable(1)
        for _ in range(20): model(dict(batch), 50)
        torch.cuda.synchronize()
    ms, n = (ctypes.c_double * 2)(), (ctypes.c_int64 * 2)()
    lib.pmhc_profile_read(ms, n)
    print(f"engines={os.environ.get('PMHC_TC3_ENGINES', '2')} B={B}: pair kernel {ms[0] / n[0] * 1e3:.1f} us per launch", flush=True)
else:
    for eng, B in ((1, 148), (2, 296), (1, 296), (2, 592), (1, 444), (2, 888)):
        subprocess.run([sys.executable, __file__, str(B)], env={**os.environ, "PMHC_TC3_ENGINES": str(eng)})
