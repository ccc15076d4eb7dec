#!/usr/bin/env python
"""bench.py — sampled complexes/s (T = 100 reverse steps) of the denoising hot path, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one whole sampling trajectory (T = 100 x (denoiser forward + reverse step)) over one batch of
synthetic complexes: BASELINE.json configs[1] — 1 000 synthetic 9-mer complexes, M = 180 protein residues of
which 60 form the pocket, padded to the reference's pocket_maxlen = 80 — per GPU (weak scaling: complexes are
independent, each rank samples its own shard, no collective on the data path).

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM; `e2e`: through the public
DiffusionModelOptimizer.sample() call with the batch in pinned HOST memory (H2D + D2H inside the timed region);
`roofline`: the fused EGNN layer-forward kernel timed with CUDA events on its launching stream;
`cpu_baseline`: the UNMODIFIED reference (baseline/_ref, installed by baseline/install_reference.py; the oracle port when that
install is absent) on this box's host cores, bounded sample;
`train`: the second half of the metric (training complexes/s, B = 256) measured the same way, with the reference's own
optimize() timed beside it; `modes`: the other arithmetic modes on the same workload; `configs`: the other BASELINE configs.
`--impl reference` times the CPU path alone with all host threads (the reference has no GPU path).

Headline mode: precision "tc32" — tcgen05 tensor cores with fp16 hi + lo operand splits (fp32-class: it meets the fp32 parity
gate on the reference's shipped model.pth, tests/test_gpu_parity.py PARITY_MODES) — on the SHIPPED weights
(tests/golden/shipped_params.pt), the checkpoint whose attention logits reach 2.5e3.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T_STEPS = 100
N_COMPLEX = 1000
PEPTIDE_LEN = 9
POCKET_N = 60
P_PAD = 80
TRAIN_B = 256
T_TRAIN = 1000
CPU_SAMPLE_B = 16          # complexes in one CPU-baseline trajectory (bounded sample)
CPU_TRAIN_B = 64           # complexes in one CPU-baseline training step (BASELINE configs[0])
WORKLOAD = "sampling T=100, 1000 complexes/GPU, 9-mer peptide, M=180 protein / 60 pocket residues padded to 80 (BASELINE configs[1])"
METRIC = "sampled complexes/s (T=100 reverse steps)"
UNIT = "complexes/s"

# algorithmic work: 21 696 MAC per attention-carrying pair per layer
# (64x64 message layer 2 + 64x256 head hidden layers + 64x(2+4) extra inputs + 64x13 head outputs);
# message-only pairs of layer 1 (self, padded peptide slots, one shared padded-pocket message) cost 64x64.
MAC_PER_PAIR = 64 * 64 + 64 * 256 + 64 * 6 + 64 * 13
MAC_PER_MSG_PAIR = 64 * 64


# per forward mode: the fused layer kernel, what it computes in, its parity gate, and its measured DRAM traffic per launch
# (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture at B = 1000, summaries under profiles/)
KERNELS = {
    "tc32": {"kernel": "egnn_pair3_kernel<LAYER, 2>", "dtype": "f16x2-split (fp32-class)", "mma_terms": 3, "traffic": 31.1e6,
             "traffic_source": "profiles/prof_v3_r2_summary.csv",
             "gate": "max(1e-4, 2 x the reference's fp32 noise floor) on the reference fixtures, shipped weights included",
             "math": "tcgen05 f16 x f16 -> fp32 (TMEM), every operand as fp16 hi + lo, every contraction as hi.hi + lo.hi + hi.lo; geometry / softmax / updates fp32"},
    "fp16": {"kernel": "egnn_pair3_kernel<LAYER, 1>", "dtype": "f16", "mma_terms": 1, "traffic": None, "traffic_source": None,
             "gate": "1e-2 on well-conditioned weights", "math": "tcgen05 f16 x f16 -> fp32 (TMEM), single fp16 terms"},
    "bf16": {"kernel": "egnn_pair_tc_kernel<LAYER>", "dtype": "bf16", "mma_terms": 1, "traffic": 18.2e6, "traffic_source": "profiles/prof_pair_r1_summary.csv",
             "gate": "1e-2 on well-conditioned (random-init) weights; NOT met on the shipped model.pth",
             "math": "tcgen05 bf16 x bf16 -> fp32 (TMEM) for every per-pair contraction, geometry / softmax / updates fp32"},
    "fp32": {"kernel": "egnn_layer_forward_kernel<LAYER>", "dtype": "f32", "mma_terms": 1, "traffic": 15.1e6, "traffic_source": "profiles/prof_fwd_r1b_summary.csv",
             "gate": "max(1e-4, 2 x the reference's fp32 noise floor) on the reference fixtures", "math": "fp32 FFMA; tensor-pipe peak is the judged denominator"},
}


def forward_flops_per_complex(L=PEPTIDE_LEN, pocket_n=POCKET_N, n_pad=16):
    full = L * (L - 1 + pocket_n)
    msg_only = L * (n_pad - L + 1 + 1)
    return 2.0 * (2 * full * MAC_PER_PAIR + msg_only * MAC_PER_MSG_PAIR)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "hbm_gbs": p.get("hbm_gbs"), "source": "measured"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def synthetic(B, seed, P_pad=P_PAD, pocket_n=POCKET_N, L=PEPTIDE_LEN):
    from pmhc_diffusion_model_b200.synthetic import synthetic_batch   # seeded synthetic SwiftMHC-shaped complexes
    return synthetic_batch(B, L, pocket_n, P_pad=P_pad, seed=seed)


def load_weights():
    """The reference's shipped checkpoint (committed as a fixture: /root/reference does not exist on the GPU box)."""
    path = os.path.join(ROOT, "tests", "golden", "shipped_params.pt")
    if os.path.isfile(path):
        return torch.load(path, map_location="cpu"), "shipped model.pth of the reference (79 195 params)"
    from pmhc_diffusion_model_b200.synthetic import random_params
    return random_params(seed=0), "random init, reference architecture (79 195 params)"


class CpuReference:
    """The reference path on the host cores.  kind = "reference": the unmodified reference installed in baseline/_ref, imported
    through oracle/ref_shim.py (openfold -> the rigid_utils transformers ships); kind = "port": oracle/egnn_oracle.py.
    (The one place, with run_reference below, where bench.py executes oracle/: as the measured CPU baseline.)"""

    def __init__(self, params):
        self.params = params
        ref_root = os.path.join(ROOT, "baseline", "_ref")
        self.kind = "port"
        if os.path.isfile(os.path.join(ref_root, "diffusion", "optimizer.py")):
            try:
                from oracle import ref_shim
                ref_shim.REFERENCE_ROOT = ref_root
                self.ref_model, self.ref_opt, _ = ref_shim.load_reference()
                self.kind = "reference"
            except Exception as e:  # noqa: BLE001 — e.g. transformers' openfold_utils missing: fall back to the port, and say so
                print(f"baseline/_ref not importable ({e}); timing the oracle port", file=sys.stderr)
        from oracle import egnn_oracle as orc
        self.orc = orc

    def describe(self):
        return ("unmodified reference (baseline/_ref: diffusion.optimizer.DiffusionModelOptimizer on torch CPU fp32)" if self.kind == "reference"
                else "oracle/egnn_oracle.py on torch CPU fp32")

    def sample_rate(self, threads, B=CPU_SAMPLE_B, T=T_STEPS):
        """One trajectory of B complexes: DiffusionModelOptimizer.sample as test.py:60-75 drives it."""
        torch.set_num_threads(threads)
        raw = synthetic(B, seed=4242)
        if self.kind == "reference":
            model = self.ref_model.Model(16, 22, T)
            model.load_state_dict(self.params, strict=True)
            dm = self.ref_opt.DiffusionModelOptimizer(T, model, 0.0)
            torch.manual_seed(1)
            t0 = time.perf_counter()
            with torch.no_grad():
                noise = dm.gen_noise(raw["frames"].shape[:-1], device=torch.device("cpu"))
                batch = dict(raw)
                batch["frames"] = noise["frames"].to_tensor_7()
                batch["torsions"] = noise["torsions"]
                dm.sample(batch)
            dt = time.perf_counter() - t0
        else:
            orc = self.orc
            batch = orc.batch_to_frames(raw)
            g = torch.Generator().manual_seed(1)
            t0 = time.perf_counter()
            start = orc.gen_noise([B, 16], g)
            batch["frames"], batch["torsions"] = start["frames"], start["torsions"]
            orc.sample(self.params, batch, T, generator=g)
            dt = time.perf_counter() - t0
        return B / dt, dt

    def train_rate(self, threads, B=CPU_TRAIN_B, steps=3):
        """optimize() (optimizer.py:195-224) on B complexes, best of `steps`; reference kind only."""
        if self.kind != "reference":
            return None
        torch.set_num_threads(threads)
        raw = synthetic(B, seed=4343)
        model = self.ref_model.Model(16, 22, T_TRAIN)
        model.load_state_dict(self.params, strict=True)
        dm = self.ref_opt.DiffusionModelOptimizer(T_TRAIN, model, 1e-3)
        from diffusion.tools.metrics import MetricsRecord
        best = None
        for _ in range(steps):
            t0 = time.perf_counter()
            dm.optimize(dict(raw), MetricsRecord())
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return B / best, best


def io_leg(dev, dm, n=256):
    """HDF5 -> GPU-resident batches -> (sampled frames) -> PDB files on n synthetic complexes: rates of the rows around the
    hot path.  The reference's per-entry loader (file re-open + CPU eigh per entry, data.py:38, :107) and BioPython writer
    cannot run on this box (no h5py / BioPython), so only this repo's side is timed."""
    import shutil
    import tempfile
    from pmhc_diffusion_model_b200.diffusion.data import MhcpDataset, write_synthetic_hdf5
    from pmhc_diffusion_model_b200.diffusion.tools import pdb as pdbio
    tmp = tempfile.mkdtemp(prefix="pmhc_io_")
    try:
        path = os.path.join(tmp, "synthetic.hdf5")
        write_synthetic_hdf5(path, n, peptide_len=PEPTIDE_LEN, protein_len=180, pocket_n=POCKET_N, seed=3)
        t0 = time.perf_counter()
        ds = MhcpDataset(path, dev)
        ds.load_all()
        parse_s = time.perf_counter() - t0
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(10):
            batch = ds.device_batch(slice(0, n), dev)
        torch.cuda.synchronize(dev)
        batch_s = (time.perf_counter() - t0) / 10
        batch.update(ds.get_protein_positions(batch["name"][0]))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pdbio.peptide_atoms(batch)
        e0.record()
        for _ in range(20):
            pdbio.peptide_atoms(batch)
        e1.record()
        torch.cuda.synchronize(dev)
        atom_us = e0.elapsed_time(e1) * 1e3 / 20
        t0 = time.perf_counter()
        pdbio.save_batch(batch, batch["name"][0], os.path.join(tmp, "out"))
        write_s = time.perf_counter() - t0
        return {"complexes": n, "hdf5_parse_complexes_per_s": n / parse_s, "device_batch_complexes_per_s": n / batch_s,
                "atom14_kernel_us": atom_us, "pdb_files_per_s": n / write_s,
                "note": "HDF5 subset parsed in Python (no h5py in the image); device_batch = pinned slices -> H2D -> 4x4 to tensor_7 kernel; "
                        "PDB = one atom14 launch + text formatting of peptide (chain P) and 180-residue protein (chain M)"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (it has no GPU path): the unmodified
    reference from baseline/_ref when installed (kind "reference"), else its restatement in oracle/ (kind "port"); all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    params, weights = load_weights()
    cpu = CpuReference(params)
    for _ in range(min(args.warmup, 1)):
        cpu.sample_rate(threads, B=2, T=5)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        cpu.sample_rate(threads)
        done += CPU_SAMPLE_B
    dt = time.perf_counter() - t0
    value = done / dt
    train = cpu.train_rate(threads)
    sample = f"{CPU_SAMPLE_B} complexes x T={T_STEPS} per step (9-mer, pocket 60 padded to {P_PAD}), fp32, {cpu.describe()}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + ", bounded sample", "complexes_per_step": CPU_SAMPLE_B, "weights": weights},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "train": None if train is None else {"metric": "train complexes/s", "value": train[0], "unit": UNIT, "ms_per_step": train[1] * 1e3,
                                             "config": f"optimize() on B={CPU_TRAIN_B} (BASELINE configs[0] shape), best of 3, {threads} threads"},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--complexes", type=int, default=N_COMPLEX, help="complexes per GPU per step")
    ap.add_argument("--no-train", action="store_true", help="skip the training-throughput leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-io", action="store_true", help="skip the HDF5 loader / PDB writer leg")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every launch of a trajectory / training step instead of replaying a CUDA graph")
    ap.add_argument("--no-modes", action="store_true", help="skip the other arithmetic modes and the other BASELINE configs")
    ap.add_argument("--capture-allreduce", action="store_true", help="N > 1: capture the gradient all-reduce and Adam into the training step's CUDA graph (experimental)")
    ap.add_argument("--metric", default="sample", choices=["sample", "train"],
                    help="which half of BASELINE.json's metric the JSON line headlines: sampled complexes/s (default; the training record rides along under `train`) "
                         "or train complexes/s (the line is then the training step: value, e2e, roofline of the backward kernel, cpu_baseline = the reference's optimize())")
    ap.add_argument("--precision", default="tc32", choices=["fp32", "tc32", "fp16", "bf16"],
                    help="arithmetic of the denoiser's dense per-pair contractions in the headline legs (see include/pmhc_b200.h)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    # stdout carries exactly one line, the JSON: anything libraries print there (NCCL's version banner at communicator
    # creation, for one) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch.distributed as dist
    from pmhc_diffusion_model_b200 import _lib
    from pmhc_diffusion_model_b200.diffusion.model import Model
    from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.pmhc_check_device(), "pmhc_check_device")
    W, K, B = max(args.warmup, 3), args.steps, args.complexes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    params, weights_note = load_weights()
    model = Model(16, 22, T_STEPS)
    model.load_state_dict(params, strict=True)
    model = model.to(dev)
    model.precision = args.precision
    dm = DiffusionModelOptimizer(T_STEPS, model, 0.0)
    dm.sample_seed = 2024
    dm.sample_first_complex = rank * B        # Philox stream per global complex index: result independent of N
    dm.use_graph = not args.no_graph          # the trajectory's 4 T launches replayed as one CUDA graph (captured in the warm-up)

    keys = ("frames", "torsions", "features", "mask", "pocket_frames", "pocket_features", "pocket_mask")

    def make_inputs(batch_host, seed):
        host = {k: v.pin_memory() for k, v in batch_host.items() if isinstance(v, torch.Tensor)}
        torch.manual_seed(seed)
        n = host["frames"].shape[0]
        start = dm.gen_noise([n, 16], dev)                                               # z_T (test.py:71-74)
        host["frames"] = start["frames"].to_tensor_7().cpu().pin_memory()
        host["torsions"] = start["torsions"].cpu().pin_memory()
        return host, {k: host[k].to(dev) for k in keys}

    host, resident = make_inputs(synthetic(B, seed=1000 + rank), 7 + rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sample_resident(inputs=None):
        flush.zero_()                          # L2 flush between timed iterations
        return dm.sample(dict(resident if inputs is None else inputs))

    def sample_e2e():
        flush.zero_()
        batch = {k: host[k].to(dev, non_blocking=True) for k in keys}
        out = dm.sample(batch)
        return out["frames"].to_tensor_7().cpu(), out["torsions"].cpu()

    def timed(fn, warm, steps):
        """`steps` calls of fn between two events on the launching stream, barrier + synchronize on both sides; max over ranks."""
        for _ in range(warm):
            fn()
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---------------- value: inputs resident in HBM (no profiling hooks inside the timed region) ----------------
    for _ in range(W):
        sample_resident()
    barrier()
    launches0 = lib.pmhc_launch_count()
    with ClockSampler(local_rank) as clocks:
        ms = timed(sample_resident, 0, K)
    launches = lib.pmhc_launch_count() - launches0
    value = world * B * K / (ms / 1e3)

    # ---------------- e2e: host buffers through the public API ----------------
    for _ in range(2):
        sample_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        sample_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    h2d = sum(host[k].numel() * host[k].element_size() for k in keys)
    d2h = B * 16 * 21 * 4

    # ---------------- roofline of the dominant kernel (fused EGNN layer forward) ----------------
    # the same K steps once more with a CUDA event pair around every layer launch (pmhc_profile_*; kept out of `value`'s timed
    # region: creating 2 x 200 events per trajectory costs host time there)
    def kernel_times(fn, steps, owner):
        graphed, owner.use_graph = owner.use_graph, False      # (events cannot be recorded into a replayed graph: launches enqueued one by one)
        lib.pmhc_profile_enable(1)
        barrier()
        for _ in range(steps):
            fn()
        barrier()
        pm, pn = (ctypes.c_double * 2)(), (ctypes.c_int64 * 2)()
        lib.pmhc_profile_read(pm, pn)
        lib.pmhc_profile_enable(0)
        owner.use_graph = graphed
        return [pm[0], pm[1]], [pn[0], pn[1]]

    prof_ms, prof_n = kernel_times(sample_resident, K, dm)
    eager_ms = None
    if dm.use_graph and world == 1:
        dm.use_graph = False
        eager_ms = timed(sample_resident, 1, K) / K
        dm.use_graph = True
    peaks = measured_peaks()
    flops_per_launch = forward_flops_per_complex() * B / 2.0       # one layer per launch
    kernel_ms = prof_ms[0] / max(1, prof_n[0])
    achieved = flops_per_launch / (kernel_ms * 1e-3) / 1e12
    kinfo = KERNELS[args.precision]
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"],
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture under profiles/ (B = 1000)
                "traffic": kinfo["traffic"], "traffic_source": kinfo["traffic_source"], "peak_source": peaks["source"] + " bf16 sustained",
                "kernel": kinfo["kernel"], "kernel_ms": kernel_ms, "kernel_launches_timed": int(prof_n[0]),
                "kernel_share_of_step": prof_ms[0] / (ms * 1.0) if world == 1 else None,
                "math": kinfo["math"], "flops_per_launch": flops_per_launch,
                "tensor_flops_executed_per_launch": flops_per_launch * kinfo["mma_terms"],
                "note": "achieved counts ALGORITHMIC flops (43.4 kFLOP per pair per layer); this mode issues "
                        f"{kinfo['mma_terms']} MMA term(s) per contraction, so the tensor pipe does {kinfo['mma_terms']}x that"}

    # ---------------- the other arithmetic modes on the same workload (inputs resident), for the record ----------------
    modes = None
    if world == 1 and not args.no_modes:
        modes = {}
        for mode in ("fp32", "tc32", "bf16"):
            if mode == args.precision:
                continue
            model.precision = mode
            k2 = 1 if mode == "fp32" else 3
            oms = timed(sample_resident, 1, k2) / k2
            modes[mode] = {"value": B / (oms / 1e3), "unit": UNIT, "ms_per_step": oms, "parity_gate": KERNELS[mode]["gate"], "kernel": KERNELS[mode]["kernel"]}
        model.precision = args.precision

    # ---------------- training throughput (second half of the metric) ----------------
    train = None
    if not args.no_train:
        tmodel = Model(16, 22, T_TRAIN)
        tmodel.load_state_dict(params, strict=True)
        tmodel = tmodel.to(dev)
        tdm = DiffusionModelOptimizer(T_TRAIN, tmodel, 1e-3)
        tdm.use_graph = not args.no_graph      # N > 1: the gradient step is one graph, the two all-reduces and Adam follow it
        tdm.capture_grad_hook = bool(args.capture_allreduce)
        from pmhc_diffusion_model_b200.diffusion.parallel import DataParallelTrainer
        trainer = DataParallelTrainer(tdm, seed=0)   # N = 1: plain optimize(); N > 1: shared t + overlapped NCCL all-reduce
        n_train = 20

        def train_leg(tb, precision, backward):
            tmodel.precision, tmodel.backward_precision = precision, backward

            def step():
                flush.zero_()
                trainer.optimize(dict(tb), None)
            t = timed(step, W, n_train)
            tdm.check_nan()
            nb = tb["frames"].shape[0]
            return {"value": world * nb * n_train / (t / 1e3), "unit": UNIT, "ms_per_step": t / n_train, "global_batch": world * nb}, step

        tb = {k: v.to(dev) for k, v in synthetic(TRAIN_B, seed=5000 + rank).items()}
        fp32_leg, _ = train_leg(tb, "fp32", None)
        tc32_leg, _ = train_leg(tb, "tc32", None)          # tcgen05 hi/lo forward (saves the softmax statistics), fp32 FFMA backward
        bf16_leg, bf16_step = train_leg(tb, "bf16", None)  # bf16 tcgen05 forward + tcgen05 backward (fp16 operand tiles, TMEM accumulators)
        tprof_ms, tprof_n = kernel_times(bf16_step, n_train, tdm)
        # backward roofline (tensor pipe): 3 x 43.4 kFLOP per real pair per layer (recomputation + input gradients + weight
        # gradients), both layers' kernels averaged; pairs as in the forward count
        bwd_us = tprof_ms[1] / max(tprof_n[1], 1) * 1e3
        bwd_flops = TRAIN_B * 3.0 * 0.5 * forward_flops_per_complex()
        bf16_leg.update({"config": "bf16 training mode: tcgen05 bf16 forward + tcgen05 backward (fp16 operand tiles used K-major and MN-major, "
                                   "weight-gradient sums resident in tensor memory; gradient gate 1e-2 class)",
                         "backward_kernel": "egnn_layer_backward_t5_kernel<LAYER>",
                         "backward_kernel_us": bwd_us,
                         "backward_roofline": {"bound": "tensor", "achieved": bwd_flops / (bwd_us * 1e-6) / 1e12, "peak": peaks["bf16_tflops"],
                                               "unit": "TFLOP/s", "frac": bwd_flops / (bwd_us * 1e-6) / 1e12 / peaks["bf16_tflops"],
                                               "note": "algorithmic 3 x 43.4 kFLOP per pair per layer (the folded message layer executes fewer); "
                                                       "tcgen05 kind::f16, judged against the measured bf16 peak"}})
        mixed_leg, mixed_step = train_leg(tb, "tc32", "fp16")
        mixed_leg["config"] = "tcgen05 hi/lo-split forward (fp32-class outputs and loss) + tcgen05 fp16 backward (1e-2-class gradients)"
        bf16_leg["tc32_forward"] = mixed_leg
        train_e2e = None
        if args.metric == "train":
            # end to end through the public API: the batch comes from pinned host memory every step, the per-complex losses go back
            mprof_ms, mprof_n = kernel_times(mixed_step, n_train, tdm)
            mixed_leg["backward_kernel_us"] = mprof_ms[1] / max(mprof_n[1], 1) * 1e3
            tmodel.precision, tmodel.backward_precision = "tc32", "fp16"
            host_tb = {k: v.cpu().pin_memory() for k, v in tb.items()}

            def train_step_e2e():
                gbt = {k: v.to(dev, non_blocking=True) for k, v in host_tb.items()}
                trainer.optimize(gbt, None)
                return tdm.last_losses["total loss"].cpu()

            for _ in range(3):
                train_step_e2e()
            barrier()
            t0e = time.perf_counter()
            for _ in range(n_train):
                train_step_e2e()
            barrier()
            te = max_over_ranks((time.perf_counter() - t0e) * 1e3)
            train_e2e = {"value": world * TRAIN_B * n_train / (te / 1e3), "unit": UNIT, "ms_per_step": te / n_train,
                         "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host_tb.values())), "d2h_bytes_per_step": TRAIN_B * 4}
        if world == 1 and not args.no_modes:
            legacy_leg, legacy_step = train_leg(tb, "bf16", "bf16")   # the round-1 warp-level TF32 mma.sync backward, for comparison
            lprof_ms, lprof_n = kernel_times(legacy_step, n_train, tdm)
            legacy_leg.update({"config": "bf16 tcgen05 forward + warp-level TF32 mma.sync backward (round 1)",
                               "backward_kernel_us": lprof_ms[1] / max(lprof_n[1], 1) * 1e3})
            bf16_leg["tf32_mma_backward"] = legacy_leg
        tc32_leg["config"] = "fp32-class training: tcgen05 hi/lo-split forward + fp32 FFMA backward (gradient gate 1e-4, same as fp32)"
        train = {"metric": "train complexes/s", **fp32_leg, "steps": n_train,
                 "config": "B=256/GPU, 9-mer, pocket 60/80, fp32 FFMA forward + backward, noise+forward+loss+backward+Adam per step (BASELINE configs[2], [3] at N=8)",
                 "tc32": tc32_leg, "bf16": bf16_leg}
        if "tc32_forward" in bf16_leg:
            # fp32-class forward and loss + tcgen05 backward: the fast combination that also holds on the shipped checkpoint
            # (tests: ..._holds_on_the_shipped_checkpoint, ..._track_the_fp32_loss_curve)
            train["tc32_fp16_backward"] = bf16_leg["tc32_forward"]
        if world == 1 and not args.no_modes:
            # BASELINE configs[2] variant B: the whole M = 180 groove as the pocket (padded to 192)
            tb2 = {k: v.to(dev) for k, v in synthetic(TRAIN_B, seed=6000 + rank, P_pad=192, pocket_n=180).items()}
            train["full_groove_P192"] = {"fp32": train_leg(tb2, "fp32", None)[0], "tc32": train_leg(tb2, "tc32", None)[0],
                                         "bf16": train_leg(tb2, "bf16", None)[0]}
            del tb2
        tmodel.precision, tmodel.backward_precision = "fp32", None

    # ---------------- BASELINE configs[4]: mixed peptide lengths 8-15, pockets up to 400 slots (class I + II), bounded sample ----------------
    sweep = None
    if not args.no_modes:
        n5 = 2000
        from pmhc_diffusion_model_b200.synthetic import synthetic_batch
        h5, r5 = make_inputs(synthetic_batch(n5, (8, 15), (40, 400), P_pad=400, seed=9000 + rank), 99 + rank)
        dm.sample_first_complex = rank * n5
        t5 = timed(lambda: sample_resident(r5), 1, 2) / 2
        dm.sample_first_complex = rank * B
        sweep = {"value": world * n5 / (t5 / 1e3), "unit": UNIT, "ms_per_step": t5, "complexes_per_gpu": n5,
                 "config": "BASELINE configs[4] shapes: peptide length uniform 8-15, pocket uniform 40-400 of 400 slots, T=100, "
                           f"{args.precision}; {n5} complexes per GPU as a bounded sample of the 100k sweep (weak scaling, no collective)"}
        del h5, r5

    # ---------------- loader / writer rows, rank 0 at N = 1 ----------------
    io = None
    if rank == 0 and world == 1 and not args.no_io:
        io = io_leg(dev, dm)

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        ref = CpuReference(params)
        rate, secs = ref.sample_rate(threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": ref.kind,
               "sample": f"one trajectory of {CPU_SAMPLE_B} complexes x T={T_STEPS} (same shapes, same weights), {secs:.1f} s, {ref.describe()}"}
        tr = ref.train_rate(threads)
        if tr is not None and train is not None:
            train["cpu_baseline"] = {"value": tr[0], "unit": UNIT, "cores": threads, "kind": ref.kind,
                                     "sample": f"optimize() on B={CPU_TRAIN_B} (BASELINE configs[0] shape), best of 3, {tr[1]:.2f} s per step"}

    if rank == 0 and args.metric == "train" and train is not None:
        leg = train["tc32_fp16_backward"]
        bus = leg.get("backward_kernel_us")
        bflops = TRAIN_B * 3.0 * 0.5 * forward_flops_per_complex()
        line = json.dumps({
            "metric": "train complexes/s (one optimize() step: noise, forward, loss, backward, Adam)", "value": leg["value"], "unit": UNIT,
            "n_gpus": world, "steps": train["steps"], "warmup": W, "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16x2-split forward (fp32-class) + f16 tcgen05 backward, fp32 accumulation", "data": "synthetic",
            "config": {"workload": f"training step, {TRAIN_B} complexes/GPU, 9-mer peptide, M=180 protein / 60 pocket residues padded to 80 (BASELINE configs[2]; [3] at N=8)",
                       "complexes_per_gpu": TRAIN_B, "P_pad": P_PAD, "l2": "flushed between steps (256 MiB memset)",
                       "launch": "one CUDA graph per step" + ("" if world == 1 else ", then one NCCL all-reduce of the flat gradient and Adam"),
                       "weights": weights_note,
                       "precision": "tc32 forward (parity gate of the fp32 mode on the reference fixtures) + tcgen05 fp16 backward (gradient gates: 2e-2 of each tensor's "
                                    "largest entry vs the oracle's autograd, flat gradient cosine >= 0.99999; on the shipped checkpoint cosine 0.9999998 vs the fp32 backward)"},
            "clocks": clocks.summary(),
            "e2e": train_e2e,
            "gpu_launches": None,
            "roofline": None if bus is None else {"bound": "tensor", "achieved": bflops / (bus * 1e-6) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                                  "frac": bflops / (bus * 1e-6) / 1e12 / peaks["bf16_tflops"], "traffic": 12.2e6,
                                                  "traffic_source": "profiles/prof_t5_r2_summary.csv", "peak_source": peaks["source"] + " bf16 sustained",
                                                  "kernel": "egnn_layer_backward_t5_kernel<LAYER>", "kernel_ms": bus / 1e3,
                                                  "note": "algorithmic 3 x 43.4 kFLOP per pair per layer; two launches per step"},
            "cpu_baseline": train.get("cpu_baseline"),
            "train": train,
        })
        sys.stdout.flush()
        os.write(json_fd, (line + "\n").encode())
    elif rank == 0:
        line = json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": KERNELS[args.precision]["dtype"], "data": "synthetic",
            "config": {"workload": WORKLOAD, "complexes_per_gpu": B, "T": T_STEPS, "P_pad": P_PAD, "l2": "flushed between steps (256 MiB memset)",
                       "launch": ("one CUDA graph per trajectory (captured once per batch shape; seed / shard offset read from device memory)"
                                  if dm.use_graph else "every launch enqueued by pmhc_sample"),
                       "eager_ms_per_step": eager_ms,
                       "weights": weights_note, "precision": f"{args.precision}: {KERNELS[args.precision]['math']}; parity gate {KERNELS[args.precision]['gate']}"},
            "clocks": clocks.summary(),
            "e2e": {"value": world * B * K / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "modes": modes,
            "train": train,
            "sweep_config5": sweep,
            "io": io,
        })
        sys.stdout.flush()
        os.write(json_fd, (line + "\n").encode())
    if world > 1:
        if not args.no_train and args.capture_allreduce:
            tdm._step_states.clear()          # captured graphs hold the communicator: release them before it goes
            torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
