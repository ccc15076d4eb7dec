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
`cpu_baseline`: the CPU port of the reference path (oracle/) on this box's host cores, bounded sample;
`train`: the second half of the metric (training complexes/s, B = 256) measured the same way.
`--impl reference` times that CPU path alone with all host threads (the reference has no GPU path).
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
METRIC = "sampled complexes/s (T=100 reverse steps)"
UNIT = "complexes/s"

# algorithmic work: 21 696 MAC per attention-carrying pair per layer
# (64x64 message layer 2 + 64x256 head hidden layers + 64x(2+4) extra inputs + 64x13 head outputs);
# message-only pairs of layer 1 (self, padded peptide slots, one shared padded-pocket message) cost 64x64.
MAC_PER_PAIR = 64 * 64 + 64 * 256 + 64 * 6 + 64 * 13
MAC_PER_MSG_PAIR = 64 * 64


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
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
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


def cpu_trajectory_rate(params, threads, B=CPU_SAMPLE_B, T=T_STEPS, repeats=1):
    """Reference path on the host cores: oracle.sample() = the reference's DiffusionModelOptimizer.sample restated.
    (The one place, with run_reference below, where bench.py executes oracle/: as the measured CPU baseline.)"""
    from oracle import egnn_oracle as orc
    torch.set_num_threads(threads)
    batch = orc.batch_to_frames(synthetic(B, seed=4242))
    g = torch.Generator().manual_seed(1)
    start = orc.gen_noise([B, 16], g)
    batch["frames"], batch["torsions"] = start["frames"], start["torsions"]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.sample(params, batch, T, generator=g)
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
    """--impl reference: the reference's own CPU implementation of the path (it has no GPU path),
    restated in oracle/ (the reference itself cannot travel to the GPU box), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pmhc_diffusion_model_b200.synthetic import random_params
    threads = os.cpu_count() or 1
    params = random_params(seed=0)
    for _ in range(args.warmup):
        cpu_trajectory_rate(params, threads, B=2, T=5)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        cpu_trajectory_rate(params, threads)
        done += CPU_SAMPLE_B
    dt = time.perf_counter() - t0
    value = done / dt
    sample = f"{CPU_SAMPLE_B} complexes x T={T_STEPS} per step (9-mer, pocket 60 padded to {P_PAD}), fp32, torch CPU"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "sampling T=100, 9-mer peptide, M=180 protein / 60 pocket residues padded to 80 (BASELINE configs[1], bounded sample)",
                   "complexes_per_step": CPU_SAMPLE_B},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"],
                    help="arithmetic of the denoiser's two dense contractions in the sampling legs (see include/pmhc_b200.h)")
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
    from pmhc_diffusion_model_b200.synthetic import random_params
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

    params = random_params(seed=0)            # random-init weights of the reference architecture
    model = Model(16, 22, T_STEPS)
    model.load_state_dict(params, strict=True)
    model = model.to(dev)
    model.precision = args.precision
    dm = DiffusionModelOptimizer(T_STEPS, model, 0.0)
    dm.sample_seed = 2024
    dm.sample_first_complex = rank * B        # Philox stream per global complex index: result independent of N

    host = synthetic(B, seed=1000 + rank)
    host = {k: v.pin_memory() for k, v in host.items()}
    torch.manual_seed(7 + rank)
    start = dm.gen_noise([B, 16], dev)                                                   # z_T (test.py:71-74)
    host["frames"] = start["frames"].to_tensor_7().cpu().pin_memory()
    host["torsions"] = start["torsions"].cpu().pin_memory()
    keys = ("frames", "torsions", "features", "mask", "pocket_frames", "pocket_features", "pocket_mask")
    resident = {k: host[k].to(dev) for k in keys}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def sample_resident():
        flush.zero_()                          # L2 flush between timed iterations
        return dm.sample(dict(resident))

    def sample_e2e():
        flush.zero_()
        batch = {k: host[k].to(dev, non_blocking=True) for k in keys}
        out = dm.sample(batch)
        return out["frames"].to_tensor_7().cpu(), out["torsions"].cpu()

    # ---------------- value: inputs resident in HBM ----------------
    for _ in range(W):
        sample_resident()
    lib.pmhc_profile_enable(1)
    barrier()
    launches0 = lib.pmhc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(K):
            sample_resident()
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.pmhc_launch_count() - launches0
    prof_ms = (ctypes.c_double * 2)()
    prof_n = (ctypes.c_int64 * 2)()
    lib.pmhc_profile_read(prof_ms, prof_n)
    lib.pmhc_profile_enable(0)
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
    peaks = measured_peaks()
    flops_per_launch = forward_flops_per_complex() * B / 2.0       # one layer per launch
    kernel_ms = prof_ms[0] / max(1, prof_n[0])
    achieved = flops_per_launch / (kernel_ms * 1e-3) / 1e12
    tc = args.precision == "bf16"
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"],
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from profiles/ (ncu --set full, B = 1000)
                "traffic": 18.2e6 if tc else 15.1e6, "peak_source": peaks["source"] + " bf16 sustained",
                "kernel": "egnn_pair_tc_kernel" if tc else "egnn_layer_forward_kernel", "kernel_ms": kernel_ms,
                "kernel_share_of_step": prof_ms[0] / (ms * 1.0) if world == 1 else None,
                "math": ("tcgen05 bf16 x bf16 -> fp32 (TMEM) for every per-pair contraction, geometry / softmax / updates fp32" if tc
                         else "fp32 FFMA (exact-parity mode); tensor-pipe peak is the judged denominator"),
                "flops_per_launch": flops_per_launch}

    # the other precision mode, one short measurement, for the record (same workload, inputs resident)
    other = None
    if world == 1:
        model.precision = "fp32" if tc else "bf16"
        sample_resident()
        barrier()
        e0.record()
        sample_resident()
        e1.record()
        barrier()
        oms = e0.elapsed_time(e1)
        other = {"precision": model.precision, "value": B / (oms / 1e3), "unit": UNIT, "ms_per_step": oms,
                 "parity_gate": "1e-4 (fp32 FFMA, exact-parity mode)" if tc else "1e-2 (bf16 tensor cores)"}
        model.precision = args.precision

    # ---------------- training throughput (second half of the metric) ----------------
    train = None
    if not args.no_train:
        tmodel = Model(16, 22, T_TRAIN)
        tmodel.load_state_dict(params, strict=True)
        tmodel = tmodel.to(dev)
        tdm = DiffusionModelOptimizer(T_TRAIN, tmodel, 1e-3)
        tb = {k: v.to(dev) for k, v in synthetic(TRAIN_B, seed=5000 + rank).items()}
        from pmhc_diffusion_model_b200.diffusion.parallel import DataParallelTrainer
        trainer = DataParallelTrainer(tdm, seed=0)   # N = 1: plain optimize(); N > 1: shared t + overlapped NCCL all-reduce
        n_train = 20

        def time_training():
            for _ in range(W):
                trainer.optimize(dict(tb), None)
            barrier()
            e0.record()
            for _ in range(n_train):
                flush.zero_()
                trainer.optimize(dict(tb), None)
            e1.record()
            barrier()
            t = max_over_ranks(e0.elapsed_time(e1))
            tdm.check_nan()
            return t

        tms = time_training()
        tmodel.precision, tmodel.backward_precision = "bf16", "fp32"   # tensor-core forward (saves the softmax statistics), FFMA backward
        tms_bf16 = time_training()
        tmodel.backward_precision = None                                # ... and the TF32 tensor-core backward: the "bf16" training mode
        lib.pmhc_profile_enable(1)
        tms_tc = time_training()
        tprof_ms = (ctypes.c_double * 2)()
        tprof_n = (ctypes.c_int64 * 2)()
        lib.pmhc_profile_read(tprof_ms, tprof_n)
        lib.pmhc_profile_enable(0)
        tmodel.precision = "fp32"
        # backward roofline (tensor pipe): 3 x 43.4 kFLOP per real pair per layer (recomputation + input gradients + weight
        # gradients), both layers' kernels averaged; pairs as in the forward count
        bwd_us = tprof_ms[1] / max(tprof_n[1], 1) * 1e3
        bwd_flops = TRAIN_B * 3.0 * 0.5 * forward_flops_per_complex()
        train = {"metric": "train complexes/s", "value": world * TRAIN_B * n_train / (tms / 1e3), "unit": UNIT,
                 "ms_per_step": tms / n_train, "global_batch": world * TRAIN_B, "steps": n_train,
                 "config": "B=256/GPU, 9-mer, pocket 60/80, fp32, noise+forward+loss+backward+Adam per step",
                 "bf16_forward": {"value": world * TRAIN_B * n_train / (tms_bf16 / 1e3), "ms_per_step": tms_bf16 / n_train,
                                  "config": "same step with the tensor-core (bf16 operand) forward, fp32 FFMA backward"},
                 "bf16": {"value": world * TRAIN_B * n_train / (tms_tc / 1e3), "ms_per_step": tms_tc / n_train,
                          "config": "same step in the bf16 training mode: tcgen05 bf16 forward + TF32 tensor-core backward (gradient gate 1e-2 class)",
                          "backward_kernel_us": bwd_us,
                          "backward_roofline": {"bound": "tensor", "achieved": bwd_flops / (bwd_us * 1e-6) / 1e12,
                                                "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                                "frac": bwd_flops / (bwd_us * 1e-6) / 1e12 / peaks["bf16_tflops"],
                                                "note": "algorithmic 3 x 43.4 kFLOP per pair per layer; TF32 mma.sync (legacy tensor path), "
                                                        "judged against the measured bf16 peak"}}}

    # ---------------- loader / writer rows, rank 0 at N = 1 ----------------
    io = None
    if rank == 0 and world == 1 and not args.no_io:
        io = io_leg(dev, dm)

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, secs = cpu_trajectory_rate(params, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"one trajectory of {CPU_SAMPLE_B} complexes x T={T_STEPS} (same shapes), {secs:.1f} s, oracle/egnn_oracle.py on torch CPU fp32"}

    if rank == 0:
        line = json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if tc else "f32", "data": "synthetic",
            "config": {"workload": "sampling T=100, 1000 complexes/GPU, 9-mer peptide, M=180 protein / 60 pocket residues padded to 80 (BASELINE configs[1])",
                       "complexes_per_gpu": B, "T": T_STEPS, "P_pad": P_PAD, "l2": "flushed between steps (256 MiB memset)",
                       "weights": "random init, reference architecture (79 195 params)",
                       "precision": args.precision + (" (tcgen05: bf16 operands / fp32 accumulate for every per-pair contraction; parity gate 1e-2)"
                                                      if tc else " (FFMA; parity gate 1e-4)")},
            "clocks": clocks.summary(),
            "e2e": {"value": world * B * K / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "other_precision": other,
            "train": train,
            "io": io,
        })
        sys.stdout.flush()
        os.write(json_fd, (line + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
