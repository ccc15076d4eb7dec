"""Multi-GPU use of the hot path on one NVLink/NVSwitch box: one process per GPU, `torch.distributed` plumbing.

The reference is single-process; this module is the B200-side addition described in SURVEY.md §8e:

* sampling shards complexes over ranks — complexes are independent for the whole trajectory
  (optimizer.py:226-252), so there is NO collective on the data path;
* training is data-parallel by complex with ONE exchange step: an all-reduce of the flat fp32 gradient
  (79 195 floats = 317 KB; the kernels write weight gradients straight into that buffer).  The gnn2.* half is
  final first (the backward runs layer 2 before layer 1) and is reduced on a side stream while the layer-1
  backward kernel is still running; the gnn1.* half follows.  The 8 321 floats of gnn2.feature_mlp.* never carry a
  gradient (model.py:415) and are left out of the exchange; the mean over the GLOBAL batch is formed by scaling
  each rank's loss by 1 / B_global (the 1 / world factor is folded into the loss scale), so the collective is a plain SUM
  with nothing after it.  All ranks use the same noise step t per batch (optimizer.py:197 draws one t per batch) and draw
  the noise of complex g of the global batch from the same Philox counter range (key shared, `first` = the shard's first
  global complex), so DP over N shards equals one process on the concatenated batch — uneven shards included.
"""
from __future__ import annotations

import random
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from .. import _lib


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split of n_items over `world` ranks: sizes differ by at most one, order preserved."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict, rank: int, world: int) -> Tuple[Dict, int]:
    """This rank's slice of every per-complex tensor of `batch` (+ the global index of its first complex)."""
    n = batch["mask"].shape[0]
    lo, hi = shard_range(n, rank, world)
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n:
            out[k] = v[lo:hi]
        elif hasattr(v, "shape") and hasattr(v, "__getitem__") and not isinstance(v, (str, list, tuple)) and len(v.shape) > 0 and v.shape[0] == n:
            out[k] = v[lo:hi]  # Rigid
        else:
            out[k] = v
    return out, lo


def layer_split_offset() -> int:
    """First float of the gnn2.* parameters inside the flat buffer."""
    return int(_lib.load().pmhc_param_offset(24))


def shared_noise_step(T: int, step_index: int, seed: int) -> int:
    """Same t on every rank for training step `step_index` (stands in for the reference's unseeded random.randint)."""
    return random.Random(seed * 1_000_003 + step_index).randint(0, T - 1)


def unused_gradient_span() -> Tuple[int, int]:
    """[lo, hi) of gnn2.feature_mlp.* inside the flat buffer: the first four tensors of gnn2 (state-dict order), never used."""
    lib = _lib.load()
    lo = int(lib.pmhc_param_offset(24))
    hi = int(lib.pmhc_param_offset(27)) + int(lib.pmhc_param_numel(27))
    return lo, hi


def shared_noise_key(step_index: int, seed: int) -> int:
    """Same 62-bit Philox key on every rank for training step `step_index`."""
    return random.Random(seed * 7_000_003 + 2 * step_index + 1).getrandbits(62)


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    if dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over ranks of a flat gradient (SUM then scale: works on NCCL and gloo alike)."""
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat


def broadcast_parameters(model: torch.nn.Module, src: int = 0, group=None) -> None:
    """All ranks start from rank `src`'s weights (one broadcast of the flat buffer when the model has one)."""
    flat = model._flat_params() if hasattr(model, "_flat_params") else None
    if flat is not None:
        dist.broadcast(flat, src=src, group=group)
    else:
        for p in model.parameters():
            dist.broadcast(p.data, src=src, group=group)


class DataParallelTrainer:
    """Wraps a DiffusionModelOptimizer: per-rank shards of a global batch, shared t and noise key, overlapped two-bucket
    SUM all-reduce of the gradients that exist (loss pre-scaled by 1 / B_global)."""

    def __init__(self, dm, group=None, seed: int = 0, overlap: bool = True):
        self.dm, self.group, self.seed, self.step_index = dm, group, seed, 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.overlap = overlap and self.world > 1
        self._split = None
        self._event = None
        self._comm = None
        if self.world > 1:      # (a single process keeps the plain step: backward and Adam in one enqueue, graph-capturable as a whole)
            broadcast_parameters(dm.model, 0, group)
            dm.grad_hook = self._reduce
            dm.layer2_event_handle = self._event_handle

    def _event_handle(self):
        if not self.overlap:
            return None
        if self._event is None:
            dev = next(self.dm.model.parameters()).device
            self._event = torch.cuda.Event()
            self._event.record(torch.cuda.current_stream(dev))  # materialise the cudaEvent_t
            self._comm = torch.cuda.Stream(device=dev)
        return self._event.cuda_event

    def _buckets(self, flat_grad: torch.Tensor):
        """(gnn2 bucket, gnn1 bucket): views of the flat gradient that carry gradients, in the order they become final."""
        if self._split is None:
            self._split = layer_split_offset()
            self._unused = unused_gradient_span()
        lo, hi = self._unused
        assert lo == self._split, "gnn2.feature_mlp.* is expected at the start of the gnn2 half"
        return flat_grad[hi:], flat_grad[:self._split]

    def _reduce(self, flat_grad: torch.Tensor) -> None:
        if self.world <= 1:
            return
        second, first = self._buckets(flat_grad)
        if not self.overlap or self._event is None or not getattr(self.dm, "last_step_overlapped", True):
            # nothing to overlap with (e.g. the gradient step was one CUDA-graph replay): ONE latency-bound all-reduce over the whole flat
            # gradient — the span that never carries a gradient holds zeros and costs less than a second collective's launch
            allreduce_sum_(flat_grad, self.group)
            return
        main = torch.cuda.current_stream(flat_grad.device)
        comm = self._comm
        comm.wait_event(self._event)                  # gnn2.* half is final: reduce it while layer 1 still runs
        with torch.cuda.stream(comm):
            allreduce_sum_(second, self.group)
        done = torch.cuda.Event()
        done.record(main)                             # end of the layer-1 backward on the main stream
        comm.wait_event(done)
        with torch.cuda.stream(comm):
            allreduce_sum_(first, self.group)
        main.wait_stream(comm)

    def optimize(self, batch: Dict, metrics=None, noise: Optional[Dict] = None, global_batch: Optional[int] = None,
                 first_complex: Optional[int] = None) -> None:
        """One data-parallel step on this rank's shard `batch`.  global_batch = complexes over all ranks (default: world x
        this shard, i.e. equal shards); first_complex = global index of the shard's first complex (default: rank x shard)."""
        t = shared_noise_step(self.dm.noise_step_count, self.step_index, self.seed)
        key = shared_noise_key(self.step_index, self.seed)
        self.step_index += 1
        b_local = int(batch["mask"].shape[0])
        if b_local == 0:
            raise ValueError("this rank's shard is empty: drop or pad global batches smaller than the number of ranks "
                             "(every rank has to enter the gradient all-reduce)")
        if self.world <= 1:
            self.dm.optimize(batch, metrics, t=t, noise=noise)
            return
        if global_batch is None:
            global_batch = self.world * b_local
        if first_complex is None:
            first_complex = self.rank * b_local
        self.dm.optimize(batch, metrics, t=t, noise=noise, noise_key=key, noise_first_complex=first_complex,
                         loss_scale=1.0 / float(global_batch))


def all_gather_shards(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """Concatenation over ranks of contiguous shards (shard_range sizes, which may differ by one): all_gather
    needs equal shapes, so shards are padded to the largest and trimmed after."""
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in (shard_range(n_items, r, world) for r in range(world))]
    width = max(sizes)
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)])


def sample_sharded(dm, batch: Dict, rank: int, world: int, seed: int = 0, gather: bool = False, group=None) -> Dict:
    """Samples this rank's contiguous shard of `batch`; noise is keyed by GLOBAL complex index, so the sampled
    structures do not depend on the number of GPUs.  With gather=True every rank returns all complexes."""
    local, first = shard_batch(batch, rank, world)
    dm.sample_seed, dm.sample_first_complex = seed, first
    out = dm.sample(local)
    if not gather or world == 1:
        return out
    n = batch["mask"].shape[0]
    from ..rigid import Rigid
    res = dict(batch)
    res["frames"] = Rigid.from_tensor_7(all_gather_shards(out["frames"].to_tensor_7().contiguous(), n, group))
    res["torsions"] = all_gather_shards(out["torsions"].contiguous(), n, group)
    return res
