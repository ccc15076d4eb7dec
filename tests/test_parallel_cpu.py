"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, shared noise step, flat-gradient
all-reduce, parameter broadcast.  The kernels themselves are covered by the -m gpu tests."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pmhc_diffusion_model_b200.diffusion import parallel as par


def test_shard_range_is_a_balanced_ordered_partition():
    for n in (0, 1, 7, 1000, 100001):
        for world in (1, 2, 3, 8):
            ranges = [par.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.shard_range(10, 2, 2)


def test_shard_batch_slices_per_complex_tensors_only():
    batch = {"mask": torch.ones(10, 16, dtype=torch.bool), "frames": torch.arange(10 * 16 * 7.0).view(10, 16, 7),
             "name": ["x"] * 3, "scalar": torch.tensor(3.0)}
    parts = [par.shard_batch(batch, r, 3) for r in range(3)]
    assert [p[1] for p in parts] == [0, 4, 7]
    assert torch.equal(torch.cat([p[0]["frames"] for p in parts]), batch["frames"])
    assert parts[1][0]["name"] == batch["name"] and parts[1][0]["scalar"] is batch["scalar"]


def test_shared_noise_step_is_rank_independent():
    ts = [par.shared_noise_step(1000, k, seed=5) for k in range(50)]
    assert ts == [par.shared_noise_step(1000, k, seed=5) for k in range(50)]
    assert all(0 <= t < 1000 for t in ts) and len(set(ts)) > 10


def test_gradient_buckets_skip_the_parameters_that_never_get_a_gradient():
    """gnn2.feature_mlp.* (8 321 floats, model.py:415) sits at the start of the gnn2 half of the flat buffer and is left out of
    the all-reduce; the two buckets cover everything else exactly once."""
    lo, hi = par.unused_gradient_span()
    assert lo == par.layer_split_offset() and hi - lo == 128 * 64 + 64 + 64 + 1

    class Dm:
        model = None

        def __init__(self):
            self.noise_step_count = 10
    t = par.DataParallelTrainer.__new__(par.DataParallelTrainer)
    t._split = None
    flat = torch.arange(79195, dtype=torch.float32)
    second, first = t._buckets(flat)
    assert first.numel() + second.numel() + (hi - lo) == 79195
    assert first.data_ptr() == flat.data_ptr() and int(second[0]) == hi and int(first[-1]) == lo - 1
    keys = [par.shared_noise_key(k, 3) for k in range(20)]
    assert keys == [par.shared_noise_key(k, 3) for k in range(20)] and len(set(keys)) == 20 and all(0 <= k < 2 ** 62 for k in keys)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pmhc_diffusion_model_b200.diffusion.model import Model
        # flat gradient all-reduce == mean of the per-rank gradients
        g = torch.full((79195,), float(rank + 1))
        par.allreduce_mean_(g)
        assert torch.allclose(g, torch.full_like(g, (1 + world) / 2))
        # parameter broadcast through the flat buffer
        torch.manual_seed(100 + rank)
        model = Model(16, 22, 100)
        before = model._flat_params().clone()
        par.broadcast_parameters(model, 0)
        gathered = [torch.empty_like(before) for _ in range(world)]
        dist.all_gather(gathered, model._flat_params())
        assert all(torch.equal(gathered[0], x) for x in gathered)
        if rank != 0:
            assert not torch.equal(before, model._flat_params())
        assert model.gnn1.message_mlp[0].weight.data_ptr() == model._flat_params().data_ptr() + 4 * (64 * 87 + 64 + 64 * 64 + 64)
        # gather of contiguous shards restores the global order
        n = 11
        lo, hi = par.shard_range(n, rank, world)
        mine = torch.arange(lo, hi, dtype=torch.float32)
        assert torch.equal(par.all_gather_shards(mine, n), torch.arange(n, dtype=torch.float32))
        out.put((rank, "ok"))
    except Exception as e:  # surfaced in the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce_broadcast_and_gather():
    world = 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")  # spawned ranks import this module by name
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
