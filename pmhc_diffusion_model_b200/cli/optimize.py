#!/usr/bin/env python
"""Training entry point — same arguments and outputs as the reference's optimize.py (optimize.py:24-82):

    python -m pmhc_diffusion_model_b200.cli.optimize train_set.hdf5 100 model.pth [-T 1000] [-b 64] [--lr 0.001]

writes `<output_model>` (state dict, the reference's 48 keys) every 100 batches and per epoch, and one CSV row of mean
losses per epoch to `<output_model stem>.csv`.  Differences that do not change results: batches come GPU-resident from
`MhcpDataset.batches()` (`--num-workers` is accepted and ignored: there are no loader processes), the NaN flag stays on the
device (sticky; from the first NaN loss on every Adam update is skipped there) and is read before every save and at the
end of each epoch instead of once per step (`DiffusionModelOptimizer.check_nan`), `--precision` selects the arithmetic
(tc32 = tcgen05 forward with fp32-class results + fp32 backward; bf16 = tcgen05 forward + TF32 tensor-core backward).
Runs on one GPU; under torchrun each rank trains on its contiguous shard of every batch (same batch order, same t, noise drawn
per GLOBAL complex) with an NCCL gradient all-reduce; a tail batch smaller than the number of ranks is dropped on all ranks.
"""
import logging
import os
import sys
from argparse import ArgumentParser

import torch

_log = logging.getLogger(__name__)

arg_parser = ArgumentParser()
arg_parser.add_argument("train_hdf5", help="train data")
arg_parser.add_argument("epoch_count", type=int, help="number of epochs over the data")
arg_parser.add_argument("output_model", help="output model parameters file")
arg_parser.add_argument("--debug", "-d", action="store_const", const=True, default=False, help="run in debug mode")
arg_parser.add_argument("-T", type=int, help="number of noise steps", default=1000)
arg_parser.add_argument("--batch-size", "-b", type=int, help="data batch size", default=64)
arg_parser.add_argument("--num-workers", "-w", type=int, help="accepted for compatibility; batches are built on the GPU", default=4)
arg_parser.add_argument("--lr", type=float, help="learning rate", default=0.001)
arg_parser.add_argument("--precision", choices=["fp32", "tc32", "bf16"], default="tc32", help="arithmetic of the denoiser: fp32 FFMA (exact parity), tc32 = tcgen05 forward with fp16 hi/lo operand splits (fp32-class) + fp32 backward, or bf16 = tcgen05 bf16 forward + tcgen05 fp16 backward")
arg_parser.add_argument("--backward-precision", choices=["fp32", "fp16", "bf16"], default=None, help="arithmetic of the backward on its own (default: follows --precision): fp32 FFMA, fp16 = the tcgen05 backward (1e-2-class gradients; with --precision tc32 the fast combination that holds on ill-conditioned checkpoints), bf16 = the older TF32 mma.sync backward")
arg_parser.add_argument("--seed", type=int, default=None, help="seed of the batch order, noise steps and noise")
arg_parser.add_argument("--checkpoint", default=None, help="full training state (weights, Adam, random streams, epoch): written "
                        "after every epoch and resumed from when the file exists")


def main(argv=None) -> None:
    args = arg_parser.parse_args(argv)
    logging.basicConfig(stream=sys.stdout, level=logging.DEBUG if args.debug else logging.INFO)
    if not torch.cuda.is_available():
        raise SystemExit("pmhc_diffusion_model_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")

    import torch.distributed as dist
    from pmhc_diffusion_model_b200.diffusion.data import MhcpDataset
    from pmhc_diffusion_model_b200.diffusion.model import Model
    from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer
    from pmhc_diffusion_model_b200.diffusion.parallel import DataParallelTrainer, shard_batch
    from pmhc_diffusion_model_b200.diffusion.tools.metrics import MetricsRecord

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    if args.seed is not None:      # before the model is built: the initial weights come from torch's generator too
        import random
        random.seed(args.seed)
        torch.manual_seed(args.seed)
    model = Model(16, 22, args.T).to(device=device)
    if os.path.isfile(args.output_model):
        model.load_state_dict(torch.load(args.output_model, map_location=device), strict=True)
    model.precision = args.precision
    model.backward_precision = args.backward_precision
    dm = DiffusionModelOptimizer(args.T, model, args.lr)
    trainer = DataParallelTrainer(dm, seed=args.seed if args.seed is not None else 0)

    train_dataset = MhcpDataset(args.train_hdf5, device)
    order_seed = args.seed if args.seed is not None else torch.seed() % (1 << 31)
    if world > 1 and args.seed is None:      # unseeded: rank 0's draw decides, so every rank walks the same permutation
        box = torch.tensor([order_seed], dtype=torch.int64, device=device)
        dist.broadcast(box, src=0)
        order_seed = int(box.item())
        trainer.seed = order_seed            # ... and the same t / noise keys
    order = torch.Generator().manual_seed(order_seed)
    first_epoch = 0
    if args.checkpoint and os.path.isfile(args.checkpoint):
        ckpt = torch.load(args.checkpoint, map_location=device, weights_only=False)
        dm.load_state_dict(ckpt["trainer"])
        order.set_state(ckpt["order"].cpu())
        trainer.step_index = int(ckpt["step_index"])
        first_epoch = int(ckpt["epoch"]) + 1
        _log.info(f"resumed from {args.checkpoint} at epoch {first_epoch}")
    metrics_path = args.output_model.replace(".pth", ".csv")
    for epoch_index in range(first_epoch, args.epoch_count):
        _log.debug(f"starting epoch {epoch_index}")
        metrics = MetricsRecord()
        for i, batch in enumerate(train_dataset.batches(args.batch_size, device, shuffle=True, generator=order)):
            if world > 1:      # every rank draws the same order and takes its contiguous shard of the batch
                n_global = int(batch["mask"].shape[0])
                if n_global < world:
                    _log.debug(f"dropped a tail batch of {n_global} complexes (fewer than {world} ranks)")
                    continue
                batch, first = shard_batch(batch, rank, world)
                trainer.optimize(batch, metrics, global_batch=n_global, first_complex=first)
            else:
                trainer.optimize(batch, metrics)
            if i > 0 and i % 100 == 0:
                dm.check_nan()         # never overwrite a good model file with the weights of a run that has seen a NaN loss
                if rank == 0:
                    torch.save(model.state_dict(), args.output_model)
                    _log.debug(f"saved {args.output_model}")
        dm.check_nan()
        if rank == 0:
            torch.save(model.state_dict(), args.output_model)
            _log.debug(f"saved {args.output_model}")
            metrics.save(metrics_path, epoch_index)
            if args.checkpoint:
                torch.save({"trainer": dm.state_dict(), "order": order.get_state(), "step_index": trainer.step_index, "epoch": epoch_index},
                           args.checkpoint)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
