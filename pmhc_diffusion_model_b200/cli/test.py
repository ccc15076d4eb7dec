#!/usr/bin/env python
"""Sampling entry point — same arguments and outputs as the reference's test.py (test.py:20-84):

    python -m pmhc_diffusion_model_b200.cli.test model.pth test_set.hdf5 [-T 1000] [-b 64]

For every complex of the file: z_T = noise, T reverse steps (`DiffusionModelOptimizer.sample`, one C call per batch), the
full protein's atoms added, `<hdf5 stem>-sampled/<name>.pdb` written.  `--num-workers` is accepted and ignored (batches are
built on the GPU; loader batches are merged up to `--gpu-batch` complexes per sampling call); `--precision tc32` selects the tensor-core denoiser with fp32-class results (bf16: the faster 1e-2-class one).  Under torchrun the batches are dealt round-robin
over the ranks with no communication; `--seed` makes the noise of every complex independent of the number of GPUs.
"""
import logging
import os
import sys
from argparse import ArgumentParser

import torch

_log = logging.getLogger(__name__)

arg_parser = ArgumentParser()
arg_parser.add_argument("model", help="model parameters file")
arg_parser.add_argument("test_hdf5", help="test data")
arg_parser.add_argument("--debug", "-d", action="store_const", const=True, default=False, help="run in debug mode")
arg_parser.add_argument("-T", type=int, default=1000, help="number of noise steps")
arg_parser.add_argument("--batch-size", "-b", type=int, help="data batch size", default=64)
arg_parser.add_argument("--num-workers", "-w", type=int, help="accepted for compatibility; batches are built on the GPU", default=4)
arg_parser.add_argument("--precision", choices=["fp32", "tc32", "bf16"], default="tc32", help="arithmetic of the denoiser")
arg_parser.add_argument("--seed", type=int, default=None, help="noise seed (per-complex Philox streams)")
arg_parser.add_argument("--gpu-batch", type=int, default=1024, help="complexes sampled per kernel launch sequence: loader batches of "
                        "--batch-size are merged up to this size (a trajectory of 64 complexes leaves four fifths of a B200 idle); "
                        "with --seed the structures do not depend on it")


def main(argv=None) -> None:
    args = arg_parser.parse_args(argv)
    logging.basicConfig(stream=sys.stdout, level=logging.DEBUG if args.debug else logging.INFO)
    if not torch.cuda.is_available():
        raise SystemExit("pmhc_diffusion_model_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")

    from pmhc_diffusion_model_b200.diffusion.data import MhcpDataset
    from pmhc_diffusion_model_b200.diffusion.model import Model
    from pmhc_diffusion_model_b200.diffusion.optimizer import DiffusionModelOptimizer
    from pmhc_diffusion_model_b200.diffusion.tools.pdb import save_batch

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)

    model = Model(16, 22, args.T).to(device=device)
    model.load_state_dict(torch.load(args.model, map_location=device))
    model.precision = args.precision
    dm = DiffusionModelOptimizer(args.T, model, 0.0)
    dm.sample_seed = args.seed

    test_dataset = MhcpDataset(args.test_hdf5, device)
    output_path = os.path.splitext(args.test_hdf5)[0] + "-sampled"
    os.makedirs(output_path, exist_ok=True)

    with torch.no_grad():
        step = max(args.batch_size, args.gpu_batch)
        for i, true_batch in enumerate(test_dataset.batches(step, device)):
            if i % world != rank:
                continue
            names = list(true_batch["name"][0])
            dm.sample_first_complex = i * step
            # z_T (test.py:68-74): with --seed, keyed by the seed and the GLOBAL complex index — the same structures come out
            # whatever the batch size and the number of GPUs, run after run
            zkey = None if args.seed is None else (int(args.seed) * 0x9E3779B1 + 0x7A5) % (1 << 62)
            noise = dm.gen_noise(true_batch["frames"].shape[:-1], device=device, key=zkey, first_residue=i * step * 16)
            input_batch = {k: true_batch[k] for k in true_batch}
            input_batch["frames"] = noise["frames"].to_tensor_7()
            input_batch["torsions"] = noise["torsions"]
            pred_batch = dm.sample(input_batch)
            pred_batch.update(test_dataset.get_protein_positions(names))
            save_batch(pred_batch, names, output_path)
            _log.debug(f"sampled {len(names)} complexes")


if __name__ == "__main__":
    main()
