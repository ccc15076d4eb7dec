"""ctypes binding of libpmhc_b200.so — the C ABI of include/pmhc_b200.h.

The library is loaded lazily and its absence is a hard error: the product has no other path.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# development: PMHC_B200_LIB points at an alternative build of the same library (A/B experiments on the GPU box)
LIB_PATH = os.environ.get("PMHC_B200_LIB") or os.path.join(_HERE, "libpmhc_b200.so")

N = 16          # PMHC_N
NFEAT = 22      # PMHC_NFEAT
NTORS = 7       # PMHC_NTORS
HID = 64        # PMHC_HID
NPARAM = 79195  # PMHC_NPARAM
ROWSTAT = 16    # PMHC_ROWSTAT
PRECISIONS = {"fp32": 0, "bf16": 1, "tc32": 2, "fp16": 3}  # PMHC_PRECISION_*
# default backward arithmetic of a forward mode: "fp16" = the tcgen05 backward (fp16 operand tiles, fp32 accumulation in tensor memory),
# "bf16" = the older warp-level TF32 mma.sync backward (kept selectable through Model.backward_precision for A/B runs)
BACKWARD_OF = {"fp32": "fp32", "bf16": "fp16", "tc32": "fp32", "fp16": "fp16"}

EXPORTS = (
    "pmhc_last_error", "pmhc_check_device", "pmhc_param_offset", "pmhc_param_numel", "pmhc_workspace_bytes",
    "pmhc_saved_floats", "pmhc_model_forward", "pmhc_model_forward_ex", "pmhc_model_backward", "pmhc_model_backward_ex", "pmhc_gen_noise", "pmhc_noise_from_randoms",
    "pmhc_add_noise", "pmhc_remove_noise", "pmhc_loss", "pmhc_sample", "pmhc_sample_ex", "pmhc_adam_step", "pmhc_adam_step_guarded", "pmhc_step_scalars", "pmhc_upload_small",
    "pmhc_train_step_grad", "pmhc_train_step_adam", "pmhc_launch_count", "pmhc_launch_count_add",
    "pmhc_profile_enable", "pmhc_profile_read", "pmhc_frames4x4_to_tensor7", "pmhc_atom14", "pmhc_format_pdb_host",
)


class PmhcBatch(Structure):
    _fields_ = [
        ("B", c_int32), ("P", c_int32),
        ("frames", c_void_p), ("torsions", c_void_p), ("features", c_void_p), ("mask", c_void_p),
        ("pocket_frames", c_void_p), ("pocket_features", c_void_p), ("pocket_mask", c_void_p),
    ]


class PmhcStepScalars(Structure):
    _fields_ = [
        ("t_over_T", c_float), ("beta", c_float), ("alpha", c_float), ("sigma", c_float),
        ("adam_step_size", c_float), ("adam_bc2_sqrt", c_float), ("grad_scale", c_float), ("reserved", c_float),
        ("noise_seed", c_uint64), ("noise_first_residue", c_uint64),
    ]


class PmhcStepBuffers(Structure):
    _fields_ = [(n, c_void_p) for n in ("noise_frames", "noise_torsions", "zt_frames", "zt_torsions", "pred_frames", "pred_torsions",
                                        "d_frames", "d_torsions", "losses", "saved", "flat_grad", "nan_flag")]


_lib = None


def load() -> ctypes.CDLL:
    """Returns the loaded library; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). pmhc_diffusion_model_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp, f32, i64, u64 = c_void_p, c_float, c_int64, c_uint64
    lib.pmhc_last_error.restype = c_char_p
    lib.pmhc_last_error.argtypes = []
    lib.pmhc_check_device.restype = c_int
    lib.pmhc_param_offset.restype = i64
    lib.pmhc_param_offset.argtypes = [c_int]
    lib.pmhc_param_numel.restype = i64
    lib.pmhc_param_numel.argtypes = [c_int]
    lib.pmhc_workspace_bytes.restype = c_size_t
    lib.pmhc_workspace_bytes.argtypes = [c_int, c_int]
    lib.pmhc_saved_floats.restype = c_size_t
    lib.pmhc_saved_floats.argtypes = [c_int, c_int]
    lib.pmhc_model_forward.restype = c_int
    lib.pmhc_model_forward.argtypes = [vp, POINTER(PmhcBatch), f32, vp, vp, vp, vp, c_size_t, vp]
    lib.pmhc_model_forward_ex.restype = c_int
    lib.pmhc_model_forward_ex.argtypes = [vp, POINTER(PmhcBatch), f32, vp, vp, vp, vp, c_size_t, vp, c_int]
    lib.pmhc_model_backward.restype = c_int
    lib.pmhc_model_backward.argtypes = [vp, POINTER(PmhcBatch), f32, vp, vp, vp, vp, vp, c_size_t, vp, vp]
    lib.pmhc_model_backward_ex.restype = c_int
    lib.pmhc_model_backward_ex.argtypes = [vp, POINTER(PmhcBatch), f32, vp, vp, vp, vp, vp, c_size_t, vp, vp, c_int]
    lib.pmhc_gen_noise.restype = c_int
    lib.pmhc_gen_noise.argtypes = [u64, u64, i64, vp, vp, vp]
    lib.pmhc_noise_from_randoms.restype = c_int
    lib.pmhc_noise_from_randoms.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.pmhc_add_noise.restype = c_int
    lib.pmhc_add_noise.argtypes = [vp, vp, vp, vp, c_double, i64, vp, vp, vp, vp]
    lib.pmhc_remove_noise.restype = c_int
    lib.pmhc_remove_noise.argtypes = [vp, vp, vp, vp, vp, vp, c_double, c_double, i64, vp, vp, vp, vp]
    lib.pmhc_loss.restype = c_int
    lib.pmhc_loss.argtypes = [vp, vp, vp, vp, vp, vp, c_int, f32, vp, vp, vp, vp]
    lib.pmhc_sample.restype = c_int
    lib.pmhc_sample.argtypes = [vp, POINTER(PmhcBatch), vp, vp, c_int, c_double, c_double, u64, u64, vp, vp, vp, vp, c_size_t, vp, c_int]
    lib.pmhc_adam_step.restype = c_int
    lib.pmhc_adam_step.argtypes = [vp, vp, vp, vp, i64, c_double, c_double, c_double, c_double, c_int, vp]
    lib.pmhc_adam_step_guarded.restype = c_int
    lib.pmhc_adam_step_guarded.argtypes = [vp, vp, vp, vp, i64, c_double, c_double, c_double, c_double, c_int, vp, vp]
    lib.pmhc_sample_ex.restype = c_int
    lib.pmhc_sample_ex.argtypes = [vp, POINTER(PmhcBatch), vp, vp, c_int, c_double, c_double, u64, u64, vp, vp, vp, vp, vp, c_size_t, vp, c_int]
    lib.pmhc_step_scalars.restype = c_int
    lib.pmhc_step_scalars.argtypes = [c_int, c_int, c_double, c_double, c_double, c_double, c_double, c_int, c_double, u64, u64,
                                      POINTER(PmhcStepScalars)]
    lib.pmhc_upload_small.restype = c_int
    lib.pmhc_upload_small.argtypes = [vp, vp, c_int, vp]
    lib.pmhc_train_step_grad.restype = c_int
    lib.pmhc_train_step_grad.argtypes = [vp, POINTER(PmhcBatch), vp, POINTER(PmhcStepScalars), vp, POINTER(PmhcStepBuffers), c_int, vp,
                                         vp, c_size_t, vp, vp, c_int, c_int]
    lib.pmhc_train_step_adam.restype = c_int
    lib.pmhc_train_step_adam.argtypes = [vp, vp, vp, vp, c_double, c_double, c_double, POINTER(PmhcStepScalars), vp, vp, vp]
    lib.pmhc_launch_count.restype = i64
    lib.pmhc_launch_count.argtypes = []
    lib.pmhc_launch_count_add.restype = None
    lib.pmhc_launch_count_add.argtypes = [i64]
    lib.pmhc_profile_enable.restype = None
    lib.pmhc_profile_enable.argtypes = [c_int]
    lib.pmhc_profile_read.restype = c_int
    lib.pmhc_profile_read.argtypes = [POINTER(c_double), POINTER(i64)]
    lib.pmhc_frames4x4_to_tensor7.restype = c_int
    lib.pmhc_frames4x4_to_tensor7.argtypes = [vp, i64, vp, vp]
    lib.pmhc_atom14.restype = c_int
    lib.pmhc_atom14.argtypes = [vp, vp, vp, vp, c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.pmhc_format_pdb_host.restype = i64
    lib.pmhc_format_pdb_host.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, c_char_p, c_char_p, c_char_p, vp, i64]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {load().pmhc_last_error().decode()}")


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """All tensors must live on one CUDA device: the kernels have no host path."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("pmhc_diffusion_model_b200 runs on CUDA (sm_100a) tensors only; got a CPU tensor. "
                               "There is no CPU fallback — move the model and batch to the GPU.")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def f32c(t: torch.Tensor) -> torch.Tensor:
    return t.to(dtype=torch.float32).contiguous()


def u8c(t: torch.Tensor) -> torch.Tensor:
    """bool/any mask -> contiguous uint8 (1 = set)."""
    if t.dtype == torch.bool:
        return t.contiguous().view(torch.uint8)
    return (t != 0).contiguous().view(torch.uint8)


_workspaces = {}


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Grow-only per-device scratch buffer (caller-owned memory as the C ABI requires)."""
    key = (device.type, device.index)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def make_batch(frames7, torsions, features, mask, pocket_frames7, pocket_features, pocket_mask):
    """Validates shapes and builds the PmhcBatch descriptor; returns (descriptor, keep-alive tensors)."""
    B, n, _ = frames7.shape
    P = pocket_frames7.shape[1]
    if n != N or features.shape[-1] != NFEAT:
        raise ValueError(f"kernels are built for Model({N}, {NFEAT}, T); got {n} peptide slots, {features.shape[-1]} features")
    if tuple(torsions.shape) != (B, N, NTORS, 2) or tuple(mask.shape) != (B, N):
        raise ValueError("torsions must be [B,16,7,2] and mask [B,16]")
    if tuple(pocket_features.shape) != (B, P, NFEAT) or tuple(pocket_mask.shape) != (B, P) or pocket_frames7.shape[-1] != 7:
        raise ValueError("pocket_frames must be [B,P,7], pocket_features [B,P,22], pocket_mask [B,P]")
    keep = [f32c(frames7), f32c(torsions), f32c(features), u8c(mask), f32c(pocket_frames7), f32c(pocket_features), u8c(pocket_mask)]
    require_cuda(*keep)
    desc = PmhcBatch(B, P, *[t.data_ptr() for t in keep])
    return desc, keep
