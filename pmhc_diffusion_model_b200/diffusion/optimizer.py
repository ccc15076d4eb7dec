"""`DiffusionModelOptimizer` — drop-in for diffusion/optimizer.py:27-252 of the reference.

Same constructor, attributes and method names; every method runs fused sm_100a kernels through the C ABI
(csrc/diffusion_ops.cu, csrc/egnn_*.cu).  Rotations never leave quaternion form: where the reference goes
through rotation matrices and `torch.linalg.eigh` (compose_r + rot_to_quat) the kernels
compose Hamilton products, which is the same rotation up to the eigenvector's arbitrary sign.
"""
from __future__ import annotations

import ctypes
import logging
import random
from math import sqrt
from typing import Dict, List, Optional, Tuple, Union

import torch

from .. import _lib
from ..rigid import Rigid, Rotation
from .tools.metrics import MetricsRecord

_log = logging.getLogger(__name__)

LOSS_KEYS = ("total loss", "positions loss", "rotations loss", "torsions loss", "rmsd")  # optimizer.py:73-79


def linear_schedule(t: int, T: int, beta_min: float, beta_max: float) -> float:
    """optimizer.py:20-21."""
    return beta_min + (beta_max - beta_min) * (float(t) / T)


def _frames7(x: Union[Rigid, torch.Tensor]) -> torch.Tensor:
    return x if isinstance(x, torch.Tensor) else x.to_tensor_7()


def _rigid(t7: torch.Tensor) -> Rigid:
    return Rigid(Rotation(quats=t7[..., :4], normalize_quats=False), t7[..., 4:])


def _next_noise_key() -> int:
    """64-bit Philox key for one noise draw, taken from torch's default CPU generator: `torch.manual_seed` makes
    runs repeatable exactly as it does for the reference's torch.randn / torch.rand draws (no GPU sync involved)."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class _LossFn(torch.autograd.Function):
    """get_loss with a fused gradient: one kernel computes the five per-complex losses [5, B] and d total / d pred."""

    @staticmethod
    def forward(ctx, true_f, true_t, pred_f, pred_t, mask_u8, tmask_u8):
        lib = _lib.load()
        dev = pred_f.device
        B = pred_f.shape[0]
        losses = torch.empty(5, B, device=dev, dtype=torch.float32)
        need = pred_f.requires_grad or pred_t.requires_grad
        d_f = torch.empty_like(pred_f) if need else None
        d_t = torch.empty_like(pred_t) if need else None
        with torch.cuda.device(dev):
            _lib.check(lib.pmhc_loss(true_f.data_ptr(), true_t.data_ptr(), pred_f.data_ptr(), pred_t.data_ptr(),
                                     mask_u8.data_ptr(), tmask_u8.data_ptr(), B, 1.0, losses.data_ptr(),
                                     _lib.ptr(d_f), _lib.ptr(d_t), _lib.stream_ptr(dev)), "pmhc_loss")
        ctx.need = need
        if need:
            ctx.save_for_backward(d_f, d_t)
        return losses

    @staticmethod
    def backward(ctx, g):
        if not ctx.need:
            return (None,) * 6
        d_f, d_t = ctx.saved_tensors
        g0 = g[0].to(torch.float32)  # only 'total loss' is differentiable; the other rows are diagnostics
        return None, None, d_f * g0[:, None, None], d_t * g0[:, None, None, None], None, None



class FlatAdam(torch.optim.Adam):
    """`torch.optim.Adam` (what optimizer.py:33 builds) whose step over the model's flat parameter buffer is one fused kernel
    (`pmhc_adam_step`, the same single-tensor update arithmetic) instead of torch's 21 multi-tensor launches.  It stays a
    `torch.optim.Adam`: `param_groups`, `state` (per parameter `step`, `exp_avg`, `exp_avg_sq` — the moments are views into two
    flat buffers), `state_dict()` / `load_state_dict()` and `zero_grad()` behave as before, and parameters whose `.grad` is None
    (gnn2.feature_mlp) get no state and no update, as in torch."""

    def __init__(self, model, lr: float):
        super().__init__(model.parameters(), lr=lr)
        self._model = model
        self._m = self._v = None
        self._generation = -1

    def _link(self, flat: torch.Tensor) -> None:
        """(Re)build the flat moment buffers for the model's current flat parameter buffer, keeping whatever state exists."""
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        off = 0
        for p in self._model.parameters():
            n = p.numel()
            st = self.state.get(p)
            if st:
                m[off:off + n].copy_(st["exp_avg"].reshape(-1))
                v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                st["exp_avg"], st["exp_avg_sq"] = m[off:off + n].view(p.shape), v[off:off + n].view(p.shape)
            off += n
        self._m, self._v, self._generation = m, v, self._model._flat_generation

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        self._generation = -1          # the loaded moments are fresh tensors: copied into the flat buffers at the next step

    @torch.no_grad()
    def step(self, closure=None, flat_grad: Optional[torch.Tensor] = None, skip_flag: Optional[torch.Tensor] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        group = self.param_groups[0]
        if (len(self.param_groups) != 1 or group["weight_decay"] != 0 or group["amsgrad"] or group["maximize"]
                or group.get("capturable") or group.get("differentiable")):
            raise RuntimeError("FlatAdam implements plain Adam (one group, no weight decay / amsgrad / maximize), as optimizer.py:33 uses it")
        lib = _lib.load()
        model = self._model
        flat = model._flat_params()
        if self._generation != model._flat_generation or self._m is None:
            self._link(flat)
        params = list(model.parameters())
        if flat_grad is None:        # plain .step(): gather the .grad tensors (None -> that parameter is skipped)
            flat_grad = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in params])
        flat_grad = _lib.f32c(flat_grad)
        # contiguous runs of parameters that carry a gradient; all of them share one step count, as in torch when they are
        # always stepped together (a parameter that gets its first gradient later starts its own count: one run per count)
        spans, steps, off = [], [], 0
        for p in params:
            n = p.numel()
            if p.grad is not None:
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = self._m[off:off + n].view(p.shape)
                    st["exp_avg_sq"] = self._v[off:off + n].view(p.shape)
                spans.append((off, off + n))
                steps.append(st["step"])
            off += n
        runs = []
        if steps:
            torch._foreach_add_(steps, 1.0)
            for (lo, hi), k in zip(spans, torch.stack(steps).tolist()):
                k = int(k)
                if runs and runs[-1][1] == lo and runs[-1][2] == k:
                    runs[-1][1] = hi
                else:
                    runs.append([lo, hi, k])
        b1, b2 = group["betas"]
        dev = flat.device
        with torch.cuda.device(dev):
            for lo, hi, k in runs:
                # skip_flag (one device byte): non-zero = a non-finite loss was seen, leave weights and moments untouched
                _lib.check(lib.pmhc_adam_step_guarded(flat.data_ptr() + 4 * lo, flat_grad.data_ptr() + 4 * lo, self._m.data_ptr() + 4 * lo,
                                                      self._v.data_ptr() + 4 * lo, hi - lo, float(group["lr"]), float(b1), float(b2),
                                                      float(group["eps"]), k, _lib.ptr(skip_flag), _lib.stream_ptr(dev)), "pmhc_adam_step")
        return loss


    # ---- the training step's own path: one pmhc_train_step_adam call over both gradient-carrying runs ------------------------
    def next_step(self) -> int:
        """Adam step count of the update about to be made (all gradient-carrying parameters are stepped together)."""
        for name, p in self._model.named_parameters():
            if not name.startswith("gnn2.feature_mlp"):
                st = self.state.get(p)
                return int(st["step"]) + 1 if st else 1
        return 1

    def prepare_flat(self) -> None:
        """Moment buffers linked to the model's current flat buffer and state entries in place (before a step is enqueued)."""
        model = self._model
        flat = model._flat_params()
        if self._generation != model._flat_generation or self._m is None:
            self._link(flat)
        off = 0
        for name, p in model.named_parameters():
            n = p.numel()
            if not name.startswith("gnn2.feature_mlp") and len(self.state[p]) == 0:
                st = self.state[p]
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = self._m[off:off + n].view(p.shape)
                st["exp_avg_sq"] = self._v[off:off + n].view(p.shape)
            off += n

    def count_step(self) -> None:
        steps = [self.state[p]["step"] for name, p in self._model.named_parameters() if not name.startswith("gnn2.feature_mlp")]
        torch._foreach_add_(steps, 1.0)

    def hyper(self):
        group = self.param_groups[0]
        if (len(self.param_groups) != 1 or group["weight_decay"] != 0 or group["amsgrad"] or group["maximize"]
                or group.get("capturable") or group.get("differentiable")):
            raise RuntimeError("FlatAdam implements plain Adam (one group, no weight decay / amsgrad / maximize), as optimizer.py:33 uses it")
        b1, b2 = group["betas"]
        return float(group["lr"]), float(b1), float(b2), float(group["eps"])


class _StepState:
    """Device buffers of one training step for one batch shape.  Their addresses never change, so a CUDA graph captured around the
    step's two C calls can be replayed for every later batch of that shape (inputs are copied into `inputs` first)."""

    def __init__(self, B: int, P: int, dev: torch.device, nan_flag: torch.Tensor):
        lib = _lib.load()

        def f(*shape):
            return torch.empty(*shape, device=dev, dtype=torch.float32)
        self.B, self.P, self.dev = B, P, dev
        self.noise_f, self.noise_t = f(B, _lib.N, 7), f(B, _lib.N, _lib.NTORS, 2)
        self.zt_f, self.zt_t = f(B, _lib.N, 7), f(B, _lib.N, _lib.NTORS, 2)
        self.pred_f, self.pred_t = f(B, _lib.N, 7), f(B, _lib.N, _lib.NTORS, 2)
        self.d_f, self.d_t = f(B, _lib.N, 7), f(B, _lib.N, _lib.NTORS, 2)
        self.losses = f(5, B)
        self.saved = f(lib.pmhc_saved_floats(B, P))
        self.grad = f(_lib.NPARAM)
        self.buffers = _lib.PmhcStepBuffers(*[t.data_ptr() for t in (self.noise_f, self.noise_t, self.zt_f, self.zt_t, self.pred_f, self.pred_t,
                                                                      self.d_f, self.d_t, self.losses, self.saved, self.grad, nan_flag)])
        self.ws_bytes = lib.pmhc_workspace_bytes(B, P)
        # graph mode only
        self.inputs = None          # static copies of the eight batch tensors
        self.scalars_dev = None     # PmhcStepScalars in device memory
        self.graph = None
        self.graph_key = None

    def static_inputs(self, tensors):
        if self.inputs is None:
            self.inputs = [torch.empty_like(t) for t in tensors]
            self.scalars_dev = torch.zeros(48, dtype=torch.uint8, device=self.dev)
        for dst, src in zip(self.inputs, tensors):
            dst.copy_(src, non_blocking=True)
        return self.inputs


class DiffusionModelOptimizer:

    def __init__(self, noise_step_count: int, model: torch.nn.Module, lr: float):
        self.noise_step_count = noise_step_count
        self.model = model
        self.optimizer = FlatAdam(self.model, lr)  # optimizer.py:33: torch.optim.Adam, stepped by one fused kernel
        self.beta_min = 0.0
        self.beta_max = 0.8
        # parity hooks (tests only): reference z_t quaternions to align the q / -q sign with
        self.quat_sign_ref: Optional[torch.Tensor] = None
        # sharded sampling: a fixed (seed, index of this shard's first complex) makes the Philox noise of every
        # complex independent of how the complexes are split over GPUs
        self.sample_seed: Optional[int] = None
        self.sample_first_complex: int = 0
        # CUDA graphs: capture the training step (optimize) / the whole trajectory (sample) once per batch shape and replay it;
        # per-step scalars (t, noise key, Adam bias corrections; the sampling seed) are read from device memory by the kernels
        self.use_graph: bool = False
        # use_graph with a data-parallel wrapper: True captures its all-reduce and Adam into the step's graph (0.90 vs 0.96 ms per step at
        # N = 2) — off by default: with torch 2.11 / NCCL 2.28 the process then hangs in destroy_process_group() at exit
        self.capture_grad_hook: bool = False
        self._step_states: Dict = {}
        self._sample_graphs: Dict = {}
        self._nan_flag: Optional[torch.Tensor] = None

    # ---- loss ---------------------------------------------------------------------------------------------
    @staticmethod
    def get_loss(
        noise_true: Dict[str, Union[Rigid, torch.Tensor]],
        noise_pred: Dict[str, Union[Rigid, torch.Tensor]],
        residues_mask: torch.Tensor,
        torsions_mask: torch.Tensor,
    ) -> Dict[str, torch.Tensor]:
        """optimizer.py:38-79: 0.1 * masked translation MSE + masked (1 - q.q') + masked (1 - t.t')."""
        true_f = _lib.f32c(_frames7(noise_true["frames"]))
        pred_f = _lib.f32c(_frames7(noise_pred["frames"]))
        true_t = _lib.f32c(noise_true["torsions"])
        pred_t = _lib.f32c(noise_pred["torsions"])
        _lib.require_cuda(true_f, pred_f, true_t, pred_t, residues_mask, torsions_mask)
        losses = _LossFn.apply(true_f.detach(), true_t.detach(), pred_f, pred_t, _lib.u8c(residues_mask), _lib.u8c(torsions_mask))
        out = {LOSS_KEYS[0]: losses[0]}
        out.update({k: losses[i].detach() for i, k in enumerate(LOSS_KEYS) if i > 0})
        return out

    # ---- schedule -----------------------------------------------------------------------------------------
    def get_beta_alpha_sigma(self, noise_step: int) -> Tuple[float, float, float]:
        """optimizer.py:81-91."""
        beta = linear_schedule(noise_step, self.noise_step_count, self.beta_min, self.beta_max)
        return (beta, sqrt(1.0 - beta), sqrt(beta))

    # ---- noise --------------------------------------------------------------------------------------------
    @staticmethod
    def gen_noise(shape: Union[List[int], Tuple[int]], device: torch.device, key: Optional[int] = None,
                  first_residue: int = 0) -> Dict[str, Union[Rigid, torch.Tensor]]:
        """optimizer.py:93-108: translation 5*N(0,I), uniform rotation (Shoemake), 7 uniform torsion angles.
        Philox counter stream instead of torch's generator (statistically identical, not bitwise).  `key` (default: drawn from
        torch's CPU generator) and `first_residue` select the counter range: residue r of the call uses counter first_residue + r,
        so shards of one global batch (same key, first_residue = 16 x the shard's first complex) draw what one call would."""
        lib = _lib.load()
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("gen_noise runs on the GPU only (no CPU fallback)")
        shape = list(shape)
        n = 1
        for s in shape:
            n *= int(s)
        frames = torch.empty(shape + [7], device=device, dtype=torch.float32)
        tors = torch.empty(shape + [_lib.NTORS, 2], device=device, dtype=torch.float32)
        with torch.cuda.device(device):
            _lib.check(lib.pmhc_gen_noise(_next_noise_key() if key is None else int(key), int(first_residue), n, frames.data_ptr(),
                                          tors.data_ptr(), _lib.stream_ptr(device)), "pmhc_gen_noise")
        return {"frames": _rigid(frames), "torsions": tors}

    @staticmethod
    def noise_from_randoms(normal: torch.Tensor, uniform: torch.Tensor) -> Dict[str, Union[Rigid, torch.Tensor]]:
        """Same formulas as gen_noise from caller-supplied randoms: normal [*,3] ~ N(0,1), uniform [*,10] ~ U(0,1)."""
        lib = _lib.load()
        normal, uniform = _lib.f32c(normal), _lib.f32c(uniform)
        dev = _lib.require_cuda(normal, uniform)
        shape = list(normal.shape[:-1])
        n = normal.numel() // 3
        frames = torch.empty(shape + [7], device=dev, dtype=torch.float32)
        tors = torch.empty(shape + [_lib.NTORS, 2], device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(lib.pmhc_noise_from_randoms(normal.data_ptr(), uniform.data_ptr(), n, frames.data_ptr(), tors.data_ptr(),
                                                   _lib.stream_ptr(dev)), "pmhc_noise_from_randoms")
        return {"frames": _rigid(frames), "torsions": tors}

    def add_noise(self, signal: Dict, noise: Dict, t: int) -> Dict:
        """optimizer.py:110-138.  All other keys of `signal` pass through untouched (:135)."""
        lib = _lib.load()
        beta, _, _ = self.get_beta_alpha_sigma(t)
        f0, t0 = _lib.f32c(_frames7(signal["frames"])), _lib.f32c(signal["torsions"])
        fe, te = _lib.f32c(_frames7(noise["frames"])), _lib.f32c(noise["torsions"])
        dev = _lib.require_cuda(f0, t0, fe, te)
        out_f, out_t = torch.empty_like(f0), torch.empty_like(t0)
        ref = _lib.f32c(self.quat_sign_ref) if self.quat_sign_ref is not None else None
        with torch.cuda.device(dev):
            _lib.check(lib.pmhc_add_noise(f0.data_ptr(), t0.data_ptr(), fe.data_ptr(), te.data_ptr(), beta, f0.numel() // 7,
                                          _lib.ptr(ref), out_f.data_ptr(), out_t.data_ptr(), _lib.stream_ptr(dev)), "pmhc_add_noise")
        result = {k: signal[k] for k in signal}
        result["frames"] = _rigid(out_f)
        result["torsions"] = out_t
        return result

    def remove_noise(self, noised_signal: Dict, predicted_noise: Dict, t: int, s: int,
                     random_noise: Optional[Dict] = None) -> Dict:
        """optimizer.py:140-193; draws the step's fresh noise itself (:151) unless `random_noise` is given."""
        lib = _lib.load()
        beta_t, _, _ = self.get_beta_alpha_sigma(t)
        beta_s, _, _ = self.get_beta_alpha_sigma(s)
        zf, zt = _lib.f32c(_frames7(noised_signal["frames"])), _lib.f32c(noised_signal["torsions"])
        pf, pt = _lib.f32c(_frames7(predicted_noise["frames"])), _lib.f32c(predicted_noise["torsions"])
        dev = _lib.require_cuda(zf, zt, pf, pt)
        if random_noise is None:
            random_noise = DiffusionModelOptimizer.gen_noise(zf.shape[:-1], dev)
        xf, xt = _lib.f32c(_frames7(random_noise["frames"])), _lib.f32c(random_noise["torsions"])
        out_f, out_t = torch.empty_like(zf), torch.empty_like(zt)
        ref = _lib.f32c(self.quat_sign_ref) if self.quat_sign_ref is not None else None
        with torch.cuda.device(dev):
            _lib.check(lib.pmhc_remove_noise(zf.data_ptr(), zt.data_ptr(), pf.data_ptr(), pt.data_ptr(), xf.data_ptr(), xt.data_ptr(),
                                             beta_t, beta_s, zf.numel() // 7, _lib.ptr(ref), out_f.data_ptr(), out_t.data_ptr(),
                                             _lib.stream_ptr(dev)), "pmhc_remove_noise")
        result = {k: noised_signal[k] for k in noised_signal}
        result["frames"] = _rigid(out_f)
        result["torsions"] = out_t
        return result

    # ---- training step --------------------------------------------------------------------------------------
    def optimize(self, batch: Dict[str, Union[Rigid, torch.Tensor]], metrics: Optional[MetricsRecord] = None,
                 t: Optional[int] = None, noise: Optional[Dict] = None, noise_key: Optional[int] = None,
                 noise_first_complex: int = 0, loss_scale: Optional[float] = None):
        """One training step (optimizer.py:195-224): draw t and noise, noise the batch, predict, loss, backward,
        Adam.  `t` / `noise` may be pinned by the caller (parity tests, data-parallel ranks sharing one t); data-parallel ranks
        also pass the shared Philox key with their shard's first global complex, and loss_scale = 1 / B_global (default 1 / B:
        total_loss.mean(), optimizer.py:222).  Six fused launches + Adam; no autograd graph is built."""
        lib = _lib.load()
        if t is None:
            t = random.randint(0, self.noise_step_count - 1)  # optimizer.py:197
        self.optimizer.zero_grad()

        # in-place conversion of the caller's batch, as optimizer.py:201-202
        batch["frames"] = Rigid.from_tensor_7(_frames7(batch["frames"]))
        batch["pocket_frames"] = Rigid.from_tensor_7(_frames7(batch["pocket_frames"]))

        model = self.model
        desc, keep = _lib.make_batch(batch["frames"].to_tensor_7(), batch["torsions"], batch["features"], batch["mask"],
                                     batch["pocket_frames"].to_tensor_7(), batch["pocket_features"], batch["pocket_mask"])
        tmask = _lib.u8c(batch["torsions_mask"])
        B, P = desc.B, desc.P
        dev = keep[0].device
        if self._nan_flag is None or self._nan_flag.device != dev:
            self._nan_flag = torch.zeros(1, dtype=torch.uint8, device=dev)
            self._step_states.clear()
        st = self._step_states.get((B, P))
        if st is None:
            if len(self._step_states) >= 4:          # e.g. the tail batch of every epoch: keep the most recent shapes only
                self._step_states.pop(next(iter(self._step_states)))
            st = self._step_states[(B, P)] = _StepState(B, P, dev, self._nan_flag)
        flat = model._flat_params()
        opt = self.optimizer
        opt.prepare_flat()
        lr, b1, b2, eps = opt.hyper()
        key = _next_noise_key() if (noise is None and noise_key is None) else int(noise_key or 0)
        sc = _lib.PmhcStepScalars()
        _lib.check(lib.pmhc_step_scalars(int(t), int(model.T), self.beta_min, self.beta_max, lr, b1, b2, opt.next_step(),
                                         1.0 / B if loss_scale is None else float(loss_scale), key, int(noise_first_complex) * _lib.N,
                                         ctypes.byref(sc)), "pmhc_step_scalars")
        if noise is not None:
            st.noise_f.copy_(_lib.f32c(_frames7(noise["frames"])))
            st.noise_t.copy_(_lib.f32c(noise["torsions"]))
        sign = _lib.f32c(self.quat_sign_ref) if self.quat_sign_ref is not None else None
        ws = _lib.workspace(dev, st.ws_bytes)
        stream_of = lambda: _lib.stream_ptr(dev)
        hooked = type(self).grad_hook is not getattr(self.grad_hook, "__func__", None)   # a data-parallel wrapper reduces between the calls
        # (a graph-captured step cannot record an event another, non-captured stream waits on: no overlap in graph mode)
        event = None if self.use_graph else self.layer2_event_handle()
        self.last_step_overlapped = event is not None

        def enqueue(d, tm, sc_dev, with_adam):
            _lib.check(lib.pmhc_train_step_grad(flat.data_ptr(), ctypes.byref(d), tm.data_ptr(), ctypes.byref(sc), sc_dev, ctypes.byref(st.buffers),
                                                0 if noise is not None else 1, _lib.ptr(sign), ws.data_ptr(), st.ws_bytes, stream_of(), event,
                                                model.precision_code(), model.backward_precision_code()), "pmhc_train_step_grad")
            if with_adam:
                _lib.check(lib.pmhc_train_step_adam(flat.data_ptr(), st.grad.data_ptr(), opt._m.data_ptr(), opt._v.data_ptr(), b1, b2, eps,
                                                    ctypes.byref(sc), sc_dev, self._nan_flag.data_ptr(), stream_of()), "pmhc_train_step_adam")

        with torch.cuda.device(dev):
            if not self.use_graph:
                enqueue(desc, tmask, None, not hooked)
            else:
                # static copies of the inputs, the scalar block refreshed through a kernel's launch parameters, then one graph launch
                ins = st.static_inputs(keep + [tmask])
                gdesc = _lib.PmhcBatch(B, P, *[x.data_ptr() for x in ins[:7]])
                # opt-in (capture_grad_hook): a data-parallel wrapper's collective (NCCL is capturable) and the Adam step behind it join
                # the graph, so the whole step is ONE launch at N > 1 too; by default they follow the replay eagerly
                in_graph_hook = hooked and self.capture_grad_hook
                gkey = (model.precision_code(), model.backward_precision_code(), noise is not None, sign is not None and sign.data_ptr(),
                        hooked, in_graph_hook, event, flat.data_ptr(), opt._m.data_ptr(), ws.data_ptr(), lr, b1, b2, eps)
                if st.graph is None or st.graph_key != gkey:
                    if st.graph is None:
                        enqueue(gdesc, ins[7], None, False)          # warm-up outside capture: lazy kernel-attribute set-up happens here
                        if in_graph_hook:
                            self.grad_hook(st.grad)                   # ... and the communicator's first collective
                    torch.cuda.synchronize(dev)
                    g = torch.cuda.CUDAGraph()
                    n0 = lib.pmhc_launch_count()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        enqueue(gdesc, ins[7], st.scalars_dev.data_ptr(), not hooked)
                        if in_graph_hook:
                            self.grad_hook(st.grad)
                            _lib.check(lib.pmhc_train_step_adam(flat.data_ptr(), st.grad.data_ptr(), opt._m.data_ptr(), opt._v.data_ptr(), b1, b2, eps,
                                                                ctypes.byref(sc), st.scalars_dev.data_ptr(), self._nan_flag.data_ptr(), stream_of()),
                                       "pmhc_train_step_adam")
                    st.graph, st.graph_key, st.graph_launches = g, gkey, lib.pmhc_launch_count() - n0
                _lib.check(lib.pmhc_upload_small(ctypes.byref(sc), st.scalars_dev.data_ptr(), 48, stream_of()), "pmhc_upload_small")
                st.graph.replay()
                lib.pmhc_launch_count_add(st.graph_launches)
                if in_graph_hook:
                    hooked = False
            if hooked:
                self.grad_hook(st.grad)
                _lib.check(lib.pmhc_train_step_adam(flat.data_ptr(), st.grad.data_ptr(), opt._m.data_ptr(), opt._v.data_ptr(), b1, b2, eps,
                                                    ctypes.byref(sc), None, self._nan_flag.data_ptr(), stream_of()), "pmhc_train_step_adam")
        opt.count_step()

        # (views of the step's static buffers: valid until the next optimize() call on this batch shape)
        loss_dict = {k: st.losses[i] for i, k in enumerate(LOSS_KEYS)}
        if metrics is not None:
            if hasattr(metrics, "add_stacked"):
                metrics.add_stacked(LOSS_KEYS, st.losses)
            else:
                metrics.add_batch(loss_dict)
        self.last_losses = loss_dict
        self.last_prediction = (st.pred_f, st.pred_t)
        # the NaN flag is sticky and stays on the device (check_nan()); from the first NaN loss on, every Adam update is skipped
        # there, so the weights in memory (and whatever gets saved from them) stay the last finite ones — the reference raises
        # before backward() / step() (optimizer.py:217-218)
        for p, g in zip(model.parameters(), model._split_flat(st.grad)):
            p.grad = g

    def grad_hook(self, flat_grad: torch.Tensor) -> None:
        """Called with the flat gradient before the Adam step; data-parallel wrappers all-reduce here."""

    def layer2_event_handle(self):
        """cudaEvent_t (or None) the backward records once the gnn2.* gradients are final; see diffusion/parallel.py."""
        return None

    def check_nan(self) -> None:
        """The reference raises RuntimeError("NaN loss") inside optimize() (optimizer.py:217-218) at the price of a
        host sync per step; here the flag stays on the device until asked for."""
        if self._nan_flag is not None and bool(self._nan_flag.any()):
            raise RuntimeError("NaN loss")

    # ---- full training state (the reference only saves model.state_dict(), optimize.py:75-80) -----------
    def state_dict(self) -> Dict:
        """Everything a resume needs to continue the same run: weights, Adam moments and step count, and the two host random streams the
        training step draws from (Python's `random` for t, optimizer.py:197; torch's CPU generator for the noise keys)."""
        return {"model": self.model.state_dict(), "optimizer": self.optimizer.state_dict(), "noise_step_count": self.noise_step_count,
                "python_random": random.getstate(), "torch_rng": torch.get_rng_state()}

    def load_state_dict(self, state: Dict) -> None:
        if int(state["noise_step_count"]) != int(self.noise_step_count):
            raise ValueError(f"checkpoint was trained with T={state['noise_step_count']}, this optimizer has T={self.noise_step_count}")
        self.model.load_state_dict(state["model"], strict=True)
        self.optimizer.load_state_dict(state["optimizer"])
        random.setstate(state["python_random"])
        torch.set_rng_state(state["torch_rng"].cpu())       # (a checkpoint loaded with map_location=cuda moves it)

    # ---- sampling -------------------------------------------------------------------------------------------
    def sample(self, batch: Dict[str, Union[torch.Tensor, Rigid]], noise_tape: Optional[torch.Tensor] = None,
               quat_sign_tape: Optional[torch.Tensor] = None, graph: Optional[bool] = None) -> Dict:
        """optimizer.py:226-252: T sequential reverse steps, all enqueued by one pmhc_sample call.
        noise_tape [T, B, 16, 21] (tensor_7 + 14 torsion values) and quat_sign_tape [T, B, 16, 4] are parity hooks.
        graph (default: self.use_graph): replay the trajectory's 4 T + launches as ONE CUDA graph captured once per batch shape;
        the Philox (seed, first complex) pair is read from device memory, so every replay draws its own noise."""
        lib = _lib.load()
        batch["pocket_frames"] = Rigid.from_tensor_7(_frames7(batch["pocket_frames"]))
        batch["frames"] = Rigid.from_tensor_7(_frames7(batch["frames"]))
        if (self.use_graph if graph is None else graph) and noise_tape is None and quat_sign_tape is None:
            return self._sample_graphed(batch)
        frames = _lib.f32c(batch["frames"].to_tensor_7()).clone()
        tors = _lib.f32c(batch["torsions"]).clone()
        desc, keep = _lib.make_batch(frames, tors, batch["features"], batch["mask"], batch["pocket_frames"].to_tensor_7(),
                                     batch["pocket_features"], batch["pocket_mask"])
        dev = frames.device
        B, P = desc.B, desc.P
        T = self.noise_step_count
        scratch = torch.empty(2 * B * _lib.N * 21, device=dev, dtype=torch.float32)
        ws_bytes = lib.pmhc_workspace_bytes(B, P)
        ws = _lib.workspace(dev, ws_bytes)
        tape = _lib.f32c(noise_tape) if noise_tape is not None else None
        sign = _lib.f32c(quat_sign_tape) if quat_sign_tape is not None else None
        flat = self.model._flat_params()
        seed = self.sample_seed if self.sample_seed is not None else _next_noise_key()
        with torch.cuda.device(dev):
            _lib.check(lib.pmhc_sample(flat.data_ptr(), ctypes.byref(desc), frames.data_ptr(), tors.data_ptr(), T,
                                       self.beta_min, self.beta_max, seed, self.sample_first_complex, _lib.ptr(tape), _lib.ptr(sign),
                                       scratch.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr(dev),
                                       self.model.precision_code()), "pmhc_sample")
        result = {k: batch[k] for k in batch}
        result["frames"] = _rigid(frames)
        result["torsions"] = tors
        return result

    def _sample_graphed(self, batch: Dict) -> Dict:
        lib = _lib.load()
        desc, keep = _lib.make_batch(batch["frames"].to_tensor_7(), batch["torsions"], batch["features"], batch["mask"],
                                     batch["pocket_frames"].to_tensor_7(), batch["pocket_features"], batch["pocket_mask"])
        dev = keep[0].device
        B, P, T = desc.B, desc.P, self.noise_step_count
        flat = self.model._flat_params()
        ws_bytes = lib.pmhc_workspace_bytes(B, P)
        ws = _lib.workspace(dev, ws_bytes)
        key = (B, P, T, self.model.precision_code(), dev.index)
        sg = self._sample_graphs.get(key)
        valid = (flat.data_ptr(), ws.data_ptr(), self.beta_min, self.beta_max)
        with torch.cuda.device(dev):
            if sg is None or sg["valid"] != valid:
                if sg is None and len(self._sample_graphs) >= 4:
                    self._sample_graphs.pop(next(iter(self._sample_graphs)))
                ins = [torch.empty_like(t) for t in keep]
                scratch = torch.empty(2 * B * _lib.N * 21, device=dev, dtype=torch.float32)
                seed_dev = torch.zeros(2, dtype=torch.int64, device=dev)
                gdesc = _lib.PmhcBatch(B, P, *[t.data_ptr() for t in ins])

                def enqueue():
                    _lib.check(lib.pmhc_sample_ex(flat.data_ptr(), ctypes.byref(gdesc), ins[0].data_ptr(), ins[1].data_ptr(), T, self.beta_min,
                                                  self.beta_max, 0, 0, seed_dev.data_ptr(), None, None, scratch.data_ptr(), ws.data_ptr(),
                                                  ws_bytes, _lib.stream_ptr(dev), self.model.precision_code()), "pmhc_sample")
                for dst, src in zip(ins, keep):
                    dst.copy_(src)
                enqueue()                      # once outside capture: lazy kernel-attribute set-up happens here
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                n0 = lib.pmhc_launch_count()
                with torch.cuda.graph(g):
                    enqueue()
                sg = self._sample_graphs[key] = {"graph": g, "ins": ins, "scratch": scratch, "seed": seed_dev, "valid": valid, "ws": ws,
                                                 "launches": lib.pmhc_launch_count() - n0}
            for dst, src in zip(sg["ins"], keep):
                dst.copy_(src, non_blocking=True)
            seed = self.sample_seed if self.sample_seed is not None else _next_noise_key()
            pair = (ctypes.c_uint64 * 2)(int(seed) & (2 ** 64 - 1), int(self.sample_first_complex))
            _lib.check(lib.pmhc_upload_small(pair, sg["seed"].data_ptr(), 16, _lib.stream_ptr(dev)), "pmhc_upload_small")
            sg["graph"].replay()
            lib.pmhc_launch_count_add(sg["launches"])
            frames, tors = sg["ins"][0].clone(), sg["ins"][1].clone()
        result = {k: batch[k] for k in batch}
        result["frames"] = _rigid(frames)
        result["torsions"] = tors
        return result
