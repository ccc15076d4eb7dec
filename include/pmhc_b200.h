/*
 * pmhc_b200.h — C ABI of the B200-native denoising hot path of cmbi/pmhc-diffusion-model.
 *
 * The reference has no FFI layer: its boundary is the Python surface of
 * diffusion/model.py and diffusion/optimizer.py.  Each entry point below names the
 * reference interface (file:line under /root/reference) whose arithmetic it
 * replaces; pmhc_diffusion_model_b200/ binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), nothing
 *     allocates, nothing synchronises, no global mutable state;
 *   - return value: 0 = ok, < 0 = error (text in pmhc_last_error(), thread-local);
 *   - sm_100a only: there is no CPU or other-arch fallback.
 *   - N = 16 padded peptide slots, 22 node features, 7 torsions (sin, cos),
 *     frames are tensor_7 rows: quaternion (w, x, y, z) then translation (x, y, z).
 */
#ifndef PMHC_B200_H
#define PMHC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMHC_N 16          /* peptide_maxlen, data.py:15 / Model(16, ...) optimize.py:54 */
#define PMHC_NFEAT 22      /* node_input_size, optimize.py:54 */
#define PMHC_NTORS 7       /* model.py:32 */
#define PMHC_HID 64        /* transition / message / feature width, model.py:36, 367-368 */
#define PMHC_NPARAM 79195  /* total fp32 parameters of Model(16, 22, T) */
#define PMHC_ROWSTAT 16    /* floats saved per peptide row per layer for the backward pass */

/* Arithmetic of the two dense per-pair contractions of the denoiser (message_mlp.2 and the four head hidden layers):
 *   FP32: fp32 FFMA, the exact-parity mode (<= 1e-4 vs the reference);
 *   BF16: tcgen05 tensor cores, bf16 operands, fp32 accumulation in TMEM (<= 1e-2); everything else stays fp32. */
#define PMHC_PRECISION_FP32 0
#define PMHC_PRECISION_BF16 1
/*   TC32: tcgen05 tensor cores with every operand written as two fp16 terms (x = hi + lo, ~22 bits) and every contraction as
 *         hi.hi + lo.hi + hi.lo with fp32 accumulation in TMEM: fp32-class results (gate: the FP32 mode's own, max(1e-4, 2 x the
 *         reference's fp32 noise floor)) — the tensor-core mode that holds on the reference's shipped model.pth;
 *   FP16: the same pipeline with single fp16 terms (11 bits; 1e-2 class on well-conditioned weights).
 * As BACKWARD modes (pmhc_model_backward_ex, pmhc_train_step_grad): FP32 = the FFMA backward, FP16 = the backward on tcgen05
 *   (fp16 operand tiles, fp32 accumulation in tensor memory, 1e-2 class), BF16 = the older warp-level TF32 mma.sync backward;
 *   TC32 is not a backward mode (a TC32 forward pairs with the FP32 backward, or with FP16 for 1e-2-class gradients). */
#define PMHC_PRECISION_TC32 2
#define PMHC_PRECISION_FP16 3

/* One batch of complexes in the dataset's padded layout (data.py:105-117, model.py:384-390). */
typedef struct {
    int32_t B;                    /* complexes */
    int32_t P;                    /* padded pocket slots per complex (data.py:16 pocket_maxlen) */
    const float *frames;          /* [B,16,7]  peptide frames, used as given (RU:1037 from_tensor_7) */
    const float *torsions;        /* [B,16,7,2] sin, cos */
    const float *features;        /* [B,16,22] */
    const uint8_t *mask;          /* [B,16]  1 = real residue */
    const float *pocket_frames;   /* [B,P,7] */
    const float *pocket_features; /* [B,P,22] */
    const uint8_t *pocket_mask;   /* [B,P] */
} PmhcBatch;

const char *pmhc_last_error(void);

/* Library / device probe: returns 0 when the current device is sm_100 and the kernels are loadable. */
int pmhc_check_device(void);

/* Offset (in floats) of state-dict tensor `index` (0..47, state_dict order of model.py:370-371:
 * gnn1.feature_mlp.0.weight, .0.bias, .2.weight, .2.bias, message_mlp.*, attention_mlp.*, translation_mlp.*,
 * rotation_mlp.*, torsion_mlp.*, then gnn2.*) inside the flat parameter buffer; -1 if out of range. */
int64_t pmhc_param_offset(int index);
int64_t pmhc_param_numel(int index);

/* Bytes of workspace pmhc_model_forward/backward need for a batch shape. */
size_t pmhc_workspace_bytes(int B, int P);

/* Denoiser forward — replaces Model.forward (diffusion/model.py:377-421) and both EGNNLayer.forward calls
 * (model.py:83-181): two fused message-passing layers over peptide x (peptide + pocket).
 *   params      flat fp32 parameter buffer (PMHC_NPARAM floats, pmhc_param_offset layout)
 *   t_over_T    the time feature t / T (model.py:394)
 *   out_frames  [B,16,7]   unit quaternion + translation (model.py:181, 418-421)
 *   out_torsions[B,16,7,2]
 *   saved       NULL for inference; else pmhc_saved_floats(B, P) floats kept for pmhc_model_backward
 *               (per-row softmax statistics, per-pair logits, layer-1 outputs)
 * Padded peptide rows (mask 0) pass their input frame/torsions through (the reference leaves finite
 * don't-care values there). */
size_t pmhc_saved_floats(int B, int P);
int pmhc_model_forward(const float *params, const PmhcBatch *batch_host, float t_over_T,
                       float *out_frames, float *out_torsions, float *saved,
                       void *workspace, size_t workspace_bytes, void *stream);
/* Same with an explicit PMHC_PRECISION_* mode (pmhc_model_forward == PMHC_PRECISION_FP32). */
int pmhc_model_forward_ex(const float *params, const PmhcBatch *batch_host, float t_over_T,
                          float *out_frames, float *out_torsions, float *saved,
                          void *workspace, size_t workspace_bytes, void *stream, int precision);

/* Denoiser backward — the autograd of the above (what total_loss.mean().backward() does at
 * optimizer.py:222 for the model part).  Accumulates (+=) into flat_grad (PMHC_NPARAM floats, same
 * layout as params; gnn2.feature_mlp.* never receives a gradient).
 *   d_out_frames [B,16,7], d_out_torsions [B,16,7,2]: gradient of the loss w.r.t. the forward outputs. */
int pmhc_model_backward(const float *params, const PmhcBatch *batch_host, float t_over_T,
                        const float *saved, const float *d_out_frames, const float *d_out_torsions,
                        float *flat_grad, void *workspace, size_t workspace_bytes, void *stream,
                        void *layer2_done_event);
/* Same with an explicit PMHC_PRECISION_* mode (pmhc_model_backward == PMHC_PRECISION_FP32).  PMHC_PRECISION_BF16 selects the
 * tensor-core backward: every 64-wide contraction of the pair recomputation, the input gradients and the weight-gradient
 * outer products as TF32 MMAs with fp32 accumulation (operands rounded to tf32, i.e. more mantissa than the bf16 forward);
 * second layers, geometry, softmax backward and all reductions stay fp32.  Gradient gate: the 1e-2 class.
 * PMHC_PRECISION_FP16 selects the backward on the Blackwell tensor-core path (tcgen05.mma, accumulators in tensor memory): the
 * message layer folded into the head weights, fp16 operand tiles used K-major and MN-major, every weight-gradient sum of a CTA
 * resident in tensor memory, work dealt to the SMs by 128-pair pass; per layer it enqueues the setup, schedule, pair, node,
 * reduce and unfold kernels (+ layer 1's feature pre-kernel) on `stream`.  Same gate; 3.7x the TF32 kernel's speed.  When
 * layer2_done_event is given it leaves eight SMs to the collective the caller overlaps with the layer-1 launch. */
int pmhc_model_backward_ex(const float *params, const PmhcBatch *batch_host, float t_over_T,
                           const float *saved, const float *d_out_frames, const float *d_out_torsions,
                           float *flat_grad, void *workspace, size_t workspace_bytes, void *stream,
                           void *layer2_done_event, int precision);
/* layer2_done_event (nullable cudaEvent_t): recorded on `stream` as soon as the gnn2.* half of flat_grad is final
 * (the gnn1.* half follows), so a data-parallel caller can start all-reducing it on a second stream while the
 * layer-1 backward kernel is still running. */

/* Noise draw — replaces DiffusionModelOptimizer.gen_noise (optimizer.py:93-108) with random_quat /
 * shoemake_quat (angle.py:59-98) and random_sin_cos (angle.py:33-57): translation 5*N(0,I), uniform
 * rotation, 7 uniform torsion angles per residue.  Counter-based (Philox4x32-10): residue r of the call
 * uses counter (first_residue + r, draw), so results do not depend on how residues are split over GPUs.
 *   noise_frames [n,7], noise_torsions [n,7,2]. */
int pmhc_gen_noise(uint64_t seed, uint64_t first_residue, int64_t n_residues,
                   float *noise_frames, float *noise_torsions, void *stream);

/* Noise from caller-supplied randoms (parity mode): normal [n,3] ~ N(0,1), uniform [n,10] ~ U(0,1)
 * (3 Shoemake coordinates then 7 torsion angles / 2pi) -> same formulas as above. */
int pmhc_noise_from_randoms(const float *normal, const float *uniform, int64_t n_residues,
                            float *noise_frames, float *noise_torsions, void *stream);

/* Forward noising — replaces add_noise (optimizer.py:110-138; partial_sin_cos angle.py:165-174,
 * partial_rot angle.py:177-186, multiply_sin_cos angle.py:139-152, compose_r RU:525-538).
 * Rotations stay quaternions: q_t = partial_rot(eps_q, beta) * q_0 (Hamilton product), equal to the
 * reference's compose_r + rot_to_quat up to the eigh sign.  quat_sign_ref (nullable,
 * [n,4]): if given, q_t is flipped where q_t . ref < 0 (parity mode: the reference's sign tape). */
int pmhc_add_noise(const float *frames, const float *torsions, const float *noise_frames,
                   const float *noise_torsions, double beta, int64_t n_residues,
                   const float *quat_sign_ref, float *out_frames, float *out_torsions, void *stream);

/* Reverse step t -> s — replaces remove_noise (optimizer.py:140-193; inverse_sin_cos angle.py:155-162,
 * Rotation.invert RU:585-601).  fresh_* is the step's new noise (optimizer.py:151).
 * beta is the schedule value as the reference's Python float (double): the angle scalings use it rounded to
 * fp32 (tensor * python-float semantics), the translation coefficients are derived from it in double.
 * In-place safe (out may alias zt). */
int pmhc_remove_noise(const float *zt_frames, const float *zt_torsions, const float *pred_frames,
                      const float *pred_torsions, const float *fresh_frames, const float *fresh_torsions,
                      double beta_t, double beta_s, int64_t n_residues, const float *quat_sign_ref,
                      float *out_frames, float *out_torsions, void *stream);

/* Loss forward + gradient — replaces get_loss (optimizer.py:38-79) and its autograd.
 *   losses [5,B]: total, positions, rotations, torsions, rmsd (optimizer.py:73-79)
 *   d_pred_frames [B,16,7], d_pred_torsions [B,16,7,2] (nullable): d(grad_scale * sum_b total[b]) / d pred;
 *   the training step uses grad_scale = 1/B for total_loss.mean() (optimizer.py:222). */
int pmhc_loss(const float *true_frames, const float *true_torsions, const float *pred_frames,
              const float *pred_torsions, const uint8_t *mask, const uint8_t *torsions_mask, int B,
              float grad_scale, float *losses, float *d_pred_frames, float *d_pred_torsions, void *stream);

/* Whole sampling trajectory — replaces DiffusionModelOptimizer.sample (optimizer.py:226-252): T sequential
 * (denoiser forward, reverse step) pairs on `stream`, updating frames/torsions in place.
 *   frames [B,16,7], torsions [B,16,7,2]: z_T on entry (test.py:71-74), z_0 on return
 *   noise_tape (nullable): [T, B*16, 21] = per step fresh noise as tensor_7 + 14 torsion values;
 *                          NULL -> Philox with (seed, first_complex)
 *   quat_sign_tape (nullable): [T, B*16, 4] reference quaternions of z after each step (parity mode)
 *   scratch: 2 * B*16*21 floats for (pred, fresh). */
int pmhc_sample(const float *params, const PmhcBatch *batch_host, float *frames, float *torsions,
                int T, double beta_min, double beta_max, uint64_t seed, uint64_t first_complex,
                const float *noise_tape, const float *quat_sign_tape, float *scratch,
                void *workspace, size_t workspace_bytes, void *stream, int precision);

/* Same trajectory with the Philox (seed, first_complex) pair read from device memory (seed_first_dev: 2 x uint64, nullable):
 * the call only enqueues kernels, so it can be captured into a CUDA graph once per (batch shape, T, precision) and replayed
 * with new noise by rewriting those 16 bytes — DiffusionModelOptimizer.sample(graph=True). */
int pmhc_sample_ex(const float *params, const PmhcBatch *batch_host, float *frames, float *torsions,
                   int T, double beta_min, double beta_max, uint64_t seed, uint64_t first_complex,
                   const uint64_t *seed_first_dev, const float *noise_tape, const float *quat_sign_tape,
                   float *scratch, void *workspace, size_t workspace_bytes, void *stream, int precision);

/* Adam update over the flat buffers — replaces torch.optim.Adam.step (optimizer.py:33, 224), default
 * betas/eps unless given; `skip` ranges (gnn2.feature_mlp) are left untouched like grad=None params. */
int pmhc_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n,
                   double lr, double beta1, double beta2, double eps, int step, void *stream);
/* Same, guarded on the device: when *skip_flag (one byte, nullable) is non-zero nothing is written.  The training step sets
 * the flag from `total_loss.isnan().any()`: the reference raises RuntimeError("NaN loss") BEFORE backward() and step()
 * (optimizer.py:217-218), so its weights never see a NaN gradient; here the check stays on the device (no host sync per
 * step) and the update is skipped instead. */
int pmhc_adam_step_guarded(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n,
                           double lr, double beta1, double beta2, double eps, int step,
                           const uint8_t *skip_flag, void *stream);

/* ---- one training step as two enqueue-only calls — DiffusionModelOptimizer.optimize (optimizer.py:195-224) -------------
 * pmhc_train_step_grad: [gen_noise] -> add_noise -> denoiser forward -> loss + its gradient -> NaN flag -> flat_grad = 0 ->
 * denoiser backward; pmhc_train_step_adam: the guarded Adam update of every parameter that carries a gradient.  (Between the
 * two a data-parallel caller all-reduces flat_grad.)  All per-step scalars live in one 48-byte block: given by value
 * (scalars_host, always required) and, optionally, in device memory (scalars_dev): then every kernel reads the device copy, and
 * a CUDA graph captured around the two calls is replayable for any later step — refresh the block, launch the graph. */
typedef struct {
    float t_over_T;               /* time feature t / T (model.py:394) */
    float beta, alpha, sigma;     /* add_noise at step t: beta_t, sqrt(1 - beta_t), sqrt(beta_t) (optimizer.py:81-91, 110-138) */
    float adam_step_size;         /* lr / (1 - beta1^k) */
    float adam_bc2_sqrt;          /* sqrt(1 - beta2^k) */
    float grad_scale;             /* d(mean loss) / d(per-complex loss): 1 / B, or 1 / B_global under data parallelism */
    float reserved;
    uint64_t noise_seed;          /* Philox key of this step's noise */
    uint64_t noise_first_residue; /* counter of the batch's first residue (16 x its first global complex) */
} PmhcStepScalars;
/* Fills *out_host for noise step t of T and Adam step k, forming every scalar in double and rounding once (as the reference's
 * Python floats / torch's Adam do). */
int pmhc_step_scalars(int t, int T, double beta_min, double beta_max, double lr, double beta1, double beta2,
                      int adam_step, double grad_scale, uint64_t noise_seed, uint64_t noise_first_residue,
                      PmhcStepScalars *out_host);
/* Copies 1..64 bytes from host to device memory through a kernel's launch parameters (captured when the call returns, ordered
 * on `stream`): how the scalar block above — and pmhc_sample_ex's seed pair — is refreshed before a graph replay. */
void pmhc_launch_count_add(int64_t n);   /* kernels launched by a replayed CUDA graph (counted at capture) join pmhc_launch_count() */
int pmhc_upload_small(const void *src_host, void *dst_dev, int bytes, void *stream);
typedef struct {                  /* caller-owned device buffers of one step */
    float *noise_frames;          /* [B,16,7]   epsilon: written when draw_noise != 0, else read */
    float *noise_torsions;        /* [B,16,7,2] */
    float *zt_frames;             /* [B,16,7]   noised input (out) */
    float *zt_torsions;           /* [B,16,7,2] */
    float *pred_frames;           /* [B,16,7]   predicted noise (out) */
    float *pred_torsions;         /* [B,16,7,2] */
    float *d_frames;              /* [B,16,7]   loss gradient (scratch) */
    float *d_torsions;            /* [B,16,7,2] */
    float *losses;                /* [5,B]      total, positions, rotations, torsions, rmsd (out) */
    float *saved;                 /* pmhc_saved_floats(B, P) */
    float *flat_grad;             /* PMHC_NPARAM, overwritten */
    uint8_t *nan_flag;            /* nullable, 1 byte, sticky: set when any total loss is NaN */
} PmhcStepBuffers;
/* batch->frames / torsions are the CLEAN x_0; quat_sign_ref (nullable, [B*16,4]) as in pmhc_add_noise. */
int pmhc_train_step_grad(const float *params, const PmhcBatch *batch_host, const uint8_t *torsions_mask,
                         const PmhcStepScalars *scalars_host, const PmhcStepScalars *scalars_dev,
                         const PmhcStepBuffers *buffers_host, int draw_noise, const float *quat_sign_ref,
                         void *workspace, size_t workspace_bytes, void *stream, void *layer2_done_event,
                         int precision, int backward_precision);
int pmhc_train_step_adam(float *params, const float *flat_grad, float *exp_avg, float *exp_avg_sq,
                         double beta1, double beta2, double eps, const PmhcStepScalars *scalars_host,
                         const PmhcStepScalars *scalars_dev, const uint8_t *nan_flag, void *stream);

/* Loader side — replaces the per-entry Rigid.from_tensor_4x4(...).to_tensor_7() of MhcpDataset.get_entry
 * (diffusion/data.py:107, :115; RU:1004-1034 + rot_to_quat RU:184-216): n homogeneous 4x4 matrices (row-major,
 * rotation in [0:3,0:3], translation in [0:3,3]) -> tensor_7 rows.  The quaternion is unit with w >= 0; the reference's
 * eigh-based rot_to_quat returns the same quaternion up to sign. */
int pmhc_frames4x4_to_tensor7(const float *frames4x4, int64_t n, float *out7, void *stream);

/* Writer side — the coordinate arithmetic of tools/pdb.py::save (pdb.py:67-174): torsion_angles_to_frames and
 * frames_and_literature_positions_to_atom14_pos (openfold.utils.feats), backbone N / CA / C / CB from the normalised
 * frame, backbone O from the next residue's N (pdb.py:139-151), terminal O and OXT from the psi frame (pdb.py:153-174).
 *   frames [B,16,7], torsions [B,16,7,2], aatype [B,16] int64, mask [B,16]
 *   default_frames [21,8,4,4], group_idx [21,14] int32, lit_positions [21,14,3], atom_mask [21,14] u8:
 *       openfold.np.residue_constants.restype_rigid_group_default_frame / restype_atom14_to_rigid_group /
 *       restype_atom14_rigid_group_positions / restype_atom14_mask (pdb.py:47-66)
 *   positions [B,16,15,3]: atom14 order, slot 14 = OXT;  exists [B,16,15] (0 on padded residues). */
int pmhc_atom14(const float *frames, const float *torsions, const int64_t *aatype, const uint8_t *mask, int B,
                const float *default_frames, const int32_t *group_idx, const float *lit_positions,
                const uint8_t *atom_mask, float *positions, uint8_t *exists, void *stream);

/* Host-side text of one complex in the fixed PDB columns, as tools/pdb.py:206-209 gets it from BioPython's PDBIO: chain P =
 * the peptide's atoms in the order the reference adds them (N, CA, C, CB, side chain, O [, OXT]; pdb.py:112-174), chain M = the
 * protein's existing atom14 slots (pdb.py:177-204), serial numbers from 1, a TER record per chain, END.  HOST pointers.
 *   pep_*: one peptide: aatype [16], mask [16], positions [16,15,3], exists [16,15] (pmhc_atom14 layout)
 *   prot_*: aatype [n_prot], positions [n_prot,14,3], exists [n_prot,14]
 *   atom_fields [21,15,4] chars, elements [21,15] chars, res3 [21,3] chars per (residue type, atom slot), not terminated
 * Returns the bytes written to `out`, or -(bytes needed) when out_cap is too small. */
int64_t pmhc_format_pdb_host(const int64_t *pep_aatype, const uint8_t *pep_mask, const float *pep_pos, const uint8_t *pep_exists,
                             int64_t n_prot, const int64_t *prot_aatype, const float *prot_pos, const uint8_t *prot_exists,
                             const char *atom_fields, const char *elements, const char *res3, char *out, int64_t out_cap);

/* Number of kernel launches issued by this library since load (for bench.py's gpu_launches claim). */
int64_t pmhc_launch_count(void);

/* Measurement hooks (bench.py's roofline leg; no reference counterpart): while enabled, every fused EGNN layer
 * launch is bracketed by CUDA events on its stream.  pmhc_profile_read waits for them and returns, for slot 0
 * (forward layer kernels) and slot 1 (backward layer kernels), the summed device time in ms and the launch count,
 * then clears the record. */
void pmhc_profile_enable(int on);
int pmhc_profile_read(double *ms_by_slot, int64_t *launches_by_slot);

#ifdef __cplusplus
}
#endif
#endif /* PMHC_B200_H */
