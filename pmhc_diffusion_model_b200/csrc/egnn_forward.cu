// egnn_forward.cu — fused EGNN layer forward (replaces EGNNLayer.forward, diffusion/model.py:83-333, and the
// glue of Model.forward, model.py:377-421).
//
// One persistent CTA per SM loops over complexes.  Per complex: frames / features / per-node projections are
// staged in shared memory once; every (peptide row i, neighbour j) pair is owned by one thread slot which
// carries it through message MLP -> four head MLPs -> quaternion sandwich entirely in registers, reading the
// layer's weights as warp-uniform 128-bit shared-memory loads; per-pair head outputs go to a shared buffer and
// one warp per row does the masked softmax and the attention-weighted frame / torsion / translation updates.
// Nothing of size [B, N, K, *] ever touches HBM (the reference materialises ~15 such tensors per layer).
#include "egnn_common.cuh"

namespace pmhc {

// Compiler-only fence: keeps ptxas from hoisting all 64 units' shared loads of an unrolled head epilogue
// ahead of the FMAs (which spilled the 2 x 64 accumulators); placed every 8 units.
#define PMHC_SCHED_FENCE(n) do { if (((n) & 7) == 7) asm volatile("" ::: "memory"); } while (0)

// Two threads per pair: threads p and p + 256 (warps w and w + 8) own column p of the pass.  Each computes the half of every
// 64-wide result whose index lies in its half h (n in [32h, 32h + 32)); every output is still produced by ONE thread with the
// same k order as before, and the 64-term second-layer sums are continued by the second thread from the first thread's
// partial — so the results are bit-identical to the one-thread-per-pair form, at twice the warps per scheduler.
__device__ __forceinline__ void pair_sync8(int warp8) { asm volatile("bar.sync %0, 64;" ::"r"(1 + warp8) : "memory"); }

// acc[nn] += sum_k WT[k*ldw + n0 + nn] * scr[k*kScrLd + col]  for k in [0,64), nn in [0,32)
__device__ __forceinline__ void gemv32(float (&acc)[32], const float* __restrict__ WT, int ldw, const float* __restrict__ scr, int col) {
#pragma unroll 2
    for (int k = 0; k < kHid; ++k) {
        const float av = scr[k * kScrLd + col];
        const float4* w4 = reinterpret_cast<const float4*>(WT + k * ldw);
#pragma unroll
        for (int n4 = 0; n4 < 8; ++n4) {
            const float4 w = w4[n4];
            acc[4 * n4 + 0] = fmaf(w.x, av, acc[4 * n4 + 0]);
            acc[4 * n4 + 1] = fmaf(w.y, av, acc[4 * n4 + 1]);
            acc[4 * n4 + 2] = fmaf(w.z, av, acc[4 * n4 + 2]);
            acc[4 * n4 + 3] = fmaf(w.w, av, acc[4 * n4 + 3]);
        }
    }
}

// message MLP of the thread's pair: scr[:, col] <- m = W2 relu(A_i + A_j + W_e) + b2  (model.py:183-226)
__device__ __forceinline__ void message_stage(float* S, const SmemMap& M, const float* __restrict__ ajt, int Kpad, const PairRef& pr,
                                              int col, int half, int warp8) {
    float* scr = S + M.Scr;
    const int n0 = 32 * half;
    {
        const int i = pr.i, j = pr.j;
        const float* ai = S + M.Ai + i * kLdN + n0;
        const bool pep = (j >= 0 && j < kN);
        const float* we = S + M.We + (pep ? (kN - 1 + i - j) : 0) * kLdN + n0;
        // all L2 loads of the A_j^T column are issued before the first use (__ldcg is an ordered asm: mixed
        // with its consumer it serialises into dependent L2 round trips, 26 % of the kernel in the first profile)
        float aj[32];
        const float* ajc = ajt + (j >= 0 ? j : 0) + (size_t)n0 * Kpad;
#pragma unroll
        for (int k = 0; k < 32; ++k) aj[k] = __ldcg(ajc + k * Kpad);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            float v = ai[k];
            if (j >= 0) v += aj[k];
            if (pep) v += we[k];
            scr[(n0 + k) * kScrLd + col] = fmaxf(v, 0.0f);
        }
    }
    pair_sync8(warp8);          // m1 complete
    float acc[32];
#pragma unroll
    for (int n = 0; n < 32; ++n) acc[n] = S[M.PkMisc + 4 * (n0 + n) + 3];
    gemv32(acc, S + M.W2T + n0, kHid, scr, col);
    pair_sync8(warp8);          // both halves have read m1: the column may be overwritten with m
#pragma unroll
    for (int n = 0; n < 32; ++n) scr[(n0 + n) * kScrLd + col] = acc[n];
    pair_sync8(warp8);          // m complete
}

// the four heads of one pair; writes kOutPerPair floats  (model.py:228-333).  The first half seeds each 64-term second-layer
// sum with the bias and its 32 terms and parks it in the pair's output slot; the second half continues it and finishes the head.
__device__ __forceinline__ void heads_stage(float* S, const SmemMap& M, const PairRef& pr, int col, int oslot, int half, int warp8,
                                            float* __restrict__ logit_save, int Kpad) {
    const float* scr = S + M.Scr;
    const int n0 = 32 * half;
    float* out = S + M.Out + oslot * kOutPerPair;
    const bool act = pr.active;
    float acc[32];
    // ---- attention logit (model.py:238-242) ----
    // The distance / orientation inputs (-d2 reaches thousands of A^2) are added AFTER the 64-term message
    // contraction: one rounding at that magnitude instead of 64 keeps the logit within a few ulp of a blocked
    // CPU summation (logits of the shipped weights reach 2.5e3, where one fp32 ulp is already 2.4e-4).
    const float* qi_ = S + M.Q + pr.i * 4;
    const float* qj_ = S + M.Q + pr.j * 4;
    const float* xi = S + M.X + pr.i * 3;
    const float* xj = S + M.X + pr.j * 3;
    const float dx = xi[0] - xj[0], dy = xi[1] - xj[1], dz = xi[2] - xj[2];
    const float ex_d2 = dx * dx + dy * dy + dz * dz;
    const float dotq = qi_[0] * qj_[0] + qi_[1] * qj_[1] + qi_[2] * qj_[2] + qi_[3] * qj_[3];
    const float ex_qd = dotq * dotq;
#pragma unroll
    for (int n = 0; n < 32; ++n) acc[n] = S[M.PkAtt + 4 * (n0 + n) + 2];
    gemv32(acc, S + M.WhT + n0, 256, scr, col);
    float logit = S[M.Scal + SC_ATT2B];
    if (half == 1) {
        pair_sync8(warp8);
        logit = out[0];
    }
#pragma unroll
    for (int n = 0; n < 32; ++n) {
        const float4 pk = *reinterpret_cast<const float4*>(S + M.PkAtt + 4 * (n0 + n));
        const float h = acc[n] + fmaf(pk.y, ex_qd, pk.x * -ex_d2);
        logit = fmaf(pk.w, fmaxf(h, 0.0f), logit);
        PMHC_SCHED_FENCE(n);
    }
    if (act) {
        out[0] = logit;
        if (half == 1 && logit_save != nullptr) logit_save[pr.i * Kpad + pr.j] = logit;
    }
    if (half == 0) pair_sync8(warp8);
    // ---- rotation: local frame -> MLP -> sigmoid -> back to the global frame (model.py:283-296) ----
    const Quat qi{qi_[0], qi_[1], qi_[2], qi_[3]}, qj{qj_[0], qj_[1], qj_[2], qj_[3]};
    {
        const Quat lq = qmul(qinv(qj), qmul(qi, qj));
#pragma unroll
        for (int n = 0; n < 32; ++n) {
            const float4 wq = *reinterpret_cast<const float4*>(S + M.PkRotQ + 4 * (n0 + n));
            float v = S[M.PkMisc + 4 * (n0 + n) + 2];
            v = fmaf(wq.x, lq.w, v);
            v = fmaf(wq.y, lq.x, v);
            v = fmaf(wq.z, lq.y, v);
            v = fmaf(wq.w, lq.z, v);
            acc[n] = v;
            PMHC_SCHED_FENCE(n);
        }
    }
    gemv32(acc, S + M.WhT + 64 + n0, 256, scr, col);
    {
        float pre[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) pre[c] = S[M.Scal + SC_ROT2B + c];
        if (half == 1) {
            pair_sync8(warp8);
#pragma unroll
            for (int c = 0; c < 4; ++c) pre[c] = out[1 + c];
        }
#pragma unroll
        for (int n = 0; n < 32; ++n) {
            const float h = fmaxf(acc[n], 0.0f);
            const float4 w = *reinterpret_cast<const float4*>(S + M.PkRot2 + 4 * (n0 + n));
            pre[0] = fmaf(w.x, h, pre[0]);
            pre[1] = fmaf(w.y, h, pre[1]);
            pre[2] = fmaf(w.z, h, pre[2]);
            pre[3] = fmaf(w.w, h, pre[3]);
            PMHC_SCHED_FENCE(n);
        }
        if (half == 0) {
            if (act) {
#pragma unroll
                for (int c = 0; c < 4; ++c) out[1 + c] = pre[c];
            }
            pair_sync8(warp8);
        } else {
            const Quat dl{sigmoidf(pre[0]), sigmoidf(pre[1]), sigmoidf(pre[2]), sigmoidf(pre[3])};  // never normalised (T5)
            const Quat dg = qmul(qj, qmul(dl, qinv(qj)));
            if (act) {
                out[1] = dg.w; out[2] = dg.x; out[3] = dg.y; out[4] = dg.z;
            }
        }
    }
    // ---- torsion angle increments (model.py:257-260) ----
#pragma unroll
    for (int n = 0; n < 32; ++n) acc[n] = S[M.Tt + pr.i * kHid + n0 + n];
    gemv32(acc, S + M.WhT + 128 + n0, 256, scr, col);
    {
        float da[PMHC_NTORS];
#pragma unroll
        for (int c = 0; c < PMHC_NTORS; ++c) da[c] = S[M.Scal + SC_TOR2B + c];
        if (half == 1) {
            pair_sync8(warp8);
#pragma unroll
            for (int c = 0; c < PMHC_NTORS; ++c) da[c] = out[5 + c];
        }
#pragma unroll
        for (int n = 0; n < 32; ++n) {
            const float h = fmaxf(acc[n], 0.0f);
            const float4 w0 = *reinterpret_cast<const float4*>(S + M.PkTor2 + 8 * (n0 + n));
            const float4 w1 = *reinterpret_cast<const float4*>(S + M.PkTor2 + 8 * (n0 + n) + 4);
            da[0] = fmaf(w0.x, h, da[0]);
            da[1] = fmaf(w0.y, h, da[1]);
            da[2] = fmaf(w0.z, h, da[2]);
            da[3] = fmaf(w0.w, h, da[3]);
            da[4] = fmaf(w1.x, h, da[4]);
            da[5] = fmaf(w1.y, h, da[5]);
            da[6] = fmaf(w1.z, h, da[6]);
            PMHC_SCHED_FENCE(n);
        }
        if (act) {
#pragma unroll
            for (int c = 0; c < PMHC_NTORS; ++c) out[5 + c] = da[c];
        }
        if (half == 0) pair_sync8(warp8);
    }
    // ---- translation scale (model.py:325-331) ----
#pragma unroll
    for (int n = 0; n < 32; ++n) acc[n] = S[M.PkMisc + 4 * (n0 + n) + 0];
    gemv32(acc, S + M.WhT + 192 + n0, 256, scr, col);
    {
        float sc = S[M.Scal + SC_TRN2B];
        if (half == 1) {
            pair_sync8(warp8);
            sc = out[12];
        }
#pragma unroll
        for (int n = 0; n < 32; ++n) sc = fmaf(S[M.PkMisc + 4 * (n0 + n) + 1], fmaxf(acc[n], 0.0f), sc);
        if (half == 0) {
            if (act) out[12] = sc;
            pair_sync8(warp8);
        } else if (act) {
            out[12] = sc * dx;
            out[13] = sc * dy;
            out[14] = sc * dz;
        }
    }
}

// Adds this pass's messages into the per-row unmasked sums (model.py:151, trap T3).  `mult_last` is the
// multiplicity of the last pair of every row (the shared zero-feature pocket message), 1 otherwise.
__device__ __forceinline__ void accumulate_msum(float* S, const SmemMap& M, const int* I, int row0, int nrows, int W,
                                                int pass_base, int npass, int mult_last_e, float mult_last) {
    const float* scr = S + M.Scr;
    for (int idx = threadIdx.x; idx < nrows * kHid; idx += kFwdThreads) {
        int rl = idx >> 6, n = idx & 63;
        int lo = max(rl * W, pass_base), hi = min((rl + 1) * W, pass_base + npass);
        float sum = 0.0f;
        for (int gp = lo; gp < hi; ++gp) {
            float v = scr[n * kScrLd + (gp - pass_base)];
            if (gp - rl * W == mult_last_e) v *= mult_last;
            sum += v;
        }
        if (hi > lo) S[M.Msum + I[IN_ROWS + row0 + rl] * kHid + n] += sum;
    }
}

template <int LAYER>
__global__ void __launch_bounds__(kFwdThreads, 1) egnn_layer_forward_kernel(LayerArgs a) {
    extern __shared__ __align__(16) float S[];
    const SmemMap M = make_smem_map(a.Kpad);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int pcol = tid & (kPassPairs - 1), half = tid >> 8, warp8 = warp & 7;   // two threads (tid, tid + 256) per pair column
    int* I = reinterpret_cast<int*>(S + M.Ints);
    float* ajt = a.ajt_ws + (size_t)blockIdx.x * kHid * a.Kpad;

    stage_layer_weights<LAYER>(S, M, a.params);
    __syncthreads();

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const ComplexInfo ci = setup_complex<LAYER>(S, M, a, b, ajt);
        const int L = ci.L;
        const int W = (L - 1) + ci.nv;  // attention-carrying neighbours of every real row
        float* lsave = a.logit_out ? a.logit_out + (size_t)b * kN * a.Kpad : nullptr;

        // padded rows: pass the inputs through (finite don't-care values)
        for (int idx = tid; idx < (kN - L) * 21; idx += kFwdThreads) {
            int s = idx / 21, c = idx - s * 21;
            int i = I[IN_PEPX + s];
            if (c < 7) a.frames_out[((size_t)b * kN + i) * 7 + c] = a.frames_in[((size_t)b * kN + i) * 7 + c];
            else a.tors_out[((size_t)b * kN + i) * 14 + (c - 7)] = a.tors_in[((size_t)b * kN + i) * 14 + (c - 7)];
        }

        // ---------------- attention-carrying pairs, in groups of whole rows ----------------
        const int rows_per_group = W > 0 ? max(1, kCapPairs / W) : kN;
        for (int row0 = 0; row0 < L; row0 += rows_per_group) {  // (L == 0: nothing to do)
            const int nrows = min(rows_per_group, L - row0);
            const int gpairs = nrows * W;
            for (int pass_base = 0; pass_base < gpairs; pass_base += kPassPairs) {
                const int npass = min(kPassPairs, gpairs - pass_base);
                if (warp8 * 32 < npass) {  // warp pairs past the end of a short last pass have nothing to do
                    const bool act = pcol < npass;
                    const int gp = act ? pass_base + pcol : pass_base;
                    const PairRef pr = decode_full_pair(I, gp, W, L, row0, act);
                    message_stage(S, M, ajt, a.Kpad, pr, pcol, half, warp8);
                    heads_stage(S, M, pr, pcol, gp, half, warp8, lsave, a.Kpad);
                }
                if (LAYER == 0) {
                    __syncthreads();
                    accumulate_msum(S, M, I, row0, nrows, W, pass_base, npass, -1, 1.0f);
                    __syncthreads();
                }
            }
            __syncthreads();

            finalize_rows(S, M, a, I, b, row0, nrows, W);
            __syncthreads();
        }

        if (LAYER == 0) {
            // ---------------- message-only pairs: self, masked peptide slots, masked pocket slots ----------------
            const int npx = kN - L;                          // masked peptide slots
            const int W2 = 1 + npx + ci.nx + (ci.c0 > 0 ? 1 : 0);
            const int total = L * W2;
            for (int pass_base = 0; pass_base < total; pass_base += kPassPairs) {
                const int npass = min(kPassPairs, total - pass_base);
                if (warp8 * 32 < npass) {
                    PairRef pr;
                    const bool act = pcol < npass;
                    const int gp = act ? pass_base + pcol : pass_base;
                    const int rl = gp / W2, e = gp - rl * W2;
                    pr.i = I[IN_ROWS + rl];
                    pr.active = act;
                    if (e == 0) pr.j = pr.i;
                    else if (e <= npx) pr.j = I[IN_PEPX + e - 1];
                    else if (e <= npx + ci.nx) pr.j = I[IN_POCKET + a.Kpad - 1 - (e - npx - 1)];
                    else pr.j = -1;
                    message_stage(S, M, ajt, a.Kpad, pr, pcol, half, warp8);
                }
                __syncthreads();
                accumulate_msum(S, M, I, 0, L, W2, pass_base, npass, ci.c0 > 0 ? W2 - 1 : -1, (float)ci.c0);
                __syncthreads();
            }

            // ---------------- node feature update: relu(feature_mlp(cat(h_i, sum_j m_ij))) (model.py:151, :407) ----------------
            const float* f0w = a.params + param_offset(0, FEAT0_W);
            const float* f0b = a.params + param_offset(0, FEAT0_B);
            const float* f2w = a.params + param_offset(0, FEAT2_W);
            const float* f2b = a.params + param_offset(0, FEAT2_B);
            constexpr int ldf = kH1 + kHid;
            float* hid = S + M.Scr;                       // [16][65]
            float* sf0 = hid + kN * kLdN;                 // feature_mlp.0.weight [64][87], staged (pass buffers are free)
            float* sf2 = sf0 + kHid * ldf;                // feature_mlp.2.weight [64][65]
            for (int idx = tid; idx < kHid * ldf; idx += kFwdThreads) sf0[idx] = f0w[idx];
            for (int idx = tid; idx < kHid * kHid; idx += kFwdThreads) sf2[(idx >> 6) * kLdN + (idx & 63)] = f2w[idx];
            __syncthreads();
            for (int idx = tid; idx < L * kHid; idx += kFwdThreads) {
                int r = idx >> 6, n = idx & 63;
                int i = I[IN_ROWS + r];
                const float* w = sf0 + n * ldf;
                const float* h = S + M.H + i * kLdN;
                const float* ms = S + M.Msum + i * kHid;
                float acc = f0b[n];
#pragma unroll
                for (int c = 0; c < kH1; ++c) acc = fmaf(w[c], h[c], acc);
#pragma unroll 8
                for (int c = 0; c < kHid; ++c) acc = fmaf(w[kH1 + c], ms[c], acc);
                hid[r * kLdN + n] = fmaxf(acc, 0.0f);
                if (a.msum_out != nullptr) a.msum_out[((size_t)b * kN + i) * kHid + n] = ms[n];
            }
            __syncthreads();
            for (int idx = tid; idx < kN * kHid; idx += kFwdThreads) {
                int s = idx >> 6, n = idx & 63;
                float v = 0.0f;
                int i;
                if (s < L) {
                    i = I[IN_ROWS + s];
                    const float* w = sf2 + n * kLdN;
                    float acc = f2b[n];
#pragma unroll 8
                    for (int c = 0; c < kHid; ++c) acc = fmaf(w[c], hid[s * kLdN + c], acc);
                    v = fmaxf(acc, 0.0f);
                } else {
                    i = I[IN_PEPX + s - L];
                    if (a.msum_out != nullptr) a.msum_out[((size_t)b * kN + i) * kHid + n] = 0.0f;
                }
                a.feat_out[((size_t)b * kN + i) * kHid + n] = v;
            }
        }
        __syncthreads();
    }
}

}  // namespace pmhc

using namespace pmhc;

namespace pmhc {

static int g_num_sms = 0;
static int g_max_smem = 0;

int device_props() {
    if (g_num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return -1;
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    return g_num_sms > 0 ? 0 : -1;
}
int num_sms() { return g_num_sms; }
int pad_k(int P) { return ((kN + P + 31) / 32) * 32; }

// workspace layout: [ajt scratch: num_sms * 64 * Kpad] [frames1: B*16*7] [tors1: B*16*14] [feat1: B*16*64]
struct Workspace {
    float *ajt, *frames1, *tors1, *feat1;
    void* tc2;          // buffers of the second-generation tensor-core path (carve_tc2)
    void* tc3;          // buffers of the fp16-term tensor-core path (carve_tc3)
    size_t bytes;
};
size_t tc2_workspace_bytes(int B, int P);
size_t tc3_workspace_bytes(int B, int P);
Workspace carve_workspace(void* base, int B, int P) {
    Workspace w;
    size_t o = 0;
    float* p = (float*)base;
    size_t n_ajt = (size_t)(g_num_sms > 0 ? g_num_sms : 148) * kHid * pad_k(P);
    w.ajt = p + o;     o += n_ajt;
    w.frames1 = p + o; o += (size_t)B * kN * 7;
    w.tors1 = p + o;   o += (size_t)B * kN * 14;
    w.feat1 = p + o;   o += (size_t)B * kN * kHid;
    o = (o + 63) & ~(size_t)63;   // 256-byte aligned
    w.tc2 = p + o;     o += (tc2_workspace_bytes(B, P) + 3) / 4;
    o = (o + 63) & ~(size_t)63;
    w.tc3 = p + o;     o += (tc3_workspace_bytes(B, P) + 3) / 4;
    w.bytes = o * sizeof(float);
    return w;
}

template <int LAYER>
int launch_layer_forward(const LayerArgs& a, cudaStream_t stream) {
    static PerDeviceOnce configured;
    const SmemMap M = make_smem_map(a.Kpad);
    size_t smem = (size_t)M.total_floats * sizeof(float);
    PMHC_REQUIRE((int)smem <= g_max_smem, "EGNN layer needs %zu B of shared memory (P=%d), device allows %d", smem, a.P, g_max_smem);
    if (configured.needed()) {
        cudaError_t e = cudaFuncSetAttribute(egnn_layer_forward_kernel<LAYER>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem);
        PMHC_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(forward): %s", cudaGetErrorString(e));
        configured.mark();
    }
    int grid = a.B < g_num_sms ? a.B : g_num_sms;
    if (profile_enabled()) profile_mark(PROF_FWD, stream, true);
    egnn_layer_forward_kernel<LAYER><<<grid, kFwdThreads, smem, stream>>>(a);
    if (profile_enabled()) profile_mark(PROF_FWD, stream, false);
    PMHC_CHECK_LAUNCH("egnn_layer_forward");
    return 0;
}

}  // namespace pmhc

namespace pmhc {
size_t forward_workspace_bytes(int B, int P) {
    device_props();
    return carve_workspace(nullptr, B, P).bytes;
}
}  // namespace pmhc

// saved-for-backward layout (floats): rowstat1 [B,16,16] | rowstat2 [B,16,16] | frames1 [B,16,7] | tors1 [B,16,14]
// | feat1 [B,16,64] | msum1 [B,16,64] | logits1 [B,16,Kpad] | logits2 [B,16,Kpad]
extern "C" size_t pmhc_saved_floats(int B, int P) {
    return (size_t)B * kN * (2 * PMHC_ROWSTAT + 7 + 14 + kHid + kHid + 2 * pad_k(P));
}

namespace pmhc {
SavedMap carve_saved(float* saved, int B, int P) {
    SavedMap s;
    const size_t BN = (size_t)B * kN;
    s.rowstat1 = saved;
    s.rowstat2 = s.rowstat1 + BN * PMHC_ROWSTAT;
    s.frames1 = s.rowstat2 + BN * PMHC_ROWSTAT;
    s.tors1 = s.frames1 + BN * 7;
    s.feat1 = s.tors1 + BN * 14;
    s.msum1 = s.feat1 + BN * kHid;
    s.logits1 = s.msum1 + BN * kHid;
    s.logits2 = s.logits1 + BN * pad_k(P);
    return s;
}
}  // namespace pmhc

namespace pmhc {
int forward_tc2(const float* params, const PmhcBatch* bt, float t_over_T, float* frames1, float* tors1, float* out_frames,
                float* out_torsions, float* feat1_out, float* msum_out, float* rowstat1, float* rowstat2, float* logits1,
                float* logits2, void* tc2_ws, cudaStream_t stream, bool reuse_pocket_cache);
int forward_tc3(int terms, const float* params, const PmhcBatch* bt, float t_over_T, float* frames1, float* tors1, float* out_frames,
                float* out_torsions, float* feat1_out, float* msum_out, float* rowstat1, float* rowstat2, float* logits1, float* logits2,
                void* ws, cudaStream_t stream, bool reuse_pocket_cache);

int model_forward_impl(const float* params, const PmhcBatch* bt, float t_over_T, float* out_frames, float* out_torsions,
                       float* saved, void* workspace, size_t workspace_bytes, cudaStream_t stream, int precision,
                       bool reuse_pocket_cache);
}

extern "C" int pmhc_model_forward(const float* params, const PmhcBatch* bt, float t_over_T, float* out_frames,
                                  float* out_torsions, float* saved, void* workspace, size_t workspace_bytes,
                                  void* stream_) {
    return pmhc_model_forward_ex(params, bt, t_over_T, out_frames, out_torsions, saved, workspace, workspace_bytes,
                                 stream_, PMHC_PRECISION_FP32);
}

extern "C" int pmhc_model_forward_ex(const float* params, const PmhcBatch* bt, float t_over_T, float* out_frames,
                                     float* out_torsions, float* saved, void* workspace, size_t workspace_bytes,
                                     void* stream_, int precision) {
    return model_forward_impl(params, bt, t_over_T, out_frames, out_torsions, saved, workspace, workspace_bytes,
                              (cudaStream_t)stream_, precision, false);
}

// reuse_pocket_cache: the pocket projection cache in `workspace` is still valid (same batch, same weights) — the
// sampling trajectory sets it from its second step on.
int pmhc::model_forward_impl(const float* params, const PmhcBatch* bt, float t_over_T, float* out_frames,
                             float* out_torsions, float* saved, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream, int precision, bool reuse_pocket_cache) {
    PMHC_REQUIRE(precision >= PMHC_PRECISION_FP32 && precision <= PMHC_PRECISION_FP16, "unknown precision mode %d", precision);
    const bool use_tc = precision == PMHC_PRECISION_BF16;
    PMHC_REQUIRE(device_props() == 0, "no CUDA device");
    PMHC_REQUIRE(bt != nullptr && bt->B > 0, "pmhc_model_forward: empty batch");
    PMHC_REQUIRE(bt->P >= 1 && bt->P <= kMaxP, "pmhc_model_forward: pocket_maxlen %d outside [1, %d]", bt->P, kMaxP);
    Workspace w = carve_workspace(workspace, bt->B, bt->P);
    PMHC_REQUIRE(workspace != nullptr && workspace_bytes >= w.bytes, "pmhc_model_forward: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    SavedMap sv{};
    if (saved != nullptr) sv = carve_saved(saved, bt->B, bt->P);
    float* rowstat1 = sv.rowstat1;
    float* rowstat2 = sv.rowstat2;
    float* frames1 = saved ? sv.frames1 : w.frames1;
    float* tors1 = saved ? sv.tors1 : w.tors1;
    float* feat1 = saved ? sv.feat1 : w.feat1;
    float* msum1 = sv.msum1;

    if (precision == PMHC_PRECISION_TC32 || precision == PMHC_PRECISION_FP16)
        return forward_tc3(precision == PMHC_PRECISION_TC32 ? 2 : 1, params, bt, t_over_T, frames1, tors1, out_frames, out_torsions,
                           saved ? feat1 : nullptr, msum1, rowstat1, rowstat2, sv.logits1, sv.logits2, w.tc3, stream, reuse_pocket_cache);
    if (use_tc)
        return forward_tc2(params, bt, t_over_T, frames1, tors1, out_frames, out_torsions, saved ? feat1 : nullptr, msum1,
                           rowstat1, rowstat2, sv.logits1, sv.logits2, w.tc2, stream, reuse_pocket_cache);

    LayerArgs a;
    a.params = params;
    a.B = bt->B; a.P = bt->P; a.Kpad = pad_k(bt->P);
    a.t_over_T = t_over_T;
    a.t_dev = step_t_dev();
    a.frames_in = bt->frames; a.tors_in = bt->torsions; a.feat_in = bt->features; a.mask = bt->mask;
    a.pocket_frames = bt->pocket_frames; a.pocket_feat = bt->pocket_features; a.pocket_mask = bt->pocket_mask;
    a.frames_out = frames1; a.tors_out = tors1; a.feat_out = feat1; a.msum_out = msum1; a.rowstat = rowstat1;
    a.logit_out = sv.logits1;
    a.ajt_ws = w.ajt;
    a.ajt_cache = nullptr;
    a.pocket_cls = nullptr;
    int rc = launch_layer_forward<0>(a, stream);
    if (rc != 0) return rc;
    a.frames_in = frames1; a.tors_in = tors1; a.feat_in = feat1;
    a.frames_out = out_frames; a.tors_out = out_torsions; a.feat_out = nullptr; a.msum_out = nullptr; a.rowstat = rowstat2;
    a.logit_out = sv.logits2;
    return launch_layer_forward<1>(a, stream);
}
