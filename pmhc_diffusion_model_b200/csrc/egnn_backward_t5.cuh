// egnn_backward_t5.cuh — the EGNN layer backward on the Blackwell tensor-core path (tcgen05.mma, accumulators in tensor memory).
// Included by egnn_backward.cu (same skeleton: one persistent CTA per SM, 128-pair passes, per-CTA gradient partials, the shared
// per-complex prologue and node level).  It is the autograd of EGNNLayer.forward (diffusion/model.py:83-333) under
// `total_loss.mean().backward()` (diffusion/optimizer.py:222), arithmetic mode `PMHC_PRECISION_FP16` of pmhc_model_backward_ex.
//
// What differs from the warp-level (mma.sync) kernel:
//   * message_mlp.2 is folded into the head weights for the whole pass: F_h = W_h[:, :64] W2 and c_h = W_h[:, :64] b2 + b_h are formed
//     once per step (bwd_fold_weights_kernel), so a pass contracts m1 = relu(A_i + A_j + W_e) directly:
//         hid_h = relu(m1 F_h^T + c_h + extras),   dm1 = (sum_h dpre_h F_h) .* relu'(m1),   dF_h = sum_p dpre_h[p] (x) m1[p];
//     the message and its gradient are never formed.  The chain rule at the weight level (bwd_unfold_kernel, after the fixed-order
//     reduction of the per-CTA partials) gives dW_h = dF_h W2^T + db_h (x) b2, dW2 = sum_h W_h^T dF_h, db2 = sum_h W_h^T db_h;
//     layer 1's message-sum path (model.py:151) enters per node: dm1 += W2^T dMsum[i], dW2 += dMsum[i] (x) sum_j m1[i, j].
//   * operands are fp16 tiles in shared memory (128-byte swizzle, one tile = 128 pairs x 64 features = 16 KB) that serve BOTH as
//     K-major operands (hid = m1 F^T, dm1 = dpre F) and as MN-major operands (the weight-gradient sums contract over the PAIRS of a
//     pass: dF = dpre^T m1, dW3 = hid^T dout, extras / bias columns = dpre^T [lq, -d2, qdot2, 1]) — checked on B200 by
//     profiles/probes/mixed16_probe.cu.  (kind::tf32 cannot do that: its MN-major form needs a different swizzle than its K-major
//     form, and kind::f16 rejects mixed fp16 x bf16 operands — profiles/probes/tf32_probe.cu.)  Gradient-like operands (dout, dpre)
//     are scaled by a power of two s = 16 * 2^-floor(log2 max|upstream gradient|) so they sit in fp16's normal range; accumulators
//     are read back with 1 / s.
//   * every weight-gradient sum of the CTA stays in tensor memory for the whole launch (dF: two M = 128 accumulators over head pairs
//     (rotation, torsion) and (translation, attention), dW3 and the extras / bias columns: 16 columns each) and is written to the
//     CTA's partial once, at the end; a pass needs three MMA batches (hidden layers of all four heads; pair A; pair B) plus a small one
//     for the per-row sums, issued by a dedicated warp, and the 256 compute threads work as two threads per pair, one HEAD each
//     (rotation | torsion, then translation | attention), walking the 64 hidden units of their TMEM lane in chunks of 16.
//   * passes are COLUMN-major and dealt to the CTAs by pass (bwd_schedule_kernel); everything of a complex that is per node rather
//     than per pair runs in its own kernels: bwd_setup_pre_kernel, bwd_feature_pre_kernel (layer 1), bwd_node_post_kernel.
#pragma once

#include <cuda_fp16.h>

#include "tcgen05.cuh"

namespace pmhc {

constexpr int kT5Threads = 288;   // 8 compute warps (two threads per pair) + 1 MMA-issuing warp
constexpr int kT5Compute = 256;
// operand tiles: byte offsets from the 1024-byte aligned base of dynamic shared memory
constexpr int T5_M1 = 0, T5_M1B = 16384, T5_F = 32768, T5_HID = 65536, T5_DPRE = 98304, T5_DX = 131072, T5_TILE_BYTES = 147456;
// (two m1 tiles: pass p uses the one of parity p & 1, so the next pass's tile is staged under this pass's last MMA batches)
__host__ __device__ constexpr int t5_m1_tile(uint32_t np) { return (np & 1u) ? T5_M1B : T5_M1; }
constexpr int T5_RED = T5_HID;                       // fp32 [64][132] tile of dm1 over the (then free) hid / dpre tiles
constexpr int T5_DM1H = T5_DPRE + 16384;             // the same as an fp16 operand tile (per-row sums on the tensor core)
constexpr int T5_RED2 = T5_HID + 64 * kLdc * 4;      // message-only passes: fp32 [64][132] tile of mult * m1
static_assert(T5_RED + 64 * kLdc * 4 <= T5_DM1H, "the fp32 and fp16 dm1 tiles must not overlap");
static_assert(T5_DM1H > T5_M1B, "the stacked [m1 | dm1] operand needs the dm1 tile behind both m1 tiles");
static_assert(T5_HID / 4 + 2 * kHid * kLdt + kHid * PMHC_NFEAT + 96 + kN * kHid <= T5_TILE_BYTES / 4, "staging views must fit behind the folded weights");
static_assert(T5_RED2 + 64 * kLdc * 4 <= T5_TILE_BYTES, "reduction tiles must fit over the hid / dpre / dx tiles");
// DX tile columns (fp16): second-layer output gradients of head pair A / B, the extra-input block shared by all heads
// DX_OUT (16): rotation 0..3, translation 4, attention 5, torsion 8..14 — pair A and pair B write the same columns (the other pair's zero);
// DX_EXT_A (16): lq 0..3, -d2 4, qdot2 5, one 6; DX_EXT_B = the window 8 columns further: zeros 0..7, -d2 8, qdot2 9, one 10; DX_SEL (16): one-hot row
constexpr int DX_OUT = 0, DX_EXT_A = 16, DX_EXT_B = 24, DX_SEL = 48;
// tensor memory columns
constexpr int TM_HID = 0, TM_DM1 = 256, TM_DFA = 320, TM_DFB = 384, TM_DW3 = 448, TM_EXT = 464, TM_SEL = 480, TM_ROW = 496;
// image of one layer's folded weights (global memory, written by bwd_fold_weights_kernel): 4 fp16 SW128 tiles [n][k] in the order
// rotation, torsion, translation, attention, then c_h as [4][64] floats
constexpr int kFoldTailFloats = 4 * kHid + 4 * 4 * kHid + 8 * kHid + 16;   // c_h, then the per-unit parameter packs and scalars as the pair kernel keeps them
constexpr int kFoldImageBytes = 4 * 8192 + kFoldTailFloats * 4;
enum { F_ROT = 0, F_TOR = 1, F_TRN = 2, F_ATT = 3 };

struct T5Map {
    BwdMap b;          // float offsets (shared prologue / node-level code reads these)
    int Cvec, Dl, B3, G, S1, Bars, Rec, total_bytes;
};

// The per-complex record the setup pre-kernel leaves in global memory (A_i with the first-layer bias, T_t with the torsion head's folded
// constant, torsions, quaternions, translations, row / neighbour lists and counts), laid out as the pair kernel keeps it in shared memory.
__host__ __device__ inline int t5_layout_record(int o, SmemMap& f, int Kpad) {
    f.Ai = o;       o += kN * kLdN;
    f.Tt = o;       o += kN * kHid;
    f.Tors = o;     o += kN * 2 * PMHC_NTORS;
    f.Q = o;        o += Kpad * 4;
    f.X = o;        o += Kpad * 3;
    f.Ints = o;     o += Kpad + 64;
    return o;
}
__host__ __device__ inline int t5_record_floats(int Kpad) { return kN * kLdN + kN * kHid + kN * 2 * PMHC_NTORS + 8 * Kpad + 64; }

// The per-complex gradient accumulators, laid out identically in the pair kernel and in the node kernel: the pair kernel writes the
// block [S1, grads_end) to global memory once per complex (kAccFloats floats), the node kernel (bwd_node_post_kernel) loads it back.
__host__ __device__ inline int t5_layout_accumulators(int o, BwdMap& m, int& S1) {
    S1 = o;         o += kN * kHid;                     // layer 1: sum over all neighbour slots of m1[i, .]
    m.dAi = o;      o += kN * kLdN;
    m.dAjPep = o;   o += kN * kLdN;
    m.dWe = o;      o += kEdge * kLdN;
    m.dTt = o;      o += kN * kHid;
    m.dMsum = o;    o += kN * kHid;
    m.RowG = o;     o += kN * 16;
    m.dQ = o;       o += kN * 4;
    m.dX = o;       o += kN * 3;
    m.dTors = o;    o += kN * 14;
    m.grads_end = o;
    return o;
}
// (dQ, dX, dTors are adjacent: a later segment of a shared complex zeroes them as one block)
constexpr int kAccFloats = kN * kHid + 2 * kN * kLdN + kEdge * kLdN + 2 * kN * kHid + kN * 16 + kN * 4 + kN * 3 + kN * 14;

struct PostMap {
    BwdMap b;
    int S1, total_bytes;
};
__host__ __device__ inline PostMap make_post_map() {
    PostMap t;
    BwdMap& m = t.b;
    int o = 0;
    m.W2 = m.Wh = m.W3i = m.Wx = m.Dx = m.Ex = m.Pl = -1;
    m.BufA = o;     o += kHid * kLdt;                   // staging: pocket dA_j^T chunk, then message_mlp.0's node columns
    m.BufB = o;     o += kHid * kLdt;                   // staging: pocket feature chunk, then torsion_mlp.0's torsion columns
    m.Dout = o;     o += kHid * PMHC_NFEAT + 96;
    m.f.H = o;      o += kN * kLdN;
    o = (o + 3) & ~3;
    m.f.Tors = o;   o += kN * 2 * PMHC_NTORS;
    o = t5_layout_accumulators(o, m, t.S1);
    m.total_floats = o;
    t.total_bytes = o * 4;
    return t;
}

__host__ __device__ inline T5Map make_t5_map(int Kpad) {
    T5Map t;
    BwdMap& m = t.b;
    int o = T5_TILE_BYTES / 4;
    m.W2 = m.Wh = m.W3i = m.Wx = m.Dx = m.Ex = -1;
    m.f.W2T = m.f.WhT = m.f.We = m.f.Out = -1;
    // staging views of the shared prologue / node level inside the tile area (free outside the passes)
    // (the folded weights at [T5_F, T5_HID) stay resident: everything transient lives in the 80 KB behind them)
    m.BufA = T5_HID / 4;
    m.BufB = m.BufA + kHid * kLdt;
    m.Dout = m.BufB + kHid * kLdt;
    m.f.Scr = m.BufA;                                   // setup_complex: <= (480 + 64) * 23 floats at the largest pocket
    m.f.Msum = m.Dout + kHid * PMHC_NFEAT + 96;         // layer 1: the saved message sums, read by the prologue only
    t.Cvec = o;     o += 4 * kHid;                      // must follow the tiles: the weight image is one bulk copy
    m.f.PkAtt = o;  o += 4 * kHid;
    m.f.PkRotQ = o; o += 4 * kHid;
    m.f.PkRot2 = o; o += 4 * kHid;
    m.f.PkMisc = o; o += 4 * kHid;
    m.f.PkTor2 = o; o += 8 * kHid;
    m.f.Scal = o;   o += 16;
    t.Dl = o;       o += 2 * kBwdPairs;                 // dL/dw shares of the rotation and translation threads
    t.B3 = o;       o += 16;                            // second-layer bias gradients of the CTA
    t.Bars = o;     o += 32;                            // mbarriers + the TMEM base address
    m.Pl = o;       o += kBwdPairs + 4;
    m.f.H = -1;
    t.G = o;        o += kN * kHid;                     // layer 1: W2^T dMsum[i]
    o = t5_layout_accumulators(o, m, t.S1);
    o = (o + 3) & ~3;
    t.Rec = o;                                          // the complex's record from bwd_setup_pre_kernel: one bulk copy
    o = t5_layout_record(o, m.f, Kpad);
    m.total_floats = o;
    m.f.total_floats = o;
    t.total_bytes = o * 4 + 1024;                       // + slack for the 1024-byte alignment of the base
    return t;
}

struct T5Args {
    const uint8_t* wimg;   // this layer's folded-weight image (kFoldImageBytes)
    const unsigned* max_bits;   // device: bit pattern of max |upstream gradient| of this launch (bwd_grad_max_kernel)
    float* acc;            // [B][kAccFloats] per-complex gradient accumulators, consumed by bwd_node_post_kernel
    float* dajt_all;       // [B][64][Kpad] per-complex dL / dA_j^T
    const float* ajt_all;  // [B][64][Kpad] neighbour projections A_j^T per complex (bwd_setup_pre_kernel)
    const float* rec_all;  // [B][t5_record_floats(Kpad)] per-complex records (bwd_setup_pre_kernel)
    const int* units;      // [B] work units of a complex = max(its attention-carrying passes, 1) (bwd_setup_pre_kernel)
    const int4* sched;     // [gridDim] {first complex, first unit inside it, units of this CTA, accumulator slot of its first segment}
    const int2* segs;      // [B] {first accumulator slot, segments} of a complex (bwd_schedule_kernel)
    const float* dmsum_g;  // layer 1: [B][2][16][64] dL / d(message sum) and W2^T of it, from bwd_feature_pre_kernel (which also zeroed the partials)
};

// ---------------------------------------------------------------------------------------------------------------------
// once per step: F_h = W_h[:, :64] W2 (fp16, SW128 tile), c_h = W_h[:, :64] b2 + b_h (the torsion head's bias lives in T_t)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bwd_fold_weights_kernel(const float* __restrict__ params, uint8_t* __restrict__ img_all,
                                                               unsigned* __restrict__ max_bits) {
    // grid: 2 layers x 4 heads x 8 blocks of 8 hidden units
    const int layer = blockIdx.x >> 5, head = (blockIdx.x >> 3) & 3, n0 = (blockIdx.x & 7) * 8;
    __shared__ float sW2[kHid * kHid];
    __shared__ float sWh[8 * kHid];
    auto off = [&](int id) { return layer == 0 ? param_offset(0, id) : param_offset(1, id); };
    int wid, bid, ld;
    if (head == F_ROT)      { wid = ROT0_W; bid = ROT0_B; ld = 68; }
    else if (head == F_TOR) { wid = TOR0_W; bid = -1;     ld = 78; }
    else if (head == F_TRN) { wid = TRN0_W; bid = TRN0_B; ld = 64; }
    else                    { wid = ATT0_W; bid = ATT0_B; ld = 66; }
    const float* W2 = params + off(MSG2_W);
    const float* b2 = params + off(MSG2_B);
    const float* Wh = params + off(wid);
    if (blockIdx.x == 0 && threadIdx.x < 8) max_bits[threadIdx.x] = 0u;      // the step's |upstream gradient| maxima start at zero
#pragma unroll 8
    for (int idx = threadIdx.x; idx < kHid * kHid; idx += 256) sW2[idx] = __ldg(W2 + idx);
    for (int idx = threadIdx.x; idx < 8 * kHid; idx += 256) sWh[idx] = __ldg(Wh + (n0 + (idx >> 6)) * ld + (idx & 63));
    __syncthreads();
    uint8_t* img = img_all + (size_t)layer * kFoldImageBytes;
    for (int idx = threadIdx.x; idx < 8 * kHid; idx += 256) {
        const int nl = idx >> 6, c = idx & 63;
        float acc = 0.0f;
#pragma unroll 16
        for (int k = 0; k < kHid; ++k) acc = fmaf(sWh[nl * kHid + k], sW2[k * kHid + c], acc);
        *reinterpret_cast<__half*>(img + head * 8192 + tc::sw128_offset(n0 + nl, c)) = __float2half_rn(acc);
    }
    float* tail = reinterpret_cast<float*>(img + 4 * 8192);
    if (threadIdx.x < 8) {
        const int n = n0 + threadIdx.x;
        float acc = bid >= 0 ? params[off(bid) + n] : 0.0f;
        for (int k = 0; k < kHid; ++k) acc = fmaf(sWh[threadIdx.x * kHid + k], __ldg(b2 + k), acc);
        tail[head * kHid + n] = acc;
    }
    if ((blockIdx.x & 31) == 0) {
        // the per-hidden-unit packs and scalars (stage_packs' layout: PkAtt, PkRotQ, PkRot2, PkMisc, PkTor2, Scal), once per layer
        float* pk = tail + 4 * kHid;
        const float* att0 = params + off(ATT0_W);
        const float* rot0 = params + off(ROT0_W);
        for (int idx = threadIdx.x; idx < 4 * kHid; idx += 256) {
            const int n = idx >> 2, c = idx & 3;
            pk[idx] = c == 0 ? att0[n * 66 + 64] : c == 1 ? att0[n * 66 + 65] : c == 2 ? params[off(ATT0_B) + n] : params[off(ATT2_W) + n];
            pk[4 * kHid + idx] = rot0[n * 68 + 64 + c];
            pk[8 * kHid + idx] = params[off(ROT2_W) + c * kHid + n];
            pk[12 * kHid + idx] = c == 0 ? params[off(TRN0_B) + n] : c == 1 ? params[off(TRN2_W) + n] : c == 2 ? params[off(ROT0_B) + n] : params[off(MSG2_B) + n];
        }
        for (int idx = threadIdx.x; idx < 8 * kHid; idx += 256) {
            const int n = idx >> 3, c = idx & 7;
            pk[16 * kHid + idx] = c < PMHC_NTORS ? params[off(TOR2_W) + c * kHid + n] : 0.0f;
        }
        if (threadIdx.x < 16) {
            const int t = threadIdx.x;
            float v = 0.0f;
            if (t == SC_ATT2B) v = params[off(ATT2_B)];
            else if (t == SC_TRN2B) v = params[off(TRN2_B)];
            else if (t >= SC_ROT2B && t < SC_ROT2B + 4) v = params[off(ROT2_B) + t - SC_ROT2B];
            else if (t >= SC_TOR2B && t < SC_TOR2B + PMHC_NTORS) v = params[off(TOR2_B) + t - SC_TOR2B];
            pk[24 * kHid + t] = v;
        }
    }
}

// max |x| over the upstream gradients of a layer launch (frames and torsions), as the bit pattern of a non-negative float (atomicMax on
// unsigned orders those like the floats).  The pair kernel turns it into s = 2^-floor(log2 max): the largest gradient-like operand
// entry then lies within a few binades of 1.
__global__ void __launch_bounds__(256) bwd_grad_max_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb,
                                                           unsigned* __restrict__ out_bits) {
    float mx = 0.0f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < na; i += gridDim.x * 256) mx = fmaxf(mx, fabsf(__ldg(a + i)));
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nb; i += gridDim.x * 256) mx = fmaxf(mx, fabsf(__ldg(b + i)));
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(mx));
}
// The largest upstream gradient is scaled into [2^-h, 2^(1-h)).  h = -4 (x16): measured, the LOW end matters more than the high one — with
// h = +6 the many small per-pair gradients fall into fp16's subnormal range (relative L2 of the flat gradient 1.1e-3 instead of 3.9e-4,
// loss-curve deviation 1.1-1.8 % instead of 0.6-0.9 %), while nothing came near fp16's 65 504 even at h = -8 (conversions saturate)
constexpr int kT5ScaleHeadroom = -4;
__device__ __forceinline__ float t5_scale_from_max(float mx) {
    float s = 1.0f;
    if (mx > 0.0f && mx < INFINITY) {
        int e;
        frexpf(mx, &e);                 // mx = f 2^e, f in [0.5, 1)
        e = min(max(-(e - 1) - kT5ScaleHeadroom, -100), 100);
        s = ldexpf(1.0f, e);            // s mx in [1, 2)
    }
    return s;
}

// After the fixed-order reduction of the per-CTA partials into `red` (parameter layout of the layer, the head first-layer message
// columns holding dF_h): the chain rule through the folded weights, added into the caller's gradient.
//   dW_h[n][k] = sum_c dF_h[n][c] W2[k][c] + db_h[n] b2[k],   dW2[k][c] += sum_h sum_n W_h[n][k] dF_h[n][c],   db2[k] += sum_h sum_n W_h[n][k] db_h[n]
// Grid: ceil(numel / 256) pass-through CTAs (every other parameter: grad += red), then 16 CTAs for the head matrices (head x 16 rows),
// 16 for message_mlp.2.weight (4 rows each) and one for its bias; the products read their operands from shared memory.
template <int LAYER>
__global__ void __launch_bounds__(256) bwd_unfold_kernel(const float* __restrict__ params, const float* __restrict__ red, float* __restrict__ grad) {
    constexpr int base = param_offset(LAYER, 0);
    constexpr int numel = param_offset(LAYER + 1, 0) - base;
    constexpr int n_pass = (numel + 255) / 256;
    constexpr int wid[4] = {ROT0_W, TOR0_W, TRN0_W, ATT0_W};
    constexpr int bid[4] = {ROT0_B, TOR0_B, TRN0_B, ATT0_B};
    constexpr int lds[4] = {68, 78, 64, 66};
    constexpr int oW2 = param_offset(LAYER, MSG2_W), oB2 = param_offset(LAYER, MSG2_B);
    __shared__ float sA[kHid * (kHid + 1)];
    __shared__ float sB[kHid * (kHid + 1)];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < n_pass) {
        const int p = blockIdx.x * 256 + tid;
        if (p >= numel) return;
        const int q = p + base;
        bool folded = (q >= oW2 && q < oW2 + kHid * kHid) || (q >= oB2 && q < oB2 + kHid);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int o = param_offset(LAYER, wid[h]);
            if (q >= o && q < o + kHid * lds[h] && (q - o) % lds[h] < kHid) folded = true;
        }
        if (!folded) grad[q] += red[p];
        return;
    }
    const int job = blockIdx.x - n_pass;
    if (job < 16) {
        // head h, rows n0 .. n0 + 15: sA = W2 [k][c] (padded), sB = dF_h rows
        const int h = job >> 2, n0 = (job & 3) * 16;
        const int ow = h == 0 ? param_offset(LAYER, ROT0_W) : h == 1 ? param_offset(LAYER, TOR0_W) : h == 2 ? param_offset(LAYER, TRN0_W) : param_offset(LAYER, ATT0_W);
        const int ob = h == 0 ? param_offset(LAYER, ROT0_B) : h == 1 ? param_offset(LAYER, TOR0_B) : h == 2 ? param_offset(LAYER, TRN0_B) : param_offset(LAYER, ATT0_B);
        const int ld = h == 0 ? 68 : h == 1 ? 78 : h == 2 ? 64 : 66;
#pragma unroll 4
        for (int idx = tid; idx < kHid * kHid; idx += 256) sA[(idx >> 6) * (kHid + 1) + (idx & 63)] = __ldg(params + oW2 + idx);
        for (int idx = tid; idx < 16 * kHid; idx += 256) sB[(idx >> 6) * (kHid + 1) + (idx & 63)] = __ldcg(red + (ow - base) + (n0 + (idx >> 6)) * ld + (idx & 63));
        __syncthreads();
        for (int idx = tid; idx < 16 * kHid; idx += 256) {
            const int nl = idx >> 6, k = idx & 63;
            float acc = __ldcg(red + (ob - base) + n0 + nl) * __ldg(params + oB2 + k);
#pragma unroll 16
            for (int c = 0; c < kHid; ++c) acc = fmaf(sB[nl * (kHid + 1) + c], sA[k * (kHid + 1) + c], acc);
            grad[ow + (n0 + nl) * ld + k] += acc;
        }
    } else if (job < 32) {
        // message_mlp.2.weight rows k0 .. k0 + 3: per head, sA = dF_h [n][c], sB = W_h[n][k0 .. k0 + 3]
        const int k0 = (job - 16) * 4;
        const int kk = tid >> 6, c = tid & 63;
        float acc = __ldcg(red + (oW2 - base) + (k0 + kk) * kHid + c);
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
            const int ow = h == 0 ? param_offset(LAYER, ROT0_W) : h == 1 ? param_offset(LAYER, TOR0_W) : h == 2 ? param_offset(LAYER, TRN0_W) : param_offset(LAYER, ATT0_W);
            const int ld = h == 0 ? 68 : h == 1 ? 78 : h == 2 ? 64 : 66;
            __syncthreads();
#pragma unroll 4
            for (int idx = tid; idx < kHid * kHid; idx += 256) sA[idx] = __ldcg(red + (ow - base) + (idx >> 6) * ld + (idx & 63));
            sB[tid] = __ldg(params + ow + (tid >> 2) * ld + k0 + (tid & 3));
            __syncthreads();
            float part = 0.0f;
#pragma unroll 16
            for (int n = 0; n < kHid; ++n) part = fmaf(sB[n * 4 + kk], sA[n * kHid + c], part);
            acc += part;
        }
        grad[oW2 + (k0 + kk) * kHid + c] += acc;
    } else {
        // message_mlp.2.bias: thread (h, k) forms sum_n W_h[n][k] db_h[n], the four heads are added in order
        const int h = tid >> 6, k = tid & 63;
        const int ow = h == 0 ? param_offset(LAYER, ROT0_W) : h == 1 ? param_offset(LAYER, TOR0_W) : h == 2 ? param_offset(LAYER, TRN0_W) : param_offset(LAYER, ATT0_W);
        const int ob = h == 0 ? param_offset(LAYER, ROT0_B) : h == 1 ? param_offset(LAYER, TOR0_B) : h == 2 ? param_offset(LAYER, TRN0_B) : param_offset(LAYER, ATT0_B);
        const int ld = h == 0 ? 68 : h == 1 ? 78 : h == 2 ? 64 : 66;
        float part = 0.0f;
#pragma unroll 16
        for (int n = 0; n < kHid; ++n) part = fmaf(__ldg(params + ow + n * ld + k), __ldcg(red + (ob - base) + n), part);
        sA[tid] = part;
        __syncthreads();
        if (tid < kHid) grad[oB2 + tid] += __ldcg(red + (oB2 - base) + tid) + (((sA[tid] + sA[64 + tid]) + sA[128 + tid]) + sA[192 + tid]);
    }
}

// partial -> red (plain store).  blockDim = (64, 4): the CTA rows are split in four contiguous groups summed by four threads of a
// parameter, each in ascending CTA order, and the four sums are added in group order — a fixed order, independent of timing.
__global__ void __launch_bounds__(256) reduce_partials_to_kernel(const float* __restrict__ partial, int stride, int n_cta, int numel,
                                                                 float* __restrict__ red) {
    __shared__ float part[4][64];
    const int p = blockIdx.x * 64 + threadIdx.x;
    const int per = (n_cta + 3) / 4;
    const int c_lo = threadIdx.y * per, c_hi = min(n_cta, c_lo + per);
    float acc = 0.0f;
    if (p < numel) {
        for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = c0 + u < c_hi ? __ldcg(partial + (size_t)(c0 + u) * stride + kTileFloats + p) : 0.0f;
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (c0 + u < c_hi) acc += v[u];
        }
    }
    part[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && p < numel) red[p] = ((part[0][threadIdx.x] + part[1][threadIdx.x]) + part[2][threadIdx.x]) + part[3][threadIdx.x];
}

// ---------------------------------------------------------------------------------------------------------------------
// device helpers of the pass
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void t5_bar_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// one 16-byte chunk (8 fp16 columns starting at column 8 * chunk) of row p
__device__ __forceinline__ void t5_store_chunk(uint8_t* tile, int p, int chunk, float v0, float v1, float v2, float v3, float v4, float v5,
                                               float v6, float v7) {
    uint4 u;
    u.x = pack_h2_sat(v0, v1); u.y = pack_h2_sat(v2, v3); u.z = pack_h2_sat(v4, v5); u.w = pack_h2_sat(v6, v7);
    *reinterpret_cast<uint4*>(tile + (p >> 3) * 1024 + (p & 7) * 128 + ((chunk ^ (p & 7)) << 4)) = u;
}
// ---- the MMA-issuing warp: the three batches of one attention-carrying pass (np = running pass count of the CTA, q = pass of the complex) ----
__device__ __forceinline__ void t5_issue_pass(uint32_t sbase, uint32_t tmem, uint64_t* bars, uint32_t np, int q) {
    const uint32_t par = np & 1u;
    const uint32_t keep = np == 0 ? 0u : 1u;      // accumulate flag of the CTA-resident weight-gradient sums at their first step
    const uint32_t keep_c = q == 0 ? 0u : 1u;     // ... of the complex-resident per-row sums
    uint64_t* rdy = bars;        // [0..3]
    uint64_t* done = bars + 4;   // [0..3]
    constexpr uint32_t id_hid = tc::idesc_f16_f32(128, 128);
    constexpr uint32_t id_w16 = tc::idesc_f16_f32_major(128, 16, 1, 1);
    constexpr uint32_t id_w64 = tc::idesc_f16_f32_major(128, 64, 1, 1);
    constexpr uint32_t id_dm1 = tc::idesc_f16_f32_major(128, 64, 0, 1);
    // batch 0: hidden layers of the four heads (two N = 128 contractions over the folded weights)
    tc::mbar_wait_suspend(rdy + 0, par);
    tc::fence_after_thread_sync();
    if (tc::elect_one()) {
        const uint64_t da = tc::smem_desc_sw128(sbase + t5_m1_tile(np));
#pragma unroll
        for (int pair = 0; pair < 2; ++pair) {
            const uint64_t db = tc::smem_desc_sw128(sbase + T5_F + pair * 16384);
#pragma unroll
            for (int s = 0; s < 4; ++s) tc::mma_bf16(tmem + TM_HID + 128 * pair, da + 2 * s, db + 2 * s, id_hid, s > 0);
        }
        tc::mma_commit(done + 0);
    }
    __syncwarp();
    // batches 1, 2: head pairs (rotation, torsion) and (translation, attention)
#pragma unroll 1
    for (int pair = 0; pair < 2; ++pair) {
        tc::mbar_wait_suspend(rdy + 1 + pair, par);
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
            // dm1 += dpre_h F_h first (the compute threads wait for it): A K-major, B = the folded weight tile read MN-major
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint64_t da = tc::smem_desc_sw128(sbase + T5_DPRE + h * 16384);
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    tc::mma_bf16(tmem + TM_DM1, da + 2 * s, tc::smem_desc(sbase + T5_F + (2 * pair + h) * 8192 + s * 2048, 8192, 1024, 2), id_dm1,
                                 (pair | h | s) > 0);
            }
            // weight-gradient sums over the pairs of the pass (K = 128 pairs = 8 steps of 16 rows), A = the two heads' tiles stacked in M
#pragma unroll
            for (int s = 0; s < 8; ++s)
                tc::mma_bf16(tmem + (pair == 0 ? TM_DFA : TM_DFB), tc::smem_desc(sbase + T5_DPRE + s * 2048, 16384, 1024, 2),
                             tc::smem_desc(sbase + t5_m1_tile(np) + s * 2048, 16384, 1024, 2), id_w64, s > 0 ? 1u : keep);
#pragma unroll
            for (int s = 0; s < 8; ++s)
                tc::mma_bf16(tmem + TM_DW3, tc::smem_desc(sbase + T5_HID + s * 2048, 16384, 1024, 2),
                             tc::smem_desc(sbase + T5_DX + DX_OUT * 2 + s * 2048, 16384, 1024, 2), id_w16, (s > 0 || pair > 0) ? 1u : keep);
#pragma unroll
            for (int s = 0; s < 8; ++s)
                tc::mma_bf16(tmem + TM_EXT, tc::smem_desc(sbase + T5_DPRE + s * 2048, 16384, 1024, 2),
                             tc::smem_desc(sbase + T5_DX + (pair == 0 ? DX_EXT_A : DX_EXT_B) * 2 + s * 2048, 16384, 1024, 2), id_w16,
                             (s > 0 || pair > 0) ? 1u : keep);
            if (pair == 0) {
                // per-row sums of the torsion head's hidden-layer gradient (rows 64.. of the stacked tiles) against the row selector
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    tc::mma_bf16(tmem + TM_SEL, tc::smem_desc(sbase + T5_DPRE + s * 2048, 16384, 1024, 2),
                                 tc::smem_desc(sbase + T5_DX + DX_SEL * 2 + s * 2048, 16384, 1024, 2), id_w16, s > 0 ? 1u : keep_c);
            }
            tc::mma_commit(done + 1 + pair);
        }
        __syncwarp();
    }
    // batch 3: per-row sums of m1 (rows 0..63: layer 1's message-sum path) and of dm1 (rows 64..127: dL / dA_i) against the row selector
    tc::mbar_wait_suspend(rdy + 3, par);
    tc::fence_after_thread_sync();
    if (tc::elect_one()) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
            tc::mma_bf16(tmem + TM_ROW, tc::smem_desc(sbase + t5_m1_tile(np) + s * 2048, (uint32_t)(T5_DM1H - t5_m1_tile(np)), 1024, 2),
                         tc::smem_desc(sbase + T5_DX + DX_SEL * 2 + s * 2048, 16384, 1024, 2), id_w16, s > 0 ? 1u : keep_c);
        tc::mma_commit(done + 3);
    }
    __syncwarp();
}

#ifdef PMHC_T5_STAMPS
__device__ long long t5_dbg[2][16];
#define T5_PSTAMP(k) do { if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 128)) { long long now_ = clock64(); t5_dbg[threadIdx.x >> 7][k] += now_ - pst_; pst_ = now_; } } while (0)
#else
#define T5_PSTAMP(k) do { } while (0)
#endif

// 16 accumulator columns of the thread's TMEM lane, split in issue and wait so that the next chunk's load runs under the current
// chunk's arithmetic; the wait takes the registers as in/out operands: no use of them can be scheduled above it
__device__ __forceinline__ void t5_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void t5_ld16_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
// chunk ch of a head's 64 accumulator columns -> hv, with chunk ch + 1 requested behind it (nx carries the request between calls)
__device__ __forceinline__ void t5_next16(uint32_t tbase, int ch, uint32_t (&nx)[16], float (&hv)[16]) {
    t5_ld16_wait(nx);
#pragma unroll
    for (int e = 0; e < 16; ++e) hv[e] = __uint_as_float(nx[e]);
    if (ch < 3) t5_ld16_issue(tbase + 16 * (ch + 1), nx);
}
// 16 floats (scaled) -> two 16-byte chunks (2 ch, 2 ch + 1) of row p of a SW128 fp16 tile
__device__ __forceinline__ void t5_store16(uint8_t* tile, int p, int ch, const float (&v)[16], float scale) {
    uint8_t* row = tile + (p >> 3) * 1024 + (p & 7) * 128;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        uint4 u;
        u.x = pack_h2_sat(v[8 * q + 0] * scale, v[8 * q + 1] * scale);
        u.y = pack_h2_sat(v[8 * q + 2] * scale, v[8 * q + 3] * scale);
        u.z = pack_h2_sat(v[8 * q + 4] * scale, v[8 * q + 5] * scale);
        u.w = pack_h2_sat(v[8 * q + 6] * scale, v[8 * q + 7] * scale);
        *reinterpret_cast<uint4*>(row + (((2 * ch + q) ^ (p & 7)) << 4)) = u;
    }
}

// the A_j^T column half (32 features) and the saved attention logit of a pair: 33 L2 loads in flight
__device__ __forceinline__ void t5_request_pair(float (&aj)[32], float& logit, const BwdArgs& g, int b, const float* __restrict__ ajt, int Kpad,
                                                const PairRef pr, int half) {
    const float* ajc = ajt + (pr.j >= 0 ? pr.j : 0) + (size_t)(32 * half) * Kpad;
#pragma unroll
    for (int k = 0; k < 32; ++k) aj[k] = __ldcg(ajc + k * Kpad);
    logit = __ldcg(g.logits + ((size_t)b * kN + pr.i) * Kpad + pr.j);
}

// m1 = relu(A_i + A_j + W_e) of pair `pr` (this thread's half of the features) -> fp16 row of the m1 tile of pass np; returns the relu mask
template <int LAYER>
__device__ __forceinline__ uint32_t t5_stage_m1(uint8_t* sb, const float* S, const BwdMap& M, const LayerArgs& a, const PairRef pr, int p, int half,
                                                uint32_t np, float (&aj)[32]) {
    constexpr int H = layer_H(LAYER);
    constexpr int ld1 = 2 * H + kEdge;
    const int k0 = 32 * half;
    const float* ai = S + M.f.Ai + pr.i * kLdN + k0;
    if (pr.j >= 0 && pr.j < kN) {
        const float* we = a.params + param_offset(LAYER, MSG0_W) + 2 * H + (kN - 1 + pr.i - pr.j) + (size_t)k0 * ld1;
#pragma unroll
        for (int k = 0; k < 32; ++k) aj[k] += __ldg(we + k * ld1);
    }
    uint32_t mask = 0u;
    uint8_t* row = sb + t5_m1_tile(np) + (p >> 3) * 1024 + (p & 7) * 128;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float m[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            m[e] = fmaxf(ai[8 * q + e] + aj[8 * q + e], 0.0f);
            mask |= (m[e] > 0.0f ? 1u : 0u) << (8 * q + e);
        }
        uint4 u;
        u.x = pack_h2_sat(m[0], m[1]); u.y = pack_h2_sat(m[2], m[3]); u.z = pack_h2_sat(m[4], m[5]); u.w = pack_h2_sat(m[6], m[7]);
        *reinterpret_cast<uint4*>(row + (((4 * half + q) ^ (p & 7)) << 4)) = u;
    }
    return mask;
}

// ---- compute threads: one attention-carrying pass of up to 128 pairs ----
// Two threads per pair, one head each: threads 0..127 rotation then translation, threads 128..255 torsion then attention.  A head is
// walked in four chunks of 16 hidden units (runtime loops: the pass code stays inside the instruction cache): chunk loop 1 forms the
// hidden units and the second layer, chunk loop 2 the hidden-layer gradient; only a 64-bit relu mask is kept between them.
template <int LAYER>
__device__ __forceinline__ void t5_heads_pass(uint8_t* sb, float* S, const T5Map& T, const BwdArgs& g, const PairRef pr, int b,
                                              const float* __restrict__ ajt, float* __restrict__ dajt, const int* I, int L, int rl,
                                              int e0, int ncols, uint32_t np, uint32_t tmem, float gs, float inv_gs,
                                              float (&aj)[32], float& logit_io, const PairRef nxt, const bool has_next, uint2& m1mask_io) {
    // aj / logit_io: this pass's A_j^T column half and saved logit, loaded by the caller for the first pass of a segment; m1mask_io =
    // {relu mask of this pass's m1, staged flag}: set by the previous pass when it staged this pass's m1 tile, and set for the next here
    // Pairs of a pass are COLUMN-major: pair p = (neighbour column e0 + p / L, peptide row p % L), whole columns only.  A pocket
    // column's L pairs are then adjacent (its dL / dA_j is complete inside the pass: a plain store), and a row's pairs are L apart
    // (per-row sums go through the one-hot row selector on the tensor core; shared-memory atomics see at most 4 lanes per row).
    constexpr bool IN_GRADS = (LAYER == 1);
    const BwdMap& M = T.b;
    const LayerArgs& a = g.a;
    const int tid = threadIdx.x;
    const int p = tid & (kBwdPairs - 1);
    const int half = tid >> 7;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int i = pr.i, j = pr.j;
    const bool act = pr.active;
    const bool pep = (j >= 0 && j < kN);
    const int Kpad = a.Kpad;
    const uint32_t par = np & 1u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(S + T.Bars);
    uint64_t* rdy = bars;
    uint64_t* done = bars + 4;
    float* sDl = S + T.Dl;
    float* sB3 = S + T.B3;
    int* sPl = reinterpret_cast<int*>(S + M.Pl);
    const float* cvec = S + T.Cvec;
#ifdef PMHC_T5_STAMPS
    long long pst_ = clock64();
#endif
    // ---- m1 (my half of its features) -> fp16 tile, unless the previous pass staged it already; geometry; extras; row selector ----
    if (np > 0) tc::mbar_wait_suspend(done + 3, (np - 1) & 1u);   // the previous pass's row sums have read their m1 tile and the selector
    const bool prestaged = m1mask_io.y != 0u;
    const uint32_t m1mask = prestaged ? m1mask_io.x : t5_stage_m1<LAYER>(sb, S, M, a, pr, p, half, np, aj);
    const float* rg = S + M.RowG + i * 16;
    const float lse = rg[15], c_i = rg[14];
    const float logit = logit_io;
    const float w = act ? expf(logit - lse) : 0.0f;
    const float* pqi = S + M.f.Q + i * 4;
    const float* pqj = S + M.f.Q + j * 4;
    const Quat qi{pqi[0], pqi[1], pqi[2], pqi[3]}, qj{pqj[0], pqj[1], pqj[2], pqj[3]};
    const float rx = S[M.f.X + i * 3] - S[M.f.X + j * 3], ry = S[M.f.X + i * 3 + 1] - S[M.f.X + j * 3 + 1],
                rz = S[M.f.X + i * 3 + 2] - S[M.f.X + j * 3 + 2];
    const Quat qinvj = qinv(qj);
    const Quat v = qmul(qi, qj);
    const Quat lq = qmul(qinvj, v);
    const float d2 = rx * rx + ry * ry + rz * rz;
    const float dotq = qdot(qi, qj);
    const float qd = dotq * dotq;
    if (half == 0) {
        // extras block of pair A: lq (0..3), -d2, qdot2, 1 at its columns 0..6; pair B reads a window shifted by 8 columns, so the same
        // three inputs appear again at columns 16..18 (its columns 8..10) and its columns 0..7 are zero
        t5_store_chunk(sb + T5_DX, p, DX_EXT_A / 8, lq.w, lq.x, lq.y, lq.z, -d2, qd, 1.0f, 0.0f);
        t5_store_chunk(sb + T5_DX, p, DX_EXT_A / 8 + 1, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f);
        t5_store_chunk(sb + T5_DX, p, DX_EXT_A / 8 + 2, -d2, qd, 1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f);
    } else {
        // one-hot selector of the pair's peptide row (list position): the per-row sums of a pass are one small MMA
        const float o = act ? 1.0f : 0.0f;
        t5_store_chunk(sb + T5_DX, p, DX_SEL / 8, rl == 0 ? o : 0.0f, rl == 1 ? o : 0.0f, rl == 2 ? o : 0.0f, rl == 3 ? o : 0.0f, rl == 4 ? o : 0.0f,
                       rl == 5 ? o : 0.0f, rl == 6 ? o : 0.0f, rl == 7 ? o : 0.0f);
        t5_store_chunk(sb + T5_DX, p, DX_SEL / 8 + 1, rl == 8 ? o : 0.0f, rl == 9 ? o : 0.0f, rl == 10 ? o : 0.0f, rl == 11 ? o : 0.0f,
                       rl == 12 ? o : 0.0f, rl == 13 ? o : 0.0f, rl == 14 ? o : 0.0f, rl == 15 ? o : 0.0f);
    }
    if (tid == 0) sPl[0] = 0;      // (every reader of the previous pass's list is behind that pass's last barrier)
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    if (!prestaged) tc::mbar_arrive(rdy + 0);
    T5_PSTAMP(0);

    float gi[7] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};   // layer 2: this thread's share of dL / d (q_i, x_i)
    float dLdw_own = 0.0f;                                       // torsion thread: its share of dL / dw
    float gd_keep = 0.0f, gq_keep = 0.0f;                        // attention thread: dL / d(-d2), dL / d(qdot2)
    float hv[16];
    uint32_t nx[16];

    tc::mbar_wait_suspend(done + 0, par);
    tc::fence_after_thread_sync();
    T5_PSTAMP(1);

    // ======================= head pair A: rotation (half 0) | torsion (half 1) =======================
    if (half == 0) {
        float pre[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) pre[c] = S[M.f.Scal + SC_ROT2B + c];
        unsigned long long on = 0ull;
        t5_ld16_issue(tlane + TM_HID, nx);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            t5_next16(tlane + TM_HID, ch, nx, hv);
            uint32_t bits = 0u;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = 16 * ch + e;
                const float4 wq = *reinterpret_cast<const float4*>(S + M.f.PkRotQ + 4 * n);
                const float s = hv[e] + cvec[F_ROT * kHid + n] + wq.x * lq.w + wq.y * lq.x + wq.z * lq.y + wq.w * lq.z;
                const float h = fmaxf(s, 0.0f);
                hv[e] = h;
                bits |= (h > 0.0f ? 1u : 0u) << e;
                const float4 w2 = *reinterpret_cast<const float4*>(S + M.f.PkRot2 + 4 * n);
                pre[0] = fmaf(w2.x, h, pre[0]); pre[1] = fmaf(w2.y, h, pre[1]);
                pre[2] = fmaf(w2.z, h, pre[2]); pre[3] = fmaf(w2.w, h, pre[3]);
            }
            on |= (unsigned long long)bits << (16 * ch);
            t5_store16(sb + T5_HID, p, ch, hv, 1.0f);
        }
        const Quat dl{sigmoidf(pre[0]), sigmoidf(pre[1]), sigmoidf(pre[2]), sigmoidf(pre[3])};
        const Quat u = qmul(dl, qinvj);
        const Quat dg = qmul(qj, u);
        const Quat dG{rg[0], rg[1], rg[2], rg[3]};
        sDl[p] = qdot(dG, dg);
        const Quat ddg = qscale(dG, w);
        const Quat du = qmul_grad_b(qj, ddg);         // dg = qj * u
        const Quat ddl = qmul_grad_a(du, qinvj);      // u = dl * qinvj
        const float dp2[4] = {ddl.w * dl.w * (1.0f - dl.w), ddl.x * dl.x * (1.0f - dl.x), ddl.y * dl.y * (1.0f - dl.y),
                              ddl.z * dl.z * (1.0f - dl.z)};
        t5_store_chunk(sb + T5_DX, p, DX_OUT / 8, dp2[0] * gs, dp2[1] * gs, dp2[2] * gs, dp2[3] * gs, 0.0f, 0.0f, 0.0f, 0.0f);
        float dlq[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            const uint32_t bits = (uint32_t)(on >> (16 * ch));
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = 16 * ch + e;
                const float4 w2 = *reinterpret_cast<const float4*>(S + M.f.PkRot2 + 4 * n);
                float dp = w2.x * dp2[0] + w2.y * dp2[1] + w2.z * dp2[2] + w2.w * dp2[3];
                dp = ((bits >> e) & 1u) ? dp : 0.0f;
                hv[e] = dp;
                if (IN_GRADS) {
                    const float4 wq = *reinterpret_cast<const float4*>(S + M.f.PkRotQ + 4 * n);
                    dlq[0] = fmaf(wq.x, dp, dlq[0]); dlq[1] = fmaf(wq.y, dp, dlq[1]);
                    dlq[2] = fmaf(wq.z, dp, dlq[2]); dlq[3] = fmaf(wq.w, dp, dlq[3]);
                }
            }
            t5_store16(sb + T5_DPRE, p, ch, hv, gs);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float sum = warp_sum(dp2[c]);
            if (lane == 0) atomicAdd(sB3 + c, sum);
        }
        if (IN_GRADS && act) {
            const Quat dlqq{dlq[0], dlq[1], dlq[2], dlq[3]};
            Quat dqinv = qmul_grad_a(dlqq, v);             // lq = qinvj * v
            const Quat dv = qmul_grad_b(qinvj, dlqq);
            const Quat dqi = qmul_grad_a(dv, qj);          // v = qi * qj
            Quat dqj = qmul_grad_b(qi, dv);
            dqj = qadd(dqj, qmul_grad_a(ddg, u));          // dg = qj * u
            dqinv = qadd(dqinv, qmul_grad_b(dl, du));      // u = dl * qinvj
            dqj = qadd(dqj, qinv_grad(qj, dqinv));
            gi[0] += dqi.w; gi[1] += dqi.x; gi[2] += dqi.y; gi[3] += dqi.z;
            if (pep) atomic_add_quat(S + M.dQ + j * 4, dqj);
        }
    } else {
        float da[PMHC_NTORS];
#pragma unroll
        for (int c = 0; c < PMHC_NTORS; ++c) da[c] = S[M.f.Scal + SC_TOR2B + c];
        const float* tt = S + M.f.Tt + i * kHid;
        unsigned long long on = 0ull;
        t5_ld16_issue(tlane + TM_HID + 64, nx);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            t5_next16(tlane + TM_HID + 64, ch, nx, hv);
            uint32_t bits = 0u;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = 16 * ch + e;
                const float h = fmaxf(hv[e] + tt[n], 0.0f);
                hv[e] = h;
                bits |= (h > 0.0f ? 1u : 0u) << e;
                const float4 w0 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n);
                const float4 w1 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n + 4);
                da[0] = fmaf(w0.x, h, da[0]); da[1] = fmaf(w0.y, h, da[1]); da[2] = fmaf(w0.z, h, da[2]);
                da[3] = fmaf(w0.w, h, da[3]); da[4] = fmaf(w1.x, h, da[4]); da[5] = fmaf(w1.y, h, da[5]);
                da[6] = fmaf(w1.z, h, da[6]);
            }
            on |= (unsigned long long)bits << (16 * ch);
            t5_store16(sb + T5_HID + 16384, p, ch, hv, 1.0f);
        }
        float dda[PMHC_NTORS];
#pragma unroll
        for (int c = 0; c < PMHC_NTORS; ++c) {
            dLdw_own = fmaf(rg[4 + c], da[c], dLdw_own);
            dda[c] = w * rg[4 + c];
        }
        t5_store_chunk(sb + T5_DX, p, DX_OUT / 8 + 1, dda[0] * gs, dda[1] * gs, dda[2] * gs, dda[3] * gs, dda[4] * gs, dda[5] * gs, dda[6] * gs, 0.0f);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            const uint32_t bits = (uint32_t)(on >> (16 * ch));
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = 16 * ch + e;
                const float4 w0 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n);
                const float4 w1 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n + 4);
                const float dp = w0.x * dda[0] + w0.y * dda[1] + w0.z * dda[2] + w0.w * dda[3] + w1.x * dda[4] + w1.y * dda[5] + w1.z * dda[6];
                hv[e] = ((bits >> e) & 1u) ? dp : 0.0f;
            }
            t5_store16(sb + T5_DPRE + 16384, p, ch, hv, gs);
        }
#pragma unroll
        for (int c = 0; c < PMHC_NTORS; ++c) {
            const float sum = warp_sum(dda[c]);
            if (lane == 0) atomicAdd(sB3 + 4 + c, sum);
        }
    }
    tc::fence_proxy_async_smem();
    tc::mbar_arrive(rdy + 1);
    T5_PSTAMP(2);

    // ======================= head pair B: translation (half 0) | attention (half 1) =======================
    // chunk loop 1 touches no tile (pair A's MMAs are still reading them); loop 2 recomputes the hidden units from the accumulators
    if (half == 0) {
        float sc = S[M.f.Scal + SC_TRN2B];
        t5_ld16_issue(tlane + TM_HID + 128, nx);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            t5_next16(tlane + TM_HID + 128, ch, nx, hv);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = 16 * ch + e;
                sc = fmaf(S[M.f.PkMisc + 4 * n + 1], fmaxf(hv[e] + cvec[F_TRN * kHid + n], 0.0f), sc);
            }
        }
        const float dXr = rg[11] * rx + rg[12] * ry + rg[13] * rz;
        sDl[kBwdPairs + p] = sc * dXr;
        const float ds = w * dXr;
        if (IN_GRADS && act) {
            const float f = w * sc;
            gi[4] += f * rg[11]; gi[5] += f * rg[12]; gi[6] += f * rg[13];
            if (pep) {
                atomicAdd(S + M.dX + j * 3 + 0, -f * rg[11]); atomicAdd(S + M.dX + j * 3 + 1, -f * rg[12]); atomicAdd(S + M.dX + j * 3 + 2, -f * rg[13]);
            }
        }
        const float sum = warp_sum(ds);
        if (lane == 0) atomicAdd(sB3 + 11, sum);
        T5_PSTAMP(3);
        t5_bar_compute();                       // dL/dw shares published
        T5_PSTAMP(4);
        if (act && pep) sPl[4 + atomicAdd(sPl, 1)] = p | (j << 8) | ((kN - 1 + i - j) << 16);   // read behind a later barrier
        tc::mbar_wait_suspend(done + 1, par);   // pair A's MMAs have read the hid / dpre tiles and the second-layer block
        T5_PSTAMP(5);
        t5_ld16_issue(tlane + TM_HID + 128, nx);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            t5_next16(tlane + TM_HID + 128, ch, nx, hv);
            float dp[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = 16 * ch + e;
                const float h = fmaxf(hv[e] + cvec[F_TRN * kHid + n], 0.0f);
                hv[e] = h;
                dp[e] = h > 0.0f ? S[M.f.PkMisc + 4 * n + 1] * ds : 0.0f;
            }
            t5_store16(sb + T5_HID, p, ch, hv, 1.0f);
            t5_store16(sb + T5_DPRE, p, ch, dp, gs);
        }
        // second-layer block of pair B: rotation's columns 0..3 zero, translation at 4 (attention writes 5 of the same chunk: see below)
    } else {
        unsigned long long on = 0ull;
        t5_ld16_issue(tlane + TM_HID + 192, nx);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            t5_next16(tlane + TM_HID + 192, ch, nx, hv);
            uint32_t bits = 0u;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = 16 * ch + e;
                const float4 pk = *reinterpret_cast<const float4*>(S + M.f.PkAtt + 4 * n);
                const float h = fmaxf((cvec[F_ATT * kHid + n] + hv[e]) + fmaf(pk.y, qd, pk.x * -d2), 0.0f);
                bits |= (h > 0.0f ? 1u : 0u) << e;
            }
            on |= (unsigned long long)bits << (16 * ch);
        }
        T5_PSTAMP(3);
        t5_bar_compute();
        T5_PSTAMP(4);
        const float dLdw = sDl[p] + dLdw_own + sDl[kBwdPairs + p];
        // softmax backward with the saved row statistics; a fully saturated row has dlogit = 0 exactly (see pair_pass)
        const float dlogit = (w == 1.0f) ? 0.0f : w * (dLdw - c_i);
        tc::mbar_wait_suspend(done + 1, par);
        T5_PSTAMP(5);
        float gd = 0.0f, gq = 0.0f;
        t5_ld16_issue(tlane + TM_HID + 192, nx);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            t5_next16(tlane + TM_HID + 192, ch, nx, hv);
            const uint32_t bits = (uint32_t)(on >> (16 * ch));
            float dp[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = 16 * ch + e;
                const float4 pk = *reinterpret_cast<const float4*>(S + M.f.PkAtt + 4 * n);
                hv[e] = fmaxf((cvec[F_ATT * kHid + n] + hv[e]) + fmaf(pk.y, qd, pk.x * -d2), 0.0f);
                dp[e] = ((bits >> e) & 1u) ? pk.w * dlogit : 0.0f;
                gd = fmaf(pk.x, dp[e], gd);
                gq = fmaf(pk.y, dp[e], gq);
            }
            t5_store16(sb + T5_HID + 16384, p, ch, hv, 1.0f);
            t5_store16(sb + T5_DPRE + 16384, p, ch, dp, gs);
        }
        // second-layer block of pair B: the translation thread's ds travels through shared memory so that ONE thread writes the chunk
        {
            const float dXr = rg[11] * rx + rg[12] * ry + rg[13] * rz;
            t5_store_chunk(sb + T5_DX, p, DX_OUT / 8, 0.0f, 0.0f, 0.0f, 0.0f, w * dXr * gs, dlogit * gs, 0.0f, 0.0f);
            t5_store_chunk(sb + T5_DX, p, DX_OUT / 8 + 1, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f);
        }
        gd_keep = gd; gq_keep = gq;
        const float sum = warp_sum(dlogit);
        if (lane == 0) atomicAdd(sB3 + 12, sum);
    }
    tc::fence_proxy_async_smem();
    tc::mbar_arrive(rdy + 2);
    T5_PSTAMP(6);
    if (has_next) {
        // the next pass of this segment: its operands from L2, its m1 tile (the other buffer) and its batch-0 signal — all under
        // pair B's MMAs, so the next pass starts at its hidden-layer wait
        t5_request_pair(aj, logit_io, g, b, ajt, Kpad, nxt, half);
        m1mask_io.x = t5_stage_m1<LAYER>(sb, S, M, a, nxt, p, half, np + 1, aj);
        m1mask_io.y = 1u;
        tc::fence_proxy_async_smem();
        tc::fence_before_thread_sync();
        tc::mbar_arrive(rdy + 0);
    } else {
        m1mask_io.y = 0u;
    }
    if (IN_GRADS && act && half == 1) {
        // attention head's input gradients (under pair B's MMAs)
        const float f = -gd_keep * 2.0f;                 // d(-d2) = gd
        gi[4] += f * rx; gi[5] += f * ry; gi[6] += f * rz;
        const float fq = gq_keep * 2.0f * dotq;
        gi[0] += fq * qj.w; gi[1] += fq * qj.x; gi[2] += fq * qj.y; gi[3] += fq * qj.z;
        if (pep) {
            atomicAdd(S + M.dX + j * 3 + 0, -f * rx); atomicAdd(S + M.dX + j * 3 + 1, -f * ry); atomicAdd(S + M.dX + j * 3 + 2, -f * rz);
            atomic_add_quat(S + M.dQ + j * 4, qscale(qi, fq));
        }
    }
    if (IN_GRADS) {
        // this thread's share of dL / d (q_i, x_i): a row's pairs are L lanes apart — for L >= 8 the (at most four) lanes of a warp
        // that share a row are folded into the first of them before the shared-memory atomics
        const bool fold = L >= 8;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            float x = act ? gi[c] : 0.0f;
            if (fold) {
                const float y = __shfl_down_sync(0xffffffffu, x, L);
                if (lane + L < 32) x += y;
                const float z = __shfl_down_sync(0xffffffffu, x, (2 * L) & 31);
                if (lane + 2 * L < 32) x += z;
            }
            gi[c] = x;
        }
        if (!fold || lane < L) {
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(S + M.dQ + i * 4 + c, gi[c]);
#pragma unroll
            for (int c = 0; c < 3; ++c) atomicAdd(S + M.dX + i * 3 + c, gi[4 + c]);
        }
    }

    // ======================= dm1 = (sum_h dpre_h F_h [+ W2^T dMsum_i]) .* relu'(m1) -> reduction tile =======================
    tc::mbar_wait_suspend(done + 2, par);
    tc::fence_after_thread_sync();
    T5_PSTAMP(7);
    float* red = reinterpret_cast<float*>(sb + T5_RED);
    {
        uint32_t r[32];
        tc::tmem_ld32_nowait(tlane + TM_DM1 + 32 * half, r);
        tc::tmem_wait_ld();
        const float* gv = S + T.G + i * kHid + 32 * half;
        float x[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            float t = __uint_as_float(r[k]) * inv_gs;
            if (LAYER == 0) t += gv[k];
            x[k] = (act && ((m1mask >> k) & 1u)) ? t : 0.0f;
            red[(32 * half + k) * kLdc + p] = x[k];
        }
        uint8_t* row = sb + T5_DM1H + (p >> 3) * 1024 + (p & 7) * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_h2_sat(x[8 * q + 0] * gs, x[8 * q + 1] * gs); u.y = pack_h2_sat(x[8 * q + 2] * gs, x[8 * q + 3] * gs);
            u.z = pack_h2_sat(x[8 * q + 4] * gs, x[8 * q + 5] * gs); u.w = pack_h2_sat(x[8 * q + 6] * gs, x[8 * q + 7] * gs);
            *reinterpret_cast<uint4*>(row + (((4 * half + q) ^ (p & 7)) << 4)) = u;
        }
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    tc::mbar_arrive(rdy + 3);
    T5_PSTAMP(8);
    t5_bar_compute();
    T5_PSTAMP(9);
    {   // peptide neighbours: dA_j[j] and dW_e[rel] get the column
        const int n_pl = sPl[0];
        for (int idx = tid; idx < n_pl * kHid; idx += kT5Compute) {
            const int e = sPl[4 + (idx >> 6)], k = idx & 63;
            const float x = red[k * kLdc + (e & 255)];
            atomicAdd(S + M.dAjPep + ((e >> 8) & 255) * kLdN + k, x);
            atomicAdd(S + M.dWe + (e >> 16) * kLdN + k, x);
        }
    }
    T5_PSTAMP(10);
    {   // pocket columns: dA_j^T[k][j] = the sum over the column's L adjacent pairs — complete inside this pass, a plain store
        const int c_lo = max(e0, L - 1) - e0;          // first pocket column of the pass (local index)
        const int npc = ncols - c_lo;
        for (int idx = tid; idx < npc * kHid; idx += kT5Compute) {
            const int k = idx / npc, c = c_lo + (idx - k * npc);
            const float* src = red + k * kLdc + c * L;
            float sum = 0.0f;
            for (int r = 0; r < L; ++r) sum += src[r];
            dajt[k * Kpad + I[IN_POCKET + (e0 + c - (L - 1))]] = sum;
        }
    }
    T5_PSTAMP(11);
    t5_bar_compute();
    T5_PSTAMP(12);
}

// ---- compute threads, layer 1: one pass of message-only pairs (model.py:151: the unmasked message sum) — no contraction at all:
//      dm1 = mult W2^T dMsum[i] .* relu'(m1), and mult m1 joins the per-row sums behind dW2 ----
__device__ __forceinline__ void t5_message_only_pass(uint8_t* sb, float* S, const T5Map& T, const BwdArgs& g, const PairRef pr, float mult,
                                                     const float* __restrict__ ajt, float* __restrict__ dajt, const int* I, int L, int Wr,
                                                     int pass_base, int npass) {
    const BwdMap& M = T.b;
    const LayerArgs& a = g.a;
    const int tid = threadIdx.x;
    const int p = tid & (kBwdPairs - 1), half = tid >> 7, n0 = 32 * half;
    const int i = pr.i, j = pr.j;
    const bool act = pr.active;
    const bool pep = (j >= 0 && j < kN);
    const int Kpad = a.Kpad;
    float* red = reinterpret_cast<float*>(sb + T5_RED);
    float* red2 = reinterpret_cast<float*>(sb + T5_RED2);
    int* sPl = reinterpret_cast<int*>(S + M.Pl);
    float m1h[32];
    compute_m1<0, 32>(m1h, n0, S, M, a.params, ajt, Kpad, i, j);
    const float* gv = S + T.G + i * kHid + n0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        red[(n0 + k) * kLdc + p] = (act && m1h[k] > 0.0f) ? mult * gv[k] : 0.0f;
        red2[(n0 + k) * kLdc + p] = act ? mult * m1h[k] : 0.0f;
    }
    if (tid == 0) sPl[0] = 0;
    t5_bar_compute();
    if (half == 0 && act && pep) sPl[4 + atomicAdd(sPl, 1)] = p | (j << 8) | ((kN - 1 + i - j) << 16);
    t5_bar_compute();
    {
        const int n_pl = sPl[0];
        for (int idx = tid; idx < n_pl * kHid; idx += kT5Compute) {
            const int e = sPl[4 + (idx >> 6)], k = idx & 63;
            const float x = red[k * kLdc + (e & 255)];
            atomicAdd(S + M.dAjPep + ((e >> 8) & 255) * kLdN + k, x);
            atomicAdd(S + M.dWe + (e >> 16) * kLdN + k, x);
        }
    }
    if (act && j >= kN) {
        // masked pocket slot with non-zero features (rare): straight to the A_j^T gradient scratch
        for (int k = 0; k < 32; ++k) atomicAdd(dajt + (n0 + k) * Kpad + j, red[(n0 + k) * kLdc + p]);
    }
    for (int idx = tid; idx < L * kHid; idx += kT5Compute) {
        const int rl = idx >> 6, n = idx & 63;
        const int lo = max(rl * Wr, pass_base), hi = min((rl + 1) * Wr, pass_base + npass);
        if (hi <= lo) continue;
        float sum = 0.0f, s1 = 0.0f;
        for (int gp = lo; gp < hi; ++gp) {
            sum += red[n * kLdc + (gp - pass_base)];
            s1 += red2[n * kLdc + (gp - pass_base)];
        }
        const int ri = I[IN_ROWS + rl];
        S[M.dAi + ri * kLdN + n] += sum;
        S[T.S1 + ri * kHid + n] += s1;
    }
    t5_bar_compute();
}

template <int LAYER>
__global__ void __launch_bounds__(kT5Threads, 1) egnn_layer_backward_t5_kernel(BwdArgs g, T5Args x) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* sb = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    float* S = reinterpret_cast<float*>(sb);
    const LayerArgs& a = g.a;
    const T5Map T = make_t5_map(a.Kpad);
    const BwdMap& M = T.b;
    constexpr int base = param_offset(LAYER, 0);
    constexpr int layer_numel = param_offset(LAYER + 1, 0) - base;
#ifdef PMHC_T5_STAMPS
    const long long st_entry = clock64();
#endif
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool mma_warp = warp == 8;
    const int Kpad = a.Kpad, P = a.P;
    int* I = reinterpret_cast<int*>(S + M.f.Ints);
    float* direct = g.partial + (size_t)blockIdx.x * g.partial_stride + kTileFloats;
    uint64_t* bars = reinterpret_cast<uint64_t*>(S + T.Bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(S + T.Bars + 24);
    const uint32_t sbase = tc::smem_u32(sb);

    if (tid < 16) S[T.B3 + tid] = 0.0f;
    if (tid == 0) reinterpret_cast<int*>(S + M.Pl)[0] = 0;
    if (mma_warp) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        for (int q = 0; q < 4; ++q) { tc::mbar_init(bars + q, kT5Compute); tc::mbar_init(bars + 4 + q, 1); }
        tc::mbar_init(bars + 8, 1);
        tc::mbar_init(bars + 9, 1);
        tc::mbar_fence_init();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    // the folded weights, c_h and the parameter packs: two TMA bulk copies, in flight under the zeroing of the partial row
    if (tid == 0) {
        tc::mbar_expect_tx(bars + 8, kFoldImageBytes);
        tc::bulk_g2s(sb + T5_F, x.wimg, 4 * 8192, bars + 8);
        tc::bulk_g2s(S + T.Cvec, x.wimg + 4 * 8192, kFoldTailFloats * 4, bars + 8);
    }
    if (LAYER == 1) {   // (layer 1's partial rows were zeroed by bwd_feature_pre_kernel, which already added feature_mlp's gradients)
        float4* d4 = reinterpret_cast<float4*>(g.partial + (size_t)blockIdx.x * g.partial_stride + kTileFloats);   // 16-byte aligned: stride and kTileFloats are multiples of 4
        for (int idx = tid; idx < layer_numel / 4; idx += kT5Threads) d4[idx] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int idx = (layer_numel & ~3) + tid; idx < layer_numel; idx += kT5Threads) direct[idx] = 0.0f;
    }
    tc::mbar_wait_suspend(bars + 8, 0);
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    // gradient-like operands (dout, dpre, dm1) are scaled by the power of two s that brings the largest upstream gradient into [16, 32)
    const float gs = t5_scale_from_max(__uint_as_float(__ldg(x.max_bits))), inv_gs = 1.0f / gs;
    uint32_t np = 0;          // attention-carrying passes of this CTA so far (mbarrier phase, accumulate flag)
    uint32_t nrec = 0;        // records loaded so far (mbarrier phase)
#ifdef PMHC_T5_STAMPS
    long long st_acc[6] = {0, 0, 0, 0, 0, 0}, st_t = clock64();
    const long long st_first = st_t;
#define T5_STAMP(k) do { long long now_ = clock64(); st_acc[k] += now_ - st_t; st_t = now_; } while (0)
#else
#define T5_STAMP(k) do { } while (0)
#endif

    const int4 my = x.sched[blockIdx.x];
    int b = my.x, q0 = my.y, remaining = my.z, slot = my.w;
    for (; remaining > 0; ++b, q0 = 0) {
        T5_STAMP(5);
        if (q0 == 0 && b != my.x) slot = x.segs[b].x;      // a complex this CTA starts gets the complex's first accumulator slot
        // the complex's record (projections, geometry, lists) by one TMA bulk copy; its A_j^T stays in global memory (L2)
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)t5_record_floats(Kpad) * 4u;
            tc::mbar_expect_tx(bars + 9, bytes);
            tc::bulk_g2s(S + T.Rec, x.rec_all + (size_t)b * t5_record_floats(Kpad), bytes, bars + 9);
        }
        tc::mbar_wait_suspend(bars + 9, nrec & 1u);
        ++nrec;
        ComplexInfo ci;
        ci.L = I[IN_POCKET + Kpad + 0]; ci.nv = I[IN_POCKET + Kpad + 1]; ci.nx = I[IN_POCKET + Kpad + 2]; ci.c0 = I[IN_POCKET + Kpad + 3];
        const float* ajt = x.ajt_all + (size_t)b * kHid * Kpad;
        float* dajt = x.dajt_all + (size_t)b * kHid * Kpad;
        T5_STAMP(0);
        const int L = ci.L;
        const int W = (L - 1) + ci.nv;
        for (int idx = tid; idx < M.grads_end - T.G; idx += kT5Threads) S[T.G + idx] = 0.0f;
        __syncthreads();
        bwd_prologue<LAYER, false>(S, M, g, b, I, L, W, direct, kT5Threads);
        if (LAYER == 1 && q0 > 0) {
            // a later segment of a shared complex: the row-level terms of the input gradients belong to its first segment
            __syncthreads();
            for (int idx = tid; idx < kN * (4 + 3 + 14); idx += kT5Threads) S[M.dQ + idx] = 0.0f;
        }
        if (LAYER == 0) {
            // dL / d(message sum) and G[i] = W2^T dMsum[i] (added to the message gradient of every pair of row i), from the node pre-kernel
            const float* rec = x.dmsum_g + (size_t)b * 2 * kN * kHid;
            for (int idx = tid; idx < kN * kHid; idx += kT5Threads) {
                S[M.dMsum + idx] = __ldcg(rec + idx);
                S[T.G + idx] = __ldcg(rec + kN * kHid + idx);
            }
        }
        __syncthreads();
        // ---------------- attention-carrying pairs ----------------
        T5_STAMP(1);
        const int cpp = L > 0 ? kBwdPairs / L : 1;                  // whole neighbour columns per pass
        const int npasses_all = L > 0 ? (W + cpp - 1) / cpp : 0;
        const int units_b = npasses_all > 0 ? npasses_all : 1;
        const int q1 = min(units_b, q0 + remaining);                // this segment: units [q0, q1) of the complex
        const bool last_seg = q1 == units_b;
        const int npasses = min(q1, npasses_all) - q0;              // attention-carrying passes of the segment
        remaining -= q1 - q0;
        if (mma_warp) {
            for (int q = 0; q < npasses; ++q) t5_issue_pass(sbase, tmem, bars, np + q, q);
        } else {
            const int pcol = tid & (kBwdPairs - 1);
            const int pc = L > 0 ? pcol / L : 0, prl = L > 0 ? pcol - pc * L : 0;   // this thread's column (local) and peptide row
            auto decode = [&](int q, PairRef& pr, int& rl, int& e0, int& ncols) {
                e0 = (q0 + q) * cpp;
                ncols = min(cpp, W - e0);
                pr.active = pc < ncols;
                const int e = pr.active ? e0 + pc : e0;
                rl = pr.active ? prl : 0;
                pr.i = I[IN_ROWS + rl];
                pr.j = e < L - 1 ? I[IN_ROWS + (e < rl ? e : e + 1)] : I[IN_POCKET + (e - (L - 1))];
            };
            float aj[32], logit = 0.0f;
            uint2 m1state = make_uint2(0u, 0u);
            PairRef pr, nxt;
            int rl = 0, e0 = 0, ncols = 0, rl_n = 0, e0_n = 0, ncols_n = 0;
            if (npasses > 0) {
                decode(0, pr, rl, e0, ncols);
                t5_request_pair(aj, logit, g, b, ajt, Kpad, pr, tid >> 7);
            }
            for (int q = 0; q < npasses; ++q) {
                const bool has_next = q + 1 < npasses;
                if (has_next) decode(q + 1, nxt, rl_n, e0_n, ncols_n);
                else nxt = pr;
                t5_heads_pass<LAYER>(sb, S, T, g, pr, b, ajt, dajt, I, L, rl, e0, ncols, np + q, tmem, gs, inv_gs, aj, logit, nxt, has_next, m1state);
                pr = nxt; rl = rl_n; e0 = e0_n; ncols = ncols_n;
            }
            if (npasses > 0) tc::mbar_wait_suspend(bars + 4 + 3, (np + npasses - 1) & 1u);   // the last pass's per-row sums
            // ---------------- layer 1: message-only pairs (self, masked peptide / pocket slots) ----------------
            T5_STAMP(2);
            if (LAYER == 0 && L > 0 && last_seg) {
                const int npx = kN - L;
                const int W2 = 1 + npx + ci.nx + (ci.c0 > 0 ? 1 : 0);
                const int total2 = L * W2;
                for (int pass_base = 0; pass_base < total2; pass_base += kBwdPairs) {
                    const int npass = min(kBwdPairs, total2 - pass_base);
                    const bool act = pcol < npass;
                    const int gp = act ? pass_base + pcol : pass_base;
                    const int rl = gp / W2, e = gp - rl * W2;
                    PairRef pr;
                    pr.i = I[IN_ROWS + rl];
                    pr.active = act;
                    float mult = 1.0f;
                    if (e == 0) pr.j = pr.i;
                    else if (e <= npx) pr.j = I[IN_PEPX + e - 1];
                    else if (e <= npx + ci.nx) pr.j = I[IN_POCKET + Kpad - 1 - (e - npx - 1)];
                    else { pr.j = -1; mult = (float)ci.c0; }
                    t5_message_only_pass(sb, S, T, g, pr, mult, ajt, dajt, I, L, W2, pass_base, npass);
                }
            }
        }
        if (npasses > 0 && tid >= 64 && tid < 128) {
            // dL / dT_t[i][n] = the torsion head's hidden-layer gradient summed over the pairs of row i (TM_SEL rows 64..127)
            float sel[16], row[16];
            tc::fence_after_thread_sync();
            tc::tmem_ld16(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + TM_SEL, sel);
            tc::tmem_ld16(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + TM_ROW, row);
#pragma unroll
            for (int rl = 0; rl < kN; ++rl)
                if (rl < L) {
                    S[M.dTt + I[IN_ROWS + rl] * kHid + (tid - 64)] = sel[rl] * inv_gs;
                    S[M.dAi + I[IN_ROWS + rl] * kLdN + (tid - 64)] += row[rl] * inv_gs;   // (+=: layer 1's message-only passes add to it)
                }
            tc::fence_before_thread_sync();
        }
        if (LAYER == 0 && npasses > 0 && tid < 64) {
            float row[16];
            tc::fence_after_thread_sync();
            tc::tmem_ld16(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + TM_ROW, row);
#pragma unroll
            for (int rl = 0; rl < kN; ++rl)
                if (rl < L) S[T.S1 + I[IN_ROWS + rl] * kHid + tid] += row[rl];
            tc::fence_before_thread_sync();
        }
        np += npasses;
        __syncthreads();
        T5_STAMP(3);

        // the complex's accumulators leave for the node kernel (message_mlp.0, the torsion columns, biases, layer 2's input gradients)
        {
            float* rec = x.acc + (size_t)slot * kAccFloats;
            // (dL / d(message sum) is the same in every segment of a shared complex: only the first one hands it on)
            const int dm_lo = M.dMsum - T.S1, dm_hi = dm_lo + kN * kHid;
            for (int idx = tid; idx < kAccFloats; idx += kT5Threads) rec[idx] = (q0 > 0 && idx >= dm_lo && idx < dm_hi) ? 0.0f : S[T.S1 + idx];
            ++slot;
        }
        __syncthreads();
        T5_STAMP(4);
    }

    // ---------------- the CTA-resident weight-gradient sums: tensor memory -> this CTA's partial (parameter layout) ----------------
    tc::fence_after_thread_sync();
    if (!mma_warp && np > 0) {
        const int half = tid >> 7;
        const int r = tid & 127;                       // accumulator row = TMEM lane: head (r >> 6) of the pair, hidden unit n
        const int n = r & 63, hsel = r >> 6;
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
        for (int pair = 0; pair < 2; ++pair) {
            const int head = 2 * pair + hsel;          // F_ROT, F_TOR | F_TRN, F_ATT
            const int wid = head == F_ROT ? ROT0_W : head == F_TOR ? TOR0_W : head == F_TRN ? TRN0_W : ATT0_W;
            const int ld = head == F_ROT ? 68 : head == F_TOR ? 78 : head == F_TRN ? 64 : 66;
            const int woff = (head == F_ROT ? param_offset(LAYER, ROT0_W) : head == F_TOR ? param_offset(LAYER, TOR0_W)
                              : head == F_TRN ? param_offset(LAYER, TRN0_W) : param_offset(LAYER, ATT0_W)) - base;
            (void)wid;
            (void)ld; (void)woff;
            {   // through a padded shared-memory tile, so that the stores to the partial row are coalesced
                float* stage = reinterpret_cast<float*>(sb + T5_HID);      // [128][65]
                uint32_t v[32];
                tc::tmem_ld32_nowait(tlane + (pair == 0 ? TM_DFA : TM_DFB) + 32 * half, v);
                tc::tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) stage[r * (kHid + 1) + 32 * half + k] = __uint_as_float(v[k]) * inv_gs;
                t5_bar_compute();
                for (int idx = tid; idx < 2 * kHid * kHid; idx += kT5Compute) {
                    const int rr = idx >> 6, k = idx & 63, hd = 2 * pair + (rr >> 6), nn = rr & 63;
                    const int ldh = hd == F_ROT ? 68 : hd == F_TOR ? 78 : hd == F_TRN ? 64 : 66;
                    const int wo = (hd == F_ROT ? param_offset(LAYER, ROT0_W) : hd == F_TOR ? param_offset(LAYER, TOR0_W)
                                    : hd == F_TRN ? param_offset(LAYER, TRN0_W) : param_offset(LAYER, ATT0_W)) - base;
                    direct[wo + nn * ldh + k] = stage[rr * (kHid + 1) + k];
                }
                t5_bar_compute();
            }
            if (half == 0 && pair == 0) {
                // rows 0..63: rotation (pair A) and translation (pair B) share the accumulator rows; 64..127: torsion and attention
                float w3[16], ex[16];
                tc::tmem_ld16(tlane + TM_DW3, w3);
                tc::tmem_ld16(tlane + TM_EXT, ex);
                if (hsel == 0) {
                    const int wrot = param_offset(LAYER, ROT0_W) - base;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        direct[(param_offset(LAYER, ROT2_W) - base) + c * kHid + n] = w3[c] * inv_gs;
                        direct[wrot + n * 68 + 64 + c] = ex[c] * inv_gs;
                    }
                    direct[(param_offset(LAYER, ROT0_B) - base) + n] = ex[6] * inv_gs;
                    direct[(param_offset(LAYER, TRN2_W) - base) + n] = w3[4] * inv_gs;
                    direct[(param_offset(LAYER, TRN0_B) - base) + n] = ex[10] * inv_gs;
                } else {
                    const int watt = param_offset(LAYER, ATT0_W) - base;
#pragma unroll
                    for (int c = 0; c < PMHC_NTORS; ++c) direct[(param_offset(LAYER, TOR2_W) - base) + c * kHid + n] = w3[8 + c] * inv_gs;
                    direct[(param_offset(LAYER, TOR0_B) - base) + n] = ex[6] * inv_gs;
                    direct[(param_offset(LAYER, ATT2_W) - base) + n] = w3[5] * inv_gs;
                    direct[watt + n * 66 + 64] = ex[8] * inv_gs;
                    direct[watt + n * 66 + 65] = ex[9] * inv_gs;
                    direct[(param_offset(LAYER, ATT0_B) - base) + n] = ex[10] * inv_gs;
                }
            }
        }
        if (tid < 4) direct[(param_offset(LAYER, ROT2_B) - base) + tid] = S[T.B3 + tid];
        else if (tid < 4 + PMHC_NTORS) direct[(param_offset(LAYER, TOR2_B) - base) + (tid - 4)] = S[T.B3 + tid];
        else if (tid == 11) direct[param_offset(LAYER, TRN2_B) - base] = S[T.B3 + 11];
        else if (tid == 12) direct[param_offset(LAYER, ATT2_B) - base] = S[T.B3 + 12];
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    T5_STAMP(5);
#ifdef PMHC_T5_STAMPS
    if (tid == 0) {
        long long tot = 0;
        for (int k = 0; k < 6; ++k) tot += st_acc[k];
        printf("t5cta<%d> %d passes %u total %lld passes_cycles %lld entry_to_end %lld pre_loop %lld tail %lld\n", LAYER, (int)blockIdx.x, np, tot, st_acc[2], clock64() - st_entry, st_first - st_entry, st_acc[5]);
    }
    if (blockIdx.x == 0 && tid == 0)
        for (int hh = 0; hh < 2; ++hh) {
            printf("t5<%d> thread %d:", LAYER, 128 * hh);
            for (int k = 0; k < 14; ++k) { printf(" %lld", t5_dbg[hh][k] / (np ? np : 1)); t5_dbg[hh][k] = 0; }
            printf("\n");
        }
    if (blockIdx.x == 0 && tid == 0) {
        printf("t5<%d> node level:", LAYER);
        for (int k = 0; k < 16; ++k) { printf(" %lld", node_dbg[k]); node_dbg[k] = 0; }
        printf("\n");
    }
    if (blockIdx.x == 0 && tid == 0)
        printf("t5<%d> cta0: setup %lld | prologue+G+image %lld | head passes %lld (%u) | message-only %lld | node level %lld | other %lld cycles\n", LAYER,
               st_acc[0], st_acc[1], st_acc[2], np, st_acc[3], st_acc[4], st_acc[5]);
#endif
    if (mma_warp) tc::tmem_dealloc(tmem, 512);
}


// ---------------------------------------------------------------------------------------------------------------------
// Work split of the pair kernel: a complex is max(passes, 1) work units; CTA c takes the units [c T / G, (c + 1) T / G) of the batch's
// unit stream, so a complex may be shared by consecutive CTAs (each of them a SEGMENT with its own accumulator slot, summed by the node
// kernel in slot order).  One warp; thread 0 walks the complexes (B + G steps).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bwd_schedule_kernel(const int* __restrict__ units, int B, int G, int4* __restrict__ sched, int2* __restrict__ segs) {
    // P[b] = units before complex b (exclusive scan), start(c) = floor(c T / Ga) for the Ga = min(G, T) CTAs that get work (their starts
    // are distinct).  Complex b is cut at every start strictly inside (P[b], P[b + 1]); its segments get consecutive slots.
    extern __shared__ int sm[];              // [B + 1] P, then [B + 1] slot0
    __shared__ int wsum[32];
    __shared__ int carry_s;
    int* P = sm;
    int* S0 = sm + B + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto block_excl_scan = [&](auto value_of, int* out) {      // out[b] = sum of value_of(b') for b' < b, out[B] = total
        if (tid == 0) carry_s = 0;
        __syncthreads();
        for (int base = 0; base < B; base += 1024) {
            const int b = base + tid;
            const int v = b < B ? value_of(b) : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                int w = wsum[lane], winc = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
                wsum[lane] = winc - w;
            }
            __syncthreads();
            const int carry = carry_s;
            if (b < B) out[b] = carry + wsum[warp] + inc - v;
            __syncthreads();
            if (tid == 1023) carry_s = carry + wsum[31] + inc;
            __syncthreads();
        }
        if (tid == 0) out[B] = carry_s;
        __syncthreads();
    };
    block_excl_scan([&](int b) { return __ldg(units + b); }, P);
    const long long T = P[B];
    const int Ga = (int)(T < G ? T : G);
    auto start = [&](int c) { return (int)((long long)c * T / Ga); };               // c in [0, Ga]
    auto cnt_lt = [&](int x) {        // number of c in [0, Ga) with start(c) < x   (x in [0, T])
        if (x <= 0) return 0;
        int c = (int)(((long long)x * Ga) / T);                                     // start(c) <= x
        if (c > Ga) c = Ga;
        while (c > 0 && start(c - 1) >= x) --c;
        while (c < Ga && start(c) < x) ++c;
        return c;
    };
    // segments of a complex = 1 + starts strictly inside it
    block_excl_scan([&](int b) { return 1 + cnt_lt(P[b + 1]) - cnt_lt(P[b] + 1); }, S0);
    for (int b = tid; b < B; b += 1024) segs[b] = make_int2(S0[b], S0[b + 1] - S0[b]);
    for (int c = tid; c < G; c += 1024) {
        int4 v = make_int4(0, 0, 0, 0);
        if (c < Ga) {
            const int st = start(c), en = start(c + 1);
            int lo = 0, hi = B - 1;                       // the complex with P[b] <= st < P[b + 1]
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (P[mid] <= st) lo = mid; else hi = mid - 1; }
            while (P[lo + 1] <= st) ++lo;                 // (complexes always have >= 1 unit; defensive)
            v = make_int4(lo, st - P[lo], en - st, S0[lo] + (cnt_lt(st + 1) - cnt_lt(P[lo] + 1)));
        }
        sched[c] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Node kernel: everything of a complex that is per NODE, not per pair — message_mlp.0 (its node, pocket and relative-position
// columns), torsion_mlp.0's torsion columns, the biases, layer 1's message-sum term of message_mlp.2 and layer 2's input gradients —
// from the accumulators the pair kernel left in global memory.  Same code as the other backward kernels' node level
// (bwd_node_level), but with 32 warps per SM instead of the pair kernel's 9: these are short shared-memory-bound products.
// CTA c adds into row c of the per-CTA partials after the pair kernel's CTA c, so the fixed-order reduction is unchanged.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kPostThreads = 1024;
template <int LAYER>
__global__ void __launch_bounds__(kPostThreads, 1) bwd_node_post_kernel(BwdArgs g, const float* __restrict__ acc, float* __restrict__ dajt_all,
                                                                        const int2* __restrict__ segs) {
    extern __shared__ __align__(16) float S[];
    const LayerArgs& a = g.a;
    const PostMap T = make_post_map();
    const BwdMap& M = T.b;
    constexpr int base = param_offset(LAYER, 0);
    const int tid = threadIdx.x;
    const int Kpad = a.Kpad, P = a.P;
    float* direct = g.partial + (size_t)blockIdx.x * g.partial_stride + kTileFloats;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        for (int idx = tid; idx < kN * kHid; idx += kPostThreads) {
            const int i = idx >> 6, c = idx & 63;
            float v;
            if (LAYER == 0) v = (c < PMHC_NFEAT) ? a.feat_in[((size_t)b * kN + i) * PMHC_NFEAT + c] : (c == PMHC_NFEAT ? time_feature(a) : 0.0f);
            else v = a.feat_in[((size_t)b * kN + i) * kHid + c];
            S[M.f.H + i * kLdN + c] = v;
        }
        for (int idx = tid; idx < kN * 14; idx += kPostThreads) S[M.f.Tors + idx] = a.tors_in[(size_t)b * kN * 14 + idx];
        const int2 sg = segs[b];
        for (int u = 0; u < sg.y; ++u) {                      // segments in slot order; the loads of a segment are independent
            const float* rec = acc + (size_t)(sg.x + u) * kAccFloats;
#pragma unroll 8
            for (int idx = tid; idx < kAccFloats; idx += kPostThreads) {
                const float v = __ldcg(rec + idx);
                S[T.S1 + idx] = u == 0 ? v : S[T.S1 + idx] + v;
            }
        }
        __syncthreads();
        if (LAYER == 0) {
            // message_mlp.2 through the message sum: dW2[k][c] += sum_i dMsum[i][k] S1[i][c], db2[k] += (16 + P) sum_i dMsum[i][k]
            float* dW2 = direct + (param_offset(LAYER, MSG2_W) - base);
            rmw_batched<4>(dW2, kHid * kHid, kPostThreads, [&](int idx) {
                const int k = idx >> 6, c = idx & 63;
                float sum = 0.0f;
#pragma unroll
                for (int i = 0; i < kN; ++i) sum = fmaf(S[M.dMsum + i * kHid + k], S[T.S1 + i * kHid + c], sum);
                return sum;
            });
            for (int k = tid; k < kHid; k += kPostThreads) {
                float sum = 0.0f;
                for (int i = 0; i < kN; ++i) sum += S[M.dMsum + i * kHid + k];      // (rows that are not real hold zeros)
                direct[(param_offset(LAYER, MSG2_B) - base) + k] += (float)(kN + P) * sum;
            }
        }
        bwd_node_level<LAYER, kLdt, false, true>(S, M, g, b, nullptr, dajt_all + (size_t)b * kHid * Kpad, direct, kPostThreads);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Layer 1's node pre-kernel: the node feature MLP backward (model.py:151, :407) of every complex at 32 warps per SM — zeroes the
// per-CTA partial rows of the layer, adds feature_mlp.{0,2}'s gradients to them, and leaves dL / d(message sum) and W2^T of it in
// global memory for the pair kernel.
// ---------------------------------------------------------------------------------------------------------------------
struct PreMap {
    BwdMap b;
    int W2s, G, total_bytes;
};
__host__ __device__ inline PreMap make_pre_map() {
    PreMap t;
    BwdMap& m = t.b;
    int o = 0;
    m.BufA = o;     o += 2 * kN * kLdN + kHid * (kH1 + kHid);     // hid, dO, feature_mlp.0.weight
    o = (o + 3) & ~3;
    m.BufB = o;     o += kN * kLdN + kHid * kHid;                 // dhid, feature_mlp.2.weight
    o = (o + 3) & ~3;
    m.f.H = o;      o += kN * kLdN;
    o = (o + 3) & ~3;
    m.f.Msum = o;   o += kN * kHid;
    m.dMsum = o;    o += kN * kHid;
    t.G = o;        o += kN * kHid;
    t.W2s = o;      o += kHid * kHid;
    m.f.Ints = o;   o += 64;
    m.total_floats = o;
    t.total_bytes = o * 4;
    return t;
}
__global__ void __launch_bounds__(kPostThreads, 1) bwd_feature_pre_kernel(BwdArgs g, float* __restrict__ dmsum_g) {
    extern __shared__ __align__(16) float S[];
    const LayerArgs& a = g.a;
    const PreMap T = make_pre_map();
    const BwdMap& M = T.b;
    constexpr int base = param_offset(0, 0);
    constexpr int layer_numel = param_offset(1, 0) - base;
    const int tid = threadIdx.x, lane = tid & 31;
    int* I = reinterpret_cast<int*>(S + M.f.Ints);
    float* direct = g.partial + (size_t)blockIdx.x * g.partial_stride + kTileFloats;
    for (int idx = tid; idx < layer_numel; idx += kPostThreads) direct[idx] = 0.0f;
    const float* W2 = a.params + param_offset(0, MSG2_W);
    for (int idx = tid; idx < kHid * kHid; idx += kPostThreads) S[T.W2s + idx] = __ldg(W2 + idx);
    __syncthreads();
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        if (tid < 32) {
            const bool real = lane < kN && a.mask[(size_t)b * kN + lane] != 0;
            const unsigned bal = __ballot_sync(0xffffffffu, real);
            if (real) I[IN_ROWS + __popc(bal & ((1u << lane) - 1u))] = lane;
            if (lane == 0) I[IN_PEPX] = __popc(bal);
        }
        for (int idx = tid; idx < kN * kHid; idx += kPostThreads) {
            const int i = idx >> 6, c = idx & 63;
            S[M.f.H + i * kLdN + c] = (c < PMHC_NFEAT) ? a.feat_in[((size_t)b * kN + i) * PMHC_NFEAT + c] : (c == PMHC_NFEAT ? time_feature(a) : 0.0f);
            S[M.f.Msum + idx] = g.msum[(size_t)b * kN * kHid + idx];
            S[M.dMsum + idx] = 0.0f;
        }
        __syncthreads();
        const int L = I[IN_PEPX];
        bwd_feature_mlp(S, M, g, b, I, L, direct, kPostThreads, /*weights_staged=*/b != (int)blockIdx.x);
        __syncthreads();
        float* rec = dmsum_g + (size_t)b * 2 * kN * kHid;
        for (int idx = tid; idx < kN * kHid; idx += kPostThreads) {
            const int i = idx >> 6, c = idx & 63;
            float sum = 0.0f;
#pragma unroll 16
            for (int k = 0; k < kHid; ++k) sum = fmaf(S[T.W2s + k * kHid + c], S[M.dMsum + i * kHid + k], sum);
            rec[idx] = S[M.dMsum + idx];
            rec[kN * kHid + idx] = sum;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Setup pre-kernel: per complex, the per-node projections of message_mlp.0 (A_i with the bias, A_j^T for peptide and pocket slots),
// T_t = torsion_mlp.0's torsion columns applied to the row's torsions (+ the head's folded constant), geometry and the row /
// neighbour lists — setup_complex() of the other kernels, run here with one CTA per complex (several per SM) instead of serially
// inside the persistent pair kernel.
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline SmemMap make_setup_map(int Kpad, int P, int H) {
    SmemMap m;
    int o = 0;
    m.W2T = m.WhT = m.We = m.PkAtt = m.PkRotQ = m.PkRot2 = m.PkMisc = m.PkTor2 = m.Scal = m.Out = -1;
    const int s1 = ((P * 23 + 3) & ~3) + kHid * 23, s2 = kHid * (2 * H + 1) + kHid * 15;
    m.Scr = o;      o += ((s1 > s2 ? s1 : s2) + 3) & ~3;
    m.Msum = o;     o += kN * kHid;
    m.H = o;        o += kN * kLdN;
    o = (o + 3) & ~3;
    o = t5_layout_record(o, m, Kpad);
    m.total_floats = o;
    return m;
}
constexpr int kSetupThreads = 512;
template <int LAYER>
__global__ void __launch_bounds__(kSetupThreads) bwd_setup_pre_kernel(LayerArgs a, float* __restrict__ ajt_all, float* __restrict__ rec_all,
                                                            const uint8_t* __restrict__ wimg, float* __restrict__ dajt_all, int* __restrict__ units) {
    extern __shared__ __align__(16) float S[];
    const SmemMap M = make_setup_map(a.Kpad, a.P, layer_H(LAYER));
    const int b = blockIdx.x;
    setup_complex<LAYER>(S, M, a, b, ajt_all + (size_t)b * kHid * a.Kpad);
    const float* cvec_tor = reinterpret_cast<const float*>(wimg + 4 * 8192) + F_TOR * kHid;
    for (int idx = threadIdx.x; idx < kN * kHid; idx += kSetupThreads) S[M.Tt + idx] += __ldg(cvec_tor + (idx & 63));
    __syncthreads();
    const int n = t5_record_floats(a.Kpad);
    float* rec = rec_all + (size_t)b * n;
    for (int idx = threadIdx.x; idx < n; idx += kSetupThreads) rec[idx] = S[M.Ai + idx];
    // dL / dA_j^T of the complex starts at zero (the pair kernel's segments store disjoint pocket columns)
    float* dajt = dajt_all + (size_t)b * kHid * a.Kpad;
    for (int idx = threadIdx.x; idx < kHid * a.Kpad; idx += kSetupThreads) dajt[idx] = 0.0f;
    if (threadIdx.x == 0) {
        const int* I = reinterpret_cast<const int*>(S + M.Ints);
        const int L = I[IN_POCKET + a.Kpad + 0], W = L - 1 + I[IN_POCKET + a.Kpad + 1];
        const int cpp = L > 0 ? kBwdPairs / L : 1;
        const int npasses = L > 0 ? (W + cpp - 1) / cpp : 0;
        units[b] = npasses > 0 ? npasses : 1;
    }
}

}  // namespace pmhc
