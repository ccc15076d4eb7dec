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
//     are scaled by a power of two s = 2^-floor(log2 max|upstream gradient|) so they sit in fp16's normal range; accumulators
//     are read back with 1 / s.
//   * every weight-gradient sum of the CTA stays in tensor memory for the whole launch (dF: two M = 128 accumulators over head pairs
//     (rotation, torsion) and (translation, attention), dW3 and the extras / bias columns: 16 columns each) and is written to the
//     CTA's partial once, at the end; a pass needs three MMA batches (hidden layers of all four heads; pair A; pair B) issued by a
//     dedicated warp, and the 256 compute threads work as two threads per pair, one HEAD each (rotation | torsion, then
//     translation | attention), each on the full 64-wide hidden vector read from its TMEM lane.
#pragma once

#include <cuda_fp16.h>

#include "tcgen05.cuh"

namespace pmhc {

constexpr int kT5Threads = 288;   // 8 compute warps (two threads per pair) + 1 MMA-issuing warp
constexpr int kT5Compute = 256;
// operand tiles: byte offsets from the 1024-byte aligned base of dynamic shared memory
constexpr int T5_M1 = 0, T5_F = 16384, T5_HID = 49152, T5_DPRE = 81920, T5_DX = 114688, T5_TILE_BYTES = 131072;
constexpr int T5_RED = T5_HID;                       // fp32 [78][132] reduction tile over the (then free) hid / dpre tiles
constexpr int T5_RED2 = T5_HID + 78 * kLdc * 4;      // message-only passes: fp32 [64][132] tile of mult * m1
static_assert(T5_HID / 4 + 2 * kHid * kLdt + kHid * PMHC_NFEAT + 96 + kN * kHid <= T5_TILE_BYTES / 4, "staging views must fit behind the folded weights");
static_assert(T5_RED2 + 64 * kLdc * 4 <= T5_TILE_BYTES, "reduction tiles must fit over the hid / dpre / dx tiles");
// DX tile columns (fp16): second-layer output gradients of head pair A / B, the extra-input block shared by all heads
constexpr int DX_A = 0 /* rotation 0..3, torsion 8..14 */, DX_B = 16 /* translation 16, attention 24 */, DX_EXT = 32 /* lq 0..3, -d2, qdot2, 1 */;
// tensor memory columns
constexpr int TM_HID = 0, TM_DM1 = 256, TM_DFA = 320, TM_DFB = 384, TM_DW3A = 448, TM_DW3B = 464, TM_EXTA = 480, TM_EXTB = 496;
// image of one layer's folded weights (global memory, written by bwd_fold_weights_kernel): 4 fp16 SW128 tiles [n][k] in the order
// rotation, torsion, translation, attention, then c_h as [4][64] floats
constexpr int kFoldImageBytes = 4 * 8192 + 4 * kHid * 4;
enum { F_ROT = 0, F_TOR = 1, F_TRN = 2, F_ATT = 3 };

struct T5Map {
    BwdMap b;          // float offsets (shared prologue / node-level code reads these)
    int Cvec, Dl, B3, G, S1, Bars, total_bytes;
};

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
    m.f.Ai = o;     o += kN * kLdN;
    o = (o + 3) & ~3;
    m.f.Tt = o;     o += kN * kHid;
    m.f.H = o;      o += kN * kLdN;
    o = (o + 3) & ~3;
    m.f.Tors = o;   o += kN * 2 * PMHC_NTORS;
    t.G = o;        o += kN * kHid;                     // layer 1: W2^T dMsum[i]
    t.S1 = o;       o += kN * kHid;                     // layer 1: sum over all neighbour slots of m1[i, .]
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
    o = (o + 3) & ~3;
    m.f.Q = o;      o += Kpad * 4;
    m.f.X = o;      o += Kpad * 3;
    o = (o + 3) & ~3;
    m.f.Ints = o;   o += Kpad + 64;
    m.total_floats = o;
    m.f.total_floats = o;
    t.total_bytes = o * 4 + 1024;                       // + slack for the 1024-byte alignment of the base
    return t;
}

struct T5Args {
    const uint8_t* wimg;   // this layer's folded-weight image (kFoldImageBytes)
    const float* scale;    // device: [0] = s (power of two applied to gradient-like operands), [1] = 1 / s
};

// ---------------------------------------------------------------------------------------------------------------------
// once per step: F_h = W_h[:, :64] W2 (fp16, SW128 tile), c_h = W_h[:, :64] b2 + b_h (the torsion head's bias lives in T_t)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bwd_fold_weights_kernel(const float* __restrict__ params, uint8_t* __restrict__ img_all) {
    const int layer = blockIdx.x >> 2, head = blockIdx.x & 3;
    __shared__ float sW2[kHid * kHid];
    __shared__ float sWh[kHid * (kHid + 1)];
    const int base = layer == 0 ? param_offset(0, 0) : param_offset(1, 0);
    auto off = [&](int id) { return layer == 0 ? param_offset(0, id) : param_offset(1, id); };
    (void)base;
    int wid, bid, ld;
    if (head == F_ROT)      { wid = ROT0_W; bid = ROT0_B; ld = 68; }
    else if (head == F_TOR) { wid = TOR0_W; bid = -1;     ld = 78; }
    else if (head == F_TRN) { wid = TRN0_W; bid = TRN0_B; ld = 64; }
    else                    { wid = ATT0_W; bid = ATT0_B; ld = 66; }
    const float* W2 = params + off(MSG2_W);
    const float* b2 = params + off(MSG2_B);
    const float* Wh = params + off(wid);
    for (int idx = threadIdx.x; idx < kHid * kHid; idx += 256) {
        sW2[idx] = W2[idx];
        sWh[(idx >> 6) * (kHid + 1) + (idx & 63)] = Wh[(idx >> 6) * ld + (idx & 63)];
    }
    __syncthreads();
    uint8_t* img = img_all + (size_t)layer * kFoldImageBytes;
    for (int idx = threadIdx.x; idx < kHid * kHid; idx += 256) {
        const int n = idx >> 6, c = idx & 63;
        float acc = 0.0f;
#pragma unroll 8
        for (int k = 0; k < kHid; ++k) acc = fmaf(sWh[n * (kHid + 1) + k], sW2[k * kHid + c], acc);
        *reinterpret_cast<__half*>(img + head * 8192 + tc::sw128_offset(n, c)) = __float2half_rn(acc);
    }
    if (threadIdx.x < kHid) {
        const int n = threadIdx.x;
        float acc = bid >= 0 ? params[off(bid) + n] : 0.0f;
        for (int k = 0; k < kHid; ++k) acc = fmaf(sWh[n * (kHid + 1) + k], b2[k], acc);
        reinterpret_cast<float*>(img + 4 * 8192)[head * kHid + n] = acc;
    }
}

// s = 2^-floor(log2 max|x|) over the upstream gradients of a layer launch (frames and torsions): the largest gradient-like
// operand entry then lies within a few binades of 1.  One block.
__global__ void __launch_bounds__(1024) bwd_grad_scale_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb,
                                                              float* __restrict__ out) {
    __shared__ float red[32];
    float mx = 0.0f;
    for (int i = threadIdx.x; i < na; i += 1024) mx = fmaxf(mx, fabsf(a[i]));
    for (int i = threadIdx.x; i < nb; i += 1024) mx = fmaxf(mx, fabsf(b[i]));
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        mx = warp_max(red[threadIdx.x]);
        if (threadIdx.x == 0) {
            float s = 1.0f;
            if (mx > 0.0f && mx < INFINITY) {
                int e;
                frexpf(mx, &e);                 // mx = f 2^e, f in [0.5, 1)
                e = min(max(-(e - 1), -100), 100);
                s = ldexpf(1.0f, e);            // s mx in [1, 2)
            }
            out[0] = s;
            out[1] = 1.0f / s;
        }
    }
}

// After the fixed-order reduction of the per-CTA partials into `red` (parameter layout of the layer, the head first-layer message
// columns holding dF_h): the chain rule through the folded weights, added into the caller's gradient.
template <int LAYER>
__global__ void __launch_bounds__(256) bwd_unfold_kernel(const float* __restrict__ params, const float* __restrict__ red, float* __restrict__ grad) {
    constexpr int base = param_offset(LAYER, 0);
    constexpr int numel = param_offset(LAYER + 1, 0) - base;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= numel) return;
    const int q = p + base;
    const float* W2 = params + param_offset(LAYER, MSG2_W);
    const float* b2 = params + param_offset(LAYER, MSG2_B);
    constexpr int wid[4] = {ROT0_W, TOR0_W, TRN0_W, ATT0_W};
    constexpr int bid[4] = {ROT0_B, TOR0_B, TRN0_B, ATT0_B};
    constexpr int lds[4] = {68, 78, 64, 66};
    float v = red[p];
    bool done = false;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const int o = param_offset(LAYER, wid[h]);
        if (!done && q >= o && q < o + kHid * lds[h]) {
            const int n = (q - o) / lds[h], k = (q - o) - n * lds[h];
            if (k < kHid) {
                // dW_h[n][k] = sum_c dF_h[n][c] W2[k][c] + db_h[n] b2[k]
                const float* dF = red + (o - base) + n * lds[h];
                float acc = red[param_offset(LAYER, bid[h]) - base + n] * b2[k];
#pragma unroll 8
                for (int c = 0; c < kHid; ++c) acc = fmaf(dF[c], __ldg(W2 + k * kHid + c), acc);
                v = acc;
            }
            done = true;
        }
    }
    if (!done && q >= param_offset(LAYER, MSG2_W) && q < param_offset(LAYER, MSG2_W) + kHid * kHid) {
        // dW2[k][c] += sum_h sum_n W_h[n][k] dF_h[n][c]
        const int k = (q - param_offset(LAYER, MSG2_W)) >> 6, c = (q - param_offset(LAYER, MSG2_W)) & 63;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float* Wh = params + param_offset(LAYER, wid[h]);
            const float* dF = red + (param_offset(LAYER, wid[h]) - base);
            float acc = 0.0f;
#pragma unroll 8
            for (int n = 0; n < kHid; ++n) acc = fmaf(__ldg(Wh + n * lds[h] + k), dF[n * lds[h] + c], acc);
            v += acc;
        }
        done = true;
    }
    if (!done && q >= param_offset(LAYER, MSG2_B) && q < param_offset(LAYER, MSG2_B) + kHid) {
        const int k = q - param_offset(LAYER, MSG2_B);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float* Wh = params + param_offset(LAYER, wid[h]);
            const float* db = red + (param_offset(LAYER, bid[h]) - base);
            float acc = 0.0f;
            for (int n = 0; n < kHid; ++n) acc = fmaf(__ldg(Wh + n * lds[h] + k), db[n], acc);
            v += acc;
        }
    }
    grad[q] += v;
}

// partial -> red (plain store): same fixed CTA order as reduce_partials_kernel
__global__ void reduce_partials_to_kernel(const float* __restrict__ partial, int stride, int n_cta, int numel, float* __restrict__ red) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= numel) return;
    float acc = 0.0f;
    for (int c0 = 0; c0 < n_cta; c0 += 16) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = c0 + u < n_cta ? __ldcg(partial + (size_t)(c0 + u) * stride + kTileFloats + p) : 0.0f;
#pragma unroll
        for (int u = 0; u < 16; ++u)
            if (c0 + u < n_cta) acc += v[u];
    }
    red[p] = acc;
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
// write 64 floats (scaled) as one fp16 row of a SW128 tile: 8 chunks of 16 bytes at their swizzled positions
__device__ __forceinline__ void t5_store_row64(uint8_t* tile, int p, const float (&v)[kHid], float scale) {
    uint8_t* row = tile + (p >> 3) * 1024 + (p & 7) * 128;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        uint4 u;
        u.x = pack_h2_sat(v[8 * q + 0] * scale, v[8 * q + 1] * scale);
        u.y = pack_h2_sat(v[8 * q + 2] * scale, v[8 * q + 3] * scale);
        u.z = pack_h2_sat(v[8 * q + 4] * scale, v[8 * q + 5] * scale);
        u.w = pack_h2_sat(v[8 * q + 6] * scale, v[8 * q + 7] * scale);
        *reinterpret_cast<uint4*>(row + ((q ^ (p & 7)) << 4)) = u;
    }
}
// one 16-byte chunk (8 fp16 columns starting at column 8 * chunk) of row p
__device__ __forceinline__ void t5_store_chunk(uint8_t* tile, int p, int chunk, float v0, float v1, float v2, float v3, float v4, float v5,
                                               float v6, float v7) {
    uint4 u;
    u.x = pack_h2_sat(v0, v1); u.y = pack_h2_sat(v2, v3); u.z = pack_h2_sat(v4, v5); u.w = pack_h2_sat(v6, v7);
    *reinterpret_cast<uint4*>(tile + (p >> 3) * 1024 + (p & 7) * 128 + ((chunk ^ (p & 7)) << 4)) = u;
}
// the thread's 64 accumulator columns of TMEM lane (pair) p
__device__ __forceinline__ void t5_load64(uint32_t taddr, float (&v)[kHid]) {
    uint32_t r0[32], r1[32];
    tc::tmem_ld32_nowait(taddr, r0);
    tc::tmem_ld32_nowait(taddr + 32, r1);
    tc::tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 32; ++c) { v[c] = __uint_as_float(r0[c]); v[32 + c] = __uint_as_float(r1[c]); }
}

// ---- the MMA-issuing warp: the three batches of one attention-carrying pass (np = running pass count of the CTA) ----
__device__ __forceinline__ void t5_issue_pass(uint32_t sbase, uint32_t tmem, uint64_t* bars, uint32_t np) {
    const uint32_t par = np & 1u;
    const uint32_t first = np == 0 ? 0u : 1u;     // accumulate flag of the CTA-resident weight-gradient sums
    uint64_t* rdy = bars;        // [0..2]
    uint64_t* done = bars + 3;   // [0..2]
    constexpr uint32_t id_hid = tc::idesc_f16_f32(128, 128);
    constexpr uint32_t id_w16 = tc::idesc_f16_f32_major(128, 16, 1, 1);
    constexpr uint32_t id_w64 = tc::idesc_f16_f32_major(128, 64, 1, 1);
    constexpr uint32_t id_dm1 = tc::idesc_f16_f32_major(128, 64, 0, 1);
    // batch 0: hidden layers of the four heads (two N = 128 contractions over the folded weights)
    tc::mbar_wait_suspend(rdy + 0, par);
    tc::fence_after_thread_sync();
    if (tc::elect_one()) {
        const uint64_t da = tc::smem_desc_sw128(sbase + T5_M1);
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
#pragma unroll
    for (int pair = 0; pair < 2; ++pair) {
        tc::mbar_wait_suspend(rdy + 1 + pair, par);
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
            const uint32_t dx = sbase + T5_DX + (pair == 0 ? DX_A : DX_B) * 2;
            // weight-gradient sums over the pairs of the pass (K = 128 pairs = 8 steps of 16 rows), A = the two heads' tiles stacked in M
#pragma unroll
            for (int s = 0; s < 8; ++s)
                tc::mma_bf16(tmem + (pair == 0 ? TM_DW3A : TM_DW3B), tc::smem_desc(sbase + T5_HID + s * 2048, 16384, 1024, 2),
                             tc::smem_desc(dx + s * 2048, 16384, 1024, 2), id_w16, s > 0 ? 1u : first);
#pragma unroll
            for (int s = 0; s < 8; ++s)
                tc::mma_bf16(tmem + (pair == 0 ? TM_EXTA : TM_EXTB), tc::smem_desc(sbase + T5_DPRE + s * 2048, 16384, 1024, 2),
                             tc::smem_desc(sbase + T5_DX + DX_EXT * 2 + s * 2048, 16384, 1024, 2), id_w16, s > 0 ? 1u : first);
#pragma unroll
            for (int s = 0; s < 8; ++s)
                tc::mma_bf16(tmem + (pair == 0 ? TM_DFA : TM_DFB), tc::smem_desc(sbase + T5_DPRE + s * 2048, 16384, 1024, 2),
                             tc::smem_desc(sbase + T5_M1 + s * 2048, 16384, 1024, 2), id_w64, s > 0 ? 1u : first);
            // dm1 += dpre_h F_h: A K-major, B = the folded weight tile read MN-major (K = hidden units, 4 steps of 16 rows)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint64_t da = tc::smem_desc_sw128(sbase + T5_DPRE + h * 16384);
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    tc::mma_bf16(tmem + TM_DM1, da + 2 * s, tc::smem_desc(sbase + T5_F + (2 * pair + h) * 8192 + s * 2048, 8192, 1024, 2), id_dm1,
                                 (pair | h | s) > 0);
            }
            tc::mma_commit(done + 1 + pair);
        }
        __syncwarp();
    }
}

// ---- compute threads: one attention-carrying pass of up to 128 pairs ----
template <int LAYER>
__device__ __forceinline__ void t5_heads_pass(uint8_t* sb, float* S, const T5Map& T, const BwdArgs& g, const PairRef pr, int b,
                                              const float* __restrict__ ajt, float* __restrict__ dajt, const int* I, int L, int Wr,
                                              int pass_base, int npass, int n_pocket_cols, int pocket_e0, uint32_t np, uint32_t tmem,
                                              float gs, float inv_gs) {
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
    uint64_t* done = bars + 3;
    float* sDl = S + T.Dl;
    float* sB3 = S + T.B3;
    int* sPl = reinterpret_cast<int*>(S + M.Pl);
    const float* cvec = S + T.Cvec;

    // ---- m1 (my half of its features) -> fp16 tile; geometry; the extra-input block ----
    uint32_t m1mask = 0u;
    {
        float m1h[32];
        compute_m1<LAYER, 32>(m1h, 32 * half, S, M, a.params, ajt, Kpad, i, j);
        uint8_t* row = sb + T5_M1 + (p >> 3) * 1024 + (p & 7) * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_h2_sat(m1h[8 * q + 0], m1h[8 * q + 1]); u.y = pack_h2_sat(m1h[8 * q + 2], m1h[8 * q + 3]);
            u.z = pack_h2_sat(m1h[8 * q + 4], m1h[8 * q + 5]); u.w = pack_h2_sat(m1h[8 * q + 6], m1h[8 * q + 7]);
            *reinterpret_cast<uint4*>(row + (((4 * half + q) ^ (p & 7)) << 4)) = u;
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) m1mask |= (m1h[k] > 0.0f ? 1u : 0u) << k;
    }
    const float* rg = S + M.RowG + i * 16;
    const float lse = rg[15], c_i = rg[14];
    const float logit = g.logits[((size_t)b * kN + i) * Kpad + j];
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
        t5_store_chunk(sb + T5_DX, p, 4, lq.w, lq.x, lq.y, lq.z, -d2, qd, 1.0f, 0.0f);
        t5_store_chunk(sb + T5_DX, p, 5, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f);
    }
    if (tid == 0) sPl[0] = 0;      // (every reader of the previous pass's list is behind that pass's last barrier)
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    tc::mbar_arrive(rdy + 0);

    float gi[7] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};   // layer 2: this thread's share of dL / d (q_i, x_i)
    float dLdw_own = 0.0f;                                       // torsion thread: its share of dL / dw
    float hid[kHid];

    tc::mbar_wait_suspend(done + 0, par);
    tc::fence_after_thread_sync();

    // ======================= head pair A: rotation (half 0) | torsion (half 1) =======================
    if (half == 0) {
        t5_load64(tlane + TM_HID + 0, hid);
        float pre[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) pre[c] = S[M.f.Scal + SC_ROT2B + c];
#pragma unroll
        for (int n = 0; n < kHid; ++n) {
            const float4 wq = *reinterpret_cast<const float4*>(S + M.f.PkRotQ + 4 * n);
            const float s = hid[n] + cvec[F_ROT * kHid + n] + wq.x * lq.w + wq.y * lq.x + wq.z * lq.y + wq.w * lq.z;
            const float h = fmaxf(s, 0.0f);
            hid[n] = h;
            const float4 w2 = *reinterpret_cast<const float4*>(S + M.f.PkRot2 + 4 * n);
            pre[0] = fmaf(w2.x, h, pre[0]); pre[1] = fmaf(w2.y, h, pre[1]);
            pre[2] = fmaf(w2.z, h, pre[2]); pre[3] = fmaf(w2.w, h, pre[3]);
        }
        t5_store_row64(sb + T5_HID, p, hid, 1.0f);
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
        t5_store_chunk(sb + T5_DX, p, 0, dp2[0] * gs, dp2[1] * gs, dp2[2] * gs, dp2[3] * gs, 0.0f, 0.0f, 0.0f, 0.0f);
        float dlq[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int n = 0; n < kHid; ++n) {
            const float4 w2 = *reinterpret_cast<const float4*>(S + M.f.PkRot2 + 4 * n);
            float dp = w2.x * dp2[0] + w2.y * dp2[1] + w2.z * dp2[2] + w2.w * dp2[3];
            dp = hid[n] > 0.0f ? dp : 0.0f;
            hid[n] = dp;
            if (IN_GRADS) {
                const float4 wq = *reinterpret_cast<const float4*>(S + M.f.PkRotQ + 4 * n);
                dlq[0] = fmaf(wq.x, dp, dlq[0]); dlq[1] = fmaf(wq.y, dp, dlq[1]);
                dlq[2] = fmaf(wq.z, dp, dlq[2]); dlq[3] = fmaf(wq.w, dp, dlq[3]);
            }
        }
        t5_store_row64(sb + T5_DPRE, p, hid, gs);
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
        t5_load64(tlane + TM_HID + 64, hid);
        float da[PMHC_NTORS];
#pragma unroll
        for (int c = 0; c < PMHC_NTORS; ++c) da[c] = S[M.f.Scal + SC_TOR2B + c];
        const float* tt = S + M.f.Tt + i * kHid;
#pragma unroll
        for (int n = 0; n < kHid; ++n) {
            const float h = fmaxf(hid[n] + tt[n], 0.0f);
            hid[n] = h;
            const float4 w0 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n);
            const float4 w1 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n + 4);
            da[0] = fmaf(w0.x, h, da[0]); da[1] = fmaf(w0.y, h, da[1]); da[2] = fmaf(w0.z, h, da[2]);
            da[3] = fmaf(w0.w, h, da[3]); da[4] = fmaf(w1.x, h, da[4]); da[5] = fmaf(w1.y, h, da[5]);
            da[6] = fmaf(w1.z, h, da[6]);
        }
        t5_store_row64(sb + T5_HID + 16384, p, hid, 1.0f);
        float dda[PMHC_NTORS];
#pragma unroll
        for (int c = 0; c < PMHC_NTORS; ++c) {
            dLdw_own = fmaf(rg[4 + c], da[c], dLdw_own);
            dda[c] = w * rg[4 + c];
        }
        t5_store_chunk(sb + T5_DX, p, 1, dda[0] * gs, dda[1] * gs, dda[2] * gs, dda[3] * gs, dda[4] * gs, dda[5] * gs, dda[6] * gs, 0.0f);
#pragma unroll
        for (int n = 0; n < kHid; ++n) {
            const float4 w0 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n);
            const float4 w1 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n + 4);
            float dp = w0.x * dda[0] + w0.y * dda[1] + w0.z * dda[2] + w0.w * dda[3] + w1.x * dda[4] + w1.y * dda[5] + w1.z * dda[6];
            hid[n] = hid[n] > 0.0f ? dp : 0.0f;
        }
        t5_store_row64(sb + T5_DPRE + 16384, p, hid, gs);
#pragma unroll
        for (int c = 0; c < PMHC_NTORS; ++c) {
            const float sum = warp_sum(dda[c]);
            if (lane == 0) atomicAdd(sB3 + 4 + c, sum);
        }
    }
    tc::fence_proxy_async_smem();
    tc::mbar_arrive(rdy + 1);
    t5_bar_compute();
    if (half == 0 && act && pep) sPl[4 + atomicAdd(sPl, 1)] = p | (j << 8) | ((kN - 1 + i - j) << 16);   // read behind a later barrier
    // per-row sums of the torsion head's hidden-layer gradient (dL / dT_t[i]) from its tile (fp16, scaled)
    for (int idx = tid; idx < L * kHid; idx += kT5Compute) {
        const int rl = idx >> 6, n = idx & 63;
        const int lo = max(rl * Wr, pass_base), hi = min((rl + 1) * Wr, pass_base + npass);
        if (hi <= lo) continue;
        float sum = 0.0f;
        for (int gp = lo; gp < hi; ++gp) {
            const int pp = gp - pass_base;
            sum += __half2float(*reinterpret_cast<const __half*>(sb + T5_DPRE + 16384 + tc::sw128_offset(pp, n)));
        }
        S[M.dTt + I[IN_ROWS + rl] * kHid + n] += sum * inv_gs;
    }

    // ======================= head pair B: translation (half 0) | attention (half 1) =======================
    if (half == 0) {
        t5_load64(tlane + TM_HID + 128, hid);
        float sc = S[M.f.Scal + SC_TRN2B];
#pragma unroll
        for (int n = 0; n < kHid; ++n) {
            const float h = fmaxf(hid[n] + cvec[F_TRN * kHid + n], 0.0f);
            hid[n] = h;
            sc = fmaf(S[M.f.PkMisc + 4 * n + 1], h, sc);
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
        t5_bar_compute();                       // dL/dw shares published; every read of pair A's torsion tile is done
        tc::mbar_wait_suspend(done + 1, par);   // pair A's MMAs have read the hid / dpre tiles
        t5_store_row64(sb + T5_HID, p, hid, 1.0f);
#pragma unroll
        for (int n = 0; n < kHid; ++n) hid[n] = hid[n] > 0.0f ? S[M.f.PkMisc + 4 * n + 1] * ds : 0.0f;
        t5_store_row64(sb + T5_DPRE, p, hid, gs);
        t5_store_chunk(sb + T5_DX, p, 2, ds * gs, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f);
    } else {
        t5_load64(tlane + TM_HID + 192, hid);
#pragma unroll
        for (int n = 0; n < kHid; ++n) {
            const float4 pk = *reinterpret_cast<const float4*>(S + M.f.PkAtt + 4 * n);
            hid[n] = fmaxf((cvec[F_ATT * kHid + n] + hid[n]) + fmaf(pk.y, qd, pk.x * -d2), 0.0f);
        }
        t5_bar_compute();
        const float dLdw = sDl[p] + dLdw_own + sDl[kBwdPairs + p];
        // softmax backward with the saved row statistics; a fully saturated row has dlogit = 0 exactly (see pair_pass)
        const float dlogit = (w == 1.0f) ? 0.0f : w * (dLdw - c_i);
        tc::mbar_wait_suspend(done + 1, par);
        t5_store_row64(sb + T5_HID + 16384, p, hid, 1.0f);
        float gd = 0.0f, gq = 0.0f;
#pragma unroll
        for (int n = 0; n < kHid; ++n) {
            const float4 pk = *reinterpret_cast<const float4*>(S + M.f.PkAtt + 4 * n);
            const float dp = hid[n] > 0.0f ? pk.w * dlogit : 0.0f;
            hid[n] = dp;
            gd = fmaf(pk.x, dp, gd);
            gq = fmaf(pk.y, dp, gq);
        }
        t5_store_row64(sb + T5_DPRE + 16384, p, hid, gs);
        t5_store_chunk(sb + T5_DX, p, 3, dlogit * gs, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f);
        if (IN_GRADS && act) {
            const float f = -gd * 2.0f;                 // d(-d2) = gd
            gi[4] += f * rx; gi[5] += f * ry; gi[6] += f * rz;
            const float fq = gq * 2.0f * dotq;
            gi[0] += fq * qj.w; gi[1] += fq * qj.x; gi[2] += fq * qj.y; gi[3] += fq * qj.z;
            if (pep) {
                atomicAdd(S + M.dX + j * 3 + 0, -f * rx); atomicAdd(S + M.dX + j * 3 + 1, -f * ry); atomicAdd(S + M.dX + j * 3 + 2, -f * rz);
                atomic_add_quat(S + M.dQ + j * 4, qscale(qi, fq));
            }
        }
        const float sum = warp_sum(dlogit);
        if (lane == 0) atomicAdd(sB3 + 12, sum);
    }
    tc::fence_proxy_async_smem();
    tc::mbar_arrive(rdy + 2);

    // ======================= dm1 = (sum_h dpre_h F_h [+ W2^T dMsum_i]) .* relu'(m1) -> reduction tile =======================
    tc::mbar_wait_suspend(done + 2, par);
    tc::fence_after_thread_sync();
    float* red = reinterpret_cast<float*>(sb + T5_RED);
    {
        uint32_t r[32];
        tc::tmem_ld32_nowait(tlane + TM_DM1 + 32 * half, r);
        tc::tmem_wait_ld();
        const float* gv = S + T.G + i * kHid + 32 * half;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            float x = __uint_as_float(r[k]) * inv_gs;
            if (LAYER == 0) x += gv[k];
            red[(32 * half + k) * kLdc + p] = (act && ((m1mask >> k) & 1u)) ? x : 0.0f;
        }
        if (IN_GRADS) {
#pragma unroll
            for (int c = 0; c < 7; ++c) red[(64 + 7 * half + c) * kLdc + p] = act ? gi[c] : 0.0f;
        }
    }
    tc::fence_before_thread_sync();
    t5_bar_compute();
    if (IN_GRADS) {
        for (int idx = tid; idx < L * 7; idx += kT5Compute) {
            const int rl = idx / 7, c = idx - rl * 7;
            const int lo = max(rl * Wr, pass_base), hi = min((rl + 1) * Wr, pass_base + npass);
            float sum = 0.0f;
            for (int gp = lo; gp < hi; ++gp) sum += red[(64 + c) * kLdc + (gp - pass_base)] + red[(71 + c) * kLdc + (gp - pass_base)];
            const int ri = I[IN_ROWS + rl];
            if (hi > lo) {
                if (c < 4) S[M.dQ + ri * 4 + c] += sum;
                else S[M.dX + ri * 3 + (c - 4)] += sum;
            }
        }
    }
    {   // peptide neighbours: dA_j[j] and dW_e[rel] get the column
        const int n_pl = sPl[0];
        for (int idx = tid; idx < n_pl * kHid; idx += kT5Compute) {
            const int e = sPl[4 + (idx >> 6)], k = idx & 63;
            const float x = red[k * kLdc + (e & 255)];
            atomicAdd(S + M.dAjPep + ((e >> 8) & 255) * kLdN + k, x);
            atomicAdd(S + M.dWe + (e >> 16) * kLdN + k, x);
        }
    }
    for (int idx = tid; idx < L * kHid; idx += kT5Compute) {
        const int rl = idx >> 6, n = idx & 63;
        const int lo = max(rl * Wr, pass_base), hi = min((rl + 1) * Wr, pass_base + npass);
        if (hi <= lo) continue;
        float sum = 0.0f, s1 = 0.0f;
        for (int gp = lo; gp < hi; ++gp) {
            sum += red[n * kLdc + (gp - pass_base)];
            if (LAYER == 0) s1 += __half2float(*reinterpret_cast<const __half*>(sb + T5_M1 + tc::sw128_offset(gp - pass_base, n)));
        }
        const int ri = I[IN_ROWS + rl];
        S[M.dAi + ri * kLdN + n] += sum;
        if (LAYER == 0) S[T.S1 + ri * kHid + n] += s1;
    }
    {   // dA_j^T[k][j] += sum over rows of dm1 for the valid pocket columns of this pass
        const int rl_lo = pass_base / Wr, rl_hi = (pass_base + npass - 1) / Wr;
        const int n_items = n_pocket_cols * kHid;
        for (int idx0 = tid; idx0 < n_items; idx0 += 4 * kT5Compute) {   // four L2 read-modify-writes in flight per thread
            float old[4];
            int addr[4], kk[4], ee[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = idx0 + u * kT5Compute;
                kk[u] = idx / n_pocket_cols;
                ee[u] = idx - kk[u] * n_pocket_cols;
                addr[u] = idx < n_items ? kk[u] * Kpad + I[IN_POCKET + ee[u]] : -1;
                old[u] = addr[u] >= 0 ? __ldcg(dajt + addr[u]) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (addr[u] < 0) continue;
                float sum = 0.0f;
                for (int rl = rl_lo; rl <= rl_hi; ++rl) {
                    const int col = rl * Wr + pocket_e0 + ee[u] - pass_base;
                    if (col >= 0 && col < npass) sum += red[kk[u] * kLdc + col];
                }
                dajt[addr[u]] = old[u] + sum;
            }
        }
    }
    t5_bar_compute();
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
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool mma_warp = warp == 8;
    const int Kpad = a.Kpad, P = a.P;
    int* I = reinterpret_cast<int*>(S + M.f.Ints);
    float* ajt = a.ajt_ws + (size_t)blockIdx.x * kHid * Kpad;
    float* dajt = g.dajt_ws + (size_t)blockIdx.x * kHid * Kpad;
    float* direct = g.partial + (size_t)blockIdx.x * g.partial_stride + kTileFloats;
    uint64_t* bars = reinterpret_cast<uint64_t*>(S + T.Bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(S + T.Bars + 16);
    const uint32_t sbase = tc::smem_u32(sb);

    for (int idx = tid; idx < layer_numel; idx += kT5Threads) direct[idx] = 0.0f;
    stage_packs<LAYER>(S, M.f, a.params);
    if (tid < 16) S[T.B3 + tid] = 0.0f;
    if (tid == 0) reinterpret_cast<int*>(S + M.Pl)[0] = 0;
    if (mma_warp) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        for (int q = 0; q < 3; ++q) { tc::mbar_init(bars + q, kT5Compute); tc::mbar_init(bars + 3 + q, 1); }
        tc::mbar_init(bars + 6, 1);
        tc::mbar_fence_init();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const float gs = __ldg(x.scale), inv_gs = __ldg(x.scale + 1);
    uint32_t np = 0;          // attention-carrying passes of this CTA so far (mbarrier phase, accumulate flag)
    bool image_loaded = false;

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const ComplexInfo ci = setup_complex<LAYER>(S, M.f, a, b, ajt);
        const int L = ci.L;
        const int W = (L - 1) + ci.nv;
        for (int idx = tid; idx < M.grads_end - T.G; idx += kT5Threads) S[T.G + idx] = 0.0f;
        for (int idx = tid; idx < kHid * Kpad; idx += kT5Threads) dajt[idx] = 0.0f;
        if (LAYER == 0)
            for (int idx = tid; idx < kN * kHid; idx += kT5Threads) S[M.f.Msum + idx] = g.msum[(size_t)b * kN * kHid + idx];
        __syncthreads();
        bwd_prologue<LAYER>(S, M, g, b, I, L, W, direct, kT5Threads);
        __syncthreads();
        if (LAYER == 0) {
            // G[i][c] = sum_k W2[k][c] dMsum[i][k]: the message-sum gradient behind message_mlp.2, added to every pair of row i
            const float* W2 = a.params + param_offset(LAYER, MSG2_W);
            for (int idx = tid; idx < kN * kHid; idx += kT5Threads) {
                const int i = idx >> 6, c = idx & 63;
                float acc = 0.0f;
#pragma unroll 8
                for (int k = 0; k < kHid; ++k) acc = fmaf(__ldg(W2 + k * kHid + c), S[M.dMsum + i * kHid + k], acc);
                S[T.G + idx] = acc;
            }
        }
        if (!image_loaded) {
            // the folded weights (+ c_h right behind the tiles... c_h is copied separately: it lives behind the pass tiles)
            if (tid == 0) {
                tc::mbar_expect_tx(bars + 6, kFoldImageBytes);
                tc::bulk_g2s(sb + T5_F, x.wimg, 4 * 8192, bars + 6);
                tc::bulk_g2s(S + T.Cvec, x.wimg + 4 * 8192, 4 * kHid * 4, bars + 6);
            }
            tc::mbar_wait_suspend(bars + 6, 0);
            image_loaded = true;
        }
        __syncthreads();
        // the torsion head's per-row extras carry its folded constant
        for (int idx = tid; idx < kN * kHid; idx += kT5Threads) S[M.f.Tt + idx] += S[T.Cvec + F_TOR * kHid + (idx & 63)];
        __syncthreads();

        // ---------------- attention-carrying pairs ----------------
        const int total = L > 0 ? L * W : 0;
        const int npasses = (total + kBwdPairs - 1) / kBwdPairs;
        if (mma_warp) {
            for (int q = 0; q < npasses; ++q) t5_issue_pass(sbase, tmem, bars, np + q);
        } else {
            const int pcol = tid & (kBwdPairs - 1);
            for (int q = 0; q < npasses; ++q) {
                const int pass_base = q * kBwdPairs;
                const int npass = min(kBwdPairs, total - pass_base);
                const bool act = pcol < npass;
                const PairRef pr = decode_full_pair(I, act ? pass_base + pcol : pass_base, W, L, 0, act);
                t5_heads_pass<LAYER>(sb, S, T, g, pr, b, ajt, dajt, I, L, W, pass_base, npass, ci.nv, L - 1, np + q, tmem, gs, inv_gs);
            }
            // ---------------- layer 1: message-only pairs (self, masked peptide / pocket slots) ----------------
            if (LAYER == 0 && L > 0) {
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
        np += npasses;
        __syncthreads();

        if (LAYER == 0) {
            // message_mlp.2 through the message sum: dW2[k][c] += sum_i dMsum[i][k] S1[i][c], db2[k] += (16 + P) sum_i dMsum[i][k]
            float* dW2 = direct + (param_offset(LAYER, MSG2_W) - base);
            for (int idx = tid; idx < kHid * kHid; idx += kT5Threads) {
                const int k = idx >> 6, c = idx & 63;
                const float old = __ldcg(dW2 + idx);
                float acc = 0.0f;
#pragma unroll
                for (int i = 0; i < kN; ++i) acc = fmaf(S[M.dMsum + i * kHid + k], S[T.S1 + i * kHid + c], acc);
                dW2[idx] = old + acc;
            }
            for (int k = tid; k < kHid; k += kT5Threads) {
                float acc = 0.0f;
                for (int r = 0; r < L; ++r) acc += S[M.dMsum + I[IN_ROWS + r] * kHid + k];
                direct[(param_offset(LAYER, MSG2_B) - base) + k] += (float)(kN + P) * acc;
            }
        }
        bwd_node_level<LAYER, kLdt, false>(S, M, g, b, I, dajt, direct, kT5Threads);
        __syncthreads();
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
            uint32_t v[32];
            tc::tmem_ld32_nowait(tlane + (pair == 0 ? TM_DFA : TM_DFB) + 32 * half, v);
            tc::tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) direct[woff + n * ld + 32 * half + k] = __uint_as_float(v[k]) * inv_gs;
            if (half == 0) {
                float w3[16], ex[16];
                tc::tmem_ld16(tlane + (pair == 0 ? TM_DW3A : TM_DW3B), w3);
                tc::tmem_ld16(tlane + (pair == 0 ? TM_EXTA : TM_EXTB), ex);
                if (head == F_ROT) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        direct[(param_offset(LAYER, ROT2_W) - base) + c * kHid + n] = w3[c] * inv_gs;
                        direct[woff + n * ld + 64 + c] = ex[c] * inv_gs;
                    }
                    direct[(param_offset(LAYER, ROT0_B) - base) + n] = ex[6] * inv_gs;
                } else if (head == F_TOR) {
#pragma unroll
                    for (int c = 0; c < PMHC_NTORS; ++c) direct[(param_offset(LAYER, TOR2_W) - base) + c * kHid + n] = w3[8 + c] * inv_gs;
                    direct[(param_offset(LAYER, TOR0_B) - base) + n] = ex[6] * inv_gs;
                } else if (head == F_TRN) {
                    direct[(param_offset(LAYER, TRN2_W) - base) + n] = w3[0] * inv_gs;
                    direct[(param_offset(LAYER, TRN0_B) - base) + n] = ex[6] * inv_gs;
                } else {
                    direct[(param_offset(LAYER, ATT2_W) - base) + n] = w3[8] * inv_gs;
                    direct[woff + n * ld + 64] = ex[4] * inv_gs;
                    direct[woff + n * ld + 65] = ex[5] * inv_gs;
                    direct[(param_offset(LAYER, ATT0_B) - base) + n] = ex[6] * inv_gs;
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
    if (mma_warp) tc::tmem_dealloc(tmem, 512);
}

}  // namespace pmhc
