// egnn_forward_tc.cu — fused EGNN layer forward on the 5th-generation tensor cores (bf16 mode).
//
// Same fusion as egnn_forward.cu (nothing of size [B, N, K, *] touches HBM, one persistent CTA per SM, per-complex
// node data in shared memory, one warp per row for the softmax), but the two dense per-pair contractions — the
// true GEMMs of the reference's layer (message_mlp.2: 64x64; the four head hidden layers on the shared message:
// 64x256; model.py:47-81) — run as tcgen05.mma on tiles of 128 pair rows:
//
//   CUDA cores   m1 = relu(A_i + A_j + W_e)            -> bf16, 128x64 K-major SW128 tile in shared memory
//   tcgen05      D1[128x64]  = m1 . W2^T                (fp32 accumulator in TMEM columns 0..63)
//   CUDA cores   m = D1 + b2 (tcgen05.ld)               -> bf16 tile; layer 1: segmented row sums of m (fp32)
//   tcgen05      D2[128x256] = m . [W_att|W_rot|W_tor|W_trn]^T   (TMEM columns 64..319)
//   CUDA cores   hidden = relu(D2 + fp32 extras: -d2, qdot2, local quaternion, torsions, biases),
//                second layers (64 -> 1, 4, 7, 1) in fp32, sigmoid, quaternion sandwich -> per-pair outputs
//
// bf16 rounding applies ONLY to the operands of those two contractions; geometry, the large-magnitude extra
// inputs, the second layers, softmax and every update stay fp32.  Tolerance of this mode: 1e-2 (north_star).
// Weights stay resident in shared memory as bf16 B operands (40 KB per layer).
#include "egnn_common.cuh"
#include "tcgen05.cuh"

namespace pmhc {

constexpr int kTcThreads = 256;   // 8 warps: warp w and w+4 share TMEM lanes 32*(w%4).., each takes half the columns
constexpr int kTileRows = 128;    // UMMA M
constexpr int kTmemCols = 512;    // D1: 0..63, D2: 64..319
constexpr int kAjLd = 68;         // row stride (floats) of the shared A_j copy: 16-byte aligned rows, 4-bank skew

struct TcMap {
    SmemMap f;           // float-offset fields used by finalize_rows / packs
    int W2b, Whb, A1b, A2b;   // BYTE offsets of the bf16 SW128 tiles (1024-B aligned)
    int WqHb;            // BYTE offset: resident message_mlp.0[:, 0:2H] as bf16 [64][LDQ]
    int TorX;            // float offset: torsion_mlp.0[:, 64:78] + bias as [64][15]
    int AjS;             // float offset of the shared A_j copy [Kpad][kAjLd], or -1 when it does not fit (read from L2)
    int Bar, TmemPtr;    // float offsets of the two mbarriers (4 floats) and the TMEM base address
    int total_bytes;
};

// row length (bf16 elements) of the resident message_mlp.0[:, 0:2H] copy: an odd number of 32-bit words per row
__host__ __device__ constexpr int tc_ldq(int layer) { return layer == 0 ? 50 : 130; }

__host__ __device__ inline TcMap make_tc_map(int Kpad, int layer, bool aj_in_smem) {
    TcMap m;
    m.W2b = 0;
    m.Whb = m.W2b + 64 * 128;
    m.A1b = m.Whb + 256 * 128;
    m.A2b = m.A1b + kTileRows * 128;
    int ob = m.A2b + kTileRows * 128;
    int o = ob / 4;                             // float offsets from here on
    m.f.W2T = m.f.WhT = -1;
    m.f.Scr = m.A1b / 4;                        // scratch for the feature MLP: A1 | A2 | Out (contiguous, free then)
    m.f.Out = o;    o += kCapPairs * kOutPerPair;
    m.f.PkAtt = o;  o += 4 * kHid;
    m.f.PkRotQ = o; o += 4 * kHid;
    m.f.PkRot2 = o; o += 4 * kHid;
    m.f.PkMisc = o; o += 4 * kHid;
    m.f.PkTor2 = o; o += 8 * kHid;
    m.f.Scal = o;   o += 16;
    m.f.We = o;     o += kEdge * kLdN + 1;
    m.f.Ai = o;     o += kN * kLdN;
    o = (o + 3) & ~3;
    m.f.Tt = o;     o += kN * kHid;
    m.f.Msum = o;   o += kN * kHid;
    m.f.H = o;      o += kN * kLdN;
    o = (o + 3) & ~3;
    m.f.Tors = o;   o += kN * 2 * PMHC_NTORS;
    m.f.Q = o;      o += Kpad * 4;
    m.f.X = o;      o += Kpad * 3;
    o = (o + 3) & ~3;
    m.f.Ints = o;   o += Kpad + 64;
    o = (o + 3) & ~3;
    m.TorX = o;     o += kHid * 15;
    o = (o + 3) & ~3;
    m.WqHb = o * 4; o += (kHid * tc_ldq(layer) * 2 + 3) / 4;
    o = (o + 3) & ~3;
    m.AjS = -1;
    if (aj_in_smem) { m.AjS = o; o += Kpad * kAjLd; }
    m.Bar = o;      o += 4;
    m.TmemPtr = o;  o += 4;
    m.f.total_floats = o;
    m.total_bytes = o * 4;
    return m;
}

// one-time staging of the layer's weights: bf16 B operands (K-major, 128-byte swizzle), packs, edge weights,
// the peptide block of message_mlp.0 (bf16) and the torsion-input block of torsion_mlp.0
template <int LAYER>
__device__ inline void stage_weights_tc(uint8_t* smem, const TcMap& M, const float* __restrict__ params) {
    constexpr int L = LAYER;
    constexpr int H = layer_H(L);
    constexpr int ld1 = 2 * H + kEdge;
    constexpr int LDQ = tc_ldq(L);
    const int tid = threadIdx.x;
    float* S = reinterpret_cast<float*>(smem);
    const float* msg0 = params + param_offset(L, MSG0_W);
    const float* msg2 = params + param_offset(L, MSG2_W);
    const float* head[4] = {params + param_offset(L, ATT0_W), params + param_offset(L, ROT0_W),
                            params + param_offset(L, TOR0_W), params + param_offset(L, TRN0_W)};
    const int ldh[4] = {66, 68, 78, 64};
    for (int idx = tid; idx < 64 * 32; idx += kTcThreads) {          // W2: B[n][k] = message_mlp.2.weight[n][k]
        int n = idx >> 5, k = (idx & 31) * 2;
        *reinterpret_cast<uint32_t*>(smem + M.W2b + tc::sw128_offset(n, k)) = tc::pack_bf16x2(msg2[n * 64 + k], msg2[n * 64 + k + 1]);
    }
    for (int idx = tid; idx < 256 * 32; idx += kTcThreads) {         // heads: rows 64h + n, message columns only
        int row = idx >> 5, k = (idx & 31) * 2;
        int h = row >> 6, n = row & 63;
        const float* w = head[h] + n * ldh[h] + k;
        *reinterpret_cast<uint32_t*>(smem + M.Whb + tc::sw128_offset(row, k)) = tc::pack_bf16x2(w[0], w[1]);
    }
    for (int idx = tid; idx < kHid * kEdge; idx += kTcThreads) {
        int k = idx / kEdge, r = idx - k * kEdge;
        S[M.f.We + r * kLdN + k] = msg0[k * ld1 + 2 * H + r];
    }
    __nv_bfloat16* wq = reinterpret_cast<__nv_bfloat16*>(smem + M.WqHb);
    for (int idx = tid; idx < kHid * 2 * H; idx += kTcThreads) {
        int k = idx / (2 * H), c = idx - k * (2 * H);
        wq[k * LDQ + c] = __float2bfloat16_rn(msg0[k * ld1 + c]);
    }
    const float* tor0 = params + param_offset(L, TOR0_W);
    for (int idx = tid; idx < kHid * 14; idx += kTcThreads) {
        int n = idx / 14, c = idx - n * 14;
        S[M.TorX + n * 15 + c] = tor0[n * 78 + 64 + c];
    }
    for (int n = tid; n < kHid; n += kTcThreads) S[M.TorX + n * 15 + 14] = params[param_offset(L, TOR0_B) + n];
    stage_packs<LAYER>(S, M.f, params);
}

// Per-complex setup with the pocket projections cached (pocket_projection_kernel): geometry, torsions, node
// features, lists from the cached slot classes, the pocket rows of A_j copied into shared memory (when they fit),
// A_i / A_j of the 16 peptide slots (register-blocked, bf16 weights, fp32 features and sums) and T_t.
// `ajg` is this complex's [Kpad][64] block of the cache; without the shared copy its peptide rows are rewritten.
template <int LAYER>
__device__ inline ComplexInfo setup_tc(uint8_t* smem, const TcMap& MM, const LayerArgs& a, int b, float* ajg) {
    constexpr int L = LAYER;
    constexpr int H = layer_H(L);
    constexpr int LDQ = tc_ldq(L);
    float* S = reinterpret_cast<float*>(smem);
    const SmemMap& M = MM.f;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.P, K = kN + P, Kpad = a.Kpad;
    int* I = reinterpret_cast<int*>(S + M.Ints);
    const float* msg0b = a.params + param_offset(L, MSG0_B);
    const __nv_bfloat16* wq = reinterpret_cast<const __nv_bfloat16*>(smem + MM.WqHb);
    const float* torx = S + MM.TorX;

    if (MM.AjS >= 0) {
        // pocket rows of A_j: L2 -> shared, 16 floats per row chunk (all loads of a thread in flight together)
        const float4* src = reinterpret_cast<const float4*>(ajg + kN * kHid);
        for (int idx = tid; idx < P * 16; idx += kTcThreads) {
            int j = idx >> 4, q = idx & 15;
            float4 v = __ldcg(src + idx);
            *reinterpret_cast<float4*>(S + MM.AjS + (kN + j) * kAjLd + 4 * q) = v;
        }
    }
    for (int idx = tid; idx < K * 7; idx += kTcThreads) {
        int j = idx / 7, c = idx - j * 7;
        float v = (j < kN) ? a.frames_in[((size_t)b * kN + j) * 7 + c] : a.pocket_frames[((size_t)b * P + (j - kN)) * 7 + c];
        if (c < 4) S[M.Q + j * 4 + c] = v;
        else S[M.X + j * 3 + (c - 4)] = v;
    }
    for (int idx = tid; idx < kN * 14; idx += kTcThreads) S[M.Tors + idx] = a.tors_in[(size_t)b * kN * 14 + idx];
    for (int idx = tid; idx < kN * kHid; idx += kTcThreads) {
        int i = idx >> 6, c = idx & 63;
        float v;
        if (L == 0) v = (c < PMHC_NFEAT) ? a.feat_in[((size_t)b * kN + i) * PMHC_NFEAT + c] : (c == PMHC_NFEAT ? a.t_over_T : 0.0f);
        else v = a.feat_in[((size_t)b * kN + i) * kHid + c];
        S[M.H + i * kLdN + c] = v;
    }
    for (int idx = tid; idx < kN * kHid; idx += kTcThreads) S[M.Msum + idx] = 0.0f;
    if (warp == 0) {
        bool real = lane < kN && a.mask[(size_t)b * kN + lane] != 0;
        unsigned bal = __ballot_sync(0xffffffffu, real);
        int pos = __popc(bal & ((1u << lane) - 1u));
        int Lr = __popc(bal);
        if (lane < kN) {
            if (real) I[IN_ROWS + pos] = lane;
            else I[IN_PEPX + (lane - pos)] = lane;
        }
        int nv = 0, nx = 0, c0 = 0;
        for (int base = 0; base < P; base += 32) {
            int j = base + lane;
            int cls = j < P ? (int)a.pocket_cls[(size_t)b * P + j] : 3;
            unsigned bv = __ballot_sync(0xffffffffu, cls == 0);
            unsigned bx = __ballot_sync(0xffffffffu, cls == 2);
            unsigned bz = __ballot_sync(0xffffffffu, cls == 1);
            if (cls == 0) I[IN_POCKET + nv + __popc(bv & ((1u << lane) - 1u))] = kN + j;
            if (cls == 2) I[IN_POCKET + Kpad - 1 - (nx + __popc(bx & ((1u << lane) - 1u)))] = kN + j;
            nv += __popc(bv);
            nx += __popc(bx);
            c0 += __popc(bz);
        }
        if (lane == 0) {
            I[IN_POCKET + Kpad + 0] = Lr;
            I[IN_POCKET + Kpad + 1] = nv;
            I[IN_POCKET + Kpad + 2] = nx;
            I[IN_POCKET + Kpad + 3] = c0;
        }
    }
    __syncthreads();
    {   // peptide projections: thread = (k, 4 peptide slots)
        const int k = tid & 63, i0 = (tid >> 6) * 4;
        const __nv_bfloat16* w = wq + k * LDQ;
        float ai[4], aj[4];
        const float bias = msg0b[k];
#pragma unroll
        for (int u = 0; u < 4; ++u) { ai[u] = bias; aj[u] = 0.0f; }
#pragma unroll 4
        for (int c = 0; c < H; ++c) {
            const float wi = __bfloat162float(w[c]), wj = __bfloat162float(w[H + c]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float hv = S[M.H + (i0 + u) * kLdN + c];
                ai[u] = fmaf(wi, hv, ai[u]);
                aj[u] = fmaf(wj, hv, aj[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            S[M.Ai + (i0 + u) * kLdN + k] = ai[u];
            if (MM.AjS >= 0) S[MM.AjS + (i0 + u) * kAjLd + k] = aj[u];
            else ajg[(i0 + u) * kHid + k] = aj[u];
        }
    }
    for (int idx = tid; idx < kN * kHid; idx += kTcThreads) {
        int i = idx >> 6, n = idx & 63;
        const float* w = torx + n * 15;
        const float* t = S + M.Tors + i * 14;
        float acc = w[14];
#pragma unroll
        for (int c = 0; c < 14; ++c) acc = fmaf(w[c], t[c], acc);
        S[M.Tt + idx] = acc;
    }
    __syncthreads();
    ComplexInfo ci;
    ci.L = I[IN_POCKET + Kpad + 0];
    ci.nv = I[IN_POCKET + Kpad + 1];
    ci.nx = I[IN_POCKET + Kpad + 2];
    ci.c0 = I[IN_POCKET + Kpad + 3];
    return ci;
}

struct TileCtx {
    uint32_t tmem;            // TMEM base address
    uint32_t phase1, phase2;  // parities of the two mbarriers (MMA 1 / MMA 2 completions)
};

struct TcTile {
    uint8_t* smem;
    const TcMap& M;
    const LayerArgs& a;
    const float* ajg;   // this complex's cache block (used when the shared A_j copy is absent)

    __device__ __forceinline__ float* S() const { return reinterpret_cast<float*>(smem); }
    __device__ __forceinline__ uint64_t* bar(int which) const { return reinterpret_cast<uint64_t*>(S() + M.Bar) + which; }

    // m1 = relu(A_i + A_j + W_e) -> bf16 A1 tile (this thread: row tid&127, 32 of the 64 columns); then publish
    __device__ __forceinline__ void stage_a1(const PairRef& pr) const {
        const int tid = threadIdx.x, r = tid & 127, h = tid >> 7;
        const int i = pr.i, j = pr.j;
        const bool pep = (j >= 0 && j < kN);
        const float* Sf = S();
        const float* ai = Sf + M.f.Ai + i * kLdN + 32 * h;
        const float* we = Sf + M.f.We + (pep ? (kN - 1 + i - j) : 0) * kLdN + 32 * h;
        float4 aj[8];
        if (j >= 0) {
            if (M.AjS >= 0) {
                const float4* src = reinterpret_cast<const float4*>(Sf + M.AjS + j * kAjLd + 32 * h);
#pragma unroll
                for (int q = 0; q < 8; ++q) aj[q] = src[q];
            } else {
                const float4* src = reinterpret_cast<const float4*>(ajg + (size_t)j * kHid + 32 * h);
#pragma unroll
                for (int q = 0; q < 8; ++q) aj[q] = __ldcg(src + q);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) aj[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 lo = aj[2 * q], hi = aj[2 * q + 1];
            float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                v[e] += ai[8 * q + e];
                if (pep) v[e] += we[8 * q + e];
                v[e] = fmaxf(v[e], 0.0f);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) pk[e] = tc::pack_bf16x2(v[2 * e], v[2 * e + 1]);
            *reinterpret_cast<uint4*>(smem + M.A1b + tc::sw128_offset(r, 32 * h + 8 * q)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        tc::fence_proxy_async_smem();
        tc::fence_before_thread_sync();
    }

    // D1[128x64] = A1 . W2^T (one thread), completion -> mbarrier 0
    __device__ __forceinline__ void issue_mma1(const TileCtx& ctx) const {
        if (threadIdx.x == 0) {
            tc::fence_after_thread_sync();
            const uint64_t da = tc::smem_desc_sw128(tc::smem_u32(smem + M.A1b));
            const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + M.W2b));
            constexpr uint32_t idesc = tc::idesc_bf16_f32(128, 64);
#pragma unroll
            for (int s = 0; s < 4; ++s) tc::mma_bf16(ctx.tmem, da + 2 * s, db + 2 * s, idesc, s > 0);  // +32 B per K = 16 step
            tc::mma_commit(bar(0));
        }
    }
    // D2[128x256] = A2 . [W_att | W_rot | W_tor | W_trn]^T, completion -> mbarrier 1
    __device__ __forceinline__ void issue_mma2(const TileCtx& ctx) const {
        if (threadIdx.x == 0) {
            tc::fence_after_thread_sync();
            const uint64_t da = tc::smem_desc_sw128(tc::smem_u32(smem + M.A2b));
            const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + M.Whb));
            constexpr uint32_t idesc = tc::idesc_bf16_f32(128, 256);
#pragma unroll
            for (int s = 0; s < 4; ++s) tc::mma_bf16(ctx.tmem + 64, da + 2 * s, db + 2 * s, idesc, s > 0);
            tc::mma_commit(bar(1));
        }
    }

    // epilogue 1: m = D1 + b2 ; layer 1: unmasked message sums ; bf16 A2 tile for the heads (HEADS) ; then publish
    template <int LAYER, bool HEADS>
    __device__ __forceinline__ void epilogue1(const TileCtx& ctx, const PairRef& pr, float mult) const {
        const int tid = threadIdx.x, lane = tid & 31, r = tid & 127, h = tid >> 7;
        const uint32_t lane_base = (uint32_t)(((tid >> 5) & 3) * 32) << 16;
        float* Sf = S();
        const int i = pr.i;
        const bool act = pr.active;
        float m[32];
        tc::tmem_ld32(ctx.tmem + lane_base + 32 * h, m);
#pragma unroll
        for (int c = 0; c < 32; ++c) m[c] += Sf[M.f.PkMisc + 4 * (32 * h + c) + 3];
        if (HEADS) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) pk[e] = tc::pack_bf16x2(m[8 * q + 2 * e], m[8 * q + 2 * e + 1]);
                *reinterpret_cast<uint4*>(smem + M.A2b + tc::sw128_offset(r, 32 * h + 8 * q)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
        if (LAYER == 0) {
            // unmasked message sums (model.py:151): rows of one peptide residue are contiguous, so a segmented
            // shuffle reduction leaves each segment's sum in its first lane; the four row-warps then add their
            // segment sums in a fixed order (plain adds, no atomics) -> bit-reproducible, shard-independent
            const int seg = act ? i : -1;
            int seg_at[5];
#pragma unroll
            for (int s = 0; s < 5; ++s) seg_at[s] = __shfl_down_sync(0xffffffffu, seg, 1 << s);
            const int seg_prev = __shfl_up_sync(0xffffffffu, seg, 1);
            const bool head_lane = act && (lane == 0 || seg_prev != seg);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float v = act ? m[c] * mult : 0.0f;
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    const float o = __shfl_down_sync(0xffffffffu, v, 1 << s);
                    if (lane + (1 << s) < 32 && seg_at[s] == seg) v += o;
                }
                m[c] = v;
            }
#pragma unroll 1
            for (int wq = 0; wq < 4; ++wq) {
                if (((tid >> 5) & 3) == wq && head_lane) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) Sf[M.f.Msum + i * kHid + 32 * h + c] += m[c];
                }
                __syncthreads();
            }
        }
        tc::fence_proxy_async_smem();
        tc::fence_before_thread_sync();
    }

    // epilogue 2: fp32 extras + relu + second layers + geometry -> per-pair outputs
    __device__ __forceinline__ void epilogue2(const TileCtx& ctx, const PairRef& pr, int out_slot, float* __restrict__ lsave) const {
        const int tid = threadIdx.x, h = tid >> 7;
        const uint32_t lane_base = (uint32_t)(((tid >> 5) & 3) * 32) << 16;
        float* Sf = S();
        const int i = pr.i, j = pr.j;
        const bool act = pr.active;
        float* out = Sf + M.f.Out + out_slot * kOutPerPair;
        const float* pqi = Sf + M.f.Q + i * 4;
        const float* pqj = Sf + M.f.Q + j * 4;
        const Quat qi{pqi[0], pqi[1], pqi[2], pqi[3]}, qj{pqj[0], pqj[1], pqj[2], pqj[3]};
        const float rx = Sf[M.f.X + i * 3] - Sf[M.f.X + j * 3], ry = Sf[M.f.X + i * 3 + 1] - Sf[M.f.X + j * 3 + 1],
                    rz = Sf[M.f.X + i * 3 + 2] - Sf[M.f.X + j * 3 + 2];
        float v[32];
        if (h == 0) {
            // attention logit (model.py:238-242): large-magnitude -d2 / qdot2 inputs enter in fp32 after the contraction
            const float d2 = rx * rx + ry * ry + rz * rz;
            const float dq = qdot(qi, qj);
            const float qd = dq * dq;
            float lg0 = Sf[M.f.Scal + SC_ATT2B], lg1 = 0.0f;   // two chains: halves the dependent-FMA latency
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                tc::tmem_ld32(ctx.tmem + lane_base + 64 + 32 * half, v);
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const float4 p0 = *reinterpret_cast<const float4*>(Sf + M.f.PkAtt + 4 * (32 * half + c));
                    const float4 p1 = *reinterpret_cast<const float4*>(Sf + M.f.PkAtt + 4 * (32 * half + c + 1));
                    const float h0 = (v[c] + p0.z) + fmaf(p0.y, qd, p0.x * -d2);
                    const float h1 = (v[c + 1] + p1.z) + fmaf(p1.y, qd, p1.x * -d2);
                    lg0 = fmaf(p0.w, fmaxf(h0, 0.0f), lg0);
                    lg1 = fmaf(p1.w, fmaxf(h1, 0.0f), lg1);
                }
            }
            const float logit = lg0 + lg1;
            // rotation (model.py:283-296)
            const Quat qinvj = qinv(qj);
            const Quat lq = qmul(qinvj, qmul(qi, qj));
            float pre[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) pre[c] = Sf[M.f.Scal + SC_ROT2B + c];
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                tc::tmem_ld32(ctx.tmem + lane_base + 64 + 64 + 32 * half, v);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int n = 32 * half + c;
                    const float4 wq = *reinterpret_cast<const float4*>(Sf + M.f.PkRotQ + 4 * n);
                    float hid = v[c] + Sf[M.f.PkMisc + 4 * n + 2];
                    hid += wq.x * lq.w + wq.y * lq.x + wq.z * lq.y + wq.w * lq.z;
                    hid = fmaxf(hid, 0.0f);
                    const float4 w2 = *reinterpret_cast<const float4*>(Sf + M.f.PkRot2 + 4 * n);
                    pre[0] = fmaf(w2.x, hid, pre[0]); pre[1] = fmaf(w2.y, hid, pre[1]);
                    pre[2] = fmaf(w2.z, hid, pre[2]); pre[3] = fmaf(w2.w, hid, pre[3]);
                }
            }
            const Quat dl{sigmoidf(pre[0]), sigmoidf(pre[1]), sigmoidf(pre[2]), sigmoidf(pre[3])};  // never normalised (T5)
            const Quat dg = qmul(qj, qmul(dl, qinvj));
            if (act) {
                out[0] = logit;
                out[1] = dg.w; out[2] = dg.x; out[3] = dg.y; out[4] = dg.z;
                if (lsave != nullptr) lsave[i * a.Kpad + j] = logit;
            }
        } else {
            // torsion increments (model.py:257-260)
            float da[PMHC_NTORS];
#pragma unroll
            for (int c = 0; c < PMHC_NTORS; ++c) da[c] = Sf[M.f.Scal + SC_TOR2B + c];
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                tc::tmem_ld32(ctx.tmem + lane_base + 64 + 128 + 32 * half, v);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int n = 32 * half + c;
                    const float hid = fmaxf(v[c] + Sf[M.f.Tt + i * kHid + n], 0.0f);
                    const float4 w0 = *reinterpret_cast<const float4*>(Sf + M.f.PkTor2 + 8 * n);
                    const float4 w1 = *reinterpret_cast<const float4*>(Sf + M.f.PkTor2 + 8 * n + 4);
                    da[0] = fmaf(w0.x, hid, da[0]); da[1] = fmaf(w0.y, hid, da[1]); da[2] = fmaf(w0.z, hid, da[2]);
                    da[3] = fmaf(w0.w, hid, da[3]); da[4] = fmaf(w1.x, hid, da[4]); da[5] = fmaf(w1.y, hid, da[5]);
                    da[6] = fmaf(w1.z, hid, da[6]);
                }
            }
            // translation scale (model.py:325-331)
            float sc0 = Sf[M.f.Scal + SC_TRN2B], sc1 = 0.0f;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                tc::tmem_ld32(ctx.tmem + lane_base + 64 + 192 + 32 * half, v);
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const int n = 32 * half + c;
                    sc0 = fmaf(Sf[M.f.PkMisc + 4 * n + 1], fmaxf(v[c] + Sf[M.f.PkMisc + 4 * n], 0.0f), sc0);
                    sc1 = fmaf(Sf[M.f.PkMisc + 4 * n + 5], fmaxf(v[c + 1] + Sf[M.f.PkMisc + 4 * n + 4], 0.0f), sc1);
                }
            }
            const float sc = sc0 + sc1;
            if (act) {
#pragma unroll
                for (int c = 0; c < PMHC_NTORS; ++c) out[5 + c] = da[c];
                out[12] = sc * rx; out[13] = sc * ry; out[14] = sc * rz;
            }
        }
        tc::fence_before_thread_sync();
    }
};

template <int LAYER>
__global__ void __launch_bounds__(kTcThreads, 1) egnn_layer_forward_tc_kernel(LayerArgs a, int aj_in_smem) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need a 1024-byte aligned base in the shared window: round up (1 KB slack is allocated)
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    float* S = reinterpret_cast<float*>(smem);
    const TcMap M = make_tc_map(a.Kpad, LAYER, aj_in_smem != 0);
    const int tid = threadIdx.x, warp = tid >> 5;
    int* I = reinterpret_cast<int*>(S + M.f.Ints);
    uint64_t* bars = reinterpret_cast<uint64_t*>(S + M.Bar);

    if (warp == 0) tc::tmem_alloc(reinterpret_cast<uint32_t*>(S + M.TmemPtr), kTmemCols);
    if (tid == 32) {
        tc::mbar_init(bars + 0, 1);
        tc::mbar_init(bars + 1, 1);
        tc::mbar_fence_init();
    }
    stage_weights_tc<LAYER>(smem, M, a.params);
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    TileCtx ctx;
    ctx.tmem = *reinterpret_cast<volatile uint32_t*>(S + M.TmemPtr);
    ctx.phase1 = ctx.phase2 = 0;

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        float* ajg = a.ajt_cache + ((size_t)b * 2 + LAYER) * kHid * a.Kpad;
        const ComplexInfo ci = setup_tc<LAYER>(smem, M, a, b, ajg);
        const TcTile T{smem, M, a, ajg};
        const int L = ci.L;
        const int W = (L - 1) + ci.nv;
        float* lsave = a.logit_out ? a.logit_out + (size_t)b * kN * a.Kpad : nullptr;
        for (int idx = tid; idx < (kN - L) * 21; idx += kTcThreads) {   // padded rows: pass-through (T4)
            int s = idx / 21, c = idx - s * 21;
            int i = I[IN_PEPX + s];
            if (c < 7) a.frames_out[((size_t)b * kN + i) * 7 + c] = a.frames_in[((size_t)b * kN + i) * 7 + c];
            else a.tors_out[((size_t)b * kN + i) * 14 + (c - 7)] = a.tors_in[((size_t)b * kN + i) * 14 + (c - 7)];
        }

        const int rows_per_group = W > 0 ? max(1, kCapPairs / W) : kN;
        for (int row0 = 0; row0 < L; row0 += rows_per_group) {
            const int nrows = min(rows_per_group, L - row0);
            const int gpairs = nrows * W;
            const int ntiles = (gpairs + kTileRows - 1) / kTileRows;
            // Software pipeline over the tiles of this row group: while the tensor core runs MMA 2 of tile t, the
            // CUDA cores stage the A operand of tile t+1 and MMA 1 of tile t+1 runs under epilogue 2 of tile t.
            auto pair_of = [&](int t) {
                const int gp = t * kTileRows + (tid & 127);
                const bool act = gp < gpairs;
                return decode_full_pair(I, act ? gp : t * kTileRows, W, L, row0, act);
            };
            PairRef cur = pair_of(0);
            if (ntiles > 0) {
                T.stage_a1(cur);
                __syncthreads();
                T.issue_mma1(ctx);
            }
            for (int t = 0; t < ntiles; ++t) {
                tc::mbar_wait(bars + 0, ctx.phase1);
                ctx.phase1 ^= 1;
                tc::fence_after_thread_sync();
                T.template epilogue1<LAYER, true>(ctx, cur, 1.0f);
                __syncthreads();
                T.issue_mma2(ctx);
                PairRef nxt = cur;
                if (t + 1 < ntiles) {
                    nxt = pair_of(t + 1);
                    T.stage_a1(nxt);            // A1 is free: MMA 1 of tile t has completed
                    __syncthreads();
                    T.issue_mma1(ctx);
                }
                tc::mbar_wait(bars + 1, ctx.phase2);
                ctx.phase2 ^= 1;
                tc::fence_after_thread_sync();
                T.epilogue2(ctx, cur, t * kTileRows + (tid & 127), lsave);
                cur = nxt;
            }
            __syncthreads();
            finalize_rows(S, M.f, a, I, b, row0, nrows, W);
            __syncthreads();
        }

        if (LAYER == 0) {
            // message-only pairs: self, masked peptide slots, masked pocket slots (one shared zero-feature message)
            const int npx = kN - L;
            const int W2 = 1 + npx + ci.nx + (ci.c0 > 0 ? 1 : 0);
            const int total = L * W2;
            for (int tile_base = 0; tile_base < total; tile_base += kTileRows) {
                const int gp0 = tile_base + (tid & 127);
                const bool act = gp0 < total;
                const int gp = act ? gp0 : tile_base;
                const int rl = gp / W2, e = gp - rl * W2;
                PairRef pr;
                pr.i = I[IN_ROWS + rl];
                pr.active = act;
                float mult = 1.0f;
                if (e == 0) pr.j = pr.i;
                else if (e <= npx) pr.j = I[IN_PEPX + e - 1];
                else if (e <= npx + ci.nx) pr.j = I[IN_POCKET + a.Kpad - 1 - (e - npx - 1)];
                else { pr.j = -1; mult = (float)ci.c0; }
                T.stage_a1(pr);
                __syncthreads();
                T.issue_mma1(ctx);
                tc::mbar_wait(bars + 0, ctx.phase1);
                ctx.phase1 ^= 1;
                tc::fence_after_thread_sync();
                T.template epilogue1<LAYER, false>(ctx, pr, mult);
                __syncthreads();
            }

            // node feature update: relu(feature_mlp(cat(h_i, sum_j m_ij))) (model.py:151, :407), fp32
            const float* f0w = a.params + param_offset(0, FEAT0_W);
            const float* f0b = a.params + param_offset(0, FEAT0_B);
            const float* f2w = a.params + param_offset(0, FEAT2_W);
            const float* f2b = a.params + param_offset(0, FEAT2_B);
            constexpr int ldf = kH1 + kHid;
            float* hid = S + M.f.Scr;
            float* sf0 = hid + kN * kLdN;
            float* sf2 = sf0 + kHid * ldf;
            for (int idx = tid; idx < kHid * ldf; idx += kTcThreads) sf0[idx] = f0w[idx];
            for (int idx = tid; idx < kHid * kHid; idx += kTcThreads) sf2[(idx >> 6) * kLdN + (idx & 63)] = f2w[idx];
            __syncthreads();
            for (int idx = tid; idx < L * kHid; idx += kTcThreads) {
                int r = idx >> 6, n = idx & 63;
                int i = I[IN_ROWS + r];
                const float* w = sf0 + n * ldf;
                const float* hh = S + M.f.H + i * kLdN;
                const float* ms = S + M.f.Msum + i * kHid;
                float acc = f0b[n];
#pragma unroll
                for (int c = 0; c < kH1; ++c) acc = fmaf(w[c], hh[c], acc);
#pragma unroll 8
                for (int c = 0; c < kHid; ++c) acc = fmaf(w[kH1 + c], ms[c], acc);
                hid[r * kLdN + n] = fmaxf(acc, 0.0f);
                if (a.msum_out != nullptr) a.msum_out[((size_t)b * kN + i) * kHid + n] = ms[n];
            }
            __syncthreads();
            for (int idx = tid; idx < kN * kHid; idx += kTcThreads) {
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

    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(ctx.tmem, kTmemCols);
}

template <int LAYER>
int launch_layer_forward_tc(const LayerArgs& a, cudaStream_t stream) {
    static bool configured = false;
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // keep the complex's neighbour projections A_j in shared memory when they fit next to everything else
    bool aj_smem = (size_t)make_tc_map(a.Kpad, LAYER, true).total_bytes + 1024 <= (size_t)max_smem;
    const TcMap M = make_tc_map(a.Kpad, LAYER, aj_smem);
    const size_t smem = (size_t)M.total_bytes + 1024;   // slack so the tile base can be rounded up to 1024 B
    PMHC_REQUIRE((int)smem <= max_smem, "EGNN tensor-core layer needs %zu B of shared memory (P=%d), device allows %d", smem, a.P, max_smem);
    PMHC_REQUIRE(a.ajt_cache != nullptr && a.pocket_cls != nullptr, "tensor-core layer needs the pocket projection cache");
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(egnn_layer_forward_tc_kernel<LAYER>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        PMHC_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(forward_tc): %s", cudaGetErrorString(e));
        configured = true;
    }
    const int grid = a.B < num_sms() ? a.B : num_sms();
    if (profile_enabled()) profile_mark(PROF_FWD, stream, true);
    egnn_layer_forward_tc_kernel<LAYER><<<grid, kTcThreads, smem, stream>>>(a, aj_smem ? 1 : 0);
    if (profile_enabled()) profile_mark(PROF_FWD, stream, false);
    PMHC_CHECK_LAUNCH("egnn_layer_forward_tc");
    return 0;
}

// Step-invariant pocket side of message_mlp.0 for both layers (model.py:401, 411-412: pocket nodes carry no time
// feature and layer 2 sees the same 22 features zero-padded), computed once per batch / sampling trajectory:
//   ajt_cache[b][l][16 + p][k] = W1_l[k, H_l : H_l + 22] . pocket_features[b][p]      (zero beyond P; rows 0..15 = peptide)
//   cls[b][p] = 0 valid, 1 masked + all-zero features (shares one message), 2 masked + non-zero features
__global__ void pocket_projection_kernel(const float* __restrict__ params, const float* __restrict__ pocket_feat,
                                         const uint8_t* __restrict__ pocket_mask, int P, int Kpad,
                                         float* __restrict__ ajt_cache, uint8_t* __restrict__ cls) {
    extern __shared__ __align__(16) float sp[];
    const int b = blockIdx.x, layer = blockIdx.y, tid = threadIdx.x;
    const int H = layer == 0 ? kH1 : kH2, ld1 = 2 * H + kEdge;
    constexpr int FS = 23;
    float* feat = sp;
    float* w = sp + ((P * FS + 3) & ~3);
    const float* msg0 = params + (layer == 0 ? param_offset(0, MSG0_W) : param_offset(1, MSG0_W));
    for (int idx = tid; idx < P * PMHC_NFEAT; idx += blockDim.x) {
        int j = idx / PMHC_NFEAT, c = idx - j * PMHC_NFEAT;
        feat[j * FS + c] = pocket_feat[(size_t)b * P * PMHC_NFEAT + idx];
    }
    for (int idx = tid; idx < kHid * PMHC_NFEAT; idx += blockDim.x) {
        int k = idx / PMHC_NFEAT, c = idx - k * PMHC_NFEAT;
        w[k * FS + c] = msg0[k * ld1 + H + c];
    }
    __syncthreads();
    float* out = ajt_cache + ((size_t)b * 2 + layer) * kHid * Kpad;
    const int ncol = Kpad - kN;
    for (int idx = tid; idx < (kHid / 4) * ncol; idx += blockDim.x) {
        int kq = idx / ncol, pj = idx - kq * ncol;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        if (pj < P) {
            const float* h = feat + pj * FS;
            const float* ww = w + (4 * kq) * FS;
#pragma unroll
            for (int c = 0; c < PMHC_NFEAT; ++c) {
                float hv = h[c];
                a0 = fmaf(ww[c], hv, a0);
                a1 = fmaf(ww[FS + c], hv, a1);
                a2 = fmaf(ww[2 * FS + c], hv, a2);
                a3 = fmaf(ww[3 * FS + c], hv, a3);
            }
        }
        *reinterpret_cast<float4*>(out + (size_t)(kN + pj) * kHid + 4 * kq) = make_float4(a0, a1, a2, a3);
    }
    if (layer == 0) {
        for (int j = tid; j < P; j += blockDim.x) {
            uint8_t c = 0;
            if (pocket_mask[(size_t)b * P + j] == 0) {
                bool nz = false;
                for (int q = 0; q < PMHC_NFEAT; ++q) nz |= (feat[j * FS + q] != 0.0f);
                c = nz ? 2 : 1;
            }
            cls[(size_t)b * P + j] = c;
        }
    }
}

int launch_pocket_projection(const LayerArgs& a, cudaStream_t stream) {
    const size_t smem = (size_t)(((a.P * 23 + 3) & ~3) + kHid * 23) * sizeof(float);
    static bool configured = false;
    if (!configured && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(pocket_projection_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        PMHC_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(pocket_projection): %s", cudaGetErrorString(e));
        configured = true;
    }
    pocket_projection_kernel<<<dim3(a.B, 2), 128, smem, stream>>>(a.params, a.pocket_feat, a.pocket_mask, a.P, a.Kpad, a.ajt_cache,
                                                                 const_cast<uint8_t*>(a.pocket_cls));
    PMHC_CHECK_LAUNCH("pocket_projection");
    return 0;
}

int launch_layer_forward_tc_layer(int layer, const LayerArgs& a, cudaStream_t stream) {
    return layer == 0 ? launch_layer_forward_tc<0>(a, stream) : launch_layer_forward_tc<1>(a, stream);
}

}  // namespace pmhc
