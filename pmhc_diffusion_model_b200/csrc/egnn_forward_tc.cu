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

struct TcMap {
    SmemMap f;           // float-offset fields used by setup_complex / finalize_rows / packs
    int W2b, Whb, A1b, A2b;   // BYTE offsets of the bf16 SW128 tiles (1024-B aligned)
    int Wq, TorX;        // float offsets: resident message_mlp.0[:, 0:2H] as [64][2H+1] and torsion_mlp.0[:, 64:78]+bias as [64][15]
    int Bar, TmemPtr;    // float offsets of the mbarrier (2 floats) and the TMEM base address
    int total_bytes;
};

__host__ __device__ inline TcMap make_tc_map(int Kpad, int layer) {
    TcMap m;
    m.W2b = 0;
    m.Whb = m.W2b + 64 * 128;
    m.A1b = m.Whb + 256 * 128;
    m.A2b = m.A1b + kTileRows * 128;
    int o = (m.A2b + kTileRows * 128) / 4;      // float offsets from here on
    m.f.W2T = m.f.WhT = -1;
    m.f.Scr = m.A1b / 4;                        // staging area of setup_complex: A1 | A2 | Out (contiguous, free then)
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
    m.Wq = o;       o += kHid * (2 * layer_H(layer) + 1);
    m.TorX = o;     o += kHid * 15;
    o = (o + 3) & ~3;
    m.Bar = o;      o += 4;
    m.TmemPtr = o;  o += 4;
    m.f.total_floats = o;
    m.total_bytes = o * 4;
    return m;
}

// one-time staging of the layer's weights: bf16 B operands (K-major, 128-byte swizzle), packs, edge weights
template <int LAYER>
__device__ inline void stage_weights_tc(uint8_t* smem, const TcMap& M, const float* __restrict__ params) {
    constexpr int L = LAYER;
    constexpr int H = layer_H(L);
    constexpr int ld1 = 2 * H + kEdge;
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
    constexpr int LDQ = 2 * H + 1;
    for (int idx = tid; idx < kHid * 2 * H; idx += kTcThreads) {
        int k = idx / (2 * H), c = idx - k * (2 * H);
        S[M.Wq + k * LDQ + c] = msg0[k * ld1 + c];
    }
    const float* tor0 = params + param_offset(L, TOR0_W);
    for (int idx = tid; idx < kHid * 14; idx += kTcThreads) {
        int n = idx / 14, c = idx - n * 14;
        S[M.TorX + n * 15 + c] = tor0[n * 78 + 64 + c];
    }
    for (int n = tid; n < kHid; n += kTcThreads) S[M.TorX + n * 15 + 14] = params[param_offset(L, TOR0_B) + n];
    stage_packs<LAYER>(S, M.f, params);
}

struct TileCtx {
    uint32_t tmem;       // TMEM base address
    uint32_t phase;      // mbarrier parity to wait for next
};

// Runs one tile of up to 128 pairs.  HEADS: attention-carrying pairs (both GEMMs + head epilogue -> Out buffer);
// otherwise message-only pairs of layer 1 (first GEMM only, contributes to the unmasked message sums).
template <int LAYER, bool HEADS>
__device__ __forceinline__ void run_tile(uint8_t* smem, const TcMap& M, const LayerArgs& a, TileCtx& ctx,
                                         const float* __restrict__ ajt, const PairRef pr, float mult, int out_slot,
                                         float* __restrict__ lsave) {
    float* S = reinterpret_cast<float*>(smem);
    const int tid = threadIdx.x, lane = tid & 31;
    const int r = tid & 127, h = tid >> 7;            // tile row, column half
    const uint32_t lane_base = (uint32_t)(((tid >> 5) & 3) * 32) << 16;
    uint64_t* bar = reinterpret_cast<uint64_t*>(S + M.Bar);
    const int i = pr.i, j = pr.j;
    const bool act = pr.active;
    const bool pep = (j >= 0 && j < kN);
    const int Kpad = a.Kpad;

    // ---------------- S1: m1 = relu(A_i + A_j + W_e) -> bf16 A tile (this thread: 32 of the 64 columns) ----------------
    {
        float aj[32];
        const float* ajc = ajt + (size_t)(32 * h) * Kpad + (j >= 0 ? j : 0);
#pragma unroll
        for (int k = 0; k < 32; ++k) aj[k] = __ldcg(ajc + k * Kpad);
        const float* ai = S + M.f.Ai + i * kLdN + 32 * h;
        const float* we = S + M.f.We + (pep ? (kN - 1 + i - j) : 0) * kLdN + 32 * h;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k0 = 8 * q + 2 * e;
                float v0 = ai[k0], v1 = ai[k0 + 1];
                if (j >= 0) { v0 += aj[k0]; v1 += aj[k0 + 1]; }
                if (pep) { v0 += we[k0]; v1 += we[k0 + 1]; }
                pk[e] = tc::pack_bf16x2(fmaxf(v0, 0.0f), fmaxf(v1, 0.0f));
            }
            *reinterpret_cast<uint4*>(smem + M.A1b + tc::sw128_offset(r, 32 * h + 8 * q)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    // ---------------- MMA 1: D1[128x64] = A1 . W2^T ----------------
    if (tid == 0) {
        tc::fence_after_thread_sync();
        const uint64_t da = tc::smem_desc_sw128(tc::smem_u32(smem + M.A1b));
        const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + M.W2b));
        constexpr uint32_t idesc = tc::idesc_bf16_f32(128, 64);
#pragma unroll
        for (int s = 0; s < 4; ++s) tc::mma_bf16(ctx.tmem, da + 2 * s, db + 2 * s, idesc, s > 0);  // +32 B per K = 16 step
        tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, ctx.phase);
    ctx.phase ^= 1;
    tc::fence_after_thread_sync();

    // ---------------- epilogue 1: m = D1 + b2 ; layer 1: row sums ; bf16 A tile for the heads ----------------
    {
        float m[32];
        tc::tmem_ld32(ctx.tmem + lane_base + 32 * h, m);
#pragma unroll
        for (int c = 0; c < 32; ++c) m[c] += S[M.f.PkMisc + 4 * (32 * h + c) + 3];
        float red[32];
        bool head_lane = false;
        if (LAYER == 0) {
            // unmasked message sums (model.py:151): rows of one peptide residue are contiguous, so a segmented
            // shuffle reduction leaves each segment's sum in its first lane
            const int seg = act ? i : -1;
            int seg_at[5];
#pragma unroll
            for (int s = 0; s < 5; ++s) seg_at[s] = __shfl_down_sync(0xffffffffu, seg, 1 << s);
            const int seg_prev = __shfl_up_sync(0xffffffffu, seg, 1);
            head_lane = act && (lane == 0 || seg_prev != seg);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float v = act ? m[c] * mult : 0.0f;
#pragma unroll
                for (int s = 0; s < 5; ++s) {
                    const float o = __shfl_down_sync(0xffffffffu, v, 1 << s);
                    if (lane + (1 << s) < 32 && seg_at[s] == seg) v += o;
                }
                red[c] = v;
            }
        }
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
            // the four row-warps add their segment sums in a fixed order (plain adds, no atomics): the message sums,
            // hence everything downstream, are bit-reproducible from run to run and independent of batch sharding
#pragma unroll 1
            for (int wq = 0; wq < 4; ++wq) {
                if (((tid >> 5) & 3) == wq && head_lane) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) S[M.f.Msum + i * kHid + 32 * h + c] += red[c];
                }
                __syncthreads();
            }
        }
    }
    if (!HEADS) {
        tc::fence_before_thread_sync();
        return;
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    // ---------------- MMA 2: D2[128x256] = A2 . [W_att | W_rot | W_tor | W_trn]^T ----------------
    if (tid == 0) {
        tc::fence_after_thread_sync();
        const uint64_t da = tc::smem_desc_sw128(tc::smem_u32(smem + M.A2b));
        const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + M.Whb));
        constexpr uint32_t idesc = tc::idesc_bf16_f32(128, 256);
#pragma unroll
        for (int s = 0; s < 4; ++s) tc::mma_bf16(ctx.tmem + 64, da + 2 * s, db + 2 * s, idesc, s > 0);
        tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, ctx.phase);
    ctx.phase ^= 1;
    tc::fence_after_thread_sync();

    // ---------------- epilogue 2: fp32 extras + relu + second layers + geometry -> per-pair outputs ----------------
    float* out = S + M.f.Out + out_slot * kOutPerPair;
    const float* pqi = S + M.f.Q + i * 4;
    const float* pqj = S + M.f.Q + j * 4;
    const Quat qi{pqi[0], pqi[1], pqi[2], pqi[3]}, qj{pqj[0], pqj[1], pqj[2], pqj[3]};
    const float rx = S[M.f.X + i * 3] - S[M.f.X + j * 3], ry = S[M.f.X + i * 3 + 1] - S[M.f.X + j * 3 + 1],
                rz = S[M.f.X + i * 3 + 2] - S[M.f.X + j * 3 + 2];
    float v[32];
    if (h == 0) {
        // attention logit (model.py:238-242): large-magnitude -d2 / qdot2 inputs enter in fp32 after the contraction
        const float d2 = rx * rx + ry * ry + rz * rz;
        const float dq = qdot(qi, qj);
        const float qd = dq * dq;
        float logit = S[M.f.Scal + SC_ATT2B];
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {  // rolled: keeps the epilogue inside the instruction cache
            tc::tmem_ld32(ctx.tmem + lane_base + 64 + 32 * half, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float4 pk = *reinterpret_cast<const float4*>(S + M.f.PkAtt + 4 * (32 * half + c));
                const float hid = (v[c] + pk.z) + fmaf(pk.y, qd, pk.x * -d2);
                logit = fmaf(pk.w, fmaxf(hid, 0.0f), logit);
            }
        }
        // rotation (model.py:283-296)
        const Quat qinvj = qinv(qj);
        const Quat lq = qmul(qinvj, qmul(qi, qj));
        float pre[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) pre[c] = S[M.f.Scal + SC_ROT2B + c];
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {  // rolled: keeps the epilogue inside the instruction cache
            tc::tmem_ld32(ctx.tmem + lane_base + 64 + 64 + 32 * half, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int n = 32 * half + c;
                const float4 wq = *reinterpret_cast<const float4*>(S + M.f.PkRotQ + 4 * n);
                float hid = v[c] + S[M.f.PkMisc + 4 * n + 2];
                hid += wq.x * lq.w + wq.y * lq.x + wq.z * lq.y + wq.w * lq.z;
                hid = fmaxf(hid, 0.0f);
                const float4 w2 = *reinterpret_cast<const float4*>(S + M.f.PkRot2 + 4 * n);
                pre[0] = fmaf(w2.x, hid, pre[0]); pre[1] = fmaf(w2.y, hid, pre[1]);
                pre[2] = fmaf(w2.z, hid, pre[2]); pre[3] = fmaf(w2.w, hid, pre[3]);
            }
        }
        const Quat dl{sigmoidf(pre[0]), sigmoidf(pre[1]), sigmoidf(pre[2]), sigmoidf(pre[3])};  // never normalised (T5)
        const Quat dg = qmul(qj, qmul(dl, qinvj));
        if (act) {
            out[0] = logit;
            out[1] = dg.w; out[2] = dg.x; out[3] = dg.y; out[4] = dg.z;
            if (lsave != nullptr) lsave[i * Kpad + j] = logit;
        }
    } else {
        // torsion increments (model.py:257-260)
        float da[PMHC_NTORS];
#pragma unroll
        for (int c = 0; c < PMHC_NTORS; ++c) da[c] = S[M.f.Scal + SC_TOR2B + c];
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {  // rolled: keeps the epilogue inside the instruction cache
            tc::tmem_ld32(ctx.tmem + lane_base + 64 + 128 + 32 * half, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int n = 32 * half + c;
                const float hid = fmaxf(v[c] + S[M.f.Tt + i * kHid + n], 0.0f);
                const float4 w0 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n);
                const float4 w1 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n + 4);
                da[0] = fmaf(w0.x, hid, da[0]); da[1] = fmaf(w0.y, hid, da[1]); da[2] = fmaf(w0.z, hid, da[2]);
                da[3] = fmaf(w0.w, hid, da[3]); da[4] = fmaf(w1.x, hid, da[4]); da[5] = fmaf(w1.y, hid, da[5]);
                da[6] = fmaf(w1.z, hid, da[6]);
            }
        }
        // translation scale (model.py:325-331)
        float sc = S[M.f.Scal + SC_TRN2B];
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {  // rolled: keeps the epilogue inside the instruction cache
            tc::tmem_ld32(ctx.tmem + lane_base + 64 + 192 + 32 * half, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int n = 32 * half + c;
                sc = fmaf(S[M.f.PkMisc + 4 * n + 1], fmaxf(v[c] + S[M.f.PkMisc + 4 * n], 0.0f), sc);
            }
        }
        if (act) {
#pragma unroll
            for (int c = 0; c < PMHC_NTORS; ++c) out[5 + c] = da[c];
            out[12] = sc * rx; out[13] = sc * ry; out[14] = sc * rz;
        }
    }
    tc::fence_before_thread_sync();
}

template <int LAYER>
__global__ void __launch_bounds__(kTcThreads, 1) egnn_layer_forward_tc_kernel(LayerArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need a 1024-byte aligned base in the shared window: round up (1 KB slack is allocated)
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    float* S = reinterpret_cast<float*>(smem);
    const TcMap M = make_tc_map(a.Kpad, LAYER);
    const int tid = threadIdx.x, warp = tid >> 5;
    int* I = reinterpret_cast<int*>(S + M.f.Ints);
    uint64_t* bar = reinterpret_cast<uint64_t*>(S + M.Bar);

    if (warp == 0) tc::tmem_alloc(reinterpret_cast<uint32_t*>(S + M.TmemPtr), kTmemCols);
    if (tid == 32) {
        tc::mbar_init(bar, 1);
        tc::mbar_fence_init();
    }
    stage_weights_tc<LAYER>(smem, M, a.params);
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    TileCtx ctx;
    ctx.tmem = *reinterpret_cast<volatile uint32_t*>(S + M.TmemPtr);
    ctx.phase = 0;

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        float* ajt = a.ajt_cache + ((size_t)b * 2 + LAYER) * kHid * a.Kpad;
        const ComplexInfo ci = setup_complex_cached<LAYER>(S, M.f, S + M.Wq, S + M.TorX, a, b, ajt);
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
            for (int tile_base = 0; tile_base < gpairs; tile_base += kTileRows) {
                const int gp = tile_base + (tid & 127);
                const bool act = gp < gpairs;
                const PairRef pr = decode_full_pair(I, act ? gp : tile_base, W, L, row0, act);
                run_tile<LAYER, true>(smem, M, a, ctx, ajt, pr, 1.0f, gp, lsave);
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
                run_tile<LAYER, false>(smem, M, a, ctx, ajt, pr, mult, 0, nullptr);
            }
            __syncthreads();

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
    const TcMap M = make_tc_map(a.Kpad, LAYER);
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
    egnn_layer_forward_tc_kernel<LAYER><<<grid, kTcThreads, smem, stream>>>(a);
    if (profile_enabled()) profile_mark(PROF_FWD, stream, false);
    PMHC_CHECK_LAUNCH("egnn_layer_forward_tc");
    return 0;
}

// Step-invariant pocket side of message_mlp.0 for both layers (model.py:401, 411-412: pocket nodes carry no time
// feature and layer 2 sees the same 22 features zero-padded), computed once per batch / sampling trajectory:
//   ajt_cache[b][l][k][16 + p] = W1_l[k, H_l : H_l + 22] . pocket_features[b][p]      (zero beyond P)
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
        out[(4 * kq + 0) * Kpad + kN + pj] = a0;
        out[(4 * kq + 1) * Kpad + kN + pj] = a1;
        out[(4 * kq + 2) * Kpad + kN + pj] = a2;
        out[(4 * kq + 3) * Kpad + kN + pj] = a3;
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
