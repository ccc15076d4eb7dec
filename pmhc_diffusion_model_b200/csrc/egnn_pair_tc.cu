// egnn_pair_tc.cu — fused EGNN layer forward, bf16 tensor-core mode, second generation.
//
// Every dense contraction of the layer (model.py:47-81, 183-333) runs on tcgen05 tensor cores; the CUDA cores only
// build operands and do the geometry.  Per 128-pair tile (one pair = one TMEM lane = one thread of an engine):
//
//   CUDA   m1 = relu(A_i + A_j + W_e)                          packed bf16x2 arithmetic -> SW128 tile in shared memory
//   MMA 1  D1[128x64]  = m1 . W2^T                             (layer 1 also: Dsum += m1^T . Sel, the unmasked message
//                                                               sums of model.py:151 as a tensor-core column sum)
//   CUDA   A2 = bf16(D1) written back to TENSOR MEMORY, plus 32 "extra" K columns per pair: -d2 (hi/lo), qdot2,
//          the local quaternion, the row's torsions and 1.0 (biases, hi/lo) -> the geometry inputs and all first-layer
//          biases of the four heads ride inside the contraction (message bias b2 is folded into them: W_h b2)
//   MMA 2  D2[128x128] = [A2 | extras] . [W_att | W_rot]^T     then, same columns, [W_tor | W_trn]: A operand from TMEM
//   CUDA   A3 = relu(D2) -> packed bf16x2, in place in TMEM (cvt.rn.relu.bf16x2.f32)
//   MMA 3  D3[128x16] += A3 . W3^T                             second layers of the four heads (64 -> 1, 4, 7, 1)
//   CUDA   sigmoid, quaternion sandwich, translation scale -> 15 floats per pair; quarter-warp softmax per row
//
// A CTA holds the layer's weights once (bf16 operand tiles, 57 KB) and runs TWO independent engines of 4 warps; each
// engine owns half of tensor memory (256 columns), its own operand tile and per-complex arrays, and walks its own
// complexes, so one engine's CUDA-core stages overlap the other's MMAs.  MMAs are issued by an elected lane of the
// engine's first warp from warp-uniform code (see tc::elect_one).
//
// Per-node work is not done here: node_pre_kernel (once per batch / sampling trajectory) projects the pocket and the
// peptide's static features through message_mlp.0; node_mid_kernel (between the layers, tensor cores, 128 nodes per
// CTA) turns layer 1's message sums into layer 2's node features and projections.
//
// Precision (PMHC_PRECISION_BF16, gate 1e-2): bf16 operands, fp32 accumulation in TMEM; geometry, softmax and the
// frame / torsion updates in fp32.
#include "egnn_common.cuh"
#include "tcgen05.cuh"

namespace pmhc {
namespace tc2 {

constexpr int kEngines = 2;
constexpr int kEngThreads = 128;                 // compute threads of one engine (one per pair row / TMEM lane)
constexpr int kComputeThreads = kEngines * kEngThreads;
constexpr int kThreads = kComputeThreads + 128;  // + one warpgroup: warp 8 / 9 issue the MMAs of engine 0 / 1, two warps idle
// setmaxnreg budgets.  An increase only succeeds out of the registers the CTA was LAUNCHED with (384 threads x the 168 ptxas
// allots under __launch_bounds__(384, 1) = 64 512): 256 * 232 + 128 * 40 = 64 512 uses exactly that pool; a larger sum makes
// the increase wait for ever.
constexpr int kRegsCompute = 232, kRegsIssue = 40;
static_assert(kComputeThreads * kRegsCompute + 128 * kRegsIssue <= kThreads * 168, "setmaxnreg budgets exceed the launch pool");
constexpr int kTile = 128;
constexpr int kEngCols = 256;                 // TMEM columns per engine
// TMEM columns of one engine, relative to its base
constexpr int TM_A2 = 0;                      // D1 [0,64) -> A2 = bf16 message, 32 columns
constexpr int TM_XA = 32;                     // extras of the attention / rotation heads, 8 columns = 16 bf16
constexpr int TM_XB = 40;                     // extras of the torsion / translation heads
constexpr int TM_D3 = 48;                     // second-layer outputs, 16 columns
constexpr int TM_X = 64;                      // head buffers: hidden pre-activations D2 (64 columns), then the
constexpr int TM_Y = 128;                     //   second-layer operand A3 in place; X: attention then translation,
constexpr int TM_Z = 192;                     //   Y: rotation, Z: torsion (layer 1: also the per-tile message column sums)
// mbarriers of one engine
constexpr int BM = 0, BX = 1, BY = 2, BZ = 3;

struct PairArgs {
    int B, P, Kpad;
    float t_over_T;
    const float* t_dev;              // nullable: t / T in device memory (see time_feature)
    const float* frames_in;          // [B,16,7]
    const float* tors_in;            // [B,16,14]
    const uint8_t* mask;             // [B,16]
    const float* pocket_frames;      // [B,P,7]
    const uint8_t* pocket_cls;       // [B,cls_stride]
    const __nv_bfloat16* pk_cache;   // [B,2,P,64] pocket rows of A_j per layer
    const __nv_bfloat16* aij;        // [B,16,128] (A_i + b1 | A_j) of this layer's peptide nodes (layer 1: without the time term)
    const uint8_t* wimage;           // this layer's operand tiles in their shared-memory layout (weight_image_kernel)
    int cls_stride;                  // row stride of pocket_cls (P rounded up to 16)
    long long* dbg;                  // development: per-phase clock64 stamps of engine 0 / CTA 0 (nullable)
    float* frames_out;               // [B,16,7]
    float* tors_out;                 // [B,16,14]
    float* ssum_out;                 // layer 1: [B,16,64] sum_j m1_ij (before message_mlp.2), zero for padded rows
    float* rowstat;                  // training only
    float* logit_out;                // training only: [B,16,Kpad]
    const int* order;                // [B] complexes sorted by pair count, largest first (work item -> complex); nullable
    int cap_pairs;                   // pairs buffered per row group
    int aj_rows;                     // rows of A_j kept in shared memory: 16 (peptide only) or 16 + P
};

// Work of one engine.  Complexes — taken in the order of `order`, largest first, so that a round holds complexes of
// similar size when peptide lengths and pocket sizes are mixed — are dealt round-robin over the E = 2 * gridDim.x engines; when the last round is only
// partly filled (rem < E complexes left), each of its complexes is split by peptide rows into `parts` pieces handled by
// different engines, so the tail of the launch costs a fraction of a complex instead of a whole one.  Engine slots are
// ordered "first engine of every CTA, then the second", so a thin tail spreads over the SMs.
constexpr int kMaxParts = 4;
struct Work {
    int b, part, parts;
};
__device__ __forceinline__ bool get_work(int k, int cta, int eng, int ctas, int B, const int* __restrict__ order, Work& w) {
    const int E = ctas * kEngines;
    const int full = B / E, rem = B - full * E;
    int item;
    if (k < full) {
        item = k * E + cta * kEngines + eng;
        w.part = 0;
        w.parts = 1;
    } else {
        if (k > full || rem == 0) return false;
        int S = E / rem;
        S = S > kMaxParts ? kMaxParts : S;
        const int slot = eng * ctas + cta;
        if (slot >= rem * S) return false;
        item = full * E + slot / S;
        w.part = slot - (slot / S) * S;
        w.parts = S;
    }
    w.b = order != nullptr ? __ldg(order + item) : item;
    return true;
}
// real peptide rows [beg, end) of a part
__device__ __forceinline__ void part_rows(const Work& w, int L, int& beg, int& end) {
    beg = (L * w.part) / w.parts;
    end = (L * (w.part + 1)) / w.parts;
}

struct Map {
    int W2b, Whb, W3b, Wxb, Web, Misc, Bar, TmemPtr, cta_bytes;       // CTA-shared, byte offsets
    int A1, Sel, AjS, Out, Ai, Q, X, Tors, TorsB, Ints, Cls, eng_bytes;   // per engine, byte offsets from the engine base
    int total_bytes;
};
// Misc words: [0,16) second-layer biases (fp32) in D3 order; [16,80) layer 1: time column of message_mlp.0 as bf16x2
// (A_i part | A_j part); [80,144) attention_mlp.2.weight in fp32 (the attention logit is finished on the CUDA cores)
constexpr int MISC_B2ND = 0, MISC_TIME = 16, MISC_ATT2 = 80, MISC_FLOATS = 144;

__host__ __device__ inline Map make_map(int Kpad, int cap_pairs, int aj_rows) {
    Map m;
    int o = 0;
    m.W2b = o; o += 64 * 128;
    m.Whb = o; o += 256 * 128;
    m.W3b = o; o += 5 * 2048;
    m.Wxb = o; o += 256 * 32;
    m.Web = o; o += 32 * 128;
    m.Misc = o; o += MISC_FLOATS * 4;
    m.Bar = o; o += 64;   // [0, Bar) is the weight image; 4 mbarriers per engine
    m.TmemPtr = o; o += 32;
    o = (o + 1023) & ~1023;
    m.cta_bytes = o;
    int e = 0;
    m.A1 = e; e += kTile * 128;
    m.Sel = e; e += 32 * 128;
    m.AjS = e; e += aj_rows * 128;
    m.Out = e; e += cap_pairs * kOutPerPair * 4;
    m.Ai = e; e += kN * 128;
    m.Q = e; e += Kpad * 16;
    m.X = e; e += Kpad * 16;
    m.Tors = e; e += kN * 14 * 4;
    m.TorsB = e; e += kN * 32;
    m.Ints = e; e += (Kpad + 64) * 4;
    m.Cls = e; e += Kpad + 32;            // [0,16) peptide mask, [16, ...) pocket slot classes
    e = (e + 1023) & ~1023;
    m.eng_bytes = e;
    m.total_bytes = m.cta_bytes + kEngines * m.eng_bytes;
    return m;
}

// ---------------------------------------------------------------------------------------------------------------
// operand tiles of one layer, in the shared-memory layout of the pair kernel (Map offsets [0, Bar))
// ---------------------------------------------------------------------------------------------------------------
template <int LAYER>
__device__ inline void build_weight_image(uint8_t* smem, const Map& M, const float* __restrict__ params, int tid, int kThreads) {
    constexpr int L = LAYER;
    constexpr int H = layer_H(L);
    constexpr int ld1 = 2 * H + kEdge;
    const float* msg0 = params + param_offset(L, MSG0_W);
    const float* msg2 = params + param_offset(L, MSG2_W);
    const float* msg2b = params + param_offset(L, MSG2_B);
    const float* head[4] = {params + param_offset(L, ATT0_W), params + param_offset(L, ROT0_W),
                            params + param_offset(L, TOR0_W), params + param_offset(L, TRN0_W)};
    const float* headb[4] = {params + param_offset(L, ATT0_B), params + param_offset(L, ROT0_B),
                             params + param_offset(L, TOR0_B), params + param_offset(L, TRN0_B)};
    const int ldh[4] = {66, 68, 78, 64};
    for (int idx = tid; idx < 64 * 32; idx += kThreads) {            // W2: B[n][k] = message_mlp.2.weight[n][k]
        int n = idx >> 5, k = (idx & 31) * 2;
        *reinterpret_cast<uint32_t*>(smem + M.W2b + tc::sw128_offset(n, k)) = tc::pack_bf16x2(msg2[n * 64 + k], msg2[n * 64 + k + 1]);
    }
    for (int idx = tid; idx < 256 * 32; idx += kThreads) {           // heads: rows 64h + n, message columns
        int row = idx >> 5, k = (idx & 31) * 2;
        int h = row >> 6, n = row & 63;
        const float* w = head[h] + n * ldh[h] + k;
        *reinterpret_cast<uint32_t*>(smem + M.Whb + tc::sw128_offset(row, k)) = tc::pack_bf16x2(w[0], w[1]);
    }
    {   // second layers as one [16 x 256] operand: row 0 attention, 1..4 rotation, 5..11 torsion, 12 translation
        const float* att2 = params + param_offset(L, ATT2_W);
        const float* rot2 = params + param_offset(L, ROT2_W);
        const float* tor2 = params + param_offset(L, TOR2_W);
        const float* trn2 = params + param_offset(L, TRN2_W);
        for (int idx = tid; idx < 16 * 256; idx += kThreads) {
            int n = idx >> 8, k = idx & 255;
            int h = k >> 6, c = k & 63;
            float v = 0.0f;
            if (h == 0 && n == 0) v = att2[c];
            else if (h == 1 && n >= 1 && n <= 4) v = rot2[(n - 1) * 64 + c];
            else if (h == 2 && n >= 5 && n <= 11) v = tor2[(n - 5) * 64 + c];
            else if (h == 3 && n == 12) v = trn2[c];
            *reinterpret_cast<__nv_bfloat16*>(smem + M.W3b + h * 2048 + tc::sw128_offset(n, c)) = __float2bfloat16_rn(v);
            // block 4: low half of the attention second layer (its hidden units see -d2 and reach 1e2..1e3, so the
            // logit is computed as hi.hi + lo.hi + hi.lo, ~16 mantissa bits on both operands)
            if (h == 0) *reinterpret_cast<__nv_bfloat16*>(smem + M.W3b + 4 * 2048 + tc::sw128_offset(n, c)) =
                            __float2bfloat16_rn(n == 0 ? v - tc::bf16_round(v) : 0.0f);
        }
    }
    for (int row = tid; row < 256; row += kThreads) {   // extras operand, one row per hidden unit of the four heads
        const int h = row >> 6, n = row & 63;            // (no-swizzle K-major core matrices, K = 16)
        float bias = headb[h][n];
        const float* w = head[h] + n * ldh[h];
        for (int k = 0; k < 64; ++k) bias = fmaf(w[k], msg2b[k], bias);   // message bias folded in: W_h b2
        const float bh = tc::bf16_round(bias), bl = bias - bh;
        float x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = 0.0f;
        if (h == 0) {          // inputs: -d2 hi, -d2 lo, -d2 hi, qd hi | . . . . | 1, 1, qd lo, qd hi
            const float wd = w[64], wq = w[65];
            const float wdh = tc::bf16_round(wd), wqh = tc::bf16_round(wq);
            x[0] = wdh; x[1] = wdh; x[2] = wd - wdh; x[3] = wqh;
            x[8] = bh; x[9] = bl; x[10] = wqh; x[11] = wq - wqh;
        } else if (h == 1) {   // inputs 4..7: local quaternion
            x[4] = w[64]; x[5] = w[65]; x[6] = w[66]; x[7] = w[67];
            x[8] = bh; x[9] = bl;
        } else if (h == 2) {   // inputs 0..13: the row's torsions
#pragma unroll
            for (int c = 0; c < 14; ++c) x[c] = w[64 + c];
            x[14] = bh; x[15] = bl;
        } else {
            x[14] = bh; x[15] = bl;
        }
        uint8_t* dst = smem + M.Wxb + (row >> 3) * 256 + (row & 7) * 16;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) pk[e] = tc::pack_bf16x2(x[8 * half + 2 * e], x[8 * half + 2 * e + 1]);
            *reinterpret_cast<uint4*>(dst + half * 128) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
    for (int idx = tid; idx < kEdge * 32; idx += kThreads) {         // edge one-hot columns of message_mlp.0
        int r = idx >> 5, k = (idx & 31) * 2;
        // 16-byte chunks of a row are XOR-swizzled by the row: neighbouring relative positions read different banks
        *reinterpret_cast<uint32_t*>(smem + M.Web + r * 128 + ((((k >> 3) ^ r) & 7) << 4) + (k & 7) * 2) =
            tc::pack_bf16x2(msg0[k * ld1 + 2 * H + r], msg0[(k + 1) * ld1 + 2 * H + r]);
    }
    float* misc = reinterpret_cast<float*>(smem + M.Misc);
    if (tid < 16) {
        float v = 0.0f;
        if (tid == 0) v = params[param_offset(L, ATT2_B)];
        else if (tid <= 4) v = params[param_offset(L, ROT2_B) + tid - 1];
        else if (tid <= 11) v = params[param_offset(L, TOR2_B) + tid - 5];
        else if (tid == 12) v = params[param_offset(L, TRN2_B)];
        misc[MISC_B2ND + tid] = v;
    }
    for (int n = tid; n < 64; n += kThreads) misc[MISC_ATT2 + n] = params[param_offset(L, ATT2_W) + n];
    if (LAYER == 0) {
        for (int w = tid; w < 64; w += kThreads) {
            const int k = 2 * (w & 31), col = w < 32 ? PMHC_NFEAT : H + PMHC_NFEAT;
            reinterpret_cast<uint32_t*>(misc)[MISC_TIME + w] = tc::pack_bf16x2(msg0[k * ld1 + col], msg0[(k + 1) * ld1 + col]);
        }
    }
}

template <int LAYER>
__global__ void __launch_bounds__(256) weight_image_kernel(const float* __restrict__ params, uint8_t* __restrict__ image) {
    const Map M = make_map(32, 256, kN);
    build_weight_image<LAYER>(image, M, params, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// ---------------------------------------------------------------------------------------------------------------
// one engine: 128 threads, 256 TMEM columns, its own shared-memory block
// ---------------------------------------------------------------------------------------------------------------
struct Engine {
    uint8_t* smem;      // CTA base (weights)
    uint8_t* es;        // engine base
    const Map& M;
    const PairArgs& a;
    int eng, et;        // engine index, thread within the engine (= pair row = TMEM lane)
    uint32_t tmem;      // engine's TMEM base (column offset applied)
    uint32_t lane_base; // (32 * warp-in-engine) << 16
    uint64_t* bar;      // this engine's four MMA completion mbarriers (BM, BX, BY, BZ)
    uint32_t phase;     // bit k: parity of the next completion of barrier k
    uint32_t smem_u, es_u;   // shared-window addresses of the CTA block and the engine block

    // value the compiler cannot see through: keeps the operand descriptors from being hoisted out of the issue
    // branch into ~40 long-lived registers (they are recomputed by the one issuing lane, a handful of integer ops)
    static __device__ __forceinline__ uint32_t opaque(uint32_t x) {
        uint32_t r;
        asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(x));
        return r;
    }

    // named barriers of the engine: two alternating request barriers (compute threads arrive, the issuing warp syncs),
    // one among the compute threads, one among all 160 threads
    uint32_t req;       // requests made (compute threads) / served (issuing warp) so far
    __device__ __forceinline__ int bar_id(int k) const { return 1 + 4 * eng + k; }
    __device__ __forceinline__ void sync() const { tc::named_bar_sync(bar_id(2), kEngThreads); }
    __device__ __forceinline__ void sync_all() const { tc::named_bar_sync(bar_id(3), kEngThreads + 32); }
    // every compute thread waits for every committed completion exactly once, in the same order
    __device__ __forceinline__ void wait(int k) {
        tc::mbar_wait(bar + k, (phase >> k) & 1u);
        phase ^= 1u << k;
        tc::fence_after_thread_sync();
    }
    // compute thread: my TMEM / shared-memory operand writes are done -> ask the issuing warp for the next MMA batch.
    // Between two requests every compute thread waits for a completion caused by the earlier one, so a thread is never
    // more than one request ahead and two alternating barriers suffice.
    __device__ __forceinline__ void request() {
        tc::tmem_wait_st();
        tc::fence_before_thread_sync();
        asm volatile("bar.arrive %0, %1;" ::"r"(bar_id((int)(req & 1u))), "r"(kEngThreads + 32) : "memory");
        ++req;
    }
    // issuing warp: wait for the request, then one elected lane issues f (which commits to the barriers it wants)
    template <class F>
    __device__ __forceinline__ void serve(F&& f) {
        tc::named_bar_sync(bar_id((int)(req & 1u)), kEngThreads + 32);
        ++req;
        tc::fence_after_thread_sync();
        if (tc::elect_one()) f();
        __syncwarp();
    }
    __device__ __forceinline__ void commit(int k) const { tc::mma_commit(bar + k); }
    __device__ __forceinline__ const int* ints() const { return reinterpret_cast<const int*>(es + M.Ints); }
};

// m1 row of this thread's pair -> A1 tile
template <int LAYER, bool FENCE = true>
__device__ __forceinline__ void stage_a1(const Engine& E, const PairRef& pr, int b) {
    const int r = E.et, i = pr.i, j = pr.j;
    const bool pep = (j >= 0 && j < kN);
    const uint4* ai = reinterpret_cast<const uint4*>(E.es + E.M.Ai) + i * 8;
    uint4 bj[8];
    if (j >= 0 && j < E.a.aj_rows) {
        const uint4* src = reinterpret_cast<const uint4*>(E.es + E.M.AjS) + j * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) bj[c] = src[c ^ (j & 7)];
    } else if (j >= 0) {
        const uint4* src = reinterpret_cast<const uint4*>(E.a.pk_cache + (((size_t)b * 2 + LAYER) * E.a.P + (j - kN)) * kHid);
#pragma unroll
        for (int c = 0; c < 8; ++c) bj[c] = __ldg(src + c);
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) bj[c] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (pep) {
        const int rel = kN - 1 + i - j;
        const uint4* we = reinterpret_cast<const uint4*>(E.smem + E.M.Web) + rel * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 w = we[c ^ (rel & 7)];
            bj[c].x = tc::add_bf16x2(bj[c].x, w.x); bj[c].y = tc::add_bf16x2(bj[c].y, w.y);
            bj[c].z = tc::add_bf16x2(bj[c].z, w.z); bj[c].w = tc::add_bf16x2(bj[c].w, w.w);
        }
    }
    uint8_t* row = E.es + E.M.A1 + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 av = ai[c];
        uint4 v;
        v.x = tc::add_relu_bf16x2(av.x, bj[c].x); v.y = tc::add_relu_bf16x2(av.y, bj[c].y);
        v.z = tc::add_relu_bf16x2(av.z, bj[c].z); v.w = tc::add_relu_bf16x2(av.w, bj[c].w);
        *reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4)) = v;
    }
    if (FENCE) tc::fence_proxy_async_smem();
}

// layer 1: this pair's entry of the column-sum selector (Sel[16 * (r / 64) + i][r % 64] = multiplicity).  A thread only
// ever touches column r % 64 of row block r / 64, so writing and clearing need no synchronisation between threads.
__device__ __forceinline__ void write_sel(const Engine& E, const PairRef& pr, float mult) {
    if (pr.active) {
        const int r = E.et;
        *reinterpret_cast<__nv_bfloat16*>(E.es + E.M.Sel + tc::sw128_offset(16 * (r >> 6) + pr.i, r & 63)) = __float2bfloat16_rn(mult);
    }
}
__device__ __forceinline__ void clear_sel(const Engine& E, const PairRef& pr) { write_sel(E, pr, 0.0f); }
__device__ __forceinline__ void zero_sel(const Engine& E) {
    uint4* s = reinterpret_cast<uint4*>(E.es + E.M.Sel);
    s[E.et] = make_uint4(0u, 0u, 0u, 0u);
    s[E.et + kEngThreads] = make_uint4(0u, 0u, 0u, 0u);
}

// D1 = m1 . W2^T  (+ layer 1: the tile's message column sums into buffer Z), completion -> BM
template <int LAYER>
__device__ __forceinline__ void mma_first(const Engine& E, bool with_d1) {
    const uint32_t cta = Engine::opaque(E.smem_u), tm = Engine::opaque(E.tmem);
    const uint32_t a1 = Engine::opaque(E.es_u) + E.M.A1;
    if (with_d1) {
        const uint64_t da = tc::smem_desc_sw128(a1);
        const uint64_t db = tc::smem_desc_sw128(cta + E.M.W2b);
        constexpr uint32_t id = tc::idesc_bf16_f32(128, 64);
#pragma unroll
        for (int s = 0; s < 4; ++s) tc::mma_bf16(tm + TM_A2, da + 2 * s, db + 2 * s, id, s > 0);
    }
    if (LAYER == 0) {
        // message column sums: A = the pair tile read MN-major (M = 2 atoms of 64 features for pairs 0..63 | 64..127,
        // K = 64 pairs), B = the selector
        const uint64_t db = tc::smem_desc_sw128(a1 - E.M.A1 + E.M.Sel);
        constexpr uint32_t id = tc::idesc_bf16_f32_major(128, 32, 1, 0);
#pragma unroll
        for (int s = 0; s < 4; ++s) tc::mma_bf16(tm + TM_Z, tc::smem_desc(a1 + s * 2048, 8192, 1024, 2), db + 2 * s, id, s > 0);
    }
    E.commit(BM);
}
// hidden pre-activations of head h: D2 = [A2 | extras] . W_h^T  (h: 0 attention, 1 rotation, 2 torsion, 3 translation)
__device__ __forceinline__ void mma_head(const Engine& E, uint32_t cta, uint32_t tm, int h, int dst) {
    const uint64_t db = tc::smem_desc_sw128(cta + E.M.Whb + h * 64 * 128);
    const uint64_t dx = tc::smem_desc(cta + E.M.Wxb + h * 64 * 32, 128, 256, 0);
    constexpr uint32_t id = tc::idesc_bf16_f32(128, 64);
#pragma unroll
    for (int s = 0; s < 4; ++s) tc::mma_bf16_ts(tm + dst, tm + TM_A2 + 8 * s, db + 2 * s, id, s > 0);
    tc::mma_bf16_ts(tm + dst, tm + (h < 2 ? TM_XA : TM_XB), dx, id, 1);
}
// second layer of head h (1 rotation, 2 torsion, 3 translation) on its packed hidden units: D3 (+)= A3 . W3_h^T.  The
// attention head's second layer (64 -> 1) is a dot product on the CUDA cores, in fp32 (attention_logit below).
__device__ __forceinline__ void mma_second(const Engine& E, uint32_t cta, uint32_t tm, int h, int src, bool first) {
    constexpr uint32_t id = tc::idesc_bf16_f32(128, 16);
    const uint32_t w3 = cta + E.M.W3b;
#pragma unroll
    for (int s = 0; s < 4; ++s) tc::mma_bf16_ts(tm + TM_D3, tm + src + 8 * s, tc::smem_desc_sw128(w3 + h * 2048) + 2 * s, id, (first && s == 0) ? 0u : 1u);
}

// extras of the (attention, rotation) half for this pair
__device__ __forceinline__ void pair_extras(const Engine& E, const PairRef& pr, uint32_t (&xa)[8]) {
    const float4* Q = reinterpret_cast<const float4*>(E.es + E.M.Q);
    const float4* X = reinterpret_cast<const float4*>(E.es + E.M.X);
    const float4 qi4 = Q[pr.i], qj4 = Q[pr.j], xi = X[pr.i], xj = X[pr.j];
    const Quat qi{qi4.x, qi4.y, qi4.z, qi4.w}, qj{qj4.x, qj4.y, qj4.z, qj4.w};
    const float rx = xi.x - xj.x, ry = xi.y - xj.y, rz = xi.z - xj.z;
    const float nd2 = -(rx * rx + ry * ry + rz * rz);               // attention input -d2 (model.py:238)
    const float dq = qdot(qi, qj);
    const float qd = dq * dq;                                       // (q_i . q_j)^2 (model.py:239)
    const float in2 = __fdividef(1.0f, qdot(qj, qj));
    const Quat lq = qmul(Quat{qj.w * in2, -qj.x * in2, -qj.y * in2, -qj.z * in2}, qmul(qi, qj));   // local quaternion (model.py:283-287)
    const float dh = tc::bf16_round(nd2), qh = tc::bf16_round(qd);
    xa[0] = tc::pack_bf16x2(dh, nd2 - dh);
    xa[1] = tc::pack_bf16x2(dh, qh);
    xa[2] = tc::pack_bf16x2(lq.w, lq.x);
    xa[3] = tc::pack_bf16x2(lq.y, lq.z);
    xa[4] = 0x3F803F80u;                                            // (1, 1): bias hi / lo
    xa[5] = tc::pack_bf16x2(qd - qh, qh);
    xa[6] = 0u;
    xa[7] = 0u;
}

// D1 -> A2 (bf16 message) + both extras blocks, in tensor memory
__device__ __forceinline__ void epilogue1(const Engine& E, const PairRef& pr, const uint32_t (&xa)[8]) {
    uint32_t lo[32], hi[32];
    tc::tmem_ld32_nowait(E.tmem + E.lane_base + TM_A2, lo);
    tc::tmem_ld32_nowait(E.tmem + E.lane_base + TM_A2 + 32, hi);
    uint32_t xs[16];
#pragma unroll
    for (int c = 0; c < 8; ++c) xs[c] = xa[c];
    const uint4* tb = reinterpret_cast<const uint4*>(E.es + E.M.TorsB) + pr.i * 2;
    const uint4 t0 = tb[0], t1 = tb[1];
    xs[8] = t0.x; xs[9] = t0.y; xs[10] = t0.z; xs[11] = t0.w;
    xs[12] = t1.x; xs[13] = t1.y; xs[14] = t1.z; xs[15] = t1.w;
    tc::tmem_wait_ld();
    uint32_t pk[32];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        pk[c] = tc::pack_bf16x2(__uint_as_float(lo[2 * c]), __uint_as_float(lo[2 * c + 1]));
        pk[16 + c] = tc::pack_bf16x2(__uint_as_float(hi[2 * c]), __uint_as_float(hi[2 * c + 1]));
    }
    tc::tmem_st32(E.tmem + E.lane_base + TM_A2, pk);
    tc::tmem_st16(E.tmem + E.lane_base + TM_XA, xs);
}

// hidden units of one head (64 columns) -> relu -> packed bf16x2 in place (32 columns)
__device__ __forceinline__ void epilogue2(const Engine& E, int buf) {
    uint32_t v0[32], v1[32];
    tc::tmem_ld32_nowait(E.tmem + E.lane_base + buf, v0);
    tc::tmem_ld32_nowait(E.tmem + E.lane_base + buf + 32, v1);
    tc::tmem_wait_ld();
    uint32_t pk[32];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        pk[c] = tc::pack_bf16x2_relu(__uint_as_float(v0[2 * c]), __uint_as_float(v0[2 * c + 1]));
        pk[16 + c] = tc::pack_bf16x2_relu(__uint_as_float(v1[2 * c]), __uint_as_float(v1[2 * c + 1]));
    }
    tc::tmem_st32(E.tmem + E.lane_base + buf, pk);
}
// attention head: logit = attention_mlp.2 . relu(hidden) + bias (model.py:241-243) straight from the fp32 accumulators.  The
// hidden units see -d2 and reach 1e2..1e3, so a bf16 second layer would need a hi + lo split of both operands (three MMAs
// and a 64-column repack); 64 FMAs per pair on the CUDA cores are cheaper than that repack, exact in fp32, and take one MMA
// round trip out of the tile.  The buffer is free for the next head as soon as the loads have landed.
__device__ __forceinline__ float attention_logit(const Engine& E, int buf) {
    uint32_t v0[32], v1[32];
    tc::tmem_ld32_nowait(E.tmem + E.lane_base + buf, v0);
    tc::tmem_ld32_nowait(E.tmem + E.lane_base + buf + 32, v1);
    const float4* w = reinterpret_cast<const float4*>(E.smem + E.M.Misc) + MISC_ATT2 / 4;
    tc::tmem_wait_ld();
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 a = w[c], b = w[8 + c];
        s0 = fmaf(a.x, fmaxf(__uint_as_float(v0[4 * c + 0]), 0.0f), s0);
        s1 = fmaf(a.y, fmaxf(__uint_as_float(v0[4 * c + 1]), 0.0f), s1);
        s2 = fmaf(a.z, fmaxf(__uint_as_float(v0[4 * c + 2]), 0.0f), s2);
        s3 = fmaf(a.w, fmaxf(__uint_as_float(v0[4 * c + 3]), 0.0f), s3);
        s0 = fmaf(b.x, fmaxf(__uint_as_float(v1[4 * c + 0]), 0.0f), s0);
        s1 = fmaf(b.y, fmaxf(__uint_as_float(v1[4 * c + 1]), 0.0f), s1);
        s2 = fmaf(b.z, fmaxf(__uint_as_float(v1[4 * c + 2]), 0.0f), s2);
        s3 = fmaf(b.w, fmaxf(__uint_as_float(v1[4 * c + 3]), 0.0f), s3);
    }
    return (s0 + s1) + (s2 + s3);
}

// D3 -> per-pair outputs: logit, global delta quaternion, delta angles, scale * (x_i - x_j)   (model.py:243-331)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ void epilogue3(const Engine& E, const PairRef& pr, int slot, float* __restrict__ lsave, float att) {
    float o[16];
    tc::tmem_ld16(E.tmem + E.lane_base + TM_D3, o);
    const float4* b2nd = reinterpret_cast<const float4*>(E.smem + E.M.Misc) + MISC_B2ND / 4;
    const float4 b0 = b2nd[0], b1 = b2nd[1], b2 = b2nd[2], b3 = b2nd[3];
    const float4* Q = reinterpret_cast<const float4*>(E.es + E.M.Q);
    const float4* X = reinterpret_cast<const float4*>(E.es + E.M.X);
    const float4 qj4 = Q[pr.j], xi = X[pr.i], xj = X[pr.j];
    const Quat qj{qj4.x, qj4.y, qj4.z, qj4.w};
    const float in2 = __fdividef(1.0f, qdot(qj, qj));
    const Quat qinvj{qj.w * in2, -qj.x * in2, -qj.y * in2, -qj.z * in2};
    const Quat dl{fast_sigmoid(o[1] + b0.y), fast_sigmoid(o[2] + b0.z), fast_sigmoid(o[3] + b0.w), fast_sigmoid(o[4] + b1.x)};   // never normalised (T5)
    const Quat dg = qmul(qj, qmul(dl, qinvj));
    const float sc = o[12] + b3.x;
    if (pr.active) {
        float* out = reinterpret_cast<float*>(E.es + E.M.Out) + slot * kOutPerPair;
        const float logit = att + b0.x;
        out[0] = logit;
        out[1] = dg.w; out[2] = dg.x; out[3] = dg.y; out[4] = dg.z;
        out[5] = o[5] + b1.y; out[6] = o[6] + b1.z; out[7] = o[7] + b1.w;
        out[8] = o[8] + b2.x; out[9] = o[9] + b2.y; out[10] = o[10] + b2.z; out[11] = o[11] + b2.w;
        out[12] = sc * (xi.x - xj.x); out[13] = sc * (xi.y - xj.y); out[14] = sc * (xi.z - xj.z);
        if (lsave != nullptr) lsave[pr.i * E.a.Kpad + pr.j] = logit;
    }
}

// masked softmax over the buffered pair outputs + the weighted updates; one quarter-warp (8 lanes) per row
__device__ inline void finalize_rows_engine(const Engine& E, int b, int row0, int nrows, int W) {
    const PairArgs& a = E.a;
    const int lane = E.et & 31, l8 = lane & 7, q = lane >> 3, ew = E.et >> 5;
    const int* I = E.ints();
    const float* Out = reinterpret_cast<const float*>(E.es + E.M.Out);
    const float4* Q = reinterpret_cast<const float4*>(E.es + E.M.Q);
    const float4* X = reinterpret_cast<const float4*>(E.es + E.M.X);
    const float* Tors = reinterpret_cast<const float*>(E.es + E.M.Tors);
    for (int base = 0; base < nrows; base += 16) {
        const int rl = base + 4 * ew + q;
        const bool live = rl < nrows;
        const int i = I[IN_ROWS + row0 + (live ? rl : 0)];
        const float* out = Out + (size_t)(live ? rl : 0) * W * kOutPerPair;
        const int Wl = live ? W : 0;
        float mx = -INFINITY;
        for (int e = l8; e < Wl; e += 8) mx = fmaxf(mx, out[e * kOutPerPair]);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.0f, ws[14];
#pragma unroll
        for (int c = 0; c < 14; ++c) ws[c] = 0.0f;
        for (int e = l8; e < Wl; e += 8) {
            const float* o = out + e * kOutPerPair;
            const float p = expf(o[0] - mx);
            se += p;
#pragma unroll
            for (int c = 0; c < 14; ++c) ws[c] = fmaf(p, o[1 + c], ws[c]);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
#pragma unroll
            for (int c = 0; c < 14; ++c) ws[c] += __shfl_xor_sync(0xffffffffu, ws[c], o);
        }
        if (!live) continue;
        const float inv = W > 0 ? 1.0f / se : 0.0f;
#pragma unroll
        for (int c = 0; c < 14; ++c) ws[c] *= inv;
        const size_t node = (size_t)b * kN + i;
        if (l8 == 0) {
            const float4 qi = Q[i], xi = X[i];
            const Quat G{ws[0], ws[1], ws[2], ws[3]};
            const Quat gq = W > 0 ? qnormalize(G) : Quat{1.0f, 0.0f, 0.0f, 0.0f};   // model.py:301-306
            const Quat qo = qunit(qmul(gq, Quat{qi.x, qi.y, qi.z, qi.w}));          // model.py:310, :181
            float* fo = a.frames_out + node * 7;
            fo[0] = qo.w; fo[1] = qo.x; fo[2] = qo.y; fo[3] = qo.z;
            fo[4] = xi.x + ws[11]; fo[5] = xi.y + ws[12]; fo[6] = xi.z + ws[13];
            if (a.rowstat != nullptr) {
                float* rs = a.rowstat + node * PMHC_ROWSTAT;
                rs[0] = W > 0 ? mx + logf(se) : 0.0f;
#pragma unroll
                for (int c = 0; c < 14; ++c) rs[1 + c] = ws[c];
                rs[15] = 0.0f;
            }
        } else {
            // torsions' = (sin dA, cos dA) (x) torsions (model.py:263-269); lanes 1..7 take one torsion each
            const int tq = l8 - 1;
            float da = 0.0f;
#pragma unroll
            for (int c = 0; c < PMHC_NTORS; ++c) da = (tq == c) ? ws[4 + c] : da;
            float sn, cs;
            sincosf(da, &sn, &cs);
            const float* t = Tors + i * 14 + 2 * tq;
            const SinCos o = scmul(SinCos{sn, cs}, SinCos{t[0], t[1]});
            a.tors_out[node * 14 + 2 * tq] = o.s;
            a.tors_out[node * 14 + 2 * tq + 1] = o.c;
        }
    }
}

// per-complex setup of one engine.  Every global read is a cp.async issued up front (one memory round trip per complex
// instead of one per array); then the lists and the torsion extras rows are derived from the shared copies.
__device__ inline ComplexInfo setup_engine(const Engine& E, int b, bool layer1) {
    const PairArgs& a = E.a;
    const Map& M = E.M;
    const int et = E.et, lane = et & 31;
    const int P = a.P, K = kN + P, Kpad = a.Kpad;
    int* I = reinterpret_cast<int*>(E.es + M.Ints);
    float* Q = reinterpret_cast<float*>(E.es + M.Q);
    float* X = reinterpret_cast<float*>(E.es + M.X);
    float* Tors = reinterpret_cast<float*>(E.es + M.Tors);
    uint8_t* Cls = E.es + M.Cls;
    for (int idx = et; idx < K * 7; idx += kEngThreads) {
        const int j = idx / 7, c = idx - j * 7;
        const float* f = (j < kN) ? a.frames_in + ((size_t)b * kN + j) * 7 + c : a.pocket_frames + ((size_t)b * P + (j - kN)) * 7 + c;
        tc::cp_async_4(c < 4 ? Q + 4 * j + c : X + 4 * j + (c - 4), f);
    }
    for (int idx = et; idx < kN * 14 / 4; idx += kEngThreads) tc::cp_async_16(Tors + 4 * idx, a.tors_in + (size_t)b * kN * 14 + 4 * idx);
    if (et == 0) tc::cp_async_16(Cls, a.mask + (size_t)b * kN);
    for (int idx = et; idx < a.cls_stride / 16; idx += kEngThreads) tc::cp_async_16(Cls + 16 + 16 * idx, a.pocket_cls + (size_t)b * a.cls_stride + 16 * idx);
    {   // peptide rows of A_i (row-major) and A_j (chunk-swizzled like the pocket rows)
        uint4* ai = reinterpret_cast<uint4*>(E.es + M.Ai);
        uint4* aj = reinterpret_cast<uint4*>(E.es + M.AjS);
        const uint4* src = reinterpret_cast<const uint4*>(a.aij + (size_t)b * kN * 128);
        for (int idx = et; idx < kN * 16; idx += kEngThreads) {
            const int i = idx >> 4, ch = idx & 15;
            tc::cp_async_16(ch < 8 ? ai + i * 8 + ch : aj + i * 8 + ((ch - 8) ^ (i & 7)), src + idx);
        }
        if (a.aj_rows > kN) {
            const uint4* ps = reinterpret_cast<const uint4*>(a.pk_cache + ((size_t)b * 2 + (layer1 ? 0 : 1)) * P * kHid);
            for (int idx = et; idx < P * 8; idx += kEngThreads) {
                const int j = kN + (idx >> 3), c = idx & 7;
                tc::cp_async_16(aj + j * 8 + (c ^ (j & 7)), ps + idx);
            }
        }
    }
    if (layer1) zero_sel(E);
    tc::cp_async_wait_all();
    E.sync();
    if (layer1) {
        // time feature (model.py:394): A_i += (t/T) w_ti, A_j += (t/T) w_tj on the 16 peptide rows, packed bf16x2 FMAs
        const uint32_t* tw = reinterpret_cast<const uint32_t*>(E.smem + M.Misc) + MISC_TIME;
        const float tt = time_feature(a);
        const uint32_t t2 = tc::pack_bf16x2(tt, tt);
        uint4* ai = reinterpret_cast<uint4*>(E.es + M.Ai);
        uint4* aj = reinterpret_cast<uint4*>(E.es + M.AjS);
        for (int idx = et; idx < kN * 16; idx += kEngThreads) {
            const int i = idx >> 4, ch = idx & 15;
            uint4* p = ch < 8 ? ai + i * 8 + ch : aj + i * 8 + ((ch - 8) ^ (i & 7));
            const uint32_t* w = tw + 4 * ch;
            uint4 v = *p;
            v.x = tc::fma_bf16x2(w[0], t2, v.x); v.y = tc::fma_bf16x2(w[1], t2, v.y);
            v.z = tc::fma_bf16x2(w[2], t2, v.z); v.w = tc::fma_bf16x2(w[3], t2, v.w);
            *p = v;
        }
    }
    {   // torsion extras rows: 14 bf16 (sin, cos) + (1, 1) for the biases
        const int i = et >> 3, w = et & 7;
        reinterpret_cast<uint32_t*>(E.es + M.TorsB)[et] = w < 7 ? tc::pack_bf16x2(Tors[i * 14 + 2 * w], Tors[i * 14 + 2 * w + 1]) : 0x3F803F80u;
    }
    if ((et >> 5) == 0) {
        const bool real = lane < kN && Cls[lane] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, real);
        const int pos = __popc(bal & ((1u << lane) - 1u));
        const int Lr = __popc(bal);
        if (lane < kN) {
            if (real) I[IN_ROWS + pos] = lane;
            else I[IN_PEPX + (lane - pos)] = lane;
        }
        int nv = 0, nx = 0, c0 = 0;
        for (int base = 0; base < P; base += 32) {
            const int j = base + lane;
            const int cls = j < P ? (int)Cls[16 + j] : 3;
            const unsigned bv = __ballot_sync(0xffffffffu, cls == 0);
            const unsigned bx = __ballot_sync(0xffffffffu, cls == 2);
            const unsigned bz = __ballot_sync(0xffffffffu, cls == 1);
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
    E.sync_all();   // + the issuing warp, which reads the lists' counts
    ComplexInfo ci;
    ci.L = I[IN_POCKET + Kpad + 0];
    ci.nv = I[IN_POCKET + Kpad + 1];
    ci.nx = I[IN_POCKET + Kpad + 2];
    ci.c0 = I[IN_POCKET + Kpad + 3];
    return ci;
}

// tiles and message-only tiles of one complex, identically derived by the compute threads and the issuing warp
struct ComplexPlan {
    int L, W, rows_per_group, msg_w, msg_tiles;
};
__device__ __forceinline__ ComplexPlan plan_complex(const ComplexInfo& ci, int cap_pairs, bool layer1) {
    ComplexPlan p;
    p.L = ci.L;
    p.W = (ci.L - 1) + ci.nv;
    p.rows_per_group = p.W > 0 ? max(1, cap_pairs / p.W) : kN;
    // message-only pairs (model.py:151 sums over ALL K slots): self, masked peptide slots, masked pocket slots with their
    // own features, and one shared message for the c0 zero-feature masked pocket slots (multiplicity <= 256 stays exact in bf16)
    const int nshared = ci.c0 > 256 ? 2 : (ci.c0 > 0 ? 1 : 0);
    p.msg_w = 1 + (kN - ci.L) + ci.nx + nshared;
    p.msg_tiles = layer1 ? (ci.L * p.msg_w + kTile - 1) / kTile : 0;
    return p;
}

template <int LAYER>
__global__ void __launch_bounds__(kThreads, 1) egnn_pair_tc_kernel(PairArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const Map M = make_map(a.Kpad, a.cap_pairs, a.aj_rows);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + M.Bar);

    if (warp == 0) tc::tmem_alloc(reinterpret_cast<uint32_t*>(smem + M.TmemPtr), 512);
    if (tid == 32) {
        for (int k = 0; k < 4 * kEngines; ++k) tc::mbar_init(bars + k, 1);
        tc::mbar_fence_init();
    }
    {   // the layer's operand tiles, prepared by weight_image_kernel: one coalesced copy, all of a thread's 16-byte pieces
        // in flight at once (cp.async: one memory round trip instead of one per piece)
        const uint4* src = reinterpret_cast<const uint4*>(a.wimage);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int idx = tid; idx < M.Bar / 16; idx += kThreads) tc::cp_async_16(dst + idx, src + idx);
        tc::cp_async_wait_all();
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + M.TmemPtr);

    if (tid >= kComputeThreads) {
        // =============================== MMA issue: warp 8 -> engine 0, warp 9 -> engine 1 ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsIssue));
        const int eng = warp - kComputeThreads / 32;
        if (eng < kEngines) {
            Engine E{smem, smem + M.cta_bytes + eng * M.eng_bytes, M, a, eng, kEngThreads, tmem_base + (uint32_t)(eng * kEngCols), 0u,
                     bars + 4 * eng, 0u, tc::smem_u32(smem), tc::smem_u32(smem + M.cta_bytes + eng * M.eng_bytes), 0u};
            const int* I = E.ints();
            Work wk;
            for (int k = 0; get_work(k, blockIdx.x, eng, gridDim.x, a.B, a.order, wk); ++k) {
                E.sync_all();   // the compute threads have set the complex up
                ComplexInfo ci;
                ci.L = I[IN_POCKET + a.Kpad + 0];
                ci.nv = I[IN_POCKET + a.Kpad + 1];
                ci.nx = I[IN_POCKET + a.Kpad + 2];
                ci.c0 = I[IN_POCKET + a.Kpad + 3];
                const ComplexPlan cp = plan_complex(ci, a.cap_pairs, LAYER == 0);
                int rbeg, rend;
                part_rows(wk, cp.L, rbeg, rend);
                for (int row0 = rbeg; row0 < rend; row0 += cp.rows_per_group) {
                    const int ntiles = (min(cp.rows_per_group, rend - row0) * cp.W + kTile - 1) / kTile;
                    for (int t = 0; t < ntiles; ++t) {
                        E.serve([&] { mma_first<LAYER>(E, true); });
                        E.serve([&] {
                            const uint32_t cta = Engine::opaque(E.smem_u), tm = Engine::opaque(E.tmem);
                            mma_head(E, cta, tm, 0, TM_X); E.commit(BX);
                            mma_head(E, cta, tm, 1, TM_Y); E.commit(BY);
                            mma_head(E, cta, tm, 2, TM_Z); E.commit(BZ);
                        });
                        E.serve([&] {   // the compute threads have read the attention hidden units out of X: X is free
                            const uint32_t cta = Engine::opaque(E.smem_u), tm = Engine::opaque(E.tmem);
                            mma_second(E, cta, tm, 1, TM_Y, true);
                            mma_head(E, cta, tm, 3, TM_X); E.commit(BX);
                        });
                        E.serve([&] {
                            const uint32_t cta = Engine::opaque(E.smem_u), tm = Engine::opaque(E.tmem);
                            mma_second(E, cta, tm, 2, TM_Z, false); E.commit(BZ);
                        });
                        E.serve([&] {
                            const uint32_t cta = Engine::opaque(E.smem_u), tm = Engine::opaque(E.tmem);
                            mma_second(E, cta, tm, 3, TM_X, false); E.commit(BM);
                        });
                    }
                }
                const int msg_tiles = LAYER == 0 ? ((rend - rbeg) * cp.msg_w + kTile - 1) / kTile : 0;
                for (int t = 0; t < msg_tiles; ++t) E.serve([&] { mma_first<LAYER>(E, false); });
            }
        }
    } else {
        // ======================================= compute: one thread per pair row =======================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCompute));
        const int eng = tid >> 7, et = tid & 127;
        Engine E{smem, smem + M.cta_bytes + eng * M.eng_bytes, M, a, eng, et, tmem_base + (uint32_t)(eng * kEngCols),
                 (uint32_t)(((et >> 5) & 3) * 32) << 16, bars + 4 * eng, 0u, tc::smem_u32(smem),
                 tc::smem_u32(smem + M.cta_bytes + eng * M.eng_bytes), 0u};
        const int* I = E.ints();
        int ts_n = 0;
        const bool ts_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
#define PMHC_TS() do { if (ts_on && ts_n < 120) a.dbg[ts_n++] = clock64(); } while (0)
        PMHC_TS();

        Work wk;
        for (int k = 0; get_work(k, blockIdx.x, eng, gridDim.x, a.B, a.order, wk); ++k) {
            const int b = wk.b;
            PMHC_TS();
            const ComplexInfo ci = setup_engine(E, b, LAYER == 0);
            PMHC_TS();
            const ComplexPlan cp = plan_complex(ci, a.cap_pairs, LAYER == 0);
            const int L = cp.L, W = cp.W;
            int rbeg, rend;
            part_rows(wk, L, rbeg, rend);
            float* lsave = a.logit_out ? a.logit_out + (size_t)b * kN * a.Kpad : nullptr;
            if (wk.part == 0) {   // padded rows: pass-through (T4)
                for (int idx = et; idx < (kN - L) * 21; idx += kEngThreads) {
                    const int s = idx / 21, c = idx - s * 21;
                    const int i = I[IN_PEPX + s];
                    if (c < 7) a.frames_out[((size_t)b * kN + i) * 7 + c] = a.frames_in[((size_t)b * kN + i) * 7 + c];
                    else a.tors_out[((size_t)b * kN + i) * 14 + (c - 7)] = a.tors_in[((size_t)b * kN + i) * 14 + (c - 7)];
                }
            }
            // layer 1: this thread's share of the message column sums, sum_j m1_ij: thread 64 h + f holds feature f of
            // rows i = 0..15 summed over the tile halves h (the tensor core delivers them per tile in buffer Z)
            float ssum[kN];
#pragma unroll
            for (int i = 0; i < kN; ++i) ssum[i] = 0.0f;
            auto add_tile_sums = [&] {
                float v[16];
                tc::tmem_ld16(E.tmem + E.lane_base + TM_Z + 16 * (et >> 6), v);
#pragma unroll
                for (int i = 0; i < kN; ++i) ssum[i] += v[i];
            };

            for (int row0 = rbeg; row0 < rend; row0 += cp.rows_per_group) {
                const int nrows = min(cp.rows_per_group, rend - row0);
                const int gpairs = nrows * W;
                const int ntiles = (gpairs + kTile - 1) / kTile;
                // Pair order inside a row group: the pocket neighbours of all its rows first, then the peptide neighbours
                // (they need the relative-position term, a divergent extra step: keeping them together confines it to a few
                // warps).  The output slot stays row-major, rl * W + e.  (row, entry) of this thread's pocket pair advance by
                // 128 pairs per tile without a division.
                const int nvp = W - (L - 1), npocket = nrows * nvp;
                const int adv_q = nvp > 0 ? kTile / nvp : 0, adv_r = nvp > 0 ? kTile - adv_q * nvp : 0;
                int cur_rl = nvp > 0 ? et / nvp : 0, cur_e = nvp > 0 ? et - cur_rl * nvp : 0;
                int slot = 0, nslot = 0;
                auto decode = [&](int t) {
                    PairRef p;
                    const int gp = t * kTile + et;
                    p.active = gp < gpairs;
                    int rl = 0, e = 0;
                    if (gp < npocket) {
                        rl = cur_rl;
                        e = (L - 1) + cur_e;
                        cur_rl += adv_q;
                        cur_e += adv_r;
                        if (cur_e >= nvp) { cur_e -= nvp; ++cur_rl; }
                    } else if (p.active) {
                        const int g2 = gp - npocket;
                        rl = g2 / (L - 1);
                        e = g2 - rl * (L - 1);
                    } else if (nvp == 0) {
                        e = 0;   // no pocket: pair (row0, first peptide neighbour) stands in for the idle lanes
                    } else {
                        e = L - 1;
                    }
                    const int r = row0 + rl;
                    p.i = I[IN_ROWS + r];
                    p.j = e < L - 1 ? I[IN_ROWS + (e < r ? e : e + 1)] : I[IN_POCKET + (e - (L - 1))];
                    nslot = rl * W + e;
                    return p;
                };
                PairRef pr = decode(0);
                slot = nslot;
                if (ntiles > 0) {
                    if (LAYER == 0) write_sel(E, pr, 1.0f);
                    stage_a1<LAYER>(E, pr, b);
                }
                // One tile = a chain of small MMAs issued by the engine's issuing warp on request.  The four heads go through
                // three 64-column TMEM buffers (X, Y, Z), so the tensor core has the next head's contraction in flight while
                // these threads convert the previous one; the pair tile of tile t + 1 is staged under the head contractions.
                for (int t = 0; t < ntiles; ++t) {
                    PMHC_TS();   // 0
                    E.request();           // message layer 2 (+ layer 1: message column sums)
                    uint32_t xa[8];
                    pair_extras(E, pr, xa);
                    PMHC_TS();   // 1 extras
                    E.wait(BM);
                    PMHC_TS();   // 2 D1 ready
                    if (LAYER == 0) {
                        clear_sel(E, pr);
                        add_tile_sums();
                    }
                    epilogue1(E, pr, xa);
                    E.request();           // attention -> X, rotation -> Y, torsion -> Z
                    PMHC_TS();   // 3 ep1
                    PairRef nxt = pr;
                    if (t + 1 < ntiles) {
                        nxt = decode(t + 1);
                        PMHC_TS();   // decode
                        if (LAYER == 0) write_sel(E, nxt, 1.0f);
                        stage_a1<LAYER, false>(E, nxt, b);
                        PMHC_TS();   // stores issued
                        tc::fence_proxy_async_smem();
                    } else { PMHC_TS(); PMHC_TS(); }
                    PMHC_TS();   // 4 next tile staged
                    E.wait(BX);
                    PMHC_TS();   // 5 attention hidden ready
                    const float att = attention_logit(E, TM_X);   // second layer on the CUDA cores; X is free once read
                    PMHC_TS();   // 6
                    E.wait(BY);
                    PMHC_TS();   // 7 rotation hidden ready
                    epilogue2(E, TM_Y);
                    E.request();           // rotation second layer, translation -> X
                    PMHC_TS();   // 8
                    E.wait(BZ);
                    PMHC_TS();   // 9 torsion hidden ready
                    epilogue2(E, TM_Z);
                    E.wait(BX);            // translation hidden ready
                    E.request();           // torsion second layer
                    PMHC_TS();   // 10
                    epilogue2(E, TM_X);
                    E.wait(BZ);
                    E.request();           // translation second layer
                    PMHC_TS();   // 11
                    E.wait(BM);
                    PMHC_TS();   // 12 second layers done
                    epilogue3(E, pr, slot, lsave, att);
                    PMHC_TS();   // 13
                    pr = nxt;
                    slot = nslot;
                }
                E.sync();
                finalize_rows_engine(E, b, row0, nrows, W);
                E.sync();
            }

            if (LAYER == 0) {
                const int npx = kN - L, W2 = cp.msg_w, total = (rend - rbeg) * W2;
                for (int tile_base = 0; tile_base < total; tile_base += kTile) {
                    const int gp0 = tile_base + et;
                    const bool act = gp0 < total;
                    const int gp = act ? gp0 : tile_base;
                    const int rl = gp / W2, e = gp - rl * W2;
                    PairRef pr;
                    pr.i = I[IN_ROWS + rbeg + rl];
                    pr.active = act;
                    float mult = 1.0f;
                    if (e == 0) pr.j = pr.i;
                    else if (e <= npx) pr.j = I[IN_PEPX + e - 1];
                    else if (e <= npx + ci.nx) pr.j = I[IN_POCKET + a.Kpad - 1 - (e - npx - 1)];
                    else {
                        pr.j = -1;
                        const int which = e - (npx + ci.nx + 1);
                        mult = (float)(which == 0 ? min(ci.c0, 256) : ci.c0 - 256);
                    }
                    write_sel(E, pr, mult);
                    stage_a1<LAYER>(E, pr, b);
                    E.request();
                    E.wait(BM);
                    clear_sel(E, pr);
                    add_tile_sums();
                }
                // thread 64 h + f holds the sums of tile half h: add the two halves through shared memory
                float* scr = reinterpret_cast<float*>(E.es + M.Out);
#pragma unroll
                for (int i = 0; i < kN; ++i) scr[((et >> 6) * kN + i) * 64 + (et & 63)] = ssum[i];
                E.sync();
                // this part's real rows; part 0 also zeroes the padded rows
                for (int idx = et; idx < (rend - rbeg) * kHid; idx += kEngThreads) {
                    const int o = I[IN_ROWS + rbeg + (idx >> 6)] * kHid + (idx & 63);
                    a.ssum_out[(size_t)b * kN * kHid + o] = scr[o] + scr[kN * kHid + o];
                }
                if (wk.part == 0)
                    for (int idx = et; idx < npx * kHid; idx += kEngThreads)
                        a.ssum_out[(size_t)b * kN * kHid + I[IN_PEPX + (idx >> 6)] * kHid + (idx & 63)] = 0.0f;
            }
            E.sync();
        }
        PMHC_TS();
        if (ts_on) a.dbg[127] = ts_n;
#undef PMHC_TS
    }

    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// order_kernel — once per batch / trajectory: complexes sorted by their number of attention-carrying pairs
// L (L - 1 + valid pocket slots), largest first (a counting sort by one CTA; equal keys keep their order, so the result is
// deterministic).  The pair kernels deal work items in this order: with mixed peptide lengths and pocket sizes every round
// then holds complexes of similar size, and the small ones end up in the row-split tail.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kOrderBins = kN * (kN - 1 + kMaxP) + 1;
__global__ void __launch_bounds__(1024) order_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ pocket_mask, int B, int P,
                                                     int n_engines, int* __restrict__ keys, int* __restrict__ order) {
    extern __shared__ int bins[];      // [kOrderBins] counts, then starting offsets (descending key)
    __shared__ int carry;
    const int tid = threadIdx.x;
    if (B <= n_engines) {
        // a single round: every complex has an engine to itself whatever the order (the usual training batch), and the
        // serial scan below (30 us) would be a visible part of a training step
        for (int b = tid; b < B; b += blockDim.x) order[b] = b;
        return;
    }
    for (int k = tid; k < kOrderBins; k += blockDim.x) bins[k] = 0;
    __syncthreads();
    for (int b = tid; b < B; b += blockDim.x) {
        int L = 0, nv = 0;
        for (int i = 0; i < kN; ++i) L += mask[(size_t)b * kN + i] != 0;
        for (int j = 0; j < P; ++j) nv += pocket_mask[(size_t)b * P + j] != 0;
        const int key = L * (L - 1 + nv);
        keys[b] = key;
        atomicAdd(&bins[key], 1);
    }
    __syncthreads();
    if (tid == 0) {                    // exclusive scan from the largest key down (7 921 bins: microseconds, once per trajectory)
        int run = 0;
        for (int k = kOrderBins - 1; k >= 0; --k) {
            const int c = bins[k];
            bins[k] = run;
            run += c;
        }
        carry = run;
    }
    __syncthreads();
    if (B <= 4096) {
        // stable placement: complex b goes after the complexes with the same key and a smaller index (O(B) per complex)
        for (int b = tid; b < B; b += blockDim.x) {
            const int key = keys[b];
            int before = 0;
            for (int c = 0; c < b; ++c) before += keys[c] == key;
            order[bins[key] + before] = b;
        }
    } else {
        // very large batches: first come, first placed within a key (only the schedule depends on it, not the results)
        for (int b = tid; b < B; b += blockDim.x) order[atomicAdd(&bins[keys[b]], 1)] = b;
    }
    (void)carry;
}

// ---------------------------------------------------------------------------------------------------------------
// node_pre_kernel — step-invariant first-layer projections, once per batch / sampling trajectory (model.py:401,
// 411-412: pocket nodes carry no time feature; layer 2 sees the pocket's 22 features zero-padded to 64):
//   pk_cache[b][l][p][k] = bf16( W1_l[k, H_l : H_l + 22] . pocket_features[b][p] )
//   cls[b][p]            = 0 valid, 1 masked + all-zero features (one shared message), 2 masked + non-zero features
//   aij1[b][i][0:64]     = bf16( b1 + W1_0[k, 0:22] . features[b][i] )   (A_i of layer 1 without the time term, which the
//   aij1[b][i][64:128]   = bf16(      W1_0[k, 23:45] . features[b][i] )   pair kernel adds per step)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) node_pre_kernel(const float* __restrict__ params, const float* __restrict__ feat,
                                                       const float* __restrict__ pocket_feat, const uint8_t* __restrict__ pocket_mask,
                                                       int P, int cls_stride, __nv_bfloat16* __restrict__ pk_cache,
                                                       uint8_t* __restrict__ cls, __nv_bfloat16* __restrict__ aij1) {
    extern __shared__ __align__(16) float sp[];
    const int b = blockIdx.x, tid = threadIdx.x;
    constexpr int FS = 23;
    float* pf = sp;                                   // [P][23]
    float* w = pf + ((P * FS + 3) & ~3);              // [2][64][23] pocket blocks of both layers
    float* nf = w + 2 * kHid * FS;                    // [16][23] peptide features
    float* wp = nf + kN * FS + 1;                     // [128][23] layer-1 A_i | A_j blocks
    for (int idx = tid; idx < P * PMHC_NFEAT; idx += blockDim.x) {
        int j = idx / PMHC_NFEAT, c = idx - j * PMHC_NFEAT;
        pf[j * FS + c] = pocket_feat[(size_t)b * P * PMHC_NFEAT + idx];
    }
    for (int idx = tid; idx < 2 * kHid * PMHC_NFEAT; idx += blockDim.x) {
        int l = idx / (kHid * PMHC_NFEAT), r = idx - l * kHid * PMHC_NFEAT;
        int k = r / PMHC_NFEAT, c = r - k * PMHC_NFEAT;
        const int H = l == 0 ? kH1 : kH2, ld1 = 2 * H + kEdge;
        w[(l * kHid + k) * FS + c] = params[param_offset(l, MSG0_W) + k * ld1 + H + c];
    }
    for (int idx = tid; idx < kN * PMHC_NFEAT; idx += blockDim.x) {
        int i = idx / PMHC_NFEAT, c = idx - i * PMHC_NFEAT;
        nf[i * FS + c] = feat[(size_t)b * kN * PMHC_NFEAT + idx];
    }
    for (int idx = tid; idx < 128 * PMHC_NFEAT; idx += blockDim.x) {
        int k = idx / PMHC_NFEAT, c = idx - k * PMHC_NFEAT;
        constexpr int ld1 = 2 * kH1 + kEdge;
        wp[k * FS + c] = params[param_offset(0, MSG0_W) + (k & 63) * ld1 + (k < 64 ? 0 : kH1) + c];
    }
    __syncthreads();
    for (int idx = tid; idx < 2 * P * (kHid / 8); idx += blockDim.x) {
        const int l = idx / (P * 8), r = idx - l * P * 8;
        const int p = r >> 3, k0 = (r & 7) * 8;
        const float* h = pf + p * FS;
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.0f;
#pragma unroll
        for (int c = 0; c < PMHC_NFEAT; ++c) {
            const float hv = h[c];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = fmaf(w[(l * kHid + k0 + u) * FS + c], hv, acc[u]);
        }
        uint4 o;
        o.x = tc::pack_bf16x2(acc[0], acc[1]); o.y = tc::pack_bf16x2(acc[2], acc[3]);
        o.z = tc::pack_bf16x2(acc[4], acc[5]); o.w = tc::pack_bf16x2(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(pk_cache + (((size_t)b * 2 + l) * P + p) * kHid + k0) = o;
    }
    for (int j = tid; j < P; j += blockDim.x) {
        uint8_t c = 0;
        if (pocket_mask[(size_t)b * P + j] == 0) {
            bool nz = false;
            for (int q = 0; q < PMHC_NFEAT; ++q) nz |= (pf[j * FS + q] != 0.0f);
            c = nz ? 2 : 1;
        }
        cls[(size_t)b * cls_stride + j] = c;
    }
    for (int idx = tid; idx < kN * 64; idx += blockDim.x) {
        const int i = idx >> 6, k = 2 * (idx & 63);
        float acc[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            acc[u] = k + u < 64 ? params[param_offset(0, MSG0_B) + k + u] : 0.0f;
#pragma unroll
            for (int c = 0; c < PMHC_NFEAT; ++c) acc[u] = fmaf(wp[(k + u) * FS + c], nf[i * FS + c], acc[u]);
        }
        reinterpret_cast<uint32_t*>(aij1)[(size_t)b * kN * 64 + idx] = tc::pack_bf16x2(acc[0], acc[1]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// node_mid_kernel — between the layers, 128 nodes (8 complexes) per CTA, four chained tensor-core GEMMs with the
// A operand in tensor memory (thread = node = TMEM lane):
//   Msum = S . W2^T + (16 + P) b2                  (S = sum_j m1_ij from layer 1; hi/lo bf16 split, K = 128)
//   hid  = relu([Msum | h | t/T] . W_f0^T + b)     (feature_mlp.0, model.py:151)
//   o1   = relu(hid . W_f2^T + b)                  (feature_mlp.2 and the ReLU of model.py:407)
//   [A_i | A_j] = o1 . [W1_i ; W1_j]^T (+ b1)      (layer 2's message_mlp.0 peptide blocks) -> bf16
// ---------------------------------------------------------------------------------------------------------------
struct NodeMidArgs {
    const uint8_t* image;     // node_mid_image_kernel's output
    int B, P;
    float t_over_T;
    const float* ssum;        // [B,16,64]
    const float* feat;        // [B,16,22]
    const uint8_t* mask;      // [B,16]
    __nv_bfloat16* aij2;      // [B,16,128]
    float* feat1_out;         // nullable: [B,16,64] relu(o1)   (saved for the backward pass)
    float* msum_out;          // nullable: [B,16,64]
    const float* t_dev;       // nullable: t / T in device memory (see time_feature)
};
constexpr int NM_W2 = 0, NM_WF0 = 8192, NM_WF2 = 24576, NM_W1 = 32768, NM_BIAS = 49152, NM_BAR = 50432, NM_TPTR = 50448,
              NM_BYTES = 50464 + 1024;

// operand tiles + biases of node_mid_kernel in its shared-memory layout ([0, NM_BAR)), built once per batch / trajectory
__global__ void __launch_bounds__(256) node_mid_image_kernel(const float* __restrict__ params, int P, uint8_t* __restrict__ smem) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    float* bias = reinterpret_cast<float*>(smem + NM_BIAS);   // [0,64) b2 * (16 + P) | f0b | f2b | b1(layer 2) | 64 zeros
    const float* msg2 = params + param_offset(0, MSG2_W);
    const float* f0w = params + param_offset(0, FEAT0_W);   // [64][23 + 64]: columns 0..22 node features (+time), 23.. message sum
    const float* f2w = params + param_offset(0, FEAT2_W);
    const float* w1 = params + param_offset(1, MSG0_W);     // [64][2*64 + 31]
    constexpr int ldf = kH1 + kHid, ld1 = 2 * kH2 + kEdge;
    for (int idx = tid; idx < 64 * 32; idx += nthr) {
        const int n = idx >> 5, k = (idx & 31) * 2;
        *reinterpret_cast<uint32_t*>(smem + NM_W2 + tc::sw128_offset(n, k)) = tc::pack_bf16x2(msg2[n * 64 + k], msg2[n * 64 + k + 1]);
        *reinterpret_cast<uint32_t*>(smem + NM_WF2 + tc::sw128_offset(n, k)) = tc::pack_bf16x2(f2w[n * 64 + k], f2w[n * 64 + k + 1]);
        // feature_mlp.0, K block 0: the 64 message-sum columns
        *reinterpret_cast<uint32_t*>(smem + NM_WF0 + tc::sw128_offset(n, k)) =
            tc::pack_bf16x2(f0w[n * ldf + kH1 + k], f0w[n * ldf + kH1 + k + 1]);
        // K block 1: 22 features, t/T hi, t/T lo (same weight), zeros
        float v0 = 0.0f, v1 = 0.0f;
        if (k < PMHC_NFEAT) { v0 = f0w[n * ldf + k]; v1 = f0w[n * ldf + k + 1]; }        // k even, k + 1 <= 21
        else if (k == PMHC_NFEAT) v0 = v1 = f0w[n * ldf + PMHC_NFEAT];                   // columns 22, 23: t/T hi, t/T lo
        *reinterpret_cast<uint32_t*>(smem + NM_WF0 + 8192 + tc::sw128_offset(n, k)) = tc::pack_bf16x2(v0, v1);
    }
    for (int idx = tid; idx < 128 * 32; idx += nthr) {
        const int n = idx >> 5, k = (idx & 31) * 2;
        const float* src = w1 + (n & 63) * ld1 + (n < 64 ? 0 : kH2) + k;
        *reinterpret_cast<uint32_t*>(smem + NM_W1 + tc::sw128_offset(n, k)) = tc::pack_bf16x2(src[0], src[1]);
    }
    if (tid < 64) {
        bias[tid] = params[param_offset(0, MSG2_B) + tid] * (float)(kN + P);
        bias[64 + tid] = params[param_offset(0, FEAT0_B) + tid];
        bias[128 + tid] = params[param_offset(0, FEAT2_B) + tid];
        bias[192 + tid] = params[param_offset(1, MSG0_B) + tid];
        bias[256 + tid] = 0.0f;
    }
}

__global__ void __launch_bounds__(128, 1) node_mid_kernel(NodeMidArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NM_BAR);
    const float* bias = reinterpret_cast<const float*>(smem + NM_BIAS);
    if (warp == 0) tc::tmem_alloc(reinterpret_cast<uint32_t*>(smem + NM_TPTR), 512);
    if (tid == 32) {
        tc::mbar_init(bar, 1);
        tc::mbar_fence_init();
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.image);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int idx = tid; idx < NM_BAR / 16; idx += 128) tc::cp_async_16(dst + idx, src + idx);   // all pieces in flight at once
        tc::cp_async_wait_all();
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + NM_TPTR);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t phase = 0;
    // TMEM columns: A_S [0,64) | D [64,128) | A_f [128,176) | D [192,256) | A_hid [256,288) | D [320,384) | A_o1 [384,416) | D4 [0,128)
    const int64_t node = (int64_t)blockIdx.x * 128 + tid;
    const bool in = node < (int64_t)a.B * kN;
    const bool real = in && a.mask[node] != 0;

    {   // S -> hi / lo bf16 (K = 128) ; node features + time -> K block 1 of the feature MLP operand
        uint32_t hi[32], lo[32];
        const float4* src = reinterpret_cast<const float4*>(a.ssum + (in ? node : 0) * kHid);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float4 v = in ? __ldg(src + c) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            const float h0 = tc::bf16_round(v.x), h1 = tc::bf16_round(v.y), h2 = tc::bf16_round(v.z), h3 = tc::bf16_round(v.w);
            hi[2 * c] = tc::pack_bf16x2(h0, h1); hi[2 * c + 1] = tc::pack_bf16x2(h2, h3);
            lo[2 * c] = tc::pack_bf16x2(v.x - h0, v.y - h1); lo[2 * c + 1] = tc::pack_bf16x2(v.z - h2, v.w - h3);
        }
        tc::tmem_st32(tmem + lane_base + 0, hi);
        tc::tmem_st32(tmem + lane_base + 32, lo);
        uint32_t hf[16];
        const float* f = a.feat + (in ? node : 0) * PMHC_NFEAT;
        const float tt = time_feature(a);
        const float th = tc::bf16_round(tt);
#pragma unroll
        for (int c = 0; c < 11; ++c) hf[c] = in ? tc::pack_bf16x2(f[2 * c], f[2 * c + 1]) : 0u;
        hf[11] = tc::pack_bf16x2(th, tt - th);
#pragma unroll
        for (int c = 12; c < 16; ++c) hf[c] = 0u;
        tc::tmem_st16(tmem + lane_base + 160, hf);
    }
    auto issue = [&](auto&& f) {
        tc::tmem_wait_st();
        tc::fence_before_thread_sync();
        __syncthreads();
        if (warp == 0) {
            if (tc::elect_one()) {
                tc::fence_after_thread_sync();
                f();
                tc::mma_commit(bar);
            }
            __syncwarp();
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1;
        tc::fence_after_thread_sync();
    };
    constexpr uint32_t id64 = tc::idesc_bf16_f32(128, 64), id128 = tc::idesc_bf16_f32(128, 128);
    // G1: Msum
    issue([&] {
        const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + NM_W2));
#pragma unroll
        for (int s = 0; s < 8; ++s) tc::mma_bf16_ts(tmem + 64, tmem + 8 * s, db + 2 * (s & 3), id64, s > 0);
    });
    // epilogue with a per-column bias: D columns [col, col+64) -> out[64] (fp32), optional relu
    auto load64 = [&](int col, const float* bv, bool relu, float (&out)[64]) {
        uint32_t v0[32], v1[32];
        tc::tmem_ld32_nowait(tmem + lane_base + col, v0);
        tc::tmem_ld32_nowait(tmem + lane_base + col + 32, v1);
        tc::tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            float x0 = __uint_as_float(v0[c]) + bv[c], x1 = __uint_as_float(v1[c]) + bv[32 + c];
            out[c] = relu ? fmaxf(x0, 0.0f) : x0;
            out[32 + c] = relu ? fmaxf(x1, 0.0f) : x1;
        }
    };
    auto store_bf16 = [&](int col, const float (&v)[64]) {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) pk[c] = tc::pack_bf16x2(v[2 * c], v[2 * c + 1]);
        tc::tmem_st32(tmem + lane_base + col, pk);
    };
    auto store_global = [&](float* dst, const float (&v)[64], bool keep) {
        if (dst == nullptr || !in) return;
        float4* d4 = reinterpret_cast<float4*>(dst + node * kHid);
#pragma unroll
        for (int c = 0; c < 16; ++c)
            d4[c] = keep ? make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    };
    float v[64];
    load64(64, bias + 0, false, v);
    store_global(a.msum_out, v, real);
    store_bf16(128, v);
    // G2: feature_mlp.0 over [Msum (64) | features, time (32)]
    issue([&] {
        const uint64_t d0 = tc::smem_desc_sw128(tc::smem_u32(smem + NM_WF0));
        const uint64_t d1 = tc::smem_desc_sw128(tc::smem_u32(smem + NM_WF0 + 8192));
#pragma unroll
        for (int s = 0; s < 4; ++s) tc::mma_bf16_ts(tmem + 192, tmem + 128 + 8 * s, d0 + 2 * s, id64, s > 0);
#pragma unroll
        for (int s = 0; s < 2; ++s) tc::mma_bf16_ts(tmem + 192, tmem + 160 + 8 * s, d1 + 2 * s, id64, 1);
    });
    load64(192, bias + 64, true, v);
    store_bf16(256, v);
    // G3: feature_mlp.2 (+ the ReLU between the layers)
    issue([&] {
        const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + NM_WF2));
#pragma unroll
        for (int s = 0; s < 4; ++s) tc::mma_bf16_ts(tmem + 320, tmem + 256 + 8 * s, db + 2 * s, id64, s > 0);
    });
    load64(320, bias + 128, true, v);
    store_global(a.feat1_out, v, real);
    store_bf16(384, v);
    // G4: layer 2's A_i | A_j
    issue([&] {
        const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + NM_W1));
#pragma unroll
        for (int s = 0; s < 4; ++s) tc::mma_bf16_ts(tmem + 0, tmem + 384 + 8 * s, db + 2 * s, id128, s > 0);
    });
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        load64(64 * half, bias + (half == 0 ? 192 : 256), false, v);
        if (in) {
            uint4* dst = reinterpret_cast<uint4*>(a.aij2 + node * 128 + 64 * half);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint4 o;
                o.x = tc::pack_bf16x2(v[8 * c], v[8 * c + 1]); o.y = tc::pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
                o.z = tc::pack_bf16x2(v[8 * c + 4], v[8 * c + 5]); o.w = tc::pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
                dst[c] = o;
            }
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace tc2

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
long long* g_tc2_dbg = nullptr;   // development hook (pmhc_debug_set_stamps)

struct Tc2Workspace {
    __nv_bfloat16* pk_cache;   // [B,2,P,64]
    float* ssum;               // [B,16,64]
    __nv_bfloat16* aij1;       // [B,16,128]
    __nv_bfloat16* aij2;       // [B,16,128]
    uint8_t* cls;              // [B,cls_stride]
    uint8_t* wimage;           // [2][image_bytes]
    uint8_t* nm_image;         // node_mid operand tiles
    int* order;                // [B] work order of the pair kernels
    int* keys;                 // [B] scratch of order_kernel
    int cls_stride, image_bytes;
    size_t bytes;
};
Tc2Workspace carve_tc2(void* base, int B, int P) {
    Tc2Workspace w;
    uint8_t* p = (uint8_t*)base;
    size_t o = 0;
    auto take = [&](size_t n) { size_t at = o; o += (n + 255) & ~(size_t)255; return p + at; };
    w.cls_stride = (P + 15) & ~15;
    w.image_bytes = tc2::make_map(32, 256, kN).Bar;
    w.pk_cache = (__nv_bfloat16*)take((size_t)B * 2 * P * kHid * 2);
    w.ssum = (float*)take((size_t)B * kN * kHid * 4);
    w.aij1 = (__nv_bfloat16*)take((size_t)B * kN * 128 * 2);
    w.aij2 = (__nv_bfloat16*)take((size_t)B * kN * 128 * 2);
    w.cls = (uint8_t*)take((size_t)B * w.cls_stride);
    w.wimage = (uint8_t*)take((size_t)2 * w.image_bytes);
    w.nm_image = (uint8_t*)take((size_t)tc2::NM_BAR);
    w.order = (int*)take((size_t)B * sizeof(int));
    w.keys = (int*)take((size_t)B * sizeof(int));
    w.bytes = o;
    return w;
}
size_t tc2_workspace_bytes(int B, int P) { return carve_tc2(nullptr, B, P).bytes; }

template <int LAYER>
static int launch_pair(tc2::PairArgs& a, cudaStream_t stream) {
    static PerDeviceOnce configured;
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // buffer as many pairs per row group as fit (fewer partial tiles); keep the pocket rows of A_j in shared memory
    // when there is room, else read them from L2
    const int W_max = kN - 1 + a.P;
    int cap = 0, aj_rows = kN;
    const int caps[4] = {640, 512, 384, 256};
    for (int with_aj = 1; with_aj >= 0 && cap == 0; --with_aj)
        for (int c = 0; c < 4 && cap == 0; ++c) {
            const int cc = caps[c] > W_max ? caps[c] : ((W_max + 127) / 128) * 128;
            const int rows = with_aj ? kN + a.P : kN;
            if ((size_t)tc2::make_map(a.Kpad, cc, rows).total_bytes + 1024 <= (size_t)max_smem) { cap = cc; aj_rows = rows; }
        }
    PMHC_REQUIRE(cap > 0, "EGNN tensor-core layer does not fit in shared memory (P=%d, device allows %d B)", a.P, max_smem);
    a.cap_pairs = cap;
    a.aj_rows = aj_rows;
    const tc2::Map M = tc2::make_map(a.Kpad, cap, aj_rows);
    const size_t smem = (size_t)M.total_bytes + 1024;
    if (configured.needed()) {
        cudaError_t e = cudaFuncSetAttribute(tc2::egnn_pair_tc_kernel<LAYER>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        PMHC_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(pair_tc): %s", cudaGetErrorString(e));
        configured.mark();
    }
    const int want = (a.B + tc2::kEngines - 1) / tc2::kEngines;
    const int grid = want < num_sms() ? want : num_sms();
    if (profile_enabled()) profile_mark(PROF_FWD, stream, true);
    tc2::egnn_pair_tc_kernel<LAYER><<<grid, tc2::kThreads, smem, stream>>>(a);
    if (profile_enabled()) profile_mark(PROF_FWD, stream, false);
    PMHC_CHECK_LAUNCH("egnn_pair_tc");
    return 0;
}

// Denoiser forward in bf16 tensor-core mode: [node_pre] -> pair<0> -> node_mid -> pair<1>.
int forward_tc2(const float* params, const PmhcBatch* bt, float t_over_T, float* frames1, float* tors1, float* out_frames,
                float* out_torsions, float* feat1_out, float* msum_out, float* rowstat1, float* rowstat2, float* logits1,
                float* logits2, void* tc2_ws, cudaStream_t stream, bool reuse_pocket_cache) {
    const int B = bt->B, P = bt->P;
    Tc2Workspace w = carve_tc2(tc2_ws, B, P);
    if (!reuse_pocket_cache) {
        // step-invariant: operand tiles of both layers, pocket projections, static peptide projections
        tc2::weight_image_kernel<0><<<32, 256, 0, stream>>>(params, w.wimage);
        PMHC_CHECK_LAUNCH("weight_image");
        tc2::weight_image_kernel<1><<<32, 256, 0, stream>>>(params, w.wimage + w.image_bytes);
        PMHC_CHECK_LAUNCH("weight_image");
        tc2::node_mid_image_kernel<<<32, 256, 0, stream>>>(params, P, w.nm_image);
        PMHC_CHECK_LAUNCH("node_mid_image");
        const size_t smem = (size_t)(((P * 23 + 3) & ~3) + 2 * kHid * 23 + kN * 23 + 1 + 128 * 23) * sizeof(float);
        static PerDeviceOnce configured;
        if (configured.needed()) {
            cudaError_t e = cudaFuncSetAttribute(tc2::node_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            PMHC_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(node_pre): %s", cudaGetErrorString(e));
            configured.mark();
        }
        tc2::node_pre_kernel<<<B, 256, smem, stream>>>(params, bt->features, bt->pocket_features, bt->pocket_mask, P, w.cls_stride,
                                                       w.pk_cache, w.cls, w.aij1);
        PMHC_CHECK_LAUNCH("node_pre");
        tc2::order_kernel<<<1, 1024, tc2::kOrderBins * sizeof(int), stream>>>(bt->mask, bt->pocket_mask, B, P, 2 * num_sms(), w.keys, w.order);
        PMHC_CHECK_LAUNCH("order");
    }
    tc2::PairArgs a{};
    a.B = B; a.P = P; a.Kpad = pad_k(P);
    a.t_over_T = t_over_T;
    a.t_dev = step_t_dev();
    a.frames_in = bt->frames; a.tors_in = bt->torsions; a.mask = bt->mask;
    a.pocket_frames = bt->pocket_frames; a.pocket_cls = w.cls; a.cls_stride = w.cls_stride; a.pk_cache = w.pk_cache;
    a.aij = w.aij1; a.wimage = w.wimage;
    a.order = w.order;
    a.dbg = g_tc2_dbg;
    a.frames_out = frames1; a.tors_out = tors1; a.ssum_out = w.ssum;
    a.rowstat = rowstat1; a.logit_out = logits1;
    int rc = launch_pair<0>(a, stream);
    if (rc != 0) return rc;
    {
        static PerDeviceOnce configured;
        if (configured.needed()) {
            cudaError_t e = cudaFuncSetAttribute(tc2::node_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::NM_BYTES);
            PMHC_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(node_mid): %s", cudaGetErrorString(e));
            configured.mark();
        }
        tc2::NodeMidArgs n{w.nm_image, B, P, t_over_T, w.ssum, bt->features, bt->mask, w.aij2, feat1_out, msum_out, step_t_dev()};
        const int grid = (B * kN + 127) / 128;
        tc2::node_mid_kernel<<<grid, 128, tc2::NM_BYTES, stream>>>(n);
        PMHC_CHECK_LAUNCH("node_mid");
    }
    a.frames_in = frames1; a.tors_in = tors1;
    a.aij = w.aij2; a.wimage = w.wimage + w.image_bytes;
    a.dbg = g_tc2_dbg ? g_tc2_dbg + 128 : nullptr;
    a.frames_out = out_frames; a.tors_out = out_torsions; a.ssum_out = nullptr;
    a.rowstat = rowstat2; a.logit_out = logits2;
    return launch_pair<1>(a, stream);
}

}  // namespace pmhc

// development hook: device buffer of 256 int64 receiving clock64 stamps of CTA 0 / thread 0 (layer 1: [0,128), layer 2: [128,256))
extern "C" void pmhc_debug_set_stamps(long long* dev_buf) { pmhc::g_tc2_dbg = dev_buf; }
