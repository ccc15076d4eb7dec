// egnn_backward.cu — fused EGNN layer backward: weight gradients of every MLP of the layer and (layer 2) the
// gradients w.r.t. the layer's input frames / torsions / node features.  This is the autograd of
// EGNNLayer.forward (diffusion/model.py:83-333) that `total_loss.mean().backward()` (optimizer.py:222) runs
// as ~700 ATen backward ops in the reference.
//
// Same persistent one-CTA-per-SM structure as the forward.  One thread owns one (i, j) pair per pass of 128
// pairs: it recomputes the pair's forward (message, head hidden layers) in registers from the staged weights,
// back-propagates through the heads / softmax / quaternion sandwich in registers, and leaves its activation
// and gradient columns in two shared [64][128] tiles; after each head the CTA turns those tiles into the
// weight-gradient outer products (a 64x64x128 register-tiled GEMM per matrix) accumulated in a per-CTA
// partial buffer, so no gradient needs an atomic on global memory and the result is run-to-run deterministic.
// Softmax statistics and logits saved by the forward make the per-pair backward single-pass.
#include "egnn_common.cuh"

namespace pmhc {

constexpr int kLdc = 132;                     // column-tile row stride (floats): 16-byte aligned rows
constexpr int kLdt = 136;                     // same in tensor-core mode: 8 g + t is a conflict-free bank pattern for MMA fragments
// tensor-core mode: second-layer weights as swizzled [c][n] row images (rotation 4 rows, torsion 8 (7 used), translation 1,
// attention 1) and the extra-input columns of the first layers as [e][n] images (rotation: local quaternion 4, attention: 2)
constexpr int kW3Rot = 0, kW3Tor = 4 * kHid, kW3Trn = 12 * kHid, kW3Att = 13 * kHid, kW3Floats = 14 * kHid;
constexpr int kWxRot = 0, kWxAtt = 4 * kHid, kWxFloats = 6 * kHid;
constexpr int kTileFloats = 5 * kHid * kHid;  // tile-owner-layout partials of the five 64x64 matrices
enum { T_W2 = 0, T_ATT = 1, T_ROT = 2, T_TOR = 3, T_TRN = 4 };
enum { HD_ATT = 0, HD_ROT = 1, HD_TOR = 2, HD_TRN = 3 };

struct BwdArgs {
    LayerArgs a;                // the forward's inputs of this layer (outputs unused)
    const float* rowstat;       // [B,16,16] saved by the forward
    const float* logits;        // [B,16,Kpad] saved by the forward
    const float* msum;          // layer 1: [B,16,64] saved unmasked message sums
    const float* feat_post;     // layer 1: [B,16,64] relu(o1) (for the relu mask)
    const float* d_frames_out;  // [B,16,7]   dL / d (unit quaternion, translation) of this layer's output
    const float* d_tors_out;    // [B,16,14]
    const float* d_feat_out;    // layer 1: [B,16,64] dL / d relu(o1)
    float* d_frames_in;         // layer 2: [B,16,7]  dL / d layer inputs
    float* d_tors_in;           // layer 2: [B,16,14]
    float* d_feat_in;           // layer 2: [B,16,64]
    float* partial;             // [gridDim][kTileFloats + layer params] per-CTA gradient partial sums
    float* dajt_ws;             // [gridDim][64][Kpad] per-CTA scratch: dL / d A_j^T
    int partial_stride;
};

struct BwdMap {
    SmemMap f;  // the fields setup_complex() uses (Scr, Ai, Tt, Msum, H, Tors, Q, X, Ints) and the packs
    int W2, Wh, BufA, BufB, Dout, Ex;
    int Pl;
    int W3i, Wx, Dx;            // tensor-core mode: second-layer weight images, extra-input weight images, per-pair extras gradients
    int dAi, dAjPep, dWe, dTt, dMsum, RowG, dQ, dX, dTors, grads_end;
    int total_floats;
};

__host__ __device__ inline BwdMap make_bwd_map(int Kpad, bool tc = false) {
    BwdMap m;
    int o = 0;
    const int ldc = tc ? kLdt : kLdc;
    m.W2 = o;       o += kHid * kHid;        // message_mlp.2.weight as stored: [n][k]
    m.Wh = o;       o += 4 * kHid * kHid;    // head first layers, message columns only: [head][n][k]
    m.f.W2T = m.f.WhT = m.f.We = -1;
    m.f.PkAtt = o;  o += 4 * kHid;
    m.f.PkRotQ = o; o += 4 * kHid;
    m.f.PkMisc = o; o += 4 * kHid;
    m.W3i = m.Wx = m.Dx = -1;
    if (tc) {
        m.f.PkRot2 = m.f.PkTor2 = -1;
        m.W3i = o;  o += kW3Floats;
        m.Wx = o;   o += kWxFloats;
    } else {
        m.f.PkRot2 = o; o += 4 * kHid;
        m.f.PkTor2 = o; o += 8 * kHid;
    }
    m.f.Scal = o;   o += 16;
    m.BufA = o;     o += kHid * ldc;
    m.BufB = o;     o += kHid * ldc;
    m.f.Scr = m.BufA;                         // setup_complex stages pocket features in BufA..BufB
    m.f.Out = -1;
    m.Dout = o;     o += 8 * ldc;
    m.Ex = o;       o += (tc ? 10 : 6) * ldc;   // tensor-core mode: the per-pair geometry tile (8 MMA rows + 2 fp32 rows)
    if (tc) { m.Dx = o; o += 6 * ldc; }          // dL / d (local quaternion (4), -d2, qdot2) per pair column
    m.Pl = -1;
    if (tc) { m.Pl = o; o += kBwdPairs + 4; }    // list of the pass's peptide-neighbour pair columns (+ its length)
    m.f.Ai = o;     o += kN * kLdN;
    o = (o + 3) & ~3;
    m.f.Tt = o;     o += kN * kHid;
    m.f.Msum = o;   o += kN * kHid;
    m.f.H = o;      o += kN * kLdN;
    o = (o + 3) & ~3;
    m.f.Tors = o;   o += kN * 2 * PMHC_NTORS;
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
    return m;
}

// tensor-core mode: the five 64 x 64 matrices are stored tf32-rounded with an XOR swizzle of the column index, so that both
// A[m][k] = W[m][k] (forward) and A[m][k] = W[k][m] (input gradients) fragment loads of mma.m16n8k8 are bank-conflict free
// without padding: element (n, k) lives at n * 64 + (k ^ (((n & 3) << 3) | (n & 4))).
__device__ __forceinline__ int wswz(int n, int k) { return n * 64 + (k ^ (((n & 3) << 3) | (n & 4))); }
// The five 64 x 64 matrices additionally store column k at position kperm(k) inside its block of 8 — MMA k-slots (t, t + 4) of a
// k-step sit next to each other — so the forward fragments (a0, a2) are one 64-bit load.
__device__ __forceinline__ int kperm(int k) { return (k & ~7) | ((k & 3) << 1) | ((k >> 2) & 1); }
__device__ __forceinline__ int wswzp(int n, int k) { return wswz(n, kperm(k)); }
__device__ __forceinline__ float tf32r(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

template <int LAYER, bool TC = false>
__device__ inline void stage_layer_weights_bwd(float* S, const BwdMap& M, const float* __restrict__ params) {
    constexpr int L = LAYER;
    const int tid = threadIdx.x;
    const float* msg2 = params + param_offset(L, MSG2_W);
    const float* att0 = params + param_offset(L, ATT0_W);
    const float* rot0 = params + param_offset(L, ROT0_W);
    const float* tor0 = params + param_offset(L, TOR0_W);
    const float* trn0 = params + param_offset(L, TRN0_W);
#pragma unroll 8
    for (int idx = tid; idx < kHid * kHid; idx += kBwdThreads) {
        if (TC) {
            const int o = wswzp(idx >> 6, idx & 63);
            S[M.W2 + o] = tf32r(msg2[idx]);
            S[M.Wh + HD_TRN * 4096 + o] = tf32r(trn0[idx]);
        } else {
            S[M.W2 + idx] = msg2[idx];
            S[M.Wh + HD_TRN * 4096 + idx] = trn0[idx];
        }
    }
#pragma unroll 8
    for (int idx = tid; idx < kHid * 66; idx += kBwdThreads) {
        int n = idx / 66, k = idx - n * 66;
        float v = att0[idx];
        if (k < 64) S[M.Wh + HD_ATT * 4096 + (TC ? wswzp(n, k) : n * 64 + k)] = TC ? tf32r(v) : v;
        else S[M.f.PkAtt + 4 * n + (k - 64)] = v;
    }
#pragma unroll 8
    for (int idx = tid; idx < kHid * 68; idx += kBwdThreads) {
        int n = idx / 68, k = idx - n * 68;
        float v = rot0[idx];
        if (k < 64) S[M.Wh + HD_ROT * 4096 + (TC ? wswzp(n, k) : n * 64 + k)] = TC ? tf32r(v) : v;
        else S[M.f.PkRotQ + 4 * n + (k - 64)] = v;
    }
#pragma unroll 8
    for (int idx = tid; idx < kHid * 78; idx += kBwdThreads) {
        int n = idx / 78, k = idx - n * 78;
        if (k < 64) S[M.Wh + HD_TOR * 4096 + (TC ? wswzp(n, k) : n * 64 + k)] = TC ? tf32r(tor0[idx]) : tor0[idx];
    }
    for (int n = tid; n < kHid; n += kBwdThreads) {
        S[M.f.PkAtt + 4 * n + 2] = params[param_offset(L, ATT0_B) + n];
        S[M.f.PkAtt + 4 * n + 3] = params[param_offset(L, ATT2_W) + n];
        S[M.f.PkMisc + 4 * n + 0] = params[param_offset(L, TRN0_B) + n];
        S[M.f.PkMisc + 4 * n + 1] = params[param_offset(L, TRN2_W) + n];
        S[M.f.PkMisc + 4 * n + 2] = params[param_offset(L, ROT0_B) + n];
        S[M.f.PkMisc + 4 * n + 3] = params[param_offset(L, MSG2_B) + n];
        if (!TC) S[M.f.PkTor2 + 8 * n + 7] = 0.0f;
        if (TC) {
            S[M.W3i + kW3Tor + wswz(7, n)] = 0.0f;
            S[M.W3i + kW3Trn + wswz(0, n)] = tf32r(params[param_offset(L, TRN2_W) + n]);
            S[M.W3i + kW3Att + wswz(0, n)] = tf32r(params[param_offset(L, ATT2_W) + n]);
            for (int e = 0; e < 4; ++e) S[M.Wx + kWxRot + wswz(e, n)] = tf32r(rot0[n * 68 + 64 + e]);
            for (int e = 0; e < 2; ++e) S[M.Wx + kWxAtt + wswz(e, n)] = tf32r(att0[n * 66 + 64 + e]);
        }
    }
    for (int idx = tid; idx < 4 * kHid; idx += kBwdThreads) {
        int c = idx >> 6, n = idx & 63;
        if (TC) S[M.W3i + kW3Rot + wswz(c, n)] = tf32r(params[param_offset(L, ROT2_W) + idx]);
        else S[M.f.PkRot2 + 4 * n + c] = params[param_offset(L, ROT2_W) + idx];
    }
    for (int idx = tid; idx < PMHC_NTORS * kHid; idx += kBwdThreads) {
        int c = idx >> 6, n = idx & 63;
        if (TC) S[M.W3i + kW3Tor + wswz(c, n)] = tf32r(params[param_offset(L, TOR2_W) + idx]);
        else S[M.f.PkTor2 + 8 * n + c] = params[param_offset(L, TOR2_W) + idx];
    }
    if (tid == 0) {
        S[M.f.Scal + SC_ATT2B] = params[param_offset(L, ATT2_B)];
        S[M.f.Scal + SC_TRN2B] = params[param_offset(L, TRN2_B)];
        for (int c = 0; c < 4; ++c) S[M.f.Scal + SC_ROT2B + c] = params[param_offset(L, ROT2_B) + c];
        for (int c = 0; c < PMHC_NTORS; ++c) S[M.f.Scal + SC_TOR2B + c] = params[param_offset(L, TOR2_B) + c];
    }
}

// sum_k wrow[k] * v[k], wrow a warp-uniform shared row; four independent chains
__device__ __forceinline__ float dot64(const float* __restrict__ wrow, const float (&v)[kHid]) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
    for (int k4 = 0; k4 < kHid / 4; ++k4) {
        float4 w = w4[k4];
        s0 = fmaf(w.x, v[4 * k4 + 0], s0);
        s1 = fmaf(w.y, v[4 * k4 + 1], s1);
        s2 = fmaf(w.z, v[4 * k4 + 2], s2);
        s3 = fmaf(w.w, v[4 * k4 + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
// acc[k] += wrow[k] * a
__device__ __forceinline__ void axpy64(float (&acc)[kHid], const float* __restrict__ wrow, float a) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
#pragma unroll
    for (int k4 = 0; k4 < kHid / 4; ++k4) {
        float4 w = w4[k4];
        acc[4 * k4 + 0] = fmaf(w.x, a, acc[4 * k4 + 0]);
        acc[4 * k4 + 1] = fmaf(w.y, a, acc[4 * k4 + 1]);
        acc[4 * k4 + 2] = fmaf(w.z, a, acc[4 * k4 + 2]);
        acc[4 * k4 + 3] = fmaf(w.w, a, acc[4 * k4 + 3]);
    }
}

// L2 loads the compiler must not sink to their use: issued where written (volatile), so their round trip overlaps the
// barrier / MMA work in between
__device__ __forceinline__ float4 ldcg_early4(const float* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldcg_early(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// tile[n][k] += sum_p bufN[n][p] * bufK[k][p] over the 128 columns of the pass: register-tiled 64x64x128 GEMM.
// Thread t owns k in {t&15 + 16a}, n in {t>>4 + 16b}, a, b < 4; its 16 sums live contiguously in the tile-owner layout.
__device__ __forceinline__ void coop_outer(const float* __restrict__ bufN, const float* __restrict__ bufK,
                                           float* __restrict__ tile) {
    const int tid = threadIdx.x, kk = tid & 15, nn = tid >> 4;
    // the running sums (L2) are loaded first and consumed after the products, so their round trip hides under them
    float4 old[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) old[a] = ldcg_early4(tile + tid * 16 + 4 * a);
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
#pragma unroll 2
    for (int p4 = 0; p4 < 32; ++p4) {
        float4 kv[4], nv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) kv[a] = *reinterpret_cast<const float4*>(bufK + (kk + 16 * a) * kLdc + 4 * p4);
#pragma unroll
        for (int b = 0; b < 4; ++b) nv[b] = *reinterpret_cast<const float4*>(bufN + (nn + 16 * b) * kLdc + 4 * p4);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                acc[a][b] = fmaf(kv[a].x, nv[b].x, acc[a][b]);
                acc[a][b] = fmaf(kv[a].y, nv[b].y, acc[a][b]);
                acc[a][b] = fmaf(kv[a].z, nv[b].z, acc[a][b]);
                acc[a][b] = fmaf(kv[a].w, nv[b].w, acc[a][b]);
            }
    }
    float4* dst = reinterpret_cast<float4*>(tile + tid * 16);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float4 v = old[a];
        v.x += acc[a][0];
        v.y += acc[a][1];
        v.z += acc[a][2];
        v.w += acc[a][3];
        dst[a] = v;
    }
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// second-layer weights of a head: dWo[c][n] += sum_p dout[c][p] * hid[n][p], dbo[c] += sum_p dout[c][p]
template <int LD = kLdc>
__device__ __forceinline__ void coop_dwo(const float* __restrict__ hid, const float* __restrict__ dout, int C,
                                         float* __restrict__ dwo, float* __restrict__ dbo) {
    for (int idx = threadIdx.x; idx < C * kHid + C; idx += kBwdThreads) {
        float sum = 0.0f;
        if (idx < C * kHid) {
            int c = idx >> 6, n = idx & 63;
            for (int p4 = 0; p4 < 32; ++p4)
                sum += dot4(*reinterpret_cast<const float4*>(hid + n * LD + 4 * p4),
                            *reinterpret_cast<const float4*>(dout + c * LD + 4 * p4));
            dwo[idx] += sum;
        } else {
            int c = idx - C * kHid;
            for (int p4 = 0; p4 < 32; ++p4) {
                float4 v = *reinterpret_cast<const float4*>(dout + c * LD + 4 * p4);
                sum += (v.x + v.y) + (v.z + v.w);
            }
            dbo[c] += sum;
        }
    }
}

// first-layer bias and "extra input" weights of a head: db[n] += sum_p dpre[n][p]; dwx[n*ldx + e] += sum_p dpre[n][p] * ex[e][p]
__device__ __forceinline__ void coop_bias_extras(const float* __restrict__ dpre, const float* __restrict__ ex, int nEx,
                                                 float* __restrict__ db, float* __restrict__ dwx, int ldx) {
    const int n = threadIdx.x;
    if (n >= kHid) return;
    float sb = 0.0f, se[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int p4 = 0; p4 < 32; ++p4) {
        float4 v = *reinterpret_cast<const float4*>(dpre + n * kLdc + 4 * p4);
        sb += (v.x + v.y) + (v.z + v.w);
        for (int e = 0; e < nEx; ++e) se[e] += dot4(v, *reinterpret_cast<const float4*>(ex + e * kLdc + 4 * p4));
    }
    db[n] += sb;
    for (int e = 0; e < nEx; ++e) dwx[n * ldx + e] += se[e];
}

// dst[row(rl)][n] += sum over the pass's pairs of row rl of buf[n][col]
template <int LD = kLdc>
__device__ __forceinline__ void accumulate_rows(const float* __restrict__ buf, float* __restrict__ dst, int ld,
                                                const int* I, int L, int Wr, int pass_base, int npass) {
    for (int idx = threadIdx.x; idx < L * kHid; idx += kBwdThreads) {
        int rl = idx >> 6, n = idx & 63;
        int lo = max(rl * Wr, pass_base), hi = min((rl + 1) * Wr, pass_base + npass);
        if (hi <= lo) continue;
        float sum = 0.0f;
        for (int gp = lo; gp < hi; ++gp) sum += buf[n * LD + (gp - pass_base)];
        dst[I[IN_ROWS + rl] * ld + n] += sum;
    }
}

__device__ __forceinline__ void atomic_add_quat(float* p, const Quat& q) {
    atomicAdd(p + 0, q.w); atomicAdd(p + 1, q.x); atomicAdd(p + 2, q.y); atomicAdd(p + 3, q.z);
}

// Two threads per pair: threads p and p + 128 (warps w and w + 4) own column p of the pass.  Each computes the half of
// every 64-wide result whose index lies in its half h (hidden units n in [32h, 32h + 32) on the way forward, input features
// k in the same range on the way back), so the long GEMV chains are split without partial sums to exchange; the cheap
// per-pair geometry and the 64 x (1..7) second layers are computed by both.  The halves meet at a 64-thread named barrier.
__device__ __forceinline__ void pair_sync(int warp4) { asm volatile("bar.sync %0, 64;" ::"r"(1 + warp4) : "memory"); }

// acc[kk] += wrow[32 h + kk] * a, kk < 32
__device__ __forceinline__ void axpy32(float (&acc)[32], const float* __restrict__ wrow_half, float a) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow_half);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        float4 w = w4[k4];
        acc[4 * k4 + 0] = fmaf(w.x, a, acc[4 * k4 + 0]);
        acc[4 * k4 + 1] = fmaf(w.y, a, acc[4 * k4 + 1]);
        acc[4 * k4 + 2] = fmaf(w.z, a, acc[4 * k4 + 2]);
        acc[4 * k4 + 3] = fmaf(w.w, a, acc[4 * k4 + 3]);
    }
}

// m1[k] = relu(A_i[k] + A_j[k] + W_e[k]) for k in [k0, k0 + NK)
template <int LAYER, int NK>
__device__ __forceinline__ void compute_m1(float (&m1)[NK], int k0, const float* S, const BwdMap& M, const float* __restrict__ params,
                                           const float* __restrict__ ajt, int Kpad, int i, int j) {
    constexpr int H = layer_H(LAYER);
    constexpr int ld1 = 2 * H + kEdge;
    const float* ai = S + M.f.Ai + i * kLdN + k0;
    const bool pep = (j >= 0 && j < kN);
    const float* we = params + param_offset(LAYER, MSG0_W) + 2 * H + (pep ? (kN - 1 + i - j) : 0) + (size_t)k0 * ld1;
    // 16 features at a time: all their L2 loads (the A_j^T column, the relative-position column of peptide pairs) are in
    // flight before the first use
    const float* ajc = ajt + (j >= 0 ? j : 0) + (size_t)k0 * Kpad;
#pragma unroll
    for (int kb = 0; kb < NK; kb += 16) {
        float aj[16], wr[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) aj[k] = __ldcg(ajc + (kb + k) * Kpad);
#pragma unroll
        for (int k = 0; k < 16; ++k) wr[k] = pep ? __ldg(we + (kb + k) * ld1) : 0.0f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float v = ai[kb + k];
            if (j >= 0) v += aj[k];
            if (pep) v += wr[k];
            m1[kb + k] = fmaxf(v, 0.0f);
        }
    }
}

// One pass of up to 128 pairs.  HEADS = attention-carrying pairs (all four heads); otherwise message-only
// pairs of layer 1 whose only gradient source is the unmasked message sum (model.py:151).
template <int LAYER, bool HEADS>
__device__ __forceinline__ void pair_pass(float* S, const BwdMap& M, const BwdArgs& g, const PairRef pr, float mult,
                                          int b, const float* __restrict__ ajt, float* __restrict__ dajt,
                                          float* __restrict__ tiles, float* __restrict__ direct, const int* I,
                                          int L, int Wr, int pass_base, int npass, int n_pocket_cols, int pocket_e0) {
    constexpr bool IN_GRADS = (LAYER == 1);
    const LayerArgs& a = g.a;
    const int tid = threadIdx.x;
    const int p = tid & (kBwdPairs - 1);          // this thread's pair column
    const int half = tid >> 7, n0 = 32 * half;    // its half of every 64-wide result
    const int warp4 = (tid >> 5) & 3;
    const bool owner = half == 0;                 // per-pair side effects (atomics, the small dout / extras columns) happen once
    const int i = pr.i, j = pr.j;
    const bool act = pr.active;
    const bool pep = (j >= 0 && j < kN);
    float* bufA = S + M.BufA;
    float* bufB = S + M.BufB;
    float* sDout = S + M.Dout;
    float* sEx = S + M.Ex;
    const int Kpad = a.Kpad;
    // offsets of this layer's tensors inside `direct` (relative to the layer's first parameter)
    constexpr int base = param_offset(LAYER, 0);
    float dm[32];                                  // dL / d message[n0 + kk]

    if (HEADS) {
        {   // ---- recompute the message: m1 (registers) -> m (BufA column), my half of the outputs ----
            float m1[kHid];
            compute_m1<LAYER, kHid>(m1, 0, S, M, a.params, ajt, Kpad, i, j);
#pragma unroll 2
            for (int nn = 0; nn < 32; ++nn) {
                const int n = n0 + nn;
                bufA[n * kLdc + p] = S[M.f.PkMisc + 4 * n + 3] + dot64(S + M.W2 + n * kHid, m1);
            }
        }
        pair_sync(warp4);
        float m[kHid];
#pragma unroll
        for (int k = 0; k < kHid; ++k) m[k] = bufA[k * kLdc + p];
#pragma unroll
        for (int k = 0; k < 32; ++k) dm[k] = 0.0f;

        const float* rg = S + M.RowG + i * 16;
        const float lse = rg[15], c_i = rg[14];
        const float logit = g.logits[((size_t)b * kN + i) * Kpad + j];
        const float w = act ? expf(logit - lse) : 0.0f;
        const float* pqi = S + M.f.Q + i * 4;
        const float* pqj = S + M.f.Q + j * 4;
        const Quat qi{pqi[0], pqi[1], pqi[2], pqi[3]}, qj{pqj[0], pqj[1], pqj[2], pqj[3]};
        const float rx = S[M.f.X + i * 3] - S[M.f.X + j * 3], ry = S[M.f.X + i * 3 + 1] - S[M.f.X + j * 3 + 1],
                    rz = S[M.f.X + i * 3 + 2] - S[M.f.X + j * 3 + 2];
        float dLdw = 0.0f;
        // layer 2: this pair's share of dL / d (q_i, x_i); every pair of a row adds to the same node, so the shares go through
        // one column per pair and a row reduction below instead of shared-memory atomics that serialise a whole row
        float gi[7] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};

        // ================= rotation head (model.py:283-296) =================
        {
            const Quat qinvj = qinv(qj);
            const Quat v = qmul(qi, qj);
            const Quat lq = qmul(qinvj, v);
#pragma unroll 2
            for (int nn = 0; nn < 32; ++nn) {
                const int n = n0 + nn;
                const float4 wq = *reinterpret_cast<const float4*>(S + M.f.PkRotQ + 4 * n);
                float s = S[M.f.PkMisc + 4 * n + 2] + wq.x * lq.w + wq.y * lq.x + wq.z * lq.y + wq.w * lq.z;
                s += dot64(S + M.Wh + HD_ROT * 4096 + n * kHid, m);
                bufB[n * kLdc + p] = fmaxf(s, 0.0f);
            }
            pair_sync(warp4);
            float pre[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) pre[c] = S[M.f.Scal + SC_ROT2B + c];
            unsigned long long on = 0ull;
#pragma unroll 4
            for (int n = 0; n < kHid; ++n) {
                const float h = bufB[n * kLdc + p];
                on |= (unsigned long long)(h > 0.0f) << n;
                const float4 w2 = *reinterpret_cast<const float4*>(S + M.f.PkRot2 + 4 * n);
                pre[0] = fmaf(w2.x, h, pre[0]); pre[1] = fmaf(w2.y, h, pre[1]);
                pre[2] = fmaf(w2.z, h, pre[2]); pre[3] = fmaf(w2.w, h, pre[3]);
            }
            const Quat dl{sigmoidf(pre[0]), sigmoidf(pre[1]), sigmoidf(pre[2]), sigmoidf(pre[3])};
            const Quat u = qmul(dl, qinvj);
            const Quat dg = qmul(qj, u);
            const Quat dG{rg[0], rg[1], rg[2], rg[3]};
            dLdw += qdot(dG, dg);
            const Quat ddg = qscale(dG, w);
            const Quat du = qmul_grad_b(qj, ddg);         // dg = qj * u
            const Quat ddl = qmul_grad_a(du, qinvj);      // u = dl * qinvj
            float dp2[4] = {ddl.w * dl.w * (1.0f - dl.w), ddl.x * dl.x * (1.0f - dl.x), ddl.y * dl.y * (1.0f - dl.y),
                            ddl.z * dl.z * (1.0f - dl.z)};
            if (owner) {
#pragma unroll
                for (int c = 0; c < 4; ++c) sDout[c * kLdc + p] = dp2[c];
                sEx[0 * kLdc + p] = lq.w; sEx[1 * kLdc + p] = lq.x; sEx[2 * kLdc + p] = lq.y; sEx[3 * kLdc + p] = lq.z;
            }
            __syncthreads();
            coop_dwo(bufB, sDout, 4, direct + (param_offset(LAYER, ROT2_W) - base), direct + (param_offset(LAYER, ROT2_B) - base));
            __syncthreads();
            float dlq[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 2
            for (int n = 0; n < kHid; ++n) {
                const float4 w2 = *reinterpret_cast<const float4*>(S + M.f.PkRot2 + 4 * n);
                float dp = w2.x * dp2[0] + w2.y * dp2[1] + w2.z * dp2[2] + w2.w * dp2[3];
                dp = ((on >> n) & 1ull) ? dp : 0.0f;
                if ((n >> 5) == half) bufB[n * kLdc + p] = dp;
                axpy32(dm, S + M.Wh + HD_ROT * 4096 + n * kHid + n0, dp);
                const float4 wq = *reinterpret_cast<const float4*>(S + M.f.PkRotQ + 4 * n);
                dlq[0] = fmaf(wq.x, dp, dlq[0]); dlq[1] = fmaf(wq.y, dp, dlq[1]);
                dlq[2] = fmaf(wq.z, dp, dlq[2]); dlq[3] = fmaf(wq.w, dp, dlq[3]);
            }
            if (IN_GRADS && act && owner) {
                const Quat dlqq{dlq[0], dlq[1], dlq[2], dlq[3]};
                Quat dqinv = qmul_grad_a(dlqq, v);             // lq = qinvj * v
                const Quat dv = qmul_grad_b(qinvj, dlqq);
                Quat dqi = qmul_grad_a(dv, qj);                // v = qi * qj
                Quat dqj = qmul_grad_b(qi, dv);
                dqj = qadd(dqj, qmul_grad_a(ddg, u));          // dg = qj * u
                dqinv = qadd(dqinv, qmul_grad_b(dl, du));      // u = dl * qinvj
                dqj = qadd(dqj, qinv_grad(qj, dqinv));
                gi[0] += dqi.w; gi[1] += dqi.x; gi[2] += dqi.y; gi[3] += dqi.z;
                if (pep) atomic_add_quat(S + M.dQ + j * 4, dqj);
            }
            __syncthreads();
            coop_outer(bufB, bufA, tiles + T_ROT * 4096);
            coop_bias_extras(bufB, sEx, 4, direct + (param_offset(LAYER, ROT0_B) - base),
                             direct + (param_offset(LAYER, ROT0_W) - base) + 64, 68);
            __syncthreads();
        }

        // ================= torsion head (model.py:257-263) =================
        {
#pragma unroll 2
            for (int nn = 0; nn < 32; ++nn) {
                const int n = n0 + nn;
                const float s = S[M.f.Tt + i * kHid + n] + dot64(S + M.Wh + HD_TOR * 4096 + n * kHid, m);
                bufB[n * kLdc + p] = fmaxf(s, 0.0f);
            }
            pair_sync(warp4);
            float da[PMHC_NTORS];
#pragma unroll
            for (int c = 0; c < PMHC_NTORS; ++c) da[c] = S[M.f.Scal + SC_TOR2B + c];
            unsigned long long on = 0ull;
#pragma unroll 4
            for (int n = 0; n < kHid; ++n) {
                const float h = bufB[n * kLdc + p];
                on |= (unsigned long long)(h > 0.0f) << n;
                const float4 w0 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n);
                const float4 w1 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n + 4);
                da[0] = fmaf(w0.x, h, da[0]); da[1] = fmaf(w0.y, h, da[1]); da[2] = fmaf(w0.z, h, da[2]);
                da[3] = fmaf(w0.w, h, da[3]); da[4] = fmaf(w1.x, h, da[4]); da[5] = fmaf(w1.y, h, da[5]);
                da[6] = fmaf(w1.z, h, da[6]);
            }
            float dda[PMHC_NTORS];
#pragma unroll
            for (int c = 0; c < PMHC_NTORS; ++c) {
                dLdw = fmaf(rg[4 + c], da[c], dLdw);
                dda[c] = w * rg[4 + c];
                if (owner) sDout[c * kLdc + p] = dda[c];
            }
            __syncthreads();
            coop_dwo(bufB, sDout, PMHC_NTORS, direct + (param_offset(LAYER, TOR2_W) - base), direct + (param_offset(LAYER, TOR2_B) - base));
            __syncthreads();
#pragma unroll 2
            for (int n = 0; n < kHid; ++n) {
                const float4 w0 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n);
                const float4 w1 = *reinterpret_cast<const float4*>(S + M.f.PkTor2 + 8 * n + 4);
                float dp = w0.x * dda[0] + w0.y * dda[1] + w0.z * dda[2] + w0.w * dda[3] + w1.x * dda[4] + w1.y * dda[5] + w1.z * dda[6];
                dp = ((on >> n) & 1ull) ? dp : 0.0f;
                if ((n >> 5) == half) bufB[n * kLdc + p] = dp;
                axpy32(dm, S + M.Wh + HD_TOR * 4096 + n * kHid + n0, dp);
            }
            __syncthreads();
            coop_outer(bufB, bufA, tiles + T_TOR * 4096);
            accumulate_rows(bufB, S + M.dTt, kHid, I, L, Wr, pass_base, npass);
            __syncthreads();
        }

        // ================= translation head (model.py:325-331) =================
        {
#pragma unroll 2
            for (int nn = 0; nn < 32; ++nn) {
                const int n = n0 + nn;
                const float s = S[M.f.PkMisc + 4 * n + 0] + dot64(S + M.Wh + HD_TRN * 4096 + n * kHid, m);
                bufB[n * kLdc + p] = fmaxf(s, 0.0f);
            }
            pair_sync(warp4);
            float sc = S[M.f.Scal + SC_TRN2B];
            unsigned long long on = 0ull;
#pragma unroll 4
            for (int n = 0; n < kHid; ++n) {
                const float h = bufB[n * kLdc + p];
                on |= (unsigned long long)(h > 0.0f) << n;
                sc = fmaf(S[M.f.PkMisc + 4 * n + 1], h, sc);
            }
            const float dXr = rg[11] * rx + rg[12] * ry + rg[13] * rz;
            dLdw = fmaf(sc, dXr, dLdw);
            const float ds = w * dXr;
            if (owner) sDout[p] = ds;
            if (IN_GRADS && act && owner) {
                const float f = w * sc;
                gi[4] += f * rg[11]; gi[5] += f * rg[12]; gi[6] += f * rg[13];
                if (pep) {
                    atomicAdd(S + M.dX + j * 3 + 0, -f * rg[11]); atomicAdd(S + M.dX + j * 3 + 1, -f * rg[12]); atomicAdd(S + M.dX + j * 3 + 2, -f * rg[13]);
                }
            }
            __syncthreads();
            coop_dwo(bufB, sDout, 1, direct + (param_offset(LAYER, TRN2_W) - base), direct + (param_offset(LAYER, TRN2_B) - base));
            __syncthreads();
#pragma unroll 2
            for (int n = 0; n < kHid; ++n) {
                const float dp = ((on >> n) & 1ull) ? S[M.f.PkMisc + 4 * n + 1] * ds : 0.0f;
                if ((n >> 5) == half) bufB[n * kLdc + p] = dp;
                axpy32(dm, S + M.Wh + HD_TRN * 4096 + n * kHid + n0, dp);
            }
            __syncthreads();
            coop_outer(bufB, bufA, tiles + T_TRN * 4096);
            coop_bias_extras(bufB, sEx, 0, direct + (param_offset(LAYER, TRN0_B) - base), nullptr, 0);
            __syncthreads();
        }

        // ================= attention head (model.py:238-243) =================
        {
            const float d2 = rx * rx + ry * ry + rz * rz;
            const float dotq = qdot(qi, qj);
            const float qd = dotq * dotq;
#pragma unroll 2
            for (int nn = 0; nn < 32; ++nn) {
                const int n = n0 + nn;
                const float4 pk = *reinterpret_cast<const float4*>(S + M.f.PkAtt + 4 * n);
                const float s = (pk.z + dot64(S + M.Wh + HD_ATT * 4096 + n * kHid, m)) + fmaf(pk.y, qd, pk.x * -d2);  // same order as the forward
                bufB[n * kLdc + p] = fmaxf(s, 0.0f);
            }
            pair_sync(warp4);
            unsigned long long on = 0ull;
#pragma unroll 4
            for (int n = 0; n < kHid; ++n) on |= (unsigned long long)(bufB[n * kLdc + p] > 0.0f) << n;
            // softmax backward with the saved row statistics.  A fully saturated row (w == 1 exactly, every other
            // weight underflowed) has g - sum_k w_k g_k == 0 exactly in the reference's autograd; here c_i comes from
            // the forward's aggregates, so the same difference would be rounding noise (~1e-7 |g|) that the d2 input
            // of attention_mlp.0 (thousands of A^2) then amplifies — define it as the exact zero it is.
            const float dlogit = (w == 1.0f) ? 0.0f : w * (dLdw - c_i);
            if (owner) {
                sDout[p] = dlogit;
                sEx[0 * kLdc + p] = -d2;
                sEx[1 * kLdc + p] = qd;
            }
            __syncthreads();
            coop_dwo(bufB, sDout, 1, direct + (param_offset(LAYER, ATT2_W) - base), direct + (param_offset(LAYER, ATT2_B) - base));
            __syncthreads();
            float gd = 0.0f, gq = 0.0f;
#pragma unroll 2
            for (int n = 0; n < kHid; ++n) {
                const float4 pk = *reinterpret_cast<const float4*>(S + M.f.PkAtt + 4 * n);
                const float dp = ((on >> n) & 1ull) ? pk.w * dlogit : 0.0f;
                if ((n >> 5) == half) bufB[n * kLdc + p] = dp;
                axpy32(dm, S + M.Wh + HD_ATT * 4096 + n * kHid + n0, dp);
                gd = fmaf(pk.x, dp, gd);
                gq = fmaf(pk.y, dp, gq);
            }
            if (IN_GRADS && act && owner) {
                const float f = -gd * 2.0f;                 // d(-d2) = gd
                gi[4] += f * rx; gi[5] += f * ry; gi[6] += f * rz;
                const float fq = gq * 2.0f * dotq;
                gi[0] += fq * qj.w; gi[1] += fq * qj.x; gi[2] += fq * qj.y; gi[3] += fq * qj.z;
                if (pep) {
                    atomicAdd(S + M.dX + j * 3 + 0, -f * rx); atomicAdd(S + M.dX + j * 3 + 1, -f * ry); atomicAdd(S + M.dX + j * 3 + 2, -f * rz);
                    atomic_add_quat(S + M.dQ + j * 4, qscale(qi, fq));
                }
            }
            __syncthreads();
            coop_outer(bufB, bufA, tiles + T_ATT * 4096);
            coop_bias_extras(bufB, sEx, 2, direct + (param_offset(LAYER, ATT0_B) - base),
                             direct + (param_offset(LAYER, ATT0_W) - base) + 64, 66);
            __syncthreads();
        }
        if (LAYER == 0) {
#pragma unroll
            for (int k = 0; k < 32; ++k) dm[k] += S[M.dMsum + i * kHid + n0 + k];
        }
        if (IN_GRADS && owner) {
#pragma unroll
            for (int c = 0; c < 7; ++c) sDout[c * kLdc + p] = act ? gi[c] : 0.0f;     // (Dout is free behind the last head's barrier)
        }
    } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) dm[k] = mult * S[M.dMsum + i * kHid + n0 + k];
    }

    // ---- message MLP backward: dW2 += dm (x) m1, dm1 = relu'(.) W2^T dm, then the per-node reductions ----
    float dm1[32];
    {
        float m1h[32];
        compute_m1<LAYER, 32>(m1h, n0, S, M, a.params, ajt, Kpad, i, j);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            bufB[(n0 + k) * kLdc + p] = act ? dm[k] : 0.0f;
            bufA[(n0 + k) * kLdc + p] = m1h[k];
        }
        pair_sync(warp4);
#pragma unroll
        for (int k = 0; k < 32; ++k) dm1[k] = 0.0f;
#pragma unroll 2
        for (int n = 0; n < kHid; ++n) axpy32(dm1, S + M.W2 + n * kHid + n0, bufB[n * kLdc + p]);
#pragma unroll
        for (int k = 0; k < 32; ++k) dm1[k] = m1h[k] > 0.0f ? dm1[k] : 0.0f;
    }
    __syncthreads();
    if (HEADS && IN_GRADS) {
        for (int idx = tid; idx < L * 7; idx += kBwdThreads) {
            const int rl = idx / 7, c = idx - rl * 7;
            const int lo = max(rl * Wr, pass_base), hi = min((rl + 1) * Wr, pass_base + npass);
            float sum = 0.0f;
            for (int gp = lo; gp < hi; ++gp) sum += sDout[c * kLdc + (gp - pass_base)];
            const int ri = I[IN_ROWS + rl];
            if (hi > lo) {
                if (c < 4) S[M.dQ + ri * 4 + c] += sum;
                else S[M.dX + ri * 3 + (c - 4)] += sum;
            }
        }
    }
    coop_outer(bufB, bufA, tiles + T_W2 * 4096);
    coop_bias_extras(bufB, sEx, 0, direct + (param_offset(LAYER, MSG2_B) - base), nullptr, 0);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) bufB[(n0 + k) * kLdc + p] = dm1[k];
    if (act && pep) {
        float* dj = S + M.dAjPep + j * kLdN + n0;
        float* de = S + M.dWe + (kN - 1 + i - j) * kLdN + n0;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            atomicAdd(dj + k, dm1[k]);
            atomicAdd(de + k, dm1[k]);
        }
    } else if (act && !HEADS && j >= kN) {
        // masked pocket slot with non-zero features (rare): straight to the A_j^T gradient scratch
        for (int k = 0; k < 32; ++k) atomicAdd(dajt + (n0 + k) * Kpad + j, dm1[k]);
    }
    __syncthreads();
    accumulate_rows(bufB, S + M.dAi, kLdN, I, L, Wr, pass_base, npass);
    if (HEADS) {
        // dA_j^T[k][j] += sum over rows of dm1 for the valid pocket columns of this pass
        const int rl_lo = pass_base / Wr, rl_hi = (pass_base + npass - 1) / Wr;
        const int n_items = n_pocket_cols * kHid;
        for (int idx0 = tid; idx0 < n_items; idx0 += 4 * kBwdThreads) {   // four L2 read-modify-writes in flight per thread
            float old[4];
            int addr[4], kk[4], ee[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = idx0 + u * kBwdThreads;
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
                    if (col >= 0 && col < npass) sum += bufB[kk[u] * kLdc + col];
                }
                dajt[addr[u]] = old[u] + sum;
            }
        }
    }
    __syncthreads();
}

// =====================================================================================================================
// Tensor-core mode (PMHC_PRECISION_BF16 training): every 64-wide contraction of the pass — the forward recomputation
// (message layer 2, the four head hidden layers), the input gradients (W_h^T dpre, W2^T dm) and the weight-gradient outer
// products — runs as a CTA-level TF32 GEMM on the warp-level tensor-core path (mma.sync.m16n8k8, fp32 accumulate) over
// the same [feature][pair] column tiles; the per-pair second layers, geometry and softmax backward stay fp32 scalar code.
// Operands are rounded to tf32 (cvt.rna) once where they are written; gate: the 1e-2 class of the tensor-core forward.
// =====================================================================================================================
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// acc[mt][nt] += A (64 x 64) * X (64 x this warp's 16 pair columns).  A[m][k] = W[m][k] (TRANS = false) or W[k][m] (true),
// W swizzled (wswz); X a [64][kLdt] column tile.  Fragment (mt, nt) element e: row 16 mt + g + 8 (e >> 1),
// pair column 16 warp + 8 nt + 2 t + (e & 1), g = lane >> 2, t = lane & 3.
template <bool TRANS>
__device__ __forceinline__ void gemm64_tf32(const float* __restrict__ W, const float* __restrict__ X, float (&acc)[4][2][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const uint32_t* Wu = reinterpret_cast<const uint32_t*>(W);
    const uint32_t* xb = reinterpret_cast<const uint32_t*>(X) + 16 * warp + g + t * kLdt;
    // non-transposed: (a0, a2) of row n = 16 mt + g + 8 h are the adjacent stored columns 8 ks + 2 t, + 1 (kperm):
    //                 index = n * 64 + (((2 t) ^ sw) ^ 8 ks), sw = ((g & 3) << 3) | (g & 4)
    // transposed:     A[m][k] = W[k][m], stored at k * 64 + (kperm(m) ^ swz(k)) with k = 8 ks + t + 4 c, m = 16 mt + g + 8 h:
    //                 index = (8 ks + t + 4 c) * 64 + ((pg | t << 3) ^ ((16 mt + 8 h) ^ 4 c)), pg = 2 (g & 3) + (g >> 2)
    const int base = TRANS ? t * 64 : g * 64;
    const int x = TRANS ? ((2 * (g & 3) + (g >> 2)) | (t << 3)) : ((2 * t) ^ (((g & 3) << 3) | (g & 4)));
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        uint32_t b[2][2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            b[nt][0] = xb[(8 * ks) * kLdt + 8 * nt];
            b[nt][1] = xb[(8 * ks + 4) * kLdt + 8 * nt];
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            uint32_t a[4];
            if (TRANS) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int h = e & 1, c = e >> 1;   // a0: (g, t) a1: (g + 8, t) a2: (g, t + 4) a3: (g + 8, t + 4)
                    a[e] = Wu[base + (8 * ks + 4 * c) * 64 + (x ^ ((16 * mt + 8 * h) ^ (4 * c)))];
                }
            } else {
                const uint2 lo = *reinterpret_cast<const uint2*>(Wu + base + (16 * mt) * 64 + (x ^ (8 * ks)));
                const uint2 hi = *reinterpret_cast<const uint2*>(Wu + base + (16 * mt + 8) * 64 + (x ^ (8 * ks)));
                a[0] = lo.x; a[2] = lo.y; a[1] = hi.x; a[3] = hi.y;
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) mma_tf32(acc[mt][nt], a[0], a[1], a[2], a[3], b[nt][0], b[nt][1]);
        }
    }
}

// Weight-gradient outer products of one pass: tile[n][k] += sum_p Bn[n][p] Bk[k][p] over the 128 pair columns, plus the
// same rows against the 8-row geometry tile (lq 0..3, -d2, qdot2, 1, 0): columns [e0, e0 + ne) of that product go to the
// "extra input" weights xw[n * ldw + e - e0], column 6 (the ones row) to the first-layer bias xb[n].
// Warp w owns rows n in [16 (w >> 1), +16) and columns k in [32 (w & 1), +32); its 16 sums per thread sit contiguously in
// the tile-owner layout (tile + 16 tid); the even warps also own the extras of their rows.
__device__ __forceinline__ void outer_tf32(const float* __restrict__ Bn, const float* __restrict__ Bk, const float* __restrict__ Geo,
                                           float* __restrict__ tile, const float4 (&old)[4], float* __restrict__ xw, int ldw,
                                           int e0, int ne, float* __restrict__ xb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int mt = warp >> 1, kh = warp & 1;
    // the K index (pair column) of a k-step is permuted — MMA slots (t, t + 4) hold pair columns (2 t, 2 t + 1) in both operands —
    // so every fragment pair is one 64-bit shared load
    const uint2* an = reinterpret_cast<const uint2*>(Bn + (16 * mt + g) * kLdt) + t;
    const uint2* bk = reinterpret_cast<const uint2*>(Bk + (32 * kh + g) * kLdt) + t;
    const uint2* ge = reinterpret_cast<const uint2*>(Geo + g * kLdt) + t;
    float acc[4][4], accx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float4* dst = reinterpret_cast<float4*>(tile + tid * 16);
    // the running sums come from L2: `old` was loaded by the caller ahead of the barrier before this call (outer_prefetch)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.0f;
#pragma unroll 4
    for (int ks = 0; ks < 16; ++ks) {
        const uint2 a02 = an[4 * ks], a13 = an[4 * kLdt + 4 * ks];     // rows g and g + 8 (8 rows = 4 kLdt uint2)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const uint2 bb = bk[(4 * nt) * kLdt + 4 * ks];               // row + 8 nt
            mma_tf32(acc[nt], a02.x, a13.x, a02.y, a13.y, bb.x, bb.y);
        }
        if (kh == 0) {
            const uint2 gg = ge[4 * ks];
            mma_tf32(accx, a02.x, a13.x, a02.y, a13.y, gg.x, gg.y);
        }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) dst[nt] = make_float4(old[nt].x + acc[nt][0], old[nt].y + acc[nt][1], old[nt].z + acc[nt][2], old[nt].w + acc[nt][3]);
    if (kh == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int n = 16 * mt + g + 8 * (e >> 1), col = 2 * t + (e & 1);
            if (xw != nullptr && col >= e0 && col < e0 + ne) xw[n * ldw + (col - e0)] += accx[e];
            if (xb != nullptr && col == 6) xb[n] += accx[e];
        }
    }
}
__device__ __forceinline__ void outer_prefetch(const float* __restrict__ tile, float4 (&old)[4]) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) old[nt] = ldcg_early4(tile + threadIdx.x * 16 + 4 * nt);
}
// (n, k) of element e of thread tid in the tensor-core tile-owner layout
__device__ __forceinline__ void tc_tile_coord(int tid, int e, int& n, int& k) {
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int nt = e >> 2, r = e & 3;
    n = 16 * (warp >> 1) + g + 8 * (r >> 1);
    k = 32 * (warp & 1) + 8 * nt + 2 * t + (r & 1);
}

__device__ __forceinline__ void zero_acc(float (&acc)[4][2][4]) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.0f;
}

enum { GEO_LQ = 0, GEO_D2 = 4, GEO_QD = 5, GEO_ONE = 6, GEO_ZERO = 7, GEO_D2F = 8, GEO_QDF = 9 };

// out rows: acc2[nt] = sum_n Wimg[c][n] X[n][.] for the C (<= 8) rows c = g of a swizzled row image, over this warp's 16 pair
// columns: element e < 2 of acc2[nt] is (row g, column 16 warp + 8 nt + 2 t + e); the m16 tile's rows g + 8 are zero.
__device__ __forceinline__ void rows8_tf32(const float* __restrict__ Wimg, int C, const float* __restrict__ X, float (&acc2)[2][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const uint32_t* wu = reinterpret_cast<const uint32_t*>(Wimg) + g * 64;
    const uint32_t* xb = reinterpret_cast<const uint32_t*>(X) + 16 * warp + g + t * kLdt;
    const bool rowok = g < C;
    const int x = t | ((g & 3) << 3) | (g & 4);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc2[nt][e] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const uint32_t a0 = rowok ? wu[x ^ (8 * ks)] : 0u, a2 = rowok ? wu[x ^ (8 * ks + 4)] : 0u;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) mma_tf32(acc2[nt], a0, 0u, a2, 0u, xb[(8 * ks) * kLdt + 8 * nt], xb[(8 * ks + 4) * kLdt + 8 * nt]);
    }
}
// acc[mt][nt] = sum_{c < C} Wimg[c][16 mt + .] D[c][.]: the hidden-layer gradient before the relu mask (one k-step, c = t, t + 4)
__device__ __forceinline__ void cols8_tf32(const float* __restrict__ Wimg, int C, const float* __restrict__ D, float (&acc)[4][2][4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const uint32_t* wu = reinterpret_cast<const uint32_t*>(Wimg);
    const uint32_t* du = reinterpret_cast<const uint32_t*>(D) + 16 * warp + g;
    const bool lo = t < C, hi = t + 4 < C;
    uint32_t b[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        b[nt][0] = lo ? du[t * kLdt + 8 * nt] : 0u;
        b[nt][1] = hi ? du[(t + 4) * kLdt + 8 * nt] : 0u;
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        // row c = t: swizzle ((t & 3) << 3) | (t & 4) = t << 3; row t + 4: (t << 3) | 4
        const uint32_t a0 = lo ? wu[t * 64 + ((16 * mt + g) ^ (t << 3))] : 0u;
        const uint32_t a1 = lo ? wu[t * 64 + ((16 * mt + g + 8) ^ (t << 3))] : 0u;
        const uint32_t a2 = hi ? wu[(t + 4) * 64 + ((16 * mt + g) ^ ((t << 3) | 4))] : 0u;
        const uint32_t a3 = hi ? wu[(t + 4) * 64 + ((16 * mt + g + 8) ^ ((t << 3) | 4))] : 0u;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.0f;
            mma_tf32(acc[mt][nt], a0, a1, a2, a3, b[nt][0], b[nt][1]);
        }
    }
}
// second-layer weight gradients: dwo[c * 64 + n] += sum_p D[c][p] Hid[n][p], dbo[c] += sum_p D[c][p] over the 128 pair columns;
// warp w owns the hidden units n in [8 w, 8 w + 8), warp 0 also the bias (a B operand of ones)
__device__ __forceinline__ void dwo_tf32(const float* __restrict__ D, int C, const float* __restrict__ Hid,
                                         float* __restrict__ dwo, float* __restrict__ dbo, float old0, float old1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const bool rowok = g < C;
    const uint32_t* du = reinterpret_cast<const uint32_t*>(D) + g * kLdt + t;
    const uint32_t* hu = reinterpret_cast<const uint32_t*>(Hid) + (8 * warp + g) * kLdt + t;
    float* dst = dwo + g * 64 + 8 * warp + 2 * t;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f}, accb[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const uint32_t one = __float_as_uint(1.0f);
#pragma unroll 4
    for (int ks = 0; ks < 16; ++ks) {
        const uint32_t a0 = rowok ? du[8 * ks] : 0u, a2 = rowok ? du[8 * ks + 4] : 0u;
        mma_tf32(acc, a0, 0u, a2, 0u, hu[8 * ks], hu[8 * ks + 4]);
        if (warp == 0) mma_tf32(accb, a0, 0u, a2, 0u, one, one);
    }
    if (rowok) { dst[0] = old0 + acc[0]; dst[1] = old1 + acc[1]; }
    if (warp == 0 && rowok && t == 0) dbo[g] += accb[0];
}

// One pass of up to 128 pairs, tensor-core mode.  Same contract as pair_pass.  Per head: hidden layer (GEMM) -> second
// layer (MMA, own columns) | barrier | per-pair geometry and output gradients (threads 0..127, one per pair) | barrier |
// second-layer weight gradients (MMA) | barrier | hidden-layer gradient (MMA + relu mask, in place), input gradient GEMM
// into the running dL/dm fragments, extras gradients | barrier | first-layer weight-gradient outer products | barrier.
template <int LAYER, bool HEADS>
__device__ __forceinline__ void pair_pass_tc(float* S, const BwdMap& M, const BwdArgs& g, const PairRef pr, float mult,
                                             int b, const float* __restrict__ ajt, float* __restrict__ dajt,
                                             float* __restrict__ tiles, float* __restrict__ direct, const int* I,
                                             int L, int Wr, int pass_base, int npass, int n_pocket_cols, int pocket_e0) {
    constexpr bool IN_GRADS = (LAYER == 1);
    const LayerArgs& a = g.a;
    const int tid = threadIdx.x;
    const int p = tid & (kBwdPairs - 1);          // this thread's pair column in the per-pair stages
    const int half = tid >> 7, n0 = 32 * half;
    const bool owner = half == 0;
    const int lane = tid & 31, warp = tid >> 5, fg = lane >> 2, ft = lane & 3;   // fragment coordinates in the GEMM stages
    const int i = pr.i, j = pr.j;
    const bool act = pr.active;
    const bool pep = (j >= 0 && j < kN);
    float* bufA = S + M.BufA;
    float* bufB = S + M.BufB;
    float* sDout = S + M.Dout;
    float* sGeo = S + M.Ex;
    float* sDx = S + M.Dx;
    int* sPI = reinterpret_cast<int*>(sDout + 7 * kLdt);   // row node of every pair column of the pass
    const float* W2s = S + M.W2;
    const float* Whs = S + M.Wh;
    const float* W3s = S + M.W3i;
    const float* Wxs = S + M.Wx;
    const int Kpad = a.Kpad;
    constexpr int base = param_offset(LAYER, 0);
    float acc[4][2][4], o2[2][4];

    // store the valid rows (g < C) of a rows8 result into a [.][kLdt] tile at this warp's columns
    auto store_rows = [&](float* dst, int C) {
        if (fg < C) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
                *reinterpret_cast<float2*>(dst + fg * kLdt + 16 * warp + 8 * nt + 2 * ft) = make_float2(o2[nt][0], o2[nt][1]);
        }
    };
    // hidden-layer gradient: acc (before the mask) -> relu mask from the hidden activations in BufB -> BufB in place (own columns)
    auto mask_store_dpre = [&]() {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = 16 * mt + fg + 8 * h;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    float2* q = reinterpret_cast<float2*>(bufB + n * kLdt + 16 * warp + 8 * nt + 2 * ft);
                    const float2 hv = *q;
                    *q = make_float2(hv.x > 0.0f ? tf32r(acc[mt][nt][2 * h]) : 0.0f, hv.y > 0.0f ? tf32r(acc[mt][nt][2 * h + 1]) : 0.0f);
                }
            }
    };

    int* sPl = reinterpret_cast<int*>(S + M.Pl);           // [0]: count, [4 + slot]: column | j << 8 | relative position << 16
    if (owner) {
        sPI[p] = i;
        sGeo[GEO_ONE * kLdt + p] = 1.0f;
        sGeo[GEO_ZERO * kLdt + p] = 0.0f;
    }
    if (tid == 0) sPl[0] = 0;      // (every reader of the previous pass's list is behind that pass's last barrier)
    // called behind the pass's first barrier; the list is read behind a later one
    auto list_peptide_columns = [&]() {
        if (owner && act && pep) sPl[4 + atomicAdd(sPl, 1)] = p | (j << 8) | ((kN - 1 + i - j) << 16);
    };
    if (HEADS) {
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
        float dLdw = 0.0f;
        float gi[7] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};   // layer 2: this pair's share of dL / d (q_i, x_i), summed per row below
        Quat dqj_part{0.0f, 0.0f, 0.0f, 0.0f}, dqinv_part{0.0f, 0.0f, 0.0f, 0.0f};   // rotation head: the input-gradient terms that do not wait for dlq
        {   // ---- m1 (my half of its features) -> BufB; the pair's geometry row -> sGeo ----
            float m1h[32];
            compute_m1<LAYER, 32>(m1h, n0, S, M, a.params, ajt, Kpad, i, j);
#pragma unroll
            for (int k = 0; k < 32; ++k) bufB[(n0 + k) * kLdt + p] = tf32r(m1h[k]);
            if (owner) {
                sGeo[(GEO_LQ + 0) * kLdt + p] = tf32r(lq.w); sGeo[(GEO_LQ + 1) * kLdt + p] = tf32r(lq.x);
                sGeo[(GEO_LQ + 2) * kLdt + p] = tf32r(lq.y); sGeo[(GEO_LQ + 3) * kLdt + p] = tf32r(lq.z);
                sGeo[GEO_D2 * kLdt + p] = tf32r(-d2); sGeo[GEO_QD * kLdt + p] = tf32r(qd);
                sGeo[GEO_D2F * kLdt + p] = -d2;       sGeo[GEO_QDF * kLdt + p] = qd;
            }
        }
        __syncthreads();
        list_peptide_columns();
        // ---- message = W2 m1 + b2 -> BufA ----
        zero_acc(acc);
        gemm64_tf32<false>(W2s, bufB, acc);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = 16 * mt + fg + 8 * h;
                const float b2 = S[M.f.PkMisc + 4 * n + 3];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int pc = 16 * warp + 8 * nt + 2 * ft;
                    *reinterpret_cast<float2*>(bufA + n * kLdt + pc) =
                        make_float2(tf32r(acc[mt][nt][2 * h] + b2), tf32r(acc[mt][nt][2 * h + 1] + b2));
                }
            }
        __syncthreads();
        float dm[4][2][4];                         // dL / d message, fragment layout (row = feature, column = pair)
        zero_acc(dm);

        // ================= the four heads: one runtime loop (rotation, torsion, translation, attention last — its gradient
        // needs dL/dw from the other three), so the GEMM / outer-product code exists once per kernel and stays in the
        // instruction cache; the per-head pieces are the epilogue of the hidden layer and the per-pair geometry =================
        Quat ddg{0.0f, 0.0f, 0.0f, 0.0f}, u{1.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {
            const int head = (it + 1) & 3;                         // HD_ROT, HD_TOR, HD_TRN, HD_ATT
            int C, w3o, dwo_w, dwo_b, xw_off = -1, ldw = 0, e0 = 0, ne = 0, xb_off = -1;
            if (head == HD_ROT)      { C = 4; w3o = kW3Rot; dwo_w = param_offset(LAYER, ROT2_W); dwo_b = param_offset(LAYER, ROT2_B);
                                       xw_off = param_offset(LAYER, ROT0_W) + 64; ldw = 68; e0 = GEO_LQ; ne = 4; xb_off = param_offset(LAYER, ROT0_B); }
            else if (head == HD_TOR) { C = PMHC_NTORS; w3o = kW3Tor; dwo_w = param_offset(LAYER, TOR2_W); dwo_b = param_offset(LAYER, TOR2_B); }
            else if (head == HD_TRN) { C = 1; w3o = kW3Trn; dwo_w = param_offset(LAYER, TRN2_W); dwo_b = param_offset(LAYER, TRN2_B);
                                       xb_off = param_offset(LAYER, TRN0_B); }
            else                     { C = 1; w3o = kW3Att; dwo_w = param_offset(LAYER, ATT2_W); dwo_b = param_offset(LAYER, ATT2_B);
                                       xw_off = param_offset(LAYER, ATT0_W) + 64; ldw = 66; e0 = GEO_D2; ne = 2; xb_off = param_offset(LAYER, ATT0_B); }
            const float* Wh = Whs + head * 4096;
            // this thread's two running second-layer weight-gradient sums (L2): loaded now, used four phases later
            float dwo_old0 = 0.0f, dwo_old1 = 0.0f;
            if (fg < C) {
                const float* q = direct + (dwo_w - base) + fg * 64 + 8 * warp + 2 * ft;
                dwo_old0 = ldcg_early(q);
                dwo_old1 = ldcg_early(q + 1);
            }

            // ---- hidden layer: relu(W_h m + per-head extras) -> BufB (own columns) ----
            zero_acc(acc);
            gemm64_tf32<false>(Wh, bufA, acc);
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = 16 * mt + fg + 8 * h;
                    // per-unit extras: rotation W_q (4) + bias; attention w_d, w_q, bias; translation bias; torsion T_t[i] per column
                    float4 wx = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    float bias = 0.0f;
                    if (head == HD_ROT) { wx = *reinterpret_cast<const float4*>(S + M.f.PkRotQ + 4 * n); bias = S[M.f.PkMisc + 4 * n + 2]; }
                    else if (head == HD_ATT) { wx = *reinterpret_cast<const float4*>(S + M.f.PkAtt + 4 * n); }
                    else if (head == HD_TRN) { bias = S[M.f.PkMisc + 4 * n + 0]; }
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const int pc = 16 * warp + 8 * nt + 2 * ft;
                        float o[2];
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const float av = acc[mt][nt][2 * h + c];
                            float s;
                            if (head == HD_ROT) {
                                const float* ge = sGeo + pc + c;
                                s = av + bias + wx.x * ge[0] + wx.y * ge[kLdt] + wx.z * ge[2 * kLdt] + wx.w * ge[3 * kLdt];
                            } else if (head == HD_ATT) {
                                s = (wx.z + av) + fmaf(wx.y, sGeo[GEO_QDF * kLdt + pc + c], wx.x * sGeo[GEO_D2F * kLdt + pc + c]);
                            } else if (head == HD_TOR) {
                                s = av + S[M.f.Tt + sPI[pc + c] * kHid + n];
                            } else {
                                s = av + bias;
                            }
                            o[c] = tf32r(fmaxf(s, 0.0f));
                        }
                        *reinterpret_cast<float2*>(bufB + n * kLdt + pc) = make_float2(o[0], o[1]);
                    }
                }
            if (head != HD_ATT) {
                // second layer on the warp's own columns (the attention logit was saved by the forward)
                __syncwarp();
                rows8_tf32(W3s + w3o, C, bufB, o2);
                store_rows(sDout, C);
            }
            __syncthreads();

            // ---- per pair: head outputs -> geometry -> dL / d outputs (one thread per pair) ----
            if (owner) {
                if (head == HD_ROT) {
                    const Quat dl{sigmoidf(sDout[0 * kLdt + p] + S[M.f.Scal + SC_ROT2B + 0]), sigmoidf(sDout[1 * kLdt + p] + S[M.f.Scal + SC_ROT2B + 1]),
                                  sigmoidf(sDout[2 * kLdt + p] + S[M.f.Scal + SC_ROT2B + 2]), sigmoidf(sDout[3 * kLdt + p] + S[M.f.Scal + SC_ROT2B + 3])};
                    u = qmul(dl, qinvj);
                    const Quat dg = qmul(qj, u);
                    const Quat dG{rg[0], rg[1], rg[2], rg[3]};
                    dLdw += qdot(dG, dg);
                    ddg = qscale(dG, w);
                    const Quat du = qmul_grad_b(qj, ddg);         // dg = qj * u
                    const Quat ddl = qmul_grad_a(du, qinvj);      // u = dl * qinvj
                    sDout[0 * kLdt + p] = tf32r(ddl.w * dl.w * (1.0f - dl.w));
                    sDout[1 * kLdt + p] = tf32r(ddl.x * dl.x * (1.0f - dl.x));
                    sDout[2 * kLdt + p] = tf32r(ddl.y * dl.y * (1.0f - dl.y));
                    sDout[3 * kLdt + p] = tf32r(ddl.z * dl.z * (1.0f - dl.z));
                    if (IN_GRADS) {
                        dqj_part = qmul_grad_a(ddg, u);            // dg = qj * u
                        dqinv_part = qmul_grad_b(dl, du);          // u = dl * qinvj
                    }
                } else if (head == HD_TOR) {
                    if (IN_GRADS && act) {
                        // rotation head, second part: dL / d local quaternion arrived (sDx rows 0..3)
                        const Quat dlqq{sDx[0 * kLdt + p], sDx[1 * kLdt + p], sDx[2 * kLdt + p], sDx[3 * kLdt + p]};
                        const Quat dqinv = qadd(qmul_grad_a(dlqq, v), dqinv_part);   // lq = qinvj * v
                        const Quat dv = qmul_grad_b(qinvj, dlqq);
                        const Quat dqi = qmul_grad_a(dv, qj);                         // v = qi * qj
                        Quat dqj = qadd(qmul_grad_b(qi, dv), dqj_part);
                        dqj = qadd(dqj, qinv_grad(qj, dqinv));
                        gi[0] += dqi.w; gi[1] += dqi.x; gi[2] += dqi.y; gi[3] += dqi.z;
                        if (pep) atomic_add_quat(S + M.dQ + j * 4, dqj);
                    }
#pragma unroll
                    for (int c = 0; c < PMHC_NTORS; ++c) {
                        const float da = sDout[c * kLdt + p] + S[M.f.Scal + SC_TOR2B + c];
                        dLdw = fmaf(rg[4 + c], da, dLdw);
                        sDout[c * kLdt + p] = tf32r(w * rg[4 + c]);
                    }
                } else if (head == HD_TRN) {
                    const float sc = sDout[p] + S[M.f.Scal + SC_TRN2B];
                    const float dXr = rg[11] * rx + rg[12] * ry + rg[13] * rz;
                    dLdw = fmaf(sc, dXr, dLdw);
                    sDout[p] = tf32r(w * dXr);
                    if (IN_GRADS && act) {
                        const float f = w * sc;
                        gi[4] += f * rg[11]; gi[5] += f * rg[12]; gi[6] += f * rg[13];
                        if (pep) {
                            atomicAdd(S + M.dX + j * 3 + 0, -f * rg[11]); atomicAdd(S + M.dX + j * 3 + 1, -f * rg[12]); atomicAdd(S + M.dX + j * 3 + 2, -f * rg[13]);
                        }
                    }
                } else {
                    // softmax backward with the saved row statistics (see pair_pass for the saturated-row rule)
                    sDout[p] = tf32r((w == 1.0f) ? 0.0f : w * (dLdw - c_i));
                }
            }
            __syncthreads();
            dwo_tf32(sDout, C, bufB, direct + (dwo_w - base), direct + (dwo_b - base), dwo_old0, dwo_old1);
            __syncthreads();
            cols8_tf32(W3s + w3o, C, sDout, acc);
            mask_store_dpre();
            __syncwarp();
            gemm64_tf32<true>(Wh, bufB, dm);
            if (IN_GRADS && (head == HD_ROT || head == HD_ATT)) {
                // dL / d local quaternion (rotation: sDx rows 0..3) or dL / d (-d2, qdot2) (attention: rows 4, 5)
                rows8_tf32(Wxs + (head == HD_ROT ? kWxRot : kWxAtt), head == HD_ROT ? 4 : 2, bufB, o2);
                store_rows(sDx + (head == HD_ROT ? 0 : 4 * kLdt), head == HD_ROT ? 4 : 2);
            }
            float4 told[4];
            outer_prefetch(tiles + (head + 1) * 4096, told);
            __syncthreads();
            outer_tf32(bufB, bufA, sGeo, tiles + (head + 1) * 4096, told, xw_off >= 0 ? direct + (xw_off - base) : nullptr, ldw, e0, ne,
                       xb_off >= 0 ? direct + (xb_off - base) : nullptr);
            if (head == HD_TOR) accumulate_rows<kLdt>(bufB, S + M.dTt, kHid, I, L, Wr, pass_base, npass);
            __syncthreads();
        }
        if (IN_GRADS && owner) {
            if (act) {
                const float gd = sDx[4 * kLdt + p], gq = sDx[5 * kLdt + p];
                const float f = -gd * 2.0f;                 // d(-d2) = gd
                gi[4] += f * rx; gi[5] += f * ry; gi[6] += f * rz;
                const float fq = gq * 2.0f * dotq;
                gi[0] += fq * qj.w; gi[1] += fq * qj.x; gi[2] += fq * qj.y; gi[3] += fq * qj.z;
                if (pep) {
                    atomicAdd(S + M.dX + j * 3 + 0, -f * rx); atomicAdd(S + M.dX + j * 3 + 1, -f * ry); atomicAdd(S + M.dX + j * 3 + 2, -f * rz);
                    atomic_add_quat(S + M.dQ + j * 4, qscale(qi, fq));
                }
            }
            // the i side of the input gradients: one column per pair, reduced per row after the next barrier (every pair of
            // a row adds to the same q_i / x_i: shared-memory atomics would serialise a whole row)
#pragma unroll
            for (int c = 0; c < 7; ++c) sDout[c * kLdt + p] = act ? gi[c] : 0.0f;
        }
        // ---- dm (+ the message-sum gradient of layer 1) -> BufB, inactive columns zero ----
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int pc = 16 * warp + 8 * nt + 2 * ft + c;
                const bool on_col = pc < npass;
                const float* dms = S + M.dMsum + sPI[pc] * kHid;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int n = 16 * mt + fg + 8 * h;
                        float v = dm[mt][nt][2 * h + c];
                        if (LAYER == 0) v += dms[n];
                        bufB[n * kLdt + pc] = on_col ? tf32r(v) : 0.0f;
                    }
            }
    } else {
        if (owner) {
#pragma unroll
            for (int r = 0; r < 6; ++r) sGeo[r * kLdt + p] = 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) bufB[(n0 + k) * kLdt + p] = act ? tf32r(mult * S[M.dMsum + i * kHid + n0 + k]) : 0.0f;
    }

    // ---- message MLP backward: dW2 += dm (x) m1, dm1 = relu'(.) W2^T dm, then the per-node reductions ----
    {
        float m1h[32];
        compute_m1<LAYER, 32>(m1h, n0, S, M, a.params, ajt, Kpad, i, j);
#pragma unroll
        for (int k = 0; k < 32; ++k) bufA[(n0 + k) * kLdt + p] = tf32r(m1h[k]);
    }
    float4 told[4];
    outer_prefetch(tiles + T_W2 * 4096, told);
    __syncthreads();
    if (!HEADS) list_peptide_columns();
    if (HEADS && IN_GRADS) {
        for (int idx = tid; idx < L * 7; idx += kBwdThreads) {
            const int rl = idx / 7, c = idx - rl * 7;
            const int lo = max(rl * Wr, pass_base), hi = min((rl + 1) * Wr, pass_base + npass);
            float sum = 0.0f;
            for (int gp = lo; gp < hi; ++gp) sum += sDout[c * kLdt + (gp - pass_base)];
            const int ri = I[IN_ROWS + rl];
            if (hi > lo) {
                if (c < 4) S[M.dQ + ri * 4 + c] += sum;
                else S[M.dX + ri * 3 + (c - 4)] += sum;
            }
        }
    }
    zero_acc(acc);
    gemm64_tf32<true>(W2s, bufB, acc);
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 16 * mt + fg + 8 * h;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int pc = 16 * warp + 8 * nt + 2 * ft;
                const float2 m1v = *reinterpret_cast<const float2*>(bufA + k * kLdt + pc);
                if (!(m1v.x > 0.0f)) acc[mt][nt][2 * h] = 0.0f;
                if (!(m1v.y > 0.0f)) acc[mt][nt][2 * h + 1] = 0.0f;
            }
        }
    outer_tf32(bufB, bufA, sGeo, tiles + T_W2 * 4096, told, nullptr, 0, 0, 0, direct + (param_offset(LAYER, MSG2_B) - base));
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 16 * mt + fg + 8 * h;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int pc = 16 * warp + 8 * nt + 2 * ft;
                *reinterpret_cast<float2*>(bufB + k * kLdt + pc) = make_float2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]);
            }
        }
    __syncthreads();
    {   // peptide neighbours: dA_j[j] and dW_e[rel] get the column; the few such columns are spread over all threads
        const int n_pl = sPl[0];
        for (int idx = tid; idx < n_pl * kHid; idx += kBwdThreads) {
            const int e = sPl[4 + (idx >> 6)], k = idx & 63;
            const float v = bufB[k * kLdt + (e & 255)];
            atomicAdd(S + M.dAjPep + ((e >> 8) & 255) * kLdN + k, v);
            atomicAdd(S + M.dWe + (e >> 16) * kLdN + k, v);
        }
    }
    if (act && !HEADS && j >= kN) {
        // masked pocket slot with non-zero features (rare): straight to the A_j^T gradient scratch
        for (int k = 0; k < 32; ++k) atomicAdd(dajt + (n0 + k) * Kpad + j, bufB[(n0 + k) * kLdt + p]);
    }
    accumulate_rows<kLdt>(bufB, S + M.dAi, kLdN, I, L, Wr, pass_base, npass);
    if (HEADS) {
        // dA_j^T[k][j] += sum over rows of dm1 for the valid pocket columns of this pass
        const int rl_lo = pass_base / Wr, rl_hi = (pass_base + npass - 1) / Wr;
        const int n_items = n_pocket_cols * kHid;
        for (int idx0 = tid; idx0 < n_items; idx0 += 4 * kBwdThreads) {   // four L2 read-modify-writes in flight per thread
            float old[4];
            int addr[4], kk[4], ee[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = idx0 + u * kBwdThreads;
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
                    if (col >= 0 && col < npass) sum += bufB[kk[u] * kLdt + col];
                }
                dajt[addr[u]] = old[u] + sum;
            }
        }
    }
    __syncthreads();
}

#ifdef PMHC_T5_STAMPS
__device__ long long node_dbg[24];
#define NSTAMP(k) do { __syncthreads(); if (blockIdx.x == 0 && threadIdx.x == 0) { long long now_ = clock64(); node_dbg[k] += now_ - nst_; nst_ = now_; } } while (0)
#else
#define NSTAMP(k) do { } while (0)
#endif
// dst[idx] += f(idx) for idx < n on the CTA's L2-resident partial: U read-modify-writes of a thread are in flight together (their
// round trips overlap each other and the products), instead of one L2 latency per element
template <int U, class F>
__device__ __forceinline__ void rmw_batched(float* __restrict__ dst, int n, const int NT, F f) {
    for (int idx0 = threadIdx.x; idx0 < n; idx0 += U * NT) {
        float old[U];
#pragma unroll
        for (int u = 0; u < U; ++u) old[u] = idx0 + u * NT < n ? ldcg_early(dst + idx0 + u * NT) : 0.0f;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (idx0 + u * NT < n) dst[idx0 + u * NT] = old[u] + f(idx0 + u * NT);
    }
}

// Layer 1's node feature MLP backward (model.py:151, :407) for one complex: dL / d relu(o1) -> feature_mlp.{0,2} weight gradients
// (into the CTA's partial) and dL / d(message sum) in S[M.dMsum].  Needs S[M.f.H], S[M.f.Msum] and the row list I[IN_ROWS..].
__device__ __forceinline__ void bwd_feature_mlp(float* S, const BwdMap& M, const BwdArgs& g, int b, const int* I, int L,
                                                float* __restrict__ direct, const int NT, const bool weights_staged = false) {
    const LayerArgs& a = g.a;
    constexpr int base = param_offset(0, 0);
    const int tid = threadIdx.x;
#ifdef PMHC_T5_STAMPS
    long long nst_ = clock64();
#endif
    const float* f0w = a.params + param_offset(0, FEAT0_W);
    const float* f0b = a.params + param_offset(0, FEAT0_B);
    const float* f2w = a.params + param_offset(0, FEAT2_W);
    constexpr int ldf = kH1 + kHid;
    float* hid = S + M.BufA;                 // [16][65] relu(feature_mlp.0(...))
    float* dO = S + M.BufA + kN * kLdN;      // [16][65] dL/do (after the relu mask)
    float* dhid = S + M.BufB;                // [16][65]
    // both weight matrices of the feature MLP staged behind those (coalesced copies, many loads in flight): every product
    // below then reads shared memory instead of walking an L2-latency chain per term
    NSTAMP(8);
    float* f0s = S + M.BufA + 2 * kN * kLdN;  // [64][87]
    float* f2s = S + M.BufB + kN * kLdN;      // [64][64]
    if (!weights_staged) {      // (a caller that keeps BufA / BufB untouched between complexes stages them once)
#pragma unroll 8
        for (int idx = tid; idx < kHid * ldf; idx += NT) f0s[idx] = __ldg(f0w + idx);
#pragma unroll 8
        for (int idx = tid; idx < kHid * kHid; idx += NT) f2s[idx] = __ldg(f2w + idx);
    }
    __syncthreads();
    NSTAMP(9);
    for (int idx = tid; idx < L * kHid; idx += NT) {
        int r = idx >> 6, n = idx & 63;
        int i = I[IN_ROWS + r];
        const float* w = f0s + n * ldf;
        const float* h = S + M.f.H + i * kLdN;
        const float* ms = S + M.f.Msum + i * kHid;
        float acc = f0b[n];
#pragma unroll
        for (int c = 0; c < kH1; ++c) acc = fmaf(w[c], h[c], acc);
#pragma unroll 16
        for (int c = 0; c < kHid; ++c) acc = fmaf(w[kH1 + c], ms[c], acc);
        hid[r * kLdN + n] = fmaxf(acc, 0.0f);
        const size_t node = (size_t)b * kN + i;
        dO[r * kLdN + n] = g.feat_post[node * kHid + n] > 0.0f ? g.d_feat_out[node * kHid + n] : 0.0f;
    }
    __syncthreads();
    NSTAMP(10);
    for (int idx = tid; idx < L * kHid; idx += NT) {
        int r = idx >> 6, n = idx & 63;
        float acc = 0.0f;
#pragma unroll 16
        for (int n2 = 0; n2 < kHid; ++n2) acc = fmaf(f2s[n2 * kHid + n], dO[r * kLdN + n2], acc);
        dhid[r * kLdN + n] = hid[r * kLdN + n] > 0.0f ? acc : 0.0f;
    }
    NSTAMP(11);
    // feature_mlp.2: dW[n2][n] += sum_r dO[r][n2] hid[r][n]; db[n2] += sum_r dO[r][n2]
    static_assert(param_offset(0, FEAT2_B) == param_offset(0, FEAT2_W) + kHid * kHid, "feature_mlp.2 weight and bias are adjacent");
    rmw_batched<8>(direct + (param_offset(0, FEAT2_W) - base), kHid * kHid + kHid, NT, [&](int idx) {
        float acc = 0.0f;
        if (idx < kHid * kHid) {
            const int n2 = idx >> 6, n = idx & 63;
            for (int r = 0; r < L; ++r) acc = fmaf(dO[r * kLdN + n2], hid[r * kLdN + n], acc);
        } else {
            const int n2 = idx - kHid * kHid;
            for (int r = 0; r < L; ++r) acc += dO[r * kLdN + n2];
        }
        return acc;
    });
    __syncthreads();
    NSTAMP(12);
    // feature_mlp.0: dW[n][c] += sum_r dhid[r][n] cat(h, msum)[r][c]; db[n] += sum_r dhid[r][n]
    static_assert(param_offset(0, FEAT0_B) == param_offset(0, FEAT0_W) + kHid * ldf, "feature_mlp.0 weight and bias are adjacent");
    rmw_batched<8>(direct + (param_offset(0, FEAT0_W) - base), kHid * ldf + kHid, NT, [&](int idx) {
        float acc = 0.0f;
        if (idx < kHid * ldf) {
            const int n = idx / ldf, c = idx - n * ldf;
            for (int r = 0; r < L; ++r) {
                const int i = I[IN_ROWS + r];
                const float x = c < kH1 ? S[M.f.H + i * kLdN + c] : S[M.f.Msum + i * kHid + (c - kH1)];
                acc = fmaf(dhid[r * kLdN + n], x, acc);
            }
        } else {
            const int n = idx - kHid * ldf;
            for (int r = 0; r < L; ++r) acc += dhid[r * kLdN + n];
        }
        return acc;
    });
    NSTAMP(13);
    for (int idx = tid; idx < L * kHid; idx += NT) {
        int r = idx >> 6, k = idx & 63;
        float acc = 0.0f;
#pragma unroll 16
        for (int n = 0; n < kHid; ++n) acc = fmaf(f0s[n * ldf + kH1 + k], dhid[r * kLdN + n], acc);
        S[M.dMsum + I[IN_ROWS + r] * kHid + k] = acc;
    }

}

// Per-complex prologue shared by every backward kernel: the row-level backward of the output normalisation / torsion rotation /
// translation (RowG), and for layer 1 the node-feature-MLP backward (model.py:151, :407) that produces dL / d(message sum).
// Every thread of the CTA takes part (NT = blockDim.x); the caller synchronises afterwards.
template <int LAYER, bool FEAT = true>
__device__ __forceinline__ void bwd_prologue(float* S, const BwdMap& M, const BwdArgs& g, int b, const int* I, int L, int W,
                                             float* __restrict__ direct, const int NT) {
    const LayerArgs& a = g.a;
    constexpr bool IN_GRADS = (LAYER == 1);
    constexpr int base = param_offset(LAYER, 0);
    const int tid = threadIdx.x;
#ifdef PMHC_T5_STAMPS
    long long nst_ = clock64();
#endif
    // ---------------- row level: output normalisation, q' = g * q_i, torsion rotation, x' = x + Xa ----------------
    if (tid < L) {
        const int i = I[IN_ROWS + tid];
        const size_t node = (size_t)b * kN + i;
        const float* rs = g.rowstat + node * PMHC_ROWSTAT;
        const float* dof = g.d_frames_out + node * 7;
        const float* dot_ = g.d_tors_out + node * 14;
        const Quat G{rs[1], rs[2], rs[3], rs[4]};
        const Quat qi{S[M.f.Q + i * 4], S[M.f.Q + i * 4 + 1], S[M.f.Q + i * 4 + 2], S[M.f.Q + i * 4 + 3]};
        const bool hasnb = W > 0;
        const Quat gq = hasnb ? qnormalize(G) : Quat{1.0f, 0.0f, 0.0f, 0.0f};
        const Quat qp = qmul(gq, qi);
        const Quat dqp = qnormalize_grad(qp, Quat{dof[0], dof[1], dof[2], dof[3]});
        const Quat dgq = qmul_grad_a(dqp, qi);
        const Quat dG = hasnb ? qnormalize_grad(G, dgq) : Quat{0.0f, 0.0f, 0.0f, 0.0f};
        float* rg = S + M.RowG + i * 16;
        rg[0] = dG.w; rg[1] = dG.x; rg[2] = dG.y; rg[3] = dG.z;
        float cacc = qdot(dG, G);
        if (IN_GRADS) {
            const Quat dqi = qmul_grad_b(gq, dqp);
            S[M.dQ + i * 4 + 0] += dqi.w; S[M.dQ + i * 4 + 1] += dqi.x; S[M.dQ + i * 4 + 2] += dqi.y; S[M.dQ + i * 4 + 3] += dqi.z;
        }
        for (int c = 0; c < PMHC_NTORS; ++c) {
            float sn, cs;
            sincosf(rs[5 + c], &sn, &cs);
            const float ts = S[M.f.Tors + i * 14 + 2 * c], tc = S[M.f.Tors + i * 14 + 2 * c + 1];
            const float ds_ = dot_[2 * c], dc_ = dot_[2 * c + 1];
            const float dS = ds_ * tc - dc_ * ts, dC = ds_ * ts + dc_ * tc;
            const float dDa = dS * cs - dC * sn;
            rg[4 + c] = dDa;
            cacc = fmaf(dDa, rs[5 + c], cacc);
            if (IN_GRADS) {
                S[M.dTors + i * 14 + 2 * c] += ds_ * cs - dc_ * sn;
                S[M.dTors + i * 14 + 2 * c + 1] += ds_ * sn + dc_ * cs;
            }
        }
        for (int c = 0; c < 3; ++c) {
            rg[11 + c] = dof[4 + c];
            cacc = fmaf(dof[4 + c], rs[12 + c], cacc);
            if (IN_GRADS) S[M.dX + i * 3 + c] += dof[4 + c];
        }
        rg[14] = cacc;
        rg[15] = rs[0];
    }

    if (LAYER == 0 && FEAT) bwd_feature_mlp(S, M, g, b, I, L, direct, NT);
}

// Per-complex node level shared by every backward kernel: message_mlp.0, torsion_mlp.0[:, 64:78] and biases from the per-node
// accumulators; layer 2 also writes the input gradients.  LDC = row stride of the staging tile at M.BufA; TORB = torsion_mlp.0.bias
// from the per-row sums (false when the caller gets it from its own bias column).
template <int LAYER, int LDC, bool TORB, bool TILED = false>
__device__ __forceinline__ void bwd_node_level(float* S, const BwdMap& M, const BwdArgs& g, int b, const int* I, float* __restrict__ dajt,
                                               float* __restrict__ direct, const int NT) {
    const LayerArgs& a = g.a;
    constexpr bool IN_GRADS = (LAYER == 1);
    constexpr int H = layer_H(LAYER);
    constexpr int ld1 = 2 * H + kEdge;
    constexpr int base = param_offset(LAYER, 0);
    const int tid = threadIdx.x;
    const int Kpad = a.Kpad, P = a.P;
#ifdef PMHC_T5_STAMPS
    long long nst_ = clock64();
#endif
    // ---------------- node level: message_mlp.0, torsion_mlp.0[:, 64:78] and biases; input gradients ----------------
    {
        // neighbour-feature columns fed by the pocket (cc < 22): sum_j dA_j[k] h_j[cc] over the peptide, then over the
        // pocket slots in ascending order; the pocket's dA_j^T (L2) and features are staged through shared memory in
        // chunks of 128 slots so the 80..480-term chains run from shared memory instead of one L2 round trip per term
        constexpr int ldc = LDC;
        float* scr = S + M.Dout;                 // [64][22] running sums (Dout + Ex are free outside the passes)
        float* stA = S + M.BufA;                 // [64][ldc] dA_j^T chunk
        float* stB = S + M.BufB;                 // [128][22] pocket features chunk
        for (int idx = tid; idx < kHid * PMHC_NFEAT; idx += NT) {
            const int k = idx / PMHC_NFEAT, cc = idx - k * PMHC_NFEAT;
            float acc = 0.0f;
            for (int j = 0; j < kN; ++j) acc = fmaf(S[M.dAjPep + j * kLdN + k], S[M.f.H + j * kLdN + cc], acc);
            scr[idx] = acc;
        }
        for (int p0 = 0; p0 < P; p0 += 128) {
            const int n = min(128, P - p0);
            __syncthreads();
            for (int idx = tid; idx < kHid * n; idx += NT) {
                const int k = idx / n, pp = idx - k * n;
                stA[k * ldc + pp] = __ldcg(dajt + k * Kpad + kN + p0 + pp);
            }
            const float* pf = a.pocket_feat + ((size_t)b * P + p0) * PMHC_NFEAT;
            for (int idx = tid; idx < n * PMHC_NFEAT; idx += NT) stB[idx] = __ldg(pf + idx);
            __syncthreads();
            if (TILED) {
                // register tile of 4 features x 2 pocket-feature columns over a quarter of the chunk's slots (5 shared loads per 8
                // products instead of 16); the four quarter sums are added in order afterwards
                float* part = stB + 128 * PMHC_NFEAT;     // [4][64 * 22]
                for (int item = tid; item < 16 * 11 * 4; item += NT) {
                    const int cc2 = item % 11, q = (item / 11) & 3, k4 = item / 44;
                    const int lo = (n * q) >> 2, hi = (n * (q + 1)) >> 2;
                    float acc[4][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}, {0.0f, 0.0f}, {0.0f, 0.0f}};
#pragma unroll 4
                    for (int pp = lo; pp < hi; ++pp) {
                        const float2 bv = *reinterpret_cast<const float2*>(stB + pp * PMHC_NFEAT + 2 * cc2);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float av = stA[(4 * k4 + u) * ldc + pp];
                            acc[u][0] = fmaf(av, bv.x, acc[u][0]);
                            acc[u][1] = fmaf(av, bv.y, acc[u][1]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        part[q * (kHid * PMHC_NFEAT) + (4 * k4 + u) * PMHC_NFEAT + 2 * cc2] = acc[u][0];
                        part[q * (kHid * PMHC_NFEAT) + (4 * k4 + u) * PMHC_NFEAT + 2 * cc2 + 1] = acc[u][1];
                    }
                }
                __syncthreads();
                for (int idx = tid; idx < kHid * PMHC_NFEAT; idx += NT)
                    scr[idx] += ((part[idx] + part[kHid * PMHC_NFEAT + idx]) + part[2 * kHid * PMHC_NFEAT + idx]) + part[3 * kHid * PMHC_NFEAT + idx];
            } else {
                for (int idx = tid; idx < kHid * PMHC_NFEAT; idx += NT) {
                    const int k = idx / PMHC_NFEAT, cc = idx - k * PMHC_NFEAT;
                    float acc = scr[idx];
#pragma unroll 8
                    for (int pp = 0; pp < n; ++pp) acc = fmaf(stA[k * ldc + pp], stB[pp * PMHC_NFEAT + cc], acc);
                    scr[idx] = acc;
                }
            }
        }
        __syncthreads();
        NSTAMP(0);
        float* dW1 = direct + (param_offset(LAYER, MSG0_W) - base);
        if (TILED) {
            // 4 features per item: the per-node gradients as 16-byte aligned rows (one broadcast 128-bit load feeds four products)
            float* dAiA = S + M.BufB;                // [16][64]
            float* dAjA = dAiA + kN * kHid;          // [16][64]
            for (int idx = tid; idx < kN * kHid; idx += NT) {
                dAiA[idx] = S[M.dAi + (idx >> 6) * kLdN + (idx & 63)];
                dAjA[idx] = S[M.dAjPep + (idx >> 6) * kLdN + (idx & 63)];
            }
            __syncthreads();
            for (int item = tid; item < 16 * ld1; item += NT) {
                const int k4 = item / ld1, c = item - k4 * ld1;
                float old[4], acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int u = 0; u < 4; ++u) old[u] = ldcg_early(dW1 + (4 * k4 + u) * ld1 + c);
                if (c < H || (c < 2 * H && c - H >= PMHC_NFEAT)) {
                    const float* av = (c < H ? dAiA : dAjA) + 4 * k4;
                    const float* hv = S + M.f.H + (c < H ? c : c - H);
#pragma unroll
                    for (int i = 0; i < kN; ++i) {
                        const float4 a4 = *reinterpret_cast<const float4*>(av + i * kHid);
                        const float h = hv[i * kLdN];
                        acc[0] = fmaf(a4.x, h, acc[0]); acc[1] = fmaf(a4.y, h, acc[1]);
                        acc[2] = fmaf(a4.z, h, acc[2]); acc[3] = fmaf(a4.w, h, acc[3]);
                    }
                } else if (c < 2 * H) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[u] = scr[(4 * k4 + u) * PMHC_NFEAT + (c - H)];
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[u] = S[M.dWe + (c - 2 * H) * kLdN + 4 * k4 + u];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) dW1[(4 * k4 + u) * ld1 + c] = old[u] + acc[u];
            }
        } else
        rmw_batched<8>(dW1, kHid * ld1, NT, [&](int idx) {
            const int k = idx / ld1, c = idx - k * ld1;
            float acc = 0.0f;
            if (c < H) {
#pragma unroll
                for (int i = 0; i < kN; ++i) acc = fmaf(S[M.dAi + i * kLdN + k], S[M.f.H + i * kLdN + c], acc);
            } else if (c < 2 * H) {
                const int cc = c - H;
                if (cc < PMHC_NFEAT) {
                    acc = scr[k * PMHC_NFEAT + cc];
                } else {
#pragma unroll
                    for (int j = 0; j < kN; ++j) acc = fmaf(S[M.dAjPep + j * kLdN + k], S[M.f.H + j * kLdN + cc], acc);
                }
            } else {
                acc = S[M.dWe + (c - 2 * H) * kLdN + k];
            }
            return acc;
        });
        NSTAMP(1);
        for (int k = tid; k < kHid; k += NT) {
            float acc = 0.0f, acct = 0.0f;
            for (int i = 0; i < kN; ++i) {
                acc += S[M.dAi + i * kLdN + k];
                acct += S[M.dTt + i * kHid + k];
            }
            direct[(param_offset(LAYER, MSG0_B) - base) + k] += acc;
            if (TORB) direct[(param_offset(LAYER, TOR0_B) - base) + k] += acct;
        }
        for (int idx = tid; idx < kHid * 14; idx += NT) {
            int n = idx / 14, c = idx - n * 14;
            float acc = 0.0f;
            for (int i = 0; i < kN; ++i) acc = fmaf(S[M.dTt + i * kHid + n], S[M.f.Tors + i * 14 + c], acc);
            direct[(param_offset(LAYER, TOR0_W) - base) + n * 78 + 64 + c] += acc;
        }
        if (IN_GRADS) {
            const float* tor0 = a.params + param_offset(LAYER, TOR0_W);
            const float* msg0 = a.params + param_offset(LAYER, MSG0_W);
            // the node columns of message_mlp.0 and the torsion columns of torsion_mlp.0 staged in the (now free) pass tiles:
            // coalesced copies with many loads in flight instead of one L2 round trip per term of the products below
            NSTAMP(2);
            constexpr int LDQ = 2 * H + 1;
            float* wq = S + M.BufA;                  // [64][2 H + 1]
            float* tx = S + M.BufB;                  // [64][15]
            static_assert(kHid * LDQ <= kHid * kLdc, "message_mlp.0's node columns must fit the staging tile");
            __syncthreads();
#pragma unroll 8
            for (int idx = tid; idx < kHid * 2 * H; idx += NT) {
                const int k = idx / (2 * H), c = idx - k * (2 * H);
                wq[k * LDQ + c] = __ldg(msg0 + k * ld1 + c);
            }
            for (int idx = tid; idx < kHid * 14; idx += NT) {
                const int n = idx / 14, c = idx - n * 14;
                tx[n * 15 + c] = __ldg(tor0 + n * 78 + 64 + c);
            }
            __syncthreads();
            NSTAMP(3);
            for (int idx = tid; idx < kN * 7; idx += NT) {
                int i = idx / 7, c = idx - i * 7;
                g.d_frames_in[((size_t)b * kN + i) * 7 + c] = c < 4 ? S[M.dQ + i * 4 + c] : S[M.dX + i * 3 + (c - 4)];
            }
            for (int idx = tid; idx < kN * 14; idx += NT) {
                int i = idx / 14, c = idx - i * 14;
                float acc = S[M.dTors + idx];
#pragma unroll 16
                for (int n = 0; n < kHid; ++n) acc = fmaf(tx[n * 15 + c], S[M.dTt + i * kHid + n], acc);
                g.d_tors_in[(size_t)b * kN * 14 + idx] = acc;
            }
            if (TILED) {
                // 4 nodes per item: the per-node gradients transposed ([k][16 nodes]) so that one 128-bit load covers four nodes
                float* dAiT = S + M.BufB + 1024;     // (behind tx)
                float* dAjT = dAiT + kN * kHid;
                for (int idx = tid; idx < kN * kHid; idx += NT) {
                    const int k = idx >> 4, i = idx & 15;
                    dAiT[idx] = S[M.dAi + i * kLdN + k];
                    dAjT[idx] = S[M.dAjPep + i * kLdN + k];
                }
                __syncthreads();
                for (int item = tid; item < 4 * kHid; item += NT) {
                    const int i4 = item >> 6, c = item & 63;
                    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 8
                    for (int k = 0; k < kHid; ++k) {
                        const float4 a4 = *reinterpret_cast<const float4*>(dAiT + k * kN + 4 * i4);
                        const float4 b4 = *reinterpret_cast<const float4*>(dAjT + k * kN + 4 * i4);
                        const float w0 = wq[k * LDQ + c], w1 = wq[k * LDQ + H + c];
                        acc[0] = fmaf(a4.x, w0, acc[0]); acc[0] = fmaf(b4.x, w1, acc[0]);
                        acc[1] = fmaf(a4.y, w0, acc[1]); acc[1] = fmaf(b4.y, w1, acc[1]);
                        acc[2] = fmaf(a4.z, w0, acc[2]); acc[2] = fmaf(b4.z, w1, acc[2]);
                        acc[3] = fmaf(a4.w, w0, acc[3]); acc[3] = fmaf(b4.w, w1, acc[3]);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) g.d_feat_in[((size_t)b * kN + 4 * i4 + u) * kHid + c] = acc[u];
                }
            } else
            for (int idx = tid; idx < kN * kHid; idx += NT) {
                int i = idx >> 6, c = idx & 63;
                float acc = 0.0f;
#pragma unroll 8
                for (int k = 0; k < kHid; ++k) {
                    acc = fmaf(S[M.dAi + i * kLdN + k], wq[k * LDQ + c], acc);
                    acc = fmaf(S[M.dAjPep + i * kLdN + k], wq[k * LDQ + H + c], acc);
                }
                g.d_feat_in[(size_t)b * kN * kHid + idx] = acc;
            }
            NSTAMP(4);
        }
    }
}


template <int LAYER, bool TC>
__global__ void __launch_bounds__(kBwdThreads, 1) egnn_layer_backward_kernel(BwdArgs g) {
    extern __shared__ __align__(16) float S[];
    const LayerArgs& a = g.a;
    const BwdMap M = make_bwd_map(a.Kpad, TC);
    constexpr bool IN_GRADS = (LAYER == 1);
    constexpr int H = layer_H(LAYER);
    constexpr int ld1 = 2 * H + kEdge;
    constexpr int base = param_offset(LAYER, 0);
    constexpr int layer_numel = param_offset(LAYER + 1, 0) - base;
    const int tid = threadIdx.x;
    const int Kpad = a.Kpad, P = a.P;
    int* I = reinterpret_cast<int*>(S + M.f.Ints);
    float* ajt = a.ajt_ws + (size_t)blockIdx.x * kHid * Kpad;
    float* dajt = g.dajt_ws + (size_t)blockIdx.x * kHid * Kpad;
    float* tiles = g.partial + (size_t)blockIdx.x * g.partial_stride;
    float* direct = tiles + kTileFloats;

    for (int idx = tid; idx < kTileFloats + layer_numel; idx += kBwdThreads) tiles[idx] = 0.0f;
    stage_layer_weights_bwd<LAYER, TC>(S, M, a.params);
    __syncthreads();

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const ComplexInfo ci = setup_complex<LAYER>(S, M.f, a, b, ajt);
        const int L = ci.L;
        const int W = (L - 1) + ci.nv;
        for (int idx = tid; idx < M.grads_end - M.dAi; idx += kBwdThreads) S[M.dAi + idx] = 0.0f;
        for (int idx = tid; idx < kHid * Kpad; idx += kBwdThreads) dajt[idx] = 0.0f;
        if (LAYER == 0)
            for (int idx = tid; idx < kN * kHid; idx += kBwdThreads) S[M.f.Msum + idx] = g.msum[(size_t)b * kN * kHid + idx];
        __syncthreads();

        bwd_prologue<LAYER>(S, M, g, b, I, L, W, direct, kBwdThreads);
        __syncthreads();

        // ---------------- attention-carrying pairs ----------------
        const int total = L > 0 ? L * W : 0;
        const int pcol = tid & (kBwdPairs - 1);      // two threads (tid, tid + 128) per pair column
        for (int pass_base = 0; pass_base < total; pass_base += kBwdPairs) {
            const int npass = min(kBwdPairs, total - pass_base);
            const bool act = pcol < npass;
            const PairRef pr = decode_full_pair(I, act ? pass_base + pcol : pass_base, W, L, 0, act);
            if (TC) pair_pass_tc<LAYER, true>(S, M, g, pr, 1.0f, b, ajt, dajt, tiles, direct, I, L, W, pass_base, npass, ci.nv, L - 1);
            else pair_pass<LAYER, true>(S, M, g, pr, 1.0f, b, ajt, dajt, tiles, direct, I, L, W, pass_base, npass, ci.nv, L - 1);
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
                if (TC) pair_pass_tc<LAYER, false>(S, M, g, pr, mult, b, ajt, dajt, tiles, direct, I, L, W2, pass_base, npass, 0, 0);
                else pair_pass<LAYER, false>(S, M, g, pr, mult, b, ajt, dajt, tiles, direct, I, L, W2, pass_base, npass, 0, 0);
            }
        }
        __syncthreads();

        bwd_node_level<LAYER, TC ? kLdt : kLdc, true>(S, M, g, b, I, dajt, direct, kBwdThreads);
        __syncthreads();
    }

    // ---------------- fold the tile-owner partials into the parameter layout of this CTA's `direct` region ----------------
    __syncthreads();
    // (these first-layer weight columns of `direct` are written only here, and `direct` starts at zero: plain stores; the 16 L2
    // loads of a thread per matrix are all in flight together)
    auto fold = [&](int T, int off, int ld) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldcg(tiles + T * 4096 + tid + u * kBwdThreads);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int r = tid + u * kBwdThreads;
            const int t = r >> 4, e = r & 15;
            int k = (t & 15) + 16 * (e >> 2), n = (t >> 4) + 16 * (e & 3);
            if (TC) tc_tile_coord(t, e, n, k);
            direct[(off - base) + n * ld + k] = v[u];
        }
    };
    fold(T_W2, param_offset(LAYER, MSG2_W), 64);
    fold(T_ATT, param_offset(LAYER, ATT0_W), 66);
    fold(T_ROT, param_offset(LAYER, ROT0_W), 68);
    fold(T_TOR, param_offset(LAYER, TOR0_W), 78);
    fold(T_TRN, param_offset(LAYER, TRN0_W), 64);
}

// grad[p] += sum over CTAs of direct[cta][p]
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int stride, int n_cta, int numel,
                                       float* __restrict__ grad) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= numel) return;
    float acc = 0.0f;
    // 16 L2 loads in flight, added in CTA order (the fixed order that makes the result independent of timing)
    for (int c0 = 0; c0 < n_cta; c0 += 16) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = c0 + u < n_cta ? __ldcg(partial + (size_t)(c0 + u) * stride + kTileFloats + p) : 0.0f;
#pragma unroll
        for (int u = 0; u < 16; ++u)
            if (c0 + u < n_cta) acc += v[u];
    }
    grad[p] += acc;
}

}  // namespace pmhc

#include "egnn_backward_t5.cuh"

namespace pmhc {

struct BwdWorkspace {
    float *ajt, *dajt, *partial, *d_frames1, *d_tors1, *d_feat1;
    float *red, *scale;      // tcgen05 mode: reduced partials of a layer, the two layers' operand scales
    int* sched_ws;           // tcgen05 mode: units [B], per-CTA schedule [sms] (int4), per-complex segments [B] (int2)
    float *ajt_all, *rec_all; // tcgen05 mode: per-complex neighbour projections and records from the setup pre-kernel
    float* dmsum_g;          // tcgen05 mode, layer 1: dL / d(message sum) and W2^T of it per complex, from the node pre-kernel
    float *acc, *dajt_all;   // tcgen05 mode: per-complex gradient accumulators and dL / dA_j^T, handed from the pair kernel to the node kernel
    uint8_t* wimg;           // tcgen05 mode: folded-weight images of both layers
    int partial_stride;
    size_t bytes;
};

BwdWorkspace carve_bwd_workspace(void* wsbase, size_t fwd_bytes, int B, int P) {
    BwdWorkspace w;
    const int sms = num_sms() > 0 ? num_sms() : 148;
    float* p = (float*)wsbase;
    size_t o = 0;
    w.ajt = p;  // shared with the forward's scratch (first region of the forward workspace)
    o = (fwd_bytes + 15) / 16 * 4;
    w.dajt = p + o;      o += (size_t)sms * kHid * pad_k(P);
    constexpr int max_layer = param_offset(1, 0) > (PMHC_NPARAM - param_offset(1, 0)) ? param_offset(1, 0) : (PMHC_NPARAM - param_offset(1, 0));
    w.partial_stride = ((kTileFloats + max_layer + 3) / 4) * 4;
    w.partial = p + o;   o += (size_t)sms * w.partial_stride;
    w.d_frames1 = p + o; o += (size_t)B * kN * 7;
    w.d_tors1 = p + o;   o += (size_t)B * kN * 14;
    w.d_feat1 = p + o;   o += (size_t)B * kN * kHid;
    o = (o + 3) & ~(size_t)3;
    w.red = p + o;       o += ((size_t)max_layer + 3) & ~(size_t)3;
    w.scale = p + o;     o += 8;
    w.wimg = reinterpret_cast<uint8_t*>(p + o); o += (2 * kFoldImageBytes + 3) / 4;
    o = (o + 3) & ~(size_t)3;
    w.acc = p + o;       o += (size_t)(B + sms) * kAccFloats;   // one slot per segment: at most B + (CTAs) of them
    o = (o + 3) & ~(size_t)3;
    w.dajt_all = p + o;  o += (size_t)B * kHid * pad_k(P);
    o = (o + 3) & ~(size_t)3;
    w.dmsum_g = p + o;   o += (size_t)B * 2 * kN * kHid;
    w.ajt_all = p + o;   o += (size_t)B * kHid * pad_k(P);
    w.rec_all = p + o;   o += (size_t)B * t5_record_floats(pad_k(P));
    o = (o + 3) & ~(size_t)3;
    w.sched_ws = reinterpret_cast<int*>(p + o); o += (size_t)B + 4 * (size_t)sms + 2 * (size_t)B + 8;
    w.bytes = o * sizeof(float);
    return w;
}

template <int LAYER, bool TC>
int launch_layer_backward(const BwdArgs& g, int n_cta, float* grad, cudaStream_t stream) {
    static PerDeviceOnce configured;
    int max_smem = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const BwdMap M = make_bwd_map(g.a.Kpad, TC);
    size_t smem = (size_t)M.total_floats * sizeof(float);
    PMHC_REQUIRE((int)smem <= max_smem, "EGNN backward needs %zu B of shared memory (P=%d), device allows %d", smem, g.a.P, max_smem);
    if (configured.needed()) {
        cudaError_t e = cudaFuncSetAttribute(egnn_layer_backward_kernel<LAYER, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        PMHC_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(backward): %s", cudaGetErrorString(e));
        configured.mark();
    }
    if (profile_enabled()) profile_mark(PROF_BWD, stream, true);
    egnn_layer_backward_kernel<LAYER, TC><<<n_cta, kBwdThreads, smem, stream>>>(g);
    if (profile_enabled()) profile_mark(PROF_BWD, stream, false);
    PMHC_CHECK_LAUNCH("egnn_layer_backward");
    constexpr int base = param_offset(LAYER, 0);
    constexpr int numel = param_offset(LAYER + 1, 0) - base;
    reduce_partials_kernel<<<(numel + 255) / 256, 256, 0, stream>>>(g.partial, g.partial_stride, n_cta, numel, grad + base);
    PMHC_CHECK_LAUNCH("reduce_partials");
    return 0;
}


static inline int num_sms_or(int d) { const int n = num_sms(); return n > 0 ? n : d; }

template <int LAYER>
int launch_layer_backward_t5(const BwdArgs& g, const BwdWorkspace& w, int n_cta, float* grad, cudaStream_t stream) {
    static PerDeviceOnce configured;
    int max_smem = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const T5Map T = make_t5_map(g.a.Kpad);
    const size_t smem = (size_t)T.total_bytes;
    PMHC_REQUIRE((int)smem <= max_smem, "EGNN backward needs %zu B of shared memory (P=%d), device allows %d", smem, g.a.P, max_smem);
    if (configured.needed()) {
        cudaError_t e = cudaFuncSetAttribute(egnn_layer_backward_t5_kernel<LAYER>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        PMHC_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(backward): %s", cudaGetErrorString(e));
        configured.mark();
    }
    constexpr int base = param_offset(LAYER, 0);
    constexpr int numel = param_offset(LAYER + 1, 0) - base;
    unsigned* max_bits = reinterpret_cast<unsigned*>(w.scale) + LAYER;      // zeroed by bwd_fold_weights_kernel at the start of the step
    bwd_grad_max_kernel<<<128, 256, 0, stream>>>(g.d_frames_out, g.a.B * kN * 7, g.d_tors_out, g.a.B * kN * 14, max_bits);
    PMHC_CHECK_LAUNCH("bwd_grad_max");
    if (LAYER == 0) {
        static PerDeviceOnce pre_configured;
        const PreMap PM = make_pre_map();
        if (pre_configured.needed()) {
            cudaError_t e = cudaFuncSetAttribute(bwd_feature_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PM.total_bytes);
            PMHC_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(node pre-kernel): %s", cudaGetErrorString(e));
            pre_configured.mark();
        }
        bwd_feature_pre_kernel<<<n_cta, kPostThreads, PM.total_bytes, stream>>>(g, w.dmsum_g);
        PMHC_CHECK_LAUNCH("bwd_feature_pre");
    }
    int4* sched = reinterpret_cast<int4*>(w.sched_ws + (((size_t)g.a.B + 3) & ~(size_t)3));
    int2* segs = reinterpret_cast<int2*>(reinterpret_cast<int*>(sched) + 4 * (size_t)num_sms_or(148));
    {
        static PerDeviceOnce setup_configured;
        const SmemMap SM = make_setup_map(g.a.Kpad, g.a.P, layer_H(LAYER));
        const size_t sbytes = (size_t)SM.total_floats * sizeof(float);
        PMHC_REQUIRE((int)sbytes <= max_smem, "EGNN backward setup needs %zu B of shared memory (P=%d), device allows %d", sbytes, g.a.P, max_smem);
        if (setup_configured.needed()) {
            cudaError_t e = cudaFuncSetAttribute(bwd_setup_pre_kernel<LAYER>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
            PMHC_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(setup pre-kernel): %s", cudaGetErrorString(e));
            setup_configured.mark();
        }
        int* units = w.sched_ws;
        bwd_setup_pre_kernel<LAYER><<<g.a.B, kSetupThreads, sbytes, stream>>>(g.a, w.ajt_all, w.rec_all, w.wimg + (size_t)LAYER * kFoldImageBytes,
                                                                               w.dajt_all, units);
        PMHC_CHECK_LAUNCH("bwd_setup_pre");
        const size_t sched_smem = (2 * (size_t)g.a.B + 2) * sizeof(int);
        PMHC_REQUIRE(sched_smem <= 200 * 1024, "EGNN backward: batch of %d complexes exceeds the schedule kernel's staging", g.a.B);
        static PerDeviceOnce sched_configured;
        if (sched_configured.needed()) {
            cudaFuncSetAttribute(bwd_schedule_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            sched_configured.mark();
        }
        bwd_schedule_kernel<<<1, 1024, sched_smem, stream>>>(units, g.a.B, n_cta, sched, segs);
        PMHC_CHECK_LAUNCH("bwd_schedule");
    }
    T5Args x{w.wimg + (size_t)LAYER * kFoldImageBytes, max_bits, w.acc, w.dajt_all, w.ajt_all, w.rec_all, w.sched_ws, sched, segs, w.dmsum_g};
    if (profile_enabled()) profile_mark(PROF_BWD, stream, true);
    egnn_layer_backward_t5_kernel<LAYER><<<n_cta, kT5Threads, smem, stream>>>(g, x);
    if (profile_enabled()) profile_mark(PROF_BWD, stream, false);
    PMHC_CHECK_LAUNCH("egnn_layer_backward_t5");
    {
        static PerDeviceOnce post_configured;
        const PostMap PM = make_post_map();
        if (post_configured.needed()) {
            cudaError_t e = cudaFuncSetAttribute(bwd_node_post_kernel<LAYER>, cudaFuncAttributeMaxDynamicSharedMemorySize, PM.total_bytes);
            PMHC_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(node kernel): %s", cudaGetErrorString(e));
            post_configured.mark();
        }
        bwd_node_post_kernel<LAYER><<<n_cta, kPostThreads, PM.total_bytes, stream>>>(g, w.acc, w.dajt_all, segs);
        PMHC_CHECK_LAUNCH("bwd_node_post");
    }
    reduce_partials_to_kernel<<<(numel + 63) / 64, dim3(64, 4), 0, stream>>>(g.partial, g.partial_stride, n_cta, numel, w.red);
    PMHC_CHECK_LAUNCH("reduce_partials_to");
    bwd_unfold_kernel<LAYER><<<(numel + 255) / 256 + 33, 256, 0, stream>>>(g.a.params, w.red, grad);
    PMHC_CHECK_LAUNCH("bwd_unfold");
    return 0;
}

size_t forward_workspace_bytes(int B, int P);

}  // namespace pmhc

using namespace pmhc;

extern "C" size_t pmhc_workspace_bytes(int B, int P) {
    return carve_bwd_workspace(nullptr, forward_workspace_bytes(B, P), B, P).bytes;
}

extern "C" int pmhc_model_backward_ex(const float* params, const PmhcBatch* bt, float t_over_T, const float* saved,
                                      const float* d_out_frames, const float* d_out_torsions, float* flat_grad,
                                      void* workspace, size_t workspace_bytes, void* stream_, void* layer2_done_event,
                                      int precision) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PMHC_REQUIRE(precision == PMHC_PRECISION_FP32 || precision == PMHC_PRECISION_BF16 || precision == PMHC_PRECISION_FP16,
                 "pmhc_model_backward: unknown precision %d", precision);
    const bool tc = precision == PMHC_PRECISION_BF16;
    const bool t5 = precision == PMHC_PRECISION_FP16;
    PMHC_REQUIRE(device_props() == 0, "no CUDA device");
    PMHC_REQUIRE(bt != nullptr && bt->B > 0, "pmhc_model_backward: empty batch");
    PMHC_REQUIRE(bt->P >= 1 && bt->P <= kMaxP, "pmhc_model_backward: pocket_maxlen %d outside [1, %d]", bt->P, kMaxP);
    PMHC_REQUIRE(saved != nullptr, "pmhc_model_backward: the forward must have been run with a `saved` buffer");
    const size_t fwd_bytes = forward_workspace_bytes(bt->B, bt->P);
    BwdWorkspace w = carve_bwd_workspace(workspace, fwd_bytes, bt->B, bt->P);
    PMHC_REQUIRE(workspace != nullptr && workspace_bytes >= w.bytes, "pmhc_model_backward: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    SavedMap sv = carve_saved(const_cast<float*>(saved), bt->B, bt->P);
    // (the tcgen05 mode splits complexes over CTAs by passes: every SM works, whatever the batch size)
    // (... and when the caller overlaps a collective with the layer-1 launch (layer2_done_event), eight SMs are left to it: a persistent
    // CTA that cannot become resident next to the collective's CTAs would otherwise run its whole share after everybody else)
    const int n_cta = t5 ? num_sms() - (layer2_done_event != nullptr && num_sms() > 16 ? 8 : 0) : (bt->B < num_sms() ? bt->B : num_sms());

    BwdArgs g{};
    g.a.params = params;
    g.a.B = bt->B; g.a.P = bt->P; g.a.Kpad = pad_k(bt->P);
    g.a.t_over_T = t_over_T;
    g.a.t_dev = step_t_dev();
    g.a.mask = bt->mask;
    g.a.pocket_frames = bt->pocket_frames; g.a.pocket_feat = bt->pocket_features; g.a.pocket_mask = bt->pocket_mask;
    g.a.ajt_ws = w.ajt;
    g.partial = w.partial; g.partial_stride = w.partial_stride; g.dajt_ws = w.dajt;

    // layer 2 (its inputs are layer 1's outputs)
    g.a.frames_in = sv.frames1; g.a.tors_in = sv.tors1; g.a.feat_in = sv.feat1;
    g.rowstat = sv.rowstat2; g.logits = sv.logits2; g.msum = nullptr; g.feat_post = nullptr;
    g.d_frames_out = d_out_frames; g.d_tors_out = d_out_torsions; g.d_feat_out = nullptr;
    g.d_frames_in = w.d_frames1; g.d_tors_in = w.d_tors1; g.d_feat_in = w.d_feat1;
    if (t5) {
        bwd_fold_weights_kernel<<<64, 256, 0, stream>>>(params, w.wimg, reinterpret_cast<unsigned*>(w.scale));
        PMHC_CHECK_LAUNCH("bwd_fold_weights");
    }
    int rc = t5 ? launch_layer_backward_t5<1>(g, w, n_cta, flat_grad, stream)
                : tc ? launch_layer_backward<1, true>(g, n_cta, flat_grad, stream) : launch_layer_backward<1, false>(g, n_cta, flat_grad, stream);
    if (rc != 0) return rc;
    if (layer2_done_event != nullptr) cudaEventRecord((cudaEvent_t)layer2_done_event, stream);
    // layer 1
    g.a.frames_in = bt->frames; g.a.tors_in = bt->torsions; g.a.feat_in = bt->features;
    g.rowstat = sv.rowstat1; g.logits = sv.logits1; g.msum = sv.msum1; g.feat_post = sv.feat1;
    g.d_frames_out = w.d_frames1; g.d_tors_out = w.d_tors1; g.d_feat_out = w.d_feat1;
    g.d_frames_in = nullptr; g.d_tors_in = nullptr; g.d_feat_in = nullptr;
    if (t5) return launch_layer_backward_t5<0>(g, w, n_cta, flat_grad, stream);
    return tc ? launch_layer_backward<0, true>(g, n_cta, flat_grad, stream) : launch_layer_backward<0, false>(g, n_cta, flat_grad, stream);
}

extern "C" int pmhc_model_backward(const float* params, const PmhcBatch* bt, float t_over_T, const float* saved,
                                   const float* d_out_frames, const float* d_out_torsions, float* flat_grad,
                                   void* workspace, size_t workspace_bytes, void* stream_, void* layer2_done_event) {
    return pmhc_model_backward_ex(params, bt, t_over_T, saved, d_out_frames, d_out_torsions, flat_grad, workspace, workspace_bytes,
                                  stream_, layer2_done_event, PMHC_PRECISION_FP32);
}
