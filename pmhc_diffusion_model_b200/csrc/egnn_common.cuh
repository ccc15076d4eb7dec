// egnn_common.cuh — shared-memory layout, weight staging and per-complex setup used by the fused
// EGNN layer forward and backward kernels.
//
// Factorisation used everywhere (derived from the reference's definitions, model.py:183-333):
//   message_mlp.0(cat(h_i, h_j, e_ij)) = A_i[i] + A_j[j] + W_e[:, rel(i,j)] * [j is peptide]   (bias folded in A_i)
//   attention_mlp.0(cat(m, -d2, qdot2)) = W_m m - w_d d2 + w_q qdot2 + b
//   rotation_mlp.0(cat(m, local_q))     = W_m m + W_q local_q + b
//   torsion_mlp.0(cat(m, tors_i))       = W_m m + T_t[i]          (T_t = W_t tors_i + b, per node)
//   translation_mlp.0(m)                = W_m m + b
// so per pair the dense work is 64x64 (message layer 2) and 64x256 (the four head hidden layers share m).
#pragma once

#include "common.cuh"
#include "pmhc_math.cuh"

namespace pmhc {

constexpr int kFwdThreads = 512;     // forward: 16 warps = two threads per pair column (each owns half of every 64-wide result), one CTA per SM
constexpr int kBwdThreads = 256;     // backward: 8 warps = two threads per pair column (each owns half of every 64-wide result)
constexpr int kBwdPairs = 128;       //   of a 128-pair pass: two [64][128] column tiles are all that fits next to the weights
constexpr int kPassPairs = 256;      // pairs per forward pass (two threads each)
constexpr int kScrLd = kPassPairs + 1;  // odd row stride -> conflict-free column access
constexpr int kCapPairs = 512;       // pairs whose head outputs are buffered before the row softmax
constexpr int kOutPerPair = 15;      // logit, global delta quat (4), delta angles (7), scale * (x_i - x_j) (3)
constexpr int kLdN = kHid + 1;       // padded stride of per-node rows (A_i, W_e) -> conflict-free
constexpr int kMaxP = 480;           // largest pocket_maxlen the shared-memory budget allows

struct LayerArgs {
    const float* params;         // flat parameter buffer
    int B, P, Kpad;              // complexes, pocket slots, padded neighbour count (multiple of 32)
    float t_over_T;
    const float* t_dev;          // nullable: the same scalar in device memory (read instead of t_over_T; graph-replayable steps)
    const float* frames_in;      // [B,16,7]
    const float* tors_in;        // [B,16,14]
    const float* feat_in;        // layer 1: [B,16,22] features; layer 2: [B,16,64] relu(o1)
    const uint8_t* mask;         // [B,16]
    const float* pocket_frames;  // [B,P,7]
    const float* pocket_feat;    // [B,P,22]
    const uint8_t* pocket_mask;  // [B,P]
    float* frames_out;           // [B,16,7]
    float* tors_out;             // [B,16,14]
    float* feat_out;             // layer 1 only: [B,16,64] relu(o1)
    float* msum_out;             // layer 1, training only: [B,16,64] unmasked message sums
    float* rowstat;              // training only: [B,16,16] lse, G(4), dA(7), Xa(3), pad
    float* logit_out;            // training only: [B,16,Kpad] attention logits by neighbour slot
    float* ajt_ws;               // [gridDim.x][64][Kpad] per-CTA scratch for the neighbour projections A_j^T
    // step-invariant pocket data precomputed by pocket_projection_kernel (tensor-core path; null otherwise):
    float* ajt_cache;            // [B][2][64][Kpad] A_j^T per complex and layer; pocket columns 16.. are pre-filled
    const uint8_t* pocket_cls;   // [B][P] 0 = valid, 1 = masked with all-zero features, 2 = masked with non-zero features
};

// pointers into the caller's `saved` buffer (pmhc_saved_floats)
struct SavedMap {
    float *rowstat1, *rowstat2, *frames1, *tors1, *feat1, *msum1, *logits1, *logits2;
};
SavedMap carve_saved(float* saved, int B, int P);
int pad_k(int P);
int device_props();
int num_sms();

// ---- shared memory map (float offsets); K-dependent arrays last ----
struct SmemMap {
    int W2T, WhT, We, PkAtt, PkRotQ, PkRot2, PkMisc, PkTor2, Scal;
    int Scr, Out, Ai, Tt, Msum, H, Tors, Q, X, Ints, total_floats;
};

__host__ __device__ inline SmemMap make_smem_map(int Kpad) {
    SmemMap m;
    int o = 0;
    m.W2T = o;  o += kHid * kHid;
    m.WhT = o;  o += kHid * 4 * kHid;
    m.We = o;   o += kEdge * kLdN + 1;   // +1 keeps the next array 16-byte aligned (31*65 = 2015)
    // per-hidden-unit parameter packs, one float4 (or two) per unit n so the head epilogues need one 128-bit
    // shared load per unit: PkAtt = {w_d, w_q, b_att, att2}, PkRotQ = W_q[n, 0:4], PkRot2 = rot2[0:4, n],
    // PkMisc = {b_trn, trn2, b_rot, b2}, PkTor2 = {tor2[0:7, n], 0}
    m.PkAtt = o;  o += 4 * kHid;
    m.PkRotQ = o; o += 4 * kHid;
    m.PkRot2 = o; o += 4 * kHid;
    m.PkMisc = o; o += 4 * kHid;
    m.PkTor2 = o; o += 8 * kHid;
    m.Scal = o;   o += 16;
    m.Scr = o;  o += kHid * kScrLd + 3;  // 64*257 = 16448 (+3 -> multiple of 4... keep alignment below)
    o = (o + 3) & ~3;
    m.Out = o;  o += kCapPairs * kOutPerPair;
    m.Ai = o;   o += kN * kLdN;
    o = (o + 3) & ~3;
    m.Tt = o;   o += kN * kHid;
    m.Msum = o; o += kN * kHid;
    m.H = o;    o += kN * kLdN;
    o = (o + 3) & ~3;
    m.Tors = o; o += kN * 2 * PMHC_NTORS;
    m.Q = o;    o += Kpad * 4;
    m.X = o;    o += Kpad * 3;
    o = (o + 3) & ~3;
    m.Ints = o; o += Kpad + 64;          // neighbour lists and row lists (int32)
    m.total_floats = o;
    return m;
}

// scalar slots inside Scal
enum { SC_ATT2B = 0, SC_TRN2B = 1, SC_ROT2B = 2 /*..5*/, SC_TOR2B = 6 /*..12*/ };

// int slots (relative to Ints): [0,16) real rows; [16,32) masked peptide slots; [32, 32+Kpad) pocket lists:
// valid pocket slots packed from the front, masked-with-nonzero-features packed from the back.
enum { IN_ROWS = 0, IN_PEPX = 16, IN_POCKET = 32 };

struct ComplexInfo {
    int L;    // real peptide rows
    int nv;   // valid pocket slots
    int nx;   // masked pocket slots with non-zero features (need their own message in layer 1)
    int c0;   // masked pocket slots with all-zero features (one shared message, multiplicity c0)
};

// Copy one layer's weights into shared memory in the k-major layouts the pair loops read.
template <int LAYER>
__device__ inline void stage_layer_weights(float* S, const SmemMap& M, const float* __restrict__ params) {
    constexpr int L = LAYER;
    constexpr int H = layer_H(L);
    const int tid = threadIdx.x;
    const float* msg0 = params + param_offset(L, MSG0_W);
    const float* msg2 = params + param_offset(L, MSG2_W);
    const float* att0 = params + param_offset(L, ATT0_W);
    const float* rot0 = params + param_offset(L, ROT0_W);
    const float* tor0 = params + param_offset(L, TOR0_W);
    const float* trn0 = params + param_offset(L, TRN0_W);
    constexpr int ld1 = 2 * H + kEdge;
    // global reads are coalesced along the weight rows (k fastest); shared writes are transposed
    for (int idx = tid; idx < kHid * kHid; idx += blockDim.x) {
        int n = idx >> 6, k = idx & 63;
        S[M.W2T + k * kHid + n] = msg2[idx];
        S[M.WhT + k * 256 + 192 + n] = trn0[idx];
    }
    for (int idx = tid; idx < kHid * 66; idx += blockDim.x) {
        int n = idx / 66, k = idx - n * 66;
        float v = att0[idx];
        if (k < 64) S[M.WhT + k * 256 + n] = v;
        else S[M.PkAtt + 4 * n + (k - 64)] = v;       // w_d, w_q
    }
    for (int idx = tid; idx < kHid * 68; idx += blockDim.x) {
        int n = idx / 68, k = idx - n * 68;
        float v = rot0[idx];
        if (k < 64) S[M.WhT + k * 256 + 64 + n] = v;
        else S[M.PkRotQ + 4 * n + (k - 64)] = v;
    }
    for (int idx = tid; idx < kHid * 78; idx += blockDim.x) {
        int n = idx / 78, k = idx - n * 78;
        if (k < 64) S[M.WhT + k * 256 + 128 + n] = tor0[idx];
    }
    for (int idx = tid; idx < kHid * kEdge; idx += blockDim.x) {
        int k = idx / kEdge, r = idx - k * kEdge;
        S[M.We + r * kLdN + k] = msg0[k * ld1 + 2 * H + r];
    }
    for (int n = tid; n < kHid; n += blockDim.x) {
        S[M.PkAtt + 4 * n + 2] = params[param_offset(L, ATT0_B) + n];
        S[M.PkAtt + 4 * n + 3] = params[param_offset(L, ATT2_W) + n];
        S[M.PkMisc + 4 * n + 0] = params[param_offset(L, TRN0_B) + n];
        S[M.PkMisc + 4 * n + 1] = params[param_offset(L, TRN2_W) + n];
        S[M.PkMisc + 4 * n + 2] = params[param_offset(L, ROT0_B) + n];
        S[M.PkMisc + 4 * n + 3] = params[param_offset(L, MSG2_B) + n];
        S[M.PkTor2 + 8 * n + 7] = 0.0f;
    }
    for (int idx = tid; idx < 4 * kHid; idx += blockDim.x) {
        int c = idx >> 6, n = idx & 63;
        S[M.PkRot2 + 4 * n + c] = params[param_offset(L, ROT2_W) + idx];
    }
    for (int idx = tid; idx < PMHC_NTORS * kHid; idx += blockDim.x) {
        int c = idx >> 6, n = idx & 63;
        S[M.PkTor2 + 8 * n + c] = params[param_offset(L, TOR2_W) + idx];
    }
    if (tid == 0) {
        S[M.Scal + SC_ATT2B] = params[param_offset(L, ATT2_B)];
        S[M.Scal + SC_TRN2B] = params[param_offset(L, TRN2_B)];
        for (int c = 0; c < 4; ++c) S[M.Scal + SC_ROT2B + c] = params[param_offset(L, ROT2_B) + c];
        for (int c = 0; c < PMHC_NTORS; ++c) S[M.Scal + SC_TOR2B + c] = params[param_offset(L, TOR2_B) + c];
    }
}

// Per-complex setup: frames, torsions, node features, row / neighbour lists, and the per-node projections
// A_i (shared), A_j^T (global scratch, [64][Kpad]) and T_t (shared).  Ends with a __syncthreads().
template <int LAYER>
__device__ inline ComplexInfo setup_complex(float* S, const SmemMap& M, const LayerArgs& a, int b, float* ajt) {
    constexpr int L = LAYER;
    constexpr int H = layer_H(L);
    constexpr int ld1 = 2 * H + kEdge;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.P, K = kN + P, Kpad = a.Kpad;
    int* I = reinterpret_cast<int*>(S + M.Ints);
    float* feat_stage = S + M.Scr;  // pocket features staged here: [P][23] (scr is free during setup)
    constexpr int FS = 23;

    const float* msg0 = a.params + param_offset(L, MSG0_W);
    const float* msg0b = a.params + param_offset(L, MSG0_B);
    const float* tor0 = a.params + param_offset(L, TOR0_W);
    const float* tor0b = a.params + param_offset(L, TOR0_B);
    // Temporary staging area (the pass buffers are free during setup).  Phase 1: pocket features [P][23] and the
    // pocket block of message_mlp.0 (W1[:, H:H+22]) as [64][23].  Phase 2 (overwrites phase 1): W1[:, 0:2H] as
    // [64][2H+1] and torsion_mlp.0[:, 64:78] + bias as [64][15].  All weight reads of the projections below hit
    // shared memory: with one warp per scheduler, L2-latency loads in these loops cost ~25 % of the kernel.
    float* wp = feat_stage + ((P * FS + 3) & ~3);
    constexpr int LDQ = 2 * H + 1;
    float* wq = feat_stage;
    float* torx = wq + kHid * LDQ;

    // -- geometry of all K slots --
    for (int idx = tid; idx < K * 7; idx += blockDim.x) {
        int j = idx / 7, c = idx - j * 7;
        float v = (j < kN) ? a.frames_in[((size_t)b * kN + j) * 7 + c] : a.pocket_frames[((size_t)b * P + (j - kN)) * 7 + c];
        if (c < 4) S[M.Q + j * 4 + c] = v;
        else S[M.X + j * 3 + (c - 4)] = v;
    }
    for (int idx = tid; idx < kN * 14; idx += blockDim.x) S[M.Tors + idx] = a.tors_in[(size_t)b * kN * 14 + idx];
    // -- node features: layer 1 = (22 features, t/T); layer 2 = 64 learned features --
    for (int idx = tid; idx < kN * kHid; idx += blockDim.x) {
        int i = idx >> 6, c = idx & 63;
        float v;
        if (L == 0) v = (c < PMHC_NFEAT) ? a.feat_in[((size_t)b * kN + i) * PMHC_NFEAT + c] : (c == PMHC_NFEAT ? time_feature(a) : 0.0f);
        else v = a.feat_in[((size_t)b * kN + i) * kHid + c];
        S[M.H + i * kLdN + c] = v;
    }
#pragma unroll 8
    for (int idx = tid; idx < P * PMHC_NFEAT; idx += blockDim.x) {
        int j = idx / PMHC_NFEAT, c = idx - j * PMHC_NFEAT;
        feat_stage[j * FS + c] = __ldg(a.pocket_feat + (size_t)b * P * PMHC_NFEAT + idx);
    }
#pragma unroll 8
    for (int idx = tid; idx < kHid * PMHC_NFEAT; idx += blockDim.x) {
        int k = idx / PMHC_NFEAT, c = idx - k * PMHC_NFEAT;
        wp[k * FS + c] = __ldg(msg0 + k * ld1 + H + c);
    }
    __syncthreads();

    // -- lists (warp 0): real rows, masked peptide slots, valid pocket slots, masked non-zero pocket slots --
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
            bool in = j < P;
            bool valid = in && a.pocket_mask[(size_t)b * P + j] != 0;
            bool nonzero = false;
            if (in && !valid) {
                for (int c = 0; c < PMHC_NFEAT; ++c) nonzero |= (feat_stage[j * FS + c] != 0.0f);
            }
            unsigned bv = __ballot_sync(0xffffffffu, valid);
            unsigned bx = __ballot_sync(0xffffffffu, in && !valid && nonzero);
            unsigned bz = __ballot_sync(0xffffffffu, in && !valid && !nonzero);
            if (valid) I[IN_POCKET + nv + __popc(bv & ((1u << lane) - 1u))] = kN + j;
            if (in && !valid && nonzero) I[IN_POCKET + Kpad - 1 - (nx + __popc(bx & ((1u << lane) - 1u)))] = kN + j;
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

    // -- pocket part of A_j^T[k][16 + p] = W1[k, H:H+22] . pocket_feat[p]  (4 k per thread item; time / padding columns are 0) --
    for (int idx = tid; idx < (kHid / 4) * (Kpad - kN); idx += blockDim.x) {
        int kq = idx / (Kpad - kN), pj = idx - kq * (Kpad - kN);
        float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
        if (pj < P) {
            const float* h = feat_stage + pj * FS;
            const float* w = wp + (4 * kq) * FS;
#pragma unroll
            for (int c = 0; c < PMHC_NFEAT; ++c) {
                float hv = h[c];
                acc0 = fmaf(w[c], hv, acc0);
                acc1 = fmaf(w[FS + c], hv, acc1);
                acc2 = fmaf(w[2 * FS + c], hv, acc2);
                acc3 = fmaf(w[3 * FS + c], hv, acc3);
            }
        }
        ajt[(4 * kq + 0) * Kpad + kN + pj] = acc0;
        ajt[(4 * kq + 1) * Kpad + kN + pj] = acc1;
        ajt[(4 * kq + 2) * Kpad + kN + pj] = acc2;
        ajt[(4 * kq + 3) * Kpad + kN + pj] = acc3;
    }
    __syncthreads();
    // -- phase 2 staging --
#pragma unroll 8
    for (int idx = tid; idx < kHid * 2 * H; idx += blockDim.x) {
        int k = idx / (2 * H), c = idx - k * (2 * H);
        wq[k * LDQ + c] = __ldg(msg0 + k * ld1 + c);
    }
#pragma unroll 4
    for (int idx = tid; idx < kHid * 14; idx += blockDim.x) {
        int n = idx / 14, c = idx - n * 14;
        torx[n * 15 + c] = __ldg(tor0 + n * 78 + 64 + c);
    }
    for (int n = tid; n < kHid; n += blockDim.x) torx[n * 15 + 14] = tor0b[n];
    __syncthreads();
    // -- peptide slots: A_i[i][k] = b1[k] + W1[k, 0:H] h_i (bias folded in) and A_j^T[k][i] = W1[k, H:2H] h_i --
    for (int idx = tid; idx < kN * kHid; idx += blockDim.x) {
        int k = idx >> 4, i = idx & 15;
        const float* w = wq + k * LDQ;
        const float* h = S + M.H + i * kLdN;
        float ai = msg0b[k], aj = 0.0f;
#pragma unroll 8
        for (int c = 0; c < H; ++c) {
            float hv = h[c];
            ai = fmaf(w[c], hv, ai);
            aj = fmaf(w[H + c], hv, aj);
        }
        S[M.Ai + i * kLdN + k] = ai;
        ajt[k * Kpad + i] = aj;
    }
    // T_t[i][n] = b[n] + W_t[n, 64:78] tors_i
    for (int idx = tid; idx < kN * kHid; idx += blockDim.x) {
        int i = idx >> 6, n = idx & 63;
        const float* w = torx + n * 15;
        const float* t = S + M.Tors + i * 14;
        float acc = w[14];
#pragma unroll
        for (int c = 0; c < 14; ++c) acc = fmaf(w[c], t[c], acc);
        S[M.Tt + idx] = acc;
    }
    for (int idx = tid; idx < kN * kHid; idx += blockDim.x) S[M.Msum + idx] = 0.0f;
    __syncthreads();
    ComplexInfo ci;
    ci.L = I[IN_POCKET + Kpad + 0];
    ci.nv = I[IN_POCKET + Kpad + 1];
    ci.nx = I[IN_POCKET + Kpad + 2];
    ci.c0 = I[IN_POCKET + Kpad + 3];
    return ci;
}

// acc[u][n] += sum_k WT[k*ldw + n] * scr[k*kScrLd + col[u]]  for k in [0,64), n in [0,64).
// Weight reads are warp-uniform 128-bit loads (one shared-memory wavefront feeds 4*PPT FMAs per lane).
template <int PPT>
__device__ __forceinline__ void gemv64(float (&acc)[PPT][kHid], const float* __restrict__ WT, int ldw,
                                       const float* __restrict__ scr, const int (&col)[PPT]) {
#pragma unroll 2
    for (int k = 0; k < kHid; ++k) {
        float av[PPT];
#pragma unroll
        for (int u = 0; u < PPT; ++u) av[u] = scr[k * kScrLd + col[u]];
        const float4* w4 = reinterpret_cast<const float4*>(WT + k * ldw);
#pragma unroll
        for (int n4 = 0; n4 < kHid / 4; ++n4) {
            float4 w = w4[n4];
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                acc[u][4 * n4 + 0] = fmaf(w.x, av[u], acc[u][4 * n4 + 0]);
                acc[u][4 * n4 + 1] = fmaf(w.y, av[u], acc[u][4 * n4 + 1]);
                acc[u][4 * n4 + 2] = fmaf(w.z, av[u], acc[u][4 * n4 + 2]);
                acc[u][4 * n4 + 3] = fmaf(w.w, av[u], acc[u][4 * n4 + 3]);
            }
        }
    }
}

// Decoded pair of one thread slot.
struct PairRef {
    int i;       // peptide row (slot index)
    int j;       // neighbour slot: < 16 peptide, >= 16 pocket (16 + pocket slot); -1 = shared zero-feature pocket message
    bool active;
};

// Full (attention-carrying) pairs of row i: the L-1 other real peptide residues, then the nv valid pocket slots.
__device__ __forceinline__ PairRef decode_full_pair(const int* I, int gp, int W, int L, int row0, bool active) {
    PairRef p;
    int rl = gp / W, e = gp - rl * W;
    int r = row0 + rl;
    p.i = I[IN_ROWS + r];
    if (e < L - 1) p.j = I[IN_ROWS + (e < r ? e : e + 1)];
    else p.j = I[IN_POCKET + (e - (L - 1))];
    p.active = active;
    return p;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}


// Per-row masked softmax over the buffered head outputs and the attention-weighted frame / torsion / translation
// updates (model.py:243, 263-269, 300-310, 331; output quaternion normalised, model.py:181).  One warp per row.
__device__ inline void finalize_rows(float* S, const SmemMap& M, const LayerArgs& a, const int* I, int b, int row0,
                                     int nrows, int W) {
    // one HALF-warp (16 lanes) per row: 16 rows per round with 8 warps, so a 9..16-mer is one round
    const int lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int base = 0; base < nrows; base += 2 * nwarps) {
        const int rl = base + 2 * warp + half;
        const bool live = rl < nrows;
        const int i = I[IN_ROWS + row0 + (live ? rl : 0)];
        const float* out = S + M.Out + (size_t)(live ? rl : 0) * W * kOutPerPair;
        const int Wl = live ? W : 0;
        float mx = -INFINITY;
        for (int e = l16; e < Wl; e += 16) mx = fmaxf(mx, out[e * kOutPerPair]);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.0f, ws[14];
#pragma unroll
        for (int c = 0; c < 14; ++c) ws[c] = 0.0f;
        for (int e = l16; e < Wl; e += 16) {
            const float* o = out + e * kOutPerPair;
            float p = expf(o[0] - mx);
            se += p;
#pragma unroll
            for (int c = 0; c < 14; ++c) ws[c] = fmaf(p, o[1 + c], ws[c]);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
#pragma unroll
            for (int c = 0; c < 14; ++c) ws[c] += __shfl_xor_sync(0xffffffffu, ws[c], o);
        }
        if (!live) continue;
        const float inv = W > 0 ? 1.0f / se : 0.0f;
#pragma unroll
        for (int c = 0; c < 14; ++c) ws[c] *= inv;
        const size_t node = (size_t)b * kN + i;
        if (l16 == 0) {
            const float* qi = S + M.Q + i * 4;
            const float* xi = S + M.X + i * 3;
            Quat G{ws[0], ws[1], ws[2], ws[3]};
            Quat g = W > 0 ? qnormalize(G) : Quat{1.0f, 0.0f, 0.0f, 0.0f};  // model.py:301-306
            Quat q = qunit(qmul(g, Quat{qi[0], qi[1], qi[2], qi[3]}));       // model.py:310, :181
            float* fo = a.frames_out + node * 7;
            fo[0] = q.w; fo[1] = q.x; fo[2] = q.y; fo[3] = q.z;
            fo[4] = xi[0] + ws[11]; fo[5] = xi[1] + ws[12]; fo[6] = xi[2] + ws[13];
            if (a.rowstat != nullptr) {
                float* rs = a.rowstat + node * PMHC_ROWSTAT;
                rs[0] = W > 0 ? mx + logf(se) : 0.0f;
#pragma unroll
                for (int c = 0; c < 14; ++c) rs[1 + c] = ws[c];
                rs[15] = 0.0f;
            }
        }
        if (l16 >= 1 && l16 <= PMHC_NTORS) {
            // torsions' = (sin dA, cos dA) (x) torsions (model.py:263-269)
            const int tc_ = l16 - 1;
            float da = 0.0f;
#pragma unroll
            for (int c = 0; c < PMHC_NTORS; ++c) da = (tc_ == c) ? ws[4 + c] : da;
            float sn, cs;
            sincosf(da, &sn, &cs);
            const float* t = S + M.Tors + i * 14 + 2 * tc_;
            SinCos o = scmul(SinCos{sn, cs}, SinCos{t[0], t[1]});
            a.tors_out[node * 14 + 2 * tc_] = o.s;
            a.tors_out[node * 14 + 2 * tc_ + 1] = o.c;
        }
    }
}

// Per-complex setup when the pocket projections are cached (pocket_projection_kernel): only the peptide side is
// recomputed — geometry, torsions, node features, lists from the cached slot classes, A_i / A_j for the 16 peptide
// slots from the RESIDENT first-layer weights (wq: [64][2H+1], torx: [64][15]) and T_t.  `ajt` is this complex's
// [64][Kpad] block of the cache; its peptide columns 0..15 are rewritten here.  Ends with a __syncthreads().
template <int LAYER>
__device__ inline ComplexInfo setup_complex_cached(float* S, const SmemMap& M, const float* wq, const float* torx,
                                                   const LayerArgs& a, int b, float* ajt) {
    constexpr int L = LAYER;
    constexpr int H = layer_H(L);
    constexpr int LDQ = 2 * H + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = a.P, K = kN + P, Kpad = a.Kpad;
    int* I = reinterpret_cast<int*>(S + M.Ints);
    const float* msg0b = a.params + param_offset(L, MSG0_B);

    for (int idx = tid; idx < K * 7; idx += blockDim.x) {
        int j = idx / 7, c = idx - j * 7;
        float v = (j < kN) ? a.frames_in[((size_t)b * kN + j) * 7 + c] : a.pocket_frames[((size_t)b * P + (j - kN)) * 7 + c];
        if (c < 4) S[M.Q + j * 4 + c] = v;
        else S[M.X + j * 3 + (c - 4)] = v;
    }
    for (int idx = tid; idx < kN * 14; idx += blockDim.x) S[M.Tors + idx] = a.tors_in[(size_t)b * kN * 14 + idx];
    for (int idx = tid; idx < kN * kHid; idx += blockDim.x) {
        int i = idx >> 6, c = idx & 63;
        float v;
        if (L == 0) v = (c < PMHC_NFEAT) ? a.feat_in[((size_t)b * kN + i) * PMHC_NFEAT + c] : (c == PMHC_NFEAT ? time_feature(a) : 0.0f);
        else v = a.feat_in[((size_t)b * kN + i) * kHid + c];
        S[M.H + i * kLdN + c] = v;
    }
    for (int idx = tid; idx < kN * kHid; idx += blockDim.x) S[M.Msum + idx] = 0.0f;
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
    for (int idx = tid; idx < kN * kHid; idx += blockDim.x) {
        int k = idx >> 4, i = idx & 15;
        const float* w = wq + k * LDQ;
        const float* h = S + M.H + i * kLdN;
        float ai = msg0b[k], aj = 0.0f;
#pragma unroll 8
        for (int c = 0; c < H; ++c) {
            float hv = h[c];
            ai = fmaf(w[c], hv, ai);
            aj = fmaf(w[H + c], hv, aj);
        }
        S[M.Ai + i * kLdN + k] = ai;
        ajt[k * Kpad + i] = aj;
    }
    for (int idx = tid; idx < kN * kHid; idx += blockDim.x) {
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

// The per-hidden-unit parameter packs and scalars only (see SmemMap), for kernels that stage the big matrices
// in their own format.
template <int LAYER>
__device__ inline void stage_packs(float* S, const SmemMap& M, const float* __restrict__ params) {
    constexpr int L = LAYER;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const float* att0 = params + param_offset(L, ATT0_W);
    const float* rot0 = params + param_offset(L, ROT0_W);
    for (int n = tid; n < kHid; n += nthr) {
        S[M.PkAtt + 4 * n + 0] = att0[n * 66 + 64];
        S[M.PkAtt + 4 * n + 1] = att0[n * 66 + 65];
        S[M.PkAtt + 4 * n + 2] = params[param_offset(L, ATT0_B) + n];
        S[M.PkAtt + 4 * n + 3] = params[param_offset(L, ATT2_W) + n];
        for (int c = 0; c < 4; ++c) S[M.PkRotQ + 4 * n + c] = rot0[n * 68 + 64 + c];
        for (int c = 0; c < 4; ++c) S[M.PkRot2 + 4 * n + c] = params[param_offset(L, ROT2_W) + c * kHid + n];
        S[M.PkMisc + 4 * n + 0] = params[param_offset(L, TRN0_B) + n];
        S[M.PkMisc + 4 * n + 1] = params[param_offset(L, TRN2_W) + n];
        S[M.PkMisc + 4 * n + 2] = params[param_offset(L, ROT0_B) + n];
        S[M.PkMisc + 4 * n + 3] = params[param_offset(L, MSG2_B) + n];
        for (int c = 0; c < PMHC_NTORS; ++c) S[M.PkTor2 + 8 * n + c] = params[param_offset(L, TOR2_W) + c * kHid + n];
        S[M.PkTor2 + 8 * n + 7] = 0.0f;
    }
    if (tid == 0) {
        S[M.Scal + SC_ATT2B] = params[param_offset(L, ATT2_B)];
        S[M.Scal + SC_TRN2B] = params[param_offset(L, TRN2_B)];
        for (int c = 0; c < 4; ++c) S[M.Scal + SC_ROT2B + c] = params[param_offset(L, ROT2_B) + c];
        for (int c = 0; c < PMHC_NTORS; ++c) S[M.Scal + SC_TOR2B + c] = params[param_offset(L, TOR2_B) + c];
    }
}

}  // namespace pmhc
