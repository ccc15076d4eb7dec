// structure_ops.cu — the two per-residue geometry kernels either side of the denoising path:
//
//   frames4x4_to_tensor7_kernel   loader side: backbone_rigid_tensor (4x4 homogeneous matrices, README.md:24, :30) ->
//                                 tensor_7 rows (quaternion w-first + translation), what data.py:107, :115 get from
//                                 Rigid.from_tensor_4x4(...).to_tensor_7() one entry at a time on the CPU (eigh per call)
//   atom14_kernel                 writer side: sampled frames + torsions -> peptide heavy-atom coordinates, the arithmetic
//                                 of tools/pdb.py:67-174 (torsion_angles_to_frames, frames_and_literature_positions_to_
//                                 atom14_pos, backbone O from the neighbouring N, terminal O / OXT from the psi frame)
//
// Both are HBM-bound elementwise kernels: one thread per residue, every input read once, every output written once.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "pmhc_math.cuh"

namespace pmhc {

struct Mat3 {
    float m[3][3];
};
struct Frame {
    Mat3 r;
    float t[3];
};

__device__ __forceinline__ Mat3 matmul3(const Mat3& a, const Mat3& b) {
    Mat3 c;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return c;
}
__device__ __forceinline__ void apply3(const Frame& f, const float* p, float* o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = f.r.m[i][0] * p[0] + f.r.m[i][1] * p[1] + f.r.m[i][2] * p[2] + f.t[i];
}
// Rigid.compose (RU:870-885): rotation product, translation rotated then added
__device__ __forceinline__ Frame compose(const Frame& a, const Frame& b) {
    Frame c;
    c.r = matmul3(a.r, b.r);
    apply3(a, b.t, c.t);
    return c;
}
// quat_to_rot (RU:161-181): the quadratic form of a (not necessarily unit) quaternion
__device__ __forceinline__ Mat3 quat_to_rot(const Quat& q) {
    const float ww = q.w * q.w, xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
    const float wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z, xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z;
    Mat3 r;
    r.m[0][0] = ww + xx - yy - zz; r.m[0][1] = 2.0f * (xy - wz);   r.m[0][2] = 2.0f * (xz + wy);
    r.m[1][0] = 2.0f * (xy + wz);  r.m[1][1] = ww - xx + yy - zz;  r.m[1][2] = 2.0f * (yz - wx);
    r.m[2][0] = 2.0f * (xz - wy);  r.m[2][1] = 2.0f * (yz + wx);   r.m[2][2] = ww - xx - yy + zz;
    return r;
}

// Rotation matrix -> unit quaternion.  The reference takes the top eigenvector of the 4x4 K matrix (rot_to_quat,
// RU:184-216), whose sign is arbitrary; for a proper rotation that eigenvector is the quaternion below
// (largest-component branch, numerically stable), here with the sign fixed to w >= 0.
__global__ void frames4x4_to_tensor7_kernel(const float* __restrict__ m, int64_t n, float* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float4* src = reinterpret_cast<const float4*>(m + r * 16);
    const float4 r0 = __ldg(src), r1 = __ldg(src + 1), r2 = __ldg(src + 2);
    const float m00 = r0.x, m01 = r0.y, m02 = r0.z, m10 = r1.x, m11 = r1.y, m12 = r1.z, m20 = r2.x, m21 = r2.y, m22 = r2.z;
    const float tr = m00 + m11 + m22;
    float w, x, y, z;
    if (tr > 0.0f) {
        const float s = sqrtf(tr + 1.0f) * 2.0f;
        w = 0.25f * s; x = (m21 - m12) / s; y = (m02 - m20) / s; z = (m10 - m01) / s;
    } else if (m00 > m11 && m00 > m22) {
        const float s = sqrtf(1.0f + m00 - m11 - m22) * 2.0f;
        w = (m21 - m12) / s; x = 0.25f * s; y = (m01 + m10) / s; z = (m02 + m20) / s;
    } else if (m11 > m22) {
        const float s = sqrtf(1.0f + m11 - m00 - m22) * 2.0f;
        w = (m02 - m20) / s; x = (m01 + m10) / s; y = 0.25f * s; z = (m12 + m21) / s;
    } else {
        const float s = sqrtf(1.0f + m22 - m00 - m11) * 2.0f;
        w = (m10 - m01) / s; x = (m02 + m20) / s; y = (m12 + m21) / s; z = 0.25f * s;
    }
    const float inv = rsqrtf(w * w + x * x + y * y + z * z) * (w < 0.0f ? -1.0f : 1.0f);
    float* o = out + r * 7;
    o[0] = w * inv; o[1] = x * inv; o[2] = y * inv; o[3] = z * inv;
    o[4] = r0.w; o[5] = r1.w; o[6] = r2.w;
}

struct Atom14Tables {
    const float* default_frames;   // [21,8,4,4] restype_rigid_group_default_frame
    const int32_t* group_idx;      // [21,14]    restype_atom14_to_rigid_group
    const float* lit_positions;    // [21,14,3]  restype_atom14_rigid_group_positions
    const uint8_t* atom_mask;      // [21,14]    restype_atom14_mask
};

__device__ __forceinline__ Frame load_default(const float* d) {
    Frame f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float4 row = __ldg(reinterpret_cast<const float4*>(d) + i);
        f.r.m[i][0] = row.x; f.r.m[i][1] = row.y; f.r.m[i][2] = row.z; f.t[i] = row.w;
    }
    return f;
}
// default frame composed with the torsion rotation about x (feats.torsion_angles_to_frames: rows [1,0,0], [0,c,-s], [0,s,c])
__device__ __forceinline__ Frame torsion_frame(const float* d, float s, float c) {
    const Frame f = load_default(d);
    Frame o;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        o.r.m[i][0] = f.r.m[i][0];
        o.r.m[i][1] = f.r.m[i][1] * c + f.r.m[i][2] * s;
        o.r.m[i][2] = -f.r.m[i][1] * s + f.r.m[i][2] * c;
        o.t[i] = f.t[i];
    }
    return o;
}
__device__ __forceinline__ void normalize3(float* v) {
    // torch.nn.functional.normalize: v / max(|v|, 1e-12)
    const float n = fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12f);
    v[0] /= n; v[1] /= n; v[2] /= n;
}

// positions [B,16,15,3]: atom14 order (N, CA, C, O, CB, side chain ...) + OXT in slot 14; exists [B,16,15].
__global__ void atom14_kernel(const float* __restrict__ frames, const float* __restrict__ tors, const int64_t* __restrict__ aatype,
                              const uint8_t* __restrict__ mask, int64_t n_res, Atom14Tables tb, float* __restrict__ pos,
                              uint8_t* __restrict__ exists) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_res) return;
    float* po = pos + r * 45;
    uint8_t* ex = exists + r * 15;
    if (mask[r] == 0) {
        for (int c = 0; c < 45; ++c) po[c] = 0.0f;
        for (int c = 0; c < 15; ++c) ex[c] = 0;
        return;
    }
    const int i = (int)(r % kN);
    int aa = (int)aatype[r];
    aa = aa < 0 ? 20 : (aa > 20 ? 20 : aa);
    const float* f7 = frames + r * 7;
    const Quat q{f7[0], f7[1], f7[2], f7[3]};
    // The torsion frames use the quaternion as stored (pdb.py:67-72); the backbone atoms use it normalised (pdb.py:99-102).
    Frame bb, bbn;
    bb.r = quat_to_rot(q);
    bbn.r = quat_to_rot(qnormalize(q));
#pragma unroll
    for (int c = 0; c < 3; ++c) bb.t[c] = bbn.t[c] = f7[4 + c];

    // rigid groups 0..7: backbone, pre-omega, phi, psi, chi1..chi4; chi2.. chain through chi1 (feats.torsion_angles_to_frames)
    const float* df = tb.default_frames + (size_t)aa * 8 * 16;
    const float* t = tors + r * 14;
    Frame g[8];
    g[0] = compose(bb, torsion_frame(df, 0.0f, 1.0f));
#pragma unroll
    for (int k = 1; k <= 4; ++k) g[k] = compose(bb, torsion_frame(df + k * 16, t[2 * (k - 1)], t[2 * (k - 1) + 1]));
    {
        Frame chain = torsion_frame(df + 4 * 16, t[6], t[7]);
#pragma unroll
        for (int k = 5; k < 8; ++k) {
            chain = compose(chain, torsion_frame(df + k * 16, t[2 * (k - 1)], t[2 * (k - 1) + 1]));
            g[k] = compose(bb, chain);
        }
    }
    const float* lit = tb.lit_positions + (size_t)aa * 42;
    const int32_t* gi = tb.group_idx + aa * 14;
    const uint8_t* am = tb.atom_mask + aa * 14;
    float N[3], CA[3], C[3];
#pragma unroll
    for (int a = 0; a < 14; ++a) {
        float p[3] = {0.0f, 0.0f, 0.0f};
        const int grp = gi[a];
        if (am[a]) {
            const float l[3] = {lit[3 * a], lit[3 * a + 1], lit[3 * a + 2]};
            if (grp == 0) apply3(bbn, l, p);                      // N, CA, C, CB: the normalised backbone frame (pdb.py:99-121)
            else {
                Frame f = g[0];
#pragma unroll
                for (int k = 1; k < 8; ++k) if (grp == k) f = g[k];
                apply3(f, l, p);
            }
        }
        if (a == 0) { N[0] = p[0]; N[1] = p[1]; N[2] = p[2]; }
        if (a == 1) { CA[0] = p[0]; CA[1] = p[1]; CA[2] = p[2]; }
        if (a == 2) { C[0] = p[0]; C[1] = p[1]; C[2] = p[2]; }
        if (a != 3) {
            po[3 * a] = p[0]; po[3 * a + 1] = p[1]; po[3 * a + 2] = p[2];
            ex[a] = am[a];
        }
    }
    (void)N;
    // backbone oxygen
    const bool terminal = (i + 1 >= kN) || mask[r + 1] == 0;
    float cac[3] = {C[0] - CA[0], C[1] - CA[1], C[2] - CA[2]};
    normalize3(cac);
    float O[3], OXT[3] = {0.0f, 0.0f, 0.0f};
    if (!terminal) {
        // bisector of CA->C and N(next)->C, 1.24 A from C (pdb.py:139-151); N(next) from the next residue's normalised frame
        const float* f7n = f7 + 7;
        int aan = (int)aatype[r + 1];
        aan = aan < 0 ? 20 : (aan > 20 ? 20 : aan);
        Frame nb;
        nb.r = quat_to_rot(qnormalize(Quat{f7n[0], f7n[1], f7n[2], f7n[3]}));
#pragma unroll
        for (int c = 0; c < 3; ++c) nb.t[c] = f7n[4 + c];
        const float* ln = tb.lit_positions + (size_t)aan * 42;
        const float l[3] = {ln[0], ln[1], ln[2]};
        float Nn[3];
        apply3(nb, l, Nn);
        float nc[3] = {C[0] - Nn[0], C[1] - Nn[1], C[2] - Nn[2]};
        normalize3(nc);
        float co[3] = {cac[0] + nc[0], cac[1] + nc[1], cac[2] + nc[2]};
        normalize3(co);
#pragma unroll
        for (int c = 0; c < 3; ++c) O[c] = C[c] + co[c] * 1.24f;
    } else {
        // psi frame applied to the literature O; OXT mirrors the C-O bond in the CA-C axis (pdb.py:153-174)
        const float l[3] = {lit[9], lit[10], lit[11]};
        apply3(g[3], l, O);
        const float co[3] = {O[0] - C[0], O[1] - C[1], O[2] - C[2]};
        const float d = co[0] * cac[0] + co[1] * cac[1] + co[2] * cac[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float proj = cac[c] * d;
            OXT[c] = C[c] + proj - (co[c] - proj);
        }
    }
    po[9] = O[0]; po[10] = O[1]; po[11] = O[2];
    ex[3] = 1;
    po[42] = OXT[0]; po[43] = OXT[1]; po[44] = OXT[2];
    ex[14] = terminal ? 1 : 0;
}

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

}  // namespace pmhc

using namespace pmhc;

extern "C" int pmhc_frames4x4_to_tensor7(const float* frames4x4, int64_t n, float* out7, void* stream) {
    if (n <= 0) return 0;
    PMHC_REQUIRE(frames4x4 != nullptr && out7 != nullptr, "pmhc_frames4x4_to_tensor7: null buffer");
    frames4x4_to_tensor7_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(frames4x4, n, out7);
    PMHC_CHECK_LAUNCH("pmhc_frames4x4_to_tensor7");
    return 0;
}

extern "C" int pmhc_atom14(const float* frames, const float* torsions, const int64_t* aatype, const uint8_t* mask, int B,
                           const float* default_frames, const int32_t* group_idx, const float* lit_positions,
                           const uint8_t* atom_mask, float* positions, uint8_t* exists, void* stream) {
    if (B <= 0) return 0;
    PMHC_REQUIRE(frames && torsions && aatype && mask && default_frames && group_idx && lit_positions && atom_mask && positions && exists,
                 "pmhc_atom14: null buffer");
    Atom14Tables tb{default_frames, group_idx, lit_positions, atom_mask};
    const int64_t n = (int64_t)B * kN;
    atom14_kernel<<<grid_for(n, 64), 64, 0, (cudaStream_t)stream>>>(frames, torsions, aatype, mask, n, tb, positions, exists);
    PMHC_CHECK_LAUNCH("pmhc_atom14");
    return 0;
}

// Host-side text formatting of one complex in the fixed PDB columns (what tools/pdb.py:206-209 gets from BioPython's PDBIO):
// chain P = peptide atoms in the order the reference adds them (N, CA, C, CB, side chain, O [, OXT], pdb.py:112-174), chain M =
// the protein's existing atom14 slots (pdb.py:177-204), serial numbers from 1 running through one TER record per chain, END.
// All pointers are HOST pointers.  atom_fields [21][15][4] / elements [21][15] / res3 [21][3]: the padded atom-name column,
// element letter and residue name per (residue type, atom slot).  Returns the number of bytes written, or -(bytes needed)
// when `out` is too small.
extern "C" int64_t pmhc_format_pdb_host(const int64_t* pep_aatype, const uint8_t* pep_mask, const float* pep_pos, const uint8_t* pep_exists,
                                        int64_t n_prot, const int64_t* prot_aatype, const float* prot_pos, const uint8_t* prot_exists,
                                        const char* atom_fields, const char* elements, const char* res3, char* out, int64_t out_cap) {
    static const int order[15] = {0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 3, 14};
    const int64_t max_lines = 16 * 15 + n_prot * 14 + 3;
    if (out_cap < max_lines * 96 + 1) return -(max_lines * 96 + 1);
    char* w = out;
    int serial = 0;
    // "%8.3f" without printf (glibc's float formatting is ~0.3 us a number): x * 1000 rounded half-to-even is exact for float
    // inputs in the PDB range (a float's distance from a decimal tie is either 0 or >> double rounding error), so the digits
    // are the ones printf gives.  A line with a number that does not fit its column goes through snprintf whole (and comes
    // out wider than 80, as it does from PDBIO).
    auto fits = [](float x) { const double v = (double)x; return v > -999.9994 && v < 9999.9994; };   // false for NaN too
    auto f83 = [](char* d, float x) {
        const double v = (double)x;
        long long m = (long long)nearbyint(fabs(v) * 1000.0);
        char buf[8];
        int p = 7;
        for (int k = 0; k < 3; ++k) { buf[p--] = (char)('0' + m % 10); m /= 10; }
        buf[p--] = '.';
        do { buf[p--] = (char)('0' + m % 10); m /= 10; } while (m > 0);
        if (signbit(v)) buf[p--] = '-';      // printf keeps the sign of -0.0 and of -0.0004 -> "-0.000"
        while (p >= 0) buf[p--] = ' ';
        memcpy(d, buf, 8);
    };
    auto atom = [&](int aa, int slot, char chain, int resseq, const float* p) {
        ++serial;
        // columns: 1-6 record, 7-11 serial, 13-16 name, 18-20 residue, 22 chain, 23-26 number, 31-54 x y z, 55-60 occupancy,
        // 61-66 B factor, 77-78 element
        int n = snprintf(w, 82, "ATOM  %5d %.4s %.3s %c%4d    ", serial, atom_fields + (aa * 15 + slot) * 4, res3 + aa * 3, chain, resseq);
        if (n != 30 || !fits(p[0]) || !fits(p[1]) || !fits(p[2])) {   // something beyond its column: let printf lay the line out
            w += snprintf(w, 96, "ATOM  %5d %.4s %.3s %c%4d    %8.3f%8.3f%8.3f%6.2f%6.2f          %2c  \n", serial, atom_fields + (aa * 15 + slot) * 4,
                          res3 + aa * 3, chain, resseq, (double)p[0], (double)p[1], (double)p[2], 1.0, 0.0, elements[aa * 15 + slot]);
            return;
        }
        f83(w + 30, p[0]);
        f83(w + 38, p[1]);
        f83(w + 46, p[2]);
        memcpy(w + 54, "  1.00  0.00           ", 23);
        w[77] = elements[aa * 15 + slot];
        w[78] = ' ';
        w[79] = ' ';
        w[80] = '\n';
        w += 81;
    };
    auto ter = [&](int aa, char chain, int resseq) {
        ++serial;
        const int n = snprintf(w, 82, "TER   %5d      %.3s %c%4d", serial, res3 + aa * 3, chain, resseq);
        memset(w + n, ' ', 80 - n);
        w[80] = '\n';
        w += 81;
    };
    int last_aa = -1, last_seq = 0;
    for (int i = 0; i < 16; ++i) {
        if (!pep_mask[i]) continue;
        const int aa = pep_aatype[i] < 0 || pep_aatype[i] > 20 ? 20 : (int)pep_aatype[i];
        for (int k = 0; k < 15; ++k) {
            const int a = order[k];
            if (pep_exists[i * 15 + a]) atom(aa, a, 'P', i + 1, pep_pos + (i * 15 + a) * 3);
        }
        last_aa = aa;
        last_seq = i + 1;
    }
    if (last_aa >= 0) ter(last_aa, 'P', last_seq);
    last_aa = -1;
    for (int64_t i = 0; i < n_prot; ++i) {
        const int aa = prot_aatype[i] < 0 || prot_aatype[i] > 20 ? 20 : (int)prot_aatype[i];
        for (int a = 0; a < 14; ++a)
            if (prot_exists[i * 14 + a]) atom(aa, a, 'M', (int)i + 1, prot_pos + (i * 14 + a) * 3);
        last_aa = aa;
        last_seq = (int)i + 1;
    }
    if (last_aa >= 0) ter(last_aa, 'M', last_seq);
    memcpy(w, "END   \n", 7);
    w += 7;
    return (int64_t)(w - out);
}
