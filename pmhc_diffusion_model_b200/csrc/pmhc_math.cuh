// pmhc_math.cuh — scalar geometry shared by every kernel (and by the host-side math tests).
//
// Quaternions are (w, x, y, z).  Each function cites the reference arithmetic it restates:
// RU = OpenFold rigid_utils (the reference's openfold 0.0.1 dependency), angle.py =
// diffusion/tools/angle.py, optimizer.py = diffusion/optimizer.py.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PMHC_HD __host__ __device__ __forceinline__
#else
#define PMHC_HD inline
#endif

namespace pmhc {

constexpr float kTwoPi = 6.283185307179586f;
constexpr float kNormEps = 1e-12f;  // torch.nn.functional.normalize eps

struct Quat {
    float w, x, y, z;
};
struct Vec3 {
    float x, y, z;
};
struct SinCos {
    float s, c;
};

// Hamilton product (RU:205-232 quat_multiply).
PMHC_HD Quat qmul(const Quat& a, const Quat& b) {
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return r;
}
PMHC_HD Quat qconj(const Quat& a) { return Quat{a.w, -a.x, -a.y, -a.z}; }
PMHC_HD float qdot(const Quat& a, const Quat& b) { return a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z; }
PMHC_HD Quat qscale(const Quat& a, float s) { return Quat{a.w * s, a.x * s, a.y * s, a.z * s}; }
PMHC_HD Quat qadd(const Quat& a, const Quat& b) { return Quat{a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z}; }
// conj(q) / |q|^2 (RU:246-250 invert_quat).
PMHC_HD Quat qinv(const Quat& a) {
    float n2 = qdot(a, a);
    return Quat{a.w / n2, -a.x / n2, -a.y / n2, -a.z / n2};
}
// q / max(|q|, eps) (torch.nn.functional.normalize).
PMHC_HD Quat qnormalize(const Quat& a) {
    float n = fmaxf(sqrtf(qdot(a, a)), kNormEps);
    return Quat{a.w / n, a.x / n, a.y / n, a.z / n};
}
// q / |q| (RU:283-287 Rotation.__init__ normalize_quats).
PMHC_HD Quat qunit(const Quat& a) {
    float n = sqrtf(qdot(a, a));
    return Quat{a.w / n, a.x / n, a.y / n, a.z / n};
}

// Gradient of L through c = a * b:  dL/da = dc * conj(b),  dL/db = conj(a) * dc.
PMHC_HD Quat qmul_grad_a(const Quat& dc, const Quat& b) { return qmul(dc, qconj(b)); }
PMHC_HD Quat qmul_grad_b(const Quat& a, const Quat& dc) { return qmul(qconj(a), dc); }
// Gradient through r = qinv(a): r = conj(a)/n2.
PMHC_HD Quat qinv_grad(const Quat& a, const Quat& dr) {
    float n2 = qdot(a, a);
    Quat c = qconj(a);
    float dn2 = -qdot(dr, c) / (n2 * n2);
    Quat dc = qscale(dr, 1.0f / n2);
    Quat da = qconj(dc);
    return qadd(da, qscale(a, 2.0f * dn2));
}
// Gradient through r = a / max(|a|, eps) (or a/|a|).
PMHC_HD Quat qnormalize_grad(const Quat& a, const Quat& dr) {
    float n = fmaxf(sqrtf(qdot(a, a)), kNormEps);
    Quat r = qscale(a, 1.0f / n);
    float proj = qdot(r, dr);
    return Quat{(dr.w - r.w * proj) / n, (dr.x - r.x * proj) / n, (dr.y - r.y * proj) / n, (dr.z - r.z * proj) / n};
}

// Complex product with sin = imaginary, cos = real (angle.py:139-152 multiply_sin_cos).
PMHC_HD SinCos scmul(const SinCos& a, const SinCos& b) { return SinCos{a.s * b.c + a.c * b.s, a.c * b.c - a.s * b.s}; }
// angle.py:155-162 inverse_sin_cos.
PMHC_HD SinCos scinv(const SinCos& a) {
    float n2 = a.s * a.s + a.c * a.c;
    return SinCos{-a.s / n2, a.c / n2};
}
// angle.py:165-174 partial_sin_cos: normalise, angle = +-acos(cos), scale, back to (sin, cos).
PMHC_HD SinCos scpartial(const SinCos& a, float amount) {
    float n = fmaxf(sqrtf(a.s * a.s + a.c * a.c), kNormEps);
    float s = a.s / n, c = a.c / n;
    float ang = acosf(fminf(fmaxf(c, -1.0f), 1.0f));
    if (s < 0.0f) ang = -ang;
    float so, co;
#if defined(__CUDA_ARCH__)
    sincosf(ang * amount, &so, &co);
#else
    so = sinf(ang * amount);
    co = cosf(ang * amount);
#endif
    return SinCos{so, co};
}
// angle.py:177-186 partial_rot: normalise q, half-angle = acos(w), axis = normalised vector part.
PMHC_HD Quat qpartial(const Quat& q, float amount) {
    Quat u = qnormalize(q);
    float half = acosf(fminf(fmaxf(u.w, -1.0f), 1.0f));
    float vn = fmaxf(sqrtf(u.x * u.x + u.y * u.y + u.z * u.z), kNormEps);
    float so, co;
#if defined(__CUDA_ARCH__)
    sincosf(half * amount, &so, &co);
#else
    so = sinf(half * amount);
    co = cosf(half * amount);
#endif
    return Quat{co, so * (u.x / vn), so * (u.y / vn), so * (u.z / vn)};
}
// angle.py:70-98 shoemake_quat on (u0, u1, u2) in [0,1], then RU:283-287 normalisation.
PMHC_HD Quat shoemake(float u0, float u1, float u2) {
    u0 = fminf(fmaxf(u0, 0.0f), 1.0f);
    u1 = fminf(fmaxf(u1, 0.0f), 1.0f);
    u2 = fminf(fmaxf(u2, 0.0f), 1.0f);
    float th1 = kTwoPi * u1, th2 = kTwoPi * u2;
    float r1 = sqrtf(1.0f - u0), r2 = sqrtf(u0);
    Quat q{r2 * cosf(th2), r1 * sinf(th1), r1 * cosf(th1), r2 * sinf(th2)};
    return qunit(q);
}

PMHC_HD float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- Philox4x32-10 counter-based generator (Salmon et al. 2011), for the perf-mode noise draw ----
struct Philox4 {
    uint32_t v[4];
};
PMHC_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}
PMHC_HD Philox4 philox4x32_10(uint64_t counter_lo, uint64_t counter_hi, uint64_t key) {
    uint32_t c0 = (uint32_t)counter_lo, c1 = (uint32_t)(counter_lo >> 32);
    uint32_t c2 = (uint32_t)counter_hi, c3 = (uint32_t)(counter_hi >> 32);
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
        philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{{c0, c1, c2, c3}};
}
// uint32 -> (0,1): never 0 or 1, so log() and sqrt(1-u) are safe.
PMHC_HD float u32_to_unit(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

}  // namespace pmhc
