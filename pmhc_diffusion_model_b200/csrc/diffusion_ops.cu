// diffusion_ops.cu — noise draw, forward noising, reverse step, loss (+gradient) and Adam.
//
// All HBM-bound, one thread per peptide residue (loss: one half-warp per complex), no shared-memory
// staging needed: every input element is read once, every output written once.
#include <string.h>

#include "common.cuh"
#include "pmhc_math.cuh"

namespace pmhc {

__device__ __forceinline__ Quat load_quat(const float* p) { return Quat{p[0], p[1], p[2], p[3]}; }
__device__ __forceinline__ void store_frame(float* p, const Quat& q, float x, float y, float z) {
    p[0] = q.w; p[1] = q.x; p[2] = q.y; p[3] = q.z; p[4] = x; p[5] = y; p[6] = z;
}
__device__ __forceinline__ Quat align_sign(const Quat& q, const float* ref) {
    if (ref == nullptr) return q;
    float d = q.w * ref[0] + q.x * ref[1] + q.y * ref[2] + q.z * ref[3];
    return d < 0.0f ? Quat{-q.w, -q.x, -q.y, -q.z} : q;
}

// Noise formulas shared by the Philox and the caller-supplied-randoms paths
// (optimizer.py:97-103, angle.py:33-98): translation 5 * N(0,1), Shoemake rotation, sin/cos of 2*pi*u.
__device__ __forceinline__ void write_noise(float* fr, float* tors, float n0, float n1, float n2, float u0,
                                            float u1, float u2, const float* ua) {
    Quat q = shoemake(u0, u1, u2);
    store_frame(fr, q, n0 * 5.0f, n1 * 5.0f, n2 * 5.0f);
#pragma unroll
    for (int c = 0; c < PMHC_NTORS; ++c) {
        float a = ua[c] * kTwoPi;
        tors[2 * c] = sinf(a);
        tors[2 * c + 1] = cosf(a);
    }
}

// One residue's noise from the counter-based stream: counter = global residue index, 4 Philox blocks (a: the three
// normals, b: Shoemake coordinates + torsion 0, c: torsions 1..4, d: torsions 5, 6).  Split by output part so that the
// per-residue kernels and the 8-lanes-per-residue fused reverse step run the very same arithmetic.
__device__ __forceinline__ void philox_translation_noise(uint64_t seed, uint64_t ctr, float* x) {
    Philox4 a = philox4x32_10(ctr, 0, seed);
    // Box-Muller on (a0,a1) and (a2,a3)
    float r0 = sqrtf(-2.0f * logf(u32_to_unit(a.v[0]))), r1 = sqrtf(-2.0f * logf(u32_to_unit(a.v[2])));
    float s0, c0, s1, c1;
    sincosf(kTwoPi * u32_to_unit(a.v[1]), &s0, &c0);
    sincosf(kTwoPi * u32_to_unit(a.v[3]), &s1, &c1);
    x[0] = (r0 * c0) * 5.0f; x[1] = (r0 * s0) * 5.0f; x[2] = (r1 * c1) * 5.0f;
}
__device__ __forceinline__ Quat philox_rotation_noise(uint64_t seed, uint64_t ctr) {
    Philox4 b = philox4x32_10(ctr, 1, seed);
    return shoemake(u32_to_unit(b.v[0]), u32_to_unit(b.v[1]), u32_to_unit(b.v[2]));
}
__device__ __forceinline__ void philox_frame_noise(uint64_t seed, uint64_t ctr, float* fr) {
    float x[3];
    philox_translation_noise(seed, ctr, x);
    store_frame(fr, philox_rotation_noise(seed, ctr), x[0], x[1], x[2]);
}
__device__ __forceinline__ SinCos philox_torsion_noise(uint64_t seed, uint64_t ctr, int c) {
    uint32_t u;
    if (c == 0) u = philox4x32_10(ctr, 1, seed).v[3];
    else if (c <= 4) u = philox4x32_10(ctr, 2, seed).v[c - 1];
    else u = philox4x32_10(ctr, 3, seed).v[c - 5];
    float a = u32_to_unit(u) * kTwoPi;
    return SinCos{sinf(a), cosf(a)};
}
__device__ __forceinline__ void philox_noise(uint64_t seed, uint64_t ctr, float* fr, float* tors) {
    philox_frame_noise(seed, ctr, fr);
#pragma unroll
    for (int c = 0; c < PMHC_NTORS; ++c) {
        SinCos t = philox_torsion_noise(seed, ctr, c);
        tors[2 * c] = t.s;
        tors[2 * c + 1] = t.c;
    }
}

// `sc` (nullable): the step's scalars in device memory, read instead of the by-value arguments (graph-replayable steps)
__global__ void gen_noise_kernel(uint64_t seed, uint64_t first, int64_t n, float* __restrict__ frames,
                                 float* __restrict__ tors, const PmhcStepScalars* __restrict__ sc) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (sc != nullptr) { seed = sc->noise_seed; first = sc->noise_first_residue; }
    philox_noise(seed, first + (uint64_t)r, frames + r * 7, tors + r * 14);
}

__global__ void noise_from_randoms_kernel(const float* __restrict__ normal, const float* __restrict__ uniform,
                                          int64_t n, float* __restrict__ frames, float* __restrict__ tors) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float* u = uniform + r * 10;
    float ua[PMHC_NTORS];
#pragma unroll
    for (int c = 0; c < PMHC_NTORS; ++c) ua[c] = u[3 + c];
    write_noise(frames + r * 7, tors + r * 14, normal[r * 3], normal[r * 3 + 1], normal[r * 3 + 2], u[0], u[1], u[2], ua);
}

// add_noise (optimizer.py:110-138)
__global__ void add_noise_kernel(const float* __restrict__ frames, const float* __restrict__ tors,
                                 const float* __restrict__ nframes, const float* __restrict__ ntors, float beta,
                                 float alpha, float sigma, int64_t n, const float* __restrict__ sign_ref,
                                 float* __restrict__ oframes, float* __restrict__ otors, const PmhcStepScalars* __restrict__ sc) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (sc != nullptr) { beta = sc->beta; alpha = sc->alpha; sigma = sc->sigma; }
    const float* f = frames + r * 7;
    const float* e = nframes + r * 7;
    Quat q = qunit(qmul(qpartial(load_quat(e), beta), load_quat(f)));
    q = align_sign(q, sign_ref ? sign_ref + r * 4 : nullptr);
    store_frame(oframes + r * 7, q, f[4] * alpha + e[4] * sigma, f[5] * alpha + e[5] * sigma, f[6] * alpha + e[6] * sigma);
#pragma unroll
    for (int c = 0; c < PMHC_NTORS; ++c) {
        SinCos eps{ntors[r * 14 + 2 * c], ntors[r * 14 + 2 * c + 1]};
        SinCos t0{tors[r * 14 + 2 * c], tors[r * 14 + 2 * c + 1]};
        SinCos o = scmul(scpartial(eps, beta), t0);
        otors[r * 14 + 2 * c] = o.s;
        otors[r * 14 + 2 * c + 1] = o.c;
    }
}

// remove_noise (optimizer.py:140-193)
struct ReverseCoef {
    float beta_t, beta_s, alpha_ts, var_ts, denom, sigma_t2s;
};

__device__ __forceinline__ Quat reverse_step_rotation(const Quat& z, const Quat& p, const Quat& x, const ReverseCoef& k,
                                                       const float* sign_ref) {
    Quat undo = qinv(qpartial(p, k.beta_t));
    Quat q = qunit(qmul(qpartial(x, k.beta_s), qmul(undo, z)));
    return align_sign(q, sign_ref);
}
__device__ __forceinline__ float reverse_step_coordinate(float z, float p, float x, const ReverseCoef& k) {
    return z / k.alpha_ts - (p * k.var_ts) / k.denom + k.sigma_t2s * x;
}
__device__ __forceinline__ void reverse_step_frame(const float* zf, const float* pf, const float* xf, const ReverseCoef& k,
                                                   const float* sign_ref, float* of) {
    Quat q = reverse_step_rotation(load_quat(zf), load_quat(pf), load_quat(xf), k, sign_ref);
    float px = reverse_step_coordinate(zf[4], pf[4], xf[4], k);
    float py = reverse_step_coordinate(zf[5], pf[5], xf[5], k);
    float pz = reverse_step_coordinate(zf[6], pf[6], xf[6], k);
    store_frame(of, q, px, py, pz);
}
__device__ __forceinline__ SinCos reverse_step_torsion(const SinCos& z, const SinCos& p, const SinCos& x, const ReverseCoef& k) {
    return scmul(scpartial(x, k.beta_s), scmul(scinv(scpartial(p, k.beta_t)), z));
}
__device__ __forceinline__ void reverse_step_residue(const float* zf, const float* zt, const float* pf, const float* pt,
                                                     const float* xf, const float* xt, const ReverseCoef& k,
                                                     const float* sign_ref, float* of, float* ot) {
    float tmp[14];
#pragma unroll
    for (int c = 0; c < PMHC_NTORS; ++c) {
        SinCos o = reverse_step_torsion(SinCos{zt[2 * c], zt[2 * c + 1]}, SinCos{pt[2 * c], pt[2 * c + 1]},
                                        SinCos{xt[2 * c], xt[2 * c + 1]}, k);
        tmp[2 * c] = o.s;
        tmp[2 * c + 1] = o.c;
    }
    reverse_step_frame(zf, pf, xf, k, sign_ref, of);
#pragma unroll
    for (int c = 0; c < 14; ++c) ot[c] = tmp[c];
}

__global__ void remove_noise_kernel(const float* zf, const float* zt, const float* __restrict__ pf,
                                    const float* __restrict__ pt, const float* __restrict__ xf,
                                    const float* __restrict__ xt, ReverseCoef k, int64_t n,
                                    const float* __restrict__ sign_ref, float* of, float* ot) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    reverse_step_residue(zf + r * 7, zt + r * 14, pf + r * 7, pt + r * 14, xf + r * 7, xt + r * 14, k,
                         sign_ref ? sign_ref + r * 4 : nullptr, of + r * 7, ot + r * 14);
}

// Fused noise draw + reverse step of the sampling loop: the step's fresh noise (optimizer.py:151) never leaves
// registers.  The kernel is a latency chain (Philox, Box-Muller, Shoemake, acos, sincos), not a throughput problem, and
// divergent roles inside a warp would run one after the other — so a CTA of nine warps takes 32 residues and gives every
// warp ONE role: warp 0 the rotations, warp 1 the translations, warps 2..8 one torsion each.  Bit-identical to
// gen_noise_kernel followed by remove_noise_kernel.
constexpr int kRevThreads = 9 * 32;
__global__ void __launch_bounds__(kRevThreads) reverse_step_philox_kernel(const float* zf, const float* zt, const float* __restrict__ pf,
                                                                          const float* __restrict__ pt, uint64_t seed, uint64_t first,
                                                                          ReverseCoef k, int64_t n, const float* __restrict__ sign_ref,
                                                                          float* of, float* ot, const uint64_t* __restrict__ seed_first_dev) {
    const int role = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
    if (r >= n) return;
    // (seed, first complex) from device memory: `seed` then carries only the per-step key offset (graph-replayable trajectories)
    if (seed_first_dev != nullptr) { seed += seed_first_dev[0]; first = seed_first_dev[1] * 16; }
    const uint64_t ctr = first + (uint64_t)r;
    if (role == 0) {
        const Quat q = reverse_step_rotation(load_quat(zf + r * 7), load_quat(pf + r * 7), philox_rotation_noise(seed, ctr), k,
                                             sign_ref ? sign_ref + r * 4 : nullptr);
        float* o = of + r * 7;
        o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z;
    } else if (role == 1) {
        float x[3];
        philox_translation_noise(seed, ctr, x);
        const float z4 = zf[r * 7 + 4], z5 = zf[r * 7 + 5], z6 = zf[r * 7 + 6];      // (of may alias zf)
        of[r * 7 + 4] = reverse_step_coordinate(z4, pf[r * 7 + 4], x[0], k);
        of[r * 7 + 5] = reverse_step_coordinate(z5, pf[r * 7 + 5], x[1], k);
        of[r * 7 + 6] = reverse_step_coordinate(z6, pf[r * 7 + 6], x[2], k);
    } else {
        const int c = role - 2;
        const SinCos x = philox_torsion_noise(seed, ctr, c);
        const SinCos o = reverse_step_torsion(SinCos{zt[r * 14 + 2 * c], zt[r * 14 + 2 * c + 1]},
                                              SinCos{pt[r * 14 + 2 * c], pt[r * 14 + 2 * c + 1]}, x, k);
        ot[r * 14 + 2 * c] = o.s;
        ot[r * 14 + 2 * c + 1] = o.c;
    }
}

// get_loss + gradient (optimizer.py:38-79): one half-warp (16 lanes = 16 residues) per complex.
__global__ void loss_kernel(const float* __restrict__ tf, const float* __restrict__ tt, const float* __restrict__ pf,
                            const float* __restrict__ pt, const uint8_t* __restrict__ mask,
                            const uint8_t* __restrict__ tmask, int B, float gscale, float* __restrict__ losses,
                            float* __restrict__ dpf, float* __restrict__ dpt, const PmhcStepScalars* __restrict__ sc) {
    if (sc != nullptr) gscale = sc->grad_scale;
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int b = gid >> 4, i = gid & 15;
    bool live = b < B;
    int64_t r = (int64_t)(live ? b : 0) * 16 + i;
    float m = live ? (float)mask[r] : 0.0f;
    const float* a = tf + r * 7;
    const float* p = pf + r * 7;
    float dx = a[4] - p[4], dy = a[5] - p[5], dz = a[6] - p[6];
    float sq = (dx * dx + dy * dy + dz * dz) * m;
    Quat qa = qnormalize(load_quat(a)), qp_raw = load_quat(p), qp = qnormalize(qp_raw);
    float rdev = (1.0f - qdot(qa, qp)) * m;
    float tdev = 0.0f, tcnt = 0.0f;
    float tm[PMHC_NTORS];
    SinCos ta[PMHC_NTORS], tp_raw[PMHC_NTORS];
#pragma unroll
    for (int c = 0; c < PMHC_NTORS; ++c) {
        tm[c] = live ? (float)tmask[r * 7 + c] : 0.0f;
        float as = tt[r * 14 + 2 * c], ac = tt[r * 14 + 2 * c + 1];
        float an = fmaxf(sqrtf(as * as + ac * ac), kNormEps);
        ta[c] = SinCos{as / an, ac / an};
        tp_raw[c] = SinCos{pt[r * 14 + 2 * c], pt[r * 14 + 2 * c + 1]};
        float pn = fmaxf(sqrtf(tp_raw[c].s * tp_raw[c].s + tp_raw[c].c * tp_raw[c].c), kNormEps);
        tdev += (1.0f - (ta[c].s * tp_raw[c].s + ta[c].c * tp_raw[c].c) / pn) * tm[c];
        tcnt += tm[c];
    }
    float nres = m;
    // reduce over the 16 lanes of this complex (xor offsets < 16 stay inside the half-warp)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        rdev += __shfl_xor_sync(0xffffffffu, rdev, o);
        tdev += __shfl_xor_sync(0xffffffffu, tdev, o);
        tcnt += __shfl_xor_sync(0xffffffffu, tcnt, o);
        nres += __shfl_xor_sync(0xffffffffu, nres, o);
    }
    if (!live) return;
    float pos_loss = sq / nres, rot_loss = rdev / nres, tors_loss = tdev / tcnt;
    if (i == 0) {
        losses[b] = 0.1f * pos_loss + rot_loss + tors_loss;
        losses[B + b] = pos_loss;
        losses[2 * B + b] = rot_loss;
        losses[3 * B + b] = tors_loss;
        losses[4 * B + b] = sqrtf(pos_loss);
    }
    if (dpf != nullptr) {
        float* g = dpf + r * 7;
        float wq = -gscale * m / nres;
        Quat dq = qnormalize_grad(qp_raw, qscale(qa, wq));
        float wx = gscale * 0.1f * m * 2.0f / nres;
        g[0] = dq.w; g[1] = dq.x; g[2] = dq.y; g[3] = dq.z;
        g[4] = -wx * dx; g[5] = -wx * dy; g[6] = -wx * dz;
#pragma unroll
        for (int c = 0; c < PMHC_NTORS; ++c) {
            float w = -gscale * tm[c] / tcnt;
            float ps = tp_raw[c].s, pc = tp_raw[c].c;
            float pn = fmaxf(sqrtf(ps * ps + pc * pc), kNormEps);
            float us = ps / pn, uc = pc / pn;
            float gs = w * ta[c].s, gc = w * ta[c].c;
            float proj = us * gs + uc * gc;
            dpt[r * 14 + 2 * c] = (gs - us * proj) / pn;
            dpt[r * 14 + 2 * c + 1] = (gc - uc * proj) / pn;
        }
    }
}

// torch.optim.Adam single-tensor update (no amsgrad, no weight decay) over the flat buffers.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float one_minus_b1, float b2, float one_minus_b2, float eps,
                            float step_size, float bc2_sqrt, const uint8_t* __restrict__ skip, const PmhcStepScalars* __restrict__ sc) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (sc != nullptr) { step_size = sc->adam_step_size; bc2_sqrt = sc->adam_bc2_sqrt; }
    if (skip != nullptr && *skip != 0) return;   // a non-finite loss was seen: weights and moments stay as they are
    float gi = g[i];
    float mi = m[i] + (gi - m[i]) * one_minus_b1;
    float vi = v[i] * b2 + one_minus_b2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
}

}  // namespace pmhc

using namespace pmhc;

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

extern "C" int pmhc_gen_noise(uint64_t seed, uint64_t first_residue, int64_t n, float* frames, float* tors, void* stream) {
    if (n <= 0) return 0;
    gen_noise_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(seed, first_residue, n, frames, tors, nullptr);
    PMHC_CHECK_LAUNCH("pmhc_gen_noise");
    return 0;
}

extern "C" int pmhc_noise_from_randoms(const float* normal, const float* uniform, int64_t n, float* frames,
                                       float* tors, void* stream) {
    if (n <= 0) return 0;
    noise_from_randoms_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(normal, uniform, n, frames, tors);
    PMHC_CHECK_LAUNCH("pmhc_noise_from_randoms");
    return 0;
}

extern "C" int pmhc_add_noise(const float* frames, const float* torsions, const float* nframes, const float* ntors,
                              double beta, int64_t n, const float* sign_ref, float* oframes, float* otors,
                              void* stream) {
    if (n <= 0) return 0;
    PMHC_REQUIRE(beta >= 0.0 && beta <= 1.0, "pmhc_add_noise: beta %f outside [0,1]", beta);
    float alpha = (float)sqrt(1.0 - beta), sigma = (float)sqrt(beta);
    add_noise_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(frames, torsions, nframes, ntors, (float)beta, alpha,
                                                                          sigma, n, sign_ref, oframes, otors, nullptr);
    PMHC_CHECK_LAUNCH("pmhc_add_noise");
    return 0;
}

namespace pmhc {
// host-side coefficients of one reverse step, in double like the reference's Python floats
// (optimizer.py:148-157; sigma_ts^2 = sigma_t^2 - sigma_s^2 * alpha_ts as written)
ReverseCoef reverse_coef(double beta_t, double beta_s) {
    double alpha_t = sqrt(1.0 - beta_t), alpha_s = sqrt(1.0 - beta_s);
    double sigma_t = sqrt(beta_t), sigma_s = sqrt(beta_s);
    double alpha_ts = alpha_t / alpha_s;
    double var_ts = sigma_t * sigma_t - sigma_s * sigma_s * alpha_ts;
    double sigma_ts = sqrt(var_ts);
    ReverseCoef k;
    k.beta_t = (float)beta_t;
    k.beta_s = (float)beta_s;
    k.alpha_ts = (float)alpha_ts;
    k.var_ts = (float)var_ts;
    k.denom = (float)(alpha_ts * sigma_t);
    k.sigma_t2s = (float)(sigma_ts * sigma_s / sigma_t);
    return k;
}

int launch_remove_noise(const float* zf, const float* zt, const float* pf, const float* pt, const float* xf,
                        const float* xt, double beta_t, double beta_s, int64_t n, const float* sign_ref, float* of,
                        float* ot, cudaStream_t stream) {
    PMHC_REQUIRE(beta_t > 0.0 && beta_t < 1.0 && beta_s >= 0.0 && beta_s < beta_t,
                 "pmhc_remove_noise: need 0 <= beta_s < beta_t < 1 (got %f, %f)", beta_s, beta_t);
    ReverseCoef k = reverse_coef(beta_t, beta_s);
    remove_noise_kernel<<<grid_for(n, 128), 128, 0, stream>>>(zf, zt, pf, pt, xf, xt, k, n, sign_ref, of, ot);
    PMHC_CHECK_LAUNCH("pmhc_remove_noise");
    return 0;
}
int launch_reverse_step_philox(const float* zf, const float* zt, const float* pf, const float* pt, uint64_t seed,
                               uint64_t first, double beta_t, double beta_s, int64_t n, const float* sign_ref, float* of,
                               float* ot, cudaStream_t stream, const uint64_t* seed_first_dev) {
    PMHC_REQUIRE(beta_t > 0.0 && beta_t < 1.0 && beta_s >= 0.0 && beta_s < beta_t,
                 "reverse step: need 0 <= beta_s < beta_t < 1 (got %f, %f)", beta_s, beta_t);
    ReverseCoef k = reverse_coef(beta_t, beta_s);
    reverse_step_philox_kernel<<<grid_for(n, 32), kRevThreads, 0, stream>>>(zf, zt, pf, pt, seed, first, k, n, sign_ref, of, ot, seed_first_dev);
    PMHC_CHECK_LAUNCH("reverse_step_philox");
    return 0;
}
}  // namespace pmhc

extern "C" int pmhc_remove_noise(const float* zf, const float* zt, const float* pf, const float* pt, const float* xf,
                                 const float* xt, double beta_t, double beta_s, int64_t n, const float* sign_ref,
                                 float* of, float* ot, void* stream) {
    if (n <= 0) return 0;
    return launch_remove_noise(zf, zt, pf, pt, xf, xt, beta_t, beta_s, n, sign_ref, of, ot, (cudaStream_t)stream);
}

extern "C" int pmhc_loss(const float* tf, const float* tt, const float* pf, const float* pt, const uint8_t* mask,
                         const uint8_t* tmask, int B, float gscale, float* losses, float* dpf, float* dpt,
                         void* stream) {
    if (B <= 0) return 0;
    PMHC_REQUIRE((dpf == nullptr) == (dpt == nullptr), "pmhc_loss: pass both gradient buffers or neither");
    loss_kernel<<<grid_for((int64_t)B * 16, 128), 128, 0, (cudaStream_t)stream>>>(tf, tt, pf, pt, mask, tmask, B, gscale,
                                                                                  losses, dpf, dpt, nullptr);
    PMHC_CHECK_LAUNCH("pmhc_loss");
    return 0;
}

extern "C" int pmhc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double b1, double b2,
                              double eps, int step, void* stream) {
    return pmhc_adam_step_guarded(p, g, m, v, n, lr, b1, b2, eps, step, nullptr, stream);
}

extern "C" int pmhc_adam_step_guarded(float* p, const float* g, float* m, float* v, int64_t n, double lr, double b1, double b2,
                                      double eps, int step, const uint8_t* skip_flag, void* stream) {
    if (n <= 0) return 0;
    PMHC_REQUIRE(step >= 1, "pmhc_adam_step: step counts from 1");
    // the scalars are formed in double and rounded once, as torch does with its Python-float hyperparameters
    double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
    adam_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2),
                                                                   (float)eps, (float)(lr / bc1), (float)sqrt(bc2), skip_flag, nullptr);
    PMHC_CHECK_LAUNCH("pmhc_adam_step");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// One training step as two enqueue-only calls (optimizer.py:195-224).  Every per-step scalar (t / T, the noising coefficients,
// the Philox key, the loss scale, Adam's bias corrections) comes either by value from `sc_host` or — when `sc_dev` is given —
// from device memory, so a CUDA graph captured around the calls can be replayed step after step: the caller refreshes the 48-byte
// block (a host-to-device copy node of the same graph) and launches the graph.
// ---------------------------------------------------------------------------------------------------------------
namespace pmhc {
int model_forward_impl(const float* params, const PmhcBatch* bt, float t_over_T, float* out_frames, float* out_torsions,
                       float* saved, void* workspace, size_t workspace_bytes, cudaStream_t stream, int precision,
                       bool reuse_pocket_cache);

__global__ void nan_flag_kernel(const float* __restrict__ total_loss, int B, uint8_t* __restrict__ flag) {
    bool bad = false;
    for (int b = threadIdx.x; b < B; b += blockDim.x) bad = bad || isnan(total_loss[b]);
    if (__syncthreads_or(bad) && threadIdx.x == 0) *flag = 1;     // sticky: never cleared here
}
}  // namespace pmhc

namespace pmhc {
struct SmallBlock { unsigned char b[64]; };
__global__ void upload_small_kernel(SmallBlock v, unsigned char* __restrict__ dst, int bytes) {
    if ((int)threadIdx.x < bytes) dst[threadIdx.x] = v.b[threadIdx.x];
}
}  // namespace pmhc

// Up to 64 bytes from host to device memory THROUGH THE LAUNCH PARAMETERS of a one-warp kernel: the bytes are captured when the
// call returns (no pinned staging buffer whose reuse could race with an earlier, still queued copy) and the write is ordered
// on `stream` like any kernel — what refreshes the scalar block of a replayed training-step / trajectory graph.
extern "C" int pmhc_upload_small(const void* src_host, void* dst_dev, int bytes, void* stream) {
    PMHC_REQUIRE(src_host != nullptr && dst_dev != nullptr && bytes > 0 && bytes <= 64, "pmhc_upload_small: 1..64 bytes");
    SmallBlock v;
    memset(v.b, 0, sizeof(v.b));
    memcpy(v.b, src_host, (size_t)bytes);
    upload_small_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(v, (unsigned char*)dst_dev, bytes);
    PMHC_CHECK_LAUNCH("pmhc_upload_small");
    return 0;
}

extern "C" int pmhc_step_scalars(int t, int T, double beta_min, double beta_max, double lr, double beta1, double beta2,
                                 int adam_step, double grad_scale, uint64_t noise_seed, uint64_t noise_first_residue,
                                 PmhcStepScalars* out_host) {
    PMHC_REQUIRE(out_host != nullptr && T > 0 && adam_step >= 1, "pmhc_step_scalars: bad arguments");
    const double beta = beta_min + (beta_max - beta_min) * ((double)t / (double)T);     // linear_schedule, optimizer.py:20-21
    PMHC_REQUIRE(beta >= 0.0 && beta <= 1.0, "pmhc_step_scalars: beta %f outside [0,1]", beta);
    out_host->t_over_T = (float)((double)t / (double)T);      // the Python float t / T, rounded once
    out_host->beta = (float)beta;
    out_host->alpha = (float)sqrt(1.0 - beta);
    out_host->sigma = (float)sqrt(beta);
    out_host->adam_step_size = (float)(lr / (1.0 - pow(beta1, adam_step)));
    out_host->adam_bc2_sqrt = (float)sqrt(1.0 - pow(beta2, adam_step));
    out_host->grad_scale = (float)grad_scale;
    out_host->reserved = 0.0f;
    out_host->noise_seed = noise_seed;
    out_host->noise_first_residue = noise_first_residue;
    return 0;
}

extern "C" int pmhc_train_step_grad(const float* params, const PmhcBatch* bt, const uint8_t* torsions_mask,
                                    const PmhcStepScalars* sc_host, const PmhcStepScalars* sc_dev, const PmhcStepBuffers* buf,
                                    int draw_noise, const float* quat_sign_ref, void* workspace, size_t workspace_bytes,
                                    void* stream_, void* layer2_done_event, int precision, int backward_precision) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PMHC_REQUIRE(bt != nullptr && bt->B > 0 && sc_host != nullptr && buf != nullptr && torsions_mask != nullptr,
                 "pmhc_train_step_grad: empty batch or missing arguments");
    const int B = bt->B;
    const int64_t n = (int64_t)B * kN;
    if (draw_noise) {
        gen_noise_kernel<<<grid_for(n, 128), 128, 0, stream>>>(sc_host->noise_seed, sc_host->noise_first_residue, n, buf->noise_frames,
                                                                buf->noise_torsions, sc_dev);
        PMHC_CHECK_LAUNCH("gen_noise");
    }
    add_noise_kernel<<<grid_for(n, 128), 128, 0, stream>>>(bt->frames, bt->torsions, buf->noise_frames, buf->noise_torsions, sc_host->beta,
                                                            sc_host->alpha, sc_host->sigma, n, quat_sign_ref, buf->zt_frames,
                                                            buf->zt_torsions, sc_dev);
    PMHC_CHECK_LAUNCH("add_noise");
    PmhcBatch zt = *bt;
    zt.frames = buf->zt_frames;
    zt.torsions = buf->zt_torsions;
    set_step_t_dev(sc_dev ? &sc_dev->t_over_T : nullptr);
    int rc = model_forward_impl(params, &zt, sc_host->t_over_T, buf->pred_frames, buf->pred_torsions, buf->saved, workspace,
                                workspace_bytes, stream, precision, false);
    if (rc == 0) {
        loss_kernel<<<grid_for(n, 128), 128, 0, stream>>>(buf->noise_frames, buf->noise_torsions, buf->pred_frames, buf->pred_torsions,
                                                           bt->mask, torsions_mask, B, sc_host->grad_scale, buf->losses, buf->d_frames,
                                                           buf->d_torsions, sc_dev);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("loss: %s", cudaGetErrorString(e)); rc = -2; }
    }
    if (rc == 0 && buf->nan_flag != nullptr) nan_flag_kernel<<<1, 256, 0, stream>>>(buf->losses, B, buf->nan_flag);
    if (rc == 0) {
        cudaError_t e = cudaMemsetAsync(buf->flat_grad, 0, sizeof(float) * PMHC_NPARAM, stream);
        if (e != cudaSuccess) { set_error("memset(flat_grad): %s", cudaGetErrorString(e)); rc = -2; }
    }
    if (rc == 0)
        rc = pmhc_model_backward_ex(params, &zt, sc_host->t_over_T, buf->saved, buf->d_frames, buf->d_torsions, buf->flat_grad, workspace,
                                    workspace_bytes, stream_, layer2_done_event, backward_precision);
    set_step_t_dev(nullptr);
    return rc;
}

extern "C" int pmhc_train_step_adam(float* params, const float* flat_grad, float* exp_avg, float* exp_avg_sq, double beta1,
                                    double beta2, double eps, const PmhcStepScalars* sc_host, const PmhcStepScalars* sc_dev,
                                    const uint8_t* nan_flag, void* stream) {
    PMHC_REQUIRE(sc_host != nullptr, "pmhc_train_step_adam: scalars are required");
    // gnn2.feature_mlp.* (state-dict tensors 24..27) never carries a gradient (model.py:415): no state, no update, as torch's
    // Adam treats grad = None
    const int64_t lo = pmhc_param_offset(24), hi = pmhc_param_offset(27) + pmhc_param_numel(27);
    const int64_t spans[2][2] = {{0, lo}, {hi, PMHC_NPARAM}};
    for (int k = 0; k < 2; ++k) {
        const int64_t o = spans[k][0], n = spans[k][1] - o;
        adam_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(params + o, flat_grad + o, exp_avg + o, exp_avg_sq + o, n,
                                                                       (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
                                                                       sc_host->adam_step_size, sc_host->adam_bc2_sqrt, nan_flag, sc_dev);
        PMHC_CHECK_LAUNCH("adam");
    }
    return 0;
}
