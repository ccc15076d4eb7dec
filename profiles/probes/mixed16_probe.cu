// mixed16_probe.cu — development probe (not part of the library): tcgen05 kind::f16 with MIXED operand formats (bf16 x fp16) in the
// operand forms the tcgen05 backward uses, against a CPU computation on the rounded operands.
//   U1  A bf16 K-major, B fp16 K-major                     D[p][n] = sum_k X[p][k] F[n][k]         (M = 128, N = 64, K = 64)
//   U2  A bf16 K-major, B fp16 MN-major (same F tile)       D[p][k] = sum_n Y[p][n] F[n][k]
//   U3  A bf16 MN-major, B fp16 MN-major, M = 64            D[n][k] = sum_p Y[p][n] X[p][k]         (K = 128 pairs)
//   U3b same, M = 128 over two adjacent A tiles (Y | Y2)    D[64 a + n][k]
//   U4  A fp16 MN-major (M = 64), B bf16 MN-major N = 32 at column 0 / column 32 of a [128][64] tile
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mixed16_probe mixed16_probe.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../pmhc_diffusion_model_b200/csrc/tcgen05.cuh"

using namespace pmhc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// kind::f16 instruction descriptor with separate A / B formats (0 = f16, 1 = bf16)
__host__ __device__ constexpr uint32_t idesc16(int M, int N, int afmt, int bfmt, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Ptrs {
    const uint16_t *Xb, *Xh, *Yb, *Y2b, *Fh, *Eb;   // X as bf16 and fp16, Y / Y2 bf16, F fp16 [64][64], E bf16 [128][64]
    float *D1, *D2, *D3, *D3b, *D4a, *D4b;
    long long* cyc;
    int mask, fg, fa;   // fg: format of the gradient-like operands (Xb, Yb, Y2b, Eb), fa: of the activation-like ones (Xh, Fh); 0 = f16, 1 = bf16
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Ptrs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int oXb = 0, oXh = 16384, oYb = 32768, oY2b = 49152, oF = 65536, oE = 73728, oBar = 90112, oTp = 90144;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + oBar);
    if (warp == 0) tc::tmem_alloc(reinterpret_cast<uint32_t*>(smem + oTp), 512);
    if (tid == 32) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
    for (int idx = tid; idx < 128 * 64; idx += 128) {
        int r = idx >> 6, c = idx & 63;
        *reinterpret_cast<uint16_t*>(smem + oXb + tc::sw128_offset(r, c)) = p.Xb[idx];
        *reinterpret_cast<uint16_t*>(smem + oXh + tc::sw128_offset(r, c)) = p.Xh[idx];
        *reinterpret_cast<uint16_t*>(smem + oYb + tc::sw128_offset(r, c)) = p.Yb[idx];
        *reinterpret_cast<uint16_t*>(smem + oY2b + tc::sw128_offset(r, c)) = p.Y2b[idx];
        *reinterpret_cast<uint16_t*>(smem + oE + tc::sw128_offset(r, c)) = p.Eb[idx];
    }
    for (int idx = tid; idx < 64 * 64; idx += 128) {
        int n = idx >> 6, k = idx & 63;
        *reinterpret_cast<uint16_t*>(smem + oF + tc::sw128_offset(n, k)) = p.Fh[idx];
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + oTp);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t phase = 0;
    const uint32_t sXb = tc::smem_u32(smem + oXb), sXh = tc::smem_u32(smem + oXh), sYb = tc::smem_u32(smem + oYb), sF = tc::smem_u32(smem + oF),
                   sE = tc::smem_u32(smem + oE);
    {
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(-777.0f);
        for (int q = 0; q < 8; ++q) tc::tmem_st32(tmem + lane_base + 128 + 32 * q, r);
        tc::tmem_wait_st();
        tc::fence_before_thread_sync();
        __syncthreads();
    }
    auto dump = [&](int col, int n, float* dst) {
        tc::fence_after_thread_sync();
        for (int h = 0; h < n / 16; ++h) {
            float v[16];
            tc::tmem_ld16(tmem + lane_base + col + 16 * h, v);
            for (int c = 0; c < 16; ++c) dst[(size_t)tid * n + 16 * h + c] = v[c];
        }
        tc::fence_before_thread_sync();
        __syncthreads();
    };
    long long t0 = 0;
    // U1
    if (warp == 0 && (p.mask & 1)) {
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
            t0 = clock64();
            for (int s = 0; s < 4; ++s) tc::mma_bf16(tmem + 0, tc::smem_desc_sw128(sXb) + 2 * s, tc::smem_desc_sw128(sF) + 2 * s, idesc16(128, 64, p.fg, p.fa, 0, 0), s > 0);
            tc::mma_commit(bar);
        }
        __syncwarp();
    }
    if (p.mask & 1) { tc::mbar_wait(bar, phase); phase ^= 1; }
    if (tid == 0) p.cyc[0] = clock64() - t0;
    dump(0, 64, p.D1);
    // U2: B MN-major: k-step s covers hidden units n in [16 s, 16 s + 16) = two 8-row groups, 1024 B apart
    if (warp == 0 && (p.mask & 2)) {
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
            for (int s = 0; s < 4; ++s)
                tc::mma_bf16(tmem + 64, tc::smem_desc_sw128(sYb) + 2 * s, tc::smem_desc(sF + s * 2048, 8192, 1024, 2), idesc16(128, 64, p.fg, p.fa, 0, 1), s > 0);
            tc::mma_commit(bar);
        }
        __syncwarp();
    }
    if (p.mask & 2) { tc::mbar_wait(bar, phase); phase ^= 1; }
    dump(64, 64, p.D2);
    // U3: M = 64, both MN-major, K = 128 pairs (8 k-steps of 16 pairs)
    if (warp == 0 && (p.mask & 4)) {
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
            t0 = clock64();
            for (int s = 0; s < 8; ++s)
                tc::mma_bf16(tmem + 128, tc::smem_desc(sYb + s * 2048, 16384, 1024, 2), tc::smem_desc(sXh + s * 2048, 16384, 1024, 2),
                             idesc16(64, 64, p.fg, p.fa, 1, 1), s > 0);
            tc::mma_commit(bar);
        }
        __syncwarp();
    }
    if (p.mask & 4) { tc::mbar_wait(bar, phase); phase ^= 1; }
    if (tid == 0) p.cyc[1] = clock64() - t0;
    dump(128, 64, p.D3);
    // U3b: M = 128 over (Y | Y2): MN atoms 16384 B apart
    if (warp == 0 && (p.mask & 8)) {
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
            t0 = clock64();
            for (int s = 0; s < 8; ++s)
                tc::mma_bf16(tmem + 192, tc::smem_desc(sYb + s * 2048, 16384, 1024, 2), tc::smem_desc(sXh + s * 2048, 16384, 1024, 2),
                             idesc16(128, 64, p.fg, p.fa, 1, 1), s > 0);
            tc::mma_commit(bar);
        }
        __syncwarp();
    }
    if (p.mask & 8) { tc::mbar_wait(bar, phase); phase ^= 1; }
    if (tid == 0) p.cyc[2] = clock64() - t0;
    dump(192, 64, p.D3b);
    // U4: A fp16 MN-major (X), B = E columns [0, 32) and [32, 64)
    if (warp == 0 && (p.mask & 16)) {
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
            t0 = clock64();
            for (int s = 0; s < 8; ++s)
                tc::mma_bf16(tmem + 256, tc::smem_desc(sXh + s * 2048, 16384, 1024, 2), tc::smem_desc(sE + s * 2048, 16384, 1024, 2),
                             idesc16(64, 32, p.fa, p.fg, 1, 1), s > 0);
            for (int s = 0; s < 8; ++s)
                tc::mma_bf16(tmem + 288, tc::smem_desc(sXh + s * 2048, 16384, 1024, 2), tc::smem_desc(sE + s * 2048 + 64, 16384, 1024, 2),
                             idesc16(64, 32, p.fa, p.fg, 1, 1), s > 0);
            tc::mma_commit(bar);
        }
        __syncwarp();
    }
    if (p.mask & 16) { tc::mbar_wait(bar, phase); phase ^= 1; }
    if (tid == 0) p.cyc[3] = clock64() - t0;
    dump(256, 32, p.D4a);
    dump(288, 32, p.D4b);
    // issue rates from an elected lane: 64 MMAs of three shapes
    for (int form = 0; form < 3; ++form) {
        if (!(p.mask & 32)) break;
        if (warp == 0) {
            tc::fence_after_thread_sync();
            if (tc::elect_one()) {
                t0 = clock64();
                for (int it = 0; it < 64; ++it) {
                    const int s = it & 3, s8 = it & 7;
                    if (form == 0) tc::mma_bf16(tmem + 320, tc::smem_desc_sw128(sXb) + 2 * s, tc::smem_desc_sw128(sF) + 2 * s, idesc16(128, 64, p.fg, p.fa, 0, 0), 1);
                    else if (form == 1) tc::mma_bf16(tmem + 320, tc::smem_desc(sYb + s8 * 2048, 16384, 1024, 2), tc::smem_desc(sXh + s8 * 2048, 16384, 1024, 2), idesc16(64, 64, p.fg, p.fa, 1, 1), 1);
                    else tc::mma_bf16(tmem + 320, tc::smem_desc(sXh + s8 * 2048, 16384, 1024, 2), tc::smem_desc(sE + s8 * 2048, 16384, 1024, 2), idesc16(64, 32, p.fa, p.fg, 1, 1), 1);
                }
                p.cyc[8 + form] = clock64() - t0;
                tc::mma_commit(bar);
            }
            __syncwarp();
        }
        tc::mbar_wait(bar, phase); phase ^= 1;
        if (tid == 0) p.cyc[4 + form] = clock64() - t0;
        tc::fence_after_thread_sync();
        __syncthreads();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static uint16_t f2bf(float x) { __nv_bfloat16 b = __float2bfloat16_rn(x); uint16_t u; memcpy(&u, &b, 2); return u; }
static uint16_t f2h(float x) { __half h = __float2half_rn(x); uint16_t u; memcpy(&u, &h, 2); return u; }
static float bf2f(uint16_t u) { __nv_bfloat16 b; memcpy(&b, &u, 2); return __bfloat162float(b); }
static float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
static int g_fg = 1, g_fa = 0;
static float gf(uint16_t u) { return g_fg ? bf2f(u) : h2f(u); }
static float af(uint16_t u) { return g_fa ? bf2f(u) : h2f(u); }

int main(int argc, char** argv) {
    int mask = argc > 1 ? atoi(argv[1]) : 63;
    int fg = argc > 2 ? atoi(argv[2]) : 1, fa = argc > 3 ? atoi(argv[3]) : 0;
    srand(11);
    g_fg = argc > 2 ? atoi(argv[2]) : 1; g_fa = argc > 3 ? atoi(argv[3]) : 0;
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.0f - 1.0f; };
    std::vector<float> X(128 * 64), Y(128 * 64), Y2(128 * 64), F(64 * 64), E(128 * 64);
    for (auto& v : X) v = rnd(); for (auto& v : Y) v = rnd(); for (auto& v : Y2) v = rnd(); for (auto& v : F) v = rnd(); for (auto& v : E) v = rnd();
    std::vector<uint16_t> Xb(X.size()), Xh(X.size()), Yb(Y.size()), Y2b(Y2.size()), Fh(F.size()), Eb(E.size());
    auto cg = [&](float x) { return fg ? f2bf(x) : f2h(x); };
    auto ca = [&](float x) { return fa ? f2bf(x) : f2h(x); };
    for (size_t i = 0; i < X.size(); ++i) { Xb[i] = cg(X[i]); Xh[i] = ca(X[i]); Yb[i] = cg(Y[i]); Y2b[i] = cg(Y2[i]); Eb[i] = cg(E[i]); }
    for (size_t i = 0; i < F.size(); ++i) Fh[i] = ca(F[i]);
    auto up = [](const std::vector<uint16_t>& h) { uint16_t* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice)); return d; };
    auto outb = [](size_t n) { float* d; CK(cudaMalloc(&d, n * 4)); CK(cudaMemset(d, 0, n * 4)); return d; };
    Ptrs p{};
    p.Xb = up(Xb); p.Xh = up(Xh); p.Yb = up(Yb); p.Y2b = up(Y2b); p.Fh = up(Fh); p.Eb = up(Eb);
    p.D1 = outb(128 * 64); p.D2 = outb(128 * 64); p.D3 = outb(128 * 64); p.D3b = outb(128 * 64); p.D4a = outb(128 * 32); p.D4b = outb(128 * 32);
    CK(cudaMalloc(&p.cyc, 16 * 8)); CK(cudaMemset(p.cyc, 0, 128)); p.mask = mask; p.fg = fg; p.fa = fa;
    printf("mask %d formats gradient-like %s activation-like %s\n", mask, fg ? "bf16" : "f16", fa ? "bf16" : "f16");
    const int smem = 90112 + 64 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(p);
    CK(cudaDeviceSynchronize());
    auto down = [](float* d, size_t n) { std::vector<float> h(n); CK(cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost)); return h; };
    auto D1 = down(p.D1, 128 * 64), D2 = down(p.D2, 128 * 64), D3 = down(p.D3, 128 * 64), D3b = down(p.D3b, 128 * 64), D4a = down(p.D4a, 128 * 32), D4b = down(p.D4b, 128 * 32);
    long long cyc[16]; CK(cudaMemcpy(cyc, p.cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    double e = 0;
    for (int r = 0; r < 128; ++r) for (int n = 0; n < 64; ++n) { double s = 0; for (int k = 0; k < 64; ++k) s += (double)gf(Xb[r * 64 + k]) * af(Fh[n * 64 + k]); e = fmax(e, fabs(D1[r * 64 + n] - s)); }
    printf("U1 bf16 x fp16, K-major both        : max err %.3e\n", e);
    e = 0;
    for (int r = 0; r < 128; ++r) for (int k = 0; k < 64; ++k) { double s = 0; for (int n = 0; n < 64; ++n) s += (double)gf(Yb[r * 64 + n]) * af(Fh[n * 64 + k]); e = fmax(e, fabs(D2[r * 64 + k] - s)); }
    printf("U2 B MN-major (F^T)                 : max err %.3e\n", e);
    auto lane64 = [](int n) { return (n >> 4) * 32 + (n & 15); };
    e = 0;
    for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) { double s = 0; for (int q = 0; q < 128; ++q) s += (double)gf(Yb[q * 64 + n]) * af(Xh[q * 64 + k]); e = fmax(e, fabs(D3[lane64(n) * 64 + k] - s)); }
    printf("U3 A, B MN-major, M = 64            : max err %.3e (rows at lanes 32 (n / 16) + n %% 16)\n", e);
    e = 0;
    for (int a = 0; a < 2; ++a) for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) {
        double s = 0; for (int q = 0; q < 128; ++q) s += (double)gf((a ? Y2b : Yb)[q * 64 + n]) * af(Xh[q * 64 + k]);
        e = fmax(e, fabs(D3b[(64 * a + n) * 64 + k] - s)); }
    printf("U3b M = 128 over two A tiles        : max err %.3e\n", e);
    for (int part = 0; part < 2; ++part) {
        auto& D4 = part ? D4b : D4a; e = 0;
        for (int n = 0; n < 64; ++n) for (int c = 0; c < 32; ++c) { double s = 0; for (int q = 0; q < 128; ++q) s += (double)af(Xh[q * 64 + n]) * gf(Eb[q * 64 + 32 * part + c]); e = fmax(e, fabs(D4[lane64(n) * 32 + c] - s)); }
        printf("U4 fp16^T x bf16, N = 32 at column %2d: max err %.3e\n", 32 * part, e);
    }
    printf("cycles issue -> completion seen: U1 (4 MMA) %lld | U3 (8 MMA M=64) %lld | U3b (8 MMA M=128) %lld | U4 (16 MMA N=32) %lld\n", cyc[0], cyc[1], cyc[2], cyc[3]);
    printf("64 MMAs to completion: K-major N=64 %lld | M=64 MN/MN N=64 %lld | M=64 N=32 %lld ; issue only %lld | %lld | %lld\n", cyc[4], cyc[5], cyc[6], cyc[8], cyc[9], cyc[10]);
    return 0;
}
