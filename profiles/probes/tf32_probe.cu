// tf32_probe.cu — development probe (not part of the library): checks the tcgen05 kind::tf32 operand forms the tcgen05
// backward relies on against a CPU computation, and times them on the box.
//   T1  A, B SW128 K-major fp32 tiles (two 32-float atom columns), M = 128, N = 64, K = 64
//   T2  B MN-major from the SAME weight tile (input-gradient GEMM: D[p][k] = sum_n Y[p][n] F[n][k])
//   T3  A and B MN-major from two pair tiles, M = 64, N = 64, K = 128 pairs (weight-gradient sums); raw TMEM dump -> D layout of M = 64
//   T4  same A, B = the first / second 16 columns of a [128][32] tile (N = 16, second one through a +64 B start address)
//   T5  operand rounding: full-mantissa inputs against truncated / rounded CPU operands
//   T6  timings
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tf32_probe tf32_probe.cu
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

struct Ptrs {
    const float *X, *Y;   // [128][64] pair tiles
    const float *F;       // [64][64]  weight tile (n, k)
    const float *E;       // [128][32] small tile
    float *D1, *D2;       // [128][64]
    float *D3;            // [128][64] raw dump (lane, column)
    float *D4a, *D4b;     // [128][16] raw dumps
    float *D5;            // [128][64] T1 again with M = 128 wgrad trick? (unused)
    long long *cyc;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Ptrs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int oX = 0, oY = 32768, oF = 65536, oE = 81920, oBar = 98304, oTp = 98336;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + oBar);
    if (warp == 0) tc::tmem_alloc(reinterpret_cast<uint32_t*>(smem + oTp), 512);
    if (tid == 32) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
    for (int idx = tid; idx < 128 * 64; idx += 128) {
        int r = idx >> 6, c = idx & 63;
        *reinterpret_cast<float*>(smem + oX + (c >> 5) * 16384 + tc::sw128_offset_f32(r, c & 31)) = p.X[idx];
        *reinterpret_cast<float*>(smem + oY + (c >> 5) * 16384 + tc::sw128_offset_f32(r, c & 31)) = p.Y[idx];
    }
    for (int idx = tid; idx < 64 * 64; idx += 128) {
        int n = idx >> 6, k = idx & 63;
        *reinterpret_cast<float*>(smem + oF + (k >> 5) * 8192 + tc::sw128_offset_f32(n, k & 31)) = p.F[idx];
    }
    for (int idx = tid; idx < 128 * 32; idx += 128) {
        int r = idx >> 5, c = idx & 31;
        *reinterpret_cast<float*>(smem + oE + tc::sw128_offset_f32(r, c)) = p.E[idx];
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + oTp);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t phase = 0;
    const uint32_t sX = tc::smem_u32(smem + oX), sY = tc::smem_u32(smem + oY), sF = tc::smem_u32(smem + oF), sE = tc::smem_u32(smem + oE);
    long long t0 = 0, t1 = 0;

    // sentinel in columns 128..255 (T3 / T4 destinations)
    {
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(-777.0f);
        for (int q = 0; q < 4; ++q) tc::tmem_st32(tmem + lane_base + 128 + 32 * q, r);
        tc::tmem_wait_st();
        tc::fence_before_thread_sync();
        __syncthreads();
    }
    // ---- T1: D[p][n] = sum_k X[p][k] F[n][k]; columns 0..63 ----
    if (tid == 0) {
        tc::fence_after_thread_sync();
        t0 = clock64();
        for (int s = 0; s < 8; ++s) {
            const uint64_t da = tc::smem_desc_sw128(sX + (s >> 2) * 16384) + 2 * (s & 3);
            const uint64_t db = tc::smem_desc_sw128(sF + (s >> 2) * 8192) + 2 * (s & 3);
            tc::mma_tf32(tmem + 0, da, db, tc::idesc_tf32(128, 64, 0, 0), s > 0);
        }
        tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, phase); phase ^= 1;
    if (tid == 0) { t1 = clock64(); p.cyc[0] = t1 - t0; }
    tc::fence_after_thread_sync();
    {
        float v[32];
        for (int h = 0; h < 2; ++h) {
            tc::tmem_ld32(tmem + lane_base + 32 * h, v);
            for (int c = 0; c < 32; ++c) p.D1[(size_t)tid * 64 + 32 * h + c] = v[c];
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    // ---- T2: D[p][k] = sum_n Y[p][n] F[n][k]: B MN-major over the same F tile; columns 64..127 ----
    if (tid == 0) {
        tc::fence_after_thread_sync();
        t0 = clock64();
        for (int s = 0; s < 8; ++s) {   // k-step s: hidden units n in [8 s, 8 s + 8)
            const uint64_t da = tc::smem_desc_sw128(sY + (s >> 2) * 16384) + 2 * (s & 3);
            const uint64_t db = tc::smem_desc(sF + s * 1024, 8192, 1024, 2);
            tc::mma_tf32(tmem + 64, da, db, tc::idesc_tf32(128, 64, 0, 1), s > 0);
        }
        tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, phase); phase ^= 1;
    if (tid == 0) { t1 = clock64(); p.cyc[1] = t1 - t0; }
    tc::fence_after_thread_sync();
    {
        float v[32];
        for (int h = 0; h < 2; ++h) {
            tc::tmem_ld32(tmem + lane_base + 64 + 32 * h, v);
            for (int c = 0; c < 32; ++c) p.D2[(size_t)tid * 64 + 32 * h + c] = v[c];
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    // ---- T3: D[n][k] = sum_p Y[p][n] X[p][k]: A, B MN-major, M = 64, N = 64, K = 128 pairs; columns 128..191 ----
    if (tid == 0) {
        tc::fence_after_thread_sync();
        t0 = clock64();
        for (int s = 0; s < 16; ++s) {   // k-step s: pairs [8 s, 8 s + 8)
            const uint64_t da = tc::smem_desc(sY + s * 1024, 16384, 1024, 2);
            const uint64_t db = tc::smem_desc(sX + s * 1024, 16384, 1024, 2);
            tc::mma_tf32(tmem + 128, da, db, tc::idesc_tf32(64, 64, 1, 1), s > 0);
        }
        tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, phase); phase ^= 1;
    if (tid == 0) { t1 = clock64(); p.cyc[2] = t1 - t0; }
    tc::fence_after_thread_sync();
    {
        float v[32];
        for (int h = 0; h < 2; ++h) {
            tc::tmem_ld32(tmem + lane_base + 128 + 32 * h, v);
            for (int c = 0; c < 32; ++c) p.D3[(size_t)tid * 64 + 32 * h + c] = v[c];
        }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    // ---- T4: D[n][c] = sum_p Y[p][n] E[p][c], c < 16 (columns 192..207) and c in [16, 32) (columns 208..223) ----
    if (tid == 0) {
        tc::fence_after_thread_sync();
        t0 = clock64();
        for (int s = 0; s < 16; ++s) {
            const uint64_t da = tc::smem_desc(sY + s * 1024, 16384, 1024, 2);
            tc::mma_tf32(tmem + 192, da, tc::smem_desc(sE + s * 1024, 16384, 1024, 2), tc::idesc_tf32(64, 16, 1, 1), s > 0);
        }
        for (int s = 0; s < 16; ++s) {
            const uint64_t da = tc::smem_desc(sY + s * 1024, 16384, 1024, 2);
            tc::mma_tf32(tmem + 208, da, tc::smem_desc(sE + s * 1024 + 64, 16384, 1024, 2), tc::idesc_tf32(64, 16, 1, 1), s > 0);
        }
        tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, phase); phase ^= 1;
    if (tid == 0) { t1 = clock64(); p.cyc[3] = t1 - t0; }
    tc::fence_after_thread_sync();
    {
        float v[16];
        tc::tmem_ld16(tmem + lane_base + 192, v);
        for (int c = 0; c < 16; ++c) p.D4a[(size_t)tid * 16 + c] = v[c];
        tc::tmem_ld16(tmem + lane_base + 208, v);
        for (int c = 0; c < 16; ++c) p.D4b[(size_t)tid * 16 + c] = v[c];
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    // ---- T6: rates: 64 MMAs back to back of each form ----
    for (int form = 0; form < 4; ++form) {
        if (tid == 0) {
            tc::fence_after_thread_sync();
            t0 = clock64();
            for (int it = 0; it < 64; ++it) {
                const int s = it & 7, s16 = it & 15;
                if (form == 0) {
                    tc::mma_tf32(tmem + 256, tc::smem_desc_sw128(sX + (s >> 2) * 16384) + 2 * (s & 3),
                                 tc::smem_desc_sw128(sF + (s >> 2) * 8192) + 2 * (s & 3), tc::idesc_tf32(128, 64, 0, 0), 1);
                } else if (form == 1) {
                    tc::mma_tf32(tmem + 256, tc::smem_desc_sw128(sY + (s >> 2) * 16384) + 2 * (s & 3), tc::smem_desc(sF + s * 1024, 8192, 1024, 2),
                                 tc::idesc_tf32(128, 64, 0, 1), 1);
                } else if (form == 2) {
                    tc::mma_tf32(tmem + 256, tc::smem_desc(sY + s16 * 1024, 16384, 1024, 2), tc::smem_desc(sX + s16 * 1024, 16384, 1024, 2),
                                 tc::idesc_tf32(64, 64, 1, 1), 1);
                } else {
                    tc::mma_tf32(tmem + 256, tc::smem_desc(sY + s16 * 1024, 16384, 1024, 2), tc::smem_desc(sE + s16 * 1024, 16384, 1024, 2),
                                 tc::idesc_tf32(64, 16, 1, 1), 1);
                }
            }
            tc::mma_commit(bar);
            t1 = clock64();
            p.cyc[8 + form] = t1 - t0;   // issue time
        }
        tc::mbar_wait(bar, phase); phase ^= 1;
        if (tid == 0) { t1 = clock64(); p.cyc[4 + form] = t1 - t0; }
        tc::fence_after_thread_sync();
        __syncthreads();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rna_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
    std::vector<float> X(128 * 64), Y(128 * 64), F(64 * 64), E(128 * 32);
    srand(7);
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.0f - 1.0f; };
    for (auto& v : X) v = rnd();
    for (auto& v : Y) v = rnd();
    for (auto& v : F) v = rnd();
    for (auto& v : E) v = rnd();
    Ptrs p{};
    float *dX, *dY, *dF, *dE, *dD1, *dD2, *dD3, *dD4a, *dD4b;
    long long* dcyc;
    CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dY, Y.size() * 4)); CK(cudaMalloc(&dF, F.size() * 4)); CK(cudaMalloc(&dE, E.size() * 4));
    CK(cudaMalloc(&dD1, 128 * 64 * 4)); CK(cudaMalloc(&dD2, 128 * 64 * 4)); CK(cudaMalloc(&dD3, 128 * 64 * 4));
    CK(cudaMalloc(&dD4a, 128 * 16 * 4)); CK(cudaMalloc(&dD4b, 128 * 16 * 4)); CK(cudaMalloc(&dcyc, 16 * 8));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dE, E.data(), E.size() * 4, cudaMemcpyHostToDevice));
    p.X = dX; p.Y = dY; p.F = dF; p.E = dE; p.D1 = dD1; p.D2 = dD2; p.D3 = dD3; p.D4a = dD4a; p.D4b = dD4b; p.cyc = dcyc;
    const int smem = 98304 + 64 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(p);
    CK(cudaDeviceSynchronize());
    std::vector<float> D1(128 * 64), D2(128 * 64), D3(128 * 64), D4a(128 * 16), D4b(128 * 16);
    long long cyc[16];
    CK(cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(D4a.data(), dD4a, D4a.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D4b.data(), dD4b, D4b.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost));

    // T1 / T5
    double e_exact = 0, e_tr = 0, e_rn = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 64; ++n) {
            double ex = 0, tr = 0, rn = 0;
            for (int k = 0; k < 64; ++k) {
                ex += (double)X[r * 64 + k] * F[n * 64 + k];
                tr += (double)trunc_tf32(X[r * 64 + k]) * trunc_tf32(F[n * 64 + k]);
                rn += (double)rna_tf32(X[r * 64 + k]) * rna_tf32(F[n * 64 + k]);
            }
            double d = D1[r * 64 + n];
            e_exact = fmax(e_exact, fabs(d - ex)); e_tr = fmax(e_tr, fabs(d - tr)); e_rn = fmax(e_rn, fabs(d - rn));
        }
    printf("T1 K-major x K-major : max err vs exact %.3e, vs truncated operands %.3e, vs rna operands %.3e\n", e_exact, e_tr, e_rn);
    double e2 = 0;
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 64; ++k) {
            double tr = 0;
            for (int n = 0; n < 64; ++n) tr += (double)trunc_tf32(Y[r * 64 + n]) * trunc_tf32(F[n * 64 + k]);
            e2 = fmax(e2, fabs(D2[r * 64 + k] - tr));
        }
    printf("T2 B MN-major (F^T)  : max err vs truncated operands %.3e\n", e2);
    // T3: expected rows
    std::vector<double> W(64 * 64);
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 64; ++k) {
            double tr = 0;
            for (int q = 0; q < 128; ++q) tr += (double)trunc_tf32(Y[q * 64 + n]) * trunc_tf32(X[q * 64 + k]);
            W[n * 64 + k] = tr;
        }
    int lane_of_row[64];
    int found = 0;
    for (int n = 0; n < 64; ++n) {
        lane_of_row[n] = -1;
        for (int l = 0; l < 128; ++l) {
            double e = 0;
            for (int k = 0; k < 64; ++k) e = fmax(e, fabs(D3[l * 64 + k] - W[n * 64 + k]));
            if (e < 2e-3) { lane_of_row[n] = l; ++found; break; }
        }
    }
    printf("T3 A,B MN-major M=64 : %d / 64 rows found; row -> lane:", found);
    for (int n = 0; n < 64; ++n) printf(" %d", lane_of_row[n]);
    printf("\n   untouched lanes (sentinel):");
    for (int l = 0; l < 128; ++l) if (D3[l * 64] == -777.0f) printf(" %d", l);
    printf("\n");
    if (found < 64) {
        printf("   D3 lane 0: "); for (int k = 0; k < 8; ++k) printf("%.4f ", D3[k]); printf("\n   W row 0  : "); for (int k = 0; k < 8; ++k) printf("%.4f ", W[k]); printf("\n");
    }
    // T4
    for (int part = 0; part < 2; ++part) {
        const std::vector<float>& D4 = part ? D4b : D4a;
        double e4 = 0; int ok = 0;
        for (int n = 0; n < 64; ++n) {
            int l = lane_of_row[n] >= 0 ? lane_of_row[n] : n;
            for (int c = 0; c < 16; ++c) {
                double tr = 0;
                for (int q = 0; q < 128; ++q) tr += (double)trunc_tf32(Y[q * 64 + n]) * trunc_tf32(E[q * 32 + 16 * part + c]);
                e4 = fmax(e4, fabs(D4[l * 16 + c] - tr));
            }
            ++ok;
        }
        printf("T4 N=16, columns %d.. : max err %.3e\n", 16 * part, e4);
    }
    printf("T6 cycles: T1 (8 MMA N=64 + commit -> wait) %lld | T2 %lld | T3 (16 MMA M=64) %lld | T4 (32 MMA N=16) %lld\n", cyc[0], cyc[1], cyc[2], cyc[3]);
    printf("   64 MMAs to completion: K-major N=64 %lld | B MN-major %lld | M=64 both MN-major N=64 %lld | M=64 N=16 %lld\n", cyc[4], cyc[5], cyc[6], cyc[7]);
    printf("   64 MMAs issue only   : %lld | %lld | %lld | %lld\n", cyc[8], cyc[9], cyc[10], cyc[11]);
    return 0;
}
