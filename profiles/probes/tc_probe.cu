// tc_probe.cu — development probe (not part of the library): checks the tcgen05 operand forms the fused EGNN kernel
// relies on against a CPU computation, and times TMEM loads / MMA issue on the box.
//   T1  A from TMEM (packed bf16x2 written with tcgen05.st), B SW128 K-major from shared memory
//   T2  B in the no-swizzle K-major core-matrix layout (K = 16 extension block)
//   T3  A MN-major from a SW128 tile (pairs as the K dimension: column sums of a pair tile via the tensor core)
//   T4  N = 16, K = 256 (second layers), A from TMEM
//   T5  timings
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_probe tc_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../pmhc_diffusion_model_b200/csrc/tcgen05.cuh"

using namespace pmhc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// A from tensor memory
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_full(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

struct Ptrs {
    const __nv_bfloat16 *A1;   // [128][64]   T1 operand A (rows = TMEM lanes)
    const __nv_bfloat16 *B1;   // [64][64]    T1 operand B (n, k)
    const __nv_bfloat16 *Ax;   // [128][16]   T2 extension block of A
    const __nv_bfloat16 *Bx;   // [64][16]    T2 extension block of B
    const __nv_bfloat16 *Sel;  // [32][64]    T3 operand B (n, pair)
    const __nv_bfloat16 *A4;   // [128][256]  T4 operand A
    const __nv_bfloat16 *B4;   // [16][256]   T4 operand B
    float *D1;                 // [128][64]   T1 result, then T2 accumulates on top -> D2
    float *D2;                 // [128][64]
    float *D3;                 // [128][32]
    float *D4;                 // [128][16]
    long long *cyc;            // timings
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Ptrs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    // layout: B1 SW128 [64x64] 8 KB @0 | A1 tile SW128 [128x64] 16 KB @8192 | Sel SW128 [32x64] 4 KB @24576 |
    //         Bx no-swizzle [64x16] 2 KB @28672 | B4 SW128 4 blocks [16x64] 8 KB @30720 | bars @38912 | tmem ptr @38944
    const int oB1 = 0, oA1 = 8192, oSel = 24576, oBx = 28672, oB4 = 30720, oBar = 38912, oTp = 38944;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + oBar);
    if (warp == 0) tc::tmem_alloc(reinterpret_cast<uint32_t*>(smem + oTp), 512);
    if (tid == 32) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
    for (int idx = tid; idx < 64 * 64; idx += 128) {
        int n = idx >> 6, k = idx & 63;
        *reinterpret_cast<__nv_bfloat16*>(smem + oB1 + tc::sw128_offset(n, k)) = p.B1[idx];
    }
    for (int idx = tid; idx < 128 * 64; idx += 128) {
        int r = idx >> 6, k = idx & 63;
        *reinterpret_cast<__nv_bfloat16*>(smem + oA1 + tc::sw128_offset(r, k)) = p.A1[idx];
    }
    for (int idx = tid; idx < 32 * 64; idx += 128) {
        int n = idx >> 6, k = idx & 63;
        *reinterpret_cast<__nv_bfloat16*>(smem + oSel + tc::sw128_offset(n, k)) = p.Sel[idx];
    }
    for (int idx = tid; idx < 64 * 16; idx += 128) {
        int n = idx >> 4, k = idx & 15;
        *reinterpret_cast<__nv_bfloat16*>(smem + oBx + (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) = p.Bx[idx];
    }
    for (int idx = tid; idx < 16 * 256; idx += 128) {
        int n = idx >> 8, k = idx & 255;
        *reinterpret_cast<__nv_bfloat16*>(smem + oB4 + (k >> 6) * 2048 + tc::sw128_offset(n, k & 63)) = p.B4[idx];
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + oTp);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t phase = 0;

    // ---- T1 + T2: A (and its extension block) in TMEM columns 0..31 and 32..39; D at columns 64..127 ----
    {
        uint32_t r[32];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.A1 + (size_t)tid * 64);
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = src[c];
        tmem_st32(tmem + lane_base + 0, r);
        uint32_t x[8];
        const uint32_t* sx = reinterpret_cast<const uint32_t*>(p.Ax + (size_t)tid * 16);
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = sx[c];
        tmem_st8(tmem + lane_base + 32, x);
        tmem_wait_st();
        tc::fence_before_thread_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_thread_sync();
            const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + oB1));
            for (int s = 0; s < 4; ++s) mma_ts(tmem + 64, tmem + 8 * s, db + 2 * s, idesc_full(128, 64, 0, 0), s > 0);
            tc::mma_commit(bar);
        }
        tc::mbar_wait(bar, phase); phase ^= 1;
        tc::fence_after_thread_sync();
        float v[32];
        for (int h = 0; h < 2; ++h) {
            tc::tmem_ld32(tmem + lane_base + 64 + 32 * h, v);
            for (int c = 0; c < 32; ++c) p.D1[(size_t)tid * 64 + 32 * h + c] = v[c];
        }
        tc::fence_before_thread_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_thread_sync();
            const uint64_t dx = smem_desc(tc::smem_u32(smem + oBx), 128, 256, 0);
            mma_ts(tmem + 64, tmem + 32, dx, idesc_full(128, 64, 0, 0), 1);
            tc::mma_commit(bar);
        }
        tc::mbar_wait(bar, phase); phase ^= 1;
        tc::fence_after_thread_sync();
        for (int h = 0; h < 2; ++h) {
            tc::tmem_ld32(tmem + lane_base + 64 + 32 * h, v);
            for (int c = 0; c < 32; ++c) p.D2[(size_t)tid * 64 + 32 * h + c] = v[c];
        }
        tc::fence_before_thread_sync();
        __syncthreads();
    }
    // ---- T3: A = the SW128 pair tile read MN-major (M = 2 atoms of 64 features, K = 64 pairs), B = Sel; D at 128..159 ----
    {
        if (tid == 0) {
            tc::fence_after_thread_sync();
            for (int s = 0; s < 4; ++s) {
                const uint64_t da = smem_desc(tc::smem_u32(smem + oA1 + s * 2048), 8192, 1024, 2);
                const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + oSel)) + 2 * s;
                tc::mma_bf16(tmem + 128, da, db, idesc_full(128, 32, 1, 0), s > 0);
            }
            tc::mma_commit(bar);
        }
        tc::mbar_wait(bar, phase); phase ^= 1;
        tc::fence_after_thread_sync();
        float v[32];
        tc::tmem_ld32(tmem + lane_base + 128, v);
        for (int c = 0; c < 32; ++c) p.D3[(size_t)tid * 32 + c] = v[c];
        tc::fence_before_thread_sync();
        __syncthreads();
    }
    // ---- T4: N = 16, K = 256: A in TMEM columns 256..383, D at 160..175 ----
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.A4 + (size_t)tid * 256);
        for (int q = 0; q < 4; ++q) {
            uint32_t r[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) r[c] = src[32 * q + c];
            tmem_st32(tmem + lane_base + 256 + 32 * q, r);
        }
        tmem_wait_st();
        tc::fence_before_thread_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_thread_sync();
            for (int s = 0; s < 16; ++s) {
                const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + oB4 + (s >> 2) * 2048)) + 2 * (s & 3);
                mma_ts(tmem + 160, tmem + 256 + 8 * s, db, idesc_full(128, 16, 0, 0), s > 0);
            }
            tc::mma_commit(bar);
        }
        tc::mbar_wait(bar, phase); phase ^= 1;
        tc::fence_after_thread_sync();
        float v[16];
        tmem_ld16(tmem + lane_base + 160, v);
        for (int c = 0; c < 16; ++c) p.D4[(size_t)tid * 16 + c] = v[c];
        tc::fence_before_thread_sync();
        __syncthreads();
    }
    // ---- T5: timings ----
    {
        float v[32], acc = 0.0f;
        __syncthreads();
        long long t0 = clock64();
        for (int it = 0; it < 16; ++it) {
#pragma unroll 1
            for (int q = 0; q < 8; ++q) {
                tc::tmem_ld32(tmem + lane_base + 256 + 32 * (q & 3), v);
                acc += v[it & 31];
            }
        }
        long long t1 = clock64();
        __syncthreads();
        if (tid == 0) p.cyc[0] = t1 - t0;   // 128 x (4 KB per warp) loads, 4 warps in parallel
        // one warp alone
        __syncthreads();
        t0 = clock64();
        if (warp == 0) {
            for (int it = 0; it < 16; ++it) {
#pragma unroll 1
                for (int q = 0; q < 8; ++q) {
                    tc::tmem_ld32(tmem + lane_base + 256 + 32 * (q & 3), v);
                    acc += v[it & 31];
                }
            }
        }
        t1 = clock64();
        if (tid == 0) p.cyc[1] = t1 - t0;
        __syncthreads();
        // stores
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(acc) + c;
        t0 = clock64();
        for (int it = 0; it < 128; ++it) tmem_st32(tmem + lane_base + 384 + 32 * (it & 3), r);
        tmem_wait_st();
        t1 = clock64();
        if (tid == 0) p.cyc[2] = t1 - t0;
        tc::fence_before_thread_sync();
        __syncthreads();
        // MMA rates (A from TMEM columns 256.., B1 / B4 tiles; results discarded)
        const int Ns[4] = {16, 64, 128, 256};
        for (int c = 0; c < 4; ++c) {
            if (tid == 0) {
                tc::fence_after_thread_sync();
                const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + oB1));
                uint32_t id = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Ns[c] >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                t0 = clock64();
                for (int s = 0; s < 64; ++s) mma_ts(tmem + 0, tmem + 256 + 8 * (s & 7), db + 2 * (s & 3), id, 1);
                tc::mma_commit(bar);
            }
            tc::mbar_wait(bar, phase); phase ^= 1;
            t1 = clock64();
            if (tid == 0) p.cyc[3 + c] = t1 - t0;
            tc::fence_after_thread_sync();
            __syncthreads();
        }
        // SS form, N = 64 (A1 tile from shared memory)
        if (tid == 0) {
            tc::fence_after_thread_sync();
            const uint64_t da = tc::smem_desc_sw128(tc::smem_u32(smem + oA1));
            const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + oB1));
            t0 = clock64();
            for (int s = 0; s < 64; ++s) tc::mma_bf16(tmem + 0, da + 2 * (s & 3), db + 2 * (s & 3), idesc_full(128, 64, 0, 0), 1);
            tc::mma_commit(bar);
        }
        tc::mbar_wait(bar, phase); phase ^= 1;
        t1 = clock64();
        if (tid == 0) p.cyc[7] = t1 - t0;

        __syncthreads();
        // independent accumulators: 4 D regions in rotation (is the small-N floor a dependent-accumulate latency?)
        for (int c = 0; c < 2; ++c) {
            const int N = c == 0 ? 16 : 64;
            if (tid == 0) {
                tc::fence_after_thread_sync();
                const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + oB1));
                uint32_t id = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                t0 = clock64();
                for (int s = 0; s < 64; ++s) mma_ts(tmem + 64 * (s & 3), tmem + 256 + 8 * (s & 7), db + 2 * (s & 3), id, 1);
                tc::mma_commit(bar);
            }
            tc::mbar_wait(bar, phase); phase ^= 1;
            t1 = clock64();
            if (tid == 0) p.cyc[8 + c] = t1 - t0;
            tc::fence_after_thread_sync();
            __syncthreads();
        }
        // same A columns every time (A re-read from the same 8 columns), N = 16
        if (tid == 0) {
            tc::fence_after_thread_sync();
            const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + oB1));
            uint32_t id = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            t0 = clock64();
            for (int s = 0; s < 64; ++s) mma_ts(tmem + 64 * (s & 3), tmem + 256, db, id, 1);
            tc::mma_commit(bar);
        }
        tc::mbar_wait(bar, phase); phase ^= 1;
        t1 = clock64();
        if (tid == 0) p.cyc[10] = t1 - t0;
        tc::fence_after_thread_sync();
        __syncthreads();
        // TMEM loads with four x32 loads in flight per wait
        {
            uint32_t q0[32], q1[32];
            t0 = clock64();
            for (int it = 0; it < 32; ++it) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(q0[0]), "=r"(q0[1]), "=r"(q0[2]), "=r"(q0[3]), "=r"(q0[4]), "=r"(q0[5]), "=r"(q0[6]), "=r"(q0[7]), "=r"(q0[8]),
                      "=r"(q0[9]), "=r"(q0[10]), "=r"(q0[11]), "=r"(q0[12]), "=r"(q0[13]), "=r"(q0[14]), "=r"(q0[15]), "=r"(q0[16]),
                      "=r"(q0[17]), "=r"(q0[18]), "=r"(q0[19]), "=r"(q0[20]), "=r"(q0[21]), "=r"(q0[22]), "=r"(q0[23]), "=r"(q0[24]),
                      "=r"(q0[25]), "=r"(q0[26]), "=r"(q0[27]), "=r"(q0[28]), "=r"(q0[29]), "=r"(q0[30]), "=r"(q0[31])
                    : "r"(tmem + lane_base + 256) : "memory");
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(q1[0]), "=r"(q1[1]), "=r"(q1[2]), "=r"(q1[3]), "=r"(q1[4]), "=r"(q1[5]), "=r"(q1[6]), "=r"(q1[7]), "=r"(q1[8]),
                      "=r"(q1[9]), "=r"(q1[10]), "=r"(q1[11]), "=r"(q1[12]), "=r"(q1[13]), "=r"(q1[14]), "=r"(q1[15]), "=r"(q1[16]),
                      "=r"(q1[17]), "=r"(q1[18]), "=r"(q1[19]), "=r"(q1[20]), "=r"(q1[21]), "=r"(q1[22]), "=r"(q1[23]), "=r"(q1[24]),
                      "=r"(q1[25]), "=r"(q1[26]), "=r"(q1[27]), "=r"(q1[28]), "=r"(q1[29]), "=r"(q1[30]), "=r"(q1[31])
                    : "r"(tmem + lane_base + 288) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c2 = 0; c2 < 32; ++c2) acc += __uint_as_float(q0[c2] ^ q1[c2]);
            }
            t1 = clock64();
            __syncthreads();
            if (tid == 0) p.cyc[11] = t1 - t0;   // 64 x 4 KB loads per warp, two in flight
        }

        __syncthreads();
        // warp-uniform issue path: one whole warp enters, an elected lane issues (no per-thread divergence handling)
        for (int c = 0; c < 3; ++c) {
            const int N = c == 0 ? 16 : (c == 1 ? 64 : 32);
            if (warp == 1) {
                uint32_t el;
                asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(el));
                tc::fence_after_thread_sync();
                const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + oB1));
                uint32_t id = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                t0 = clock64();
                if (el) {
#pragma unroll
                    for (int s = 0; s < 64; ++s) mma_ts(tmem + 0, tmem + 256 + 8 * (s & 7), db + 2 * (s & 3), id, 1);
                    tc::mma_commit(bar);
                }
                __syncwarp();
            }
            tc::mbar_wait(bar, phase); phase ^= 1;
            t1 = clock64();
            if (tid == 32) p.cyc[12 + c] = t1 - t0;
            tc::fence_after_thread_sync();
            __syncthreads();
        }
        if (acc == 123.456f) p.D1[0] = acc;
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    srand(7);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.0f - 1.0f; };
    std::vector<float> A1(128 * 64), B1(64 * 64), Ax(128 * 16), Bx(64 * 16), Sel(32 * 64), A4(128 * 256), B4(16 * 256);
    for (auto& v : A1) v = bf(rnd());
    for (auto& v : B1) v = bf(rnd());
    for (auto& v : Ax) v = bf(rnd());
    for (auto& v : Bx) v = bf(rnd());
    for (auto& v : Sel) v = bf((rand() % 3 == 0) ? 1.0f : 0.0f);
    for (auto& v : A4) v = bf(rnd());
    for (auto& v : B4) v = bf(rnd());
    auto up = [](const std::vector<float>& h) {
        std::vector<__nv_bfloat16> t(h.size());
        for (size_t i = 0; i < h.size(); ++i) t[i] = __float2bfloat16(h[i]);
        __nv_bfloat16* d;
        CK(cudaMalloc(&d, t.size() * 2));
        CK(cudaMemcpy(d, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
        return d;
    };
    Ptrs p;
    p.A1 = up(A1); p.B1 = up(B1); p.Ax = up(Ax); p.Bx = up(Bx); p.Sel = up(Sel); p.A4 = up(A4); p.B4 = up(B4);
    CK(cudaMalloc(&p.D1, 128 * 64 * 4)); CK(cudaMalloc(&p.D2, 128 * 64 * 4)); CK(cudaMalloc(&p.D3, 128 * 32 * 4));
    CK(cudaMalloc(&p.D4, 128 * 16 * 4)); CK(cudaMalloc(&p.cyc, 16 * 8));
    CK(cudaMemset(p.cyc, 0, 16 * 8));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    probe_kernel<<<1, 128, 41 * 1024>>>(p);
    CK(cudaDeviceSynchronize());
    std::vector<float> D1(128 * 64), D2(128 * 64), D3(128 * 32), D4(128 * 16);
    long long cyc[16];
    CK(cudaMemcpy(D1.data(), p.D1, D1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D2.data(), p.D2, D2.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D3.data(), p.D3, D3.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(D4.data(), p.D4, D4.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc, p.cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    double e1 = 0, e2 = 0, e3 = 0, e4 = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            double s = 0, sx = 0;
            for (int k = 0; k < 64; ++k) s += (double)A1[m * 64 + k] * B1[n * 64 + k];
            for (int k = 0; k < 16; ++k) sx += (double)Ax[m * 16 + k] * Bx[n * 16 + k];
            e1 = fmax(e1, fabs(s - D1[m * 64 + n]));
            e2 = fmax(e2, fabs(s + sx - D2[m * 64 + n]));
        }
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
            const int h = m >> 6, f = m & 63;
            double s = 0;
            for (int k = 0; k < 64; ++k) s += (double)A1[(64 * h + k) * 64 + f] * Sel[n * 64 + k];
            e3 = fmax(e3, fabs(s - D3[m * 32 + n]));
        }
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
            double s = 0;
            for (int k = 0; k < 256; ++k) s += (double)A4[m * 256 + k] * B4[n * 256 + k];
            e4 = fmax(e4, fabs(s - D4[m * 16 + n]));
        }
    printf("T1 A-from-TMEM            max err %.3e  %s\n", e1, e1 < 1e-3 ? "PASS" : "FAIL");
    printf("T2 no-swizzle K=16 block  max err %.3e  %s\n", e2, e2 < 1e-3 ? "PASS" : "FAIL");
    printf("T3 MN-major A (pair sums) max err %.3e  %s\n", e3, e3 < 1e-3 ? "PASS" : "FAIL");
    printf("T4 N=16 K=256             max err %.3e  %s\n", e4, e4 < 1e-3 ? "PASS" : "FAIL");
    printf("T5 tmem_ld32 x128, 4 warps: %lld cyc (%.1f cyc per 4 KB warp load; %.1f B/cyc/SM)\n", cyc[0], cyc[0] / 128.0, 128.0 * 4 * 4096 / cyc[0]);
    printf("T5 tmem_ld32 x128, 1 warp : %lld cyc (%.1f cyc per load; %.1f B/cyc)\n", cyc[1], cyc[1] / 128.0, 128.0 * 4096 / cyc[1]);
    printf("T5 tmem_st32 x128, 4 warps: %lld cyc (%.1f cyc per store; %.1f B/cyc/SM)\n", cyc[2], cyc[2] / 128.0, 128.0 * 4 * 4096 / cyc[2]);
    const int Ns[4] = {16, 64, 128, 256};
    for (int c = 0; c < 4; ++c) printf("T5 64 x MMA(TS) M=128 N=%3d K=16: %lld cyc (%.1f per MMA)\n", Ns[c], cyc[3 + c], cyc[3 + c] / 64.0);
    printf("T5 64 x MMA(SS) M=128 N= 64 K=16: %lld cyc (%.1f per MMA)\n", cyc[7], cyc[7] / 64.0);
    printf("T6 64 x MMA(TS) N=16, 4 independent D: %lld cyc (%.1f per MMA)\n", cyc[8], cyc[8] / 64.0);
    printf("T6 64 x MMA(TS) N=64, 4 independent D: %lld cyc (%.1f per MMA)\n", cyc[9], cyc[9] / 64.0);
    printf("T6 64 x MMA(TS) N=16, same A columns  : %lld cyc (%.1f per MMA)\n", cyc[10], cyc[10] / 64.0);
    printf("T7 tmem_ld32 x64, two in flight, 4 warps: %lld cyc (%.1f per load; %.1f B/cyc/SM)\n", cyc[11], cyc[11] / 64.0, 64.0 * 4 * 4096 / cyc[11]);
    printf("T8 64 x MMA(TS) N=16 elected lane of a uniform warp: %lld cyc (%.1f per MMA)\n", cyc[12], cyc[12] / 64.0);
    printf("T8 64 x MMA(TS) N=64 elected lane of a uniform warp: %lld cyc (%.1f per MMA)\n", cyc[13], cyc[13] / 64.0);
    printf("T8 64 x MMA(TS) N=32 elected lane of a uniform warp: %lld cyc (%.1f per MMA)\n", cyc[14], cyc[14] / 64.0);
    return (e1 < 1e-3 && e2 < 1e-3 && e3 < 1e-3 && e4 < 1e-3) ? 0 : 1;
}
