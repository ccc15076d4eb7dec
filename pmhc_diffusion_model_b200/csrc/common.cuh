// common.cuh — error handling, launch accounting and parameter layout shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/pmhc_b200.h"

namespace pmhc {

// thread-local error text behind pmhc_last_error()
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

// Opt-in kernel attributes (dynamic shared memory size) are per DEVICE: a process that runs the library on a second GPU must set
// them there too.  One bit per device ordinal; racing first uses at worst set the attribute twice.
// Step scalars from device memory (graph-replayable training step, pmhc_train_step): while an API call runs with a device
// scalar block, the launchers hand every kernel `&block->t_over_T`; kernels read it through time_feature() and ignore the
// by-value copy.  Thread-local, set and cleared inside one API call — no state survives a call.
const float* step_t_dev();
void set_step_t_dev(const float* p);
template <class Args>
__device__ __forceinline__ float time_feature(const Args& a) { return a.t_dev != nullptr ? __ldg(a.t_dev) : a.t_over_T; }

struct PerDeviceOnce {
    std::atomic<uint64_t> done{0};
    static uint64_t bit() {
        int dev = 0;
        cudaGetDevice(&dev);
        return 1ull << (dev & 63);
    }
    bool needed() const { return (done.load(std::memory_order_acquire) & bit()) == 0; }
    void mark() { done.fetch_or(bit(), std::memory_order_release); }
};

// Optional per-kernel timing (bench.py's roofline leg): when enabled, the EGNN layer launches are bracketed by
// CUDA events on the launching stream; pmhc_profile_read() synchronises and sums them.
enum ProfileSlot { PROF_FWD = 0, PROF_BWD = 1, PROF_SLOTS = 2 };
bool profile_enabled();
void profile_mark(int slot, cudaStream_t stream, bool begin);

#define PMHC_CHECK_LAUNCH(what)                                                         \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            pmhc::set_error("%s: %s", what, cudaGetErrorString(e__));                   \
            return -2;                                                                  \
        }                                                                               \
        pmhc::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
    } while (0)

#define PMHC_REQUIRE(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            pmhc::set_error(__VA_ARGS__);    \
            return -1;                       \
        }                                    \
    } while (0)

// ---- flat parameter layout: the 48 state-dict tensors of Model(16, 22, T) in state_dict order ----
// per layer: feature_mlp.{0.w,0.b,2.w,2.b}, message_mlp.*, attention_mlp.*, translation_mlp.*, rotation_mlp.*,
// torsion_mlp.* (model.py:39-81); layer 1: H = 23, O = 64; layer 2: H = 64, O = 1 (model.py:362-371).
enum ParamId {
    FEAT0_W = 0, FEAT0_B, FEAT2_W, FEAT2_B,
    MSG0_W, MSG0_B, MSG2_W, MSG2_B,
    ATT0_W, ATT0_B, ATT2_W, ATT2_B,
    TRN0_W, TRN0_B, TRN2_W, TRN2_B,
    ROT0_W, ROT0_B, ROT2_W, ROT2_B,
    TOR0_W, TOR0_B, TOR2_W, TOR2_B,
    PARAMS_PER_LAYER
};

constexpr int kN = PMHC_N;
constexpr int kHid = PMHC_HID;
constexpr int kEdge = 2 * PMHC_N - 1;  // 31, relative position one-hot depth (model.py:349)
constexpr int kH1 = PMHC_NFEAT + 1;    // 23, node features + time (model.py:362)
constexpr int kH2 = PMHC_HID;          // 64

__host__ __device__ constexpr int layer_H(int layer) { return layer == 0 ? kH1 : kH2; }
__host__ __device__ constexpr int layer_O(int layer) { return layer == 0 ? kHid : 1; }

__host__ __device__ constexpr int param_numel(int layer, int id) {
    const int H = layer_H(layer), O = layer_O(layer);
    switch (id) {
        case FEAT0_W: return kHid * (H + kHid);
        case FEAT0_B: return kHid;
        case FEAT2_W: return O * kHid;
        case FEAT2_B: return O;
        case MSG0_W: return kHid * (2 * H + kEdge);
        case MSG0_B: return kHid;
        case MSG2_W: return kHid * kHid;
        case MSG2_B: return kHid;
        case ATT0_W: return kHid * (kHid + 2);
        case ATT0_B: return kHid;
        case ATT2_W: return kHid;
        case ATT2_B: return 1;
        case TRN0_W: return kHid * kHid;
        case TRN0_B: return kHid;
        case TRN2_W: return kHid;
        case TRN2_B: return 1;
        case ROT0_W: return kHid * (kHid + 4);
        case ROT0_B: return kHid;
        case ROT2_W: return 4 * kHid;
        case ROT2_B: return 4;
        case TOR0_W: return kHid * (kHid + 2 * PMHC_NTORS);
        case TOR0_B: return kHid;
        case TOR2_W: return PMHC_NTORS * kHid;
        case TOR2_B: return PMHC_NTORS;
        default: return 0;
    }
}

__host__ __device__ constexpr int param_offset(int layer, int id) {
    int off = 0;
    for (int l = 0; l < layer; ++l)
        for (int i = 0; i < PARAMS_PER_LAYER; ++i) off += param_numel(l, i);
    for (int i = 0; i < id; ++i) off += param_numel(layer, i);
    return off;
}

static_assert(param_offset(2, 0) == PMHC_NPARAM, "flat parameter count must match the reference state dict");

}  // namespace pmhc
