// capi.cu — error text, device probe, parameter layout queries and the sampling trajectory driver.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "egnn_common.cuh"

namespace pmhc {

static thread_local char g_error[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

static thread_local const float* g_step_t_dev = nullptr;
const float* step_t_dev() { return g_step_t_dev; }
void set_step_t_dev(const float* p) { g_step_t_dev = p; }

static bool g_profile = false;
static std::mutex g_profile_mu;
static std::vector<cudaEvent_t> g_prof_events[PROF_SLOTS];  // begin, end, begin, end, ...

bool profile_enabled() { return g_profile; }
void profile_mark(int slot, cudaStream_t stream, bool begin) {
    std::lock_guard<std::mutex> lk(g_profile_mu);
    if ((g_prof_events[slot].size() % 2 == 0) != begin) return;  // keep begin/end paired
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, stream);
    g_prof_events[slot].push_back(ev);
}

int model_forward_impl(const float* params, const PmhcBatch* bt, float t_over_T, float* out_frames, float* out_torsions,
                       float* saved, void* workspace, size_t workspace_bytes, cudaStream_t stream, int precision,
                       bool reuse_pocket_cache);
int launch_reverse_step_philox(const float* zf, const float* zt, const float* pf, const float* pt, uint64_t seed,
                               uint64_t first, double beta_t, double beta_s, int64_t n, const float* sign_ref, float* of,
                               float* ot, cudaStream_t stream, const uint64_t* seed_first_dev);
int launch_remove_noise(const float* zf, const float* zt, const float* pf, const float* pt, const float* xf,
                        const float* xt, double beta_t, double beta_s, int64_t n, const float* sign_ref, float* of,
                        float* ot, cudaStream_t stream);
__global__ void unpack_noise_tape_kernel(const float* __restrict__ tape, int64_t n, float* __restrict__ frames,
                                         float* __restrict__ tors) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * 21) return;
    int64_t r = idx / 21;
    int c = (int)(idx - r * 21);
    if (c < 7) frames[r * 7 + c] = tape[idx];
    else tors[r * 14 + (c - 7)] = tape[idx];
}

}  // namespace pmhc

using namespace pmhc;

extern "C" const char* pmhc_last_error(void) { return g_error; }

extern "C" int64_t pmhc_launch_count(void) { return g_launches.load(); }

// a replayed CUDA graph launches kernels the library never sees: the caller adds the number it counted at capture
extern "C" void pmhc_launch_count_add(int64_t n) { g_launches.fetch_add(n); }

extern "C" void pmhc_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_profile_mu);
    g_profile = on != 0;
}

extern "C" int pmhc_profile_read(double* ms, int64_t* launches) {
    std::lock_guard<std::mutex> lk(g_profile_mu);
    for (int s = 0; s < PROF_SLOTS; ++s) {
        double total = 0.0;
        int64_t n = 0;
        auto& ev = g_prof_events[s];
        for (size_t k = 0; k + 1 < ev.size(); k += 2) {
            cudaEventSynchronize(ev[k + 1]);
            float t = 0.0f;
            if (cudaEventElapsedTime(&t, ev[k], ev[k + 1]) == cudaSuccess) {
                total += t;
                ++n;
            }
        }
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        ev.clear();
        ms[s] = total;
        launches[s] = n;
    }
    return 0;
}

extern "C" int pmhc_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    PMHC_REQUIRE(e == cudaSuccess, "cudaGetDevice: %s", cudaGetErrorString(e));
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    PMHC_REQUIRE(major == 10 && minor == 0, "this library is built for sm_100a only; device is sm_%d%d", major, minor);
    PMHC_REQUIRE(device_props() == 0, "could not query the device");
    return 0;
}

extern "C" int64_t pmhc_param_offset(int index) {
    if (index < 0 || index >= 2 * PARAMS_PER_LAYER) return -1;
    return param_offset(index / PARAMS_PER_LAYER, index % PARAMS_PER_LAYER);
}

extern "C" int64_t pmhc_param_numel(int index) {
    if (index < 0 || index >= 2 * PARAMS_PER_LAYER) return -1;
    return param_numel(index / PARAMS_PER_LAYER, index % PARAMS_PER_LAYER);
}

// DiffusionModelOptimizer.sample (optimizer.py:226-252): t = T..1: z <- remove_noise(z, model(z, t), t, t-1).
extern "C" int pmhc_sample(const float* params, const PmhcBatch* bt, float* frames, float* torsions, int T,
                           double beta_min, double beta_max, uint64_t seed, uint64_t first_complex,
                           const float* noise_tape, const float* quat_sign_tape, float* scratch, void* workspace,
                           size_t workspace_bytes, void* stream_, int precision) {
    return pmhc_sample_ex(params, bt, frames, torsions, T, beta_min, beta_max, seed, first_complex, nullptr, noise_tape, quat_sign_tape,
                          scratch, workspace, workspace_bytes, stream_, precision);
}

// seed_first_dev (nullable, 2 x uint64 in device memory: seed, first complex): read by the reverse-step kernels instead of
// the by-value pair, so a CUDA graph captured around this call (400+ launches) replays with whatever the block holds then.
extern "C" int pmhc_sample_ex(const float* params, const PmhcBatch* bt, float* frames, float* torsions, int T,
                              double beta_min, double beta_max, uint64_t seed, uint64_t first_complex,
                              const uint64_t* seed_first_dev, const float* noise_tape, const float* quat_sign_tape,
                              float* scratch, void* workspace, size_t workspace_bytes, void* stream_, int precision) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PMHC_REQUIRE(bt != nullptr && bt->B > 0 && T > 0, "pmhc_sample: empty batch or T <= 0");
    PMHC_REQUIRE(scratch != nullptr, "pmhc_sample: scratch is required");
    const int64_t n = (int64_t)bt->B * kN;
    float* pred_f = scratch;
    float* pred_t = pred_f + n * 7;
    float* fresh_f = pred_t + n * 14;
    float* fresh_t = fresh_f + n * 7;
    PmhcBatch step = *bt;
    step.frames = frames;
    step.torsions = torsions;
    for (int t = T; t > 0; --t) {
        const int k = T - t;
        // the pocket side of the first-layer projections is identical for all T steps: computed at the first step only
        int rc = model_forward_impl(params, &step, (float)((double)t / (double)T), pred_f, pred_t, nullptr, workspace,
                                    workspace_bytes, stream, precision, /*reuse_pocket_cache=*/t != T);
        if (rc != 0) return rc;
        // linear_schedule (optimizer.py:20-21) in double, exactly as the reference's Python floats
        double beta_t = beta_min + (beta_max - beta_min) * ((double)t / (double)T);
        double beta_s = beta_min + (beta_max - beta_min) * ((double)(t - 1) / (double)T);
        const float* sign = quat_sign_tape ? quat_sign_tape + (size_t)k * n * 4 : nullptr;
        if (noise_tape != nullptr) {
            unpack_noise_tape_kernel<<<(unsigned)((n * 21 + 255) / 256), 256, 0, stream>>>(noise_tape + (size_t)k * n * 21, n, fresh_f, fresh_t);
            PMHC_CHECK_LAUNCH("unpack_noise_tape");
            rc = launch_remove_noise(frames, torsions, pred_f, pred_t, fresh_f, fresh_t, beta_t, beta_s, n, sign, frames, torsions, stream);
        } else {
            // fused draw + reverse step; one Philox stream per (global residue index, step): the key mixes seed and step
            const uint64_t step_key = 0x9E3779B97F4A7C15ull * (uint64_t)(k + 1);
            rc = launch_reverse_step_philox(frames, torsions, pred_f, pred_t, (seed_first_dev ? 0ull : seed) + step_key,
                                            first_complex * kN, beta_t, beta_s, n, sign, frames, torsions, stream, seed_first_dev);
        }
        if (rc != 0) return rc;
    }
    return 0;
}
