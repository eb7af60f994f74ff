// Version, error reporting and host-side FFT planning helpers of the C ABI.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "thz_fft.cuh"
#include "thz_runtime.h"

static thread_local char g_last_error[512] = "";

int thz_set_error(int code, const char* msg) {
    snprintf(g_last_error, sizeof(g_last_error), "%s", msg ? msg : "");
    return code;
}

int thz_set_cuda_error(const char* what, cudaError_t e) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what ? what : "cuda", cudaGetErrorString(e));
    return THZ_E_CUDA;
}

int thz_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

ThzDeviceGuard::ThzDeviceGuard(const void* p) : prev(-1), switched(false) {
    if (!p) return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) return;
    if (cudaGetDevice(&prev) != cudaSuccess || prev == at.device) return;
    if (cudaSetDevice(at.device) == cudaSuccess) switched = true;
}
ThzDeviceGuard::~ThzDeviceGuard() {
    if (switched) cudaSetDevice(prev);
}

extern "C" int thz_version(void) { return 200; /* 0.2.0 */ }

extern "C" const char* thz_last_error(void) { return g_last_error; }

extern "C" int thz_fft_plan_info(int32_t n, int32_t* radices, int32_t* nstages) {
    if (!radices || !nstages) return thz_set_error(THZ_E_NULL, "thz_fft_plan_info: null pointer");
    FftPlan P;
    if (thz_make_plan(n, &P) != 0) return thz_set_error(THZ_E_UNSUPPORTED, "thz_fft_plan_info: length has a prime factor > 7");
    for (int s = 0; s < THZ_MAX_STAGES; ++s) radices[s] = s < P.ns ? P.radix[s] : 0;
    *nstages = P.ns;
    return THZ_OK;
}

extern "C" int thz_fft_slot_to_bin(int32_t n, int32_t* slot_to_bin) {
    if (!slot_to_bin) return thz_set_error(THZ_E_NULL, "thz_fft_slot_to_bin: null pointer");
    FftPlan P;
    if (thz_make_plan(n, &P) != 0) return thz_set_error(THZ_E_UNSUPPORTED, "thz_fft_slot_to_bin: length has a prime factor > 7");
    for (int p = 0; p < n; ++p) slot_to_bin[p] = thz_pos_to_bin(P, p);
    return THZ_OK;
}

extern "C" int thz_fft_twiddles(int32_t n, float* tw) {
    if (!tw) return thz_set_error(THZ_E_NULL, "thz_fft_twiddles: null pointer");
    if (n < 1) return thz_set_error(THZ_E_SHAPE, "thz_fft_twiddles: n < 1");
    const double w = -2.0 * M_PI / (double)n;
    for (int m = 0; m < n; ++m) {
        tw[2 * m] = (float)cos(w * m);
        tw[2 * m + 1] = (float)sin(w * m);
    }
    return THZ_OK;
}

// ------------------------------------------------------------------------------- launch accounting
// The only mutable process-wide state of the library: a launch counter and, when enabled by
// thz_profile_enable(1), a list of CUDA event pairs bracketing each kernel (bench.py reads them to
// attribute step time to kernels on the launching stream; never enabled in normal operation).
struct ProfRec {
    cudaEvent_t a, b;
    int cls;
};
static std::atomic<uint64_t> g_launches{0};
static std::atomic<uint64_t> g_launches_cls[THZ_KC_COUNT];
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static thread_local cudaEvent_t g_open_event = nullptr;

void thz_launch_begin(cudaStream_t stream, int kernel_class) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (kernel_class >= 0 && kernel_class < THZ_KC_COUNT) g_launches_cls[kernel_class].fetch_add(1, std::memory_order_relaxed);
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    cudaEvent_t a;
    if (cudaEventCreate(&a) != cudaSuccess) return;
    cudaEventRecord(a, stream);
    g_open_event = a;
}

void thz_launch_note(int kernel_class) {
    if (kernel_class >= 0 && kernel_class < THZ_KC_COUNT) g_launches_cls[kernel_class].fetch_add(1, std::memory_order_relaxed);
}

void thz_launch_end(cudaStream_t stream, int kernel_class) {
    if (!g_prof_on.load(std::memory_order_relaxed) || !g_open_event) return;
    ProfRec r;
    r.a = g_open_event;
    g_open_event = nullptr;
    r.cls = kernel_class;
    if (cudaEventCreate(&r.b) != cudaSuccess) {
        cudaEventDestroy(r.a);
        return;
    }
    cudaEventRecord(r.b, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
}

extern "C" uint64_t thz_launch_count(void) { return g_launches.load(); }
extern "C" uint64_t thz_launch_count_class(int32_t kernel_class) {
    return (kernel_class >= 0 && kernel_class < THZ_KC_COUNT) ? g_launches_cls[kernel_class].load() : 0;
}

extern "C" int thz_profile_enable(int32_t on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    g_prof_on.store(on ? 1 : 0);
    return THZ_OK;
}

extern "C" int thz_profile_read(int32_t nclasses, float* ms_sum, int32_t* count) {
    if (!ms_sum || !count) return thz_set_error(THZ_E_NULL, "thz_profile_read: null pointer");
    for (int i = 0; i < nclasses; ++i) {
        ms_sum[i] = 0.f;
        count[i] = 0;
    }
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) {
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e != cudaSuccess) return thz_set_cuda_error("cudaEventSynchronize", e);
        float ms = 0.f;
        e = cudaEventElapsedTime(&ms, r.a, r.b);
        if (e != cudaSuccess) return thz_set_cuda_error("cudaEventElapsedTime", e);
        if (r.cls >= 0 && r.cls < nclasses) {
            ms_sum[r.cls] += ms;
            count[r.cls] += 1;
        }
    }
    return THZ_OK;
}

#include "thz_asm_p2.cuh"
extern "C" int thz_fft_is_static(int32_t n) {
    const char* off = getenv("THZ_NO_P2");
    return (off && off[0] == '1') ? 0 : (thz_sp_instantiated(n) ? 1 : 0);
}

// ------------------------------------------------------------------------------- host helper: row thresholds
#include <algorithm>
#include <vector>
extern "C" int thz_tf_row_thresholds(int32_t C, int32_t Hp, int32_t Wp, const float* rowvec, const float* colvec,
                                     const float* scal, float* tau) {
    if (C < 0 || Hp <= 0 || Wp <= 0) return thz_set_error(THZ_E_SHAPE, "thz_tf_row_thresholds: bad sizes");
    if (!rowvec || !colvec || !scal || !tau) return thz_set_error(THZ_E_NULL, "thz_tf_row_thresholds: null pointer");
    std::vector<int> order(Wp);
    std::vector<float> ky2(Wp), b1(Wp), b2(Wp);
    for (int c = 0; c < C; ++c) {
        const float* cv = colvec + (size_t)c * Wp * 4;
        const float* rv = rowvec + (size_t)c * Hp * 4;
        const float klam2 = scal[2 * c];
        for (int i = 0; i < Wp; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [cv](int a, int b) { return cv[4 * a] < cv[4 * b]; });
        for (int i = 0; i < Wp; ++i) {
            ky2[i] = cv[4 * order[i]];
            b1[i] = cv[4 * order[i] + 1];
            b2[i] = cv[4 * order[i] + 2];
        }
        for (int i = 1; i < Wp; ++i)
            if (b1[i] < b1[i - 1] || b2[i] < b2[i - 1])
                return thz_set_error(THZ_E_UNSUPPORTED, "thz_tf_row_thresholds: band-limit quotients not monotone in Ky^2");
        for (int r = 0; r < Hp; ++r) {
            const float r0 = rv[4 * r], r1 = rv[4 * r + 1], r2 = rv[4 * r + 2];
            int lo = 0, hi = Wp;                       // sorted columns [0, lo) kept, [hi, Wp) dropped
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const bool keep = thz_add_rn(r1, b1[mid]) <= 1.f && thz_add_rn(r2, b2[mid]) <= 1.f &&
                                  !(thz_sub_rn(klam2, thz_add_rn(r0, ky2[mid])) < 0.f);
                if (keep) lo = mid + 1;
                else hi = mid;
            }
            tau[(size_t)c * Hp + r] = lo > 0 ? ky2[lo - 1] : -1.f;
        }
    }
    return THZ_OK;
}
