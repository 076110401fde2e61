// common.cuh -- shared host/device helpers of libsoap_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/soap_b200.h"

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ < 1000
#error "libsoap_b200 is written for sm_100a (B200) only"
#endif

extern thread_local char g_soap_err[512];

#define SOAP_FAIL(...)                                         \
    do {                                                       \
        snprintf(g_soap_err, sizeof(g_soap_err), __VA_ARGS__); \
        return -1;                                             \
    } while (0)

#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess)                                                      \
            SOAP_FAIL("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

// One growable scratch buffer per name; freed with the handle.
struct WsBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct PhaseTimer {
    std::string name;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0.f;
};

// Per-kernel device time (CUDA events on the launching stream around every launch), switched on by
// soap_kernel_timing(): what bench.py's roofline line is computed from.
struct KernelLog {
    struct Span { const char* name; cudaEvent_t e0, e1; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    std::map<std::string, std::pair<int64_t, double>> acc;  // name -> (launches, ms)
    cudaEvent_t ev() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    void begin(const char* name, cudaStream_t s) {
        Span sp{name, ev(), ev()};
        cudaEventRecord(sp.e0, s);
        spans.push_back(sp);
    }
    void end(cudaStream_t s) { cudaEventRecord(spans.back().e1, s); }
    // call after the streams have been synchronised
    void collect() {
        for (auto& sp : spans) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, sp.e0, sp.e1) == cudaSuccess) {
                auto& a = acc[sp.name];
                a.first += 1;
                a.second += t;
            } else {
                cudaGetLastError();
            }
            pool.push_back(sp.e0);
            pool.push_back(sp.e1);
        }
        spans.clear();
    }
    ~KernelLog() {
        for (auto& sp : spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
        for (auto e : pool) cudaEventDestroy(e);
    }
};

struct soap_handle {
    int device = 0;
    int sm_count = 148;
    int64_t launches = 0;
    bool ktime = false;
    KernelLog klog;
    // side streams for kernels of one phase that work on disjoint halos (fork / join around them with events)
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    int side_init() {
        if (side[0]) return 0;
        for (int i = 0; i < 3; i++) {
            if (cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking) != cudaSuccess) return -1;
            if (cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming) != cudaSuccess) return -1;
        }
        return cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) == cudaSuccess ? 0 : -1;
    }
    // the small-halo tiers run on their own two streams, concurrently with the general path's first round
    static constexpr int TIER_EV = 8;
    cudaStream_t tstream[2] = {nullptr, nullptr};
    cudaEvent_t ev_tfork = nullptr, ev_tjoin[2] = {nullptr, nullptr}, ev_tround[TIER_EV] = {};
    int tier_init() {
        if (tstream[0]) return 0;
        for (int i = 0; i < 2; i++) {
            if (cudaStreamCreateWithFlags(&tstream[i], cudaStreamNonBlocking) != cudaSuccess) return -1;
            if (cudaEventCreateWithFlags(&ev_tjoin[i], cudaEventDisableTiming) != cudaSuccess) return -1;
        }
        for (int i = 0; i < TIER_EV; i++)
            if (cudaEventCreateWithFlags(&ev_tround[i], cudaEventDisableTiming) != cudaSuccess) return -1;
        return cudaEventCreateWithFlags(&ev_tfork, cudaEventDisableTiming) == cudaSuccess ? 0 : -1;
    }
    void streams_destroy() {
        for (int i = 0; i < 3; i++) {
            if (side[i]) cudaStreamDestroy(side[i]);
            if (ev_join[i]) cudaEventDestroy(ev_join[i]);
        }
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (int i = 0; i < 2; i++) {
            if (tstream[i]) cudaStreamDestroy(tstream[i]);
            if (ev_tjoin[i]) cudaEventDestroy(ev_tjoin[i]);
        }
        for (int i = 0; i < TIER_EV; i++)
            if (ev_tround[i]) cudaEventDestroy(ev_tround[i]);
        if (ev_tfork) cudaEventDestroy(ev_tfork);
    }
    std::map<std::string, WsBuf> ws;
    // returns nullptr on failure (error string set)
    void* get(const char* name, size_t bytes) {
        WsBuf& b = ws[name];
        if (b.bytes >= bytes && b.p) return b.p;
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.bytes = 0;
        size_t want = bytes + bytes / 4 + 256;  // grow with slack: steady state has no allocs
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            // retry without slack
            want = bytes;
            e = cudaMalloc(&b.p, want);
        }
        if (e != cudaSuccess) {
            snprintf(g_soap_err, sizeof(g_soap_err), "workspace '%s': cudaMalloc(%zu) failed: %s",
                     name, want, cudaGetErrorString(e));
            b.p = nullptr;
            return nullptr;
        }
        b.bytes = want;
        return b.p;
    }
    void release(const char* name) {
        auto it = ws.find(name);
        if (it != ws.end()) {
            if (it->second.p) cudaFree(it->second.p);
            ws.erase(it);
        }
    }
};

#define WS_GET(var, type, handle, name, count)                               \
    type* var = (type*)(handle)->get(name, sizeof(type) * (size_t)(count)); \
    if (!var) return -1;

#define LAUNCH_N(h, name, kernel, grid, block, smem, stream, ...)                 \
    do {                                                                          \
        if ((h)->ktime) (h)->klog.begin(name, stream);                            \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);               \
        if ((h)->ktime) (h)->klog.end(stream);                                    \
        (h)->launches++;                                                          \
        cudaError_t _e = cudaGetLastError();                                      \
        if (_e != cudaSuccess)                                                    \
            SOAP_FAIL("%s:%d launch %s -> %s", __FILE__, __LINE__, name,          \
                      cudaGetErrorString(_e));                                    \
    } while (0)
#define LAUNCH(h, kernel, grid, block, smem, stream, ...) \
    LAUNCH_N(h, #kernel, kernel, grid, block, smem, stream, __VA_ARGS__)

static inline unsigned grid_for(int64_t n, int block, int64_t cap = (1 << 30)) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (unsigned)g;
}

// ------------------------------------------------------------------ device
#ifdef __CUDACC__

// Device-side assertions of the checked build (make checked): an index that leaves its array traps the kernel
// and names the line, instead of corrupting a neighbour silently.  Compiled out of the product build.
#ifdef SOAP_CHECKS
#define SOAP_ASSERT(cond)                                                              \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            printf("SOAP_ASSERT failed: %s:%d: %s\n", __FILE__, __LINE__, #cond);      \
            __trap();                                                                  \
        }                                                                              \
    } while (0)
#else
#define SOAP_ASSERT(cond) ((void)0)
#endif

// numpy floored modulo for positive divisor L (npy_divmod): fmod, then shift
// negative remainders by L (the sum is rounded, so tiny negatives give L).
static __device__ __noinline__ double floored_mod_generic(double a, double L) {
    double m = fmod(a, L);
    if (m != 0.0) {
        if (m < 0.0) m = __dadd_rn(m, L);
    } else {
        m = 0.0;  // copysign(0, L) with L > 0
    }
    return m;
}
__device__ __forceinline__ double floored_mod(double a, double L) {
    // fast paths, bit-identical to the generic branch: for |a| < L fmod returns a
    // itself, and for L <= a < 2L it returns a - L, which is exact (Sterbenz)
    if (a > 0.0 && a < L) return a;
    if (a < 0.0 && a > -L) return __dadd_rn(a, L);
    if (a >= L && a < __dadd_rn(L, L)) return __dsub_rn(a, L);
    return floored_mod_generic(a, L);
}

// periodic minimum-image displacement of SharedMesh.query_radius_periodic
// (SOAP/core/shared_mesh.py:138-142): r2 = (dx^2 + dy^2) + dz^2, no FMA.
__device__ __forceinline__ double periodic_r2(double px, double py, double pz, double cx,
                                              double cy, double cz, double L, double halfL) {
    double dx = __dsub_rn(px, cx), dy = __dsub_rn(py, cy), dz = __dsub_rn(pz, cz);
    if (dx > halfL) dx = __dsub_rn(dx, L);
    if (dx < -halfL) dx = __dadd_rn(dx, L);
    if (dy > halfL) dy = __dsub_rn(dy, L);
    if (dy < -halfL) dy = __dadd_rn(dy, L);
    if (dz > halfL) dz = __dsub_rn(dz, L);
    if (dz < -halfL) dz = __dadd_rn(dz, L);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// halo-centred re-wrap (SOAP/core/halo_tasks.py:113-117) followed by the
// "pos - centre" of every compute_basics (e.g. SO_properties.py:336).
__device__ __forceinline__ double rewrap_rel(double p, double c, double L, double halfL) {
    double offset = __dsub_rn(c, halfL);
    double w = __dadd_rn(floored_mod(__dsub_rn(p, offset), L), offset);
    return __dsub_rn(w, c);
}

__device__ __forceinline__ double radius3(double x, double y, double z) {
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

// cell coordinate of SharedMesh.__init__ (SOAP/core/shared_mesh.py:69-72)
__device__ __forceinline__ int cell_coord(double p, double pmin, double cs, int res) {
    double f = floor(__ddiv_rn(__dsub_rn(p, pmin), cs));
    int c = (f < 0.0) ? 0 : ((f >= (double)res) ? res - 1 : (int)f);
    return c;
}

// Per-dimension cell ranges overlapped by [c-r, c+r] and its periodic copies
// (SOAP/core/shared_mesh.py:146-179).  Ranges come out ascending and disjoint.
// A relative pad makes the candidate set a superset under rounding; membership
// is always decided by the exact r2 test.
#define SOAP_MAX_RANGES 4
struct DimRanges {
    int n;
    int lo[SOAP_MAX_RANGES], hi[SOAP_MAX_RANGES];
};

__device__ inline void dim_ranges(double c, double r, double L, double pmin, double pmax,
                                  double cs, int res, DimRanges& out) {
    out.n = 0;
    int min_copy = 0, max_copy = 0;
    while (c + (min_copy - 1) * L + r >= pmin && min_copy > -16) min_copy--;
    while (c + (max_copy + 1) * L - r <= pmax && max_copy < 16) max_copy++;
    double pad = 1e-12 * (fabs(c) + fabs(r) + L + fabs(pmin) + fabs(pmax));
    for (int k = min_copy; k <= max_copy; k++) {
        double a = fmax(pmin, c + k * L - r - pad);
        double b = fmin(pmax, c + k * L + r + pad);
        if (b < a) continue;
        int ilo = (int)floor((a - pmin) / cs);
        int ihi = (int)floor((b - pmin) / cs);
        if (ilo < 0) ilo = 0;
        if (ihi > res - 1) ihi = res - 1;
        if (ilo > ihi) continue;
        if (out.n > 0 && ilo <= out.hi[out.n - 1] + 1) {
            // overlaps / touches the previous range: merge
            if (ihi > out.hi[out.n - 1]) out.hi[out.n - 1] = ihi;
            if (ilo < out.lo[out.n - 1]) out.lo[out.n - 1] = ilo;  // cannot happen (ascending)
        } else if (out.n < SOAP_MAX_RANGES) {
            out.lo[out.n] = ilo;
            out.hi[out.n] = ihi;
            out.n++;
        } else {
            out.hi[out.n - 1] = ihi;  // degenerate: swallow the gap
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// In-place ascending bitonic sort of a[0..n) by the whole CTA, any n (the
// network for the next power of two with every comparison ascending, so the
// virtual +inf padding at indices >= n never moves).  Works on shared or
// global memory; callers must have made a[] visible (__syncthreads) before.
template <typename T, typename Less>
__device__ inline void block_bitonic_sort(T* a, uint32_t n, Less less) {
    uint32_t np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (uint32_t k = 2; k <= np2; k <<= 1) {
        for (uint32_t t = threadIdx.x; t < np2; t += blockDim.x) {
            uint32_t p = t ^ (k - 1);
            if (p > t && p < n) {
                T x = a[t], y = a[p];
                if (less(y, x)) { a[t] = y; a[p] = x; }
            }
        }
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < np2; t += blockDim.x) {
                uint32_t p = t ^ j;
                if (p > t && p < n) {
                    T x = a[t], y = a[p];
                    if (less(y, x)) { a[t] = y; a[p] = x; }
                }
            }
            __syncthreads();
        }
    }
}

#endif  // __CUDACC__

// exclusive scan of u32 counts (device-wide), see scan.cu
int soap_exclusive_scan_u32(soap_handle* h, const uint32_t* in, uint32_t* out_u32,
                            int64_t* out_i64, int64_t n, uint64_t* total_dev,
                            cudaStream_t stream);
