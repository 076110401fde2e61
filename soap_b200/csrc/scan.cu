// scan.cu -- device-wide exclusive scan of uint32 counts (three launches).
// Used for SharedMesh.cell_offset (SOAP/core/shared_mesh.py:96-102) and for the
// internal fine-bin / record offsets of the halo pipeline.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v,
                                                                   unsigned long long* total,
                                                                   unsigned long long* warp_sums) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned long long w = (lane < (blockDim.x >> 5)) ? warp_sums[lane] : 0ull;
        unsigned long long winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (blockDim.x >> 5)) warp_sums[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    unsigned long long res = warp_sums[wid] + inc - v;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t* __restrict__ in,
                                                                 int64_t n,
                                                                 unsigned long long* __restrict__ sums) {
    __shared__ unsigned long long ws[32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    s = warp_sum_u64(s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long w = (threadIdx.x < (SCAN_THREADS >> 5)) ? ws[threadIdx.x] : 0ull;
        w = warp_sum_u64(w);
        if (threadIdx.x == 0) sums[blockIdx.x] = w;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(unsigned long long* sums, int64_t nb,
                                                            unsigned long long* total_out) {
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long tot;
    unsigned long long carry = 0;
    for (int64_t b0 = 0; b0 < nb; b0 += SCAN_THREADS) {
        int64_t i = b0 + threadIdx.x;
        unsigned long long v = (i < nb) ? sums[i] : 0ull;
        unsigned long long ex = block_exclusive_scan(v, &tot, ws);
        if (i < nb) sums[i] = carry + ex;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* __restrict__ in,
                                                             int64_t n,
                                                             const unsigned long long* __restrict__ sums,
                                                             uint32_t* __restrict__ out_u32,
                                                             int64_t* __restrict__ out_i64) {
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long tot;
    // each thread owns SCAN_ITEMS consecutive elements
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        v[k] = (i < n) ? in[i] : 0u;
        s += v[k];
    }
    unsigned long long ex = block_exclusive_scan(s, &tot, ws) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        if (i < n) {
            if (out_u32) out_u32[i] = (uint32_t)ex;
            if (out_i64) out_i64[i] = (int64_t)ex;
        }
        ex += v[k];
    }
}

}  // namespace

// out_u32 / out_i64: either may be NULL.  in and out_u32 may alias.
// total_dev (optional) receives the grand total (uint64).
int soap_exclusive_scan_u32(soap_handle* h, const uint32_t* in, uint32_t* out_u32,
                            int64_t* out_i64, int64_t n, uint64_t* total_dev,
                            cudaStream_t stream) {
    if (n <= 0) {
        if (total_dev) CUDA_TRY(cudaMemsetAsync(total_dev, 0, sizeof(uint64_t), stream));
        return 0;
    }
    int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    WS_GET(sums, unsigned long long, h, "scan_sums", nb);
    LAUNCH(h, k_scan_tile_sums, (unsigned)nb, SCAN_THREADS, 0, stream, in, n, sums);
    LAUNCH(h, k_scan_sums, 1, SCAN_THREADS, 0, stream, sums, nb,
           (unsigned long long*)total_dev);
    LAUNCH(h, k_scan_apply, (unsigned)nb, SCAN_THREADS, 0, stream, in, n, sums, out_u32,
           out_i64);
    return 0;
}
