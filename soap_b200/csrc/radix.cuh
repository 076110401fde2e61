// radix.cuh -- LSD radix sort of (uint32 key, uint32 value) pairs, 8 bits per pass.
// Shared by the chunk build (chunk.cu) and soap_mesh_build (mesh.cu).
#pragma once
#include "common.cuh"

#ifdef __CUDACC__
// The cell-order permutation is an LSD radix sort of (cell id, particle index)
// pairs, 8 bits per pass (3 passes for the 2^24 cells of a 256^3 mesh): per
// pass a per-block digit histogram, one device-wide exclusive scan, and a
// stable scatter (warp-level multi-split with match_any, no atomics).  Unlike
// a counting sort with one returning atomic per particle on 16 M random
// counters, every pass reads and writes its 8-byte pairs in tile-coherent runs.
constexpr int RS_TB = 256, RS_IPT = 16, RS_TILE = RS_TB * RS_IPT, RS_NB = 256;

static __global__ void __launch_bounds__(RS_TB) k_rs_hist(const uint32_t* __restrict__ key, uint32_t n, int shift,
                                                   uint32_t nblk, uint32_t* __restrict__ ghist) {
    __shared__ uint32_t hist[RS_NB];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_IPT; k++) {
        const uint32_t i = base + k * RS_TB + threadIdx.x;
        if (i < n) atomicAdd(&hist[(key[i] >> shift) & (RS_NB - 1)], 1u);
    }
    __syncthreads();
    ghist[(size_t)threadIdx.x * nblk + blockIdx.x] = hist[threadIdx.x];  // digit-major: one scan gives global offsets
}

static __global__ void __launch_bounds__(RS_TB, 4) k_rs_scatter(const uint32_t* __restrict__ key_in,
                                                         const uint32_t* __restrict__ val_in, uint32_t n, int shift,
                                                         uint32_t nblk, const uint32_t* __restrict__ goff,
                                                         uint32_t* __restrict__ key_out,
                                                         uint32_t* __restrict__ val_out) {
    constexpr int NW = RS_TB / 32;
    __shared__ uint32_t wcount[NW][RS_NB];  // running count of each digit within a warp's rows
    __shared__ uint32_t wbase[NW][RS_NB];   // exclusive prefix over the warps + the block's global offset
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NW * RS_NB; i += RS_TB) (&wcount[0][0])[i] = 0;
    __syncthreads();
    // warp w owns the contiguous elements [w * 32 * IPT, (w + 1) * 32 * IPT) of the tile, row by row.
    // Only the 16-bit ranks stay in registers across the barrier (two per register): keys and values
    // are read again from L1 / L2 for the scatter, which keeps four CTAs resident per SM.
    const uint32_t wbeg = blockIdx.x * RS_TILE + wid * (32 * RS_IPT);
    uint32_t rk2[RS_IPT / 2];
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
        const uint32_t i = wbeg + r * 32 + lane;
        const bool ok = i < n;
        const uint32_t k = ok ? key_in[i] : 0xffffffffu;
        const uint32_t d = ok ? ((k >> shift) & (RS_NB - 1)) : RS_NB;  // invalid lanes match each other only
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (ok && lane == leader) {
            old = wcount[wid][d];
            wcount[wid][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        const uint32_t rk = old + __popc(peers & ((1u << lane) - 1u));  // < 32 * RS_IPT
        if (r & 1) rk2[r >> 1] |= rk << 16; else rk2[r >> 1] = rk;
        __syncwarp();
    }
    __syncthreads();
    {
        const int d = threadIdx.x;  // RS_TB == RS_NB
        uint32_t run = goff[(size_t)d * nblk + blockIdx.x];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            wbase[w][d] = run;
            run += wcount[w][d];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
        const uint32_t i = wbeg + r * 32 + lane;
        if (i < n) {
            const uint32_t k = key_in[i], vv = val_in[i];
            const uint32_t d = (k >> shift) & (RS_NB - 1);
            const uint32_t dst = wbase[wid][d] + ((rk2[r >> 1] >> ((r & 1) * 16)) & 0xffffu);
            key_out[dst] = k;
            val_out[dst] = vv;
        }
    }
}


// Sort n pairs by the low `bits` bits of the key (stable).  key/val and key2/val2 are ping-pong
// buffers of n elements, ghist holds RS_NB * ceil(n / RS_TILE) counters.  Returns (through
// key_out/val_out) the buffers that hold the sorted result.
static inline int radix_sort_pairs(soap_handle* h, uint32_t* key, uint32_t* val, uint32_t* key2, uint32_t* val2,
                                   uint32_t* ghist, uint32_t n, int bits, uint32_t** key_out, uint32_t** val_out,
                                   cudaStream_t stream) {
    const uint32_t nblk = (n + RS_TILE - 1) / RS_TILE;
    uint32_t *ki = key, *vi = val, *ko = key2, *vo = val2;
    for (int shift = 0; shift < bits; shift += 8) {
        LAUNCH(h, k_rs_hist, nblk, RS_TB, 0, stream, ki, n, shift, nblk, ghist);
        if (soap_exclusive_scan_u32(h, ghist, ghist, nullptr, (int64_t)RS_NB * nblk, nullptr, stream)) return -1;
        LAUNCH(h, k_rs_scatter, nblk, RS_TB, 0, stream, ki, vi, n, shift, nblk, ghist, ko, vo);
        uint32_t* t1 = ki; ki = ko; ko = t1;
        uint32_t* t2 = vi; vi = vo; vo = t2;
    }
    *key_out = ki;
    *val_out = vi;
    return 0;
}
#endif  // __CUDACC__
