// chunk.cuh -- device-resident particle chunk shared by chunk.cu and halos.cu.
#pragma once
#include "common.cuh"

// Type codes used on the device: PartType0/1/4/5 -> 0/1/2/3.
__host__ __device__ inline int ptype_code(int ptype) {
    return ptype == 0 ? 0 : (ptype == 1 ? 1 : (ptype == 4 ? 2 : 3));
}

// Cells of the internal mesh are numbered block by block: blocks of 8 x 8 x 8 cells in (x, y, z)
// order, cells inside a block in (x, y, z) order.  A run of cells along x inside one block is one
// contiguous particle span, and the particles of a block (a few thousand, ~200 KB of payload) sit
// together in memory: the reorder pass reads the input, which arrives in SWIFT top-level cell order
// (swift_cells.py:551-737), one slab at a time out of L2 instead of striding across the whole
// z-plane, and a sphere's rows in neighbouring y / z are neighbours in memory too.
constexpr int BLK = 8, BLK_SHIFT = 3, BLK_CELLS = BLK * BLK * BLK;
__host__ __device__ inline uint32_t blocked_cell(int i, int j, int k, int nb) {
    return (((uint32_t)(k >> BLK_SHIFT) * (uint32_t)nb + (uint32_t)(j >> BLK_SHIFT)) * (uint32_t)nb +
            (uint32_t)(i >> BLK_SHIFT)) * (uint32_t)BLK_CELLS +
           (uint32_t)((((k & (BLK - 1)) << BLK_SHIFT) + (j & (BLK - 1))) << BLK_SHIFT) + (uint32_t)(i & (BLK - 1));
}

// Device view of the chunk: all particle types merged, SoA, in the blocked cell order
// of an internal fine mesh.
struct ChunkView {
    int64_t n;
    double L;
    int res;
    int nb;  // blocks per dimension = ceil(res / BLK)
    double pmin[3], pmax[3], cs[3];
    const uint32_t* cell_off;  // [nb^3 * BLK_CELLS + 1], indexed by blocked_cell()
    const double *px, *py, *pz;
    const float *mass, *vx, *vy, *vz;
    const int32_t *grnr, *fof;
    const uint8_t* type;  // type code 0..3
};

// Event-timed phases of one API call (cudaEvents on the launching stream).
struct PhaseLog {
    struct Span { std::string name; cudaEvent_t e0, e1; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    std::map<std::string, float> ms;  // accumulated after collect()
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
    void reset() { ms.clear(); }
    // call after the stream has been synchronised
    void collect() {
        for (auto& sp : spans) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, sp.e0, sp.e1) == cudaSuccess) ms[sp.name] += t;
            else cudaGetLastError();  // an unfinished span must not poison the next launch check
            pool.push_back(sp.e0);
            pool.push_back(sp.e1);
        }
        spans.clear();
    }
    ~PhaseLog() {
        for (auto& sp : spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
        for (auto e : pool) cudaEventDestroy(e);
    }
};

struct soap_chunk {
    soap_handle* h = nullptr;
    cudaStream_t stream = nullptr;
    ChunkView v{};
    std::vector<void*> owned;  // device allocations freed at destroy
    int64_t last_pairs = 0;
    int64_t last_small_pairs = 0;  // of which handled by the fused small-halo tiers
    int64_t last_tier_pairs[3] = {0, 0, 0};
    int64_t last_candidates = 0;
    int64_t last_count_pairs = 0, last_try_pairs = 0, last_mom_pairs = 0;
    int64_t last_rec_class[4] = {0, 0, 0, 0};
    int last_rounds = 0;
    PhaseLog create_log, halo_log;
    int type_present[4] = {0, 0, 0, 0};
};
