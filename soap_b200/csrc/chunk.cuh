// chunk.cuh -- device-resident particle chunk shared by chunk.cu and halos.cu.
#pragma once
#include "common.cuh"

// Type codes used on the device: PartType0/1/4/5 -> 0/1/2/3.
__host__ __device__ inline int ptype_code(int ptype) {
    return ptype == 0 ? 0 : (ptype == 1 ? 1 : (ptype == 4 ? 2 : 3));
}

// Device view of the chunk: all particle types merged, SoA, in the cell order
// of an internal fine mesh (cell id = i + res*j + res^2*k, so a run of cells
// along i is one contiguous particle span).
struct ChunkView {
    int64_t n;
    double L;
    int res;
    double pmin[3], pmax[3], cs[3];
    const uint32_t* cell_off;  // [res^3 + 1]
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
    uint32_t* orig = nullptr;  // [n] index within the particle's own ptype array
    int64_t last_pairs = 0;
    int64_t last_small_pairs = 0;  // of which handled by the fused small-halo tiers
    int64_t last_tier_pairs[3] = {0, 0, 0};
    int64_t last_candidates = 0;
    int64_t last_count_pairs = 0, last_try_pairs = 0, last_mom_pairs = 0;
    int last_rounds = 0;
    PhaseLog create_log, halo_log;
    int type_present[4] = {0, 0, 0, 0};
};
