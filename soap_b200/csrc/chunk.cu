// chunk.cu -- build the device-resident chunk for the batched halo path.
//
// Replaces the per-ptype SharedMesh construction of
// SOAP/core/chunk_tasks.py:299-304 for soap_process_halos: all particle types
// are merged into one SoA particle set, binned by an internal fine mesh
// (counting sort keyed by cell id, same arithmetic as shared_mesh.py:69-77) and
// physically reordered into cell order, so that the sphere gather reads
// contiguous, coalesced spans.  Membership never depends on the mesh: it is
// decided by the exact periodic r2 test of shared_mesh.py:138-142,192.
#include "chunk.cuh"
#include "radix.cuh"

int soap_mesh_bounds(soap_handle* h, const double* pos_dev, int64_t n, int resolution,
                     double pos_min[3], double pos_max[3], double cell_size[3],
                     cudaStream_t stream, bool guard);

namespace {

constexpr int TB = 256;

struct Geom {
    double pmin[3], cs[3];
    int res, nb;
};

__global__ void __launch_bounds__(TB) k_cell_keys(const double* __restrict__ pos, int64_t n, Geom g,
                                                  uint32_t* __restrict__ key, uint32_t* __restrict__ val,
                                                  uint32_t base) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = cell_coord(pos[3 * i], g.pmin[0], g.cs[0], g.res);
    int cy = cell_coord(pos[3 * i + 1], g.pmin[1], g.cs[1], g.res);
    int cz = cell_coord(pos[3 * i + 2], g.pmin[2], g.cs[2], g.res);
    key[i] = blocked_cell(cx, cy, cz, g.nb);
    val[i] = base + (uint32_t)i;
}

// cell_off[c] = first sorted position whose key is >= c (cell_off[ncell] = n)
__global__ void __launch_bounds__(TB) k_cell_offsets(const uint32_t* __restrict__ key, uint32_t n, uint32_t ncell,
                                                     uint32_t* __restrict__ cell_off) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const long long prev = i == 0 ? -1ll : (long long)key[i - 1];
    const long long cur = i == n ? (long long)ncell : (long long)key[i];
    for (long long c = prev + 1; c <= cur; c++) cell_off[c] = i;
}

struct TypeIn {
    const double* pos;
    const float* mass;
    const float* vel;
    const void* grnr;
    const void* fof;
    uint32_t base, n;
    int id64;
    uint8_t tcode;
};
struct TypesIn {
    TypeIn t[4];
    int n;
};

struct SoAOut {
    double *px, *py, *pz;
    float *mass, *vx, *vy, *vz;
    int32_t *grnr, *fof;
    uint8_t* type;
};

// cell-ordered SoA: slot d takes particle perm[d] (coalesced writes, gathered reads)
// ids_bad is set when a 64-bit group id does not fit the 32 bits kept on the device
__global__ void __launch_bounds__(TB) k_gather(TypesIn in, const uint32_t* __restrict__ perm, uint32_t n, SoAOut o,
                                               int* __restrict__ ids_bad) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    const uint32_t src = perm[d];
    SOAP_ASSERT(src < n);
    int ti = 0;
#pragma unroll
    for (int k = 1; k < 4; k++)
        if (k < in.n && src >= in.t[k].base) ti = k;
    const TypeIn& T = in.t[ti];
    const uint32_t i = src - T.base;
    o.px[d] = T.pos[3 * (size_t)i];
    o.py[d] = T.pos[3 * (size_t)i + 1];
    o.pz[d] = T.pos[3 * (size_t)i + 2];
    o.mass[d] = T.mass[i];
    o.vx[d] = T.vel[3 * (size_t)i];
    o.vy[d] = T.vel[3 * (size_t)i + 1];
    o.vz[d] = T.vel[3 * (size_t)i + 2];
    if (T.id64) {
        const int64_t g64 = ((const int64_t*)T.grnr)[i], f64 = ((const int64_t*)T.fof)[i];
        if (g64 != (int64_t)(int32_t)g64 || f64 != (int64_t)(int32_t)f64) *ids_bad = 1;
        o.grnr[d] = (int32_t)g64;
        o.fof[d] = (int32_t)f64;
    } else {
        o.grnr[d] = ((const int32_t*)T.grnr)[i];
        o.fof[d] = ((const int32_t*)T.fof)[i];
    }
    o.type[d] = T.tcode;
}

template <typename T>
int dev_alloc(soap_chunk* c, T** p, size_t count) {
    void* q = nullptr;
    // stream-ordered pool allocation: after the first chunk no driver malloc/free per step
    CUDA_TRY(cudaMallocAsync(&q, sizeof(T) * (count ? count : 1), c->stream));
    c->owned.push_back(q);
    *p = (T*)q;
    return 0;
}

}  // namespace

extern "C" {

int soap_chunk_create(soap_handle* h, const soap_ptype_arrays* types, int n_types,
                      double boxsize, int fine_ppc, soap_chunk** out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!h || !types || !out) SOAP_FAIL("soap_chunk_create: NULL argument");
    if (n_types < 1 || n_types > 4) SOAP_FAIL("soap_chunk_create: n_types=%d outside [1,4]", n_types);
    CUDA_TRY(cudaSetDevice(h->device));
    int64_t n = 0;
    for (int t = 0; t < n_types; t++) {
        if (types[t].n < 0) SOAP_FAIL("soap_chunk_create: negative particle count");
        if (types[t].ptype != 0 && types[t].ptype != 1 && types[t].ptype != 4 && types[t].ptype != 5)
            SOAP_FAIL("soap_chunk_create: ptype %d not in {0,1,4,5}", types[t].ptype);
        n += types[t].n;
    }
    if (n <= 0) SOAP_FAIL("soap_chunk_create: no particles");
    if (n >= (1ll << 32) - 2) SOAP_FAIL("soap_chunk_create: %lld particles exceed the 2^32 limit", (long long)n);
    soap_chunk* c = new soap_chunk();
    c->h = h;
    c->stream = stream;
    c->create_log.reset();
    c->create_log.begin("mesh_bounds", stream);
    // exact bounding box over all types (shared_mesh.py:35-58 per ptype, merged)
    double pmin[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
    double pmax[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
    for (int t = 0; t < n_types; t++) {
        if (types[t].n == 0) continue;
        double a[3], b[3], cs[3];
        if (soap_mesh_bounds(h, types[t].pos, types[t].n, 1, a, b, cs, stream, false)) { delete c; return -1; }
        for (int d = 0; d < 3; d++) {
            pmin[d] = a[d] < pmin[d] ? a[d] : pmin[d];
            pmax[d] = b[d] > pmax[d] ? b[d] : pmax[d];
        }
    }
    c->create_log.end(stream);
    if (fine_ppc <= 0) fine_ppc = 8;
    int res = (int)cbrt((double)n / (double)fine_ppc);
    if (res < 1) res = 1;
    if (res > 512) res = 512;
    // 256^3 blocked cell ids fit 24 bits = three 8-bit radix passes; a slightly finer mesh would pay a fourth pass over
    // every (key, index) pair for less than twice the particles per cell
    if (res > 256 && res <= 320) res = 256;
    ChunkView& v = c->v;
    v.n = n;
    v.L = boxsize;
    v.res = res;
    const int nb = (res + BLK - 1) / BLK;
    v.nb = nb;
    for (int d = 0; d < 3; d++) {
        if (pmin[d] == pmax[d]) pmax[d] = pmin[d] + 1.0;
        v.pmin[d] = pmin[d];
        v.pmax[d] = pmax[d];
        v.cs[d] = (pmax[d] - pmin[d]) / res;
    }
    const int64_t ncell = (int64_t)nb * nb * nb * BLK_CELLS;  // blocked cell ids (ids of cells beyond res stay empty)
    SoAOut o;
    uint32_t* cell_off = nullptr;
    int rc = 0;
    rc |= dev_alloc(c, &cell_off, ncell + 1);
    rc |= dev_alloc(c, &o.px, n); rc |= dev_alloc(c, &o.py, n); rc |= dev_alloc(c, &o.pz, n);
    rc |= dev_alloc(c, &o.mass, n);
    rc |= dev_alloc(c, &o.vx, n); rc |= dev_alloc(c, &o.vy, n); rc |= dev_alloc(c, &o.vz, n);
    rc |= dev_alloc(c, &o.grnr, n); rc |= dev_alloc(c, &o.fof, n);
    rc |= dev_alloc(c, &o.type, n);
    if (rc) { soap_chunk_destroy(c); return -1; }
    uint32_t* key = (uint32_t*)h->get("chunk_key", sizeof(uint32_t) * (size_t)n);
    uint32_t* val = (uint32_t*)h->get("chunk_val", sizeof(uint32_t) * (size_t)n);
    uint32_t* key2 = (uint32_t*)h->get("chunk_key2", sizeof(uint32_t) * (size_t)n);
    uint32_t* val2 = (uint32_t*)h->get("chunk_val2", sizeof(uint32_t) * (size_t)n);
    const uint32_t nblk = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    uint32_t* ghist = (uint32_t*)h->get("chunk_ghist", sizeof(uint32_t) * (size_t)RS_NB * nblk);
    if (!key || !val || !key2 || !val2 || !ghist) { soap_chunk_destroy(c); return -1; }
    Geom g;
    for (int d = 0; d < 3; d++) { g.pmin[d] = v.pmin[d]; g.cs[d] = v.cs[d]; }
    g.res = res;
    g.nb = nb;
#define CK(stmt) do { if ((stmt) != 0) { soap_chunk_destroy(c); return -1; } } while (0)
#define CKL(...) do { auto _f = [&]() -> int { __VA_ARGS__; return 0; }; if (_f() != 0) { soap_chunk_destroy(c); return -1; } } while (0)
    c->create_log.begin("mesh_keys", stream);
    TypesIn tin;
    tin.n = 0;
    int64_t base = 0;
    for (int t = 0; t < n_types; t++) {
        if (types[t].n == 0) continue;
        CKL(LAUNCH(h, k_cell_keys, grid_for(types[t].n, TB), TB, 0, stream, types[t].pos, types[t].n, g, key + base,
                   val + base, (uint32_t)base));
        uint8_t tc = (uint8_t)ptype_code(types[t].ptype);
        c->type_present[tc] = 1;
        TypeIn& T = tin.t[tin.n++];
        T.pos = types[t].pos; T.mass = types[t].mass; T.vel = types[t].vel;
        T.grnr = types[t].grnr; T.fof = types[t].fof;
        T.base = (uint32_t)base; T.n = (uint32_t)types[t].n; T.id64 = types[t].ids_are_int64; T.tcode = tc;
        base += types[t].n;
    }
    c->create_log.end(stream);
    c->create_log.begin("mesh_sort", stream);
    {
        int bits = 0;
        while ((1ll << bits) < ncell) bits++;
        CK(radix_sort_pairs(h, key, val, key2, val2, ghist, (uint32_t)n, bits, &key, &val, stream));
    }
    CKL(LAUNCH(h, k_cell_offsets, grid_for(n + 1, TB), TB, 0, stream, key, (uint32_t)n, (uint32_t)ncell, cell_off));
    c->create_log.end(stream);
    c->create_log.begin("reorder", stream);
    int* ids_bad = (int*)h->get("chunk_ids_bad", sizeof(int));
    if (!ids_bad) { soap_chunk_destroy(c); return -1; }
    CK(cudaMemsetAsync(ids_bad, 0, sizeof(int), stream) != cudaSuccess);
    CKL(LAUNCH(h, k_gather, grid_for(n, TB), TB, 0, stream, tin, val, (uint32_t)n, o, ids_bad));
    int ids_bad_host = 0;
    CK(cudaMemcpyAsync(&ids_bad_host, ids_bad, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess);
    c->create_log.end(stream);
#undef CK
#undef CKL
    v.cell_off = cell_off;
    v.px = o.px; v.py = o.py; v.pz = o.pz;
    v.mass = o.mass; v.vx = o.vx; v.vy = o.vy; v.vz = o.vz;
    v.grnr = o.grnr; v.fof = o.fof; v.type = o.type;
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        snprintf(g_soap_err, sizeof(g_soap_err), "soap_chunk_create: %s", cudaGetErrorString(e));
        soap_chunk_destroy(c);
        return -1;
    }
    c->create_log.collect();
    if (ids_bad_host) {
        // membership is an integer-exact contract: ids that differ by a multiple of 2^32 must not alias
        snprintf(g_soap_err, sizeof(g_soap_err),
                 "soap_chunk_create: a GroupNr_bound / FOFGroupIDs value does not fit in 32 bits (the device keeps "
                 "group ids as int32; re-index the groups of this chunk)");
        soap_chunk_destroy(c);
        return -1;
    }
    *out = c;
    return 0;
}

int soap_chunk_destroy(soap_chunk* c) {
    if (!c) return 0;
    for (void* p : c->owned) cudaFreeAsync(p, c->stream);
    delete c;
    return 0;
}

int64_t soap_chunk_num_particles(const soap_chunk* c) { return c ? c->v.n : 0; }
int64_t soap_chunk_last_pairs(const soap_chunk* c) { return c ? c->last_pairs : 0; }

int64_t soap_chunk_timings(const soap_chunk* c, char* buf, int64_t buflen) {
    if (!c || !buf || buflen <= 0) return 0;
    std::string s;
    char line[512];
    for (auto& kv : c->create_log.ms) {
        snprintf(line, sizeof(line), "create/%s:%.6f\n", kv.first.c_str(), kv.second);
        s += line;
    }
    for (auto& kv : c->halo_log.ms) {
        snprintf(line, sizeof(line), "halos/%s:%.6f\n", kv.first.c_str(), kv.second);
        s += line;
    }
    snprintf(line, sizeof(line),
             "stat/pairs:%lld\nstat/candidates:%lld\nstat/count_pairs:%lld\nstat/try_pairs:%lld\n"
             "stat/moment_pairs:%lld\nstat/rounds:%d\nstat/res:%d\nstat/small_pairs:%lld\nstat/small_pairs_0:%lld\nstat/small_pairs_1:%lld\nstat/small_pairs_2:%lld\n",
             (long long)c->last_pairs, (long long)c->last_candidates, (long long)c->last_count_pairs,
             (long long)c->last_try_pairs, (long long)c->last_mom_pairs, c->last_rounds, c->v.res,
             (long long)c->last_small_pairs, (long long)c->last_tier_pairs[0], (long long)c->last_tier_pairs[1],
             (long long)c->last_tier_pairs[2]);
    s += line;
    snprintf(line, sizeof(line), "stat/rec_seq:%lld\nstat/rec_cta:%lld\nstat/rec_cluster8:%lld\nstat/rec_cluster16:%lld\n",
             (long long)c->last_rec_class[0], (long long)c->last_rec_class[1], (long long)c->last_rec_class[2],
             (long long)c->last_rec_class[3]);
    s += line;
    int64_t m = (int64_t)s.size() < buflen - 1 ? (int64_t)s.size() : buflen - 1;
    memcpy(buf, s.data(), m);
    buf[m] = 0;
    return m;
}

}  // extern "C"
