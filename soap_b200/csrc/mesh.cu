// mesh.cu -- stage A (cell-list build) and the API-level stage B sphere query.
//
//   soap_box_wrap      <- box_wrap,               SOAP/core/chunk_tasks.py:48-50
//   soap_mesh_build    <- SharedMesh.__init__,    SOAP/core/shared_mesh.py:11-114
//   soap_sphere_query  <- query_radius_periodic,  SOAP/core/shared_mesh.py:122-200
//
// The build is a counting sort keyed by cell id: exact min/max -> cell id +
// histogram (the atomic's return value is the particle's rank in its cell) ->
// exclusive scan -> scatter.  HBM-bound integer work; no tensor cores.
#include "common.cuh"
#include "radix.cuh"

thread_local char g_soap_err[512] = "";

namespace {

constexpr int TB = 256;

// ------------------------------------------------------------------ box wrap
__global__ void __launch_bounds__(TB) k_box_wrap(double* __restrict__ pos, int64_t n3, double sx,
                                                 double sy, double sz, double L) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n3; i += stride) {
        int d = (int)(i % 3);
        double shift = d == 0 ? sx : (d == 1 ? sy : sz);
        pos[i] = __dadd_rn(floored_mod(__dsub_rn(pos[i], shift), L), shift);
    }
}

// -------------------------------------------------------------------- bounds
__global__ void __launch_bounds__(TB) k_bounds_partial(const double* __restrict__ pos, int64_t n,
                                                       double* __restrict__ partial) {
    double mn[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
    double mx[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
        mn[0] = fmin(mn[0], x); mx[0] = fmax(mx[0], x);
        mn[1] = fmin(mn[1], y); mx[1] = fmax(mx[1], y);
        mn[2] = fmin(mn[2], z); mx[2] = fmax(mx[2], z);
    }
    __shared__ double s[6][TB / 32];
    for (int d = 0; d < 3; d++) {
        double a = warp_min(mn[d]), b = warp_max(mx[d]);
        if ((threadIdx.x & 31) == 0) {
            s[d][threadIdx.x >> 5] = a;
            s[3 + d][threadIdx.x >> 5] = b;
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = s[threadIdx.x][0];
        for (int w = 1; w < TB / 32; w++)
            v = threadIdx.x < 3 ? fmin(v, s[threadIdx.x][w]) : fmax(v, s[threadIdx.x][w]);
        partial[(int64_t)blockIdx.x * 6 + threadIdx.x] = v;
    }
}

__global__ void k_bounds_final(const double* __restrict__ partial, int nb, double* __restrict__ out) {
    // one warp per output component
    int d = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (d >= 6) return;
    double v = d < 3 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    for (int b = lane; b < nb; b += 32) {
        double p = partial[(int64_t)b * 6 + d];
        v = d < 3 ? fmin(v, p) : fmax(v, p);
    }
    v = d < 3 ? warp_min(v) : warp_max(v);
    if (lane == 0) out[d] = v;
}

// ----------------------------------------------------- cell id + histogram
struct MeshGeom {
    double pmin[3], cs[3];
    int res;
};

__global__ void __launch_bounds__(TB) k_cell_key(const double* __restrict__ pos, int64_t n, MeshGeom g,
                                                 int32_t* __restrict__ cell_idx, uint32_t* __restrict__ key,
                                                 uint32_t* __restrict__ val) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = cell_coord(pos[3 * i], g.pmin[0], g.cs[0], g.res);
    int cy = cell_coord(pos[3 * i + 1], g.pmin[1], g.cs[1], g.res);
    int cz = cell_coord(pos[3 * i + 2], g.pmin[2], g.cs[2], g.res);
    const int32_t c = cx + g.res * cy + g.res * g.res * cz;  // shared_mesh.py:73-77
    if (cell_idx) cell_idx[i] = c;
    key[i] = (uint32_t)c;
    val[i] = (uint32_t)i;
}

// cell_offset[c] = first sorted position with key >= c; cell_count from consecutive offsets
// (np.bincount + exclusive cumsum of shared_mesh.py:80-102)
__global__ void __launch_bounds__(TB) k_cell_bounds(const uint32_t* __restrict__ key, uint32_t n, uint32_t ncell,
                                                    int64_t* __restrict__ offset_ext) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const long long prev = i == 0 ? -1ll : (long long)key[i - 1];
    const long long cur = i == n ? (long long)ncell : (long long)key[i];
    for (long long c = prev + 1; c <= cur; c++) offset_ext[c] = i;
}

__global__ void __launch_bounds__(TB) k_cell_counts(const int64_t* __restrict__ offset_ext, int64_t ncell,
                                                    int64_t* __restrict__ count, int64_t* __restrict__ offset) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    offset[c] = offset_ext[c];
    count[c] = offset_ext[c + 1] - offset_ext[c];
}

__global__ void __launch_bounds__(TB) k_widen_idx(const uint32_t* __restrict__ in, int64_t n,
                                                  int64_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// ------------------------------------------------------------- sphere query
struct QueryGeom {
    double pmin[3], pmax[3], cs[3];
    double L;
    int res;
};

template <bool FILL>
__global__ void __launch_bounds__(TB) k_sphere_query(
    const double* __restrict__ pos, QueryGeom g, const int64_t* __restrict__ cell_count,
    const int64_t* __restrict__ cell_offset, const int64_t* __restrict__ sort_idx,
    const double* __restrict__ centres, const double* __restrict__ radii,
    int64_t* __restrict__ counts, const int64_t* __restrict__ offsets, int64_t* __restrict__ idx_out,
    const float* __restrict__ mass, double* __restrict__ enclosed) {
    __shared__ DimRanges rg[3];
    __shared__ unsigned wcount[TB / 32];
    __shared__ unsigned long long base_s;
    __shared__ double msum[TB / 32];
    const int64_t q = blockIdx.x;
    const double cx = centres[3 * q], cy = centres[3 * q + 1], cz = centres[3 * q + 2];
    const double r = radii[q];
    const double r2max = __dmul_rn(r, r);
    const double halfL = 0.5 * g.L;
    if (threadIdx.x < 3) {
        double c = threadIdx.x == 0 ? cx : (threadIdx.x == 1 ? cy : cz);
        dim_ranges(c, r, g.L, g.pmin[threadIdx.x], g.pmax[threadIdx.x], g.cs[threadIdx.x], g.res,
                   rg[threadIdx.x]);
    }
    if (threadIdx.x == 0) base_s = FILL ? (unsigned long long)offsets[q] : 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long my_count = 0;
    double my_mass = 0.0;
    for (int rz = 0; rz < rg[2].n; rz++)
        for (int k = rg[2].lo[rz]; k <= rg[2].hi[rz]; k++)
            for (int ry = 0; ry < rg[1].n; ry++)
                for (int j = rg[1].lo[ry]; j <= rg[1].hi[ry]; j++)
                    for (int rx = 0; rx < rg[0].n; rx++) {
                        // a run of cells along i is one contiguous span of sort_idx
                        int64_t c0 = rg[0].lo[rx] + (int64_t)g.res * j + (int64_t)g.res * g.res * k;
                        int64_t c1 = c0 + (rg[0].hi[rx] - rg[0].lo[rx]);
                        int64_t s0 = cell_offset[c0];
                        int64_t s1 = cell_offset[c1] + cell_count[c1];
                        for (int64_t b = s0; b < s1; b += TB) {
                            int64_t t = b + threadIdx.x;
                            bool keep = false;
                            int64_t pi = -1;
                            if (t < s1) {
                                pi = sort_idx[t];
                                double r2 = periodic_r2(pos[3 * pi], pos[3 * pi + 1],
                                                        pos[3 * pi + 2], cx, cy, cz, g.L, halfL);
                                keep = r2 <= r2max;
                            }
                            if (!FILL) {
                                if (keep) {
                                    my_count++;
                                    if (mass) my_mass += (double)mass[pi];
                                }
                            } else {
                                unsigned bal = __ballot_sync(0xffffffffu, keep);
                                if (lane == 0) wcount[wid] = __popc(bal);
                                __syncthreads();
                                unsigned pre = 0, tot = 0;
                                for (int w = 0; w < TB / 32; w++) {
                                    if (w < wid) pre += wcount[w];
                                    tot += wcount[w];
                                }
                                if (keep)
                                    idx_out[base_s + pre + __popc(bal & ((1u << lane) - 1u))] = pi;
                                __syncthreads();
                                if (threadIdx.x == 0) base_s += tot;
                                __syncthreads();
                            }
                        }
                    }
    if (!FILL) {
        my_count = warp_sum_u64(my_count);
        my_mass = warp_sum(my_mass);
        __shared__ unsigned long long cs_[TB / 32];
        if (lane == 0) { cs_[wid] = my_count; msum[wid] = my_mass; }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long c = 0;
            double m = 0.0;
            for (int w = 0; w < TB / 32; w++) { c += cs_[w]; m += msum[w]; }
            counts[q] = (int64_t)c;
            if (enclosed) enclosed[q] = m;
        }
    }
}

}  // namespace

// Exact bounding box + the equal-coordinate guard (shared_mesh.py:35-66).
// Leaves pmin/pmax/cs on the host; syncs the stream.
int soap_mesh_bounds(soap_handle* h, const double* pos_dev, int64_t n, int resolution,
                     double pos_min[3], double pos_max[3], double cell_size[3],
                     cudaStream_t stream, bool guard) {
    int nb = h->sm_count * 8;
    WS_GET(partial, double, h, "bounds_partial", (size_t)nb * 6 + 6);
    double* out = partial + (size_t)nb * 6;
    LAUNCH(h, k_bounds_partial, nb, TB, 0, stream, pos_dev, n, partial);
    LAUNCH(h, k_bounds_final, 1, 192, 0, stream, partial, nb, out);
    double hb[6];
    CUDA_TRY(cudaMemcpyAsync(hb, out, sizeof(hb), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    for (int d = 0; d < 3; d++) {
        pos_min[d] = hb[d];
        pos_max[d] = hb[3 + d];
        if (guard && pos_min[d] == pos_max[d]) pos_max[d] = pos_min[d] + 1.0;  // shared_mesh.py:56-58
        cell_size[d] = (pos_max[d] - pos_min[d]) / resolution;        // shared_mesh.py:66
    }
    return 0;
}

// =================================================================== C ABI
extern "C" {

int soap_abi_version(void) { return SOAP_B200_ABI_VERSION; }
const char* soap_last_error(void) { return g_soap_err; }

int soap_create(int device, soap_handle** out) {
    if (!out) SOAP_FAIL("soap_create: out is NULL");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) SOAP_FAIL("soap_create: no CUDA device %d (have %d)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        SOAP_FAIL("soap_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                  device, prop.major, prop.minor);
    // keep freed pool memory cached (chunks are created and destroyed every step)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    soap_handle* h = new soap_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    *out = h;
    return 0;
}

int soap_destroy(soap_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    for (auto& kv : h->ws)
        if (kv.second.p) cudaFree(kv.second.p);
    h->streams_destroy();
    delete h;
    return 0;
}

int64_t soap_launch_count(const soap_handle* h) { return h ? h->launches : 0; }

int soap_kernel_timing(soap_handle* h, int on) {
    if (!h) SOAP_FAIL("soap_kernel_timing: NULL handle");
    h->ktime = on != 0;
    return 0;
}

int64_t soap_kernel_timings(soap_handle* h, char* buf, int64_t buflen) {
    if (!h || !buf || buflen <= 0) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    h->klog.collect();
    std::string s;
    char line[512];
    for (auto& kv : h->klog.acc) {
        snprintf(line, sizeof(line), "%s\t%lld\t%.6f\n", kv.first.c_str(), (long long)kv.second.first, kv.second.second);
        s += line;
    }
    h->klog.acc.clear();
    const int64_t m = (int64_t)s.size() < buflen - 1 ? (int64_t)s.size() : buflen - 1;
    memcpy(buf, s.data(), m);
    buf[m] = 0;
    return m;
}

int soap_box_wrap(soap_handle* h, double* pos_dev, int64_t n, const double ref_pos[3],
                  double boxsize, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!h) SOAP_FAIL("soap_box_wrap: NULL handle");
    if (n <= 0) return 0;
    CUDA_TRY(cudaSetDevice(h->device));
    // shift = ref_pos - 0.5 * boxsize (chunk_tasks.py:49)
    double sx = ref_pos[0] - 0.5 * boxsize, sy = ref_pos[1] - 0.5 * boxsize,
           sz = ref_pos[2] - 0.5 * boxsize;
    LAUNCH(h, k_box_wrap, grid_for(3 * n, TB, h->sm_count * 16), TB, 0, stream, pos_dev, 3 * n, sx,
           sy, sz, boxsize);
    return 0;
}

int soap_mesh_build(soap_handle* h, const double* pos_dev, int64_t n, int resolution,
                    double pos_min[3], double pos_max[3], double cell_size[3],
                    int32_t* cell_idx_dev, int64_t* cell_count_dev, int64_t* cell_offset_dev,
                    int64_t* sort_idx_dev, int stable, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!h) SOAP_FAIL("soap_mesh_build: NULL handle");
    if (n <= 0) SOAP_FAIL("soap_mesh_build: empty particle set (SharedMesh.empty: shared_mesh.py:25-29)");
    if (n >= (1ll << 32)) SOAP_FAIL("soap_mesh_build: n=%lld exceeds the 2^32 particle limit", (long long)n);
    if (resolution < 1 || resolution > 1024) SOAP_FAIL("soap_mesh_build: resolution %d outside [1,1024]", resolution);
    CUDA_TRY(cudaSetDevice(h->device));
    const int64_t ncell = (int64_t)resolution * resolution * resolution;
    if (soap_mesh_bounds(h, pos_dev, n, resolution, pos_min, pos_max, cell_size, stream, true)) return -1;
    MeshGeom g;
    for (int d = 0; d < 3; d++) { g.pmin[d] = pos_min[d]; g.cs[d] = cell_size[d]; }
    g.res = resolution;
    // argsort of the cell ids (shared_mesh.py:105-114): stable LSD radix sort of (cell id, index)
    // pairs, so sort_idx is ascending within a cell whether or not `stable` is asked for
    (void)stable;
    const uint32_t nblk = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    WS_GET(key, uint32_t, h, "mesh_key", n);
    WS_GET(val, uint32_t, h, "mesh_val", n);
    WS_GET(key2, uint32_t, h, "mesh_key2", n);
    WS_GET(val2, uint32_t, h, "mesh_val2", n);
    WS_GET(ghist, uint32_t, h, "mesh_ghist", (size_t)RS_NB * nblk);
    WS_GET(offset_ext, int64_t, h, "mesh_offset_ext", ncell + 1);
    LAUNCH(h, k_cell_key, grid_for(n, TB), TB, 0, stream, pos_dev, n, g, cell_idx_dev, key, val);
    int bits = 0;
    while ((1ll << bits) < ncell) bits++;
    uint32_t *ks = nullptr, *vs = nullptr;
    if (radix_sort_pairs(h, key, val, key2, val2, ghist, (uint32_t)n, bits, &ks, &vs, stream)) return -1;
    LAUNCH(h, k_cell_bounds, grid_for(n + 1, TB), TB, 0, stream, ks, (uint32_t)n, (uint32_t)ncell, offset_ext);
    LAUNCH(h, k_cell_counts, grid_for(ncell, TB), TB, 0, stream, offset_ext, ncell, cell_count_dev, cell_offset_dev);
    LAUNCH(h, k_widen_idx, grid_for(n, TB), TB, 0, stream, vs, n, sort_idx_dev);
    return 0;
}

int soap_sphere_query(soap_handle* h, const double* pos_dev, int64_t n, int resolution,
                      const double pos_min[3], const double pos_max[3], const double cell_size[3],
                      const int64_t* cell_count_dev, const int64_t* cell_offset_dev,
                      const int64_t* sort_idx_dev, const double* centres_dev,
                      const double* radii_dev, int64_t n_query, double boxsize,
                      int64_t* counts_dev, const int64_t* offsets_dev, int64_t* idx_dev,
                      const float* mass_dev, double* enclosed_mass_dev, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!h) SOAP_FAIL("soap_sphere_query: NULL handle");
    if (n_query <= 0) return 0;
    (void)n;
    CUDA_TRY(cudaSetDevice(h->device));
    QueryGeom g;
    for (int d = 0; d < 3; d++) { g.pmin[d] = pos_min[d]; g.pmax[d] = pos_max[d]; g.cs[d] = cell_size[d]; }
    g.L = boxsize;
    g.res = resolution;
    if (!idx_dev) {
        if (!counts_dev) SOAP_FAIL("soap_sphere_query: counts_dev is NULL in the count pass");
        LAUNCH(h, k_sphere_query<false>, (unsigned)n_query, TB, 0, stream, pos_dev, g, cell_count_dev,
               cell_offset_dev, sort_idx_dev, centres_dev, radii_dev, counts_dev, offsets_dev,
               idx_dev, mass_dev, enclosed_mass_dev);
    } else {
        if (!offsets_dev) SOAP_FAIL("soap_sphere_query: offsets_dev is NULL in the fill pass");
        LAUNCH(h, k_sphere_query<true>, (unsigned)n_query, TB, 0, stream, pos_dev, g, cell_count_dev,
               cell_offset_dev, sort_idx_dev, centres_dev, radii_dev, counts_dev, offsets_dev,
               idx_dev, mass_dev, enclosed_mass_dev);
    }
    return 0;
}

}  // extern "C"
