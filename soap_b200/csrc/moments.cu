// moments.cu -- masked moment sums of every accepted halo and its result row.
//
// Every selection of the reference's property classes is a radial cut on the
// halo-centred radius, optionally restricted to bound particles and/or a
// particle type:
//   SO                r <  R_SO          all particles (SO_properties.py:485-489)
//   Exclusive sphere  r <= R_ap, bound   (aperture_properties.py:285-288,310)
//   Inclusive sphere  r <= R_ap
//   BoundSubhalo      bound, no cut      (subhalo_properties.py:144)
// The cuts of one halo are nested, so a particle belongs to exactly one
// (shell, bound, type) bank; it is added there once and every selection is a
// prefix sum over shells at the end.  Raw moments about the halo centre are
// accumulated in float64 and the reference's central moments (velocity
// dispersion about vcom, L about vcom, kinetic energy with Hubble flow) are
// derived from them:
//   com, vcom     aperture_properties.py:1098-1126, SO_properties.py:557-571
//   L             kinematic_properties.py:222-263
//   veldisp       kinematic_properties.py:91-127
//   Ekin          subhalo_properties.py:848-858
//   spin          SO_properties.py:602-618, subhalo_properties.py:1049-1073
//   concentration SO_properties.py:2724-2790
//   tensors       inertia_tensors.py:19-132 with max_iterations=1
#include "moments.cuh"

namespace {

constexpr int TB = SWEEP_NT;

template <int V, int NTY>
__global__ void __launch_bounds__(TB, (V <= 16 && NTY == 1) ? 3 : 1) k_moments(ChunkView v, HaloArrays ha, DevCfg cfg,
                                                const Item* __restrict__ items,
                                                const unsigned int* __restrict__ n_items_dev,
                                                int gbank_stride, int priv) {
    // dynamic shared memory: [priv ? NW : 1][gbank_stride] banks, then the
    // per-warp staging tiles [NW][32][VP] and their keys [NW][32]
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = TB / 32;
    constexpr int VP = BankAcc<V>::VP;
    constexpr int NCOPY = BankAcc<V>::HALF ? 2 * NW : NW;  // private bank copies of the CTA (priv)
    double* banks = (double*)smem_raw;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* stage_w = banks + (size_t)(priv ? NCOPY : 1) * gbank_stride + (size_t)wid * 32 * VP;
    int* skey_w = (int*)(banks + (size_t)(priv ? NCOPY : 1) * gbank_stride + (size_t)NW * 32 * VP) + wid * 32;
    double* bank_w = banks + (priv ? (size_t)wid * (NCOPY / NW) * gbank_stride : 0);
    __shared__ SweepShared SW;
    __shared__ Cuts cuts;
    const unsigned int n_items = *n_items_dev;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item im = items[it];
        const uint32_t h = im.halo;
        // properties [lo, hi) of halo_prop_list were computed at this rung
        const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
        if (c_hi <= c_lo || ha.status[h] >= 2) continue;
        const ScanRes* sr = ha.sres + h;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const double R = ha.rung_r[h];
        const double r2max = __dmul_rn(R, R), halfL = 0.5 * v.L, L = v.L;
        const int64_t hidx = ha.index[h];
        const bool central = ha.central[h] == 1;
        const int n_so = central ? cfg.n_so : 0;
        __syncthreads();
        if (threadIdx.x == 32) {
            build_cuts(cuts, cfg, sr, c_lo, c_hi, n_so);
            if (im.k == 0) ha.cuts[h] = cuts;  // for k_rows
        }
        __syncthreads();
        const int ncut = cuts.n;
        const int nbank = (ncut + 1) * 2 * NTY;
        if (priv) {
            for (int w = 0; w < NCOPY; w++)
                for (int i = threadIdx.x; i < nbank * V; i += TB) banks[(size_t)w * gbank_stride + i] = 0.0;
        } else {
            for (int i = threadIdx.x; i < nbank * V; i += TB) banks[i] = 0.0;
        }
        __syncthreads();
        const int32_t cen_fof = sr->cen_fof;
        BankAcc<V> ba;
        ba.init(gbank_stride);
        sweep_item<SW_MASS | SW_VEL | SW_IDS | SW_TYPE>(v, SW, cx, cy, cz, R, im, [&](uint32_t t, bool ok) {
            bool in = false;
            int key = 0;
            double val[V];
            if (ok) {
                double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                in = r2 <= r2max;
            }
            if (in) {
                const double x = rewrap_rel(v.px[t], cx, L, halfL);
                const double y = rewrap_rel(v.py[t], cy, L, halfL);
                const double z = rewrap_rel(v.pz[t], cz, L, halfL);
                const double r = radius3(x, y, z);
                key = moment_terms<V, NTY>(cuts, ncut, cfg, x, y, z, r, (double)v.mass[t], (double)v.vx[t],
                                           (double)v.vy[t], (double)v.vz[t], v.grnr[t], hidx, v.fof[t], cen_fof,
                                           NTY == 1 ? 1u : (uint32_t)v.type[t], val);
            }
            SOAP_ASSERT(!in || (key >= 0 && key < nbank));
            ba.add(in, key, val, stage_w, skey_w, bank_w, priv, lane);
        });
        ba.flush(bank_w, priv, lane);
        __syncthreads();
        if (priv) {
            for (int i = threadIdx.x; i < nbank * V; i += TB) {
                double s = banks[i];
                for (int w = 1; w < NCOPY; w++) s += banks[(size_t)w * gbank_stride + i];
                banks[i] = s;
            }
            __syncthreads();
        }
        // the halo's banks go to global memory (several work items of one halo add theirs up);
        // k_rows turns them into the result row
        {
            double* gb = ha.gbank + ha.bank_off[h];
            if (ha.n_items[h] > 1) {
                for (int i = threadIdx.x; i < nbank * V; i += TB)
                    if (banks[i] != 0.0) atomicAdd(&gb[i], banks[i]);
            } else {
                for (int i = threadIdx.x; i < nbank * V; i += TB) gb[i] = banks[i];
            }
        }
        __syncthreads();
    }
}

// Banks of the halos of `list` -> cumulative over radial shells, in place: a selection is a prefix of the shells,
// so the row writer then reads one shell instead of summing up to a dozen.  One thread per (halo, bank column).
__global__ void __launch_bounds__(128) k_bank_prefix(HaloArrays ha, const uint32_t* __restrict__ list,
                                                     const unsigned int* __restrict__ n_list, int per_shell) {
    const unsigned int it = blockIdx.y;
    for (unsigned int hi = it; hi < *n_list; hi += gridDim.y) {
        const uint32_t h = list[hi];
        const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
        if (c_hi <= c_lo || ha.status[h] >= 2) continue;
        const int nshell = ha.cuts[h].n + 1;
        double* b = ha.gbank + ha.bank_off[h];
        for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < per_shell; col += gridDim.x * blockDim.x) {
            double run = b[col];
            for (int s = 1; s < nshell; s++) {
                run += b[(size_t)s * per_shell + col];
                b[(size_t)s * per_shell + col] = run;
            }
        }
    }
}

// Result rows: one thread per (halo, selection) -- BoundSubhalo, each SO, each aperture.  The
// selections of a warp are the same, so the row writer's branches do not diverge.
struct SelIds {
    int n;
    int id[1 + SOAP_MAX_SO + SOAP_MAX_APERTURES];
};
template <int V, int NTY>
__global__ void __launch_bounds__(128) k_rows(HaloArrays ha, DevCfg cfg, const uint32_t* __restrict__ list,
                                              const unsigned int* __restrict__ n_list, SelIds sels) {
    const unsigned int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= *n_list) return;
    const uint32_t h = list[it];
    const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
    if (c_hi <= c_lo || ha.status[h] >= 2) return;
    const int sel = sels.id[blockIdx.y];
    const bool central = ha.central[h] == 1;
    const int n_so = central ? cfg.n_so : 0;
    const Cuts& cuts = ha.cuts[h];
    write_row<V, NTY>(ha.gbank + ha.bank_off[h], cuts, cuts.n, cfg, ha, h, ha.sres + h, cfg.do_sub && c_lo == 0, n_so,
                      ha.cofp[3 * h], ha.cofp[3 * h + 1], ha.cofp[3 * h + 2], sel);
}

__global__ void __launch_bounds__(TB, 3) k_kappa(ChunkView v, HaloArrays ha, DevCfg cfg, const Item* __restrict__ items,
                                              const unsigned int* __restrict__ n_items_dev) {
    __shared__ SweepShared SW;
    __shared__ KapSel sel[KAPPA_MAX_SEL];
    __shared__ double acc[KAPPA_MAX_SEL][11];
    __shared__ int nsel;
    const unsigned int n_items = *n_items_dev;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item im = items[it];
        const uint32_t h = im.halo;
        const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
        if (c_hi <= c_lo || ha.status[h] >= 2) continue;
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const double R = ha.rung_r[h];
        const double r2max = __dmul_rn(R, R), halfL = 0.5 * v.L, L = v.L;
        const int64_t hidx = ha.index[h];
        __syncthreads();
        if (threadIdx.x == 0) nsel = kappa_build_sels(sel, cfg, ha, h, c_lo, c_hi);
        for (int i = threadIdx.x; i < KAPPA_MAX_SEL * 11; i += TB) (&acc[0][0])[i] = 0.0;
        __syncthreads();
        const int ns = nsel;
        if (ns == 0) continue;
        sweep_item<SW_MASS | SW_VEL | SW_IDS | SW_TYPE>(v, SW, cx, cy, cz, R, im, [&](uint32_t t, bool ok) {
            if (!ok) return;
            const uint32_t tc = (uint32_t)v.type[t];
            if (tc != 0u && tc != 2u) return;  // gas and stars only
            const double r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
            if (!(r2 <= r2max)) return;
            const double x = rewrap_rel(v.px[t], cx, L, halfL);
            const double y = rewrap_rel(v.py[t], cy, L, halfL);
            const double z = rewrap_rel(v.pz[t], cz, L, halfL);
            kappa_add(sel, ns, acc, x, y, z, radius3(x, y, z), (double)v.mass[t], (double)v.vx[t], (double)v.vy[t],
                      (double)v.vz[t], tc, v.grnr[t] == hidx);
        });
        __syncthreads();
        for (int i = threadIdx.x; i < ns * 11; i += TB) {
            const double a = acc[i / 11][i % 11];
            if (a != 0.0) atomicAdd(sel[i / 11].slot(i % 11), a);
        }
        __syncthreads();
    }
}

__global__ void k_kappa_finish(HaloArrays ha, DevCfg cfg, const uint32_t* __restrict__ list,
                               const unsigned int* __restrict__ n_list) {
    unsigned int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= *n_list) return;
    const uint32_t h = list[it];
    const int c_lo = ha.commit_lo[h], c_hi = ha.commit_hi[h];
    if (c_hi <= c_lo || ha.status[h] >= 2) return;
    kappa_finish_row(cfg, ha, h, c_lo, c_hi);
}

__global__ void k_write_input_cols(HaloArrays ha, int64_t nh) {
    int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    double* row = ha.out + h * ha.ncol;
    row[0] = (double)ha.status[h];
    row[1] = (double)ha.nloop[h];
    row[2] = ha.cur_r[h];
}

}  // namespace

int soap_bank_stride(const DevCfg& cfg) {
    const bool full = (cfg.flags & (PF_KIN | PF_KAPPA | PF_TENS)) != 0;
    return (cfg.n_so + cfg.n_ap + 3) * 2 * (cfg.dmo ? 1 : 4) * (full ? V_FULL : V_MIN);
}

// result rows of the halos of `list` from their global banks (ha.gbank / ha.bank_off / ha.cuts)
int soap_launch_rows(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                     const unsigned int* n_list_dev, unsigned int n_list_host, cudaStream_t stream) {
    soap_handle* h = c->h;
    if (n_list_host == 0) return 0;
    const bool full = (cfg.flags & (PF_KIN | PF_KAPPA | PF_TENS)) != 0;
    SelIds sels;
    sels.n = 0;
    if (cfg.do_sub) sels.id[sels.n++] = 0;
    for (int q = 0; q < cfg.n_so; q++) sels.id[sels.n++] = 1 + q;
    for (int a = 0; a < cfg.n_ap; a++) sels.id[sels.n++] = 1 + SOAP_MAX_SO + a;
    if (sels.n == 0) return 0;
    {
        const int per_shell = 2 * (cfg.dmo ? 1 : 4) * (full ? V_FULL : V_MIN);  // doubles of one shell: [bound][type][V]
        const dim3 pg((unsigned)((per_shell + 127) / 128), n_list_host < 65535u ? n_list_host : 65535u);
        LAUNCH(h, k_bank_prefix, pg, 128, 0, stream, ha, list, n_list_dev, per_shell);
    }
    const dim3 grid(grid_for(n_list_host, 128), (unsigned)sels.n);
    if (full && !cfg.dmo) LAUNCH(h, (k_rows<V_FULL, 4>), grid, 128, 0, stream, ha, cfg, list, n_list_dev, sels);
    else if (full) LAUNCH(h, (k_rows<V_FULL, 1>), grid, 128, 0, stream, ha, cfg, list, n_list_dev, sels);
    else if (!cfg.dmo) LAUNCH(h, (k_rows<V_MIN, 4>), grid, 128, 0, stream, ha, cfg, list, n_list_dev, sels);
    else LAUNCH(h, (k_rows<V_MIN, 1>), grid, 128, 0, stream, ha, cfg, list, n_list_dev, sels);
    return 0;
}

int soap_launch_moments(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const Item* items,
                        const unsigned int* n_items_dev, unsigned int n_items_host, unsigned int grid,
                        cudaStream_t stream) {
    soap_handle* h = c->h;
    const bool full = (cfg.flags & (PF_KIN | PF_KAPPA | PF_TENS)) != 0;
    const int nty = cfg.dmo ? 1 : 4;
    const int V = full ? V_FULL : V_MIN;
    const int stride = (cfg.n_so + cfg.n_ap + 3) * 2 * nty * V;
    const int NW = TB / 32, VP = V | 1;
    const size_t stage_bytes = (size_t)NW * 32 * VP * sizeof(double) + (size_t)NW * 32 * sizeof(int);
    // warp-private banks when they fit next to the staging tiles (V <= 16: one copy per half-warp, always private)
    const int ncopy = V <= 16 ? 2 * NW : NW;
    const int priv = (V <= 16 || (size_t)ncopy * stride * sizeof(double) + stage_bytes <= 96 * 1024) ? 1 : 0;
    const size_t smem = (size_t)(priv ? ncopy : 1) * stride * sizeof(double) + stage_bytes;
    if (smem > 220 * 1024) SOAP_FAIL("soap_process_halos: %d SO + %d aperture variations need %zu bytes of shared memory", cfg.n_so, cfg.n_ap, smem);
    unsigned int g = n_items_host < grid ? n_items_host : grid;
    if (g < 1) g = 1;
#define MOM(VV, NT)                                                                                   \
    do {                                                                                              \
        CUDA_TRY(cudaFuncSetAttribute(k_moments<VV, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                    \
        LAUNCH(h, (k_moments<VV, NT>), g, TB, smem, stream, c->v, ha, cfg, items, n_items_dev, stride, priv); \
    } while (0)
    if (full && nty == 4) MOM(V_FULL, 4);
    else if (full) MOM(V_FULL, 1);
    else if (nty == 4) MOM(V_MIN, 4);
    else MOM(V_MIN, 1);
#undef MOM
    return 0;
}

int soap_launch_kappa(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, const Item* items,
                      const unsigned int* n_items_dev, unsigned int n_items_host, const uint32_t* acc_list,
                      const unsigned int* n_acc_dev, unsigned int n_acc_host, unsigned int grid,
                      cudaStream_t stream) {
    soap_handle* h = c->h;
    if (!(cfg.flags & PF_KAPPA) || cfg.dmo) return 0;  // no gas or stars: every kappa slot stays zero
    if (!(cfg.flags & PF_KIN)) SOAP_FAIL("soap_process_halos: kappa_corot needs the kinematics group (property_flags bit 0)");
    unsigned int g = n_items_host < grid ? n_items_host : grid;
    if (g < 1) g = 1;
    LAUNCH(h, k_kappa, g, TB, 0, stream, c->v, ha, cfg, items, n_items_dev);
    LAUNCH(h, k_kappa_finish, grid_for(n_acc_host, 128), 128, 0, stream, ha, cfg, acc_list, n_acc_dev);
    return 0;
}

int soap_launch_kappa_finish(soap_handle* h, const DevCfg& cfg, const HaloArrays& ha, const uint32_t* list,
                             const unsigned int* n_list_dev, unsigned int n_list_host, cudaStream_t stream) {
    if (n_list_host == 0) return 0;
    LAUNCH(h, k_kappa_finish, grid_for(n_list_host, 128), 128, 0, stream, ha, cfg, list, n_list_dev);
    return 0;
}

int soap_write_input_cols(soap_handle* h, const HaloArrays& ha, int64_t nh, cudaStream_t stream) {
    LAUNCH(h, k_write_input_cols, grid_for(nh, 128), 128, 0, stream, ha, nh);
    return 0;
}
