// iter.cu -- iterative inertia tensors (inertia_tensors.py:19-132 with the
// default max_iterations = 20).
//
// One tensor = one (selection, reduced?) pair of a halo:
//   BoundSubhalo  TotalInertiaTensor[Reduced]    bound particles, R = 10 HalfMassRadiusTot
//                                                (subhalo_properties.py:1075-1100)
//   SO            TotalInertiaTensor[Reduced]    every particle of the search sphere (in-sphere +
//                                                "surrounding"), R = R_SO (SO_properties.py:621-648)
//   apertures     StellarInertiaTensor[Reduced]  all stars of the halo mask, R = aperture radius
//                                                (aperture_properties.py:3579-3624)
//   projected     ProjectedTotalInertiaTensor[Reduced]  per axis, bound particles in the plane,
//                                                R = aperture radius (inertia_tensors.py:226-343,
//                                                projected_aperture_properties.py:789-852)
// The reference re-selects the particles inside the current ellipsoid up to 20
// times; every pass needs the eigen-decomposition of the previous one.  Here a
// pass is one sweep of the halo's sphere (the same work items and sweep_item
// as the moment kernels) that accumulates the sums of every live tensor of the
// halo at once, followed by one thread per halo that normalises, diagonalises
// (cyclic Jacobi, float64) and applies the reference's stopping rules.  The
// particle set of a selection is the sphere of the rung at which the selection
// was committed (halo_tasks.py:121-123 keeps finished halo_prop_list entries),
// whose radius write_row parked in the first iterative slot of the block.
#include "moments.cuh"

namespace {

constexpr int TB = SWEEP_NT;
constexpr int IT_MAX = 20;                                            // inertia_tensors.py:25
constexpr int IT_MAXSEL = 1 + SOAP_MAX_SO + SOAP_MAX_APERTURES + 3 * SOAP_MAX_APERTURES;
constexpr int IT_NV = 9;  // sum w, 6 weighted second moments, particles inside, particles handed in

struct __align__(8) ItState {
    double vec[9];   // eigenvectors: column j belongs to axis j
    double axis[3];  // ellipsoid semi-axes (coordinate units)
    double iax[3];   // their reciprocals (the sweep multiplies; a division per lane and axis is the hot spot)
    double q;        // sqrt(eig_val[1] / eig_val[2]) of this pass
    double R;        // sphere radius (coordinate units)
    double Rs2;      // squared radius of the sphere the selection was committed with
    int done, pad;
};

// selection s of the halo: block, layout, particle set
struct ItSel {
    double* blk;
    int out;         // block offset of the iterative pair: [out, out + w) full, [out + w, out + 2 w) reduced
    int w;           // 6 (3-D) or 3 (projected)
    int proj;        // -1: 3-D; else the projection axis
    uint32_t tmask;  // type codes (bit t)
    int bound;
};

__device__ inline bool it_selection(const DevCfg& cfg, const HaloArrays& ha, uint32_t h, int s, ItSel& o, double& R,
                                    bool& gate) {
    double* row = ha.out + (int64_t)h * ha.ncol;
    const int off_so = cfg.do_sub ? 1 : 0, off_ap = off_so + cfg.n_so, off_pj = off_ap + cfg.n_ap;
    o.w = 6; o.proj = -1;
    if (s < off_so) {
        o.blk = row + cfg.lay.sub; o.out = cfg.lay.bsub.tens + 12; o.tmask = 0xfu; o.bound = 1;
        R = 10.0 * o.blk[cfg.lay.bsub.extra];  // HalfMassRadiusTot
        gate = o.blk[8] != 0.0;
    } else if (s < off_ap) {
        const int q = s - off_so;
        o.blk = row + cfg.lay.so[q]; o.out = cfg.lay.bso.tens + 12; o.tmask = 0xfu; o.bound = 0;
        R = o.blk[cfg.lay.bso.extra];  // SO radius
        gate = ha.central[h] == 1 && o.blk[8] != 0.0;  // an SO that does not exist never parks a radius (k_it_init)
    } else if (s < off_pj) {
        const int a = s - off_ap;
        o.blk = row + cfg.lay.ap[a]; o.out = cfg.lay.bap.tens + 12; o.tmask = 1u << 2; o.bound = cfg.ap_incl[a] == 0;
        R = cfg.ap_r[a];
        gate = o.blk[6] != 0.0;  // Mstar inside the aperture
    } else {
        const int p = (s - off_pj) / 3;
        o.proj = (s - off_pj) % 3;
        o.blk = row + cfg.lay.pj[p] + o.proj * cfg.lay.pjb; o.out = PJ_BLOCK; o.w = 3; o.tmask = 0xfu; o.bound = 1;
        R = cfg.pj_r[p];
        gate = o.blk[8] != 0.0;  // mass inside the projected aperture
    }
    return true;
}

__global__ void k_it_list(HaloArrays ha, int64_t nh, uint32_t* __restrict__ list, unsigned int* __restrict__ n_list) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    if (ha.status[h] == 0) list[atomicAdd(n_list, 1u)] = (uint32_t)h;
}

__global__ void k_it_init(HaloArrays ha, DevCfg cfg, int64_t nh, int nsel, ItState* __restrict__ state,
                          int* __restrict__ alive) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nh * nsel) return;
    const uint32_t h = (uint32_t)(i / nsel);
    const int s = (int)(i % nsel);
    ItSel sel;
    double R;
    bool gate;
    it_selection(cfg, ha, h, s, sel, R, gate);
    // the radius of the sphere this selection was computed with (parked by write_row)
    double rs = ha.cur_r[h];  // projected: bound particles only, all of them inside the final sphere
    if (sel.proj < 0) {
        rs = sel.blk[sel.out];
        sel.blk[sel.out] = 0.0;
    }
    if (!(rs > 0.0) || ha.status[h] != 0) gate = false;  // block never written (no SO, aperture skipped) / halo failed
    for (int red = 0; red < 2; red++) {
        ItState& st = state[((size_t)h * nsel + s) * 2 + red];
        for (int k = 0; k < 9; k++) st.vec[k] = (k % 4 == 0) ? 1.0 : 0.0;
        if (sel.proj >= 0) { st.vec[3] = 1.0; st.vec[4] = 0.0; }  // 2x2 identity, row-major
        st.axis[0] = st.axis[1] = st.axis[2] = R;
        st.iax[0] = st.iax[1] = st.iax[2] = 1.0 / R;
        st.q = 1.0;
        st.R = R;
        st.Rs2 = __dmul_rn(rs, rs);
        st.done = gate ? 0 : 1;
        st.pad = 0;
    }
    if (gate) alive[h] = 1;
}

__global__ void __launch_bounds__(TB, 3) k_it_accum(ChunkView v, HaloArrays ha, DevCfg cfg,
                                                 const Item* __restrict__ items,
                                                 const unsigned int* __restrict__ n_items_dev, int nsel,
                                                 const ItState* __restrict__ state, double* __restrict__ sums,
                                                 const int* __restrict__ alive) {
    __shared__ SweepShared SW;
    extern __shared__ __align__(16) unsigned char it_smem[];
    const int nt = 2 * nsel;
    ItState* S = reinterpret_cast<ItState*>(it_smem);                          // [nt]
    double* wsum = reinterpret_cast<double*>(S + nt);                          // [TB / 32][nt][IT_NV]
    ItSel* SEL = reinterpret_cast<ItSel*>(wsum + (size_t)(TB / 32) * nt * IT_NV);  // [nsel]
    const unsigned int n_items = *n_items_dev;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (unsigned int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const Item im = items[it];
        const uint32_t h = im.halo;
        if (!alive[h]) continue;
        __syncthreads();
        for (int i = threadIdx.x; i < nt; i += TB) S[i] = state[(size_t)h * nt + i];
        for (int i = threadIdx.x; i < nsel; i += TB) {
            double R;
            bool gate;
            it_selection(cfg, ha, h, i, SEL[i], R, gate);
        }
        for (int i = threadIdx.x; i < (TB / 32) * nt * IT_NV; i += TB) wsum[i] = 0.0;
        __syncthreads();
        const double cx = ha.cofp[3 * h], cy = ha.cofp[3 * h + 1], cz = ha.cofp[3 * h + 2];
        const double R = ha.cur_r[h];
        const double halfL = 0.5 * v.L, L = v.L;
        const int64_t hidx = ha.index[h];
        sweep_item<SW_MASS | SW_IDS | SW_TYPE>(v, SW, cx, cy, cz, R, im, [&](uint32_t t, bool ok) {
            double r2 = 0.0, x = 0.0, y = 0.0, z = 0.0, m = 0.0, nrm = 1.0;
            uint32_t tbit = 0;
            bool bound = false;
            if (ok) {
                r2 = periodic_r2(v.px[t], v.py[t], v.pz[t], cx, cy, cz, L, halfL);
                x = rewrap_rel(v.px[t], cx, L, halfL);
                y = rewrap_rel(v.py[t], cy, L, halfL);
                z = rewrap_rel(v.pz[t], cz, L, halfL);
                const double r = radius3(x, y, z);
                nrm = r * r;
                m = (double)v.mass[t];
                tbit = 1u << (cfg.dmo ? 1u : (uint32_t)v.type[t]);
                bound = v.grnr[t] == hidx;
            }
            const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
            for (int s = 0; s < nsel; s++) {
                if (S[2 * s].done && S[2 * s + 1].done) continue;
                const bool member = ok && r2 <= S[2 * s].Rs2 && (SEL[s].tmask & tbit) && (!SEL[s].bound || bound);
                if (!__any_sync(0xffffffffu, member)) continue;
                const int proj = SEL[s].proj;
                // in-plane coordinates of a projection: axis 0 -> (y, z), 1 -> (z, x), 2 -> (x, y)
                // (inertia_tensors.py:272-280)
                const double pa = proj == 0 ? y : (proj == 1 ? z : x), pb = proj == 0 ? z : (proj == 1 ? x : y);
                const double paa = proj == 0 ? yy : (proj == 1 ? zz : xx), pbb = proj == 0 ? zz : (proj == 1 ? xx : yy);
                const double pab = proj == 0 ? yz : (proj == 1 ? xz : xy);
                const double nrm_s = proj < 0 ? nrm : paa + pbb;
                for (int red = 0; red < 2; red++) {
                    const ItState& st = S[2 * s + red];
                    if (st.done) continue;
                    // reduced: particles at the centre are dropped first (inertia_tensors.py:61-68,282-289)
                    const bool mem2 = member && !(red && nrm_s <= 1e-8);
                    bool inside;
                    if (proj < 0) {
                        const double px = ((x * st.vec[0] + y * st.vec[3]) + z * st.vec[6]) * st.iax[0];
                        const double py = ((x * st.vec[1] + y * st.vec[4]) + z * st.vec[7]) * st.iax[1];
                        const double pz = ((x * st.vec[2] + y * st.vec[5]) + z * st.vec[8]) * st.iax[2];
                        inside = mem2 && sqrt((px * px + py * py) + pz * pz) <= 1.0;
                    } else {
                        const double p0 = (pa * st.vec[0] + pb * st.vec[2]) * st.iax[0];
                        const double p1 = (pa * st.vec[1] + pb * st.vec[3]) * st.iax[1];
                        inside = mem2 && sqrt(p0 * p0 + p1 * p1) <= 1.0;
                    }
                    const double w = inside ? m : 0.0;
                    const double wq = red ? (inside ? m / nrm_s : 0.0) : w;
                    double val[IT_NV];
                    val[0] = w;
                    if (proj < 0) {
                        val[1] = wq * xx; val[2] = wq * yy; val[3] = wq * zz; val[4] = wq * xy; val[5] = wq * xz; val[6] = wq * yz;
                    } else {
                        val[1] = wq * paa; val[2] = wq * pbb; val[3] = wq * pab; val[4] = val[5] = val[6] = 0.0;
                    }
                    // the two counts by ballot; the weighted sums by butterfly, skipped when no lane is inside
                    const unsigned b_in = __ballot_sync(0xffffffffu, inside);
                    const unsigned b_mem = __ballot_sync(0xffffffffu, member);
                    double* ws = wsum + ((size_t)wid * nt + 2 * s + red) * IT_NV;
                    if (lane == 0) { ws[7] += (double)__popc(b_in); ws[8] += (double)__popc(b_mem); }
                    if (b_in == 0u) continue;
#pragma unroll
                    for (int k = 0; k < 7; k++) {
                        if (proj >= 0 && k >= 4) continue;
                        double a = val[k];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                        if (lane == 0) ws[k] += a;
                    }
                }
            }
        });
        __syncthreads();
        for (int i = threadIdx.x; i < nt * IT_NV; i += TB) {
            double a = 0.0;
            for (int w = 0; w < TB / 32; w++) a += wsum[(size_t)w * nt * IT_NV + i];
            if (a != 0.0) atomicAdd(&sums[((size_t)h * nt + i / IT_NV) * IT_NV + i % IT_NV], a);
        }
    }
}

// eigen-decomposition of the symmetric 3x3 matrix t = [xx, yy, zz, xy, xz, yz]: cyclic Jacobi;
// eigenvalues ascending like numpy.linalg.eigh, vec column j = eigenvector j
__device__ void eigh3(const double* t, double* val, double* vec) {
    double a[3][3] = {{t[0], t[3], t[4]}, {t[3], t[1], t[5]}, {t[4], t[5], t[2]}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 60; sweep++) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double dia = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off == 0.0 || off <= 1e-36 * dia) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), sn = tt * c;
                for (int k = 0; k < 3; k++) {  // A <- A J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq;
                    a[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {  // A <- J^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk;
                    a[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    int idx[3] = {0, 1, 2};
    double e[3] = {a[0][0], a[1][1], a[2][2]};
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2 - i; j++)
            if (e[idx[j]] > e[idx[j + 1]]) { const int tmp = idx[j]; idx[j] = idx[j + 1]; idx[j + 1] = tmp; }
    for (int j = 0; j < 3; j++) {
        val[j] = e[idx[j]];
        for (int i = 0; i < 3; i++) vec[3 * i + j] = V[i][idx[j]];
    }
}

__global__ void k_it_update(HaloArrays ha, DevCfg cfg, const uint32_t* __restrict__ list,
                            const unsigned int* __restrict__ n_list, int nsel, int iter, ItState* __restrict__ state,
                            double* __restrict__ sums, int* __restrict__ alive, unsigned int* __restrict__ n_alive) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_list) return;
    const uint32_t h = list[i];
    if (!alive[h]) return;
    const int nt = 2 * nsel;
    const double k2 = cfg.kpc * cfg.kpc;
    int any = 0;
    for (int ti = 0; ti < nt; ti++) {
        ItState& st = state[(size_t)h * nt + ti];
        if (st.done) continue;
        double* sp = sums + ((size_t)h * nt + ti) * IT_NV;
        double s[IT_NV];
        for (int k = 0; k < IT_NV; k++) { s[k] = sp[k]; sp[k] = 0.0; }
        const int red = ti & 1;
        ItSel sel;
        double R;
        bool gate;
        it_selection(cfg, ha, h, ti >> 1, sel, R, gate);
        double* out = sel.blk + sel.out + sel.w * red;
        // fewer than min_particles handed in / inside the initial sphere: None (inertia_tensors.py:58,103)
        if (iter == 0 && (s[8] < 20.0 || s[7] < 20.0)) { st.done = 1; continue; }
        if (!(s[0] != 0.0)) { st.done = 1; continue; }  // nothing left inside the ellipsoid
        double T[6];
        for (int q = 0; q < 6; q++) T[q] = s[1 + q] / s[0];
        // positions in kpc (inertia_tensors.py:77-78); the reduced tensor is dimensionless
        for (int q = 0; q < sel.w; q++) out[q] = red ? T[q] : T[q] * k2;
        if (st.q == 0.0) {  // :126-128, :337-339
            for (int q = 0; q < sel.w; q++) out[q] = 0.0;
            st.done = 1;
            continue;
        }
        if (iter == IT_MAX - 1) { st.done = 1; continue; }
        if (sel.proj < 0) {
            double val[3], vec[9];
            eigh3(T, val, vec);
            for (int k = 0; k < 3; k++) val[k] = fabs(val[k]);  // :123
            const double q = sqrt(val[1] / val[2]), sa = sqrt(val[0] / val[2]), p = sqrt(val[0] / val[1]);
            if (fabs((st.q - q) / q) < 0.0001) { st.done = 1; continue; }  // converged (:89-90)
            st.q = q;
            st.axis[0] = st.R * cbrt(sa * p);
            st.axis[1] = st.R * cbrt(q / p);
            st.axis[2] = st.R * (1.0 / cbrt(q * sa));
            for (int k = 0; k < 3; k++) st.iax[k] = 1.0 / st.axis[k];
            for (int k = 0; k < 9; k++) st.vec[k] = vec[k];
        } else {
            // 2x2 symmetric [[aa, ab], [ab, bb]]: one Jacobi rotation, eigenvalues ascending (:300-341)
            const double aa = T[0], bb = T[1], ab = T[2];
            double c = 1.0, sn = 0.0, e0 = aa, e1 = bb;
            if (ab != 0.0) {
                const double theta = (bb - aa) / (2.0 * ab);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                c = 1.0 / sqrt(tt * tt + 1.0);
                sn = tt * c;
                e0 = aa - tt * ab;
                e1 = bb + tt * ab;
            }
            // eigenvectors: columns (c, -sn) for e0 and (sn, c) for e1
            double v00 = c, v10 = -sn, v01 = sn, v11 = c;
            if (e0 > e1) {
                double tmp = e0; e0 = e1; e1 = tmp;
                tmp = v00; v00 = v01; v01 = tmp;
                tmp = v10; v10 = v11; v11 = tmp;
            }
            const double q = sqrt(e0 / e1);
            if (fabs((st.q - q) / q) < 0.0001) { st.done = 1; continue; }
            st.q = q;
            st.axis[0] = st.R * sqrt(q);
            st.axis[1] = st.R * (1.0 / sqrt(q));
            st.iax[0] = 1.0 / st.axis[0]; st.iax[1] = 1.0 / st.axis[1];
            st.vec[0] = v00; st.vec[1] = v01; st.vec[2] = v10; st.vec[3] = v11;
        }
        any = 1;
    }
    alive[h] = any;
    if (any) atomicAdd(n_alive, 1u);
}

}  // namespace

// list <- halos with status 0; returns their number through n_host
int soap_iter_list(soap_handle* h, const HaloArrays& ha, int64_t nh, uint32_t* list, unsigned int* n_list_dev,
                   unsigned int* n_host, cudaStream_t stream) {
    CUDA_TRY(cudaMemsetAsync(n_list_dev, 0, sizeof(unsigned int), stream));
    LAUNCH(h, k_it_list, grid_for(nh, 128), 128, 0, stream, ha, nh, list, n_list_dev);
    CUDA_TRY(cudaMemcpyAsync(n_host, n_list_dev, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return 0;
}

int soap_launch_iter_tensors(soap_chunk* c, const DevCfg& cfg, const HaloArrays& ha, int64_t nh, const Item* items,
                             const unsigned int* n_items_dev, unsigned int n_items_host, const uint32_t* list,
                             const unsigned int* n_list_dev, unsigned int n_list_host, unsigned int grid,
                             cudaStream_t stream) {
    soap_handle* h = c->h;
    const int nsel = (cfg.do_sub ? 1 : 0) + cfg.n_so + cfg.n_ap + 3 * cfg.n_pj;
    if (nsel == 0 || n_list_host == 0) return 0;
    const size_t smem = 2 * (size_t)nsel * (sizeof(ItState) + (TB / 32) * IT_NV * sizeof(double)) + nsel * sizeof(ItSel);
    CUDA_TRY(cudaFuncSetAttribute(k_it_accum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t nt = 2 * (size_t)nsel;
    ItState* state = (ItState*)h->get("h_it_state", sizeof(ItState) * nt * (size_t)nh);
    double* sums = (double*)h->get("h_it_sums", sizeof(double) * IT_NV * nt * (size_t)nh);
    int* alive = (int*)h->get("h_it_alive", sizeof(int) * (size_t)nh);
    unsigned int* n_alive = (unsigned int*)h->get("h_it_nalive", sizeof(unsigned int) * IT_MAX);
    if (!state || !sums || !alive || !n_alive) return -1;
    CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * IT_NV * nt * (size_t)nh, stream));
    CUDA_TRY(cudaMemsetAsync(alive, 0, sizeof(int) * (size_t)nh, stream));
    CUDA_TRY(cudaMemsetAsync(n_alive, 0, sizeof(unsigned int) * IT_MAX, stream));
    LAUNCH(h, k_it_init, grid_for(nh * nsel, 128), 128, 0, stream, ha, cfg, nh, nsel, state, alive);
    unsigned int g = n_items_host < grid ? n_items_host : grid;
    if (g < 1) g = 1;
    for (int iter = 0; iter < IT_MAX; iter++) {
        LAUNCH(h, k_it_accum, g, TB, smem, stream, c->v, ha, cfg, items, n_items_dev, nsel, state, sums, alive);
        LAUNCH(h, k_it_update, grid_for(n_list_host, 128), 128, 0, stream, ha, cfg, list, n_list_dev, nsel, iter, state,
               sums, alive, n_alive + iter);
        if (iter % 4 == 3) {  // most tensors converge in a few passes: stop when none is left
            unsigned int left = 0;
            CUDA_TRY(cudaMemcpyAsync(&left, n_alive + iter, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            if (left == 0) break;
        }
    }
    return 0;
}
