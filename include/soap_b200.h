/*
 * soap_b200.h -- C ABI of the B200-native SOAP per-halo aggregation path.
 *
 * The reference (SWIFTSIM/SOAP) is pure Python; it has no FFI of its own.  Each
 * entry point below therefore replaces a Python function of the hot path and
 * is what a ctypes binding on the reference side would load.
 * Citations are file:line relative to the reference tree.
 *
 * Conventions
 *  - every `dev` pointer is DEVICE memory owned by the caller (e.g. a torch
 *    tensor's data_ptr); `host` pointers are host memory;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    calls are asynchronous on that stream unless stated ("syncs");
 *  - return value: 0 = ok, <0 = error (message via soap_last_error());
 *    no exception crosses the ABI;
 *  - the library owns only scratch memory held by the handle / chunk;
 *  - positions are float64 [N,3] row-major in *coordinate units* (comoving
 *    snap_length), already box-wrapped (soap_box_wrap); masses float32,
 *    velocities float32 [N,3]; GroupNr_bound / FOFGroupIDs int32 or int64;
 *  - every physical threshold is passed pre-converted to coordinate units by
 *    the host (the value unyt would produce at that comparison).
 */
#ifndef SOAP_B200_H
#define SOAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOAP_B200_ABI_VERSION 2

/* per-halo status codes written by soap_process_halos */
#define SOAP_HALO_OK 0
#define SOAP_HALO_RADIUS_TOO_SMALL 1 /* needs a larger read_radius: halo_tasks.py:169-181,390-402 */
#define SOAP_HALO_COUNT_MISMATCH 2   /* Ntot > nr_bound_part: subhalo_properties.py:2642-2646 (RuntimeError) */
#define SOAP_HALO_SO_NOT_FOUND 3     /* SO_properties.py:150-153,190-193 (RuntimeError beyond 20 Mpc) */
#define SOAP_HALO_ROOT_FAILED 4      /* scipy brentq ValueError (same-sign bracket) at SO_properties.py:208 */

#define SOAP_MAX_SO 8
#define SOAP_MAX_APERTURES 16
#define SOAP_MAX_PTYPES 8
#define SOAP_MAX_FILTERS 8

typedef struct soap_handle soap_handle;
typedef struct soap_mesh soap_mesh;
typedef struct soap_chunk soap_chunk;

/* ------------------------------------------------------------------ runtime */
int soap_abi_version(void);
const char* soap_last_error(void);
/* per-device context (scratch workspace).  One handle per GPU / Python thread. */
int soap_create(int device, soap_handle** out);
int soap_destroy(soap_handle* h);
/* number of kernels launched through this handle so far (bench.py gpu_launches) */
int64_t soap_launch_count(const soap_handle* h);
/* per-kernel device time: with timing on, every launch through the handle is bracketed by CUDA events on its
 * stream; soap_kernel_timings synchronises the device, writes "kernel name\tlaunches\tmilliseconds\n" lines for
 * everything launched since the last call and returns the number of bytes (bench.py's roofline line) */
int soap_kernel_timing(soap_handle* h, int on);
int64_t soap_kernel_timings(soap_handle* h, char* buf, int64_t buflen);

/* ------------------------------------------------------------------ stage A */
/* box_wrap: SOAP/core/chunk_tasks.py:48-50 (numpy floored modulo), in place. */
int soap_box_wrap(soap_handle* h, double* pos_dev, int64_t n, const double ref_pos[3],
                  double boxsize, void* stream);

/* SharedMesh.__init__: SOAP/core/shared_mesh.py:11-114.
 * Outputs pos_min/pos_max/cell_size (host, exact), cell_idx int32 [n] (optional,
 * NULL to skip), cell_count / cell_offset int64 [res^3], sort_idx int64 [n]
 * (all device).  stable != 0 orders sort_idx by ascending particle index within
 * a cell (the oracle's choice; the reference's order comes from VirgoDC's
 * parallel_sort and is unpinned).  Syncs the stream (bounds go to the host). */
int soap_mesh_build(soap_handle* h, const double* pos_dev, int64_t n, int resolution,
                    double pos_min[3], double pos_max[3], double cell_size[3],
                    int32_t* cell_idx_dev, int64_t* cell_count_dev, int64_t* cell_offset_dev,
                    int64_t* sort_idx_dev, int stable, void* stream);

/* ------------------------------------------------------------------ stage B */
/* SharedMesh.query_radius_periodic: SOAP/core/shared_mesh.py:122-200, batched
 * over n_query (centre, radius) pairs.  Two calls: with idx_dev == NULL the
 * per-query counts (int64 [n_query]) and enclosed float64 mass sums (optional,
 * needs mass_dev) are written; the caller exclusive-scans counts into
 * offsets_dev (int64 [n_query+1]) and calls again with idx_dev (int64
 * [offsets[n_query]]).  Index order per query: cells in ascending (k, j, i),
 * sort_idx order within a cell -- the reference's loop order with its python
 * sets iterated ascending. */
int soap_sphere_query(soap_handle* h, const double* pos_dev, int64_t n, int resolution,
                      const double pos_min[3], const double pos_max[3], const double cell_size[3],
                      const int64_t* cell_count_dev, const int64_t* cell_offset_dev,
                      const int64_t* sort_idx_dev, const double* centres_dev,
                      const double* radii_dev, int64_t n_query, double boxsize,
                      int64_t* counts_dev, const int64_t* offsets_dev, int64_t* idx_dev,
                      const float* mass_dev, double* enclosed_mass_dev, void* stream);

/* ------------------------------------------------- stage B+C, halo batching */
/* one particle type of a chunk, as SOAP holds it after chunk_tasks.py:256-288 */
typedef struct {
    int ptype;            /* 0 gas, 1 dm, 4 star, 5 bh */
    int ids_are_int64;    /* dtype of grnr/fof: 0 = int32, 1 = int64 */
    int64_t n;
    const double* pos;    /* dev [n,3] Coordinates */
    const float* mass;    /* dev [n]   mass_dataset(ptype): SOAP/core/dataset_names.py:7 */
    const float* vel;     /* dev [n,3] Velocities */
    const void* grnr;     /* dev [n]   GroupNr_bound */
    const void* fof;      /* dev [n]   FOFGroupIDs */
} soap_ptype_arrays;

/* Build the device-resident chunk: one merged particle set in cell order
 * (internal fine mesh; membership does not depend on the mesh).  Replaces the
 * per-ptype SharedMesh construction at SOAP/core/chunk_tasks.py:299-304 for the
 * batched path.  fine_ppc = target particles per internal cell (0 = default). */
int soap_chunk_create(soap_handle* h, const soap_ptype_arrays* types, int n_types,
                      double boxsize, int fine_ppc, soap_chunk** out, void* stream);
int soap_chunk_destroy(soap_chunk* c);
int64_t soap_chunk_num_particles(const soap_chunk* c);

typedef struct {
    /* cellgrid / unyt scalars in coordinate units */
    double boxsize;
    double G;                 /* vmax = sqrt(G M / r) with r in coordinate units */
    double H;                 /* KineticEnergy Hubble term, velocity / coordinate length */
    double kpc_per_length;    /* inertia_tensors.py:77-78 */
    double r_20mpc;           /* SO_properties.py:150 */
    double nu_density;        /* SO_properties.py:3426-3435 */
    double phys_mpc_to_coord; /* halo_tasks.py:168 */
    double softening[SOAP_MAX_PTYPES]; /* indexed by ptype */
    /* halo_tasks.py:306-317; <= 0 means "no target density" */
    double target_density;
    /* halo_prop_list, in the reference's order: BoundSubhalo, SO..., apertures */
    int do_subhalo;
    int n_so;
    double so_reference_density[SOAP_MAX_SO]; /* SO_properties.py:3494-3512 */
    int so_virial[SOAP_MAX_SO];               /* virial_definition: concentration */
    int n_apertures;                          /* ascending radius */
    double ap_radius[SOAP_MAX_APERTURES];     /* coordinate units */
    double ap_physical_mpc[SOAP_MAX_APERTURES];
    int ap_inclusive[SOAP_MAX_APERTURES];
    int n_projected;                          /* ProjectedAperture radii, ascending (needs do_subhalo) */
    double proj_radius[SOAP_MAX_APERTURES];   /* coordinate units */
    double proj_physical_mpc[SOAP_MAX_APERTURES];
    /* property groups: bit 0 kinematics (veldisp, L), bit 1 kappa_corot / DtoT and the stellar
     * rotational velocity / cylindrical dispersions (needs bit 0), bit 2 non-iterative inertia
     * tensors, bit 3 half-mass radii per type (also the projected ones), bit 4 the iterative
     * 3-D inertia tensors (inertia_tensors.py:19-132, max_iterations = 20; needs bit 2) */
    uint32_t property_flags;
    int dmo;                  /* only dark matter present / requested */
    /* CategoryFilter (SOAP/core/category_filter.py:69-110): filter f is satisfied when the sum of the
     * BoundSubhalo particle counts of the types in filter_types[f] (bit 0 gas, 1 dm, 2 star, 3 bh) is
     * >= filter_limit[f].  Index 0 is "basic" (always satisfied; its entries are ignored).  A variation
     * whose halo_filter is not satisfied is left at exact zeros and never asks for a larger radius
     * (SO_properties.py:3627, aperture_properties.py:4127, projected_aperture_properties.py:1888).
     * Filters other than 0 need do_subhalo. */
    int n_filters;                            /* including index 0; 0 or 1 = no filtering */
    int64_t filter_limit[SOAP_MAX_FILTERS];
    uint32_t filter_types[SOAP_MAX_FILTERS];
    int so_filter[SOAP_MAX_SO];
    int ap_filter[SOAP_MAX_APERTURES];
    int proj_filter[SOAP_MAX_APERTURES];
    /* skip_gt_enclose_radius (aperture_properties.py:4082-4123, projected_aperture_properties.py:1827-1888):
     * radius of the previous aperture of the list in coordinate units, or < 0 when the shortcut is off /
     * this is the first radius.  If it exceeds BoundSubhalo/EncloseRadius an inclusive or projected
     * aperture is skipped (zeros) and an exclusive one equals the previous exclusive aperture, so it is
     * computed from the particles already loaded without asking for a larger radius. */
    double ap_prev_radius[SOAP_MAX_APERTURES];
    double proj_prev_radius[SOAP_MAX_APERTURES];
    /* cross-check / measurement switches, 0 in production: bit 0 = route every halo through the general
     * (kernel-sequence) path instead of the small-halo tiers; bit 1 = launch every kernel on the caller's
     * stream, one after the other, instead of overlapping tiers, scan variants and the general path */
    uint32_t debug_flags;
} soap_halo_config;

/* Column layout of the result table for a config: writes a '\n'-separated list
 * of "name:width" into buf (host) and returns the total number of float64
 * columns (or <0).  Names follow the reference's output groups
 * ("BoundSubhalo/...", "SO/<i>/...", "Aperture/<i>/...", "ProjectedAperture/<i>/proj{x,y,z}/..."). */
int64_t soap_result_layout(const soap_halo_config* cfg, char* buf, int64_t buflen);

/* process_halos: SOAP/core/halo_tasks.py:276-430 with process_single_halo
 * (:23-273) batched on the device: radius ladder + density gate, periodic
 * gather and halo-centred re-wrap, then BoundSubhalo / SO / aperture
 * reductions.  Halo arrays are device pointers of length n_halo (cofp
 * [n_halo,3]).  out_dev is float64 [n_halo, ncol] row-major (ncol from
 * soap_result_layout), status_dev int32 [n_halo].  For halos that end with
 * SOAP_HALO_RADIUS_TOO_SMALL the updated search/read radii
 * (halo_tasks.py:166-181,390-402) are in the InputHalos columns.  Syncs:
 * the work is ordered after what is queued on `stream` and complete when the
 * call returns; internally the small-halo tiers and the scan variants run on
 * streams of the handle next to the general path on `stream` (events order
 * them; debug_flags bit 1 puts everything on `stream`).  One call at a time
 * per handle. */
int soap_process_halos(soap_chunk* c, const soap_halo_config* cfg, int64_t n_halo,
                       const double* cofp_dev, const double* search_radius_dev,
                       const double* read_radius_dev, const int64_t* index_dev,
                       const int32_t* is_central_dev, const int64_t* nr_bound_part_dev,
                       double* out_dev, int64_t ncol, int32_t* status_dev, void* stream);

/* particle-halo pairs found at the accepted radii by the last soap_process_halos
 * (the unit of BASELINE.json's pairs/s metric; halo_tasks.py:89) */
int64_t soap_chunk_last_pairs(const soap_chunk* c);
/* event-timed milliseconds of named phases of the last soap_process_halos /
 * soap_chunk_create ("mesh", "count", "collect", "sort", "scan", "moments"...);
 * writes "name:ms\n" lines; returns number of bytes. */
int64_t soap_chunk_timings(const soap_chunk* c, char* buf, int64_t buflen);

#ifdef __cplusplus
}
#endif
#endif /* SOAP_B200_H */
