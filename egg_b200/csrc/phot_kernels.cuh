// phot_kernels.cuh — launch interface between the C-ABI host code and the sm_100a kernels.
#ifndef EGGPHOT_PHOT_KERNELS_CUH
#define EGGPHOT_PHOT_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "phot_core.cuh"
#include "phot_lane.cuh"

namespace eggphot {

constexpr int kMaxLines = 64;
constexpr int kMaxGroupBands = 64;
constexpr int kMaxGroups = 16;
constexpr int kMaxRf = 16;

struct GroupDesc {
    const uint8_t *blob;   // device copy of BandGroup::blob (16-byte aligned)
    uint32_t bytes, nb;
    uint32_t base, pad; // index of the group's first band in group order
};

// sparse rest-frame weights of one rfband on one component grid
struct RfDesc {
    int32_t j0, n;        // first node and count
    int32_t off;          // offset into the weight array (elements)
    int32_t covered;      // SED covers the filter at z = 0
};

// Everything one launch of the flux kernel reads.  Passed by value (__grid_constant__).
struct FluxParams {
    // galaxy columns (device pointers; reference names, src/egg-gencat.cpp:551-603)
    const float *z, *d, *m_disk, *m_bulge, *m2lcor, *lir, *fpah, *mdust, *igm_abs, *vdisp;
    const float *line_lum_disk, *line_lum_bulge;
    const uint64_t *opt_sed_disk, *opt_sed_bulge, *ir_sed;
    uint64_t ngal;
    // outputs [ngal][nband_total] / [ngal][nrf]
    float *flux, *flux_disk, *flux_bulge, *rfmag, *rfmag_disk, *rfmag_bulge;
    double *flux_disk_f64, *flux_bulge_f64, *rfmag_disk_f64, *rfmag_bulge_f64;
    uint32_t nband_total, nrf;
    // band groups: each is staged into shared memory in turn while a batch of galaxies is resident
    uint32_t ngroups, max_blob;
    GroupDesc groups[kMaxGroups];
    int32_t do_rf;         // also compute the rest-frame magnitudes
    int32_t batch;         // galaxies per batch (batch_of(precision))
    // per-CTA scratch for the rest-frame SEDs of a batch: [grid][batch][2*strideD + strideB] real (stays in L2):
    // nD pairs {disk SED, bulge SED resampled on the disk grid}, then the bulge SED on its own grid
    void *scratch;
    // component grids
    int32_t nB, nD;        // nodes (0 if the component is absent)
    int32_t strideB, strideD; // row strides in elements (16-byte multiples)
    const double *xB, *xD; // rest wavelengths, double
    const void *gridA, *gridB; // disk grid: GridA records; fp64 GridBD records, fp32 a float array of dxo
                               // (phot_core.cuh); one padding record at the end
    const uint16_t *lutB, *lutD; // log2-spaced "last node <= v" lookups (Grid::lut)
    int32_t lutB_n, lutD_n, lutB_e0, lutD_e0;
    int32_t igmB[4], igmD[4]; // lz, l0, l1, l2
    // template rows in the kernel's real type
    const void *rowsB, *rowsDopt, *rowsDir, *rowsDpah;
    // disk grid -> optical grid maps (Prepared::u_*), for the bulge evaluated on the disk grid
    const int32_t *u_oleft;
    const void *u_ofrac;   // real
    const uint8_t *opt_use;
    uint32_t nopt_sed, nir_sed;
    int32_t ir_kind, disk_mode, has_bulge, apply_igm;
    // emission lines
    int32_t nline, line_average;
    float line_lambda[kMaxLines];   // sorted by wavelength
    uint8_t line_col[kMaxLines];    // column of the caller's line_lum arrays for each sorted line
    // rest-frame weights
    const RfDesc *rfB, *rfD;
    const void *rfW;       // real
    int32_t *error_flag;   // set to EGGPHOT_ERR_INDEX if a galaxy points outside a library
};

// dynamic shared memory the flux kernel needs for these parameters
size_t flux_kernel_smem(const FluxParams &p, int precision);
// launches the flux kernel over all band groups; returns cudaError_t as int
int launch_flux_kernel(const FluxParams &p, int precision, int grid_blocks, cudaStream_t stream);
int flux_kernel_max_blocks_per_sm(const FluxParams &p, int precision);

// ---- the lane-per-galaxy flux pipeline (phot_lane_kernels.cu) ----
// fp.groups point at the BandHeader2 blobs.  A call is worked off in chunks of `nbatch` batches of 32 galaxies (in the
// order of `perm`); the chunk's buffers: SED rows, band-table windows, build ranges, table sums, item counters.
struct LaneParams {
    FluxParams fp;
    const uint32_t *perm;     // [ngal] galaxies bucketed by redshift (zsort); batches are 32 consecutive entries
    const double2 *gxD, *gxB; // [rowsD], [rowsB] {x, 1 / (x - x of the node before)} of the disk / optical grid, two virtual
                              // nodes far to the red and padding behind the real ones (lane_grid_rows)
    int32_t gx_in_smem;       // the walk kernel keeps both in shared memory
    const uint32_t *tab_col;  // [nband + 1] offsets, then [ntab] the band tables of every output column (table = index in group order)
    const double2 *col_span;  // [nband] full wavelength span of every filter (coverage test), by output column
    uint32_t ntab;            // band tables (a filter cut into pieces has several); at most 64
    uint64_t batch0;          // first batch of the chunk (index into perm / 32)
    uint32_t nbatch;          // batches of the chunk
    uint32_t rowsD, rowsB;    // rows per batch of the disk / bulge SED arrays (lane_grid_rows of nD / nB)
    uint64_t rows_stride;     // bytes per batch of `rows`
    void *rows;               // [nbatch][rowsD + rowsB][32 lanes] reals: rest-frame SEDs, node-major
    uint32_t *win;            // [nbatch][64][2] node window of every band table on the disk / optical grid (0: nothing to do)
    int4 *plan;               // [nbatch] rows to build: disk first / last, bulge first / last
    double *res;              // [nbatch][ntab][2][32 lanes] table sums (bulge, disk)
    uint32_t *counters;       // [kMaxGroups] items handed out per band group
    float group_cost[kMaxGroups]; // relative work of the groups (deals the CTAs to the groups)
    // node sets `sed` / `filter` of [vif] sed2flux (EGGPHOT_NODES_*; 0 = merged: the walk kernel): the filters as they are
    uint32_t nodes_mode;
    const double *alt_fl, *alt_fr; // wavelengths / responses of the sorted bands, concatenated
    const uint32_t *alt_off;       // [nband + 1]
    uint32_t item_mode;       // order of a group's (batch, table) items: 0 batches outermost, 1 the heavier half of the tables for
                              // every batch and then the lighter half, 2 tables outermost (heaviest first)
};
size_t lane_kernel_smem(const LaneParams &q, int precision);
// plan + build + walk + finish of one chunk; returns cudaError_t as int
// (w0 / w1: optional events recorded on `stream` right before / after the walk kernel, the dominant one, for the bench's roofline)
int launch_lane_chunk(const LaneParams &q, int precision, int num_sms, cudaStream_t stream, cudaEvent_t w0 = nullptr, cudaEvent_t w1 = nullptr);
// counting sort of the galaxies by log2(1+z): hist is scratch of zsort_hist_bytes(), perm receives the order
size_t zsort_hist_bytes();
int launch_zsort(const float *z, uint32_t ngal, uint32_t *hist, uint32_t *perm, int num_sms, cudaStream_t stream);

// FMA micro-benchmark (scratch: at least 16 bytes of device memory); returns cudaError_t as int
int measure_fma_peak(int num_sms, void *scratch, cudaStream_t stream, double *fp32_tflops, double *fp64_tflops);

// save_sed: observed-frame lam / sed (f32) of galaxies [first, first+count) of one component
struct SedParams {
    FluxParams fp;
    uint64_t first, count;
    int32_t component; // 0 bulge, 1 disk
    float *lam, *sed;  // row g of either starts at g * row_stride floats (n for two [count][n] arrays, 2 n for the packed
    uint64_t row_stride; // [count][lam | sed] layout of the save_sed file with sed = lam + n)
};
int launch_sed_kernel(const SedParams &p, int precision, cudaStream_t stream);

} // namespace eggphot
#endif
