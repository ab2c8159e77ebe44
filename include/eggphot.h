/*
 * eggphot.h — C ABI of the B200-native photometry engine for egg-gencat.
 *
 * The reference (cschreib/egg) has no plugin/FFI layer: its photometry stage is the block
 * `if (!no_flux) {...}` of src/egg-gencat.cpp:2024-2240, i.e. the closure
 *
 *     get_flux(const vec1f& m, const vec1u& optsed, const vec2f& llum,
 *              vec2f& flux, vec2f& rfmag, bool no_ir, component cmp)   (src/egg-gencat.cpp:2097-2098)
 *
 * called once for bulges and once for disks (src/egg-gencat.cpp:2206-2212) followed by the totals
 * (src/egg-gencat.cpp:2238-2239), on libraries and filters prepared at src/egg-gencat.cpp:316-541.
 * This header is the drop-in boundary for exactly that block: plain pointers and sizes, the
 * reference's own column names, types and option spellings; no torch / C++ types.
 *
 * Ownership: the caller owns every array; the context owns device copies of the prepared
 * libraries and filters.  One context per GPU; a context may be used from one host thread at a
 * time; different contexts are independent.  There is NO CPU fallback: every compute entry point
 * fails with EGGPHOT_ERR_CUDA when no sm_100 device is usable.
 *
 * Data conditions are not errors, exactly as in the reference: a band the SED does not cover
 * yields flux 0 (src/egg-gencat.cpp:2198), a non-finite magnitude yields +99 (:2156).
 */
#ifndef EGGPHOT_H
#define EGGPHOT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGGPHOT_VERSION 100

/* status codes (0 = ok).  User errors mirror the reference's `error(...); return 1;` sites. */
enum {
    EGGPHOT_OK = 0,
    EGGPHOT_ERR_ARG = 1,         /* null pointer / bad size / unknown option value            */
    EGGPHOT_ERR_FILTER = 2,      /* empty, non-monotonic or unusable filter (gencat:334)      */
    EGGPHOT_ERR_LIBRARY = 3,     /* no valid optical SED (gencat:537-541), unknown IMF (:464) */
    EGGPHOT_ERR_GRID = 4,        /* templates of one library do not share a wavelength grid,  */
                                 /* or IGM requested on a grid that starts above 0.06 um      */
    EGGPHOT_ERR_UNSUPPORTED = 5, /* option combination not implemented on the device          */
    EGGPHOT_ERR_INDEX = 6,       /* a galaxy points to a template outside the library         */
    EGGPHOT_ERR_CUDA = 7,        /* CUDA runtime failure or no usable device                  */
    EGGPHOT_ERR_IO = 8
};

/* arithmetic mode (BASELINE.json north_star): fp32 with compensated sums (1e-5 per band) or fp64 (1e-10) */
enum { EGGPHOT_FP32C = 0, EGGPHOT_FP64 = 1 };
/* IR library schema: generic LAM/SED per unit L_IR (gencat:412-444) or cs17 LAM/DUST/PAH per unit Mdust (:401-409) */
enum { EGGPHOT_IR_GENERIC = 0, EGGPHOT_IR_CS17 = 1 };
/* [vif] sed2flux quadrature node set; the contract is MERGED */
enum { EGGPHOT_NODES_MERGED = 0, EGGPHOT_NODES_SED = 1, EGGPHOT_NODES_FILTER = 2 };
/* [vif] integrate_gauss convention for emission lines; the contract is INTEGRAL */
enum { EGGPHOT_LINE_INTEGRAL = 0, EGGPHOT_LINE_AVERAGE = 1 };

/* SED libraries as read from the FITS tables, before any preparation (gencat:401-458). */
typedef struct eggphot_libs {
    /* optical library: lam, sed are [nuv][nvj][nopt] f32, use is [nuv][nvj] (0/1) */
    uint32_t nuv, nvj, nopt;
    const float *opt_lam;
    const float *opt_sed;
    const uint8_t *opt_use;
    /* IR library: lam, sed (and pah for cs17) are [nir_sed][nir] f64 */
    int32_t ir_kind; /* EGGPHOT_IR_* */
    uint32_t nir_sed, nir;
    const double *ir_lam;
    const double *ir_sed; /* generic: SED; cs17: DUST */
    const double *ir_pah; /* cs17: PAH; else NULL */
} eggphot_libs;

/* A list of raw filter response curves in the caller's (command-line) order. */
typedef struct eggphot_filters {
    uint32_t nband;
    const uint32_t *npts;     /* [nband] */
    const double *const *lam; /* [nband] -> [npts[b]] wavelength, um */
    const double *const *res; /* [nband] -> [npts[b]] response */
} eggphot_filters;

/* Options; names are the reference's command-line keys (gencat:137-149). */
typedef struct eggphot_opts {
    int32_t precision;        /* EGGPHOT_FP32C | EGGPHOT_FP64 (new key `precision=`)            */
    const char *igm;          /* "none" | "constant" | "madau95" | "inoue14" (gencat:70)        */
    int32_t no_stellar, no_dust, no_nebular; /* gencat:59-65                                     */
    const char *opt_lib_imf;  /* "salpeter" | "chabrier" | "kroupa" (gencat:90)                  */
    double opt_lib_min_step;  /* um, gencat:91 (0.001)                                           */
    int32_t filter_flambda, filter_photons, filter_normalize, trim_filters; /* gencat:96-99       */
    int32_t sed2flux_nodes;   /* EGGPHOT_NODES_* (new key `sed2flux_nodes=`)                     */
    int32_t line_profile;     /* EGGPHOT_LINE_*                                                   */
    uint32_t nline;           /* emission lines (gencat:1834-1838); 0 = none                     */
    const float *line_lambda; /* [nline] rest wavelengths, um                                    */
    int32_t device;           /* CUDA device ordinal                                             */
    uint32_t smem_table_budget; /* bytes of filter tables per band group; 0 = default           */
} eggphot_opts;

/* Per-galaxy columns the path reads, as structure-of-arrays with the reference's names and
 * types (struct `out`, gencat:551-603).  Arrays are length ngal unless noted; uint_t is 64-bit. */
typedef struct eggphot_galaxies {
    const float *z, *d;                /* redshift; luminosity distance [Mpc]                   */
    const float *m_disk, *m_bulge;     /* log10 stellar mass per component (gencat:1381-1387)    */
    const float *m2lcor;               /* gencat:1549                                            */
    const uint64_t *opt_sed_disk, *opt_sed_bulge; /* flat index iuv*nvj+ivj (gencat:709)        */
    const uint64_t *ir_sed;            /* gencat:1706                                            */
    const float *lir;                  /* generic IR library                                     */
    const float *fpah, *mdust;         /* cs17 IR library (may be NULL for generic)              */
    const float *igm_abs;              /* [ngal][3] (gencat:1731-1824); may be NULL if IGM off   */
    const float *vdisp;                /* km/s (gencat:1878); may be NULL if no lines            */
    const float *line_lum_disk;        /* [ngal][nline] Lsun; may be NULL                        */
    const float *line_lum_bulge;       /* [ngal][nline]; NULL == identically zero (gencat:1886+) */
} eggphot_galaxies;

/* Output columns (gencat:2029-2033, 2238-2239); each may be NULL to skip it.
 * flux* are [ngal][nband] uJy, rfmag* are [ngal][nrf] AB mag at 10 pc, band axis in SORTED order. */
typedef struct eggphot_outputs {
    float *flux, *flux_disk, *flux_bulge;
    float *rfmag, *rfmag_disk, *rfmag_bulge;
    /* fp64 copies of the component columns before rounding to the f32 catalog type, for
     * accuracy tests below f32 resolution; normally NULL */
    double *flux_disk_f64, *flux_bulge_f64, *rfmag_disk_f64, *rfmag_bulge_f64;
} eggphot_outputs;

typedef struct eggphot_ctx eggphot_ctx;

/* Prepare filters (adjust + sort, gencat:316-378) and libraries (IMF + rebin, gencat:461-534),
 * build the device tables and upload them.  `rfbands` may be NULL.  On failure *ctx is still a
 * valid object (unless allocation failed) holding the message for eggphot_last_error. */
int eggphot_create(const eggphot_libs *libs, const eggphot_filters *bands, const eggphot_filters *rfbands,
                   const eggphot_opts *opts, eggphot_ctx **ctx);
void eggphot_destroy(eggphot_ctx *ctx);
const char *eggphot_last_error(const eggphot_ctx *ctx);

/* Band bookkeeping after the sort of gencat:367-378: order[k] = caller index of the k-th output
 * column, lambda[k] = its reference wavelength (vec1f `lambda`, gencat:368).  which: 0 bands, 1 rfbands. */
uint32_t eggphot_nband(const eggphot_ctx *ctx, int which);
int eggphot_band_order(const eggphot_ctx *ctx, int which, uint32_t *order, float *lambda);

/* The two get_flux passes + totals for ngal galaxies whose columns live in HOST memory
 * (pinned or pageable).  Copies in, computes on the device in chunks overlapped with the copies,
 * copies out.  This is the call a replacement of gencat:2024-2240 makes. */
int eggphot_compute(eggphot_ctx *ctx, const eggphot_galaxies *gal, size_t ngal, const eggphot_outputs *out);

/* Same computation for columns already resident in DEVICE memory of ctx's device; enqueues on
 * `stream` (a cudaStream_t; NULL = the CUDA legacy default stream) and returns without
 * synchronising.  Calls on one context must not overlap each other on the device (they share the
 * context's SED scratch): enqueue them on one stream. */
int eggphot_compute_device(eggphot_ctx *ctx, const eggphot_galaxies *gal, size_t ngal,
                           const eggphot_outputs *out, void *stream);
int eggphot_synchronize(eggphot_ctx *ctx);

/* save_sed (gencat:2164-2193): observed-frame lam then sed, f32, of one component of galaxies
 * [first, first+count) into host buffers of count*eggphot_sed_size() floats each.  component: 0 bulge, 1 disk. */
uint32_t eggphot_sed_size(const eggphot_ctx *ctx, int component);
int eggphot_get_seds(eggphot_ctx *ctx, const eggphot_galaxies *gal, size_t first, size_t count, int component,
                     float *lam, float *sed);
/* The same spectra in the layout of the save_sed file (gencat:2164-2193, read by egg-getsed): per galaxy lam then sed,
 * out = [count][2 * eggphot_sed_size()] floats in host memory (pinned memory makes the copy asynchronous).  _begin
 * enqueues one range and returns; two ranges may be in flight (the copy out of one overlaps the kernel of the next).
 * _wait(ctx, 0) returns when the range begun before the last one has landed in its buffer, _wait(ctx, 1) when all have
 * (and reports data errors). */
int eggphot_get_seds_packed_begin(eggphot_ctx *ctx, const eggphot_galaxies *gal, size_t first, size_t count, int component,
                                  float *out);
int eggphot_get_seds_packed_wait(eggphot_ctx *ctx, int which);
/* cudaHostRegister / cudaHostUnregister of a caller's buffer, for programs that do not link the CUDA runtime themselves:
 * pinned host columns make the copies of eggphot_compute asynchronous.  Failure is not fatal (the copies are then staged). */
int eggphot_host_register(void *ptr, size_t bytes);
int eggphot_host_unregister(void *ptr);


/* Rest wavelengths [um] of the prepared SED grid of one component (0 bulge = the optical library's grid after
 * re-binning, 1 disk = the merge_add union with the IR grid); x receives eggphot_sed_size(ctx, component) values. */
int eggphot_get_grid(const eggphot_ctx *ctx, int component, double *x);

/* Introspection for benchmarks: kernels launched by the last compute call, number of band groups,
 * and the algorithmic work of the last call (quadrature nodes). */
typedef struct eggphot_stats {
    uint64_t kernel_launches;
    uint32_t ngroups;
    uint32_t nodes_disk, nodes_bulge, nodes_ir; /* grid sizes after preparation */
    uint64_t table_bytes;                        /* device bytes of prepared tables */
    double last_kernel_ms;                       /* mean device time [ms] of the walk kernel (the dominant one) per
                                                  * eggphot_compute_device call since the previous eggphot_get_stats, from CUDA
                                                  * events on the launching stream; call after the stream has been synchronised */
} eggphot_stats;
int eggphot_get_stats(const eggphot_ctx *ctx, eggphot_stats *stats);

/* Measured peak of the CUDA-core FMA pipes of ctx's device, the roofline this quadrature is compared with (SURVEY.md
 * section 8d asks for a measured, not a nominal, denominator): independent FMA chains on every SM timed with CUDA
 * events, best of five launches.  TFLOP/s counting 2 flops per FMA. */
int eggphot_measure_fma_peak(eggphot_ctx *ctx, double *fp32_tflops, double *fp64_tflops);

/* The `maglim` mass-limit estimator of egg-gencat (src/egg-gencat.cpp:903-969): for every redshift slice, 1000 trial
 * masses from mmin_strict to mmax get the colours of an active galaxy without scatter (gen_blue_uvj(..., true), :935), the
 * optical template of those colours (get_sed_uvj, :939), the mass-to-light correction of the slice (:941) and their flux
 * in the selection band (get_opt_sed / lsun2uJy / sed2flux, :943-952: stellar light only, no IGM, no lines, no dust);
 * z_mmin = max(mmin_strict, first trial mass whose flux exceeds the limit - 0.2), mmax if none does (:955-962).
 *   phot : a context over the optical library with bands = {selection band}, no_dust, no_nebular, igm = "none";
 *   props: a property context (include/eggprops.h) over the same library, created with no_random;
 *   zb_lo, zb_hi [nz]: the redshift slices ([vif] bin_center = their mean); ntrial: 1000 in the reference;
 *   z_mmin [nz] receives the limits (the vec1f of :901). */
struct eggprops_ctx;
int eggphot_maglim(eggphot_ctx *phot, struct eggprops_ctx *props, const double *zb_lo, const double *zb_hi, uint32_t nz,
                   double maglim, double mmin_strict, double mmax, uint32_t ntrial, float *z_mmin);

#ifdef __cplusplus
}
#endif
#endif /* EGGPHOT_H */
