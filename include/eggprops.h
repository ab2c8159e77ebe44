/*
 * eggprops.h — C ABI of the per-galaxy property recipes of egg-gencat on the GPU.
 *
 * Replaces the block of src/egg-gencat.cpp:1360-2022: from (z, M*, passive) every other catalog column is
 * computed with element-wise formulas plus random scatter — bulge-to-total ratio and component masses
 * (:1367-1387), morphology (:1398-1449), SFR / RSB (:1459-1484), IRX (:1488-1506), UVJ colours (:1515-1546,
 * lambdas at :619-668), M/L correction (:1549, :544), optical SED indices (:1554-1571, :670-705), dust
 * temperature / IR8 / L_IR / f_PAH / M_dust and the IR SED index (:1584-1725), IGM transmission (:1731-1824) and
 * emission lines (:1829-2022).  The columns it writes are, with the reference's names and types, the inputs of
 * the photometry path (eggphot.h, eggphot_galaxies): with the *_device entry point they never leave the GPU.
 *
 * Random numbers.  The reference consumes one sequential std::mt19937 stream in array order ([vif] randomn,
 * randomu, random_coin, random_pdf); a drop-in cannot and need not reproduce its realisations (the catalog is a
 * random draw; the reference's own output changes with `seed` and with the catalog order).  Here every draw is
 * Philox4x32-10(key = seed, counter = (galaxy index, draw slot)): a galaxy's columns depend only on
 * (seed, first_index + i, z, m, passive) — reproducible across chunkings, GPUs and ranks.
 *
 * Not covered (reported as EGGPHOT_ERR_UNSUPPORTED where reachable): user-provided LIR / Tdust / IR8 columns of an
 * input catalog (:1639-1663, an O(n^2) nearest-neighbour swap), IR libraries without a calibration (:1702).
 * There is NO CPU fallback.  A context belongs to one GPU and may be used from one host thread at a time.
 */
#ifndef EGGPROPS_H
#define EGGPROPS_H

#include <stddef.h>
#include <stdint.h>

#include "eggphot.h" /* status codes */

#ifdef __cplusplus
extern "C" {
#endif

/* IR library calibration (src/egg-gencat.cpp:1680-1704, chosen there by the library's file name) */
enum { EGGPROPS_IR_SINGLE = 0, EGGPROPS_IR_CE01 = 1, EGGPROPS_IR_M12 = 2, EGGPROPS_IR_CS17 = 3 };
enum { EGGPROPS_IGM_NONE = 0, EGGPROPS_IGM_MADAU95 = 1, EGGPROPS_IGM_INOUE14 = 2 };
#define EGGPROPS_NLINE 20 /* src/egg-gencat.cpp:1834-1838 */

/* The parts of the libraries the recipes read (host pointers). */
typedef struct eggprops_libs {
    /* optical library: UVJ bins and usable templates (gencat:670-705) */
    uint32_t nuv, nvj;
    const float *buv;   /* [2][nuv] lower and upper edges of the U-V bins */
    const float *bvj;   /* [2][nvj] */
    const uint8_t *use; /* [nuv][nvj] */
    const float *av;    /* [nuv*nvj] */
    /* IR library */
    int32_t ir_lib;     /* EGGPROPS_IR_* */
    uint32_t nir_sed;
    const double *cs17_tdust, *cs17_lir_dust, *cs17_lir_pah, *cs17_l8_dust, *cs17_l8_pah; /* [nir_sed], cs17 only */
} eggprops_libs;

/* Options; names are the reference's command-line keys (gencat:56-149). */
typedef struct eggprops_opts {
    uint64_t seed;
    int32_t no_random, no_stellar, no_nebular, no_passive_lir, magdis_tdust;
    double ms_disp;        /* 0.3 */
    int32_t igm;           /* EGGPROPS_IGM_* */
    double h0, omega_m, omega_l; /* 70, 0.3, 0.7 */
    int32_t device;
} eggprops_opts;

/* Output columns with the reference's names (struct `out`, gencat:551-603); each may be NULL to skip it.
 * Length ngal unless noted.  The same struct serves host and device pointers. */
typedef struct eggprops_out {
    float *d;                                   /* luminosity distance [Mpc] (the `d` of the photometry path)  */
    float *bt, *m_bulge, *m_disk;
    float *disk_angle, *bulge_angle, *disk_ratio, *bulge_ratio, *disk_radius, *bulge_radius;
    float *sfr, *rsb, *irx, *sfrir, *sfruv;
    float *rfuv_disk, *rfvj_disk, *rfuv_bulge, *rfvj_bulge;
    float *m2lcor;
    uint64_t *opt_sed_disk, *opt_sed_bulge;
    float *av_disk, *av_bulge;
    float *tdust, *ir8, *lir, *fpah, *mdust;
    uint64_t *ir_sed;
    float *igm_abs;                             /* [ngal][3] */
    float *avlines_disk, *avlines_bulge, *vdisp;
    float *line_lum_disk, *line_lum_bulge;      /* [ngal][EGGPROPS_NLINE] */
} eggprops_out;

typedef struct eggprops_ctx eggprops_ctx;

int eggprops_create(const eggprops_libs *libs, const eggprops_opts *opts, eggprops_ctx **ctx);
void eggprops_destroy(eggprops_ctx *ctx);
const char *eggprops_last_error(const eggprops_ctx *ctx);

/* Host arrays in, host arrays out (copies inside).  passive: 0/1 bytes.  first_index: index of the first galaxy of
 * this call in the whole catalog (shards, chunks): it selects the random streams. */
int eggprops_compute(eggprops_ctx *ctx, const float *z, const float *m, const uint8_t *passive, size_t ngal,
                     uint64_t first_index, const eggprops_out *out);
/* Device pointers in and out, enqueued on `stream` (a cudaStream_t passed as void *); asynchronous. */
int eggprops_compute_device(eggprops_ctx *ctx, const float *z, const float *m, const uint8_t *passive, size_t ngal,
                            uint64_t first_index, const eggprops_out *out, void *stream);
/* rest wavelengths [um] and names of the EGGPROPS_NLINE lines, in column order */
const float *eggprops_line_lambda(void);
const char *eggprops_line_name(uint32_t l);

#ifdef __cplusplus
}
#endif
#endif
