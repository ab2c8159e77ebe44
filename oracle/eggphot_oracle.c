/*
 * eggphot_oracle.c — CPU restatement of egg-gencat's photometry path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA product in
 * egg_b200/csrc; nothing under egg_b200/ links, loads or calls it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it.
 *
 * PARITY UNPINNED at the vif boundary: the reference (cschreib/egg) delegates
 * the arithmetic of this path to the third-party library cschreib/vif, pinned
 * only as "master" (doc/scripts/install.sh:44) and absent from /root/reference,
 * and the reference ships no tests or golden catalogs for the path.  The vif
 * functions are therefore restated from their published behaviour and from the
 * reference's own call sites; each such restatement is marked [vif] below.  The
 * one known-answer test the reference's documentation offers (list_bands
 * ref-lam / FWHM for 17 filters, doc/doc-egg-gencat.tex:32-55) pins
 * orc_trapz + the filter reader and is checked in tests/test_oracle_kat.py.
 *
 * All arithmetic is IEEE double.  Inputs keep the reference's storage types
 * (f32 catalog columns and optical library, f64 IR library and filters).  The
 * reference stores its per-galaxy temporaries in vec1f; with emulate_f32 != 0
 * the oracle rounds to float at those same places so that the size of that
 * effect (a few 1e-7 relative) can be measured; the contract used for parity is
 * emulate_f32 = 0.
 *
 * File:line citations are relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_NPOS ((int64_t)-1)

/* sed2flux node sets: the default contract is A. */
enum { ORC_NODES_MERGED = 0, ORC_NODES_SED = 1, ORC_NODES_FILTER = 2 };
/* integrate_gauss conventions: bin integral (default contract) or bin average. */
enum { ORC_GAUSS_INTEGRAL = 0, ORC_GAUSS_AVERAGE = 1 };

typedef struct {
    /* optical library after orc_opt_lib_prepare: [nopt_sed][nopt], flat index = iuv*nvj+ivj
     * (mult_ids / flat_id are C order, doc/doc-egg-gencat.tex:134) */
    int64_t nopt_sed, nopt;
    const float *opt_lam, *opt_sed;
    /* IR library: kind 0 = generic LAM/SED per unit L_IR (src/egg-gencat.cpp:412-444),
     *             kind 1 = cs17 LAM/DUST/PAH per unit Mdust (src/egg-gencat.cpp:401-409) */
    int32_t ir_kind;
    int64_t nir_sed, nir;
    const double *ir_lam, *ir_sed, *ir_pah;
    /* filters, already adjusted and sorted by reference wavelength (src/egg-gencat.cpp:316-378) */
    int32_t nband;
    const int64_t *foff; /* [nband+1] offsets into flam/fres */
    const double *flam, *fres;
    int32_t nrf;
    const int64_t *rfoff;
    const double *rflam, *rfres;
    /* emission lines (src/egg-gencat.cpp:1834-1838) */
    int32_t nline;
    const float *line_lambda;
    /* options */
    int32_t apply_igm;   /* igm != "none" && igm != "constant" (src/egg-gencat.cpp:2122) */
    int32_t no_stellar;  /* src/egg-gencat.cpp:2107 */
    int32_t no_nebular;  /* src/egg-gencat.cpp:2143 */
    int32_t nodes_mode;  /* ORC_NODES_* */
    int32_t gauss_mode;  /* ORC_GAUSS_* */
    int32_t emulate_f32; /* round temporaries to float where the reference stores vec1f */
} orc_setup;

typedef struct {
    int64_t ngal;
    const float *z, *d, *m2lcor, *lir, *fpah, *mdust;
    const float *igm_abs; /* [ngal][3] */
    const float *vdisp;
    const uint64_t *ir_sed;
} orc_galaxies;

static inline double st32(int emu, double x) { return emu ? (double)(float)x : x; }

/* [vif] integrate(x, y): plain trapezoid.  Verified by KAT-1 (src/egg-gencat.cpp:195). */
double orc_trapz(int64_t n, const double *x, const double *y) {
    double s = 0.0;
    for (int64_t i = 0; i + 1 < n; ++i) s += 0.5 * (y[i] + y[i + 1]) * (x[i + 1] - x[i]);
    return s;
}

/* [vif] interpolate(y1, y2, x1, x2, x) — argument order as at src/egg-utils.hpp:36-38 */
static inline double lerp2(double y1, double y2, double x1, double x2, double x) {
    return y1 + (y2 - y1) * (x - x1) / (x2 - x1);
}

/* [vif] integrate(x, y, x0, x1): trapezoid restricted to [x0,x1], end points interpolated
 * (used as a box average at src/egg-gencat.cpp:526-527). */
double orc_trapz_range(int64_t n, const double *x, const double *y, double x0, double x1) {
    if (n < 2 || x1 <= x0) return 0.0;
    if (x0 < x[0]) x0 = x[0];
    if (x1 > x[n - 1]) x1 = x[n - 1];
    if (x1 <= x0) return 0.0;
    int64_t i = 0;
    while (i + 2 < n && x[i + 1] <= x0) ++i;
    double px = x0, py = lerp2(y[i], y[i + 1], x[i], x[i + 1], x0), s = 0.0;
    for (int64_t k = i + 1; k < n; ++k) {
        if (x[k] >= x1) {
            double ey = lerp2(y[k - 1], y[k], x[k - 1], x[k], x1);
            s += 0.5 * (py + ey) * (x1 - px);
            return s;
        }
        s += 0.5 * (py + y[k]) * (x[k] - px);
        px = x[k];
        py = y[k];
    }
    return s;
}

/* [vif] lower_bound(v, x): index of the LAST element <= x, npos if none. */
static int64_t lower_bound_d(int64_t n, const double *v, double x) {
    if (n == 0 || !(v[0] <= x)) return ORC_NPOS;
    int64_t lo = 0, hi = n; /* v[lo] <= x, (hi == n or v[hi] > x) */
    while (hi - lo > 1) {
        int64_t mid = lo + (hi - lo) / 2;
        if (v[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}
/* [vif] upper_bound(v, x): index of the FIRST element > x, npos if none. */
static int64_t upper_bound_d(int64_t n, const double *v, double x) {
    if (n == 0 || !(v[n - 1] > x)) return ORC_NPOS;
    int64_t lo = -1, hi = n - 1; /* v[hi] > x, (lo == -1 or v[lo] <= x) */
    while (hi - lo > 1) {
        int64_t mid = lo + (hi - lo) / 2;
        if (v[mid] > x) hi = mid; else lo = mid;
    }
    return hi;
}

/* Filter adjustment of src/egg-gencat.cpp:330-364 (only applied by the caller when one of
 * trim/flambda/photons is set, as at :361).  In place; returns the new length, writes rlam. */
int64_t orc_adjust_filter(int64_t n, double *lam, double *res, int trim, int flambda, int photons,
                          int normalize, double *rlam) {
    if (trim) { /* :334-340 */
        double mx = res[0];
        for (int64_t i = 1; i < n; ++i) if (res[i] > mx) mx = res[i];
        int64_t a = -1, b = -1;
        for (int64_t i = 0; i < n; ++i) if (res[i] / mx > 1e-3) { if (a < 0) a = i; b = i; }
        if (a < 0) return -1;
        memmove(lam, lam + a, (size_t)(b - a + 1) * sizeof(double));
        memmove(res, res + a, (size_t)(b - a + 1) * sizeof(double));
        n = b - a + 1;
    }
    if (flambda) for (int64_t i = 0; i < n; ++i) res[i] /= lam[i] * lam[i]; /* :343-347 */
    if (photons) for (int64_t i = 0; i < n; ++i) res[i] *= lam[i];          /* :348-352 */
    double *tmp = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t i = 0; i < n; ++i) tmp[i] = lam[i] * res[i];
    double norm = orc_trapz(n, lam, res); /* :355 */
    *rlam = orc_trapz(n, lam, tmp) / norm; /* :356 */
    free(tmp);
    if (normalize) for (int64_t i = 0; i < n; ++i) res[i] /= norm; /* :357-359 */
    return n;
}

/* [vif] get_filter: rlam = integrate(lam, lam*res) on the raw response (no renormalisation; KAT-1). */
double orc_filter_rlam(int64_t n, const double *lam, const double *res) {
    double s = 0.0;
    for (int64_t i = 0; i + 1 < n; ++i)
        s += 0.5 * (lam[i] * res[i] + lam[i + 1] * res[i + 1]) * (lam[i + 1] - lam[i]);
    return s;
}

/* Optical library preparation: IMF factor (src/egg-gencat.cpp:461-466) and re-binning where the
 * grid is finer than min_step (src/egg-gencat.cpp:469-534).  Input  lam/sed [nsed][nin] f32,
 * use [nsed].  Output buffers must hold nsed*nin floats; returns the new number of nodes
 * (== nin when nothing is rebinned).  imf_shift = 1 for chabrier/kroupa. */
int64_t orc_opt_lib_prepare(int64_t nsed, int64_t nin, const float *lam, const float *sed,
                            const uint8_t *use, int imf_shift, double min_step,
                            float *lam_out, float *sed_out) {
    const double imf = imf_shift ? pow(10.0, -0.2) : 1.0;
    int64_t first = -1;
    for (int64_t s = 0; s < nsed; ++s) if (use[s]) { first = s; break; }
    if (first < 0) return -1; /* :537-541 */
    /* sed *= e10(-0.2) is a vec3f *= double: float result */
    float *fsed = (float *)malloc((size_t)(nsed * nin) * sizeof(float));
    for (int64_t k = 0; k < nsed * nin; ++k) fsed[k] = imf_shift ? (float)(sed[k] * imf) : sed[k];

    int64_t id1 = -1, id2 = -1;
    double *olam = (double *)malloc((size_t)nin * sizeof(double));
    for (int64_t l = 0; l < nin; ++l) olam[l] = lam[first * nin + l]; /* :475-476 */
    if (min_step > 0 && nin >= 3) {
        for (int64_t l = 0; l < nin; ++l) { /* :480-484 */
            double dl = l == 0 ? olam[1] - olam[0]
                      : l == nin - 1 ? olam[nin - 1] - olam[nin - 2]
                      : 0.5 * (olam[l + 1] - olam[l - 1]);
            if (dl < min_step) { if (id1 < 0) id1 = l; id2 = l; } /* :487-488 */
        }
    }
    if (id1 < 0) {
        memcpy(lam_out, lam, (size_t)(nsed * nin) * sizeof(float));
        memcpy(sed_out, fsed, (size_t)(nsed * nin) * sizeof(float));
        free(fsed); free(olam);
        return nin;
    }
    /* new grid: head copied, uniform middle, tail copied (:491-508) */
    int64_t nnew = (int64_t)floor((olam[id2] - olam[id1]) / min_step) - 1;
    if (nnew < 0) nnew = 0;
    int64_t ntail = nin - 1 - id2;
    int64_t nout = id1 + nnew + ntail;
    double *nlam = (double *)malloc((size_t)nout * sizeof(double));
    for (int64_t l = 0; l < id1; ++l) nlam[l] = olam[l];
    for (int64_t k = 0; k < nnew; ++k) nlam[id1 + k] = olam[id1] + 0.5 * min_step + min_step * (double)k;
    for (int64_t l = 0; l < ntail; ++l) nlam[id1 + nnew + l] = olam[id2 + 1 + l];
    double *row = (double *)malloc((size_t)nin * sizeof(double));
    memset(lam_out, 0, (size_t)(nsed * nout) * sizeof(float));
    memset(sed_out, 0, (size_t)(nsed * nout) * sizeof(float));
    for (int64_t s = 0; s < nsed; ++s) {
        if (!use[s]) continue; /* :515-519: unused rows stay zero */
        for (int64_t l = 0; l < nin; ++l) row[l] = fsed[s * nin + l]; /* osed is vec3d of the f32 lib */
        for (int64_t l = 0; l < nout; ++l) lam_out[s * nout + l] = (float)nlam[l];
        for (int64_t l = 0; l < id1; ++l) sed_out[s * nout + l] = (float)row[l];
        for (int64_t k = 0; k < nnew; ++k) { /* :525-528 box average */
            double c = nlam[id1 + k];
            sed_out[s * nout + id1 + k] =
                (float)(orc_trapz_range(nin, olam, row, c - 0.5 * min_step, c + 0.5 * min_step) / min_step);
        }
        for (int64_t l = 0; l < ntail; ++l) sed_out[s * nout + id1 + nnew + l] = (float)row[id2 + 1 + l];
    }
    free(row); free(nlam); free(fsed); free(olam);
    return nout;
}

/* merge_add (src/egg-utils.hpp:5-63): union grid of two sorted curves; where a node of one
 * curve lies before the other's first node or after its last, only its own value counts. */
static int64_t merge_add(int64_t n1, const double *x1, const double *y1,
                         int64_t n2, const double *x2, const double *y2,
                         double *x, double *y, int emu) {
    int64_t i1 = 0, i2 = 0, n = 0;
    while (i1 < n1 || i2 < n2) {
        int take1 = (i2 == n2) || (i1 < n1 && x1[i1] <= x2[i2]);
        int take2 = (i1 == n1) || (i2 < n2 && x2[i2] <= x1[i1]);
        if (take1 && take2) { /* equal abscissae: one node, summed (:55-60) */
            x[n] = x1[i1]; y[n] = st32(emu, y1[i1] + y2[i2]); ++i1; ++i2;
        } else if (take1) {
            double v = y1[i1];
            if (i2 != n2 && i2 != 0) /* inside curve 2's span (:33-38); past its end: own value (:24-27) */
                v += st32(emu, lerp2(y2[i2 - 1], y2[i2], x2[i2 - 1], x2[i2], x1[i1]));
            x[n] = x1[i1]; y[n] = st32(emu, v); ++i1;
        } else {
            double v = y2[i2];
            if (i1 != n1 && i1 != 0)
                v += st32(emu, lerp2(y1[i1 - 1], y1[i1], x1[i1 - 1], x1[i1], x2[i2]));
            x[n] = x2[i2]; y[n] = st32(emu, v); ++i2;
        }
        ++n;
    }
    return n;
}

/* exported for unit tests */
int64_t orc_merge_add(int64_t n1, const double *x1, const double *y1, int64_t n2, const double *x2,
                      const double *y2, double *x, double *y) {
    return merge_add(n1, x1, y1, n2, x2, y2, x, y, 0);
}

/* [vif] integrate_gauss(x0, x1, mu, sigma, amp): Gaussian of total area amp integrated over
 * [x0,x1] (ORC_GAUSS_INTEGRAL, the contract) or averaged over it (ORC_GAUSS_AVERAGE). */
static inline double integrate_gauss(double x0, double x1, double mu, double sigma, double amp, int mode) {
    const double is = 1.0 / (M_SQRT2 * sigma);
    double v = 0.5 * amp * (erf((x1 - mu) * is) - erf((x0 - mu) * is));
    return mode == ORC_GAUSS_AVERAGE ? v / (x1 - x0) : v;
}

/* [vif] lsun2uJy(z, d, lam, lum): rest nu.L_nu [Lsun] at rest wavelength lam [um] -> observed
 * S_nu [uJy] at luminosity distance d [Mpc]. */
static const double ORC_LSUN2UJY = 1.0e32 * 3.839e26 / (2.9979e14 * 4.0 * M_PI * 3.0856e22 * 3.0856e22);
static inline double lsun2uJy(double z, double d, double lam, double lum) {
    return ORC_LSUN2UJY * (1.0 + z) * lam * lum / (d * d);
}
double orc_lsun2uJy_factor(void) { return ORC_LSUN2UJY; }

/* [vif] uJy2mag / mag2uJy: AB zero point 23.9 (doc/doc-egg-2skymaker.tex:49). */
static inline double uJy2mag(double f) { return -2.5 * log10(f) + 23.9; }
static inline double mag2uJy(double m) { return pow(10.0, 0.4 * (23.9 - m)); }

/* [vif] sed2flux(filter, lam, sed): integral of (piece-wise linear R) x (piece-wise linear S_nu)
 * over the filter span, trapezoid on the node set chosen by `mode`; NaN if the SED does not
 * cover the filter. */
double orc_sed2flux(int64_t nf, const double *fl, const double *fr, int64_t n, const double *lam,
                    const double *sed, int mode) {
    if (nf < 2 || n < 2) return NAN;
    if (!(lam[0] <= fl[0] && lam[n - 1] >= fl[nf - 1])) return NAN;
    double acc = 0.0;
    if (mode == ORC_NODES_FILTER) { /* C: filter nodes only */
        int64_t k = lower_bound_d(n, lam, fl[0]);
        double pg = 0.0, px = fl[0];
        for (int64_t i = 0; i < nf; ++i) {
            while (k + 2 < n && lam[k + 1] <= fl[i]) ++k;
            double s = lerp2(sed[k], sed[k + 1], lam[k], lam[k + 1], fl[i]);
            double g = fr[i] * s;
            if (i > 0) acc += 0.5 * (pg + g) * (fl[i] - px);
            pg = g; px = fl[i];
        }
        return acc;
    }
    /* A and B walk forward from the filter's first node */
    int64_t k = lower_bound_d(n, lam, fl[0]); /* lam[k] <= fl[0] */
    if (k >= n - 1) k = n - 2;
    int64_t fi = 0;                           /* fl[fi] <= cursor */
    double px = fl[0];
    double pg = fr[0] * lerp2(sed[k], sed[k + 1], lam[k], lam[k + 1], px);
    const double xend = fl[nf - 1];
    int64_t si = k + 1; /* next SED node strictly after the start ... */
    while (si < n && lam[si] <= px) ++si;
    int64_t fn = 1;     /* next filter node */
    while (px < xend) {
        double xs = si < n ? lam[si] : INFINITY;
        double xf, x;
        if (mode == ORC_NODES_SED) xf = xend; /* B: SED nodes + the two filter end points */
        else xf = fl[fn];
        int adv_s = 0, adv_f = 0;
        if (xs < xf) { x = xs; adv_s = 1; }
        else if (xf < xs) { x = xf; adv_f = 1; }
        else { x = xf; adv_s = adv_f = 1; }
        /* response at x */
        double r, s;
        if (adv_f && mode != ORC_NODES_SED) r = fr[fn];
        else {
            while (fi + 2 < nf && fl[fi + 1] <= x) ++fi;
            r = (x >= xend) ? fr[nf - 1] : lerp2(fr[fi], fr[fi + 1], fl[fi], fl[fi + 1], x);
        }
        if (adv_s) s = sed[si];
        else {
            int64_t j = si - 1; /* lam[j] <= x < lam[si] */
            if (j < 0) j = 0;
            int64_t j1 = j + 1 < n ? j + 1 : j;
            s = j1 == j ? sed[j] : lerp2(sed[j], sed[j1], lam[j], lam[j1], x);
        }
        double g = r * s;
        acc += 0.5 * (pg + g) * (x - px);
        px = x; pg = g;
        if (adv_s) ++si;
        if (adv_f && mode != ORC_NODES_SED) { fi = fn; ++fn; if (fn >= nf) break; }
    }
    return acc;
}

/* add_emission_line (src/egg-gencat.cpp:2078-2095) */
static void add_emission_line(int64_t n, const double *lam, double *sed, double l0, double lum,
                              double vdisp, int gauss_mode, int emu) {
    double lw = l0 * (vdisp / 2.998e5);
    int64_t b0 = lower_bound_d(n, lam, l0 - 10 * lw);
    int64_t b1 = upper_bound_d(n, lam, l0 + 10 * lw);
    if (b1 == 0 || b0 == n - 1) return;
    if (b0 == ORC_NPOS) b0 = 0;
    if (b1 == ORC_NPOS) b1 = n - 1;
    for (int64_t l = b0; l <= b1; ++l) {
        double llow = l != 0 ? 0.5 * st32(emu, lam[l - 1] + lam[l]) : lam[l] - 0.5 * st32(emu, lam[l + 1] - lam[l]);
        double lup = l != n - 1 ? 0.5 * st32(emu, lam[l] + lam[l + 1]) : lam[l] + 0.5 * st32(emu, lam[l] - lam[l - 1]);
        llow = st32(emu, llow); lup = st32(emu, lup);
        sed[l] = st32(emu, sed[l] + integrate_gauss(llow, lup, l0, lw, lum * l0, gauss_mode));
    }
}

typedef struct { double *rlam, *rsed, *olam, *osed, *ilam, *ised, *lam, *sed; } scratch_t;

/* Builds the rest-frame SED of one galaxy component exactly as get_flux does
 * (src/egg-gencat.cpp:2102-2147); returns the node count. */
static int64_t build_rest_sed(const orc_setup *S, const orc_galaxies *G, int64_t i, const float *m,
                              const uint64_t *optsed, const float *llum, int no_ir, scratch_t *w) {
    const int emu = S->emulate_f32;
    int64_t n;
    /* get_opt_sed (src/egg-gencat.cpp:708-712); the mass argument is the float difference
     * m[i]-out.m2lcor[i] of two vec1f elements (src/egg-gencat.cpp:2106, 2114) */
    const double mass = (double)(float)(m[i] - G->m2lcor[i]);
    const double scale = pow(10.0, mass);
    int64_t no = 0, ni = 0;
    if (no_ir || !S->no_stellar) {
        const float *ol = S->opt_lam + (int64_t)optsed[i] * S->nopt;
        const float *os = S->opt_sed + (int64_t)optsed[i] * S->nopt;
        no = S->nopt;
        for (int64_t l = 0; l < no; ++l) { w->olam[l] = ol[l]; w->osed[l] = st32(emu, scale * os[l]); }
    }
    if (!no_ir) { /* get_ir_sed (src/egg-gencat.cpp:2039-2057) */
        const int64_t s = (int64_t)G->ir_sed[i];
        ni = S->nir;
        const double *il = S->ir_lam + s * ni;
        if (S->ir_kind == 1) {
            const double *du = S->ir_sed + s * ni, *pa = S->ir_pah + s * ni;
            const double fp = G->fpah[i], md = G->mdust[i];
            for (int64_t l = 0; l < ni; ++l) {
                w->ilam[l] = st32(emu, il[l]);
                w->ised[l] = st32(emu, st32(emu, du[l] * (1.0 - fp) + pa[l] * fp) * md);
            }
        } else {
            const double lir = G->lir[i];
            const double *is = S->ir_sed + s * ni;
            for (int64_t l = 0; l < ni; ++l) { w->ilam[l] = st32(emu, il[l]); w->ised[l] = st32(emu, lir * is[l]); }
        }
    }
    if (no_ir) { /* :2104-2106 */
        n = no; memcpy(w->rlam, w->olam, (size_t)n * sizeof(double)); memcpy(w->rsed, w->osed, (size_t)n * sizeof(double));
    } else if (S->no_stellar) { /* :2107-2109 */
        n = ni; memcpy(w->rlam, w->ilam, (size_t)n * sizeof(double)); memcpy(w->rsed, w->ised, (size_t)n * sizeof(double));
    } else { /* :2110-2119 */
        n = merge_add(no, w->olam, w->osed, ni, w->ilam, w->ised, w->rlam, w->rsed, emu);
    }
    /* IGM (src/egg-gencat.cpp:2122-2140) */
    if (S->apply_igm) {
        int64_t lz = lower_bound_d(n, w->rlam, 0.0600), l0 = lower_bound_d(n, w->rlam, 0.0912);
        int64_t l1 = lower_bound_d(n, w->rlam, 0.1026), l2 = lower_bound_d(n, w->rlam, 0.1216);
        /* npos is size_t(-1) in the reference: range(npos) would run off the array.  The oracle
         * treats "no node <= x" as an empty leading range, which is what a grid starting above x
         * means physically; the product refuses such grids instead. */
        if (lz == ORC_NPOS) lz = 0;
        if (l0 == ORC_NPOS) l0 = 0;
        if (l1 == ORC_NPOS) l1 = 0;
        if (l2 == ORC_NPOS) l2 = 0;
        const float *ia = G->igm_abs + 3 * i;
        for (int64_t l = 0; l < lz; ++l) w->rsed[l] = 0.0;
        for (int64_t l = lz; l < l0; ++l) w->rsed[l] = st32(emu, w->rsed[l] * ia[0]);
        for (int64_t l = l0; l < l1; ++l) w->rsed[l] = st32(emu, w->rsed[l] * ia[1]);
        for (int64_t l = l1; l < l2; ++l) w->rsed[l] = st32(emu, w->rsed[l] * ia[2]);
    }
    /* emission lines (src/egg-gencat.cpp:2143-2147) */
    if (!S->no_nebular && llum) {
        for (int32_t l = 0; l < S->nline; ++l)
            add_emission_line(n, w->rlam, w->rsed, S->line_lambda[l], llum[i * S->nline + l], G->vdisp[i],
                              S->gauss_mode, emu);
    }
    return n;
}

static scratch_t scratch_new(const orc_setup *S) {
    size_t n = (size_t)(S->nopt + S->nir + 8);
    scratch_t w;
    double *b = (double *)malloc(8 * n * sizeof(double));
    w.rlam = b; w.rsed = b + n; w.olam = b + 2 * n; w.osed = b + 3 * n;
    w.ilam = b + 4 * n; w.ised = b + 5 * n; w.lam = b + 6 * n; w.sed = b + 7 * n;
    return w;
}

/* get_flux (src/egg-gencat.cpp:2097-2203) for one component over all galaxies.
 * flux [ngal][nband], rfmag [ngal][nrf] (either may be NULL when its band list is empty).
 * nthreads <= 0: all OpenMP threads. */
void orc_get_flux(const orc_setup *S, const orc_galaxies *G, const float *m, const uint64_t *optsed,
                  const float *llum, float *flux, float *rfmag, int no_ir, int nthreads) {
    const int emu = S->emulate_f32;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        scratch_t w = scratch_new(S);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (int64_t i = 0; i < G->ngal; ++i) {
            int64_t n = build_rest_sed(S, G, i, m, optsed, llum, no_ir, &w);
            /* rest-frame magnitudes (src/egg-gencat.cpp:2150-2158) */
            if (S->nrf > 0 && rfmag) {
                for (int64_t l = 0; l < n; ++l) w.sed[l] = st32(emu, lsun2uJy(0.0, 1e-5, w.rlam[l], w.rsed[l]));
                for (int32_t b = 0; b < S->nrf; ++b) {
                    const int64_t o = S->rfoff[b], nf = S->rfoff[b + 1] - o;
                    double mag = uJy2mag(orc_sed2flux(nf, S->rflam + o, S->rfres + o, n, w.rlam, w.sed, S->nodes_mode));
                    rfmag[i * S->nrf + b] = isfinite(mag) ? (float)mag : 99.0f;
                }
            }
            /* redshift (src/egg-gencat.cpp:2161-2162) */
            const double z = G->z[i], d = G->d[i];
            for (int64_t l = 0; l < n; ++l) {
                w.lam[l] = st32(emu, w.rlam[l] * (1.0 + z));
                w.sed[l] = st32(emu, lsun2uJy(z, d, w.rlam[l], w.rsed[l]));
            }
            /* band fluxes (src/egg-gencat.cpp:2196-2199) */
            if (flux) for (int32_t b = 0; b < S->nband; ++b) {
                const int64_t o = S->foff[b], nf = S->foff[b + 1] - o;
                double f = orc_sed2flux(nf, S->flam + o, S->fres + o, n, w.lam, w.sed, S->nodes_mode);
                flux[i * S->nband + b] = isfinite(f) ? (float)f : 0.0f;
            }
        }
        free(w.rlam);
    }
}

/* As above but returns the fluxes in double (no rounding to the f32 catalog column), so that the
 * 1e-10 fp64-mode tolerance can be tested below f32 resolution. */
void orc_get_flux_f64(const orc_setup *S, const orc_galaxies *G, const float *m, const uint64_t *optsed,
                      const float *llum, double *flux, double *rfmag, int no_ir, int nthreads) {
    const int emu = S->emulate_f32;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        scratch_t w = scratch_new(S);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (int64_t i = 0; i < G->ngal; ++i) {
            int64_t n = build_rest_sed(S, G, i, m, optsed, llum, no_ir, &w);
            if (S->nrf > 0 && rfmag) {
                for (int64_t l = 0; l < n; ++l) w.sed[l] = st32(emu, lsun2uJy(0.0, 1e-5, w.rlam[l], w.rsed[l]));
                for (int32_t b = 0; b < S->nrf; ++b) {
                    const int64_t o = S->rfoff[b], nf = S->rfoff[b + 1] - o;
                    double mag = uJy2mag(orc_sed2flux(nf, S->rflam + o, S->rfres + o, n, w.rlam, w.sed, S->nodes_mode));
                    rfmag[i * S->nrf + b] = isfinite(mag) ? mag : 99.0;
                }
            }
            const double z = G->z[i], d = G->d[i];
            for (int64_t l = 0; l < n; ++l) {
                w.lam[l] = st32(emu, w.rlam[l] * (1.0 + z));
                w.sed[l] = st32(emu, lsun2uJy(z, d, w.rlam[l], w.rsed[l]));
            }
            if (flux) for (int32_t b = 0; b < S->nband; ++b) {
                const int64_t o = S->foff[b], nf = S->foff[b + 1] - o;
                double f = orc_sed2flux(nf, S->flam + o, S->fres + o, n, w.lam, w.sed, S->nodes_mode);
                flux[i * S->nband + b] = isfinite(f) ? f : 0.0;
            }
        }
        free(w.rlam);
    }
}

/* Observed-frame SED of one galaxy component as save_sed writes it (src/egg-gencat.cpp:2161-2179):
 * lam then sed, both f32.  Returns the node count (buffers must hold nopt+nir floats). */
int64_t orc_get_sed(const orc_setup *S, const orc_galaxies *G, int64_t i, const float *m,
                    const uint64_t *optsed, const float *llum, int no_ir, float *lam_out, float *sed_out) {
    scratch_t w = scratch_new(S);
    int64_t n = build_rest_sed(S, G, i, m, optsed, llum, no_ir, &w);
    const double z = G->z[i], d = G->d[i];
    for (int64_t l = 0; l < n; ++l) {
        lam_out[l] = (float)(w.rlam[l] * (1.0 + z));
        sed_out[l] = (float)lsun2uJy(z, d, w.rlam[l], w.rsed[l]);
    }
    free(w.rlam);
    return n;
}

/* totals (src/egg-gencat.cpp:2238-2239): vec2f arithmetic for the fluxes, magnitudes through uJy */
void orc_totals(int64_t nflux, const float *fd, const float *fb, float *f, int64_t nmag,
                const float *md, const float *mb, float *mt) {
    for (int64_t k = 0; k < nflux; ++k) f[k] = fd[k] + fb[k];
    for (int64_t k = 0; k < nmag; ++k) mt[k] = (float)uJy2mag(mag2uJy(md[k]) + mag2uJy(mb[k]));
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
