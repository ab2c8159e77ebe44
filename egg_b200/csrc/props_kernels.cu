// Per-galaxy property recipes of egg-gencat on the GPU: C ABI of include/eggprops.h.
//
// One thread per galaxy, all arithmetic in double (the recipes are O(ngal) and cost nothing next to the
// photometry), results rounded once to the reference's column types.  Every formula cites the reference
// line it restates (src/egg-gencat.cpp); constants are the reference's calibration values.  Random scatter
// comes from Philox4x32-10 keyed by the seed with counter (galaxy index, draw slot): see eggprops.h.
// The tests check this kernel against a numpy restatement of the same lines (same generator, same draw slots).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/eggprops.h"

namespace {

// ---- Philox4x32-10 (Salmon et al. 2011) ------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };
__host__ __device__ inline U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
        c = U4{(uint32_t)(p1 >> 32) ^ c.y ^ k0, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k1, (uint32_t)p0};
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// draw slots (one Philox block each; the order is part of the definition of a realisation)
enum Slot {
    S_BT, S_ANGLE, S_BULGE_RATIO, S_BULGE_RADIUS, S_DISK_RATIO, S_DISK_RADIUS, S_SFR, S_BURST, S_IRX,
    S_DISK_UVJ_A, S_DISK_UVJ_VJ, S_DISK_UVJ_UV, S_RED_BULGE, S_BULGE_UVJ_A, S_BULGE_UVJ_VJ, S_BULGE_UVJ_UV,
    S_TDUST, S_IR8, S_IGM0, S_IGM1, S_IGM2, S_GDR, S_MH2, S_AVL_DISK, S_AVL_BULGE, S_FESC, S_VDISP,
    S_C2, S_N2_205, S_HALPHA, S_O3MN2, S_O3PN2, S_O2
};

struct Draws {
    uint64_t index;
    uint32_t k0, k1;
    __device__ double uniform(int slot) const {
        const U4 r = philox4x32_10(U4{(uint32_t)index, (uint32_t)(index >> 32), (uint32_t)slot, 0u}, k0, k1);
        return ((double)r.x + 0.5) * 2.3283064365386963e-10; // 2^-32
    }
    __device__ double normal(int slot) const { // Box-Muller on words 0 and 1
        const U4 r = philox4x32_10(U4{(uint32_t)index, (uint32_t)(index >> 32), (uint32_t)slot, 0u}, k0, k1);
        const double u1 = ((double)r.x + 0.5) * 2.3283064365386963e-10, u2 = ((double)r.y + 0.5) * 2.3283064365386963e-10;
        return sqrt(-2.0 * log(u1)) * cos(2.0 * 3.141592653589793 * u2);
    }
};

// ---- [vif] helpers ------------------------------------------------------------------------------------
// interpolate(y, x, v): linear, end segments extended
template <int N> __device__ inline double interp(const double (&y)[N], const double (&x)[N], double v) {
    int i = 0;
    while (i < N - 2 && v >= x[i + 1]) ++i;
    return y[i] + (y[i + 1] - y[i]) * (v - x[i]) / (x[i + 1] - x[i]);
}
__device__ inline double interp_n(const double *y, const double *x, int n, double v) {
    int i = 0;
    while (i < n - 2 && v >= x[i + 1]) ++i;
    return y[i] + (y[i + 1] - y[i]) * (v - x[i]) / (x[i + 1] - x[i]);
}
__device__ inline double e10(double x) { return pow(10.0, x); }
__device__ inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); } // NaN passes through
__device__ inline double poly3(double x, double c0, double c1, double c2, double c3) { return c0 + x * c1 + x * x * c2 + x * x * x * c3; }

constexpr int kNQ = 13;
// Inoue+14 quantile fits (src/egg-gencat.cpp:1776-1808), stored as the reference stores them (float)
__constant__ float c_igm_z1[3][kNQ] = {
    {0.00000f, -0.119247f, -0.836013f, 0.188466f, 5.58061f, 1.80948f, 2.23333f, 2.58904f, 3.05070f, 3.44932f, 3.64749f, 4.33884f, 4.80831f},
    {4.71243f, 4.51667f, 5.13483f, 5.51325f, 6.02321f, 6.11072f, 3.96055f, 6.16846f, 6.20577f, 4.11253f, 6.25897f, 6.30988f, 6.30475f},
    {4.29922f, 4.43374f, 5.81153f, 6.07670f, 6.19915f, 4.01990f, 6.22758f, 6.23932f, 4.18000f, 6.29191f, 6.30732f, 6.34660f, 4.29822f}};
__constant__ float c_igm_dz1[3][kNQ] = {
    {0.00000f, 0.103763f, 1.53267f, 0.831935f, 0.101257f, 1.30343f, 1.23153f, 1.18386f, 1.18305f, 1.21584f, 1.22993f, 1.25208f, 1.15790f},
    {1.65837f, 1.78514f, 1.81298f, 1.76652f, 1.33753f, 1.19458f, 2.24961f, 1.11023f, 1.07492f, 2.16499f, 1.03259f, 0.998597f, 1.00750f},
    {1.83693f, 1.92457f, 1.44038f, 1.07901f, 0.980853f, 2.13450f, 1.00173f, 0.996720f, 2.25818f, 0.979706f, 0.979425f, 0.962433f, 2.04183f}};
__constant__ float c_igm_z2[3][kNQ] = {
    {0.00000f, 0.00000f, 8.03488f, 4.25279f, 0.989755f, 6.97760f, 0.00000f, 0.00000f, 7.46371f, 5.79003f, 5.31787f, 4.35838f, 4.12159f},
    {3.13640f, 4.21302f, 4.01985f, 3.91207f, 3.85061f, 3.91916f, 6.14354f, 3.99942f, 4.05655f, 6.24326f, 4.14120f, 4.22585f, 4.28847f},
    {5.85101f, 5.60883f, 3.88240f, 3.87634f, 3.98672f, -9.52333f, 4.10508f, 4.13700f, 6.26376f, 4.22116f, 4.24243f, 4.30912f, 9.86437f}};
__constant__ float c_igm_dz2[3][kNQ] = {
    {0.00000f, 0.00000f, 0.209441f, 0.0120493f, 1.27333f, 0.0810017f, 0.00000f, 0.00000f, 5.04822f, 3.71614f, 3.30993f, 2.40054f, 2.10543f},
    {5.31391f, 5.23251f, 3.21628f, 2.77969f, 2.37366f, 2.28242f, 1.14531f, 2.22511f, 2.19319f, 1.04437f, 2.15259f, 2.11453f, 2.11378f},
    {9.48358f, 6.52196f, 2.79285f, 2.47621f, 2.31278f, -1.12091f, 2.28194f, 2.27401f, 0.988508f, 2.24063f, 2.23221f, 2.20610f, 0.233647f}};
__constant__ float c_igm_q[kNQ] = {0.0f, 0.001f, 0.023f, 0.05f, 0.16f, 0.35f, 0.50f, 0.65f, 0.84f, 0.95f, 0.977f, 0.999f, 1.0f};

const float h_line_lambda[EGGPROPS_NLINE] = {157.71f, 205.18f, 609.14f, 2600.9f, 1300.4f, 867.0f, 650.27f, 520.24f, 433.57f, 371.66f,
                                             0.65628f, 0.48613f, 0.43405f, 0.41017f, 0.65835f, 0.65480f, 0.50068f, 0.49589f, 0.37274f, 0.12157f};
const char *const h_line_name[EGGPROPS_NLINE] = {"c2_157", "n2_205", "c1_609", "co10", "co21", "co32", "co43", "co54", "co65", "co76",
                                                 "halpha", "hbeta", "hgamma", "hdelta", "n2_6583", "n2_6548", "o3_5007", "o3_4959", "o2_3727", "lyalpha"};
enum { L_C2 = 0, L_N2_205, L_C1, L_CO10, L_HA = 10, L_HB, L_HG, L_HD, L_N2_6583, L_N2_6548, L_O3_5007, L_O3_4959, L_O2, L_LYA };

struct PropsParams {
    const float *z, *m;
    const uint8_t *passive;
    uint64_t ngal, first_index;
    uint32_t k0, k1;
    int32_t no_random, no_stellar, no_nebular, no_passive_lir, magdis_tdust, igm, ir_lib;
    double ms_disp, h0, omega_m, omega_l;
    // tables (device)
    const float *buv_lo, *bvj_lo; // lower edges
    const int32_t *sed_map;       // [nuv][nvj] flat index of the closest usable SED
    const float *av;
    uint32_t nuv, nvj, nir_sed;
    const double *cs_tdust, *cs_lir_dust, *cs_lir_pah, *cs_l8_dust, *cs_l8_pah;
    double ratio_x[12], bulge_cdf[12], disk_cdf[12]; // random_pdf tables (gencat:1417-1436)
    double gl_x[32], gl_w[32];                       // Gauss-Legendre nodes for the comoving distance
    double dust_law[EGGPROPS_NLINE];                 // Calzetti k(lambda)/Rv + 1 per line, 0 for Ly-alpha (gencat:1986-2005)
    int32_t igm_band[EGGPROPS_NLINE];                // -2: no transmission, -1: unattenuated, 0..2: igm_abs column (gencat:2008-2020)
    eggprops_out out;
};

// value of x distributed as the tabulated density, from a uniform deviate ([vif] random_pdf)
__device__ double pdf_draw(const double *x, const double *cdf, double u) {
    int i = 0;
    while (i < 10 && u >= cdf[i + 1]) ++i;
    const double w = cdf[i + 1] > cdf[i] ? (u - cdf[i]) / (cdf[i + 1] - cdf[i]) : 0.0;
    return x[i] + w * (x[i + 1] - x[i]);
}
__device__ int bin_open(const float *lo, int n, float x) { // [vif] in_bin_open: first and last bins open-ended
    int a = 0, b = n; // last i with lo[i] <= x, clamped
    while (b - a > 1) {
        const int mid = (a + b) >> 1;
        if (lo[mid] <= x) a = mid; else b = mid;
    }
    return a;
}
// gen_blue_uvj / gen_red_uvj (src/egg-gencat.cpp:619-668)
__device__ void blue_uvj(double m, double z, const Draws &d, int s0, bool rnd, double &uv, double &vj) {
    const double a0 = 0.58 * erf(m - 10) + 1.39, as = -0.34 + 0.3 * fmax(m - 10.35, 0.0);
    double a = fmin(a0 + as * fmin(z, 3.3), 2.0);
    if (rnd) {
        const double amp = 0.2 + (0.25 + 0.12 * clampd((z - 0.5) / 2.0, 0, 1)) * fmax(1.0 - 2.0 * fabs(m - (10.3 + 0.4 * erf(z - 1.5))), 0.0);
        a += amp * d.normal(s0);
    }
    const double theta = atan(0.65);
    vj = a * cos(theta);
    uv = 0.45 + a * sin(theta);
    if (rnd) { vj += 0.15 * d.normal(s0 + 1); uv += 0.15 * d.normal(s0 + 2); }
    const double off = fmax((0.5 - z) / 0.5, 0.0);
    uv += 0.4 * off;
    vj += 0.2 * off;
}
__device__ void red_uvj(double m, double z, const Draws &d, int s0, bool rnd, double &uv, double &vj) {
    double ms = 0.1 * (m - 11.0);
    if (rnd) ms += 0.1 * d.normal(s0);
    ms = clampd(ms, -0.1, 0.2);
    vj = 1.25 + ms;
    uv = 1.85 + ms * 0.88;
    if (rnd) { vj += 0.1 * d.normal(s0 + 1); uv += 0.1 * d.normal(s0 + 2); }
    const double off = fmax((0.5 - z) / 0.5, 0.0);
    uv += 0.4 * off;
    vj += 0.2 * off;
}

#define PUT(col, v) do { if (p.out.col) p.out.col[g] = (float)(v); } while (0)

__global__ void __launch_bounds__(128) props_kernel(const __grid_constant__ PropsParams p) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.ngal) return;
    const double z = (double)p.z[g], m = (double)p.m[g];
    const bool passive = p.passive[g] != 0, rnd = !p.no_random;
    const Draws d{p.first_index + g, p.k0, p.k1};

    // bulge-to-total ratio (Lang+14), src/egg-gencat.cpp:1367-1387
    double bt = passive ? e10(0.1 * (m - 10.0) - 0.3) : e10(0.27 * (m - 10.0) - 0.7);
    if (rnd) bt *= e10(0.2 * d.normal(S_BT));
    bt = clampd(bt, 0.0, 1.0);
    const double mb = m + log10(bt), md = m + log10(1.0 - bt);
    PUT(bt, bt);
    PUT(m_bulge, isfinite(mb) ? mb : 0.0);
    PUT(m_disk, isfinite(md) ? md : 0.0);

    // morphology, :1398-1449
    const double az = log10(1.0 + z);
    double size_disk = z < 1.7 ? e10(0.41 - 0.22 * az + 0.2 * (fmin(m, 10.6) - 9.35)) : e10(0.62 - 0.7 * az + 0.2 * (fmin(m, 10.6) - 9.35));
    double size_bulge = z < 0.5 ? e10(0.78 - 0.6 * az + 0.56 * (m - 11.25)) : e10(0.90 - 1.3 * az + 0.56 * (m - 11.25));
    const double angle = (d.uniform(S_ANGLE) - 0.5) * 180.0;
    PUT(disk_angle, angle);
    PUT(bulge_angle, angle);
    PUT(bulge_ratio, pdf_draw(p.ratio_x, p.bulge_cdf, d.uniform(S_BULGE_RATIO)));
    PUT(disk_ratio, pdf_draw(p.ratio_x, p.disk_cdf, d.uniform(S_DISK_RATIO)));
    if (rnd) {
        size_bulge *= e10(0.2 * d.normal(S_BULGE_RADIUS));
        size_disk *= e10(0.17 * d.normal(S_DISK_RADIUS));
    }
    double dc = 0.0; // line-of-sight comoving distance [Mpc]: 32-point Gauss-Legendre of 1/E on [0, z]
    {
        const double ok = 1.0 - p.omega_m - p.omega_l;
        for (int k = 0; k < 32; ++k) {
            const double zz = 0.5 * z * (p.gl_x[k] + 1.0), a = 1.0 + zz;
            dc += p.gl_w[k] / sqrt(p.omega_m * a * a * a + ok * a * a + p.omega_l);
        }
        dc *= 2.99792458e5 / p.h0 * 0.5 * z;
    }
    const double psize = dc / (1.0 + z) * 1e3 * (3.141592653589793 / 180.0 / 3600.0); // [vif] propsize: proper kpc per arcsec
    PUT(d, dc * (1.0 + z));
    PUT(disk_radius, size_disk / psize);
    PUT(bulge_radius, size_bulge / psize);

    // SFR, :1459-1484
    const double t0 = fmax(0.0, m - 9.59 - 2.22 * az);
    const double sfrms = m - 9.67 + 1.82 * az - 0.38 * t0 * t0;
    double lsfr = passive ? fmin(sfrms, 0.5 * (m - 11) + az - 0.6) : sfrms;
    if (rnd) {
        lsfr += (passive ? 0.4 : p.ms_disp) * d.normal(S_SFR);
        if (!passive && d.uniform(S_BURST) < 0.033) lsfr += 0.72; // starbursts
    }
    const double rsb = lsfr - sfrms, sfr = e10(lsfr);

    // IRX, :1488-1506
    double irx = e10((0.45 * fmin(z, 3.0) + 0.35) * (m - 10.5) + 1.2);
    if (rnd) irx *= e10(0.4 * d.normal(S_IRX));
    if (p.no_passive_lir && passive) irx = 0.0;
    const double sfrir = sfr / (1.0 + 1.0 / irx), sfruv = sfr / (1.0 + irx);
    PUT(sfr, sfr); PUT(rsb, rsb); PUT(irx, irx); PUT(sfrir, sfrir); PUT(sfruv, sfruv);

    // UVJ colours, :1515-1546: disks blue if active, red if passive; bulges red if dominant or passive, else a coin
    double uvd, vjd, uvb, vjb;
    if (passive) red_uvj(m, z, d, S_DISK_UVJ_A, rnd, uvd, vjd); else blue_uvj(m, z, d, S_DISK_UVJ_A, rnd, uvd, vjd);
    bool red_bulge = bt > 0.6 || passive;
    if (rnd) red_bulge = red_bulge || d.uniform(S_RED_BULGE) < 0.5;
    if (red_bulge) red_uvj(m, z, d, S_BULGE_UVJ_A, rnd, uvb, vjb); else blue_uvj(m, z, d, S_BULGE_UVJ_A, rnd, uvb, vjb);
    PUT(rfuv_disk, uvd); PUT(rfvj_disk, vjd); PUT(rfuv_bulge, uvb); PUT(rfvj_bulge, vjb);
    {
        const double y[5] = {0.15, 0.15, 0.0, 0.0, -0.6}, x[5] = {0.0, 0.45, 1.3, 6.0, 8.0}; // get_m2l_cor, :544-546
        PUT(m2lcor, interp(y, x, z));
    }

    // optical SEDs from the UVJ bins, :1554-1571 and :670-705 (the colours are f32 columns when they are binned)
    double av_disk = nan(""), av_bulge = nan("");
    if (!p.no_stellar) {
        const int sd = p.sed_map[bin_open(p.buv_lo, p.nuv, (float)uvd) * p.nvj + bin_open(p.bvj_lo, p.nvj, (float)vjd)];
        const int sb = p.sed_map[bin_open(p.buv_lo, p.nuv, (float)uvb) * p.nvj + bin_open(p.bvj_lo, p.nvj, (float)vjb)];
        av_disk = (double)p.av[sd];
        av_bulge = (double)p.av[sb];
        if (p.out.opt_sed_disk) p.out.opt_sed_disk[g] = (uint64_t)sd;
        if (p.out.opt_sed_bulge) p.out.opt_sed_bulge[g] = (uint64_t)sb;
    } else {
        if (p.out.opt_sed_disk) p.out.opt_sed_disk[g] = ~(uint64_t)0; // npos
        if (p.out.opt_sed_bulge) p.out.opt_sed_bulge[g] = ~(uint64_t)0;
    }
    PUT(av_disk, av_disk); PUT(av_bulge, av_bulge);

    // IR properties, :1584-1668
    const double tdust_ms = 32.9 + 4.60 * (z - 2.0) - 0.77;
    double tdust = tdust_ms;
    if (p.magdis_tdust && z > 1) tdust = 2.0 * (clampd(z, 0.0, 2.0) - 1.0) + 27.0;
    tdust += 10.1 * rsb;                                               // starbursts are warmer
    tdust -= 1.5 * fmax(0.0, 2.0 - z) * clampd(m - 10.7, 0.0, 1.0);    // massive galaxies are colder
    double ir8 = (4.08 + 3.29 * clampd(z - 1.0, 0.0, 1.0)) * 0.81;
    ir8 *= e10(0.66 * fmax(0.0, rsb));
    ir8 *= e10(-1.0 * clampd(m - 10.0, -1, 0.0));
    if (rnd) {
        tdust += 0.12 * tdust_ms * d.normal(S_TDUST);
        ir8 *= e10(0.18 * d.normal(S_IR8));
    }
    double lir = sfrir / 1.72e-10;
    ir8 = clampd(ir8, 0.48, 27.5);
    double fpah = clampd(1.0 / (1.0 - (331 - 691 * ir8) / (193 - 6.98 * ir8)), 0.0, 1.0);
    if (!isfinite(lir)) lir = 0.0;

    // IR SED, :1680-1725
    double fir = 0.0;
    if (p.ir_lib == EGGPROPS_IR_CE01) {
        const double y[7] = {26, 26, 40, 54, 53, 52, 52}, x[7] = {0.57, 1.0, 1.5, 2.1, 2.9, 4.0, 6.0};
        fir = interp(y, x, z) + 15 * clampd(rsb / p.ms_disp, -2.0, 2.0);
    } else if (p.ir_lib == EGGPROPS_IR_M12) {
        const double x[8] = {0.0125, 0.1625, 0.4625, 0.8125, 1.15, 1.525, 2.0, 2.635}, y[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        fir = interp_n(y, x, min((int)p.nir_sed, 8), z) + clampd(rsb / p.ms_disp, -2.0, 2.0);
    } else if (p.ir_lib == EGGPROPS_IR_CS17) {
        int i = 0; // interpolate(indgen(nirsed), tdust_lib, tdust)
        while (i < (int)p.nir_sed - 2 && tdust >= p.cs_tdust[i + 1]) ++i;
        fir = i + (tdust - p.cs_tdust[i]) / (p.cs_tdust[i + 1] - p.cs_tdust[i]);
    }
    const int ir_sed = (int)clampd(floor(fir + 0.5), 0.0, (double)p.nir_sed - 1.0); // [vif] round, then clamp
    double mdust;
    if (p.ir_lib == EGGPROPS_IR_CS17) {
        const double ld = p.cs_lir_dust[ir_sed], lp = p.cs_lir_pah[ir_sed], l8d = p.cs_l8_dust[ir_sed], l8p = p.cs_l8_pah[ir_sed];
        fpah = clampd(1.0 / (1.0 - (lp - l8p * ir8) / (ld - l8d * ir8)), 0.0, 1.0);
        mdust = lir / (ld * (1.0 - fpah) + lp * fpah);
    } else {
        fpah = 0.04; // placeholders, :1722-1724
        mdust = lir / 1e3;
    }
    PUT(tdust, tdust); PUT(ir8, ir8); PUT(lir, lir); PUT(fpah, fpah); PUT(mdust, mdust);
    if (p.out.ir_sed) p.out.ir_sed[g] = (uint64_t)ir_sed;

    // IGM transmission in the three bands below Ly-alpha, :1731-1824
    double igm_abs[3] = {0.0, 0.0, 0.0};
    if (p.igm == EGGPROPS_IGM_INOUE14) {
        // transmission at quantile q of band l: product of two error-function steps in redshift (0 where the fit
        // has no width); 0.5 erfc(-x) = 0.5 (erf(x) + 1) without the cancellation at small transmissions
        auto quant = [&](int l, int q) {
            if (c_igm_dz1[l][q] == 0.0f) return 0.0;
            double t = 0.5 * erfc(-((double)c_igm_z1[l][q] - z) / (double)c_igm_dz1[l][q]);
            if (c_igm_dz2[l][q] != 0.0f) t *= 0.5 * erfc(-((double)c_igm_z2[l][q] - z) / (double)c_igm_dz2[l][q]);
            return t;
        };
        for (int l = 0; l < 3; ++l) {
            if (!rnd) igm_abs[l] = quant(l, kNQ / 2); // median transmission
            else { // the two quantiles that bracket the uniform deviate
                const double u = d.uniform(S_IGM0 + l);
                int i = 0;
                while (i < kNQ - 2 && u >= (double)c_igm_q[i + 1]) ++i;
                const double t0 = quant(l, i), t1 = quant(l, i + 1);
                igm_abs[l] = t0 + (t1 - t0) * (u - (double)c_igm_q[i]) / ((double)c_igm_q[i + 1] - (double)c_igm_q[i]);
            }
        }
    } else if (p.igm == EGGPROPS_IGM_MADAU95) { // FAST's implementation: mean of exp(-tau) over 100 wavelengths per band
        double s2 = 0.0, s1 = 0.0;
        for (int k = 0; k < 100; ++k) {
            const double f = k / 99.0;
            double tl = (1050.0 + (1170.0 - 1050.0) * f) * (1.0 + z);
            s2 += exp(-3.6e-3 * pow(tl / 1216.0, 3.46));
            tl = (920.0 + (1015.0 - 920.0) * f) * (1.0 + z);
            s1 += exp(-1.7e-3 * pow(tl / 1026.0, 3.46) - 1.2e-3 * pow(tl / 972.5, 3.46) - 9.3e-4 * pow(tl / 950.0, 3.46));
        }
        igm_abs[2] = s2 / 100.0;
        igm_abs[1] = s1 / 100.0;
    }
    if (p.out.igm_abs)
        for (int l = 0; l < 3; ++l) p.out.igm_abs[g * 3 + l] = (float)igm_abs[l];

    // emission lines, :1829-2022
    if (!p.no_nebular) {
        double L[EGGPROPS_NLINE];
        const double lsfr10 = log10(sfr);
        const double mu32 = (m - 0.2) - 0.32 * (lsfr10 - 0.2);
        double metal = (mu32 < 10.36 ? 8.9 + 0.47 * (mu32 - 10) : 9.07) - 8.69; // in solar metallicity
        if (z > 1) metal -= 0.2 * clampd(z - 1, 0.0, 1.0);                      // broken FMR
        metal = clampd(metal, -0.6, 1.0);
        double gdr = 2.23 - metal;                                              // gas-to-dust ratio
        if (rnd) gdr += 0.04 * d.normal(S_GDR);
        double mh2 = mdust * e10(gdr) * 0.3;
        if (rnd) mh2 *= e10(0.2 * d.normal(S_MH2));
        double avd = 0.5 * (log10(sfr / sfruv) * 0.95 + av_disk);
        {
            const double y[4] = {1.7, 1.3, 1.0, 1.0}, x[4] = {0, 1, 2, 100};
            avd *= interp(y, x, z);
        }
        double avb = av_bulge;
        if (rnd) { avd += 0.1 * d.normal(S_AVL_DISK); avb += 0.1 * d.normal(S_AVL_BULGE); }
        avd = clampd(avd, 0.0, 6.0);
        avb = clampd(avb, 0.0, 6.0);
        double fesc = 0.445 * e10(-0.4 * (av_disk / 4.05) * 17.8); // Ly-alpha escape fraction (Hayes+10)
        {
            const double y[5] = {-1.2, -1.2, -0.8, 0, 0}, x[5] = {0.5, 1.2, 2.2, 3.0, 4.0};
            fesc *= e10(interp(y, x, z));
        }
        if (rnd) fesc *= e10(0.4 * d.normal(S_FESC));
        fesc = clampd(fesc, 0.0, 1.0);
        double vdisp = e10(0.12 * (m - 10) + 1.78); // Stott+16
        if (rnd) vdisp *= e10(0.3 * d.normal(S_VDISP));
        L[L_C2] = e10(1.017 * (lsfr10 - 0.2) + 7.13) * (rnd ? e10(0.27 * d.normal(S_C2)) : 1.0);       // deLooze+11
        L[L_N2_205] = e10(1.053 * (lsfr10 - 0.2) + 5.31) * (rnd ? e10(0.26 * d.normal(S_N2_205)) : 1.0); // Zhao+13
        L[L_C1] = 0.00246 * mh2;   // Bothwell+17
        L[L_CO10] = 4.93e-5 * mh2; // alphaCO = 4.6
        {
            const double sled[6] = {2.3, 3.7, 4.5, 7.5, 8.5, 8.3}; // Bothwell+13 SMGs
            for (int k = 0; k < 6; ++k) L[L_CO10 + 1 + k] = L[L_CO10] * sled[k];
        }
        const double lha = e10(lsfr10 + 7.519) * (rnd ? e10(0.15 * d.normal(S_HALPHA)) : 1.0); // Kennicutt+98
        L[L_HA] = lha;
        L[L_HB] = lha * 0.35; L[L_HG] = lha * 0.163; L[L_HD] = lha * 0.0862; // case B
        double lo3mn2 = poly3(metal, 0.59518604, -2.7283054, -0.23229248, 2.0151786);
        if (rnd) lo3mn2 += poly3(metal, 0.32970659, 0.043645415, -1.2231136, -2.4836792) * d.normal(S_O3MN2);
        double lo3pn2 = poly3(lo3mn2, -0.69337981, 0.66460936, -0.37040100, 0.047935498);
        if (rnd) lo3pn2 += poly3(lo3mn2, 0.086819227, -0.078283561, 0.093276637, -0.025069864) * d.normal(S_O3PN2);
        double n2off;
        {
            const double y1[4] = {0.0, 0.0, 0.3, 0.3}, x1[4] = {0.0, 1.0, 2.3, 2.4}, y2[4] = {1.0, 1.0, 0.0, 0.0}, x2[4] = {9, 10.0, 10.5, 11.0};
            n2off = interp(y1, x1, z) * interp(y2, x2, m); // Shapley+15
        }
        const double ln2 = lha * e10(0.5 * (lo3pn2 - lo3mn2) + n2off);
        L[L_N2_6583] = 0.75 * ln2; L[L_N2_6548] = 0.25 * ln2;
        const double lhb = L[L_HB], lo3 = lhb * e10(0.5 * (lo3pn2 + lo3mn2));
        L[L_O3_5007] = 0.75 * lo3; L[L_O3_4959] = 0.25 * lo3;
        const double o3hb = log10(lo3 / lhb);
        double lo2 = lhb * e10(poly3(o3hb, 0.51869283, 0.29843257, -0.49610294, -0.11728400));
        if (rnd) lo2 *= e10(poly3(o3hb, 0.061403650, -0.0081335335, 0.078420795, -0.032656826) * d.normal(S_O2));
        L[L_O2] = lo2;
        L[L_LYA] = 8.7 * lha * fesc; // Sobral+18
        PUT(avlines_disk, avd); PUT(avlines_bulge, avb); PUT(vdisp, vdisp);
        for (int l = 0; l < EGGPROPS_NLINE; ++l) { // reddening and IGM, :2004-2021; bulges carry no lines
            const double ig = p.igm_band[l] == -2 ? 0.0 : (p.igm_band[l] < 0 ? 1.0 : igm_abs[p.igm_band[l]]);
            if (p.out.line_lum_disk) p.out.line_lum_disk[g * EGGPROPS_NLINE + l] = (float)(L[l] * e10(-0.4 * avd * p.dust_law[l]) * ig);
            if (p.out.line_lum_bulge) p.out.line_lum_bulge[g * EGGPROPS_NLINE + l] = 0.0f;
        }
    }
}

// Gauss-Legendre nodes on [-1, 1] by Newton iteration on P_n (host, at create)
void gauss_legendre(int n, double *x, double *w) {
    for (int i = 0; i < n; ++i) {
        double t = std::cos(3.141592653589793 * (i + 0.75) / (n + 0.5)), dp = 1.0;
        for (int it = 0; it < 100; ++it) {
            double p0 = 1.0, p1 = t;
            for (int k = 2; k <= n; ++k) { const double p2 = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
            dp = n * (t * p1 - p0) / (t * t - 1.0);
            const double dt = p1 / dp;
            t -= dt;
            if (std::fabs(dt) < 1e-16) break;
        }
        x[n - 1 - i] = t;
        w[n - 1 - i] = 2.0 / ((1.0 - t * t) * dp * dp);
    }
}

} // namespace

struct eggprops_ctx {
    std::string err;
    int device = 0;
    bool ready = false;
    PropsParams base{};
    std::vector<void *> dev; // device allocations
    cudaStream_t stream = nullptr;
    // staging for the host entry point
    void *stage = nullptr;
    size_t stage_bytes = 0;
    int fail(int code, const std::string &m) { err = m; return code; }
};

namespace {
template <typename T> cudaError_t upload(eggprops_ctx *c, const T *h, size_t n, const T **d) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    c->dev.push_back(p);
    *d = (const T *)p;
    return n ? cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
}
const double kRatioX[12] = {0.0, 0.1, 0.15, 0.25, 0.35, 0.45, 0.55, 0.65, 0.75, 0.85, 0.95, 1.0};
const double kBulgeP[12] = {0.0, 0.0, 165.0, 428.0, 773.0, 914.0, 1069.0, 1191.0, 1154.0, 1067.0, 639.0, 450.0}; // :1414-1417
const double kDiskP[12] = {0.0, 0.0, 313.0, 900.0, 1143.0, 1127.0, 1059.0, 898.0, 775.0, 548.0, 318.0, 200.0};   // :1427-1430
void trapezoid_cdf(const double *x, const double *p, double *cdf) {
    cdf[0] = 0.0;
    for (int i = 1; i < 12; ++i) cdf[i] = cdf[i - 1] + 0.5 * (p[i] + p[i - 1]) * (x[i] - x[i - 1]);
    for (int i = 0; i < 12; ++i) cdf[i] /= cdf[11];
}
double calzetti2000(double l) { // :1986-2002
    const double irv = 1.0 / 3.33;
    double k = l <= 0.63 ? (2.659 * irv) * (-2.156 + 1.509 / l - 0.198 * std::pow(l, -2.0) + 0.011 * std::pow(l, -3.0)) + 1.0
                         : (2.659 * irv) * (-1.857 + 1.040 / l) + 1.0;
    return k < 0 ? 0 : k;
}
} // namespace

extern "C" {

const float *eggprops_line_lambda(void) { return h_line_lambda; }
const char *eggprops_line_name(uint32_t l) { return l < EGGPROPS_NLINE ? h_line_name[l] : nullptr; }
const char *eggprops_last_error(const eggprops_ctx *c) { return c ? c->err.c_str() : "null context"; }

int eggprops_create(const eggprops_libs *libs, const eggprops_opts *opts, eggprops_ctx **out) {
    if (!out) return EGGPHOT_ERR_ARG;
    eggprops_ctx *c = new eggprops_ctx;
    *out = c;
    if (!libs || !opts) return c->fail(EGGPHOT_ERR_ARG, "null argument");
    if (opts->igm < EGGPROPS_IGM_NONE || opts->igm > EGGPROPS_IGM_INOUE14) return c->fail(EGGPHOT_ERR_ARG, "unknown IGM prescription");
    if (libs->ir_lib < EGGPROPS_IR_SINGLE || libs->ir_lib > EGGPROPS_IR_CS17)
        return c->fail(EGGPHOT_ERR_UNSUPPORTED, "no calibration code available for the IR library"); // :1702
    if (libs->nir_sed == 0) return c->fail(EGGPHOT_ERR_ARG, "empty IR library");
    if (libs->ir_lib == EGGPROPS_IR_CS17 && (!libs->cs17_tdust || !libs->cs17_lir_dust || !libs->cs17_lir_pah || !libs->cs17_l8_dust || !libs->cs17_l8_pah || libs->nir_sed < 2))
        return c->fail(EGGPHOT_ERR_ARG, "the cs17 calibration needs TDUST, LIR_DUST, LIR_PAH, L8_DUST, L8_PAH");
    if (!opts->no_stellar && (!libs->buv || !libs->bvj || !libs->use || !libs->av || libs->nuv == 0 || libs->nvj == 0))
        return c->fail(EGGPHOT_ERR_ARG, "optical library bins missing");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return c->fail(EGGPHOT_ERR_CUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    c->device = opts->device;
    if ((e = cudaSetDevice(c->device)) != cudaSuccess) return c->fail(EGGPHOT_ERR_CUDA, cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, c->device);
    if (prop.major != 10) return c->fail(EGGPHOT_ERR_CUDA, "built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));

    PropsParams &p = c->base;
    p.k0 = (uint32_t)opts->seed; p.k1 = (uint32_t)(opts->seed >> 32);
    p.no_random = opts->no_random; p.no_stellar = opts->no_stellar; p.no_nebular = opts->no_nebular;
    p.no_passive_lir = opts->no_passive_lir; p.magdis_tdust = opts->magdis_tdust; p.igm = opts->igm; p.ir_lib = libs->ir_lib;
    p.ms_disp = opts->ms_disp > 0 ? opts->ms_disp : 0.3;
    p.h0 = opts->h0 > 0 ? opts->h0 : 70.0; p.omega_m = opts->h0 > 0 ? opts->omega_m : 0.3; p.omega_l = opts->h0 > 0 ? opts->omega_l : 0.7;
    p.nuv = libs->nuv; p.nvj = libs->nvj; p.nir_sed = libs->nir_sed;
#define CK(x) do { if ((e = (x)) != cudaSuccess) return c->fail(EGGPHOT_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e)); } while (0)
    if (!opts->no_stellar) {
        // closest usable SED of every UVJ bin ([vif] astar_find, :688): smallest squared index distance, ties to the
        // smaller u, then the smaller v
        std::vector<int32_t> map((size_t)libs->nuv * libs->nvj, -1);
        bool any = false;
        for (uint32_t u = 0; u < libs->nuv; ++u)
            for (uint32_t v = 0; v < libs->nvj; ++v) {
                long best = -1; int bi = -1;
                for (uint32_t a = 0; a < libs->nuv; ++a)
                    for (uint32_t b = 0; b < libs->nvj; ++b) {
                        if (!libs->use[a * libs->nvj + b]) continue;
                        const long du = (long)a - (long)u, dv = (long)b - (long)v, dd = du * du + dv * dv;
                        if (best < 0 || dd < best) { best = dd; bi = (int)(a * libs->nvj + b); }
                    }
                if (bi >= 0) any = true;
                map[u * libs->nvj + v] = bi;
            }
        if (!any) return c->fail(EGGPHOT_ERR_LIBRARY, "could not find good SEDs in the optical library!"); // :690
        CK(upload(c, map.data(), map.size(), &p.sed_map));
        CK(upload(c, libs->buv, (size_t)libs->nuv, &p.buv_lo)); // lower edges are the first row
        CK(upload(c, libs->bvj, (size_t)libs->nvj, &p.bvj_lo));
        CK(upload(c, libs->av, (size_t)libs->nuv * libs->nvj, &p.av));
    }
    if (libs->ir_lib == EGGPROPS_IR_CS17) {
        CK(upload(c, libs->cs17_tdust, libs->nir_sed, &p.cs_tdust));
        CK(upload(c, libs->cs17_lir_dust, libs->nir_sed, &p.cs_lir_dust));
        CK(upload(c, libs->cs17_lir_pah, libs->nir_sed, &p.cs_lir_pah));
        CK(upload(c, libs->cs17_l8_dust, libs->nir_sed, &p.cs_l8_dust));
        CK(upload(c, libs->cs17_l8_pah, libs->nir_sed, &p.cs_l8_pah));
    }
    for (int i = 0; i < 12; ++i) p.ratio_x[i] = kRatioX[i];
    trapezoid_cdf(kRatioX, kBulgeP, p.bulge_cdf);
    trapezoid_cdf(kRatioX, kDiskP, p.disk_cdf);
    gauss_legendre(32, p.gl_x, p.gl_w);
    for (int l = 0; l < EGGPROPS_NLINE; ++l) {
        const double lam = (double)h_line_lambda[l];
        p.dust_law[l] = l == L_LYA ? 0.0 : calzetti2000(lam); // Ly-alpha: extinction already folded in, :2005
        p.igm_band[l] = lam < 0.0600 ? -2 : (lam < 0.0912 ? 0 : (lam < 0.1026 ? 1 : (lam < 0.1216 ? 2 : -1)));
    }
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->ready = true;
    return EGGPHOT_OK;
}

void eggprops_destroy(eggprops_ctx *c) {
    if (!c) return;
    if (c->ready) cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    for (void *p : c->dev) cudaFree(p);
    if (c->stage) cudaFree(c->stage);
    delete c;
}

int eggprops_compute_device(eggprops_ctx *c, const float *z, const float *m, const uint8_t *passive, size_t ngal,
                            uint64_t first_index, const eggprops_out *out, void *stream) {
    if (!c) return EGGPHOT_ERR_ARG;
    if (!c->ready) return c->fail(EGGPHOT_ERR_CUDA, "context not ready: " + c->err);
    if (!z || !m || !passive || !out) return c->fail(EGGPHOT_ERR_ARG, "null argument");
    if (ngal == 0) return EGGPHOT_OK;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return c->fail(EGGPHOT_ERR_CUDA, cudaGetErrorString(e));
    PropsParams p = c->base;
    p.z = z; p.m = m; p.passive = passive; p.ngal = ngal; p.first_index = first_index; p.out = *out;
    const unsigned blocks = (unsigned)((ngal + 127) / 128);
    props_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(p);
    if ((e = cudaGetLastError()) != cudaSuccess) return c->fail(EGGPHOT_ERR_CUDA, std::string("props kernel launch: ") + cudaGetErrorString(e));
    return EGGPHOT_OK;
}

int eggprops_compute(eggprops_ctx *c, const float *z, const float *m, const uint8_t *passive, size_t ngal,
                     uint64_t first_index, const eggprops_out *out) {
    if (!c) return EGGPHOT_ERR_ARG;
    if (!c->ready) return c->fail(EGGPHOT_ERR_CUDA, "context not ready: " + c->err);
    if (!z || !m || !passive || !out) return c->fail(EGGPHOT_ERR_ARG, "null argument");
    if (ngal == 0) return EGGPHOT_OK;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return c->fail(EGGPHOT_ERR_CUDA, cudaGetErrorString(e));
    // one device block: inputs then every requested output column, 256-byte aligned
    struct Col { void *host; size_t bytes; size_t off; void **slot; };
    std::vector<Col> cols;
    eggprops_out dout;
    std::memset(&dout, 0, sizeof(dout));
    size_t off = 0;
    auto place = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t oz = place(ngal * 4), om = place(ngal * 4), op = place(ngal);
#define COL(name, elem) if (out->name) cols.push_back(Col{(void *)out->name, ngal * (elem), place(ngal * (elem)), (void **)&dout.name})
    COL(d, 4); COL(bt, 4); COL(m_bulge, 4); COL(m_disk, 4); COL(disk_angle, 4); COL(bulge_angle, 4); COL(disk_ratio, 4); COL(bulge_ratio, 4);
    COL(disk_radius, 4); COL(bulge_radius, 4); COL(sfr, 4); COL(rsb, 4); COL(irx, 4); COL(sfrir, 4); COL(sfruv, 4);
    COL(rfuv_disk, 4); COL(rfvj_disk, 4); COL(rfuv_bulge, 4); COL(rfvj_bulge, 4); COL(m2lcor, 4); COL(opt_sed_disk, 8); COL(opt_sed_bulge, 8);
    COL(av_disk, 4); COL(av_bulge, 4); COL(tdust, 4); COL(ir8, 4); COL(lir, 4); COL(fpah, 4); COL(mdust, 4); COL(ir_sed, 8);
    COL(igm_abs, 12); COL(avlines_disk, 4); COL(avlines_bulge, 4); COL(vdisp, 4);
    COL(line_lum_disk, 4 * EGGPROPS_NLINE); COL(line_lum_bulge, 4 * EGGPROPS_NLINE);
#undef COL
    if (off > c->stage_bytes) {
        if (c->stage) cudaFree(c->stage);
        c->stage = nullptr; c->stage_bytes = 0;
        if ((e = cudaMalloc(&c->stage, off)) != cudaSuccess) return c->fail(EGGPHOT_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        c->stage_bytes = off;
    }
    char *base = (char *)c->stage;
    for (Col &k : cols) *k.slot = base + k.off;
    CK(cudaMemcpyAsync(base + oz, z, ngal * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(base + om, m, ngal * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(base + op, passive, ngal, cudaMemcpyHostToDevice, c->stream));
    const int rc = eggprops_compute_device(c, (const float *)(base + oz), (const float *)(base + om), (const uint8_t *)(base + op), ngal,
                                           first_index, &dout, (void *)c->stream);
    if (rc != EGGPHOT_OK) return rc;
    for (Col &k : cols) CK(cudaMemcpyAsync(k.host, base + k.off, k.bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
#undef CK
    return EGGPHOT_OK;
}

} // extern "C"
