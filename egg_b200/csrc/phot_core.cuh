// phot_core.cuh — the quadrature arithmetic shared by the device kernels, the host-side table
// builder and the CPU-tier emulation of the device algorithm (tests/host_emul).
//
// What is computed.  [vif] sed2flux(filter, lam, sed) as called at src/egg-gencat.cpp:2155 and
// :2197 is the trapezoid rule applied to g = R * S on the UNION of the filter nodes and the SED
// nodes inside the filter span, with R and S each interpolated linearly between their own nodes
// (node set A).  The reference walks that union node by node (~1000 nodes per
// galaxy-band).  Here the sum is transformed into  T = sum_j S_j W_j  over the SED nodes of the
// band's window, with weights W_j that depend on (redshift, band) only — every component of a
// galaxy shares them — and that cost one evaluation of a per-band piece-wise quadratic per node,
// whatever the number of filter nodes.
//
// Derivation.  Let t_j be the observed positions of the SED nodes, S linear on [t_j, t_{j+1}] with
// slope sigma_j; K0 = cumulative integral of R, L = cumulative integral of K0 (K0 = L = 0 left of
// the filter; K0 = tot0 and L linear right of it).  On one SED interval
//     union trapezoid = exact integral + (sigma_j/6) * sum over the union cells inside of s_cell * h^3
// (R*S is quadratic in a union cell, second derivative 2 s sigma; the trapezoid error is h^3 f''/12).
//  * exact integral, by parts:     [K0 S] - sigma_j (L(t_{j+1}) - L(t_j))
//  * union cells of the interval:  a partial filter cell at each end, whole filter cells between.
//    With G(t) = sum over whole cells left of t of s dl^3, plus s u^3 in the cell holding t
//    (u = offset of t in its cell):  sum s h^3 = G(t_{j+1}) - G(t_j) - 3 gamma_j,
//        gamma_j = s_c b u (b - u),   b = min(t_{j+1} - f_c, dl_c)
//    (c = filter cell of t_j; b - u is the distance from t_j to the next union node).
// With Lambda = L - G/6, which inside filter cell c is the QUADRATIC
//     Lambda(t) = Lambda_c + u (K0_c + u R_c/2)        (the cubic terms of L and G/6 cancel),
// and E_j = (Lambda(t_{j+1}) - Lambda(t_j) + gamma_j/2) / (t_{j+1} - t_j), the sum over the intervals of a
// window jlo..jhi (t_jlo <= f_first, t_jhi >= f_last), re-collected on the nodes, is
//
//     T = sum_{j=jlo..jhi} S_j W_j,      W_j = E_j - E_{j-1},   E_{jlo-1} = 0,  E_jhi = tot0
//
// (the [K0 S] terms telescope to tot0 S_jhi).  Lambda = 0 and u = 0 at jlo (left of the filter).  The result
// equals the union-node trapezoid in exact arithmetic; W_j is the quadrature weight of node j.
//
// Two-sided gauge.  Adding a linear function to Lambda shifts every E by a constant and leaves the
// W untouched.  Cells left of a split node m (the half-mass point of the band) keep the Lambda above,
// which vanishes towards the left end; cells right of it store Lambda - asym(t), asym(t) = lam_end +
// tot0 (t - f_last), which vanishes towards the right end (there E_jhi = 0).  A far wing of the filter
// then only ever differences small table values, however bright the SED is there (a Lyman break seen
// through the 1e-7 red wing of a blue band).  Only the SED interval [t_b, t_{b+1}] that holds the split
// (b = last node with t_b < f_m) mixes the gauges: E_b takes Lambda~(t_{b+1}) + asym(t_{b+1}), and
// node b+1 sees E_b - tot0.
//
// Numerics.  The E_j are first differences of a function that grows like tot0*W across the band
// and the W_j differences of those: they are formed in IEEE double in BOTH arithmetic modes (B200
// issues FP64 at half the FP32 rate), then rounded to the mode's real type.  The W_j of a
// non-negative response are non-negative up to the small trapezoid-error term, so sum S_j W_j has no
// cancellation: in the fp32 mode the SEDs, the gamma terms and the products S_j W_j are float, the
// per-lane partial sums are added across lanes in double.
#ifndef EGGPHOT_PHOT_CORE_CUH
#define EGGPHOT_PHOT_CORE_CUH

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define EGG_HD __host__ __device__ __forceinline__
#else
#define EGG_HD inline
#endif

#include "../../include/eggphot.h"

namespace eggphot {

// Galaxies a CTA carries through all band groups at a time (powers of two): the more, the fewer table loads per
// galaxy, as long as the SEDs of all resident batches (148 CTAs x batch x ~14 KB in fp32, twice that in fp64) stay mostly in L2.
#ifndef EGG_KBATCH_F32
#define EGG_KBATCH_F32 32
#endif
#ifndef EGG_KBATCH_F64
#define EGG_KBATCH_F64 32
#endif
constexpr int kBatchF32 = EGG_KBATCH_F32, kBatchF64 = EGG_KBATCH_F64;
constexpr int batch_of(int precision) { return precision == EGGPHOT_FP64 ? kBatchF64 : kBatchF32; }

// Band tables are structure-of-arrays in shared memory (a warp's lanes sit on neighbouring filter cells:
// 8-byte strides spread them over the banks).  Per band, n = nodes of the effective span:
//   f[n+1]   double  centred node positions lambda_c - fc; f[n] = +inf (sentinel of the cell scan)
//   lk[n]    double x2 {Lambda, K0} at the node (two-sided gauge, see above), one 16-byte record per cell;
//                    entry n-1 is the open cell right of the last node (lam = k0 = 0 in its gauge)
//   hr[n]    real    R_c / 2 (0 in the open cell)
//   sdl[n]   real x2 slope of R in the cell (0 in zero-width cells and in the open cell), cell width (huge in the
//                    open cell)
//   bkt[nbkt] uint16 last node at or below the start of each bucket
// The per-cell arrays are stored with one padding slot after every 16 entries (cell c lives in slot
// c + c/16): lanes on every 2nd / 4th / ... cell (a filter sampled more finely than the SED) then fall on
// distinct banks instead of colliding every 16 cells.
template <typename real> struct alignas(2 * sizeof(real)) Pair2 { real x, y; };
EGG_HD int cell_slot(int c) { return c + (c >> 4); }

// Header of one band inside a group blob (device layout; offsets are bytes from the blob start).
struct __attribute__((aligned(16))) BandHeader {
    double fc;                    // origin of the band-centred coordinates
    double full_first, full_last; // full filter span [um] for the coverage test ([vif] sed2flux NaN rule)
    double eff_first, eff_last;   // span with the zero-response wings removed [um]
    double f_first, f_last;       // centred ends of the effective span
    double tot0;                  // integral of R over the band
    double asym0;                 // asym(t) = asym0 + tot0 t: Lambda (left gauge) right of the filter
    double f_split;               // centred position of the split node: cells from it on use the right gauge
    float bkt_inv_w;              // buckets per unit wavelength, rounded so that the bucket index never overshoots
    float cost;                   // scheduling weight (typical number of SED nodes under the band)
    float l2first, l2last;        // log2(eff_first), log2(eff_last) in lookup buckets (x kLutPerOctave)
    uint32_t n;                   // nodes in the effective span (>= 2), or 0 if the response is identically 0
    uint32_t nbkt;                // buckets
    uint32_t f_off, lam_off, k0_off, hr_off, sdl_off, bkt_off; // byte offsets of the arrays
    uint32_t out_col;             // column in the sorted band list
    uint32_t split;               // index of the split node; bit 31: one probe settles the cell lookup (buckets are
                                  // narrower than every cell)
    float l2split;                // log2(fc + f_split) in lookup buckets: where the split node is looked up
    uint32_t pad_;
};
static_assert(sizeof(BandHeader) == 144, "BandHeader layout");

// One node of the rest-frame disk grid as phase B reads it.  GridA {x, idx}: wavelength and
// 1/(x[j+1] - x[j]) (0 for the last node), one 16-byte record per node so that a warp's loads are contiguous.
// The trapezoid-error term also needs the width x[j+1] - x[j] and the distance to the next OPTICAL node if j
// is one, else 0 (the bulge's union cells start at optical nodes only): fp64 reads both from GridBD, fp32
// reads only the latter (a float array) and takes the width as 1/idx.
struct __attribute__((aligned(16))) GridA { double x, idx; };
struct __attribute__((aligned(16))) GridBD { double dx, dxo; };

// Band seen through the kernel's real type.
struct alignas(16) LamK0 { double lam, k0; }; // Lambda and K0 of a cell: one 16-byte shared-memory load
template <typename real> struct BandView {
    const double *f;
    const LamK0 *lk;
    const real *hr;
    const Pair2<real> *sdl;
    const uint16_t *bkt;
    double fc, f_first, f_split, tot0, asym0;
    float bkt_inv_w;
    int n, nbkt;
    bool one_probe;
};
template <typename real> EGG_HD void make_band_view(const uint8_t *blob, const BandHeader &h, BandView<real> &v) {
    v.f = reinterpret_cast<const double *>(blob + h.f_off);
    v.lk = reinterpret_cast<const LamK0 *>(blob + h.lam_off);
    v.hr = reinterpret_cast<const real *>(blob + h.hr_off);
    v.sdl = reinterpret_cast<const Pair2<real> *>(blob + h.sdl_off);
    v.bkt = reinterpret_cast<const uint16_t *>(blob + h.bkt_off);
    v.fc = h.fc; v.f_first = h.f_first; v.f_split = h.f_split; v.tot0 = h.tot0; v.asym0 = h.asym0;
    v.bkt_inv_w = h.bkt_inv_w;
    v.n = (int)h.n; v.nbkt = (int)h.nbkt;
    v.one_probe = (h.split >> 31) != 0;
}

// Filter cell of the centred position t, f_first <= t: the largest c with f_c <= t, decided in double in both
// modes.  (Which side of a filter node an SED node falls on must be exact: the split node separates the two
// gauges, and next to a steep filter edge a node one cell off would weigh a bright SED with the wrong slope.
// A float decision was tried: 1 entry in 1e5 off by 5e-4.)  bkt[k] holds the last node at or below the start of
// bucket k and the bucket index never overshoots (the table is built against slightly lowered bucket edges), so
// a forward scan settles it; the sentinel f[n] = +inf ends the scan in the open cell.  When the buckets are
// narrower than every cell (one_probe) the scan is a single comparison.
template <typename real> EGG_HD int find_cell(const BandView<real> &b, double t) {
    int k = (int)((float)(t - b.f_first) * b.bkt_inv_w);
    k = k < 0 ? 0 : (k > b.nbkt - 1 ? b.nbkt - 1 : k);
    int c = b.bkt[k];
    if (b.one_probe) return c + (t >= b.f[cell_slot(c + 1)] ? 1 : 0);
    while (t >= b.f[cell_slot(c + 1)]) ++c;
    return c;
}

// Lambda(t) in cell c and the offset u of t inside it.
template <typename real> EGG_HD double lambda_in_cell(const BandView<real> &b, int c, double t, double &u) {
    const int cs = cell_slot(c);
    u = t - b.f[cs];
    const LamK0 lk = b.lk[cs];
    return fma(u, fma(u, (double)b.hr[cs], lk.k0), lk.lam);
}

// gamma = s b u (b - u), b = min(dt + u, dl): the trapezoid-error term of the union cell that starts at
// an SED node at offset u inside a filter cell of slope s and width dl, the next node of the same SED
// being dt further (0 for a node that does not start an interval of this SED: gamma = 0).
// b - u is formed as min(dt, dl - u), not as (dt + u) - u: in a wide filter cell (u >> dt) the latter loses dt's
// low bits in float (1.4e-5 on a three-node 35 um band found by fuzzing).
template <typename real> EGG_HD real gamma_term(real s, real dl, real u, real dt) {
    const real rest = dl - u, bmu = dt < rest ? dt : rest; // b - u
    return s * (u + bmu) * u * bmu;
}

// IGM factor of node j (src/egg-gencat.cpp:2128-2139): [0,lz) -> 0, [lz,l0) -> a0, [l0,l1) -> a1, [l1,l2) -> a2
template <typename real>
EGG_HD real igm_factor(int j, int lz, int l0, int l1, int l2, real a0, real a1, real a2) {
    return j >= l2 ? real(1) : (j >= l1 ? a2 : (j >= l0 ? a1 : (j >= lz ? a0 : real(0))));
}

// [vif] integrate_gauss over the wavelength bin [llow, lup] of a line at l0 of width lw and area amp
EGG_HD double line_bin(double llow, double lup, double l0, double lw, double amp, int average) {
    const double is = 0.70710678118654752440 / lw;
    const double v = 0.5 * amp * (erf((lup - l0) * is) - erf((llow - l0) * is));
    return average ? v / (lup - llow) : v;
}

// The same with the erf difference in the kernel's precision (the arguments are still formed in
// double: they are differences of nearly equal wavelengths).
template <typename real> EGG_HD double line_bin_t(double llow, double lup, double l0, double lw, double amp, int average);
template <> EGG_HD double line_bin_t<double>(double llow, double lup, double l0, double lw, double amp, int average) {
    return line_bin(llow, lup, l0, lw, amp, average);
}
template <> EGG_HD double line_bin_t<float>(double llow, double lup, double l0, double lw, double amp, int average) {
    const double is = 0.70710678118654752440 / lw;
    const float a = (float)((lup - l0) * is), b = (float)((llow - l0) * is);
    // erf(a) - erf(b): use erfc on the far side so that a bin deep in a wing keeps its relative accuracy
    float d;
    if (b >= 0.0f) d = erfcf(b) - erfcf(a);
    else if (a <= 0.0f) d = erfcf(-a) - erfcf(-b);
    else d = erff(a) - erff(b);
    const double v = 0.5 * amp * (double)d;
    return average ? v / (lup - llow) : v;
}

// Compensated accumulation (TwoSum, branch-free), used for the rest-frame dot products.
template <typename real> struct CompSum {
    real s, c;
    EGG_HD void clear() { s = real(0); c = real(0); }
    EGG_HD void add(real x) {
        const real t = s + x;
        const real z = t - s;
        c += (s - (t - z)) + (x - z);
        s = t;
    }
    EGG_HD void add(const CompSum &o) { add(o.s); c += o.c; }
    EGG_HD real value() const { return s + c; }
};

// [vif] lsun2uJy constant: 1e32 * Lsun / (c * 4 pi Mpc^2)  with Mpc = 3.0856e22 m, Lsun = 3.839e26 W,
// c = 2.9979e14 um/s  (src/egg-gencat.cpp:2152, 2162 call sites).
#define EGGPHOT_LSUN2UJY (1.0e32 * 3.839e26 / (2.9979e14 * 4.0 * 3.14159265358979323846 * 3.0856e22 * 3.0856e22))

} // namespace eggphot
#endif
