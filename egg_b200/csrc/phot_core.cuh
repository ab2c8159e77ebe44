// phot_core.cuh — the quadrature arithmetic shared by the device kernels and the host-side
// table builder (which needs the same formulas in double for the rest-frame weights).
//
// What is computed.  [vif] sed2flux(filter, lam, sed) as called at src/egg-gencat.cpp:2155 and
// :2197 is the trapezoid rule applied to g = R * S on the UNION of the filter nodes and the SED
// nodes inside the filter span, with R and S each interpolated linearly between their own nodes
// (node set A).  The reference walks that union node by node; here the sum is
// regrouped per SED interval [t_j, t_{j+1}] so that its cost no longer depends on the number of
// filter nodes.  On that interval S(u) = S_j + sigma*(u - t_j), hence
//
//     T_j = S_j * q_j + sigma_j * m_j
//     q_j = trapz_union(R)            = K0(t_{j+1}) - K0(t_j)          (R is piece-wise linear: exact)
//     m_j = trapz_union(R * (u - t_j)) = K1(t_{j+1}) - K1(t_j) - t_j*q_j - corr_j
//
// where K0(t) = I0[i] + (R_i + rho(t))/2 * (t - tf_i) and K1(t) = C1[i] + (R_i*tf_i + rho(t)*t)/2 * (t - tf_i)
// are the cumulative trapezoids of R and R*u over the filter nodes with one extra ("virtual") node
// at t, i being the filter cell that holds t.  I0/C1 are per-band prefix tables.  K1's virtual
// node at the LEFT end of an SED interval splits a filter cell that the union grid does not
// split there in the same way; because R*u is quadratic inside a cell the trapezoid is not
// additive, and the difference is the closed form
//
//     corr_j = (s_i/2) * (t_j - tf_i) * (B - t_j) * (B - tf_i),   B = min(t_{j+1}, tf_{i+1})
//
// (s_i = slope of R in cell i).  Everything is evaluated in band-centred coordinates
// u = lambda_obs - fc to keep |u| small, and the prefix tables are stored as hi+lo pairs with a
// power-of-two offset/scale into [1,2) so that differences of their high parts are exact in
// fp32 (Sterbenz) — the sigma*m term is a difference of large cumulative moments and would
// otherwise lose up to w^2/dt^2 ~ 1e5 ulps.
#ifndef EGGPHOT_PHOT_CORE_CUH
#define EGGPHOT_PHOT_CORE_CUH

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define EGG_HD __host__ __device__ __forceinline__
#else
#define EGG_HD inline
#endif

namespace eggphot {

// One filter node of a band table.  fp32: prefix values as offset/scaled hi+lo pairs.
struct __attribute__((aligned(16))) FilterNodeF {
    float tf;   // centred node position  f_i - fc
    float R;    // response at the node
    float s;    // slope of R in cell [i, i+1]  (0 for the last node and zero-width cells)
    float dl;   // cell width f_{i+1} - f_i    (0 for the last node)
    float I0h, I0l; // 1.5 + I0/NI  as hi+lo,  I0 = cumulative trapezoid of R
    float C1h, C1l; // 1.5 + C1/NC  as hi+lo,  C1 = cumulative trapezoid of R*u (centred u)
};
// fp64: plain values.
struct __attribute__((aligned(16))) FilterNodeD {
    double tf, R, s, dl, I0, C1;
};

// Header of one band inside a group blob (device layout; offsets are bytes from the blob start).
struct __attribute__((aligned(16))) BandHeader {
    double fc;                    // centre of the band-centred coordinates (exactly representable in f32)
    double full_first, full_last; // full filter span [um] for the coverage test ([vif] sed2flux NaN rule)
    double eff_first, eff_last;   // span with the zero-response wings removed [um]
    double ni, nc;                // power-of-two normalisers of the prefix tables
    double tf0, tfl;              // centred ends of the effective span (= tf of first / last node)
    double bkt_inv_w;             // buckets per unit wavelength
    double tot0, tot1;            // band totals of the two prefix tables (see `split`)
    float fcf, nif, ncf, tf0f, tflf, bkt_inv_wf; // f32 mirrors of the above
    float tot0h, tot0l, tot1h, tot1l; // tot0/NI, tot1/NC as hi+lo; hi is a multiple of 2^-23
    float inv_nif, inv_ncf, ni_ncf, cost; // 1/NI, 1/NC, NI/NC; cost = scheduling weight (typical node count)
    uint32_t n;                   // nodes in the effective span (>= 2), or 0 if the response is identically 0
    uint32_t nbkt;                // buckets
    uint32_t node_off;            // byte offset of the FilterNode array
    uint32_t bkt_off;             // byte offset of the uint16 bucket array (nbkt + 1 entries)
    uint32_t out_col;             // column in the sorted band list
    uint32_t split;               // cells 0..split accumulate from the left end, cells > split from the right
    uint32_t pad[2];
};
static_assert(sizeof(BandHeader) == 192, "BandHeader layout");

template <typename real> struct NodeOf;
template <> struct NodeOf<float> { typedef FilterNodeF type; };
template <> struct NodeOf<double> { typedef FilterNodeD type; };

// Cumulative quantities at one SED node (the only things an interval needs from its two ends).
template <typename real> struct NodeQ;
template <> struct NodeQ<float> {
    float k0h, k0l, k1h, k1l; // K0, K1 in offset/scaled units, hi+lo
};
template <> struct NodeQ<double> {
    double k0, k1;
};

// Per-band scalars the arithmetic needs.
template <typename real> struct BandScal {
    real inv_ni;   // 1/NI (power of two)
    real inv_nc;   // 1/NC (power of two)
    real ni;       // NI
    real nc;       // NC
    real ni_nc;    // NI/NC (power of two)
    real tot0h, tot0l, tot1h, tot1l; // band totals in table units (fp32: hi+lo; fp64: value in *h, 0 in *l)
    int split;     // last cell whose prefix values are measured from the left end
};

// K0/K1 at effective position te (already clamped into the band's effective span) inside cell `e`.
// Also returns u = te - tf_i.
EGG_HD void node_quantities(const FilterNodeF &e, float te, const BandScal<float> &bs, NodeQ<float> &q, float &u) {
    u = te - e.tf;
    const float rho = fmaf(e.s, u, e.R);
    const float p0 = 0.5f * (e.R + rho) * u * bs.inv_ni;
    const float p1 = 0.5f * fmaf(e.R, e.tf, rho * te) * u * bs.inv_nc;
    // FastTwoSum(|hi| >= |p|): hi parts live in [1,2), partial cells are at most the band integral / NI <= 1/4
    const float s0 = e.I0h + p0;
    q.k0h = s0;
    q.k0l = e.I0l + (p0 - (s0 - e.I0h));
    const float s1 = e.C1h + p1;
    q.k1h = s1;
    q.k1l = e.C1l + (p1 - (s1 - e.C1h));
}
EGG_HD void node_quantities(const FilterNodeD &e, double te, const BandScal<double> &, NodeQ<double> &q, double &u) {
    u = te - e.tf;
    const double rho = fma(e.s, u, e.R);
    q.k0 = e.I0 + 0.5 * (e.R + rho) * u;
    q.k1 = e.C1 + 0.5 * fma(e.R, e.tf, rho * te) * u;
}

// Contribution T_j of one SED interval.
//   left node : true position tl, effective (clamped) position tel, value Sl, quantities ql,
//               its filter cell (s, dl) and offset ul = tel - tf_i, cell index il
//   right node: tr, ter, Sr, qr, cell index ir
// Returns S_l*q + sigma*m in the band's natural units (response * wavelength * S).
EGG_HD float interval_flux(float tl, float tel, float Sl, const NodeQ<float> &ql, float s_l, float dl_l, float ul,
                           int il, float tr, float ter, float Sr, const NodeQ<float> &qr, int ir,
                           const BandScal<float> &bs) {
    const float dt = tr - tl;
#if defined(__CUDA_ARCH__)
    const float sigma = dt > 0.0f ? __fdividef(Sr - Sl, dt) : 0.0f; // 2 ulp; enters only the slope term
#else
    const float sigma = dt > 0.0f ? (Sr - Sl) / dt : 0.0f;
#endif
    // prefix values left of `split` are measured from the band's left end, right of it from the
    // right end (so that a far wing is not differenced against the whole band): an interval that
    // straddles the split adds the band total.  All high parts are multiples of 2^-23 below 1 in
    // magnitude, so these sums stay exact.
    const bool cross = il <= bs.split && ir > bs.split;
    const float d0h = (qr.k0h - ql.k0h) + (cross ? bs.tot0h : 0.0f);
    const float d0l = (qr.k0l - ql.k0l) + (cross ? bs.tot0l : 0.0f);
    const float d1h = (qr.k1h - ql.k1h) + (cross ? bs.tot1h : 0.0f);
    const float d1l = (qr.k1l - ql.k1l) + (cross ? bs.tot1l : 0.0f);
    const float q = (d0h + d0l) * bs.ni;
    // m = NC*dK1 - tl*NI*dK0 - corr, evaluated as NC*[ (dK1h - tau*dK0h) + (dK1l - tau*dK0l) ]
    const float tau = tl * bs.ni_nc;
    const float r = fmaf(-tau, d0h, d1h) + fmaf(-tau, d0l, d1l);
    const float bmt = (ir == il) ? (ter - tel) : (dl_l - ul); // B - t_l
    const float corr = 0.5f * s_l * ul * bmt * (bmt + ul);
    const float m = fmaf(r, bs.nc, -corr);
    return fmaf(Sl, q, sigma * m);
}
EGG_HD double interval_flux(double tl, double tel, double Sl, const NodeQ<double> &ql, double s_l, double dl_l,
                            double ul, int il, double tr, double ter, double Sr, const NodeQ<double> &qr, int ir,
                            const BandScal<double> &bs) {
    const double dt = tr - tl;
    const double sigma = dt > 0.0 ? (Sr - Sl) / dt : 0.0;
    const bool cross = il <= bs.split && ir > bs.split;
    const double q = (qr.k0 - ql.k0) + (cross ? bs.tot0h : 0.0);
    const double r = fma(-tl, q, (qr.k1 - ql.k1) + (cross ? bs.tot1h : 0.0));
    const double bmt = (ir == il) ? (ter - tel) : (dl_l - ul);
    const double corr = 0.5 * s_l * ul * bmt * (bmt + ul);
    return fma(Sl, q, sigma * (r - corr));
}

// Compensated accumulation (TwoSum, branch-free): the "compensated sum" of the fp32 mode.
template <typename real> struct CompSum {
    real s, c; // plain aggregate (lives in __shared__ arrays): call clear() before use
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

// Cell lookup: largest i in [0, n-2] with tf[i] <= te.  bkt[k] is the last node at or below the
// start of bucket k, so the answer is a short forward scan away (buckets are finer than the node
// spacing); the backward guard only fires when the rounding of k lands one bucket too far.
// tfget(i) returns tf of node i.  te must lie inside [tf[0], tf[n-1]].
template <typename real, typename TfGet>
EGG_HD int find_cell(TfGet tfget, const uint16_t *bkt, int nbkt, real tf0, real inv_w, int n, real te) {
    int k = (int)((te - tf0) * inv_w);
    k = k < 0 ? 0 : (k > nbkt - 1 ? nbkt - 1 : k);
    int lo = bkt[k];
    lo = lo > n - 2 ? n - 2 : lo;
    if (tfget(lo) > te) {
        do { --lo; } while (lo > 0 && tfget(lo) > te);
        return lo < 0 ? 0 : lo;
    }
    while (lo < n - 2 && tfget(lo + 1) <= te) ++lo;
    return lo;
}

// Band seen through the kernel's real type.
template <typename real> struct BandView {
    const typename NodeOf<real>::type *nodes;
    const uint16_t *bkt;
    BandScal<real> bs;
    real fc, tf0, tfl, bkt_inv_w;
    int n, nbkt;
};
EGG_HD void make_band_view(const uint8_t *blob, const BandHeader &h, BandView<float> &v) {
    v.nodes = reinterpret_cast<const FilterNodeF *>(blob + h.node_off);
    v.bkt = reinterpret_cast<const uint16_t *>(blob + h.bkt_off);
    v.bs.ni = h.nif; v.bs.nc = h.ncf; v.bs.inv_ni = h.inv_nif; v.bs.inv_nc = h.inv_ncf; v.bs.ni_nc = h.ni_ncf;
    v.bs.tot0h = h.tot0h; v.bs.tot0l = h.tot0l; v.bs.tot1h = h.tot1h; v.bs.tot1l = h.tot1l; v.bs.split = (int)h.split;
    v.fc = h.fcf; v.tf0 = h.tf0f; v.tfl = h.tflf; v.bkt_inv_w = h.bkt_inv_wf;
    v.n = (int)h.n; v.nbkt = (int)h.nbkt;
}
EGG_HD void make_band_view(const uint8_t *blob, const BandHeader &h, BandView<double> &v) {
    v.nodes = reinterpret_cast<const FilterNodeD *>(blob + h.node_off);
    v.bkt = reinterpret_cast<const uint16_t *>(blob + h.bkt_off);
    v.bs.ni = v.bs.nc = v.bs.inv_ni = v.bs.inv_nc = v.bs.ni_nc = 1.0;
    v.bs.tot0h = h.tot0; v.bs.tot1h = h.tot1; v.bs.tot0l = v.bs.tot1l = 0.0; v.bs.split = (int)h.split;
    v.fc = h.fc; v.tf0 = h.tf0; v.tfl = h.tfl; v.bkt_inv_w = h.bkt_inv_w;
    v.n = (int)h.n; v.nbkt = (int)h.nbkt;
}

// Observed-frame, band-centred position of a rest-frame node: x*(1+z) - fc.
// fp32: x and (1+z) are carried as hi+lo pairs so that the only rounding is on the (small) result.
EGG_HD float node_pos(float xhi, float xlo, float opz_hi, float opz_lo, float fc) {
    return fmaf(xhi, opz_hi, -fc) + fmaf(xhi, opz_lo, xlo * opz_hi);
}
EGG_HD double node_pos(double x, double opz, double fc) { return fma(x, opz, -fc); }

EGG_HD float clampr(float t, float lo, float hi) { return fminf(fmaxf(t, lo), hi); }
EGG_HD double clampr(double t, double lo, double hi) { return fmin(fmax(t, lo), hi); }

// Everything an interval needs to know about one of its end nodes.
template <typename real> struct NodeState {
    real t;   // true position
    real S;   // SED value (x * L) at the node
    real u;   // te - tf[cell], te = t clamped into the effective span
    real s, dl; // slope and width of the filter cell
    int cell;
    NodeQ<real> q;
};

template <typename real>
EGG_HD void eval_node(const BandView<real> &b, real t, real S, NodeState<real> &ns) {
    ns.t = t;
    ns.S = S;
    const real te = clampr(t, b.tf0, b.tfl);
    const typename NodeOf<real>::type *nodes = b.nodes;
    int c = find_cell<real>([nodes](int i) { return nodes[i].tf; }, b.bkt, b.nbkt, b.tf0, b.bkt_inv_w, b.n, te);
    const typename NodeOf<real>::type e = nodes[c];
    ns.cell = c;
    ns.s = e.s;
    ns.dl = e.dl;
    node_quantities(e, te, b.bs, ns.q, ns.u);
}

template <typename real>
EGG_HD real eval_interval(const BandView<real> &b, const NodeState<real> &l, const NodeState<real> &r) {
    const real tel = clampr(l.t, b.tf0, b.tfl);
    const real ter = clampr(r.t, b.tf0, b.tfl);
    return interval_flux(l.t, tel, l.S, l.q, l.s, l.dl, l.u, l.cell, r.t, ter, r.S, r.q, r.cell, b.bs);
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

// [vif] lsun2uJy constant: 1e32 * Lsun / (c * 4 pi Mpc^2)  with Mpc = 3.0856e22 m, Lsun = 3.839e26 W,
// c = 2.9979e14 um/s  (src/egg-gencat.cpp:2152, 2162 call sites).
#define EGGPHOT_LSUN2UJY (1.0e32 * 3.839e26 / (2.9979e14 * 4.0 * 3.14159265358979323846 * 3.0856e22 * 3.0856e22))

} // namespace eggphot
#endif
