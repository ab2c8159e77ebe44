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
    float l2first, l2last;        // log2(eff_first), log2(eff_last) in lookup buckets (x kLutPerOctave)
    uint32_t n;                   // nodes in the effective span (>= 2), or 0 if the response is identically 0
    uint32_t nbkt;                // buckets
    uint32_t node_off;            // byte offset of the FilterNode array
    uint32_t bkt_off;             // byte offset of the uint16 bucket array (nbkt + 1 entries)
    uint32_t out_col;             // column in the sorted band list
    uint32_t split;               // cells 0..split accumulate from the left end, cells > split from the right
    uint32_t tflo_off;            // fp32 tables: byte offset of the float array of low parts of tf
    uint32_t pad[3];
};
static_assert(sizeof(BandHeader) == 208, "BandHeader layout");

template <typename real> struct NodeOf;
template <> struct NodeOf<float> { typedef FilterNodeF type; };
template <> struct NodeOf<double> { typedef FilterNodeD type; };

// Cumulative quantities at one SED node (the only things an interval needs from its two ends).
template <typename real> struct NodeQ;
template <> struct NodeQ<float> {
    float i0h, i0l, c1h, c1l; // prefix values of the node's filter cell (offset/scaled, hi+lo)
    float p0;                 // integral of R over the part of the cell left of the node
    float hru;                // R_i * u^2 / 2
    float R;                  // R_i (used by the node's own interval only; never exchanged)
};
template <> struct NodeQ<double> {
    double i0, c1, p0, hru, R;
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

// What an interval needs from one of its end nodes, te being the node's effective position (clamped
// into the band's effective span), `e` its filter cell and u = te - tf_i.  With
//     K0(t) = I0_i + p0,            p0 = (R_i + rho(t))/2 * u
//     K1(t) = C1_i + p0*t - R_i*u^2/2         (the same trapezoid with a virtual node at t, rewritten)
// the moment about the LEFT node of an interval becomes
//     K1(t_r) - K1(t_l) - t_l*(K0(t_r) - K0(t_l)) = [dC1 - t_l*dI0] + p0_r*(t_r - t_l) - hru_r + hru_l
// in which only table values are differenced against each other (exactly, thanks to the hi+lo
// storage) and every other term is small and well conditioned.
EGG_HD void node_quantities(const FilterNodeF &e, float u, NodeQ<float> &q) {
    const float rho = fmaf(e.s, u, e.R);
    q.i0h = e.I0h; q.i0l = e.I0l; q.c1h = e.C1h; q.c1l = e.C1l;
    q.p0 = 0.5f * (e.R + rho) * u;
    q.hru = 0.5f * e.R * u * u;
    q.R = e.R;
}
EGG_HD void node_quantities(const FilterNodeD &e, double u, NodeQ<double> &q) {
    const double rho = fma(e.s, u, e.R);
    q.i0 = e.I0; q.c1 = e.C1;
    q.p0 = 0.5 * (e.R + rho) * u;
    q.hru = 0.5 * e.R * u * u;
    q.R = e.R;
}

// Geometry of one SED interval against the band: everything that does not depend on the SED values.
// The contribution of an SED with end values (S_l, S_r) is then
//     general   : S_l*q + sigma*m,                         sigma = (S_r - S_l)/(t_r - t_l)
//     same cell : (rho_l*S(tel) + rho_r*S(ter))/2 * dte    (both ends in one filter cell: the union grid has no
//                 node in between, the interval is a single trapezoid; evaluated directly it is exact, cannot go
//                 negative for non-negative R and S, and keeps full relative accuracy for a narrow SED feature
//                 inside a wide filter cell)
// so that several SEDs sampled on the same nodes (disk, and bulge resampled on the disk grid) share it.
template <typename real> struct IntervalW {
    real q, m, inv_dt;
    real rho_l, rho_r, dte, xs, xr; // xs / xr: effective left / right end minus the TRUE left position
    bool same;
};

//   left node : true position tl (+ low part tll in fp32), offset ul inside its filter cell, quantities ql,
//               its filter cell (s, dl), cell index il
//   right node: tr (+ trl), qr, cell index ir
//   dte = effective (clamped) interval width;  xr = effective right position minus TRUE left position
EGG_HD void interval_weights(float tl, float tll, const NodeQ<float> &ql, float s_l, float dl_l, float ul, int il,
                             float tr, float trl, const NodeQ<float> &qr, int ir, float dte, float xr,
                             const BandScal<float> &bs, IntervalW<float> &w) {
    const float dt = (tr - tl) + (trl - tll);
#if defined(__CUDA_ARCH__)
    w.inv_dt = dt > 0.0f ? __fdividef(1.0f, dt) : 0.0f; // MUFU.RCP, ~1 ulp; enters only the slope terms
#else
    w.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
#endif
    // prefix values left of `split` are measured from the band's left end, right of it from the
    // right end (so that a far wing is not differenced against the whole band): an interval that
    // straddles the split adds the band total.  All high parts are multiples of 2^-23 below 1 in
    // magnitude, so these sums stay exact.
    const bool cross = il <= bs.split && ir > bs.split;
    const float d0h = (qr.i0h - ql.i0h) + (cross ? bs.tot0h : 0.0f);
    const float d0l = (qr.i0l - ql.i0l) + (cross ? bs.tot0l : 0.0f);
    const float d1h = (qr.c1h - ql.c1h) + (cross ? bs.tot1h : 0.0f);
    const float d1l = (qr.c1l - ql.c1l) + (cross ? bs.tot1l : 0.0f);
    w.q = fmaf(d0h + d0l, bs.ni, qr.p0 - ql.p0);
    // [dC1 - t_l*dI0] in table units: NC*[ (dC1h - tau*dI0h) + (dC1l - tau*dI0l) ], tau = t_l*NI/NC in hi+lo
    const float tauh = tl * bs.ni_nc, taul = tll * bs.ni_nc;
    const float r = fmaf(-tauh, d0h, d1h) + fmaf(-tauh, d0l, fmaf(-taul, d0h, d1l));
    const float bmt = (ir == il) ? dte : (dl_l - ul); // B - t_l
    const float corr = 0.5f * s_l * ul * bmt * (bmt + ul);
    w.m = fmaf(r, bs.nc, fmaf(qr.p0, xr, (ql.hru - qr.hru) - corr));
    w.rho_l = fmaf(s_l, ul, ql.R);
    w.rho_r = fmaf(s_l, dte, w.rho_l);
    w.dte = dte; w.xr = xr; w.xs = xr - dte;
    w.same = ir == il;
}
EGG_HD void interval_weights(double tl, double, const NodeQ<double> &ql, double s_l, double dl_l, double ul, int il,
                             double tr, double, const NodeQ<double> &qr, int ir, double dte, double xr,
                             const BandScal<double> &bs, IntervalW<double> &w) {
    const double dt = tr - tl;
    w.inv_dt = dt > 0.0 ? 1.0 / dt : 0.0;
    const bool cross = il <= bs.split && ir > bs.split;
    const double d0 = (qr.i0 - ql.i0) + (cross ? bs.tot0h : 0.0);
    const double d1 = (qr.c1 - ql.c1) + (cross ? bs.tot1h : 0.0);
    w.q = d0 + (qr.p0 - ql.p0);
    const double bmt = (ir == il) ? dte : (dl_l - ul);
    const double corr = 0.5 * s_l * ul * bmt * (bmt + ul);
    w.m = fma(-tl, d0, d1) + fma(qr.p0, xr, (ql.hru - qr.hru) - corr);
    w.rho_l = fma(s_l, ul, ql.R);
    w.rho_r = fma(s_l, dte, w.rho_l);
    w.dte = dte; w.xr = xr; w.xs = xr - dte;
    w.same = ir == il;
}

// Contribution of an SED with end values (Sl, Sr) to the interval; also returns its slope.
EGG_HD float interval_apply(const IntervalW<float> &w, float Sl, float Sr, float &sigma) {
    sigma = (Sr - Sl) * w.inv_dt;
    const float gen = fmaf(Sl, w.q, sigma * w.m);
    const float Sel = fmaf(sigma, w.xs, Sl), Ser = fmaf(sigma, w.xr, Sl);
    const float same = 0.5f * fmaf(w.rho_l, Sel, w.rho_r * Ser) * w.dte;
    return w.same ? same : gen;
}
EGG_HD double interval_apply(const IntervalW<double> &w, double Sl, double Sr, double &sigma) {
    sigma = (Sr - Sl) * w.inv_dt;
    const double gen = fma(Sl, w.q, sigma * w.m);
    const double Sel = fma(sigma, w.xs, Sl), Ser = fma(sigma, w.xr, Sl);
    const double same = 0.5 * fma(w.rho_l, Sel, w.rho_r * Ser) * w.dte;
    return w.same ? same : gen;
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
    // buckets are finer than the cells: one or two steps almost always settle it
    if (lo < n - 2 && tfget(lo + 1) <= te) ++lo;
    if (lo < n - 2 && tfget(lo + 1) <= te) {
        do { ++lo; } while (lo < n - 2 && tfget(lo + 1) <= te);
    }
    return lo;
}

// Band seen through the kernel's real type.
template <typename real> struct BandView {
    const typename NodeOf<real>::type *nodes;
    const float *tflo; // fp32: low parts of the node positions
    const uint16_t *bkt;
    BandScal<real> bs;
    real fc, tf0, tfl, bkt_inv_w;
    int n, nbkt;
};
EGG_HD void make_band_view(const uint8_t *blob, const BandHeader &h, BandView<float> &v) {
    v.nodes = reinterpret_cast<const FilterNodeF *>(blob + h.node_off);
    v.tflo = reinterpret_cast<const float *>(blob + h.tflo_off);
    v.bkt = reinterpret_cast<const uint16_t *>(blob + h.bkt_off);
    v.bs.ni = h.nif; v.bs.nc = h.ncf; v.bs.inv_ni = h.inv_nif; v.bs.inv_nc = h.inv_ncf; v.bs.ni_nc = h.ni_ncf;
    v.bs.tot0h = h.tot0h; v.bs.tot0l = h.tot0l; v.bs.tot1h = h.tot1h; v.bs.tot1l = h.tot1l; v.bs.split = (int)h.split;
    v.fc = h.fcf; v.tf0 = h.tf0f; v.tfl = h.tflf; v.bkt_inv_w = h.bkt_inv_wf;
    v.n = (int)h.n; v.nbkt = (int)h.nbkt;
}
EGG_HD void make_band_view(const uint8_t *blob, const BandHeader &h, BandView<double> &v) {
    v.nodes = reinterpret_cast<const FilterNodeD *>(blob + h.node_off);
    v.tflo = nullptr;
    v.bkt = reinterpret_cast<const uint16_t *>(blob + h.bkt_off);
    v.bs.ni = v.bs.nc = v.bs.inv_ni = v.bs.inv_nc = v.bs.ni_nc = 1.0;
    v.bs.tot0h = h.tot0; v.bs.tot1h = h.tot1; v.bs.tot0l = v.bs.tot1l = 0.0; v.bs.split = (int)h.split;
    v.fc = h.fc; v.tf0 = h.tf0; v.tfl = h.tfl; v.bkt_inv_w = h.bkt_inv_w;
    v.n = (int)h.n; v.nbkt = (int)h.nbkt;
}

// Observed-frame, band-centred position of a rest-frame node: x*(1+z) - fc.
// fp32: x and (1+z) come as hi+lo pairs and the result is returned as a hi+lo pair (error-free
// product and sum), because what the quadrature needs is the position RELATIVE to filter nodes and
// to neighbouring SED nodes, i.e. differences ~1e-3 of numbers ~0.1-1: a single f32 would leave
// ~1e-5 of relative noise on every interval width, which shows on SEDs with sharp features.
EGG_HD void node_pos(float xhi, float xlo, float opz_hi, float opz_lo, float fc, float &th, float &tl) {
    const float ph = xhi * opz_hi;
    const float pe = fmaf(xhi, opz_hi, -ph);      // exact product error
    th = ph - fc;
    const float bb = th - ph;
    const float e2 = (ph - (th - bb)) + (-fc - bb); // exact sum error
    tl = fmaf(xlo, opz_hi, fmaf(xhi, opz_lo, pe + e2));
}
EGG_HD void node_pos(double x, double opz, double fc, double &th, double &tl) { th = fma(x, opz, -fc); tl = 0.0; }

EGG_HD float clampr(float t, float lo, float hi) { return fminf(fmaxf(t, lo), hi); }
EGG_HD double clampr(double t, double lo, double hi) { return fmin(fmax(t, lo), hi); }

// Everything an interval needs to know about one of its end nodes.
template <typename real> struct NodeState {
    real t, tl; // true position, hi + lo (lo is 0 in fp64)
    real S;     // SED value (x * L) at the node
    real u;     // te - tf[cell], te = t clamped into the effective span
    real s, dl; // slope and width of the filter cell
    int cell;
    NodeQ<real> q;
};

EGG_HD float cell_offset(const BandView<float> &b, int c, float tf, float te, float tel) { return (te - tf) + (tel - b.tflo[c]); }
EGG_HD double cell_offset(const BandView<double> &, int, double tf, double te, double) { return te - tf; }

template <typename real>
EGG_HD void eval_node(const BandView<real> &b, real t, real tl, real S, NodeState<real> &ns) {
    ns.t = t;
    ns.tl = tl;
    ns.S = S;
    const real te = clampr(t, b.tf0, b.tfl);
    const real tel = te == t ? tl : real(0); // a clamped node sits exactly on the span end
    const typename NodeOf<real>::type *nodes = b.nodes;
    int c = find_cell<real>([nodes](int i) { return nodes[i].tf; }, b.bkt, b.nbkt, b.tf0, b.bkt_inv_w, b.n, te);
    const typename NodeOf<real>::type e = nodes[c];
    ns.cell = c;
    ns.s = e.s;
    ns.dl = e.dl;
    ns.u = cell_offset(b, c, e.tf, te, tel);
    node_quantities(e, ns.u, ns.q);
}

template <typename real>
EGG_HD void eval_weights(const BandView<real> &b, const NodeState<real> &l, const NodeState<real> &r, IntervalW<real> &w) {
    const real tel = clampr(l.t, b.tf0, b.tfl), ter = clampr(r.t, b.tf0, b.tfl);
    const real terl = ter == r.t ? r.tl : real(0), tell = tel == l.t ? l.tl : real(0);
    const real dte = (ter - tel) + (terl - tell);
    const real xr = (ter - l.t) + (terl - l.tl);
    interval_weights(l.t, l.tl, l.q, l.s, l.dl, l.u, l.cell, r.t, r.tl, r.q, r.cell, dte, xr, b.bs, w);
}
template <typename real>
EGG_HD real eval_interval(const BandView<real> &b, const NodeState<real> &l, const NodeState<real> &r) {
    IntervalW<real> w;
    eval_weights(b, l, r, w);
    real sigma;
    return interval_apply(w, l.S, r.S, sigma);
}

// A second SED on the same weights: the bulge's optical template lives on a SUBSET of the disk grid's
// nodes (merge_add's union grid).  Resampled linearly onto the disk grid it is the same function, so its
// integral against R is unchanged; what changes is the trapezoid's error term, sum over union cells of
// s_R*sigma_S*h^3/6, because the extra (IR-only) nodes split union cells the reference does not split.
// Undoing the split at an IR-only node at distance h1 from the previous node the reference keeps (optical
// node or filter node) and h2 from the next node of the fine union grid adds  s*sigma*h1*h2*(h1+h2)/2.
template <typename real>
EGG_HD real unsplit_correction(real s_cell, real sigma, real h1, real h2) {
    return real(0.5) * s_cell * sigma * h1 * h2 * (h1 + h2);
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
