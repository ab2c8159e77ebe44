// phot_tables.cpp — see phot_tables.hpp.
#include "phot_tables.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

namespace eggphot {

namespace {

double trapz(const std::vector<double> &x, const std::vector<double> &y) {
    double s = 0;
    for (size_t i = 0; i + 1 < x.size(); ++i) s += 0.5 * (y[i] + y[i + 1]) * (x[i + 1] - x[i]);
    return s;
}

// [vif] lower_bound(v, x): index of the last element <= x, -1 if none.
int last_le(const std::vector<double> &v, double x) {
    return (int)(std::upper_bound(v.begin(), v.end(), x) - v.begin()) - 1;
}

// value at xq of the piece-wise linear curve (x, y), following merge_add's rules
// (src/egg-utils.hpp:21-60): own value on a node, interpolation strictly inside the span,
// nothing before the first or after the last node.
double curve_at(const std::vector<double> &x, const double *y, double xq) {
    const int n = (int)x.size();
    if (n == 0 || xq < x[0] || xq > x[n - 1]) return 0.0;
    int i = last_le(x, xq);
    if (x[i] == xq) return y[i];
    return y[i] + (y[i + 1] - y[i]) * (xq - x[i]) / (x[i + 1] - x[i]);
}

// average of the piece-wise linear curve over [a, b] ([vif] integrate(x, y, a, b)/(b-a), gencat:526-527)
double box_integral(const std::vector<double> &x, const std::vector<double> &y, double a, double b) {
    const int n = (int)x.size();
    a = std::max(a, x.front());
    b = std::min(b, x.back());
    if (!(b > a)) return 0.0;
    auto val = [&](double q) {
        int i = std::min(std::max(last_le(x, q), 0), n - 2);
        return y[i] + (y[i + 1] - y[i]) * (q - x[i]) / (x[i + 1] - x[i]);
    };
    double px = a, py = val(a), s = 0;
    for (int k = std::max(last_le(x, a), 0) + 1; k < n && x[k] < b; ++k) {
        s += 0.5 * (py + y[k]) * (x[k] - px);
        px = x[k];
        py = y[k];
    }
    s += 0.5 * (py + val(b)) * (b - px);
    return s;
}

void align16(std::vector<uint8_t> &b) { b.resize((b.size() + 15) & ~size_t(15), 0); }
size_t up16(size_t v) { return (v + 15) & ~size_t(15); }

// Double-precision band table, the source of both device layouts.
struct FilterCellD {
    double f, lam, k0, hr, s, dl;
};
struct BandTableD {
    BandHeader h{};
    std::vector<FilterCellD> cells; // n + 1 records: cells 0..n-2, the open cell n-1, the sentinel n
    std::vector<uint16_t> bkt;
};

Status build_band_table(const Filter &f, uint32_t out_col, int bucket_factor, BandTableD &t) {
    const int nf = (int)f.lam.size();
    t.h = BandHeader{};
    t.h.out_col = out_col;
    t.h.full_first = f.lam.front();
    t.h.full_last = f.lam.back();
    // drop the wings where the response is identically zero: they add exactly nothing to the
    // union-grid trapezoid (both factors of every node product vanish there)
    int a = -1, b = -1;
    for (int i = 0; i < nf; ++i)
        if (f.res[i] != 0.0) { if (a < 0) a = i; b = i; }
    if (a < 0) { // identically zero response
        t.h.n = 0;
        t.h.eff_first = t.h.eff_last = t.h.full_first;
        t.h.fc = t.h.full_first;
        t.cells.clear();
        t.bkt.assign(1, 0);
        t.h.nbkt = 1;
        return Status();
    }
    a = std::max(a - 1, 0);
    b = std::min(b + 1, nf - 1);
    if (b == a) { b = std::min(a + 1, nf - 1); a = b - 1; }
    const int n = b - a + 1;
    if (n > 65000) return Status::error(EGGPHOT_ERR_UNSUPPORTED, "filter with more than 65000 nodes");
    t.h.n = (uint32_t)n;
    t.h.eff_first = f.lam[a];
    t.h.eff_last = f.lam[b];
    t.h.fc = 0.5 * (f.lam[a] + f.lam[b]);
    t.cells.assign(n + 1, FilterCellD{});
    // K0 = cumulative trapezoid of R; Lambda = L - G/6 advances by dl (K0 + R dl/2) per cell (phot_core.cuh)
    long double k0 = 0, lam = 0;
    for (int i = 0; i < n; ++i) {
        FilterCellD &e = t.cells[i];
        e.f = f.lam[a + i] - t.h.fc;
        e.k0 = (double)k0;
        e.lam = (double)lam;
        if (i + 1 < n) {
            const double R = f.res[a + i], rn = f.res[a + i + 1];
            e.dl = f.lam[a + i + 1] - f.lam[a + i];
            e.s = e.dl > 0 ? (rn - R) / e.dl : 0.0;
            e.hr = 0.5 * R;
            lam += (long double)e.dl * (k0 + 0.5L * (long double)R * e.dl);
            k0 += 0.5L * ((long double)R + rn) * e.dl;
        } else { // open cell right of the last node
            e.hr = 0.0;
            e.s = 0.0;
            e.dl = 1e30;
        }
    }
    t.cells[n] = FilterCellD{1e300, 0, 0, 0, 0, 1e30};
    t.h.f_first = t.cells[0].f;
    t.h.f_last = t.cells[n - 1].f;
    t.h.tot0 = t.cells[n - 1].k0;
    t.h.asym0 = t.cells[n - 1].lam - t.h.tot0 * t.h.f_last;
    // two-sided gauge (phot_core.cuh): cells from the half-mass node on are re-expressed relative to the
    // right end, accumulated from that end itself (subtracting the asymptote would cancel in the far wing)
    int split = n - 1;
    {
        double a0 = 0, run = 0;
        for (int i = 0; i + 1 < n; ++i) a0 += 0.5 * (std::fabs(f.res[a + i]) + std::fabs(f.res[a + i + 1])) * t.cells[i].dl;
        for (int i = 0; i + 1 < n; ++i) {
            run += 0.5 * (std::fabs(f.res[a + i]) + std::fabs(f.res[a + i + 1])) * t.cells[i].dl;
            if (run >= 0.5 * a0) { split = i + 1; break; }
        }
    }
    {
        long double k0r = 0, lamr = 0;
        t.cells[n - 1].k0 = 0;
        t.cells[n - 1].lam = 0;
        for (int i = n - 2; i >= split; --i) {
            const double R = f.res[a + i], rn = f.res[a + i + 1], dl = t.cells[i].dl;
            k0r -= 0.5L * ((long double)R + rn) * dl;
            lamr -= (long double)dl * (k0r + 0.5L * (long double)R * dl);
            t.cells[i].k0 = (double)k0r;
            t.cells[i].lam = (double)lamr;
        }
    }
    t.h.split = (uint32_t)split;
    t.h.f_split = t.cells[split].f;
    // bucket table: bkt[k] = last node with f <= f_first + k*w.  When buckets narrower than every cell are
    // affordable, a bucket holds at most one node and one probe settles the lookup (bit 31 of `split`).
    const double span = t.h.f_last - t.h.f_first;
    double min_dl = span;
    for (int i = 0; i + 1 < n; ++i) min_dl = std::min(min_dl, t.cells[i].dl);
    int nbkt = std::min(std::max(bucket_factor * n, 1), 65000);
    if (min_dl > 0 && span > 0) {
        const double need = std::ceil(span / min_dl * 1.02) + 1;
        if (need <= 8.0 * n && need <= 65000.0) {
            nbkt = std::max((int)need, 1);
            t.h.split |= 0x80000000u;
        }
    }
    t.h.nbkt = (uint32_t)nbkt;
    t.h.bkt_inv_w = span > 0 ? (float)(nbkt / span) : 0.0f;
    if (!(t.h.bkt_inv_w * span < nbkt)) t.h.bkt_inv_w = std::nextafter(t.h.bkt_inv_w, 0.0f);
    t.h.l2first = (float)(std::log2(t.h.eff_first) * kLutPerOctave);
    t.h.l2last = (float)(std::log2(t.h.eff_last) * kLutPerOctave);
    t.h.l2split = (float)(std::log2(std::max(t.h.fc + t.h.f_split, 1e-300)) * kLutPerOctave);
    t.bkt.assign(nbkt, 0);
    int i = 0;
    for (int k = 0; k < nbkt; ++k) {
        // the kernel forms the bucket index as (int)((float)(t - f_first) * bkt_inv_w); its two float roundings
        // move the product by less than 2^-22 of itself, so an offset that yields index k is at least
        // k / bkt_inv_w * (1 - 2^-22): nodes up to that (slightly lowered) edge can never overshoot t
        const double edge = t.h.f_first + (t.h.bkt_inv_w > 0 ? (double)k / (double)t.h.bkt_inv_w * (1.0 - 2.4e-7) : 0.0);
        while (i + 1 < n && t.cells[i + 1].f <= edge) ++i;
        t.bkt[k] = (uint16_t)i;
    }
    return Status();
}


size_t band_bytes(const BandTableD &t, int precision) {
    const size_t n = t.h.n, rs = precision == EGGPHOT_FP64 ? 8 : 4;
    if (n == 0) return up16(t.bkt.size() * sizeof(uint16_t));
    const size_t sf = cell_slot((int)n) + 1, sc = cell_slot((int)n - 1) + 1; // padded slots (phot_core.cuh)
    return up16(sf * 8) + 2 * up16(sc * 8) + up16(sc * rs) + up16(sc * 2 * rs) + up16(t.bkt.size() * sizeof(uint16_t));
}

template <typename T, typename Get> uint32_t append_array(std::vector<uint8_t> &blob, size_t count, Get get) {
    align16(blob);
    const uint32_t off = (uint32_t)blob.size();
    for (size_t i = 0; i < count; ++i) {
        const T v = (T)get(i);
        const uint8_t *p = reinterpret_cast<const uint8_t *>(&v);
        blob.insert(blob.end(), p, p + sizeof(T));
    }
    return off;
}
// per-cell array of `width` values per cell, cell c stored in slot cell_slot(c)
template <typename T, typename Get> uint32_t append_cells(std::vector<uint8_t> &blob, size_t ncell, int width, Get get) {
    align16(blob);
    const uint32_t off = (uint32_t)blob.size();
    const size_t nslot = cell_slot((int)ncell - 1) + 1;
    blob.resize(blob.size() + nslot * width * sizeof(T), 0);
    T *dst = reinterpret_cast<T *>(blob.data() + off);
    for (size_t c = 0; c < ncell; ++c)
        for (int w = 0; w < width; ++w) dst[(size_t)cell_slot((int)c) * width + w] = (T)get(c, w);
    return off;
}

void append_band(std::vector<uint8_t> &blob, size_t header_pos, const BandTableD &t, int precision) {
    BandHeader h = t.h;
    const size_t n = t.h.n;
    const auto &c = t.cells;
    if (n > 0) {
        h.f_off = append_cells<double>(blob, n + 1, 1, [&](size_t i, int) { return c[i].f; });
        h.lam_off = append_cells<double>(blob, n, 2, [&](size_t i, int w) { return w ? c[i].k0 : c[i].lam; });
        h.k0_off = h.lam_off + 8; // interleaved with Lambda: {Lambda, K0} records
        if (precision == EGGPHOT_FP64) {
            h.hr_off = append_cells<double>(blob, n, 1, [&](size_t i, int) { return c[i].hr; });
            h.sdl_off = append_cells<double>(blob, n, 2, [&](size_t i, int w) { return w ? c[i].dl : c[i].s; });
        } else {
            h.hr_off = append_cells<float>(blob, n, 1, [&](size_t i, int) { return c[i].hr; });
            h.sdl_off = append_cells<float>(blob, n, 2, [&](size_t i, int w) { return w ? c[i].dl : c[i].s; });
        }
    }
    h.bkt_off = append_array<uint16_t>(blob, t.bkt.size(), [&](size_t i) { return t.bkt[i]; });
    align16(blob);
    std::memcpy(blob.data() + header_pos, &h, sizeof(h));
}

// Rest-frame weights of one band on one grid (flux_rf = sum_j S_j w_j): the union-node trapezoid of
// [vif] sed2flux walked directly in double at z = 0, applied to the hat function of every node.
RfWeights rf_weights(const Filter &f, const Grid &g, int nodes_mode) {
    RfWeights r;
    const int N = g.n(), nf = (int)f.lam.size();
    r.covered = N >= 2 && nf >= 2 && g.x.front() <= f.lam.front() && g.x.back() >= f.lam.back();
    if (!r.covered) return r;
    const int jlo = std::max(last_le(g.x, f.lam.front()), 0);
    int jhi = (int)(std::lower_bound(g.x.begin(), g.x.end(), f.lam.back()) - g.x.begin());
    jhi = std::min(jhi, N - 1);
    if (jhi <= jlo) return r;
    r.j0 = jlo;
    r.w.assign(jhi - jlo + 1, 0.0);
    // nodes of the trapezoid: filter nodes and the SED nodes strictly inside the filter span (merged); the two ends of
    // the filter and the SED nodes inside (sed); the filter nodes alone (filter)
    std::vector<double> u(f.lam);
    if (nodes_mode == EGGPHOT_NODES_SED) u = {f.lam.front(), f.lam.back()};
    for (int j = jlo; j <= jhi && nodes_mode != EGGPHOT_NODES_FILTER; ++j)
        if (g.x[j] > f.lam.front() && g.x[j] < f.lam.back()) u.push_back(g.x[j]);
    std::sort(u.begin(), u.end());
    u.erase(std::unique(u.begin(), u.end()), u.end());
    auto hat = [&](double q, int &j, double &al) { // S(q) = (1 - al) S_j + al S_{j+1}
        j = std::min(std::max(last_le(g.x, q), jlo), jhi - 1);
        al = (q - g.x[j]) / (g.x[j + 1] - g.x[j]);
    };
    for (size_t k = 0; k + 1 < u.size(); ++k) {
        const double h = u[k + 1] - u[k];
        // the filter cell that holds this union cell (a repeated filter wavelength is a jump of R: the
        // cell on its left ends with the first value, the cell on its right starts with the last)
        const int i = std::min(std::max(last_le(f.lam, 0.5 * (u[k] + u[k + 1])), 0), nf - 2);
        const double d = f.lam[i + 1] - f.lam[i], sl = d > 0 ? (f.res[i + 1] - f.res[i]) / d : 0.0;
        for (int e = 0; e < 2; ++e) {
            const double q = u[k + e];
            double rq = f.res[i] + sl * (q - f.lam[i]);
            if (nodes_mode == EGGPHOT_NODES_SED) {
                // a cell between SED nodes spans several filter cells: the response at the node itself
                const int iq = q >= f.lam.back() ? nf - 2 : std::min(std::max(last_le(f.lam, q), 0), nf - 2);
                const double dq = f.lam[iq + 1] - f.lam[iq];
                rq = q >= f.lam.back() ? f.res[nf - 1] : (dq > 0 ? f.res[iq] + (f.res[iq + 1] - f.res[iq]) / dq * (q - f.lam[iq]) : f.res[iq + 1]);
            }
            const double w = 0.5 * h * rq;
            int j; double al;
            hat(q, j, al);
            r.w[j - jlo] += w * (1.0 - al);
            r.w[j + 1 - jlo] += w * al;
        }
    }
    return r;
}

Status make_grid_thresholds(Grid &g, bool apply_igm, const char *what) {
    g.lz = g.l0 = g.l1 = g.l2 = 0;
    if (!apply_igm) return Status();
    const int lz = last_le(g.x, 0.0600);
    if (lz < 0)
        return Status::error(EGGPHOT_ERR_GRID, std::string("IGM absorption requested but the ") + what +
                             " wavelength grid has no node at or below 0.06 um (the reference would index past the SED, "
                             "src/egg-gencat.cpp:2123-2130); use igm=none");
    g.lz = lz;
    g.l0 = last_le(g.x, 0.0912);
    g.l1 = last_le(g.x, 0.1026);
    g.l2 = last_le(g.x, 0.1216);
    return Status();
}

// ---- tables of the lane kernel (phot_lane.cuh) --------------------------------------------------------------------
struct BandTable2 {
    BandHeader2 h{};
    std::vector<Rec> rec;
    std::vector<double> hs;
    std::vector<uint16_t> bkt;
};

// Table of the response (lam, res)[i0..i1] (a whole filter or a piece of one cut at filter nodes).
Status build_band_table2(const Filter &f, int i0, int i1, uint32_t out_col, int bucket_factor, BandTable2 &t) {
    t = BandTable2();
    t.h.out_col = out_col;
    t.h.full_first = f.lam.front();
    t.h.full_last = f.lam.back();
    t.h.f_first = t.h.f_last = f.lam[i0];
    for (int s = 0; s < kMaxSplit; ++s) t.h.msplit[s] = 0x7fffffffu;
    // drop the wings where the response is identically zero (they add exactly nothing to the union-grid trapezoid)
    int a = -1, b = -1;
    for (int i = i0; i <= i1; ++i)
        if (f.res[i] != 0.0) { if (a < 0) a = i; b = i; }
    if (a < 0) return Status(); // identically zero response: n = 0
    a = std::max(a - 1, i0);
    b = std::min(b + 1, i1);
    if (!(f.lam[b] > f.lam[a])) return Status(); // no width: nothing to integrate
    // nodes: absolute wavelengths (u = t - f is formed from the same two doubles as in the reference); a zero node at
    // an end where the response starts or stops abruptly.  (No other node may be added: every filter node is a node of
    // the union grid, and the trapezoid rule on it is not exact for the quadratic R S.)
    std::vector<double> fn, R;
    if (f.res[a] != 0.0) { fn.push_back(f.lam[a]); R.push_back(0.0); }
    for (int i = a; i <= b; ++i) { fn.push_back(f.lam[i]); R.push_back(f.res[i]); }
    if (f.res[b] != 0.0) { fn.push_back(f.lam[b]); R.push_back(0.0); }
    const int n = (int)fn.size();
    if (n > 64900) return Status::error(EGGPHOT_ERR_UNSUPPORTED, "filter piece with more than 64900 nodes");
    t.h.f_first = fn[0];
    t.h.f_last = fn[n - 1];
    std::vector<double> dl(n, 0.0), sl(n, 0.0);
    for (int i = 0; i + 1 < n; ++i) {
        dl[i] = fn[i + 1] - fn[i];
        sl[i] = dl[i] > 0 ? (R[i + 1] - R[i]) / dl[i] : 0.0;
    }
    // narrowest cell, the zero-width cell between an added end node and the first / last node of the response apart
    // (that one is what the second probe of the lookup is for); any other repeated wavelength makes the scan necessary
    const double span = fn[n - 1] - fn[0];
    double min_dl = span;
    for (int i = 0; i + 1 < n; ++i)
        if (fn[i + 1] != fn[i] || !((i == 0 && R[0] == 0.0) || (i + 2 == n && R[n - 1] == 0.0))) min_dl = std::min(min_dl, dl[i]);
    const bool can_probe = min_dl > 0;
    // ---- anchors: the two ends, and the deepest node of every interior valley of |R| (widest valleys first)
    std::vector<int> anchors;
    {
        double mx = 0;
        for (double r : R) mx = std::max(mx, std::fabs(r));
        const double thr = 1e-6 * mx;
        struct Valley { int node; double width; };
        std::vector<Valley> val;
        int first_big = -1, last_big = -1;
        for (int i = 0; i < n; ++i)
            if (std::fabs(R[i]) > thr) { if (first_big < 0) first_big = i; last_big = i; }
        for (int i = first_big + 1; i < last_big;) {
            if (std::fabs(R[i]) > thr) { ++i; continue; }
            int k = i;
            while (k + 1 < last_big && std::fabs(R[k + 1]) <= thr) ++k;
            // deepest node of the run [i, k]; the middle one among equals (an all-zero gap)
            double lo = std::fabs(R[i]);
            for (int q = i; q <= k; ++q) lo = std::min(lo, std::fabs(R[q]));
            int c0 = -1, c1 = -1;
            for (int q = i; q <= k; ++q)
                if (std::fabs(R[q]) == lo) { if (c0 < 0) c0 = q; c1 = q; }
            const int node = (c0 + c1) / 2;
            if (fn[node] > fn[0] && fn[node] < fn[n - 1]) val.push_back(Valley{node, fn[k] - fn[i]});
            i = k + 1;
        }
        // a valley matters only if the response carries weight on both sides of it: a node in it differences table values of
        // the size of the lighter side's integral (the split between two anchors sits at the half-mass node), which the
        // zero runs in the noisy wings of measured curves never do
        {
            std::vector<double> cum(n, 0.0);
            for (int i = 0; i + 1 < n; ++i) cum[i + 1] = cum[i] + 0.5 * (std::fabs(R[i]) + std::fabs(R[i + 1])) * dl[i];
            std::vector<Valley> keepv;
            for (const Valley &v : val)
                if (std::min(cum[v.node], cum[n - 1] - cum[v.node]) > 1e-4 * cum[n - 1]) keepv.push_back(v);
            val.swap(keepv);
        }
        std::stable_sort(val.begin(), val.end(), [](const Valley &x, const Valley &y) { return x.width > y.width; });
        if ((int)val.size() > kMaxSplit - 1) val.resize(kMaxSplit - 1);
        anchors.push_back(0);
        for (const Valley &v : val) anchors.push_back(v.node);
        anchors.push_back(n - 1);
        std::sort(anchors.begin(), anchors.end());
        anchors.erase(std::unique(anchors.begin(), anchors.end()), anchors.end());
    }
    // ---- split nodes: the half-mass node (of |R|) between consecutive anchors, moved if need be so that its wavelength
    //      lies strictly inside the span and no other node shares it (a split is two records at one wavelength: with an
    //      abrupt end or a jump of the response there the two-probe lookup would not reach the cell)
    std::vector<int> msp;
    bool force_scan = false;
    {
        auto lone = [&](int i) {
            if (!(fn[i] > fn[0] && fn[i] < fn[n - 1])) return false;
            return !can_probe || ((i == 0 || fn[i - 1] != fn[i]) && (i == n - 1 || fn[i + 1] != fn[i])); // the scan passes any run
        };
        std::vector<int> keep{anchors[0]};
        for (size_t s = 0; s + 1 < anchors.size(); ++s) {
            const int a0 = keep.back(), a1 = anchors[s + 1];
            double tot = 0, run = 0;
            for (int i = a0; i < a1; ++i) tot += 0.5 * (std::fabs(R[i]) + std::fabs(R[i + 1])) * dl[i];
            int m = a1;
            for (int i = a0; i < a1; ++i) {
                run += 0.5 * (std::fabs(R[i]) + std::fabs(R[i + 1])) * dl[i];
                if (run >= 0.5 * tot) { m = i + 1; break; }
            }
            int best = -1; // the lone node of (a0, a1] nearest to m
            for (int i = a0 + 1; i <= a1; ++i)
                if (lone(i) && (best < 0 || std::abs(i - m) < std::abs(best - m))) best = i;
            if (best < 0) continue; // no usable split between these anchors: a1 joins the region of a0
            msp.push_back(best);
            keep.push_back(a1);
        }
        anchors = keep;
        if (msp.empty()) {
            // no node strictly inside the span on its own wavelength (a two-node top-hat): split at the half-mass node
            // wherever it is and let the scan pass the records that share its wavelength
            double tot = 0, run = 0;
            for (int i = 0; i + 1 < n; ++i) tot += 0.5 * (std::fabs(R[i]) + std::fabs(R[i + 1])) * dl[i];
            int m = n - 2;
            for (int i = 0; i + 1 < n; ++i) {
                run += 0.5 * (std::fabs(R[i]) + std::fabs(R[i + 1])) * dl[i];
                if (run >= 0.5 * tot) { m = i + 1; break; }
            }
            msp.push_back(std::min(std::max(m, 1), n - 2));
            anchors = {0, n - 1};
            force_scan = true;
        }
    }
    const int nsplit = (int)msp.size();
    t.h.nsplit = (uint32_t)nsplit;
    // ---- records: [left sentinel][nodes; region s holds nodes msp[s-1]..msp[s], (Lambda, K0) = (0, 0) at its anchor,
    //      accumulated outwards in long double: K0 = running trapezoid of R, Lambda = L - G/6 advances by
    //      dl (K0 + R dl / 2) per cell (phot_core.cuh)][terminator]
    const int nrec = 1 + n + nsplit + 1;
    t.h.n = (uint32_t)nrec;
    t.rec.assign(nrec + 1, Rec{0, 0, 0, 0}); // (a second terminator: the lookup reads both probes at once)
    t.hs.assign(nrec + 1, 0.0);
    t.rec[0] = Rec{fn[0] - std::max(fn[n - 1] - fn[0], 1e-6), 0, 0, 0};
    t.rec[nrec - 1] = Rec{1e300, 0, 0, 0};
    t.rec[nrec] = Rec{1e300, 0, 0, 0};
    std::vector<long double> K0(n), L(n);
    std::vector<long double> Kat(nsplit), Lat(nsplit); // region s at its right end (split node msp[s])
    for (int s = 0; s <= nsplit; ++s) {
        const int lo = s == 0 ? 0 : msp[s - 1], hi = s == nsplit ? n - 1 : msp[s], an = anchors[s];
        K0[an] = 0; L[an] = 0;
        for (int i = an; i < hi; ++i) {
            L[i + 1] = L[i] + (long double)dl[i] * (K0[i] + 0.5L * (long double)R[i] * dl[i]);
            K0[i + 1] = K0[i] + 0.5L * ((long double)R[i] + R[i + 1]) * dl[i];
        }
        for (int i = an - 1; i >= lo; --i) {
            K0[i] = K0[i + 1] - 0.5L * ((long double)R[i] + R[i + 1]) * dl[i];
            L[i] = L[i + 1] - (long double)dl[i] * (K0[i] + 0.5L * (long double)R[i] * dl[i]);
        }
        for (int i = lo; i <= hi; ++i) {
            t.rec[1 + i + s] = Rec{fn[i], (double)L[i], (double)K0[i], 0.5 * R[i]};
            if (i < hi) t.hs[1 + i + s] = 0.5 * sl[i]; // the last record of a region opens no cell of it
        }
        if (s > 0) { // gauge jump across split s-1: Lambda_{s-1}(t) - Lambda_s(t) = C + D (t - F)
            t.h.splitC[s - 1] = (double)(Lat[s - 1] - L[lo]);
            t.h.splitD[s - 1] = (double)(Kat[s - 1] - K0[lo]);
            t.h.splitF[s - 1] = fn[lo];
        }
        if (s < nsplit) { Kat[s] = K0[hi]; Lat[s] = L[hi]; t.h.msplit[s] = (uint32_t)(1 + hi + s); }
    }
    // ---- bucket table: entry 0 -> the left sentinel, entries 1..nb -> bkt = last record whose node is at or below the
    //      start of the bucket, the start lowered by 1e-6 of a bucket (the kernel's index is the rounding of a double
    //      product: it can exceed the exact one only for a t within ~1e-10 of a bucket of the boundary), entry nb + 1 ->
    //      the last node (everything red of the span).  When buckets narrower than every cell are affordable two probes
    //      settle the lookup (two, because a split node or an abrupt end is two records at one wavelength).
    int nb = std::min(std::max(bucket_factor * n, 1), 64000);
    if (can_probe && !force_scan) {
        const double need = std::ceil(span / min_dl * 1.02) + 1;
        if (need <= 8.0 * n && need <= 64000.0) {
            nb = std::max((int)need, 1);
            t.h.two_probe = 1;
        }
    }
    t.h.nbkt = (uint32_t)(nb + 2);
    t.h.bk_scale = nb / span;
    t.h.bk_c0 = -fn[0] * t.h.bk_scale - 0.5 + 1.0; // entry = bucket + 1
    t.bkt.assign(nb + 2, 0);
    const int last_node = nrec - 2;
    int r = 0; // the left sentinel: bucket 0 starts (just) blue of the span
    for (int k = 0; k < nb; ++k) {
        const double edge = fn[0] + ((double)k - 1e-6) / t.h.bk_scale;
        while (t.rec[r + 1].f <= edge) ++r; // edge < f_last: never reaches the last node, which opens the cell red of the span
        t.bkt[k + 1] = (uint16_t)r;
    }
    t.bkt[0] = 0;
    t.bkt[nb + 1] = (uint16_t)last_node;
    if (nrec > (int)kBktRecMask) return Status::error(EGGPHOT_ERR_UNSUPPORTED, "a band table with more than 32767 records");
    if (t.h.two_probe) {
        // one probe settles the lookup when the entry also says how far to step: two records where the next node is stored
        // twice (a split node, an abrupt end: two records at one wavelength), one otherwise.  A bucket that starts in cell r
        // ends before the node after next (it is narrower than every cell of non-zero width).
        for (int k = 0; k < nb + 2; ++k) {
            const int rr = t.bkt[k];
            if (rr + 2 <= nrec && t.rec[rr + 1].f == t.rec[rr + 2].f) t.bkt[k] = (uint16_t)(rr | kBktStep2);
        }
    }
    return Status();
}

size_t band_bytes2(const BandTable2 &t) {
    return up16(t.rec.size() * sizeof(Rec)) + up16(t.hs.size() * sizeof(double)) + up16(t.bkt.size() * sizeof(uint16_t));
}

template <typename T> uint32_t append_raw(std::vector<uint8_t> &blob, const std::vector<T> &v, size_t align) {
    blob.resize((blob.size() + align - 1) / align * align, 0);
    const uint32_t off = (uint32_t)blob.size();
    const uint8_t *p = reinterpret_cast<const uint8_t *>(v.data());
    blob.insert(blob.end(), p, p + v.size() * sizeof(T));
    return off;
}

void append_band2(std::vector<uint8_t> &blob, size_t header_pos, const BandTable2 &t) {
    BandHeader2 h = t.h;
    h.rec_off = append_raw(blob, t.rec, 32);
    h.hs_off = append_raw(blob, t.hs, 16);
    h.bkt_off = append_raw(blob, t.bkt, 16);
    align16(blob);
    std::memcpy(blob.data() + header_pos, &h, sizeof(h));
}

// first-fit-decreasing packing of per-band byte sizes into as few groups as the budget allows, the groups kept even
std::vector<std::vector<size_t>> pack_groups(const std::vector<size_t> &size, size_t budget) {
    std::vector<size_t> bysize(size.size());
    size_t total = 0;
    for (size_t b = 0; b < size.size(); ++b) { bysize[b] = b; total += size[b]; }
    std::stable_sort(bysize.begin(), bysize.end(), [&](size_t i, size_t j) { return size[i] > size[j]; });
    std::vector<std::vector<size_t>> bins;
    for (size_t ng = std::max<size_t>((total + budget - 1) / budget, 1); !size.empty(); ++ng) {
        bins.assign(ng, {});
        std::vector<size_t> fill(ng, 0);
        bool ok = true;
        for (size_t b : bysize) {
            size_t best = ng;
            for (size_t g = 0; g < ng; ++g)
                if (fill[g] + size[b] <= budget && (best == ng || fill[g] < fill[best])) best = g;
            if (best == ng) {
                if (size[b] > budget && ng >= size.size()) { best = 0; for (size_t g = 1; g < ng; ++g) if (fill[g] < fill[best]) best = g; }
                else { ok = false; break; }
            }
            bins[best].push_back(b);
            fill[best] += size[b];
        }
        if (ok) break;
    }
    return bins;
}

Status build_lane_tables(Prepared &P, const eggphot_opts &opts) {
    // ---- the disk grid as the walk reads it
    const Grid &gd = P.grid_disk, &gb = P.grid_bulge;
    const int nD = gd.n(), nB = gb.n();
    const int nG = lane_grid_rows(nD); // nD real nodes, two virtual ones, padding to the walk's strips of 8 arrivals
    P.grid_nodes.assign((size_t)nG, GridNode{0, 0, 0, -1, -1});
    P.next_opt.assign((size_t)nG, nD);
    P.opt_node.assign((size_t)nB, 0);
    {
        int o = 0, lasto = -1;
        for (int j = 0; j < nD; ++j) {
            GridNode &g = P.grid_nodes[j];
            g.x = gd.x[j];
            g.idxp = j > 0 ? 1.0 / (gd.x[j] - gd.x[j - 1]) : 0.0;
            g.po = lasto;
            if (P.has_bulge) {
                while (o < nB && gb.x[o] < gd.x[j]) ++o;
                if (o < nB && gb.x[o] == gd.x[j]) {
                    g.o = o;
                    g.idxpo = lasto >= 0 ? 1.0 / (gd.x[j] - gd.x[lasto]) : 0.0;
                    P.opt_node[o] = j;
                    lasto = j;
                }
            }
        }
        // the virtual node: far enough to the red that it is right of every filter the last real node is right of
        GridNode &v = P.grid_nodes[nD];
        const double xl = nD ? gd.x[nD - 1] : 1.0;
        v.x = xl * 2.0 + 1.0;
        v.idxp = 1.0 / (v.x - xl);
        v.po = lasto;
        v.o = P.has_bulge ? nB : -1;
        v.idxpo = lasto >= 0 ? 1.0 / (v.x - gd.x[lasto]) : 0.0;
        // a second virtual node: the unrolled walk of the kernel may arrive one node past the end of a block
        GridNode &v2 = P.grid_nodes[nD + 1];
        v2 = v;
        v2.x = v.x * 2.0;
        v2.idxp = 1.0 / (v2.x - v.x);
        v2.po = nD;
        v2.o = P.has_bulge ? nB + 1 : -1;
        v2.idxpo = 1.0 / (v2.x - v.x);
        int nxt = nD;
        for (int j = nD; j >= 0; --j) {
            P.next_opt[j] = nxt;
            if (j < nD && P.grid_nodes[j].o >= 0) nxt = j;
        }
        P.next_opt[nD] = nD;
        P.next_opt[nD + 1] = nD + 1;
        // padding nodes: further and further to the red, never optical (the walk works in strips of 8 arrivals cut at
        // multiples of 8; nodes beyond a window weigh exactly nothing)
        for (int j = nD + 2; j < nG; ++j) {
            GridNode &w = P.grid_nodes[j];
            w.x = P.grid_nodes[j - 1].x * 2.0;
            w.idxp = 1.0 / (w.x - P.grid_nodes[j - 1].x);
            w.idxpo = 0.0;
            w.o = -1;
            w.po = nD + 1;
            P.next_opt[j] = j;
        }
        if (P.has_bulge)
            for (int o2 = 0; o2 < nB; ++o2)
                if (P.grid_nodes[P.opt_node[o2]].o != o2)
                    return Status::error(EGGPHOT_ERR_GRID, "an optical node is missing from the disk grid");
    }
    // ---- the optical grid as the bulge's walk reads it: its own nodes, two virtual ones, padding
    {
        const int nGB = lane_grid_rows(nB);
        P.grid_nodes_b.assign((size_t)nGB, GridNode{0, 0, 0, -1, -1});
        for (int o = 0; o < nGB; ++o) {
            GridNode &g = P.grid_nodes_b[o];
            if (o < nB) g.x = gb.x[o];
            else g.x = (o == nB ? (nB ? gb.x[nB - 1] : 1.0) * 2.0 + 1.0 : P.grid_nodes_b[o - 1].x * 2.0);
            g.idxp = o > 0 ? 1.0 / (g.x - P.grid_nodes_b[o - 1].x) : 0.0;
        }
    }
    if (P.nodes_mode != EGGPHOT_NODES_MERGED) {
        // node sets `sed` and `filter` (conventions to try against a real catalog): the filters go to the
        // device as they are and a plain two-cursor kernel walks them; one header per band (no records) gives the plan
        // kernel the spans it needs for the node windows
        if (P.bands.size() > 64) return Status::error(EGGPHOT_ERR_UNSUPPORTED, "more than 64 bands with sed2flux_nodes=sed|filter");
        BandGroup g;
        g.blob.assign(P.bands.size() * sizeof(BandHeader2), 0);
        P.alt_off.assign(1, 0u);
        for (size_t b = 0; b < P.bands.size(); ++b) {
            const Filter &f = P.bands[b];
            BandHeader2 h{};
            h.full_first = h.f_first = f.lam.front();
            h.full_last = h.f_last = f.lam.back();
            h.n = (uint32_t)f.lam.size();
            h.out_col = (uint32_t)b;
            h.cost = 1.0f;
            for (int s = 0; s < kMaxSplit; ++s) h.msplit[s] = 0x7fffffffu;
            std::memcpy(g.blob.data() + b * sizeof(BandHeader2), &h, sizeof(h));
            g.cols.push_back((uint32_t)b);
            P.alt_fl.insert(P.alt_fl.end(), f.lam.begin(), f.lam.end());
            P.alt_fr.insert(P.alt_fr.end(), f.res.begin(), f.res.end());
            P.alt_off.push_back((uint32_t)P.alt_fl.size());
        }
        if (!P.bands.empty()) P.groups2.push_back(std::move(g));
        return Status();
    }
    // ---- band tables in groups
    size_t budget = opts.smem_table_budget;
    if (!budget) {
        const bool fp64 = opts.precision == EGGPHOT_FP64;
        budget = kSmemPerCta - kLaneFixedSmem - 256 - lane_stage_bytes(fp64) -
                 (lane_gx_in_smem(fp64, nD, nB, kSmemPerCta) ? lane_gx_bytes(nD, nB) : 0); // the grid's {x, idxp} live behind the tables
    }
    // a filter whose table exceeds the budget is cut at filter nodes into pieces that feed the same output column
    std::vector<BandTable2> tabs;
    std::vector<size_t> size;
    for (size_t b = 0; b < P.bands.size(); ++b) {
        const Filter &f = P.bands[b];
        const int nf = (int)f.lam.size();
        for (int pieces = 1;; ++pieces) {
            if (pieces > 64) return Status::error(EGGPHOT_ERR_UNSUPPORTED, "filter " + std::to_string(P.band_order[b]) +
                                                  " has too many nodes for the shared-memory band tables");
            std::vector<BandTable2> cut(pieces);
            bool fits = true;
            for (int q = 0; q < pieces && fits; ++q) {
                const int i0 = (int)((long)(nf - 1) * q / pieces), i1 = (int)((long)(nf - 1) * (q + 1) / pieces);
                if (i1 <= i0) { fits = false; break; }
                Status st = build_band_table2(f, i0, i1, (uint32_t)b, 2, cut[q]);
                if (!st.ok()) return st;
                fits = band_bytes2(cut[q]) + sizeof(BandHeader2) + 64 <= budget;
            }
            if (!fits && nf - 1 >= 2 * (pieces + 1)) continue;
            if (!fits) return Status::error(EGGPHOT_ERR_UNSUPPORTED, "filter " + std::to_string(P.band_order[b]) +
                                            " does not fit the shared-memory band tables; raise smem_table_budget");
            for (auto &c : cut) {
                if (c.h.n == 0 && pieces > 1) continue; // a piece without response
                const int lo = last_le(gd.x, c.h.f_first / 2.0), hi = last_le(gd.x, c.h.f_last / 2.0);
                c.h.cost = (float)std::max(hi - lo, 1);
                size.push_back(band_bytes2(c) + sizeof(BandHeader2) + 64);
                tabs.push_back(std::move(c));
            }
            break;
        }
    }
    for (auto &bin : pack_groups(size, budget)) {
        if (bin.empty()) continue;
        std::stable_sort(bin.begin(), bin.end(), [&](size_t i, size_t j) { return tabs[i].h.cost > tabs[j].h.cost; });
        BandGroup g;
        g.blob.assign(bin.size() * sizeof(BandHeader2), 0);
        for (size_t k = 0; k < bin.size(); ++k) {
            g.cols.push_back(tabs[bin[k]].h.out_col);
            append_band2(g.blob, k * sizeof(BandHeader2), tabs[bin[k]]);
        }
        P.groups2.push_back(std::move(g));
    }
    return Status();
}

} // namespace

void Grid::build_lut() {
    lut.clear();
    idx.assign(x.size(), 0.0);
    for (size_t j = 0; j + 1 < x.size(); ++j) idx[j] = 1.0 / (x[j + 1] - x[j]);
    if (x.empty()) return;
    lut_e0 = (int)std::floor(std::log2(x.front())) - 1;
    const int nk = (int)std::ceil((std::log2(x.back()) - lut_e0) * kLutPerOctave) + 2;
    lut.resize(nk);
    int j = -1;
    for (int k = 0; k < nk; ++k) {
        const double edge = std::exp2(lut_e0 + (double)k / kLutPerOctave);
        while (j + 1 < n() && x[j + 1] <= edge) ++j;
        lut[k] = (uint16_t)std::max(j, 0);
    }
}

Status prepare_filters(const eggphot_filters *raw, const eggphot_opts &opts, std::vector<Filter> &sorted,
                       std::vector<uint32_t> &order, std::vector<float> &lambda) {
    sorted.clear(); order.clear(); lambda.clear();
    if (!raw || raw->nband == 0) return Status();
    if (!raw->npts || !raw->lam || !raw->res) return Status::error(EGGPHOT_ERR_ARG, "null filter arrays");
    const bool adjust = opts.trim_filters || opts.filter_flambda || opts.filter_photons; // gencat:361
    std::vector<Filter> fil(raw->nband);
    for (uint32_t b = 0; b < raw->nband; ++b) {
        const uint32_t n = raw->npts[b];
        if (n < 2 || !raw->lam[b] || !raw->res[b])
            return Status::error(EGGPHOT_ERR_FILTER, "filter " + std::to_string(b) + " is empty");
        Filter &f = fil[b];
        f.lam.assign(raw->lam[b], raw->lam[b] + n);
        f.res.assign(raw->res[b], raw->res[b] + n);
        for (uint32_t i = 0; i + 1 < n; ++i)
            if (!(f.lam[i + 1] >= f.lam[i]))
                return Status::error(EGGPHOT_ERR_FILTER, "filter " + std::to_string(b) +
                                     ": wavelengths are not sorted in increasing order");
        if (adjust) {
            if (opts.trim_filters) { // gencat:334-340
                const double mx = *std::max_element(f.res.begin(), f.res.end());
                int a = -1, e = -1;
                for (uint32_t i = 0; i < n; ++i)
                    if (f.res[i] / mx > 1e-3) { if (a < 0) a = (int)i; e = (int)i; }
                if (a < 0) return Status::error(EGGPHOT_ERR_FILTER, "filter has no usable data");
                f.lam = std::vector<double>(f.lam.begin() + a, f.lam.begin() + e + 1);
                f.res = std::vector<double>(f.res.begin() + a, f.res.begin() + e + 1);
                if (f.lam.size() < 2) return Status::error(EGGPHOT_ERR_FILTER, "filter has no usable data");
            }
            if (opts.filter_flambda) for (size_t i = 0; i < f.lam.size(); ++i) f.res[i] /= f.lam[i] * f.lam[i];
            if (opts.filter_photons) for (size_t i = 0; i < f.lam.size(); ++i) f.res[i] *= f.lam[i];
            std::vector<double> lr(f.lam.size());
            for (size_t i = 0; i < lr.size(); ++i) lr[i] = f.lam[i] * f.res[i];
            const double norm = trapz(f.lam, f.res);
            f.rlam = trapz(f.lam, lr) / norm;
            if (opts.filter_normalize) for (double &r : f.res) r /= norm;
        } else {
            std::vector<double> lr(n);
            for (uint32_t i = 0; i < n; ++i) lr[i] = f.lam[i] * f.res[i];
            f.rlam = trapz(f.lam, lr); // [vif] get_filter: raw response, no renormalisation (KAT-1)
        }
    }
    // sort by the f32 reference wavelength (vec1f lambda, gencat:367-372)
    order.resize(raw->nband);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t i, uint32_t j) { return (float)fil[i].rlam < (float)fil[j].rlam; });
    for (uint32_t k : order) {
        sorted.push_back(std::move(fil[k]));
        lambda.push_back((float)sorted.back().rlam);
    }
    return Status();
}

Status prepare_opt_lib(const eggphot_libs &L, const eggphot_opts &opts, uint32_t &nopt,
                       std::vector<float> &lam, std::vector<float> &sed) {
    const size_t nsed = (size_t)L.nuv * L.nvj, nin = L.nopt;
    if (!L.opt_lam || !L.opt_sed || !L.opt_use || nsed == 0 || nin < 2)
        return Status::error(EGGPHOT_ERR_ARG, "optical library arrays are missing");
    const std::string imf = opts.opt_lib_imf ? opts.opt_lib_imf : "salpeter";
    double fimf = 1.0;
    if (imf == "chabrier" || imf == "kroupa") fimf = std::pow(10.0, -0.2); // gencat:461-463
    else if (imf != "salpeter") return Status::error(EGGPHOT_ERR_LIBRARY, "unknown IMF '" + imf + "'");
    size_t first = nsed;
    for (size_t s = 0; s < nsed; ++s) if (L.opt_use[s]) { first = s; break; }
    if (first == nsed) // gencat:537-541
        return Status::error(EGGPHOT_ERR_LIBRARY, "the provided optical SED library does not contain any valid SED");
    std::vector<float> fs(L.opt_sed, L.opt_sed + nsed * nin);
    if (fimf != 1.0) for (float &v : fs) v = (float)(v * fimf);

    std::vector<double> olam(nin);
    for (size_t l = 0; l < nin; ++l) olam[l] = L.opt_lam[first * nin + l];
    long id1 = -1, id2 = -1;
    const double step = opts.opt_lib_min_step;
    if (step > 0 && nin >= 3) { // gencat:478-488
        for (size_t l = 0; l < nin; ++l) {
            const double d = l == 0 ? olam[1] - olam[0]
                           : l == nin - 1 ? olam[nin - 1] - olam[nin - 2]
                           : 0.5 * (olam[l + 1] - olam[l - 1]);
            if (d < step) { if (id1 < 0) id1 = (long)l; id2 = (long)l; }
        }
    }
    if (id1 < 0) {
        nopt = (uint32_t)nin;
        lam.assign(L.opt_lam, L.opt_lam + nsed * nin);
        sed = std::move(fs);
        return Status();
    }
    // gencat:491-508
    std::vector<double> nl(olam.begin(), olam.begin() + id1);
    long nnew = (long)std::floor((olam[id2] - olam[id1]) / step) - 1;
    if (nnew < 0) nnew = 0;
    for (long k = 0; k < nnew; ++k) nl.push_back(olam[id1] + 0.5 * step + step * (double)k);
    nl.insert(nl.end(), olam.begin() + id2 + 1, olam.end());
    const size_t nout = nl.size();
    nopt = (uint32_t)nout;
    lam.assign(nsed * nout, 0.0f);
    sed.assign(nsed * nout, 0.0f);
    std::vector<double> row(nin);
    for (size_t s = 0; s < nsed; ++s) {
        if (!L.opt_use[s]) continue; // gencat:515-519
        for (size_t l = 0; l < nin; ++l) row[l] = fs[s * nin + l];
        for (size_t l = 0; l < nout; ++l) lam[s * nout + l] = (float)nl[l];
        for (long l = 0; l < id1; ++l) sed[s * nout + l] = (float)row[l];
        for (long k = 0; k < nnew; ++k) {
            const double c = nl[id1 + k];
            sed[s * nout + id1 + k] = (float)(box_integral(olam, row, c - 0.5 * step, c + 0.5 * step) / step);
        }
        for (size_t l = id2 + 1; l < nin; ++l) sed[s * nout + id1 + nnew + (l - id2 - 1)] = (float)row[l];
    }
    return Status();
}

Status prepare(const eggphot_libs &L, const eggphot_filters *bands, const eggphot_filters *rfbands,
               const eggphot_opts &opts, Prepared &P) {
    P = Prepared();
    if (opts.precision != EGGPHOT_FP32C && opts.precision != EGGPHOT_FP64)
        return Status::error(EGGPHOT_ERR_ARG, "precision must be EGGPHOT_FP32C or EGGPHOT_FP64");
    P.precision = opts.precision;
    const std::string igm = opts.igm ? opts.igm : "inoue14";
    if (igm != "none" && igm != "constant" && igm != "madau95" && igm != "inoue14") // gencat:286-289
        return Status::error(EGGPHOT_ERR_ARG, "unknown IGM prescription '" + igm +
                             "'; possible values are: 'none', 'constant', 'madau95', or 'inoue14'");
    P.apply_igm = igm != "none" && igm != "constant"; // gencat:2122
    P.no_stellar = opts.no_stellar;
    P.no_dust = opts.no_dust;
    P.no_nebular = opts.no_nebular || opts.nline == 0;
    P.line_profile = opts.line_profile;
    if (opts.no_stellar && opts.no_dust) // gencat:247-249 turns the whole stage off
        return Status::error(EGGPHOT_ERR_ARG, "no_stellar and no_dust together disable the flux stage (no_flux)");
    if (opts.sed2flux_nodes != EGGPHOT_NODES_MERGED && opts.sed2flux_nodes != EGGPHOT_NODES_SED && opts.sed2flux_nodes != EGGPHOT_NODES_FILTER)
        return Status::error(EGGPHOT_ERR_ARG, "sed2flux_nodes must be merged, sed or filter");
    P.nodes_mode = opts.sed2flux_nodes;
    if (opts.nline && opts.line_lambda) P.line_lambda.assign(opts.line_lambda, opts.line_lambda + opts.nline);
    if (P.line_lambda.size() > 64) return Status::error(EGGPHOT_ERR_UNSUPPORTED, "more than 64 emission lines");

    Status st = prepare_filters(bands, opts, P.bands, P.band_order, P.band_lambda);
    if (!st.ok()) return st;
    st = prepare_filters(rfbands, opts, P.rfbands, P.rfband_order, P.rfband_lambda);
    if (!st.ok()) return st;

    // ---- libraries
    P.has_bulge = !opts.no_stellar; // gencat:2206
    P.disk_mode = opts.no_dust ? DISK_OPT : (opts.no_stellar ? DISK_IR : DISK_BOTH); // gencat:2104-2119
    std::vector<double> xo, xi;
    if (!opts.no_stellar) {
        st = prepare_opt_lib(L, opts, P.nopt, P.opt_lam, P.opt_sed);
        if (!st.ok()) return st;
        P.nopt_sed = L.nuv * L.nvj;
        P.opt_use.assign(L.opt_use, L.opt_use + P.nopt_sed);
        // all used templates must share one wavelength grid (the reference's own re-binning assumes
        // it, gencat:475-476); a per-template grid would make the merged grid galaxy-dependent
        const float *ref = nullptr;
        for (uint32_t s = 0; s < P.nopt_sed; ++s) {
            if (!P.opt_use[s]) continue;
            const float *row = P.opt_lam.data() + (size_t)s * P.nopt;
            if (!ref) ref = row;
            else if (std::memcmp(ref, row, P.nopt * sizeof(float)) != 0)
                return Status::error(EGGPHOT_ERR_GRID, "optical templates do not share one wavelength grid");
        }
        xo.assign(ref, ref + P.nopt);
        for (size_t i = 0; i + 1 < xo.size(); ++i)
            if (!(xo[i + 1] > xo[i])) return Status::error(EGGPHOT_ERR_GRID, "optical wavelength grid is not strictly increasing");
    }
    if (!opts.no_dust) {
        if (!L.ir_lam || !L.ir_sed || L.nir_sed == 0 || L.nir < 2)
            return Status::error(EGGPHOT_ERR_LIBRARY, "missing LAM and/or SED columns in the IR library"); // gencat:438-441
        if (L.ir_kind == EGGPHOT_IR_CS17 && !L.ir_pah)
            return Status::error(EGGPHOT_ERR_LIBRARY, "cs17 IR library needs the PAH column");
        P.ir_kind = L.ir_kind;
        P.nir_sed = L.nir_sed;
        P.nir = L.nir;
        for (uint32_t s = 1; s < L.nir_sed; ++s)
            if (std::memcmp(L.ir_lam, L.ir_lam + (size_t)s * L.nir, L.nir * sizeof(double)) != 0)
                return Status::error(EGGPHOT_ERR_GRID, "IR templates do not share one wavelength grid");
        xi.assign(L.ir_lam, L.ir_lam + L.nir);
        for (size_t i = 0; i + 1 < xi.size(); ++i)
            if (!(xi[i + 1] > xi[i])) return Status::error(EGGPHOT_ERR_GRID, "IR wavelength grid is not strictly increasing");
    }

    // ---- grids
    if (P.has_bulge) {
        P.grid_bulge.x = xo;
        st = make_grid_thresholds(P.grid_bulge, P.apply_igm, "optical");
        if (!st.ok()) return st;
    }
    if (P.disk_mode == DISK_OPT) P.grid_disk.x = xo;
    else if (P.disk_mode == DISK_IR) P.grid_disk.x = xi;
    else { // union grid of merge_add: equal abscissae collapse into one node (egg-utils.hpp:55-60)
        P.grid_disk.x.resize(xo.size() + xi.size());
        auto e = std::set_union(xo.begin(), xo.end(), xi.begin(), xi.end(), P.grid_disk.x.begin());
        P.grid_disk.x.resize(e - P.grid_disk.x.begin());
    }
    st = make_grid_thresholds(P.grid_disk, P.apply_igm, P.disk_mode == DISK_IR ? "IR" : "disk");
    if (!st.ok()) return st;
    if (P.grid_disk.n() > 65000 || P.grid_bulge.n() > 65000)
        return Status::error(EGGPHOT_ERR_UNSUPPORTED, "SED wavelength grid with more than 65000 nodes");
    P.grid_bulge.build_lut();
    P.grid_disk.build_lut();

    // ---- template rows  Y = x * L(x)
    auto resample = [&](const std::vector<double> &xs, const double *ys, const Grid &g, double *dst) {
        for (int j = 0; j < g.n(); ++j) dst[j] = g.x[j] * curve_at(xs, ys, g.x[j]);
    };
    std::vector<double> tmp;
    if (!opts.no_stellar) {
        const int nb = P.grid_bulge.n();
        P.rows_bulge_opt.assign((size_t)P.nopt_sed * nb, 0.0);
        if (P.disk_mode != DISK_IR) P.rows_disk_opt.assign((size_t)P.nopt_sed * P.grid_disk.n(), 0.0);
        tmp.resize(P.nopt);
        for (uint32_t s = 0; s < P.nopt_sed; ++s) {
            if (!P.opt_use[s]) continue;
            for (uint32_t l = 0; l < P.nopt; ++l) tmp[l] = P.opt_sed[(size_t)s * P.nopt + l];
            for (int j = 0; j < nb; ++j) P.rows_bulge_opt[(size_t)s * nb + j] = xo[j] * tmp[j];
            if (P.disk_mode == DISK_OPT)
                std::copy_n(&P.rows_bulge_opt[(size_t)s * nb], nb, &P.rows_disk_opt[(size_t)s * nb]);
            else if (P.disk_mode == DISK_BOTH)
                resample(xo, tmp.data(), P.grid_disk, &P.rows_disk_opt[(size_t)s * P.grid_disk.n()]);
        }
    }
    if (!opts.no_dust) {
        const int nd = P.grid_disk.n();
        P.rows_disk_ir.assign((size_t)P.nir_sed * nd, 0.0);
        if (P.ir_kind == EGGPHOT_IR_CS17) P.rows_disk_pah.assign((size_t)P.nir_sed * nd, 0.0);
        for (uint32_t s = 0; s < P.nir_sed; ++s) {
            resample(xi, L.ir_sed + (size_t)s * P.nir, P.grid_disk, &P.rows_disk_ir[(size_t)s * nd]);
            if (P.ir_kind == EGGPHOT_IR_CS17)
                resample(xi, L.ir_pah + (size_t)s * P.nir, P.grid_disk, &P.rows_disk_pah[(size_t)s * nd]);
        }
    }

    // ---- map of the disk grid onto the optical (bulge) grid
    if (P.has_bulge) {
        const int nd = P.grid_disk.n(), nb = P.grid_bulge.n();
        P.u_oleft.assign(nd, -1); P.u_ofrac.assign(nd, 0.0); P.u_isopt.assign(nd, 0);
        int o = -1;
        for (int j = 0; j < nd; ++j) {
            const double xj = P.grid_disk.x[j];
            while (o + 1 < nb && P.grid_bulge.x[o + 1] <= xj) ++o;
            P.u_oleft[j] = o;
            if (o >= 0 && P.grid_bulge.x[o] == xj) P.u_isopt[j] = 1;
            else if (o >= 0 && o + 1 < nb)
                P.u_ofrac[j] = (xj - P.grid_bulge.x[o]) / (P.grid_bulge.x[o + 1] - P.grid_bulge.x[o]);
            if (o < 0 || (o + 1 >= nb && !P.u_isopt[j])) P.u_oleft[j] = -1; // outside the optical span: no bulge light
        }
        P.u_dxo.assign(nd, 0.0);
        int next = -1; // disk-grid index of the next optical node
        for (int j = nd - 1; j >= 0; --j)
            if (P.u_isopt[j]) {
                if (next >= 0) P.u_dxo[j] = P.grid_disk.x[next] - P.grid_disk.x[j];
                next = j;
            }
    }

    // ---- band tables, packed into groups under the shared-memory budget
    // default: what the other shared-memory arrays of the flux kernel (phot_kernels.cu, smem_layout: per galaxy of a
    // batch the flux accumulators, node windows and line windows) leave of the 227 KB a CTA can have
    size_t budget = opts.smem_table_budget;
    if (!budget) {
        const size_t nb = P.bands.size(), nr = P.rfbands.size(), nl = P.no_nebular ? 0 : P.line_lambda.size();
        const size_t other = (size_t)batch_of(P.precision) * (nb * (16 + 7) + nr * 16 + nl * 32) + 8 * 128 + 4096;
        budget = other + 64u * 1024 < 232448u ? 232448u - other : 64u * 1024;
    }
    const size_t hard_cap = opts.smem_table_budget ? std::max<size_t>(budget, 204u * 1024) : budget; // of one band
    const int bfac = P.precision == EGGPHOT_FP64 ? 2 : 4;
    std::vector<BandTableD> tabs(P.bands.size());
    for (size_t b = 0; b < P.bands.size(); ++b) {
        st = build_band_table(P.bands[b], (uint32_t)b, bfac, tabs[b]);
        if (!st.ok()) return st;
        if (band_bytes(tabs[b], P.precision) + sizeof(BandHeader) > hard_cap)
            return Status::error(EGGPHOT_ERR_UNSUPPORTED, "filter " + std::to_string(P.band_order[b]) +
                                 " has too many nodes for the shared-memory band tables");
    }
    // bands need not be contiguous in a group: first-fit-decreasing into as few groups as the budget allows
    // (fewer groups = fewer table loads and barriers per batch of galaxies)
    {
        std::vector<size_t> size(tabs.size()), bysize(tabs.size());
        size_t total = 0;
        for (size_t b = 0; b < tabs.size(); ++b) {
            size[b] = band_bytes(tabs[b], P.precision) + sizeof(BandHeader) + 16;
            total += size[b];
            bysize[b] = b;
            const Grid &gd = P.grid_disk;
            const int lo = last_le(gd.x, tabs[b].h.eff_first / 2.0), hi = last_le(gd.x, tabs[b].h.eff_last / 2.0);
            tabs[b].h.cost = (float)std::max(hi - lo, 1);
        }
        std::stable_sort(bysize.begin(), bysize.end(), [&](size_t i, size_t j) { return size[i] > size[j]; });
        std::vector<std::vector<size_t>> bins;
        for (size_t ng = std::max<size_t>((total + budget - 1) / budget, 1); !tabs.empty(); ++ng) {
            bins.assign(ng, {});
            std::vector<size_t> fill(ng, 0);
            bool ok = true;
            for (size_t b : bysize) {
                size_t best = ng;
                for (size_t g = 0; g < ng; ++g) // the emptiest group that still has room: keeps the groups even
                    if (fill[g] + size[b] <= budget && (best == ng || fill[g] < fill[best])) best = g;
                if (best == ng) {
                    if (size[b] > budget && ng >= tabs.size()) { best = 0; for (size_t g = 1; g < ng; ++g) if (fill[g] < fill[best]) best = g; }
                    else { ok = false; break; }
                }
                bins[best].push_back(b);
                fill[best] += size[b];
            }
            if (ok) break;
        }
        for (auto &bin : bins) {
            if (bin.empty()) continue;
            // inside a group the bands are stored by decreasing expected work (SED nodes under the
            // filter at z = 1): warps pull items in storage order, long ones first
            std::stable_sort(bin.begin(), bin.end(), [&](size_t i, size_t j) { return tabs[i].h.cost > tabs[j].h.cost; });
            BandGroup g;
            g.blob.assign(bin.size() * sizeof(BandHeader), 0);
            for (size_t k = 0; k < bin.size(); ++k) {
                g.cols.push_back((uint32_t)bin[k]);
                append_band(g.blob, k * sizeof(BandHeader), tabs[bin[k]], P.precision);
            }
            P.groups.push_back(std::move(g));
        }
    }

    // ---- rest-frame weights
    for (size_t b = 0; b < P.rfbands.size(); ++b) {
        P.rf_bulge.push_back(P.has_bulge ? rf_weights(P.rfbands[b], P.grid_bulge, P.nodes_mode) : RfWeights());
        P.rf_disk.push_back(rf_weights(P.rfbands[b], P.grid_disk, P.nodes_mode));
    }
    return build_lane_tables(P, opts);
}

} // namespace eggphot
