// phot_lane.cuh — the "lane = galaxy" form of the quadrature, shared by the device kernel
// (phot_lane_kernels.cu), the host-side table builder and the CPU-tier emulation (tests/host_emul).
//
// Same mathematics as phot_core.cuh ([vif] sed2flux on the union of filter and SED nodes, src/egg-gencat.cpp:2155,
// :2197, as T = sum_j S_j W_j with node weights W_j = E_j - E_{j-1}), re-arranged for one thread per galaxy walking
// the SED nodes of a band in order:
//
//  * E_j, the weight density of the SED interval [t_j, t_{j+1}], is formed from two expansions of Lambda = L - G/6
//    about FILTER NODES, not from "exact integral + trapezoid-error term":
//        Lambda_c(t) = Lambda_c + v (K0_c + v R_c / 2),   v = t - f_c   (v < 0 allowed)
//    - an interval whose ends lie in different filter cells:
//          E_j = (Lambda_in(t_{j+1}) - Lambda_out(t_j)) / (t_{j+1} - t_j)
//      Lambda_in expands about the LEFT node of the cell that holds t_{j+1}, Lambda_out about the RIGHT node of the cell
//      that holds t_j.  Lambda_out is exactly  Lambda(t_j) - gamma_j / 2  of phot_core.cuh (the union cell from t_j to
//      the end of its filter cell), so the separate gamma term is gone, and with it the cancellation between the two:
//      a bright SED node just outside a filter whose true weight is 0 gets the weight 0 exactly (every table value it
//      touches is 0 in the gauge of that end), in float and in double;
//    - an interval inside one filter cell c:  E_j = K0_c + (u_j + u_{j+1}) R_c / 2 + s_c u_j u_{j+1} / 2  (direct, no
//      difference at all);
//  * gauges anchored at both ends of the span and in every deep interior valley of the response (see BandHeader2): table
//    values are small wherever the response is, exactly 0 where it is 0;
//  * the disk walks every node of the merge_add grid, the bulge only the optical nodes among them (its own SED intervals,
//    so the union grid of the bulge is filter nodes + optical nodes, as when get_flux runs on the bulge alone,
//    src/egg-gencat.cpp:2206-2209); both read the same per-node expansions.  Where two consecutive disk nodes are both
//    optical the interval is shared.
//
// All of Lambda, E and W are IEEE double in both arithmetic modes; fp32c rounds W_j to float, multiplies the float SED
// value and adds in float within a block of 64 nodes; blocks are added in double in a fixed order.  Weights of a
// non-negative response are non-negative: no cancellation in the sums.
#ifndef EGGPHOT_PHOT_LANE_CUH
#define EGGPHOT_PHOT_LANE_CUH

#include <stdint.h>
#include <math.h>
#include <string.h>

#include "phot_core.cuh"

namespace eggphot {

constexpr int kLanes = 32;       // galaxies per batch (one per lane)

// One record per filter node: wavelength, Lambda and K0 at the node in the gauge of the record, R / 2 of the node.
struct alignas(32) Rec { double f, L, k0, hr; };

// Header of one band table inside a group blob of the lane kernel (offsets are bytes from the blob start).
//
// Nodes.  The table holds the nodes of the response with its identically-zero wings removed, plus a zero-response node
// at either end where the response starts or stops abruptly (a jump is two nodes at one wavelength), so that R = 0 at
// both ends and outside.  A filter too large for the shared-memory budget is cut at filter nodes into several tables
// that feed the same output column (the union-node trapezoid is additive over such cuts).
//
// Gauges.  Adding a linear function to Lambda shifts every E by a constant and leaves the weights untouched, and a
// weight is only as accurate as the table values it differences are small.  The span of a band is therefore cut into
// regions, each with an ANCHOR node where (Lambda, K0) = (0, 0) and from which both are accumulated outwards: the two
// ends of the span are anchors (a far wing only ever differences small numbers, however bright the SED is there), and so
// is the deepest node of every interior valley of |R| below 1e-6 of its maximum (between the lobes of a two-lobe
// response the tables are exactly 0 where R is: a bright line in the gap weighs exactly nothing).  Regions meet at SPLIT
// nodes (the half-mass node between two anchors), stored twice: once in the gauge of the region on its left (closing
// the last cell of that region) and once in the gauge of the region on its right.  An SED interval that crosses split s
// adds the gauge jump  Lambda_s(t) - Lambda_{s+1}(t) = C_s + D_s (t - F_s), and the next weight sees E - D_s (1+z).
//
// Records: [left sentinel][nodes, split nodes twice][terminator].  The left sentinel opens the cell of everything blue
// of the span, the last node that of everything red of it: all their table values are 0, so an SED node outside the
// span needs no special case (its expansions are exactly 0).  bkt has one entry in front and one behind the real
// buckets for those two cells.
constexpr int kMaxSplit = 6;
constexpr uint32_t kBktRecMask = 0x7fffu, kBktStep2 = 0x8000u; // a bucket entry: record number, "step two records" flag
struct __attribute__((aligned(16))) BandHeader2 {
    double full_first, full_last; // full span of the whole filter [um] for the coverage test ([vif] sed2flux NaN rule)
    double f_first, f_last;       // span of this table [um]
    double bk_scale, bk_c0;       // bucket entry of t: round-to-nearest of fma(t, bk_scale, bk_c0), clamped to the table
    double splitC[kMaxSplit], splitD[kMaxSplit], splitF[kMaxSplit];
    uint32_t n;                   // records (0: the band contributes nothing)
    uint32_t nsplit;              // regions - 1
    uint32_t nbkt;                // bucket entries, the two end entries included
    uint32_t two_probe;           // buckets are narrower than every cell: two comparisons settle the lookup
    uint32_t rec_off, hs_off, bkt_off;
    uint32_t out_col;             // column in the sorted band list
    float cost;
    uint32_t pad_;
    uint32_t msplit[kMaxSplit];   // record of the left-gauge copy of each split node (ascending); 0x7fffffff if unused
};
static_assert(sizeof(BandHeader2) == 256, "BandHeader2 layout");

struct Band2 {
    const BandHeader2 *h;  // split constants are read from the header when an interval crosses a split
    const Rec *rec;
    const double *hs;      // half the slope of R in the cell whose left node is record r (0 in zero-width cells)
    const uint16_t *bkt;   // last record whose node is at or below the (slightly lowered) start of the bucket; bit 15
                           // (kBktStep2): the next node is two records at one wavelength
    double bk_scale, bk_c0;
    int m0, nsplit, nbkt;  // m0 = msplit[0]
    bool two_probe;
};
EGG_HD void make_band2(const uint8_t *blob, const BandHeader2 &h, Band2 &v) {
    v.h = &h;
    v.rec = reinterpret_cast<const Rec *>(blob + h.rec_off);
    v.hs = reinterpret_cast<const double *>(blob + h.hs_off);
    v.bkt = reinterpret_cast<const uint16_t *>(blob + h.bkt_off);
    v.bk_scale = h.bk_scale; v.bk_c0 = h.bk_c0;
    v.m0 = (int)h.msplit[0]; v.nsplit = (int)h.nsplit; v.nbkt = (int)h.nbkt;
    v.two_probe = h.two_probe != 0;
}

// One node of a component grid as the walk reads it (the same for every galaxy: loaded with a warp-uniform address).
// o / po / idxpo tie the disk grid to the optical grid (kept for the table builder's checks).
struct alignas(16) GridNode {
    double x;      // rest wavelength [um]
    double idxp;   // 1 / (x_j - x_{j-1}); 0 for j = 0
    double idxpo;  // optical node of the disk grid: 1 / (x_j - x of the previous optical node), 0 if there is none; else 0
    int32_t o;     // disk grid: index in the optical (bulge) grid, -1 if the node is not an optical node
    int32_t po;    // disk-grid index of the last optical node before j, -1 if none
};
constexpr int32_t kNextNotOptical = 1 << 30;
EGG_HD int grid_o(const GridNode &g) { return g.o < 0 ? -1 : (g.o & (kNextNotOptical - 1)); }
static_assert(sizeof(GridNode) == 32, "GridNode layout");
// rows of a component grid as the lane kernels hold it: the real nodes, two virtual nodes far to the red, and padding to a
// multiple of 8
EGG_HD int lane_grid_rows(int n) { return (n + 2 + 7) & ~7; }

// The walk kernel (phot_lane_kernels.cu): warps per CTA, arrivals per strip (their cell lookups run ahead of the
// arithmetic as independent chains), SED rows per staging buffer (two buffers per warp).
#ifndef EGG_WALK_THREADS
#define EGG_WALK_THREADS 640
#endif
constexpr int kWalkThreads = EGG_WALK_THREADS; // 20 warps at 96 registers
constexpr int kSub = 4;
#ifndef EGG_WALK_STAGE
#define EGG_WALK_STAGE 8
#endif
constexpr int kStage = EGG_WALK_STAGE;
// Shared memory of the walk kernel that is not band tables: the staging buffers, and {x, 1 / dx} of the two component
// grids unless they would leave less than half of the shared memory to the band tables (then they are read from global
// memory through L1).
EGG_HD size_t lane_stage_bytes(bool fp64) { return (size_t)(kWalkThreads / 32) * 2 * kStage * 32 * (fp64 ? 8 : 4); }
EGG_HD size_t lane_gx_bytes(int nD, int nB) { return ((size_t)lane_grid_rows(nD) + lane_grid_rows(nB)) * 16; }
#ifndef EGG_GX_SMEM_DIVISOR
#define EGG_GX_SMEM_DIVISOR 2   // (tuning knob: a huge value keeps the grids out of shared memory, fewer and larger band groups)
#endif
EGG_HD bool lane_gx_in_smem(bool fp64, int nD, int nB, size_t limit) { return lane_gx_bytes(nD, nB) + lane_stage_bytes(fp64) <= limit / EGG_GX_SMEM_DIVISOR; }

// Shared memory of the lane kernel: the band tables of the current group, and a little besides.
constexpr size_t kSmemPerCta = 232448;   // 227 KB
constexpr size_t kLaneFixedSmem = 2048;  // barrier, per-warp odds and ends

// What the intervals on either side need from an SED node at observed wavelength t.
struct NodeEv {
    double u;         // offset from the left node of its filter cell
    double Lin;       // expansion about the left node of its cell
    double Lout;      // expansion about the right node of its cell
    double k0, hr, hs; // of the cell: for an interval that stays inside it
    int r;            // record of the cell's left node
};

// integer nearest to y (|y| < 2^31) without a conversion instruction: the low word of y + 1.5 * 2^52
EGG_HD int round_lo(double y) {
    const double z = y + 6755399441055744.0;
#if defined(__CUDA_ARCH__)
    return __double2loint(z);
#else
    uint64_t b;
    memcpy(&b, &z, 8);
    return (int)(uint32_t)b;
#endif
}

EGG_HD void eval_node(const Band2 &b, double t, NodeEv &e) {
    int k = round_lo(fma(t, b.bk_scale, b.bk_c0));
    k = k < 0 ? 0 : (k > b.nbkt - 1 ? b.nbkt - 1 : k);
    const int be = b.bkt[k];
    int r = be & (int)kBktRecMask;
    if (b.two_probe) {
        // one probe; a split node or an abrupt end is two records at one wavelength, and the entry says so
        if (t >= b.rec[r + 1].f) r += (be & (int)kBktStep2) ? 2 : 1;
    } else {
        while (t >= b.rec[r + 1].f) ++r;
    }
    const Rec q0 = b.rec[r], q1 = b.rec[r + 1];
    const double u = t - q0.f, v = t - q1.f;
    e.u = u;
    e.Lin = fma(u, fma(u, q0.hr, q0.k0), q0.L);
    e.Lout = fma(v, fma(v, q1.hr, q1.k0), q1.L);
    e.k0 = q0.k0; e.hr = q0.hr; e.hs = b.hs[r];
    e.r = r;
}

// E (times 1+z) of the SED interval from node p to node e, in the gauge of p's region; idx = 1 / (x_e - x_p) in the
// rest frame; te = observed wavelength of e.  `jump` is what the next weight subtracts from it (D (1+z) summed over the
// splits the interval crosses): E - jump is the same quantity in the gauge of e's region.
EGG_HD double interval_E(const Band2 &b, const NodeEv &p, const NodeEv &e, double te, double idx, double opz, double &jump) {
    jump = 0.0;
    if (e.r == p.r) return opz * fma(e.hs * p.u, e.u, fma(p.u + e.u, e.hr, e.k0));
    double d = e.Lin - p.Lout;
    if (b.nsplit > 0 && p.r <= (int)b.h->msplit[b.nsplit - 1] && e.r > b.m0) {
        double dsum = 0.0;
        for (int s = 0; s < b.nsplit; ++s) {
            const int ms = (int)b.h->msplit[s];
            if (p.r <= ms && e.r > ms) { // the interval crosses split s
                d += fma(b.h->splitD[s], te - b.h->splitF[s], b.h->splitC[s]);
                dsum += b.h->splitD[s];
            }
        }
        jump = dsum * opz;
    }
    return d * idx;
}

// Weighted sum of one band table for one galaxy and ONE component grid (the disk on the merge_add grid, the bulge on the
// optical grid: its SED intervals run between optical nodes, as when get_flux runs on the bulge alone,
// src/egg-gencat.cpp:2206-2209): tot = sum_j S_j W_j.  The walk arrives at nodes js..je: arriving at node j closes the
// interval [j-1, j] and settles the weight of node j-1.  js is at or before the last node left of the table's span, je
// after the first node right of it: every node outside the span has the weight 0 exactly for every galaxy of the batch,
// so the window may be wider than a galaxy needs.  The walk starts from a state blue of every filter.  Terms are summed
// in `real` over strips of four arrivals cut at absolute node numbers ([4 k, 4 k + 3]) and the strips added in double:
// the sum does not depend on where the walk starts.
constexpr int kLaneStrip = 4;
template <typename real>
EGG_HD double walk_window(const Band2 &bv, const GridNode *grid, double opz, int js, int je, const real *sed) {
    NodeEv prev, cur;
    prev.u = 0.0; prev.Lin = 0.0; prev.Lout = 0.0; prev.k0 = 0.0; prev.hr = 0.0; prev.hs = 0.0; prev.r = -1;
    double Eprev = 0.0, tot = 0.0;
    real Sprev = real(0), acc = real(0);
    for (int j = js; j <= je; ++j) {
        const GridNode g = grid[j];
        const double t = opz * g.x; // the reference's lam = rlam * (1 + z) (src/egg-gencat.cpp:2161), in double
        eval_node(bv, t, cur);
        double jump;
        const double E = interval_E(bv, prev, cur, t, g.idxp, opz, jump);
        acc = fma((real)(E - Eprev), Sprev, acc);
        Eprev = E - jump;
        prev = cur; Sprev = sed[j];
        if ((j & (kLaneStrip - 1)) == kLaneStrip - 1 || j == je) { tot += (double)acc; acc = real(0); }
    }
    return tot;
}

} // namespace eggphot
#endif
