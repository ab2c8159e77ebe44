// phot_lane_kernels.cu — the lane-per-galaxy flux pipeline (sm_100a) and the redshift bucketing that feeds it.
//
// Replaces the two get_flux passes and the totals of src/egg-gencat.cpp:2097-2239, like flux_kernel of phot_kernels.cu,
// with the work laid out the other way round: a warp's 32 lanes are 32 GALAXIES of neighbouring redshift walking the SAME
// SED nodes of one band, not 32 nodes of one galaxy.
//   * zsort_* kernels bucket the galaxies of a call by log2(1+z) (65536 buckets, counting sort): batches of 32
//     consecutive entries of the permutation have (1+z) within ~1e-4, so their lanes look up the same or adjacent
//     filter cells (shared-memory broadcasts instead of 32-way gathers), share one node window per band and read one
//     warp-uniform grid record per node.  Results do not depend on the batching: every galaxy's sums are formed by its
//     own lane in an order fixed by absolute node numbers.
//   * lane_plan_kernel (a warp per batch): the node window of every band table for the batch, and the node ranges the
//     batch needs built.
//   * lane_build_kernel (a CTA per batch, a warp per segment of 256 nodes): rest-frame SEDs of the 32 disks
//     (get_opt_sed / get_ir_sed / merge_add / IGM, :708-712, :2039-2057, :2118-2140) and bulges, emission lines included
//     (:2078-2095), into HBM rows laid out [node][lane] (coalesced 128-byte rows).
//   * lane_walk_kernel (one persistent CTA per SM): every CTA keeps the node records of ONE band group in shared memory
//     (TMA bulk copies, SASS UBLKCP + mbarrier) and its warps draw (batch, table) items of that group from a counter; every
//     lane walks the table's window for its galaxy (phot_lane.cuh: two expansions per node, weights from their
//     differences; no shuffles, no per-lane windows).  A CTA whose group is exhausted moves to the group with the most
//     work left.  No barrier inside the walk; the table sums go to [batch][table][component][lane] doubles.
//   * lane_finish_kernel (a warp per batch): rest-frame magnitudes (:2150-2158), table sums added per output column in a
//     fixed order (bit-reproducible), data conditions (uncovered band -> 0, :2198; non-finite magnitude -> 99, :2156),
//     totals (:2238-2239), catalog columns.
#include "phot_kernels.cuh"

#include <cstdio>
#include <cstdlib>

#include "../../include/eggphot.h"
#include "phot_device.cuh"
#include "phot_lane.cuh"
#include "phot_tables.hpp"

namespace eggphot {

namespace {

// ---- redshift bucketing -----------------------------------------------------------------------------
constexpr int kZBuckets = 1 << 16;
__device__ __forceinline__ int zbucket(float z) {
    // log2(1+z) in [0, 8) -> 65536 buckets: (1+z) resolved to 8.5e-5; anything else (NaN, z < 0, z > 255) to the ends
    const float v = __log2f(1.0f + fmaxf(z, 0.0f)) * (float)(kZBuckets / 8);
    int k = (int)v;
    return k < 0 ? 0 : (k > kZBuckets - 1 ? kZBuckets - 1 : k);
}
__global__ void zsort_count(const float *__restrict__ z, uint32_t n, uint32_t *__restrict__ hist) {
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x)
        atomicAdd(&hist[zbucket(__ldg(z + g))], 1u);
}
// exclusive prefix sums of the 65536 counters, one block of 1024 threads (64 counters each)
__global__ void __launch_bounds__(1024) zsort_scan(uint32_t *__restrict__ hist) {
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;
    uint32_t s = 0;
    for (int k = 0; k < kZBuckets / 1024; ++k) s += hist[t * (kZBuckets / 1024) + k];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const uint32_t v = t >= o ? part[t - o] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - s;
    for (int k = 0; k < kZBuckets / 1024; ++k) {
        const uint32_t c = hist[t * (kZBuckets / 1024) + k];
        hist[t * (kZBuckets / 1024) + k] = run;
        run += c;
    }
}
__global__ void zsort_scatter(const float *__restrict__ z, uint32_t n, uint32_t *__restrict__ cursor, uint32_t *__restrict__ perm) {
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x)
        perm[atomicAdd(&cursor[zbucket(__ldg(z + g))], 1u)] = g;
}

// ---- FMA micro-benchmark: the measured peak of the CUDA-core FP32 / FP64 pipes (the roofline of SURVEY.md 8d) ---------
template <typename T> __global__ void __launch_bounds__(1024) fma_peak_kernel(T *out, int iters, T a, T b) {
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const T s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == (T)123456789) out[0] = s; // never true: keeps the chains alive
}

// ---- helpers ------------------------------------------------------------------------------------------
template <typename real> struct Vec16 { real v[16 / sizeof(real)]; };
__device__ __forceinline__ Vec16<float> ldg16(const float *p) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
    return Vec16<float>{{a.x, a.y, a.z, a.w}};
}
__device__ __forceinline__ Vec16<double> ldg16(const double *p) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    return Vec16<double>{{a.x, a.y}};
}
// The SED rows of a batch live in HBM by the time the bands read them (16 batches per SM are in flight): every trip of
// the walk asks L2 for the rows it will need kPrefetchRows nodes later, so that the loads themselves hit L2.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// last j in [0, n) with x[j] <= v (-1 if none) / first j with x[j] >= v (n if none): plain bisection in double
__device__ __forceinline__ int bis_last_le(const double *__restrict__ x, int n, double v) {
    int lo = -1, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(x + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int bis_first_ge(const double *__restrict__ x, int n, double v) {
    int lo = -1, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(x + mid) < v) lo = mid; else hi = mid;
    }
    return hi;
}

// ---- the walk of one band table (device form of walk_window, phot_lane.cuh: same arithmetic, scheduled by hand) ---------
// Band constants of a table, addresses as 32-bit shared addresses.
struct BandDev {
    uint32_t rec, hs, bkt, hdr; // shared addresses of the record array, the half-slope array, the bucket table, the header
    double bk_scale, bk_c0;
    int nbkt_m1, nsplit;
    uint32_t nsa0;              // shared address of the record of the first split node (0xffffffff if there is none)
};

template <typename real> __device__ __forceinline__ real to_real(double v);
template <> __device__ __forceinline__ float to_real<float>(double v) { return (float)v; }
template <> __device__ __forceinline__ double to_real<double>(double v) { return v; }

// shared-memory loads by 32-bit shared address.  Not volatile: the compiler may schedule them freely (the addresses are
// computed from table headers read after the barrier that publishes the tables).
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(uint32_t a) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return (uint32_t)v;
}

// The interval from a node in the cell at record address pra to a node at t in the cell at ra crosses at least one split:
// the gauge jumps to add to the difference of the expansions, D summed over the crossed splits, and the next split at or
// after ra.  (Returned by value: a reference parameter of a function that is not inlined would pin the walk's state in
// local memory.)
struct Cross { double dadd, dsum; uint32_t nsa; };
__device__ __noinline__ Cross cross_splits(uint32_t hb, uint32_t rec, int nsplit, uint32_t pra, uint32_t ra, double t) {
    Cross c{0.0, 0.0, 0xffffffffu};
    for (int s = nsplit - 1; s >= 0; --s) {
        uint32_t ms;
        asm("ld.shared.u32 %0, [%1];" : "=r"(ms) : "r"(hb + (uint32_t)offsetof(BandHeader2, msplit) + 4u * s));
        const uint32_t msa = rec + 32u * ms;
        if (msa >= ra) c.nsa = msa;
        if (pra <= msa && ra > msa) {
            const double C = lds_f64(hb + (uint32_t)offsetof(BandHeader2, splitC) + 8u * s);
            const double D = lds_f64(hb + (uint32_t)offsetof(BandHeader2, splitD) + 8u * s);
            const double F = lds_f64(hb + (uint32_t)offsetof(BandHeader2, splitF) + 8u * s);
            c.dadd += fma(D, t - F, C);
            c.dsum += D;
        }
    }
    return c;
}

// E of the interval from the node with state (pu, pLout, pra, pnsa) to the node at t in the cell at ra, in the gauge of the
// first node's region; Eadj = the same quantity in the gauge of the second node's region; nsa = the second node's next
// split.  (interval_E of phot_lane.cuh with record addresses for record numbers; both forms are evaluated and one is
// selected: the branch would cost more than the three extra operations.)
__device__ __forceinline__ double interval_dev(const BandDev &B, double pu, double pLout, uint32_t pra, uint32_t pnsa, uint32_t ra,
                                               double u, double Lin, double k0, double hr, double hs, double t, double idx, double opz,
                                               double &Eadj, uint32_t &nsa) {
    nsa = pnsa;
    const double d = Lin - pLout;
    const double Esame = opz * fma(hs * pu, u, fma(pu + u, hr, k0));
    double E = ra == pra ? Esame : d * idx;
    Eadj = E;
    if (ra > pnsa) { // crosses a split (once or twice per band)
        const Cross c = cross_splits(B.hdr, B.rec, B.nsplit, pra, ra, t);
        nsa = c.nsa;
        E = (d + c.dadd) * idx;
        Eadj = E - c.dsum * opz;
    }
    return E;
}

// the records of the cell of a node: {f, Lambda} and {K0, R/2} of its left and right nodes, half the slope of R in it
struct Recs { double2 a0, b0, a1, b1; double hs; };
__device__ __forceinline__ Recs load_recs(const BandDev &B, uint32_t ra) {
    Recs R;
    R.a0 = lds_v2f64(ra); R.b0 = lds_v2f64(ra + 16u); R.a1 = lds_v2f64(ra + 32u); R.b1 = lds_v2f64(ra + 48u);
    R.hs = lds_f64(B.hs + ((ra - B.rec) >> 2));
    return R;
}
__device__ __forceinline__ int lds_s16(uint32_t a) {
    int16_t v;
    asm("ld.shared.s16 %0, [%1];" : "=h"(v) : "r"(a));
    return (int)v;
}

#ifndef EGG_LANE_SUB
#define EGG_LANE_SUB 4
#endif
constexpr int kSub = EGG_LANE_SUB; // arrivals per strip of the walk, strips cut at multiples of kSub

// Walks nodes js..je of the disk grid (js a multiple of kSub, je + 1 too) for one galaxy (lane) and one band table.  Per
// strip: (1) the filter cell of every node is looked up - independent chains of shared-memory loads, no branch - and
// (2) the nodes are taken in order: records of the cell, the two expansions, the interval that ends at the node, the
// weight of the node before it.  Loads run ahead of the arithmetic: the records of the next node, the SED values two
// nodes on (they come from L2, asked for two strips ahead: the rows of a batch live in HBM by the time the bands read
// them).  The bulge's SED intervals run between optical nodes: at an optical node the interval from the last optical
// node is closed (from the state kept of that node) and settles its weight.  totd / totb receive the sums of the strips
// of 16 arrivals (each summed in `real`, the strips added in double, in order: the same numbers whichever way the window
// is cut).
template <typename real, bool TWOP, bool BULGE>
__device__ __forceinline__ void walk_table(const BandDev &B, const GridNode *__restrict__ grid, uint32_t gx_sa, uint32_t on_sa, double opz,
                                           int js, int je, const real *__restrict__ sd, const real *__restrict__ sb, double &totd, double &totb) {
    const int lane = threadIdx.x & 31;
    double Pu = 0.0, PLout = 0.0, Ou = 0.0, OLout = 0.0; // state of the previous node / of the last optical node
    uint32_t Pra = 0u, Pnsa = B.nsa0, Ora = 0u, Onsa = B.nsa0;
    double Eprev_d = 0.0, Eprev_b = 0.0, td = 0.0, tb = 0.0;
    real Sprev_d = real(0), Sprev_b = real(0), accd = real(0), accb = real(0);
    constexpr int kLinesPerRow = (int)sizeof(real) / 4; // a row is 32 lanes x sizeof(real): one or two 128-byte lines
    constexpr uint32_t kRowBytes = kLanes * sizeof(real);
    constexpr int kAhead = 16;                          // rows asked of L2 ahead of the walk
    const char *sd_rows = reinterpret_cast<const char *>(sd - lane), *sb_rows = reinterpret_cast<const char *>(sb - lane);
    const char *sdl = reinterpret_cast<const char *>(sd), *sbl = reinterpret_cast<const char *>(sb);
    const int pf_row = lane / kLinesPerRow, pf_off = 128 * (lane % kLinesPerRow);
    if (lane < kAhead * kLinesPerRow) prefetch_l2(sd_rows + (size_t)(js + pf_row) * kRowBytes + pf_off);
    int o_pf = 0; // bulge rows asked of L2 so far
    if (BULGE) {
        o_pf = __ldg(&grid[js].o) >= 0 ? grid_o(grid[js]) : (grid[js].po >= 0 ? grid_o(grid[grid[js].po]) + 1 : 0);
        if (lane < kAhead * kLinesPerRow) prefetch_l2(sb_rows + (size_t)(o_pf + pf_row) * kRowBytes + pf_off);
        o_pf += kAhead;
    }
    // the pipeline of SED values: node j (0) and node j + 1 (1); o = index in the bulge grid, -1 if the node is not optical
    int o0 = BULGE ? lds_s16(on_sa + 2u * (uint32_t)js) : -1, o1 = BULGE ? lds_s16(on_sa + 2u * (uint32_t)(js + 1)) : -1;
    real Sd0 = __ldcg(reinterpret_cast<const real *>(sdl + (uint32_t)js * kRowBytes));
    real Sd1 = __ldcg(reinterpret_cast<const real *>(sdl + (uint32_t)(js + 1) * kRowBytes));
    real Sb0 = real(0), Sb1 = real(0);
    if (BULGE && o0 >= 0) Sb0 = __ldcg(reinterpret_cast<const real *>(sbl + (uint32_t)o0 * kRowBytes));
    if (BULGE && o1 >= 0) Sb1 = __ldcg(reinterpret_cast<const real *>(sbl + (uint32_t)o1 * kRowBytes));
    double idxpo = (BULGE && o0 >= 0) ? __ldg(&grid[js].idxpo) : 0.0;
#pragma unroll 1
    for (int j0 = js; j0 <= je; j0 += kSub) {
        if ((j0 & (kAhead / 2 - 1)) == 0 && lane < (kAhead / 2) * kLinesPerRow) {
            prefetch_l2(sd_rows + (size_t)(j0 + kAhead + pf_row) * kRowBytes + pf_off);
            if (BULGE) {
                // the optical nodes run at most as fast as the disk nodes: stay kAhead rows ahead of the last one met
                const int cur = max(o0, o1);
                if (cur >= 0 && cur + kAhead > o_pf) { prefetch_l2(sb_rows + (size_t)(o_pf + pf_row) * kRowBytes + pf_off); o_pf += kAhead / 2; }
            }
        }
        // (1) filter cells
        uint32_t ra[kSub];
#pragma unroll
        for (int k = 0; k < kSub; ++k) {
            const double t = opz * lds_f64(gx_sa + 16u * (uint32_t)(j0 + k));
            int kk = round_lo(fma(t, B.bk_scale, B.bk_c0));
            kk = min(max(kk, 0), B.nbkt_m1);
            uint32_t a = B.rec + 32u * lds_u16(B.bkt + 2u * (uint32_t)kk);
            if (TWOP) {
                // a split node or an abrupt end is two records at one wavelength: two probes, both read at once
                const double f1 = lds_f64(a + 32u), f2 = lds_f64(a + 64u);
                a += (t >= f1 ? 32u : 0u) + (t >= f2 ? 32u : 0u);
            } else {
                while (t >= lds_f64(a + 32u)) a += 32u;
            }
            ra[k] = a;
        }
        // (2) intervals and weights
        Recs R = load_recs(B, ra[0]);
        double2 G = lds_v2f64(gx_sa + 16u * (uint32_t)j0);
#pragma unroll
        for (int k = 0; k < kSub; ++k) {
            const uint32_t j = (uint32_t)(j0 + k);
            // loads that run ahead
            Recs Rn = R;
            double2 Gn = G;
            if (k + 1 < kSub) {
                Rn = load_recs(B, ra[k + 1]);
                Gn = lds_v2f64(gx_sa + 16u * (j + 1u));
            }
            const int o2 = BULGE ? lds_s16(on_sa + 2u * (j + 2u)) : -1;
            const real Sd2 = __ldcg(reinterpret_cast<const real *>(sdl + (j + 2u) * kRowBytes));
            real Sb2 = real(0);
            if (BULGE && o2 >= 0) Sb2 = __ldcg(reinterpret_cast<const real *>(sbl + (uint32_t)o2 * kRowBytes));
            double idxpo_n = 0.0;
            if (BULGE && o1 >= 0) idxpo_n = __ldg(&grid[j + 1u].idxpo);
            // the node
            const double t = opz * G.x; // the reference's lam = rlam * (1 + z) (src/egg-gencat.cpp:2161), in double
            const double u = t - R.a0.x, v = t - R.a1.x;
            const double Lin = fma(u, fma(u, R.b0.y, R.b0.x), R.a0.y);
            const double Lout = fma(v, fma(v, R.b1.y, R.b1.x), R.a1.y);
            // disk: close [j - 1, j], settle node j - 1
            double Eadj_d;
            uint32_t nsa_c;
            const double E_d = interval_dev(B, Pu, PLout, Pra, Pnsa, ra[k], u, Lin, R.b0.x, R.b0.y, R.hs, t, G.y, opz, Eadj_d, nsa_c);
            accd = fma(to_real<real>(E_d - Eprev_d), Sprev_d, accd);
            Eprev_d = Eadj_d;
            Sprev_d = Sd0;
            // bulge: the interval from the last optical node ends here and settles that node
            if (BULGE && o0 >= 0) {
                double Eadj_b;
                uint32_t dummy;
                const double E_b = interval_dev(B, Ou, OLout, Ora, Onsa, ra[k], u, Lin, R.b0.x, R.b0.y, R.hs, t, idxpo, opz, Eadj_b, dummy);
                accb = fma(to_real<real>(E_b - Eprev_b), Sprev_b, accb);
                Eprev_b = Eadj_b;
                Sprev_b = Sb0;
                Ou = u; OLout = Lout; Ora = ra[k]; Onsa = nsa_c;
            }
            Pu = u; PLout = Lout; Pra = ra[k]; Pnsa = nsa_c;
            R = Rn; G = Gn; idxpo = idxpo_n;
            o0 = o1; o1 = o2; Sd0 = Sd1; Sd1 = Sd2; Sb0 = Sb1; Sb1 = Sb2;
        }
        // the sums of a strip of 16 arrivals are formed in `real`, the strips added in double: strips are cut at absolute node
        // numbers, so the result does not depend on where the walk started
        if ((j0 & 15) == 16 - kSub) {
            td += (double)accd; accd = real(0);
            if (BULGE) { tb += (double)accb; accb = real(0); }
        }
    }
    totd = td + (double)accd;
    totb = BULGE ? tb + (double)accb : 0.0;
}

// ---- the kernels --------------------------------------------------------------------------------------
constexpr int kWinStride = 64;  // window words per batch in LaneParams::win (at most 64 band tables)
constexpr int kSegRows = 256;   // rows per warp of the build kernel

// lane -> galaxy of batch bl of the chunk (lanes beyond the end of the catalog repeat the batch's first galaxy)
__device__ __forceinline__ uint64_t lane_galaxy(const LaneParams &q, uint32_t bl, int lane, int &ng) {
    const uint64_t g0 = (q.batch0 + bl) * kLanes;
    ng = (int)min((uint64_t)kLanes, q.fp.ngal - g0);
    return __ldg(q.perm + g0 + (lane < ng ? lane : 0));
}

// A warp per batch: the node window of every band table for the batch and the node ranges to build.
__global__ void __launch_bounds__(256) lane_plan_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    const int lane = threadIdx.x & 31;
    const uint32_t bl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (bl >= q.nbatch) return;
    const int ntab = (int)q.ntab, nrf = p.do_rf ? (int)p.nrf : 0;
    const int nD = p.nD, nB = p.nB;
    const bool has_bulge = p.has_bulge != 0;
    const GridNode *grid = q.grid;
    int ng;
    const uint64_t g = lane_galaxy(q, bl, lane, ng);
    double omin = 1.0 + (double)__ldg(p.z + g), omax = omin;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        omin = fmin(omin, __shfl_xor_sync(kFull, omin, o));
        omax = fmax(omax, __shfl_xor_sync(kFull, omax, o));
    }
    // node window of band table fb for the batch: from the last node at or left of the table's blue end for the largest
    // (1+z) to the first node at or right of its red end for the smallest, with a relative margin (nodes a galaxy does
    // not weigh get the weight 0 exactly); 0 = nothing to do.  jlo | jhi << 15 | 1 << 30 | bulge << 31
    auto table_window = [&](int fb) -> uint32_t {
        int grp = 0;
        while (grp + 1 < (int)p.ngroups && fb >= (int)p.groups[grp + 1].base) ++grp;
        const BandHeader2 *gh = reinterpret_cast<const BandHeader2 *>(p.groups[grp].blob) + (fb - (int)p.groups[grp].base);
        const double ff = __ldg(&gh->full_first), fl = __ldg(&gh->full_last);
        // [vif] sed2flux is NaN (-> flux 0, src/egg-gencat.cpp:2198) unless the SED covers the filter: a band no
        // galaxy of the batch covers is skipped
        const bool any_cov = nD >= 2 && __ldg(p.xD) * omin <= ff && __ldg(p.xD + nD - 1) * omax >= fl;
        if (!any_cov || __ldg(&gh->n) == 0) return 0u;
        int jlo = max(bis_last_le(p.xD, nD, __ldg(&gh->f_first) / omax * (1.0 - 1e-9)), 0);
        int jhi = min(bis_first_ge(p.xD, nD, __ldg(&gh->f_last) / omin * (1.0 + 1e-9)), nD - 1);
        const bool bulge_any = has_bulge && nB >= 2 && __ldg(p.xB) * omin <= ff && __ldg(p.xB + nB - 1) * omax >= fl;
        if (bulge_any) widen_window_for_bulge(grid, q.next_opt, nD, jlo, jhi);
        return (uint32_t)jlo | ((uint32_t)jhi << 15) | (bulge_any ? 0x80000000u : 0u) | 0x40000000u;
    };
    int jn0 = nD, jn1 = -1, on0 = nB, on1 = -1; // disk / optical nodes some band of the batch reads
    for (int r = 0; r < nrf; ++r) {             // the rest-frame bands read fixed node ranges
        if (p.rfD[r].n > 0) { jn0 = min(jn0, (int)p.rfD[r].j0); jn1 = max(jn1, (int)(p.rfD[r].j0 + p.rfD[r].n - 1)); }
        if (has_bulge && p.rfB[r].n > 0) { on0 = min(on0, (int)p.rfB[r].j0); on1 = max(on1, (int)(p.rfB[r].j0 + p.rfB[r].n - 1)); }
    }
    uint32_t *win = q.win + (size_t)bl * kWinStride;
    for (int fb = lane; fb < kWinStride; fb += 32) {
        const uint32_t w = fb < ntab ? table_window(fb) : 0u;
        win[fb] = w;
        if (w) {
            jn0 = min(jn0, (int)(w & 0x7fffu));
            jn1 = max(jn1, (int)((w >> 15) & 0x7fffu));
        }
    }
    jn0 = __reduce_min_sync(kFull, jn0); jn1 = __reduce_max_sync(kFull, jn1);
    on0 = __reduce_min_sync(kFull, on0); on1 = __reduce_max_sync(kFull, on1);
    if (lane == 0) {
        // rows to build: the walks start at the optical node before a window and end at the one after it, in strips of 8
        // arrivals cut at multiples of 8 (rows a walk touches beyond the nodes a band weighs must hold finite values: their
        // weights are exactly 0); rows at or beyond nD are zero
        int4 pl = make_int4(0, -1, 0, -1);
        if (jn1 >= jn0) {
            pl.x = max(jn0 - (int)q.row_margin, 0) & ~(kSub - 1);
            pl.y = min((jn1 + (int)q.row_margin) | (kSub - 1), (int)q.rowsD - 1);
            if (has_bulge) { // the optical nodes among them
                const int jo0 = grid[pl.x].o >= 0 ? pl.x : q.next_opt[pl.x], jo1 = grid[pl.y].o >= 0 ? pl.y : grid[pl.y].po;
                if (jo0 <= jo1 && jo1 >= 0 && grid[jo0].o >= 0) { on0 = min(on0, grid_o(grid[jo0])); on1 = max(on1, grid_o(grid[jo1])); }
            }
        }
        if (has_bulge && on1 >= on0) {
            pl.z = max(on0 - 2, 0) & ~3;
            pl.w = min((on1 + 2) | 3, (int)q.rowsB - 1);
        }
        q.plan[bl] = pl;
    }
}

// A CTA per batch, a warp per segment of 256 rows: rest-frame SEDs, lane = galaxy, 16 bytes of each template row per
// load, [node][lane] stores; then the emission lines of the segment's nodes.  The bulge SED carries no mass scale
// (applied at the end).
template <typename real>
__global__ void __launch_bounds__(512) lane_build_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const uint32_t bl = blockIdx.x;
    const int4 pl = q.plan[bl];
    const int nsegD = pl.y >= pl.x ? (pl.y - pl.x + kSegRows) / kSegRows : 0;
    const int nsegB = pl.w >= pl.z ? (pl.w - pl.z + kSegRows) / kSegRows : 0;
    if (warp >= nsegD + nsegB) return;
    const int nD = p.nD, nB = p.nB, nline = p.nline;
    const bool cs17 = p.ir_kind == EGGPHOT_IR_CS17;
    const bool has_bulge = p.has_bulge != 0;
    real *SD = reinterpret_cast<real *>(reinterpret_cast<uint8_t *>(q.rows) + (size_t)bl * q.rows_stride);
    real *SB = SD + (size_t)q.rowsD * kLanes;
    const real *rO = reinterpret_cast<const real *>(p.rowsDopt), *rI = reinterpret_cast<const real *>(p.rowsDir);
    const real *rP = reinterpret_cast<const real *>(p.rowsDpah), *rB = reinterpret_cast<const real *>(p.rowsB);
    int ng;
    const uint64_t g = lane_galaxy(q, bl, lane, ng);

    // per-lane scalars (src/egg-gencat.cpp:708-712, 2039-2057)
    // the reference passes the FLOAT difference m[i]-out.m2lcor[i] to e10 (:2106, :2114)
    const float m2l = __ldg(p.m2lcor + g);
    const double a_bulge = has_bulge ? exp10((double)(__ldg(p.m_bulge + g) - m2l)) : 0.0;
    const real a_d = p.disk_mode != 2 ? (real)exp10((double)(__ldg(p.m_disk + g) - m2l)) : real(0);
    real c1 = real(0), c2 = real(0);
    bool bad = false;
    uint64_t ri = 0, rb = 0, rd = 0;
    if (p.disk_mode != 1) {
        ri = __ldg(p.ir_sed + g);
        if (ri >= p.nir_sed) { bad = true; ri = 0; }
        if (cs17) {
            const double fp = __ldg(p.fpah + g), md = __ldg(p.mdust + g);
            c1 = (real)((1.0 - fp) * md);
            c2 = (real)(fp * md);
        } else {
            c1 = (real)(double)__ldg(p.lir + g);
        }
    }
    if (has_bulge) {
        rb = __ldg(p.opt_sed_bulge + g);
        if (rb >= p.nopt_sed || !p.opt_use[rb]) { bad = true; rb = 0; }
    }
    if (p.disk_mode != 2) {
        rd = __ldg(p.opt_sed_disk + g);
        if (rd >= p.nopt_sed || !p.opt_use[rd]) { bad = true; rd = 0; }
    }
    if (bad && warp == 0) atomicExch(p.error_flag, EGGPHOT_ERR_INDEX);
    real g0f = real(1), g1f = real(1), g2f = real(1);
    if (p.apply_igm) { g0f = __ldg(p.igm_abs + 3 * g); g1f = __ldg(p.igm_abs + 3 * g + 1); g2f = __ldg(p.igm_abs + 3 * g + 2); }
    const double vd = nline > 0 ? (double)__ldg(p.vdisp + g) / 2.998e5 : 0.0;
    constexpr int NPL = 16 / (int)sizeof(real);

    for (int seg = warp; seg < nsegD + nsegB; seg += nwarp) {
        const bool disk = seg < nsegD;
        const int s0 = disk ? pl.x + seg * kSegRows : pl.z + (seg - nsegD) * kSegRows;
        const int s1 = min(s0 + kSegRows - 1, disk ? pl.y : pl.w);
        if (disk) {
            const real *rowO = rO + (size_t)rd * p.strideD, *rowI = rI + (size_t)ri * p.strideD, *rowP = rP + (size_t)ri * p.strideD;
#pragma unroll 4
            for (int j0 = s0; j0 <= s1; j0 += NPL) { // (unrolled: a dozen row loads in flight)
                Vec16<real> o, i, pq;
#pragma unroll
                for (int k = 0; k < NPL; ++k) o.v[k] = i.v[k] = pq.v[k] = real(0);
                if (j0 < p.strideD) {
                    if (p.disk_mode != 2) o = ldg16(rowO + j0);
                    if (p.disk_mode != 1) i = ldg16(rowI + j0);
                    if (p.disk_mode != 1 && cs17) pq = ldg16(rowP + j0);
                }
#pragma unroll
                for (int k = 0; k < NPL; ++k) {
                    real v = a_d * o.v[k] + c1 * i.v[k]; // absent terms are zero
                    if (p.disk_mode != 1 && cs17) v += c2 * pq.v[k];
                    if (p.apply_igm && j0 + k < p.igmD[3])
                        v *= igm_factor<real>(j0 + k, p.igmD[0], p.igmD[1], p.igmD[2], p.igmD[3], g0f, g1f, g2f);
                    __stcg(SD + (size_t)(j0 + k) * kLanes + lane, j0 + k < nD ? v : real(0));
                }
            }
        } else {
            const real *rowB = rB + (size_t)rb * p.strideB;
#pragma unroll 4
            for (int j0 = s0; j0 <= s1; j0 += NPL) {
                Vec16<real> o;
#pragma unroll
                for (int k = 0; k < NPL; ++k) o.v[k] = real(0);
                if (j0 < p.strideB) o = ldg16(rowB + j0);
#pragma unroll
                for (int k = 0; k < NPL; ++k) {
                    real v = o.v[k];
                    if (p.apply_igm && j0 + k < p.igmB[3])
                        v *= igm_factor<real>(j0 + k, p.igmB[0], p.igmB[1], p.igmB[2], p.igmB[3], g0f, g1f, g2f);
                    __stcg(SB + (size_t)(j0 + k) * kLanes + lane, j0 + k < nB ? v : real(0));
                }
            }
        }
        // emission lines (add_emission_line, src/egg-gencat.cpp:2078-2095): every lane deposits the lines of its own
        // galaxy, in wavelength order, on the nodes of the segment (its own stores: no synchronisation needed)
        const float *ll = disk ? p.line_lum_disk : p.line_lum_bulge;
        if (nline > 0 && ll != nullptr) {
            const double *x = disk ? p.xD : p.xB;
            const int n = disk ? nD : nB;
            const int lo = s0, hi = min(s1, n - 1);
            if (hi < lo) continue;
            const uint16_t *lut = disk ? p.lutD : p.lutB;
            const int nlut = disk ? p.lutD_n : p.lutB_n, e0 = disk ? p.lutD_e0 : p.lutB_e0;
            real *S = (disk ? SD : SB) + lane;
            // the bulge SED carries no mass scale, so its line term is divided by it
            const double inv = disk ? 1.0 : (a_bulge > 0 ? 1.0 / a_bulge : 0.0);
            // a line reaches node j only if x[j+1] > l0 - 10 lw and x[j-1] <= l0 + 10 lw
            const double xa = lo > 0 ? __ldg(x + lo - 1) : -1e300, xb = hi < n - 1 ? __ldg(x + hi + 1) : 1e300;
            for (int l = 0; l < nline; ++l) {
                const double l0 = (double)p.line_lambda[l], lw = l0 * vd;
                if (!(xb > l0 - 10 * lw && xa <= l0 + 10 * lw)) continue;
                const float lum = __ldg(ll + g * nline + p.line_col[l]);
                if (lum == 0.0f) continue;
                // [vif] bounds (:2081-2084)
                const int blo = lut_last_le(x, n, lut, nlut, lut_pos(l0 - 10 * lw, e0), l0 - 10 * lw);
                const int bhi = lut_last_le(x, n, lut, nlut, lut_pos(l0 + 10 * lw, e0), l0 + 10 * lw) + 1; // first > b, n if none
                if (bhi == 0 || blo == n - 1) continue;
                const int j0 = max(blo < 0 ? 0 : blo, lo), j1 = min(bhi >= n ? n - 1 : bhi, hi);
                const double amp = (double)lum * l0;
                for (int j = j0; j <= j1; ++j) {
                    const double xj = __ldg(x + j);
                    const double xm = j > 0 ? __ldg(x + j - 1) : 0.0, xp = j < n - 1 ? __ldg(x + j + 1) : 0.0;
                    const double llow = j != 0 ? 0.5 * (xm + xj) : xj - 0.5 * (xp - xj);     // :2090
                    const double lup = j != n - 1 ? 0.5 * (xj + xp) : xj + 0.5 * (xj - xm); // :2091
                    real *sp = S + (size_t)j * kLanes;
                    __stcg(sp, __ldcg(sp) + (real)(xj * line_bin_t<real>(llow, lup, l0, lw, amp, p.line_average) * inv));
                }
            }
        }
    }
}

// One persistent CTA per SM.  Dynamic shared memory: the band tables of the CTA's current group, then {x, 1 / dx} of every
// node of the disk grid.
template <typename real>
__global__ void __launch_bounds__(512, 1) lane_walk_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_tab;
    __shared__ int s_next;

    const int tid = threadIdx.x, lane = tid & 31;
    const int ngroups = (int)p.ngroups;
    const GridNode *grid = q.grid;
    const uint32_t tab_sa = smem_u32(smem);
    double2 *gxs = reinterpret_cast<double2 *>(smem + ((p.max_blob + 127) & ~127u));
    const uint32_t gx_sa = smem_u32(gxs);
    int16_t *ons = reinterpret_cast<int16_t *>(gxs + q.rowsD + 2); // index in the bulge grid of every disk node, -1 if not optical
    const uint32_t on_sa = smem_u32(ons);
    for (int j = tid; j < (int)q.rowsD + 2; j += (int)blockDim.x) {
        const bool in = j < (int)q.rowsD;
        gxs[j] = in ? make_double2(grid[j].x, grid[j].idxp) : make_double2(0.0, 0.0);
        ons[j] = (int16_t)(in ? grid_o(grid[j]) : -1);
    }
    if (tid == 0) {
        mbar_init(&bar_tab, 1);
        fence_mbar_init();
        // the CTAs are dealt to the groups in proportion to the groups' work
        float tot = 0.f;
        for (int k = 0; k < ngroups; ++k) tot += q.group_cost[k];
        const float x = ((float)blockIdx.x + 0.5f) / (float)gridDim.x * tot;
        float cum = 0.f;
        int grp = ngroups - 1;
        for (int k = 0; k < ngroups; ++k) {
            cum += q.group_cost[k];
            if (x < cum) { grp = k; break; }
        }
        s_next = grp;
    }
    __syncthreads();
    int grp = s_next;
    uint32_t tab_phase = 0;
    while (grp >= 0) {
        if (tid == 0) {
            fence_proxy_async(); // earlier generic-proxy reads of the table region precede the async writes
            const uint32_t bytes = p.groups[grp].bytes;
            mbar_expect_tx(&bar_tab, bytes);
            for (uint32_t o = 0; o < bytes; o += 32768u)
                bulk_g2s(smem + o, p.groups[grp].blob + o, min(32768u, bytes - o), &bar_tab);
        }
        mbar_wait(&bar_tab, tab_phase);
        tab_phase ^= 1;
        const uint32_t nb = p.groups[grp].nb, gbase = p.groups[grp].base;
        const uint32_t nitems = q.nbatch * nb;
        for (;;) {
            // items of a group: (batch, table), tables fastest (neighbouring warps read the same SED rows at about the same time)
            uint32_t it = 0;
            if (lane == 0) it = atomicAdd(q.counters + grp, 1u);
            it = __shfl_sync(kFull, it, 0);
            if (it >= nitems) break;
            const uint32_t bl = it / nb, bi = it - bl * nb, fb = gbase + bi;
            const uint32_t w = __ldg(q.win + (size_t)bl * kWinStride + fb);
            if (!w) continue;
            int ng;
            const uint64_t g = lane_galaxy(q, bl, lane, ng);
            const double opz = 1.0 + (double)__ldg(p.z + g);
            const int jlo = (int)(w & 0x7fffu), jhi = (int)((w >> 15) & 0x7fffu);
            const bool bulge = (w & 0x80000000u) != 0;
            const BandHeader2 &h = reinterpret_cast<const BandHeader2 *>(smem)[bi];
            BandDev B;
            B.hdr = tab_sa + bi * (uint32_t)sizeof(BandHeader2);
            B.rec = tab_sa + h.rec_off; B.hs = tab_sa + h.hs_off; B.bkt = tab_sa + h.bkt_off;
            B.bk_scale = h.bk_scale; B.bk_c0 = h.bk_c0;
            B.nbkt_m1 = (int)h.nbkt - 1;
            B.nsplit = (int)h.nsplit;
            B.nsa0 = h.nsplit > 0 ? B.rec + 32u * h.msplit[0] : 0xffffffffu;
            const bool two_probe = h.two_probe != 0;
            int js, je;
            block_arrivals(grid, q.next_opt, bulge, jlo, jhi, js, je);
            js &= ~(kSub - 1);
            je |= kSub - 1;
            const real *SD = reinterpret_cast<const real *>(reinterpret_cast<const uint8_t *>(q.rows) + (size_t)bl * q.rows_stride) + lane;
            const real *SB = SD + (size_t)q.rowsD * kLanes;
            double totd, totb;
            if (bulge) {
                if (two_probe) walk_table<real, true, true>(B, grid, gx_sa, on_sa, opz, js, je, SD, SB, totd, totb);
                else walk_table<real, false, true>(B, grid, gx_sa, on_sa, opz, js, je, SD, SB, totd, totb);
            } else {
                if (two_probe) walk_table<real, true, false>(B, grid, gx_sa, on_sa, opz, js, je, SD, SB, totd, totb);
                else walk_table<real, false, false>(B, grid, gx_sa, on_sa, opz, js, je, SD, SB, totd, totb);
            }
            double *res = q.res + ((size_t)bl * q.ntab + fb) * 2 * kLanes + lane;
            __stcg(res + kLanes, totd);
            __stcg(res, totb);
        }
        __syncthreads(); // every warp is done with the group's tables
        if (tid == 0) {
            // the group with the most work left (items not yet handed out, weighted by the group's cost per item)
            float best = 0.f;
            int nxt = -1;
            for (int k = 0; k < ngroups; ++k) {
                const uint32_t nbk = p.groups[k].nb, tot = q.nbatch * nbk;
                const uint32_t done = *reinterpret_cast<volatile uint32_t *>(q.counters + k);
                if (k == grp || done >= tot || nbk == 0) continue;
                const float left = (float)(tot - done) * q.group_cost[k] / (float)nbk;
                if (left > best) { best = left; nxt = k; }
            }
            s_next = nxt;
        }
        __syncthreads();
        grp = s_next;
    }
}

// A warp per batch: rest-frame magnitudes; table sums added per output column in table order, data conditions, totals
// (src/egg-gencat.cpp:2238-2239), catalog columns.
template <typename real>
__global__ void __launch_bounds__(256) lane_finish_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    const int lane = threadIdx.x & 31;
    const uint32_t bl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (bl >= q.nbatch) return;
    const int nbt = (int)p.nband_total, ntab = (int)q.ntab, nrf = p.do_rf ? (int)p.nrf : 0;
    const int nD = p.nD, nB = p.nB;
    const bool has_bulge = p.has_bulge != 0;
    int ng;
    const uint64_t g = lane_galaxy(q, bl, lane, ng);
    const double opz = 1.0 + (double)__ldg(p.z + g);
    const double a_bulge = has_bulge ? exp10((double)(__ldg(p.m_bulge + g) - __ldg(p.m2lcor + g))) : 0.0;
    // rest-frame magnitudes: dot products with the z = 0 weights (src/egg-gencat.cpp:2150-2158), [vif]
    // lsun2uJy(0, 1e-5, ...) then uJy2mag; non-finite -> +99 (:2152-2156)
    if (nrf > 0) {
        const real *SD = reinterpret_cast<const real *>(reinterpret_cast<const uint8_t *>(q.rows) + (size_t)bl * q.rows_stride);
        const real *SB = SD + (size_t)q.rowsD * kLanes;
        for (int r = 0; r < nrf; ++r) {
            const double arf = EGGPHOT_LSUN2UJY / (1e-5 * 1e-5);
            double mg[2] = {0.0, 0.0}; // the bulge column stays 0 when there is no bulge pass (no_stellar)
            for (int comp = has_bulge ? 0 : 1; comp < 2; ++comp) {
                const RfDesc dsc = comp ? p.rfD[r] : p.rfB[r];
                const real *S = (comp ? SD : SB) + lane;
                const real *W = reinterpret_cast<const real *>(p.rfW) + dsc.off;
                CompSum<real> acc;
                acc.clear();
                for (int k = 0; k < dsc.n; ++k) acc.add(__ldcg(S + (size_t)(dsc.j0 + k) * kLanes) * __ldg(W + k));
                const double fl = dsc.covered ? ((double)acc.s + (double)acc.c) * (comp ? 1.0 : a_bulge) : __longlong_as_double(0x7ff8000000000000LL);
                const double m = -2.5 * log10(fl * arf) + 23.9;
                mg[comp] = isfinite(m) ? m : 99.0;
            }
            if (lane < ng) {
                const size_t oo = (size_t)g * nrf + r;
                const float md = (float)mg[1], mb = (float)mg[0];
                if (p.rfmag_disk) __stcs(p.rfmag_disk + oo, md);
                if (p.rfmag_bulge) __stcs(p.rfmag_bulge + oo, mb);
                if (p.rfmag) // uJy2mag(mag2uJy(disk) + mag2uJy(bulge)) on the f32 columns (src/egg-gencat.cpp:2239)
                    __stcs(p.rfmag + oo, (float)(-2.5 * log10(exp10(0.4 * (23.9 - (double)md)) + exp10(0.4 * (23.9 - (double)mb))) + 23.9));
                if (p.rfmag_disk_f64) __stcs(p.rfmag_disk_f64 + oo, mg[1]);
                if (p.rfmag_bulge_f64) __stcs(p.rfmag_bulge_f64 + oo, mg[0]);
            }
        }
    }
    if (lane >= ng || nbt == 0) return;
    const double d = __ldg(p.d + g);
    const double sca = EGGPHOT_LSUN2UJY * opz / (d * d) / opz; // K (1+z) / d^2, and 1 / (1+z) of the walk's units
    const bool covd_any = nD >= 2, covb_any = has_bulge && nB >= 2;
    const double xd0 = covd_any ? __ldg(p.xD) * opz : 0.0, xd1 = covd_any ? __ldg(p.xD + nD - 1) * opz : 0.0;
    const double xb0 = covb_any ? __ldg(p.xB) * opz : 0.0, xb1 = covb_any ? __ldg(p.xB + nB - 1) * opz : 0.0;
    const uint32_t *win = q.win + (size_t)bl * kWinStride;
    const double *res = q.res + (size_t)bl * q.ntab * 2 * kLanes + lane;
    for (int col = 0; col < nbt; ++col) {
        double sd = 0.0, sb = 0.0;
        for (int fb = 0; fb < ntab; ++fb)
            if ((int)__ldg(q.tab_col + fb) == col && __ldg(win + fb) != 0u) {
                sd += __ldcg(res + (size_t)(fb * 2 + 1) * kLanes);
                sb += __ldcg(res + (size_t)(fb * 2) * kLanes);
            }
        const double2 cv = __ldg(q.col_span + col);
        // [vif] sed2flux: NaN (-> flux 0, src/egg-gencat.cpp:2198) unless the SED covers the filter
        const bool covd = covd_any && xd0 <= cv.x && xd1 >= cv.y;
        const bool covb = covd && covb_any && xb0 <= cv.x && xb1 >= cv.y;
        double fd = covd ? sd * sca : 0.0;
        double fbv = covb ? sb * sca * a_bulge : 0.0;
        if (!isfinite(fd)) fd = 0.0;
        if (!isfinite(fbv)) fbv = 0.0;
        const float fdf = (float)fd, fbf = (float)fbv;
        const size_t oo = (size_t)g * nbt + col;
        // streaming stores: the catalog columns are written once
        if (p.flux_disk) __stcs(p.flux_disk + oo, fdf);
        if (p.flux_bulge) __stcs(p.flux_bulge + oo, fbf);
        if (p.flux) __stcs(p.flux + oo, fdf + fbf); // vec2f sum
        if (p.flux_disk_f64) __stcs(p.flux_disk_f64 + oo, fd);
        if (p.flux_bulge_f64) __stcs(p.flux_bulge_f64 + oo, fbv);
    }
}

size_t walk_smem_bytes(const LaneParams &q) { return ((q.fp.max_blob + 127) & ~size_t(127)) + ((size_t)q.rowsD + 2) * 18 + 16; }

template <typename real> int launch_chunk(const LaneParams &q, int num_sms, cudaStream_t st) {
    if (q.nbatch == 0) return 0;
    cudaError_t e = cudaMemsetAsync(q.counters, 0, kMaxGroups * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
    lane_plan_kernel<<<(q.nbatch + 7) / 8, 256, 0, st>>>(q);
    lane_build_kernel<real><<<q.nbatch, 512, 0, st>>>(q);
    if (q.fp.ngroups > 0 && q.ntab > 0) {
        const size_t smem = walk_smem_bytes(q);
        e = cudaFuncSetAttribute(lane_walk_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        lane_walk_kernel<real><<<num_sms, 512, smem, st>>>(q);
    }
    lane_finish_kernel<real><<<(q.nbatch + 7) / 8, 256, 0, st>>>(q);
    return (int)cudaGetLastError();
}

} // namespace

size_t lane_kernel_smem(const LaneParams &q) { return walk_smem_bytes(q); }

int lane_kernel_max_blocks_per_sm(const LaneParams &q, int precision) {
    const size_t smem = walk_smem_bytes(q);
    int nb = 0;
    cudaError_t e;
    if (precision == EGGPHOT_FP64) {
        e = cudaFuncSetAttribute(lane_walk_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lane_walk_kernel<double>, 512, smem);
    } else {
        e = cudaFuncSetAttribute(lane_walk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lane_walk_kernel<float>, 512, smem);
    }
    return e == cudaSuccess ? nb : -(int)e;
}

int launch_lane_chunk(const LaneParams &q, int precision, int num_sms, cudaStream_t stream) {
    return precision == EGGPHOT_FP64 ? launch_chunk<double>(q, num_sms, stream) : launch_chunk<float>(q, num_sms, stream);
}

int launch_zsort(const float *z, uint32_t ngal, uint32_t *hist, uint32_t *perm, int num_sms, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(hist, 0, kZBuckets * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
    if (ngal == 0) return 0;
    const uint32_t want = (ngal + 255) / 256, cap = (uint32_t)num_sms * 8;
    const int blocks = (int)(want < cap ? want : cap);
    zsort_count<<<blocks, 256, 0, st>>>(z, ngal, hist);
    zsort_scan<<<1, 1024, 0, st>>>(hist);
    zsort_scatter<<<blocks, 256, 0, st>>>(z, ngal, hist, perm);
    return (int)cudaGetLastError();
}
size_t zsort_hist_bytes() { return kZBuckets * sizeof(uint32_t); }

// TFLOP/s (2 flops per FMA) of back-to-back independent FMA chains on every SM, best of five runs
template <typename T> static int fma_peak(int num_sms, void *scratch, cudaStream_t st, double &tflops) {
    const int iters = sizeof(T) == 4 ? 4096 : 1024, blocks = num_sms * 2;
    cudaEvent_t e0, e1;
    cudaError_t e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e != cudaSuccess) return (int)e;
    double best = 0.0;
    for (int rep = 0; rep < 6 && e == cudaSuccess; ++rep) {
        cudaEventRecord(e0, st);
        fma_peak_kernel<T><<<blocks, 1024, 0, st>>>((T *)scratch, iters, (T)0.999999, (T)1e-7);
        cudaEventRecord(e1, st);
        e = cudaEventSynchronize(e1);
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 64.0 * iters * 1024.0 * blocks;
        if (rep > 0 && ms > 0) best = fmax(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    tflops = best;
    return (int)e;
}
int measure_fma_peak(int num_sms, void *scratch, cudaStream_t st, double *fp32_tflops, double *fp64_tflops) {
    int e = fma_peak<float>(num_sms, scratch, st, *fp32_tflops);
    if (e == 0) e = fma_peak<double>(num_sms, scratch, st, *fp64_tflops);
    return e;
}

} // namespace eggphot
