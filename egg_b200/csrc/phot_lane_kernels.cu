// phot_lane_kernels.cu — the lane-per-galaxy flux kernel (sm_100a) and the redshift bucketing that feeds it.
//
// lane_flux_kernel<real> replaces the two get_flux passes and the totals of src/egg-gencat.cpp:2097-2239, like
// flux_kernel of phot_kernels.cu, with the work laid out the other way round: a warp's 32 lanes are 32 GALAXIES of
// neighbouring redshift walking the SAME SED nodes of one band, not 32 nodes of one galaxy.
//   * zsort_* kernels bucket the galaxies of a call by log2(1+z) (65536 buckets, counting sort): batches of 32
//     consecutive entries of the permutation have (1+z) within ~1e-4, so their lanes look up the same or adjacent
//     filter cells (shared-memory broadcasts instead of 32-way gathers), share one node window per band and read one
//     warp-uniform grid record per node.  Results do not depend on the batching: every galaxy's sums are formed by its
//     own lane in an order fixed by absolute node numbers.
//   * Two persistent CTAs per SM (one runs while the other waits at a barrier or for its tables) take batches of 32
//     galaxies.  Phase A: per-lane scalars, one node window per band table for the batch, rest-frame SEDs of the 32
//     disks (get_opt_sed / get_ir_sed / merge_add / IGM, :708-712, :2039-2057, :2118-2140) and bulges, emission lines
//     included (:2078-2095), into a per-CTA scratch laid out [node][lane] (coalesced 128-byte rows; stays in L2).
//     Phase B: per band group, the node records of the group's filters are staged in shared memory by TMA bulk copies
//     (SASS UBLKCP + mbarrier); warps pull (table, block of 64 nodes) items; every lane walks the block for its galaxy
//     (phot_lane.cuh: two expansions per node, weights from their differences; no shuffles, no per-lane windows) and
//     writes its block sums to a slot of a small per-CTA array in L2.  Phase C: slot sums in a fixed order
//     (bit-reproducible), data conditions (uncovered band -> 0, :2198; non-finite magnitude -> 99, :2156), totals
//     (:2238-2239), catalog columns.
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
__device__ __forceinline__ uint64_t evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// scratch stores carry an L2 evict-last hint (the scratch is rewritten every batch and should outlive the streamed data)
__device__ __forceinline__ void st_keep(float *p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_keep(double *p, double v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
// The SED rows of a batch live in HBM by the time the bands read them (16 batches per SM are in flight): every trip of
// the walk asks L2 for the rows it will need kPrefetchRows nodes later, so that the loads themselves hit L2.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// shared-memory loads by 32-bit shared address (all address arithmetic of the walk stays in 32 bits)
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return (int)v;
}
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

// ---- the walk of one band table (device form of walk_block, phot_lane.cuh: same arithmetic, scheduled by hand) -----
// Band constants of a table, addresses as 32-bit shared addresses.
struct BandDev {
    uint32_t rec, hs, bkt, hdr; // shared addresses of the record array, the half-slope array, the bucket table, the header
    double bk_scale, bk_c0;
    int nbkt_m1, m0, nsplit;
    bool two_probe;
};
// what an interval needs from the node at its blue end; nsr = record of the next split node at or after the node's cell
struct NodeSt { double u, Lout; int r, nsr; };

template <typename real> __device__ __forceinline__ real to_real(double v);
template <> __device__ __forceinline__ float to_real<float>(double v) { return (float)v; }
template <> __device__ __forceinline__ double to_real<double>(double v) { return v; }

// The interval from a node in cell pr to a node at t in cell r crosses at least one split: the gauge jumps to add to the
// difference of the expansions, D summed over the crossed splits, and the next split at or after r.  (Returned by value:
// a reference parameter of a function that is not inlined would pin the walk's state in local memory.)
struct Cross { double dadd, dsum; int nsr; };
__device__ __noinline__ Cross cross_splits(uint32_t hb, int nsplit, int pr, int r, double t) {
    Cross c{0.0, 0.0, 0x7fffffff};
    for (int s = nsplit - 1; s >= 0; --s) {
        int ms;
        asm("ld.shared.s32 %0, [%1];" : "=r"(ms) : "r"(hb + (uint32_t)offsetof(BandHeader2, msplit) + 4u * s));
        if (ms >= r) c.nsr = ms;
        if (pr <= ms && r > ms) {
            const double C = lds_f64(hb + (uint32_t)offsetof(BandHeader2, splitC) + 8u * s);
            const double D = lds_f64(hb + (uint32_t)offsetof(BandHeader2, splitD) + 8u * s);
            const double F = lds_f64(hb + (uint32_t)offsetof(BandHeader2, splitF) + 8u * s);
            c.dadd += fma(D, t - F, C);
            c.dsum += D;
        }
    }
    return c;
}

// E of the interval p -> node at t, in the gauge of p's region; Eadj = the same quantity in the gauge of the node's region
__device__ __forceinline__ double interval_dev(const BandDev &B, double pu, double pLout, int pr, int pnsr, int r, double u, double Lin,
                                               double k0, double hr, double hs, double t, double idx, double opz, double &Eadj, int &nsr) {
    nsr = pnsr;
    if (r == pr) {
        const double E = opz * fma(hs * pu, u, fma(pu + u, hr, k0));
        Eadj = E;
        return E;
    }
    const double d = Lin - pLout;
    if (r > pnsr) { // crosses a split (once or twice per band)
        const Cross c = cross_splits(B.hdr, B.nsplit, pr, r, t);
        nsr = c.nsr;
        const double E = (d + c.dadd) * idx;
        Eadj = E - c.dsum * opz;
        return E;
    }
    const double E = d * idx;
    Eadj = E;
    return E;
}

// Walks nodes js..je of the disk grid for one galaxy (lane) and one band table.
// ob0 = first optical node at or after js (index in the bulge grid), obn = optical nodes per 256 disk nodes of the walk
// (only for prefetching the bulge rows ahead of the walk).
// totd / totb receive the sums of the blocks of 64 nodes (each summed in `real`, the blocks added in double, in order:
// the same numbers whichever way the window is cut).
template <typename real>
__device__ __forceinline__ void walk_dev(const BandDev &B, const GridNode *__restrict__ grid, uint32_t gx_sa, double opz, bool bulge,
                                         int js, int je, const real *__restrict__ sd, const real *__restrict__ sb, int ob0, int obn,
                                         double &totd, double &totb) {
    NodeSt s0{0.0, 0.0, -1, B.m0}, s1 = s0, so = s0;
    double Eprev_d = 0.0, Eprev_b = 0.0;
    real Sprev_d = real(0), Sprev_b = real(0), accd = real(0), accb = real(0);
    double td = 0.0, tb = 0.0;
    int o_run = ob0 - 1;   // index in the bulge grid of the last optical node the walk arrived at (they are met in order)
    bool prev_opt = false; // node j - 1 is an optical node
    // one arrival: P = state of node j - 1, C receives the state of node j
    auto step = [&](const NodeSt &P, NodeSt &C, int j) {
        // {x, idxp} from shared memory; the sign of idxp flags an optical node
        const double2 gx = lds_v2f64(gx_sa + 16u * (uint32_t)j);
        const bool optical = bulge && __double2hiint(gx.y) < 0;
        const double t = opz * gx.x; // the reference's lam = rlam * (1 + z) (src/egg-gencat.cpp:2161), in double
        // the bulge interval that ends here starts at the last optical node: the previous node (then it is the disk interval
        // too) or an earlier one, whose state was kept when the walk left it
        const bool own_interval = optical && !prev_opt;
        const double idxpo = own_interval ? __ldg(&grid[j].idxpo) : 0.0;
        if (bulge && !optical && prev_opt) { // (warp-uniform: a property of the grid)
            asm volatile("" ::: "memory");     // keep it a branch: predicated copies would run at every node
            so = P;
        }
        // Nodes outside the band's window need no mask: every interval next to them lies outside the span for every lane of
        // the batch, so their weights are exactly 0 (and the rows the walk can touch are always built: finite values).
        const real Sd = __ldcg(sd + (size_t)j * kLanes);
        real Sb = real(0);
        if (optical) { ++o_run; Sb = __ldcg(sb + (size_t)o_run * kLanes); }
        // filter cell of t
        int k = round_lo(fma(t, B.bk_scale, B.bk_c0));
        k = min(max(k, 0), B.nbkt_m1);
        uint32_t ra = B.rec + 32u * (uint32_t)lds_u16(B.bkt + 2u * (uint32_t)k);
        if (B.two_probe) {
            ra += t >= lds_f64(ra + 32u) ? 32u : 0u;
            ra += t >= lds_f64(ra + 32u) ? 32u : 0u; // a split node or an abrupt end is two records at one wavelength
        } else {
            while (t >= lds_f64(ra + 32u)) ra += 32u;
        }
        const double2 q0a = lds_v2f64(ra), q0b = lds_v2f64(ra + 16u), q1a = lds_v2f64(ra + 32u), q1b = lds_v2f64(ra + 48u);
        const int r = (int)((ra - B.rec) >> 5);
        const double hs = lds_f64(B.hs + 8u * (uint32_t)r);
        const double u = t - q0a.x, v = t - q1a.x;
        const double Lin = fma(u, fma(u, q0b.y, q0b.x), q0a.y);
        C.u = u;
        C.Lout = fma(v, fma(v, q1b.y, q1b.x), q1a.y);
        C.r = r;
        // disk: close [j - 1, j], settle node j - 1
        double Eadj_d;
        int nsr_c;
        const double E_d = interval_dev(B, P.u, P.Lout, P.r, P.nsr, r, u, Lin, q0b.x, q0b.y, hs, t, fabs(gx.y), opz, Eadj_d, nsr_c);
        C.nsr = nsr_c;
        accd = fma(to_real<real>(E_d - Eprev_d), Sprev_d, accd);
        Eprev_d = Eadj_d;
        Sprev_d = Sd;
        // bulge: its SED intervals run between optical nodes; the interval that ends here settles the optical node before
        if (optical) {
            double E_b = E_d, Eadj_b = Eadj_d;
            if (own_interval) { int dummy; E_b = interval_dev(B, so.u, so.Lout, so.r, so.nsr, r, u, Lin, q0b.x, q0b.y, hs, t, idxpo, opz, Eadj_b, dummy); }
            accb = fma(to_real<real>(E_b - Eprev_b), Sprev_b, accb);
            Eprev_b = Eadj_b;
            Sprev_b = Sb;
        }
        prev_opt = optical;
    };
    const int lane = threadIdx.x & 31;
    constexpr int kLinesPerRow = (int)sizeof(real) / 4; // a row is 32 lanes x sizeof(real): one or two 128-byte lines
    const char *sd_rows = reinterpret_cast<const char *>(sd - lane) + 128 * lane, *sb_rows = reinterpret_cast<const char *>(sb - lane) + 128 * lane;
    // the rows of the first strip, one line per lane
    if (lane < kLaneStrip * kLinesPerRow) prefetch_l2(sd_rows + (size_t)js * (kLanes * sizeof(real)));
    if (bulge && ob0 >= 0) prefetch_l2(sb_rows + (size_t)ob0 * (kLanes * sizeof(real)));
    // trips start at odd nodes: a strip of 16 arrivals (16 k, 16 k + 16] is then eight whole trips
    int j = js;
    if ((j & 1) == 0) { step(s0, s1, j); s0 = s1; ++j; }
    while (j <= je) {
        const int jend = (j + kLaneStrip - 1) & ~(kLaneStrip - 1); // last arrival of the strip
        // ask L2 for the rows of the next strip (the SED rows of a batch live in HBM by the time the bands read them)
        if (lane < kLaneStrip * kLinesPerRow) prefetch_l2(sd_rows + (size_t)(jend + 1) * (kLanes * sizeof(real)));
        if (bulge) prefetch_l2(sb_rows + (size_t)(o_run + 1 + kLaneStrip / 2) * (kLanes * sizeof(real)));
        // two arrivals per trip, the two states swapping roles (no register moves); the last trip may run one node past je,
        // which weighs nothing (the grid has two virtual nodes behind the last real one)
#pragma unroll 1
        for (; j <= min(jend, je); j += 2) {
            step(s0, s1, j);
            step(s1, s0, j + 1);
        }
        // the sums of a strip are formed in `real`, the strips added in double: strips are cut at absolute node numbers, so
        // the result does not depend on where the walk started
        td += (double)accd; accd = real(0);
        tb += (double)accb; accb = real(0);
    }
    totd = td;
    totb = tb;
}

// ---- the kernel ---------------------------------------------------------------------------------------
// Per-warp scratch in global memory: SED rows of the disk grid [nD + 2][32], of the bulge grid [nB + 2][32] (reals), then
// the sums of every band table [ntab][2][32] (doubles).
template <typename real, int NT>
__global__ void __maxnreg__((65536 / NT) & ~7) lane_flux_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_tab;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NW = (int)blockDim.x >> 5; // warps of this launch (the host picks the count that fills the last round best)
    const int nbt = (int)p.nband_total, ntab = (int)q.ntab, nrf = p.do_rf ? (int)p.nrf : 0;
    const int nD = p.nD, nB = p.nB, nline = p.nline;
    const uint8_t *blob = smem;
    const bool cs17 = p.ir_kind == EGGPHOT_IR_CS17;
    const bool has_bulge = p.has_bulge != 0;
    uint8_t *wbase = reinterpret_cast<uint8_t *>(p.scratch) + ((size_t)blockIdx.x * (NT / 32) + warp) * q.scratch_stride;
    real *SD = reinterpret_cast<real *>(wbase);
    real *SB = SD + (size_t)(nD + 2) * kLanes;
    double *RES = reinterpret_cast<double *>(wbase + q.res_offset);
    const real *rO = reinterpret_cast<const real *>(p.rowsDopt), *rI = reinterpret_cast<const real *>(p.rowsDir);
    const real *rP = reinterpret_cast<const real *>(p.rowsDpah), *rB = reinterpret_cast<const real *>(p.rowsB);
    const GridNode *grid = q.grid;
    const uint32_t tab_sa = smem_u32(smem);
    // behind the band tables: {x, idxp} of every node of the disk grid, idxp negated at optical nodes
    double2 *gxs = reinterpret_cast<double2 *>(smem + ((p.max_blob + 127) & ~127u));
    const uint32_t gx_sa = smem_u32(gxs);
    for (int j = tid; j < nD + 2; j += (int)blockDim.x) gxs[j] = make_double2(grid[j].x, grid[j].o >= 0 ? -grid[j].idxp : grid[j].idxp);

    if (tid == 0) {
        mbar_init(&bar_tab, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t tab_phase = 0;
    int loaded = -1;

    // batches are dealt round-robin, NW consecutive ones (neighbouring redshifts) per CTA and round: the cost of a batch
    // varies with redshift (bands covered, nodes under them), and every CTA gets its share of every redshift
    const uint64_t nbatch = (p.ngal + kLanes - 1) / kLanes;
    const int nround = (int)((nbatch + (uint64_t)gridDim.x * NW - 1) / ((uint64_t)gridDim.x * NW));
    const int ngl = max((int)p.ngroups, 1);
    for (int round = 0; round < nround; ++round) {
        const uint64_t batch = ((uint64_t)round * gridDim.x + blockIdx.x) * NW + warp;
        const bool active = batch < nbatch;
        const uint64_t g0 = batch * kLanes;
        const int ng = active ? (int)min((uint64_t)kLanes, p.ngal - g0) : 0;
        // lanes beyond the end of the catalog repeat the batch's first galaxy (and write nothing)
        const uint64_t g = active ? q.perm[g0 + (lane < ng ? lane : 0)] : 0;
        double opz = 1.0, omin = 1.0, omax = 1.0;
        if (active) {
            opz = 1.0 + (double)__ldcs(p.z + g);
            omin = opz; omax = opz;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                omin = fmin(omin, __shfl_xor_sync(kFull, omin, o));
                omax = fmax(omax, __shfl_xor_sync(kFull, omax, o));
            }
        }
        // node window of band table fb for the batch: from the last node at or left of the table's blue end for the largest
        // (1+z) to the first node at or right of its red end for the smallest, with a relative margin (nodes a galaxy does
        // not weigh get the weight 0 exactly); 0 = nothing to do.  jlo | jhi << 16 | bulge << 31 (jhi < 32768 * 2 ...)
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
        uint32_t win_lo = 0, win_hi = 0; // lane t: windows of tables t and t + 32

        for (int gstep = 0; gstep < ngl; ++gstep) {
            // consecutive rounds walk the groups in opposite orders: the tables of the group a round ends with are still
            // in shared memory when the next round starts with it (one table load less per round)
            const int grp = (round & 1) ? ngl - 1 - gstep : gstep;
            const int nb = grp < (int)p.ngroups ? (int)p.groups[grp].nb : 0;
            if (nb > 0 && loaded != grp) {
                if (tid == 0) {
                    fence_proxy_async(); // earlier generic-proxy reads of the table region precede the async writes
                    const uint32_t bytes = p.groups[grp].bytes;
                    mbar_expect_tx(&bar_tab, bytes);
                    for (uint32_t o = 0; o < bytes; o += 32768u)
                        bulk_g2s(smem + o, p.groups[grp].blob + o, min(32768u, bytes - o), &bar_tab);
                }
                mbar_wait(&bar_tab, tab_phase);
                tab_phase ^= 1;
                loaded = grp;
            }
            if (active) {
                if (gstep == 0) {
                    // ---- phase A: windows of all tables, rest-frame SEDs of the batch, emission lines, rest-frame magnitudes
                    win_lo = lane < ntab ? table_window(lane) : 0u;
                    win_hi = lane + 32 < ntab ? table_window(lane + 32) : 0u;
                    int jn0 = nD, jn1 = -1, on0 = nB, on1 = -1; // disk / optical nodes some band of the batch reads
                    for (int r = 0; r < nrf; ++r) { // the rest-frame bands read fixed node ranges
                        if (p.rfD[r].n > 0) { jn0 = min(jn0, (int)p.rfD[r].j0); jn1 = max(jn1, (int)(p.rfD[r].j0 + p.rfD[r].n - 1)); }
                        if (has_bulge && p.rfB[r].n > 0) { on0 = min(on0, (int)p.rfB[r].j0); on1 = max(on1, (int)(p.rfB[r].j0 + p.rfB[r].n - 1)); }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t w = h ? win_hi : win_lo;
                        if (w) {
                            const int jlo = (int)(w & 0x7fffu), jhi = (int)((w >> 15) & 0x7fffu);
                            jn0 = min(jn0, jlo); jn1 = max(jn1, jhi);
                            if (w & 0x80000000u) { // optical nodes inside the window
                                const int jo0 = grid[jlo].o >= 0 ? jlo : q.next_opt[jlo], jo1 = grid[jhi].o >= 0 ? jhi : grid[jhi].po;
                                if (jo0 <= jo1 && jo1 >= 0) { on0 = min(on0, grid_o(grid[jo0])); on1 = max(on1, grid_o(grid[jo1])); }
                            }
                        }
                    }
                    jn0 = __reduce_min_sync(kFull, jn0); jn1 = __reduce_max_sync(kFull, jn1);
                    on0 = __reduce_min_sync(kFull, on0); on1 = __reduce_max_sync(kFull, on1);

                    // per-lane scalars (src/egg-gencat.cpp:708-712, 2039-2057)
                    // the reference passes the FLOAT difference m[i]-out.m2lcor[i] to e10 (:2106, :2114)
                    const float m2l = __ldcs(p.m2lcor + g);
                    const double a_bulge = has_bulge ? exp10((double)(p.m_bulge[g] - m2l)) : 0.0;
                    const real a_d = p.disk_mode != 2 ? (real)exp10((double)(p.m_disk[g] - m2l)) : real(0);
                    real c1 = real(0), c2 = real(0);
                    bool bad = false;
                    uint64_t ri = 0, rb = 0, rd = 0;
                    if (p.disk_mode != 1) {
                        ri = p.ir_sed[g];
                        if (ri >= p.nir_sed) { bad = true; ri = 0; }
                        if (cs17) {
                            const double fp = p.fpah[g], md = p.mdust[g];
                            c1 = (real)((1.0 - fp) * md);
                            c2 = (real)(fp * md);
                        } else {
                            c1 = (real)(double)p.lir[g];
                        }
                    }
                    if (has_bulge) {
                        rb = p.opt_sed_bulge[g];
                        if (rb >= p.nopt_sed || !p.opt_use[rb]) { bad = true; rb = 0; }
                    }
                    if (p.disk_mode != 2) {
                        rd = p.opt_sed_disk[g];
                        if (rd >= p.nopt_sed || !p.opt_use[rd]) { bad = true; rd = 0; }
                    }
                    if (bad) atomicExch(p.error_flag, EGGPHOT_ERR_INDEX);
                    real g0f = real(1), g1f = real(1), g2f = real(1);
                    if (p.apply_igm) { g0f = __ldcs(p.igm_abs + 3 * g); g1f = __ldcs(p.igm_abs + 3 * g + 1); g2f = __ldcs(p.igm_abs + 3 * g + 2); }

                    // rest-frame SEDs, lane = galaxy: 16 bytes of each template row per load, [node][lane] stores.  Only the
                    // nodes some band of the batch reads are built.  The bulge SED carries no mass scale (applied at the end).
                    constexpr int NPL = 16 / (int)sizeof(real);
                    // (rows a walk can touch beyond the needed nodes - the node before a window, the optical nodes around it, the
                    //  node an unrolled trip runs over - are built as well: their weights are exactly 0, their values must be finite)
                    if (jn1 >= jn0) {
                        const real *rowO = rO + (size_t)rd * p.strideD, *rowI = rI + (size_t)ri * p.strideD, *rowP = rP + (size_t)ri * p.strideD;
                        const int jb0 = max(jn0 - (int)q.row_margin, 0) & ~3, jb1 = min(jn1 + (int)q.row_margin, nD + 1);
#pragma unroll 4
                        for (int j0 = jb0; j0 <= jb1; j0 += NPL) { // (unrolled: a dozen row loads in flight)
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
                                if (j0 + k <= nD + 1) __stcg(SD + (size_t)(j0 + k) * kLanes + lane, j0 + k < nD ? v : real(0));
                            }
                        }
                    }
                    if (has_bulge && on1 >= on0) {
                        const real *rowB = rB + (size_t)rb * p.strideB;
                        const int jb0 = max(on0 - 2, 0) & ~3, jb1 = min(on1 + 2, nB + 1);
#pragma unroll 4
                        for (int j0 = jb0; j0 <= jb1; j0 += NPL) {
                            Vec16<real> o;
#pragma unroll
                            for (int k = 0; k < NPL; ++k) o.v[k] = real(0);
                            if (j0 < p.strideB) o = ldg16(rowB + j0);
#pragma unroll
                            for (int k = 0; k < NPL; ++k) {
                                real v = o.v[k];
                                if (p.apply_igm && j0 + k < p.igmB[3])
                                    v *= igm_factor<real>(j0 + k, p.igmB[0], p.igmB[1], p.igmB[2], p.igmB[3], g0f, g1f, g2f);
                                if (j0 + k <= nB + 1) __stcg(SB + (size_t)(j0 + k) * kLanes + lane, j0 + k < nB ? v : real(0));
                            }
                        }
                    }
                    __syncwarp();
                    // emission lines (add_emission_line, src/egg-gencat.cpp:2078-2095): every lane deposits the lines of its
                    // own galaxy, in wavelength order, at the nodes some band reads
                    if (nline > 0) {
                        const double vd = (double)p.vdisp[g] / 2.998e5;
                        for (int c = has_bulge && p.line_lum_bulge != nullptr ? 0 : 1; c < 2; ++c) { // 0 bulge, 1 disk
                            const float *ll = c ? p.line_lum_disk : p.line_lum_bulge;
                            if (ll == nullptr) continue;
                            const double *x = c ? p.xD : p.xB;
                            const int n = c ? nD : nB;
                            const uint16_t *lut = c ? p.lutD : p.lutB;
                            const int nlut = c ? p.lutD_n : p.lutB_n, e0 = c ? p.lutD_e0 : p.lutB_e0;
                            const int lo = c ? jn0 : on0, hi = c ? jn1 : on1;
                            real *S = (c ? SD : SB) + lane;
                            // the bulge SED carries no mass scale, so its line term is divided by it
                            const double inv = c ? 1.0 : (a_bulge > 0 ? 1.0 / a_bulge : 0.0);
                            for (int l = 0; l < nline; ++l) {
                                const float lum = ll[g * nline + p.line_col[l]];
                                if (lum == 0.0f) continue;
                                const double l0 = (double)p.line_lambda[l], lw = l0 * vd;
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
                        __syncwarp();
                    }
                    // rest-frame magnitudes: dot products with the z = 0 weights (src/egg-gencat.cpp:2150-2158), [vif]
                    // lsun2uJy(0, 1e-5, ...) then uJy2mag; non-finite -> +99 (:2152-2156)
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

                // ---- phase B: every band table of the group, one after the other, each lane walking its own galaxy
                const int gbase = grp < (int)p.ngroups ? (int)p.groups[grp].base : 0;
                for (int bi = 0; bi < nb; ++bi) {
                    const int fb = gbase + bi;
                    const uint32_t w = __shfl_sync(kFull, fb < 32 ? win_lo : win_hi, fb & 31);
                    double totd = 0.0, totb = 0.0;
                    if (w) {
                        const int jlo = (int)(w & 0x7fffu), jhi = (int)((w >> 15) & 0x7fffu);
                        const bool bulge = (w & 0x80000000u) != 0;
                        const BandHeader2 &h = reinterpret_cast<const BandHeader2 *>(blob)[bi];
                        BandDev B;
                        B.hdr = tab_sa + (uint32_t)bi * (uint32_t)sizeof(BandHeader2);
                        B.rec = tab_sa + h.rec_off; B.hs = tab_sa + h.hs_off; B.bkt = tab_sa + h.bkt_off;
                        B.bk_scale = h.bk_scale; B.bk_c0 = h.bk_c0;
                        B.nbkt_m1 = (int)h.nbkt - 1;
                        B.nsplit = (int)h.nsplit;
                        B.m0 = (int)h.msplit[0];
                        B.two_probe = h.two_probe != 0;
                        int js, je;
                        block_arrivals(grid, q.next_opt, bulge, jlo, jhi, js, je);
                        int ob0 = -1, obn = 0;
                        if (bulge) { // optical nodes the walk visits: for the prefetch of the bulge rows
                            const int jo0 = grid[js].o >= 0 ? js : q.next_opt[js], jo1 = grid[je].o >= 0 ? je : grid[je].po;
                            if (jo0 <= jo1 && jo1 >= 0) {
                                ob0 = grid_o(grid[jo0]);
                                obn = ((grid_o(grid[jo1]) - ob0 + 1) << 8) / (je - js + 1);
                            }
                        }
                        walk_dev<real>(B, grid, gx_sa, opz, bulge, js, je, SD + lane, SB + lane, ob0, obn, totd, totb);
                    }
                    __stcg(RES + (size_t)(fb * 2 + 1) * kLanes + lane, totd);
                    __stcg(RES + (size_t)(fb * 2) * kLanes + lane, totb);
                }

                if (gstep == ngl - 1 && lane < ng) {
                    // ---- phase C: table sums added per output column in table order, data conditions, totals
                    //      (src/egg-gencat.cpp:2238-2239), catalog columns
                    const double d = __ldcs(p.d + g);
                    const double sca = EGGPHOT_LSUN2UJY * opz / (d * d) / opz; // K (1+z) / d^2, and 1 / (1+z) of the walk's units
                    const double a_bulge = has_bulge ? exp10((double)(p.m_bulge[g] - __ldcs(p.m2lcor + g))) : 0.0;
                    const bool covd_any = nD >= 2, covb_any = has_bulge && nB >= 2;
                    const double xd0 = covd_any ? __ldg(p.xD) * opz : 0.0, xd1 = covd_any ? __ldg(p.xD + nD - 1) * opz : 0.0;
                    const double xb0 = covb_any ? __ldg(p.xB) * opz : 0.0, xb1 = covb_any ? __ldg(p.xB + nB - 1) * opz : 0.0;
                    for (int col = 0; col < nbt; ++col) {
                        double sd = 0.0, sb = 0.0;
                        for (int fb = 0; fb < ntab; ++fb)
                            if ((int)__ldg(q.tab_col + fb) == col) {
                                sd += __ldcg(RES + (size_t)(fb * 2 + 1) * kLanes + lane);
                                sb += __ldcg(RES + (size_t)(fb * 2) * kLanes + lane);
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
            }
            __syncthreads(); // every warp is done with the group's tables; the table region may be overwritten
        }
    }
}

#ifndef EGG_LANE_THREADS
#define EGG_LANE_THREADS 512
#endif
// Warps per CTA trade registers per thread (128 at 16 warps, 112 at 18, 104 at 19) against batches in flight; the
// variants are built side by side and EGGPHOT_LANE_THREADS (512 | 576 | 608) selects one.
static int lane_threads() {
    static int nt = [] {
        const char *e = getenv("EGGPHOT_LANE_THREADS");
        const int v = e ? atoi(e) : 0;
        return v == 512 ? 512 : (v == 608 ? 608 : (v == 576 ? 576 : EGG_LANE_THREADS));
    }();
    return nt;
}

size_t lane_smem_bytes(const LaneParams &q) { return ((q.fp.max_blob + 127) & ~size_t(127)) + (size_t)(q.fp.nD + 2) * 16; }

// Every warp takes one batch of 32 galaxies per round and a round lasts as long as its slowest warp: the launch uses the
// fewest warps (not below three quarters of NT / 32) that still need the fewest rounds, so that the last round is as full
// as it can be and the warps share the SM with as few others as possible.
template <typename real, int NT> int launch_lane(const LaneParams &q, int blocks, cudaStream_t st) {
    const size_t smem = lane_smem_bytes(q);
    cudaError_t e = cudaFuncSetAttribute(lane_flux_kernel<real, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const uint64_t nbatch = (q.fp.ngal + kLanes - 1) / kLanes, per_cta = (nbatch + blocks - 1) / blocks;
    int nw = NT / 32;
    const uint64_t rounds = (per_cta + nw - 1) / nw;
    while (nw - 1 >= (NT / 32) * 3 / 4 && (per_cta + nw - 2) / (nw - 1) == rounds) --nw;
    lane_flux_kernel<real, NT><<<blocks, 32 * nw, smem, st>>>(q);
    return (int)cudaGetLastError();
}

template <typename real, int NT> int lane_occupancy(const LaneParams &q) {
    const size_t smem = lane_smem_bytes(q);
    int nb = 0;
    cudaError_t e = cudaFuncSetAttribute(lane_flux_kernel<real, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lane_flux_kernel<real, NT>, NT, smem);
    return e == cudaSuccess ? nb : -(int)e;
}

} // namespace

size_t lane_kernel_smem(const LaneParams &q) { return lane_smem_bytes(q); }

int lane_kernel_max_blocks_per_sm(const LaneParams &q, int precision) {
    const bool d = precision == EGGPHOT_FP64;
    switch (lane_threads()) {
    case 576: return d ? lane_occupancy<double, 576>(q) : lane_occupancy<float, 576>(q);
    case 608: return d ? lane_occupancy<double, 608>(q) : lane_occupancy<float, 608>(q);
    default: return d ? lane_occupancy<double, 512>(q) : lane_occupancy<float, 512>(q);
    }
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

int lane_kernel_threads() { return lane_threads(); }

int launch_lane_flux_kernel(const LaneParams &q, int precision, int grid_blocks, cudaStream_t stream) {
    const bool d = precision == EGGPHOT_FP64;
    switch (lane_threads()) {
    case 576: return d ? launch_lane<double, 576>(q, grid_blocks, stream) : launch_lane<float, 576>(q, grid_blocks, stream);
    case 608: return d ? launch_lane<double, 608>(q, grid_blocks, stream) : launch_lane<float, 608>(q, grid_blocks, stream);
    default: return d ? launch_lane<double, 512>(q, grid_blocks, stream) : launch_lane<float, 512>(q, grid_blocks, stream);
    }
}

} // namespace eggphot
