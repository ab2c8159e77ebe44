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
// Counting sort by log2(1+z) in [0, 8): 2^16 buckets ((1+z) resolved to 8.5e-5) for small catalogs, up to 2^22 for large
// ones (about two galaxies per bucket: the lanes of a batch then sit as close in redshift as the catalog allows).
// Anything else (NaN, z < 0, z > 255) goes to the ends.
constexpr int kZBucketsMin = 1 << 16, kZBucketsMax = 1 << 22;
constexpr int kScanBlock = 2048; // counters per block of the scan
__device__ __forceinline__ int zbucket(float z, int nbuckets) {
    const float v = __log2f(1.0f + fmaxf(z, 0.0f)) * (float)(nbuckets / 8);
    int k = (int)v;
    return k < 0 ? 0 : (k > nbuckets - 1 ? nbuckets - 1 : k);
}
__global__ void zsort_count(const float *__restrict__ z, uint32_t n, uint32_t *__restrict__ hist, int nbuckets) {
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x)
        atomicAdd(&hist[zbucket(__ldg(z + g), nbuckets)], 1u);
}
// exclusive prefix sums inside blocks of 2048 counters (256 threads, eight consecutive counters each); the block totals go
// to bsum
__global__ void __launch_bounds__(256) zsort_scan_blocks(uint32_t *__restrict__ hist, uint32_t *__restrict__ bsum) {
    __shared__ uint32_t wsum[8];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    uint4 *h4 = reinterpret_cast<uint4 *>(hist + (size_t)blockIdx.x * kScanBlock) + 2 * t;
    uint4 a = h4[0], b = h4[1];
    const uint32_t c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k];
    uint32_t inc = s; // inclusive scan over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int k = 0; k < w; ++k) base += wsum[k];
    if (t == 255) bsum[blockIdx.x] = base + inc;
    uint32_t run = base + inc - s;
    uint32_t e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { e[k] = run; run += c[k]; }
    h4[0] = make_uint4(e[0], e[1], e[2], e[3]);
    h4[1] = make_uint4(e[4], e[5], e[6], e[7]);
}
// exclusive prefix sums of the block totals (at most 2048: one block of 1024 threads, two each)
__global__ void __launch_bounds__(1024) zsort_scan_totals(uint32_t *__restrict__ bsum, int nblk) {
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;
    const uint32_t c0 = 2 * t < nblk ? bsum[2 * t] : 0u, c1 = 2 * t + 1 < nblk ? bsum[2 * t + 1] : 0u;
    part[t] = c0 + c1;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const uint32_t v = t >= o ? part[t - o] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const uint32_t run = part[t] - (c0 + c1);
    if (2 * t < nblk) bsum[2 * t] = run;
    if (2 * t + 1 < nblk) bsum[2 * t + 1] = run + c0;
}
__global__ void zsort_scatter(const float *__restrict__ z, uint32_t n, uint32_t *__restrict__ cursor, const uint32_t *__restrict__ bsum,
                              uint32_t *__restrict__ perm, int nbuckets) {
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        const int k = zbucket(__ldg(z + g), nbuckets);
        perm[__ldg(bsum + k / kScanBlock) + atomicAdd(&cursor[k], 1u)] = g;
    }
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
template <typename real> __device__ __forceinline__ real lds_real(uint32_t a);
template <> __device__ __forceinline__ float lds_real<float>(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); // (volatile: the staging buffers are rewritten by bulk copies)
    return v;
}
template <> __device__ __forceinline__ double lds_real<double>(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
// mbarrier / bulk copy by 32-bit shared address
__device__ __forceinline__ void mbar_expect_tx_sa(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_sa(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s_sa(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}


// Walks nodes js..je (js a multiple of 4, je + 1 too) of ONE component grid for one galaxy (lane) and one band table:
// the disk on the merge_add grid, the bulge on the optical grid (its SED intervals run between optical nodes, as when
// get_flux runs on the bulge alone, src/egg-gencat.cpp:2206-2209).
//  * The SED rows [node][lane] of the walk are staged in shared memory by TMA bulk copies, eight rows (1 KB in fp32) per
//    copy, two buffers per warp: lane 0 requests the next eight rows before the warp starts on the current ones; an
//    mbarrier per buffer says when they have landed.  (The rows live in HBM by the time the bands read them.)
//  * Per strip of four arrivals: (1) the filter cell of every node is looked up - four independent chains of
//    shared-memory loads, no branch - and (2) the nodes are taken in order: records of the cell, the two expansions, the
//    interval that ends at the node, the weight of the node before it.
// Returns the sum over the strips of 16 arrivals (each summed in `real`, the strips added in double, in order: the same
// number whichever way the window is cut).  gx = {x, 1 / dx} of the grid: shared address (GXS) or global pointer.
template <typename real, bool TWOP, bool GXS>
__device__ __forceinline__ double walk_comp(const BandDev &B, uint32_t gx_sa, const double2 *__restrict__ gxg, double opz, int js, int je,
                                            const real *__restrict__ rows, uint32_t stage_sa, uint32_t bar_sa, uint32_t &phases) {
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kRowBytes = kLanes * sizeof(real), kStageBytes = kStage * kRowBytes;
    // x of node j for the cell lookups, 1 / dx for the intervals: one 8-byte load each (a warp-uniform 16-byte load costs two
    // wavefronts of the shared-memory pipe, and the walk is bound by that pipe)
    auto load_x = [&](int j) -> double { return GXS ? lds_f64(gx_sa + 16u * (uint32_t)j) : __ldg(&gxg[j].x); };
    auto load_idx = [&](int j) -> double { return GXS ? lds_f64(gx_sa + 16u * (uint32_t)j + 8u) : __ldg(&gxg[j].y); };
    double Pu = 0.0, PLout = 0.0, Eprev = 0.0, tot = 0.0; // state of the previous node
    uint32_t Pra = 0u, Pnsa = B.nsa0;
    real Sprev = real(0), acc = real(0);
    const char *src = reinterpret_cast<const char *>(rows) + (size_t)js * kRowBytes;
    __syncwarp(); // every lane is done with the buffers (the previous walk of this warp)
    if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx_sa(bar_sa, kStageBytes);
        bulk_g2s_sa(stage_sa, src, kStageBytes, bar_sa);
    }
    uint32_t buf = 0;
#pragma unroll 1
    for (int jg = js; jg <= je; jg += kStage, buf ^= 1u) {
        __syncwarp(); // every lane is done with the other buffer
        if (lane == 0 && jg + kStage <= je) {
            src += kStageBytes;
            fence_proxy_async();
            mbar_expect_tx_sa(bar_sa + 8u * (buf ^ 1u), kStageBytes);
            bulk_g2s_sa(stage_sa + (buf ^ 1u) * kStageBytes, src, kStageBytes, bar_sa + 8u * (buf ^ 1u));
        }
        mbar_wait_sa(bar_sa + 8u * buf, (phases >> buf) & 1u);
        phases ^= 1u << buf;
        const uint32_t srow = stage_sa + buf * kStageBytes + (uint32_t)lane * (uint32_t)sizeof(real);
#pragma unroll
        for (int h = 0; h < kStage / kSub; ++h) {
            const int j0 = jg + h * kSub;
            if (h > 0 && j0 > je) break;
            // (1) filter cells
            uint32_t ra[kSub];
            double tn[kSub]; // observed wavelengths of the strip's nodes: the reference's lam = rlam * (1 + z) (src/egg-gencat.cpp:2161), in double
#pragma unroll
            for (int k = 0; k < kSub; ++k) {
                const double t = opz * load_x(j0 + k);
                int kk = round_lo(fma(t, B.bk_scale, B.bk_c0));
                kk = min(max(kk, 0), B.nbkt_m1);
                const uint32_t e = lds_u16(B.bkt + 2u * (uint32_t)kk);
                uint32_t a = B.rec + 32u * (e & kBktRecMask);
                if (TWOP) {
                    // one probe: the bucket entry says how far to step when t is at or beyond the next node (two records where
                    // that node is a split node or an abrupt end: two records at one wavelength)
                    const double f1 = lds_f64(a + 32u);
                    a += t >= f1 ? 32u + ((e >> 10) & 32u) : 0u;
                } else {
                    while (t >= lds_f64(a + 32u)) a += 32u;
                }
                ra[k] = a;
                tn[k] = t;
            }
            // (2) intervals and weights.  Most strips have, for every lane, each node in a filter cell of its own and no split
            // node under them: then E = (Lambda_in - Lambda_out of the node before) / dt and nothing else is needed.
            bool plain = ra[kSub - 1] <= Pnsa && ra[0] != Pra; // (the cells of a walk never go back: the last one is the reddest)
#pragma unroll
            for (int k = 1; k < kSub; ++k) plain = plain && ra[k] != ra[k - 1];
            if (__all_sync(kFull, plain)) {
#pragma unroll
                for (int k = 0; k < kSub; ++k) {
                    const double2 a0 = lds_v2f64(ra[k]), b0 = lds_v2f64(ra[k] + 16u), a1 = lds_v2f64(ra[k] + 32u), b1 = lds_v2f64(ra[k] + 48u);
                    const double idx = load_idx(j0 + k);
                    const real S = lds_real<real>(srow + (uint32_t)(h * kSub + k) * kRowBytes);
                    const double t = tn[k];
                    const double u = t - a0.x, v = t - a1.x;
                    const double Lin = fma(u, fma(u, b0.y, b0.x), a0.y);
                    const double E = (Lin - PLout) * idx;
                    PLout = fma(v, fma(v, b1.y, b1.x), a1.y);
                    acc = fma(to_real<real>(E - Eprev), Sprev, acc);
                    Eprev = E;
                    Sprev = S;
                    Pu = u;
                }
                Pra = ra[kSub - 1];
            } else if (__all_sync(kFull, ra[kSub - 1] <= Pnsa)) {
                // no split node under the strip: shared cells only
#pragma unroll
                for (int k = 0; k < kSub; ++k) {
                    const Recs R = load_recs(B, ra[k]);
                    const double idx = load_idx(j0 + k);
                    const real S = lds_real<real>(srow + (uint32_t)(h * kSub + k) * kRowBytes);
                    const double t = tn[k];
                    const double u = t - R.a0.x, v = t - R.a1.x;
                    const double Lin = fma(u, fma(u, R.b0.y, R.b0.x), R.a0.y);
                    const double Esame = opz * fma(R.hs * Pu, u, fma(Pu + u, R.b0.y, R.b0.x));
                    const double E = ra[k] == Pra ? Esame : (Lin - PLout) * idx;
                    PLout = fma(v, fma(v, R.b1.y, R.b1.x), R.a1.y);
                    acc = fma(to_real<real>(E - Eprev), Sprev, acc);
                    Eprev = E;
                    Sprev = S;
                    Pu = u; Pra = ra[k];
                }
            } else {
                // a split node under the strip (once or twice per walk)
#pragma unroll 1
                for (int k = 0; k < kSub; ++k) {
                    const uint32_t rak = k == 0 ? ra[0] : (k == 1 ? ra[1] : (k == 2 ? ra[2] : ra[3]));
                    const Recs R = load_recs(B, rak);
                    const double idx = load_idx(j0 + k);
                    const real S = lds_real<real>(srow + (uint32_t)(h * kSub + k) * kRowBytes);
                    const double t = opz * load_x(j0 + k);
                    const double u = t - R.a0.x, v = t - R.a1.x;
                    const double Lin = fma(u, fma(u, R.b0.y, R.b0.x), R.a0.y);
                    const double Lout = fma(v, fma(v, R.b1.y, R.b1.x), R.a1.y);
                    // close [j - 1, j], settle node j - 1
                    double Eadj;
                    uint32_t nsa_c;
                    const double E = interval_dev(B, Pu, PLout, Pra, Pnsa, rak, u, Lin, R.b0.x, R.b0.y, R.hs, t, idx, opz, Eadj, nsa_c);
                    acc = fma(to_real<real>(E - Eprev), Sprev, acc);
                    Eprev = Eadj;
                    Sprev = S;
                    Pu = u; PLout = Lout; Pra = rak; Pnsa = nsa_c;
                }
            }
            // the sum of a strip of four arrivals is formed in `real`, the strips are added in double: strips are cut at
            // absolute node numbers, so the result does not depend on where the walk started
            tot += (double)acc;
            acc = real(0);
        }
    }
    return tot;
}

// ---- the kernels --------------------------------------------------------------------------------------
constexpr int kWinStride = 128; // window words per batch in LaneParams::win: {disk, bulge} of at most 64 band tables

constexpr int kSegRows = 256;   // rows per warp of the build kernel
constexpr int kBuildThreads = 128;
#ifndef EGG_BUILD_MINBLOCKS
#define EGG_BUILD_MINBLOCKS 8   // (64 registers, no spills: 1.358e9 against 1.352e9 with the compiler's own 72; 10 / 12 / 16 spill and lose 2 - 5 %; tuning knobs of the build kernel: resident CTAs per SM the register allocation aims at, row-loop unroll)
#endif
#ifndef EGG_BUILD_UNROLL
#define EGG_BUILD_UNROLL 4
#endif
constexpr int kBuildUnroll = EGG_BUILD_UNROLL;
constexpr int kFinishThreads = 128;

// lane -> galaxy of batch bl of the chunk (lanes beyond the end of the catalog repeat the batch's first galaxy)
__device__ __forceinline__ uint64_t lane_galaxy(const LaneParams &q, uint32_t bl, int lane, int &ng) {
    const uint64_t g0 = (q.batch0 + bl) * kLanes;
    ng = (int)min((uint64_t)kLanes, q.fp.ngal - g0);
    return __ldg(q.perm + g0 + (lane < ng ? lane : 0));
}

// A warp per batch: the node windows of every band table for the batch, on the disk grid and on the optical grid, and
// the node ranges to build.
__global__ void __launch_bounds__(256) lane_plan_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    const int lane = threadIdx.x & 31;
    const uint32_t bl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (bl >= q.nbatch) return;
    const int ntab = (int)q.ntab, nrf = p.do_rf ? (int)p.nrf : 0;
    const int nD = p.nD, nB = p.nB;
    const bool has_bulge = p.has_bulge != 0;
    int ng;
    const uint64_t g = lane_galaxy(q, bl, lane, ng);
    double omin = 1.0 + (double)__ldg(p.z + g), omax = omin;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        omin = fmin(omin, __shfl_xor_sync(kFull, omin, o));
        omax = fmax(omax, __shfl_xor_sync(kFull, omax, o));
    }
    int jn0 = nD, jn1 = -1, on0 = nB, on1 = -1; // disk / optical nodes some band of the batch reads
    for (int r = 0; r < nrf; ++r) {             // the rest-frame bands read fixed node ranges
        if (p.rfD[r].n > 0) { jn0 = min(jn0, (int)p.rfD[r].j0); jn1 = max(jn1, (int)(p.rfD[r].j0 + p.rfD[r].n - 1)); }
        if (has_bulge && p.rfB[r].n > 0) { on0 = min(on0, (int)p.rfB[r].j0); on1 = max(on1, (int)(p.rfB[r].j0 + p.rfB[r].n - 1)); }
    }
    // window of band table fb on a grid: from the last node at or left of the table's blue end for the largest (1+z) of the
    // batch to the first node at or right of its red end for the smallest, with a relative margin (nodes a galaxy does
    // not weigh get the weight 0 exactly); the walk arrives at lo .. hi + 1.  0 = nothing to do, else lo | hi << 15 | 1 << 30
    uint32_t *win = q.win + (size_t)bl * kWinStride;
    for (int fb = lane; fb < kWinStride / 2; fb += 32) {
        uint32_t wd = 0u, wb = 0u;
        if (fb < ntab) {
            int grp = 0;
            while (grp + 1 < (int)p.ngroups && fb >= (int)p.groups[grp + 1].base) ++grp;
            const BandHeader2 *gh = reinterpret_cast<const BandHeader2 *>(p.groups[grp].blob) + (fb - (int)p.groups[grp].base);
            const double ff = __ldg(&gh->full_first), fl = __ldg(&gh->full_last);
            // [vif] sed2flux is NaN (-> flux 0, src/egg-gencat.cpp:2198) unless the SED covers the filter: a band no
            // galaxy of the batch covers is skipped
            const bool any_cov = nD >= 2 && __ldg(p.xD) * omin <= ff && __ldg(p.xD + nD - 1) * omax >= fl;
            if (any_cov && __ldg(&gh->n) != 0) {
                const double xlo = __ldg(&gh->f_first) / omax * (1.0 - 1e-9), xhi = __ldg(&gh->f_last) / omin * (1.0 + 1e-9);
                const int lo = max(bis_last_le(p.xD, nD, xlo), 0), hi = min(bis_first_ge(p.xD, nD, xhi), nD - 1);
                wd = (uint32_t)lo | ((uint32_t)hi << 15) | 0x40000000u;
                jn0 = min(jn0, lo); jn1 = max(jn1, hi + 1);
                if (has_bulge && nB >= 2 && __ldg(p.xB) * omin <= ff && __ldg(p.xB + nB - 1) * omax >= fl) {
                    const int blo = max(bis_last_le(p.xB, nB, xlo), 0), bhi = min(bis_first_ge(p.xB, nB, xhi), nB - 1);
                    wb = (uint32_t)blo | ((uint32_t)bhi << 15) | 0x40000000u;
                    on0 = min(on0, blo); on1 = max(on1, bhi + 1);
                }
            }
        }
        win[2 * fb] = wd;
        win[2 * fb + 1] = wb;
    }
    jn0 = __reduce_min_sync(kFull, jn0); jn1 = __reduce_max_sync(kFull, jn1);
    on0 = __reduce_min_sync(kFull, on0); on1 = __reduce_max_sync(kFull, on1);
    if (lane == 0) {
        // rows to build: whole strips of four arrivals (rows a walk touches beyond the nodes a band weighs must hold finite
        // values: their weights are exactly 0); rows at or beyond the end of a grid are zero
        int4 pl = make_int4(0, -1, 0, -1);
        if (jn1 >= jn0) { pl.x = jn0 & ~(kSub - 1); pl.y = min(jn1 | (kSub - 1), (int)q.rowsD - 1); }
        if (has_bulge && on1 >= on0) { pl.z = on0 & ~(kSub - 1); pl.w = min(on1 | (kSub - 1), (int)q.rowsB - 1); }
        q.plan[bl] = pl;
    }
}

// Rows s0..s1 of one component for the 32 galaxies of a warp: v = a * rowA (+ b * rowB (+ c * rowC)), times the IGM factor
// below node igm[3]; 16 bytes of every template row per load, [node][lane] stores.  The template rows are zero from the
// end of the grid to their stride, and rows beyond the stride are written as zeros.
template <typename real, int NROW>
__device__ __forceinline__ void build_rows(real *__restrict__ dst, int s0, int s1, int stride, const real *__restrict__ rowA, real a,
                                           const real *__restrict__ rowB, real b, const real *__restrict__ rowC, real c, bool apply_igm,
                                           const int32_t *igm, real g0f, real g1f, real g2f) {
    constexpr int NPL = 16 / (int)sizeof(real);
    const int igm_end = apply_igm ? igm[3] : 0;
    real *d = dst + (size_t)s0 * kLanes;
#pragma unroll kBuildUnroll
    for (int j0 = s0; j0 <= s1; j0 += NPL, d += NPL * kLanes) { // (unrolled: a dozen row loads in flight)
        real v[NPL];
#pragma unroll
        for (int k = 0; k < NPL; ++k) v[k] = real(0);
        if (j0 < stride) {
            const Vec16<real> A = ldg16(rowA + j0);
#pragma unroll
            for (int k = 0; k < NPL; ++k) v[k] = a * A.v[k];
            if (NROW >= 2) {
                const Vec16<real> B = ldg16(rowB + j0);
#pragma unroll
                for (int k = 0; k < NPL; ++k) v[k] = a * A.v[k] + b * B.v[k];
            }
            if (NROW >= 3) {
                const Vec16<real> C = ldg16(rowC + j0);
#pragma unroll
                for (int k = 0; k < NPL; ++k) v[k] += c * C.v[k];
            }
            if (j0 < igm_end) { // (a few rows in the far UV)
#pragma unroll
                for (int k = 0; k < NPL; ++k)
                    if (j0 + k < igm_end) v[k] *= igm_factor<real>(j0 + k, igm[0], igm[1], igm[2], igm[3], g0f, g1f, g2f);
            }
        }
#pragma unroll
        for (int k = 0; k < NPL; ++k) __stcg(d + k * kLanes, v[k]);
    }
}

// A CTA per batch, a warp per segment of 256 rows: rest-frame SEDs, lane = galaxy, 16 bytes of each template row per
// load, [node][lane] stores; then the emission lines of the segment's nodes.  The bulge SED carries no mass scale
// (applied at the end).
template <typename real>
__global__ void __launch_bounds__(kBuildThreads, (sizeof(real) == 4 ? EGG_BUILD_MINBLOCKS : 6)) lane_build_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    // gridDim.x / nbatch CTAs share a batch: together they have a warp for every segment the grids can have
    const uint32_t per_batch = gridDim.x / q.nbatch, bl = blockIdx.x / per_batch;
    const int lane = threadIdx.x & 31, warp = (int)(blockIdx.x - bl * per_batch) * (kBuildThreads / 32) + (threadIdx.x >> 5);
    const int nwarp = (int)per_batch * (kBuildThreads / 32);
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

    for (int seg = warp; seg < nsegD + nsegB; seg += nwarp) {
        const bool disk = seg < nsegD;
        const int s0 = disk ? pl.x + seg * kSegRows : pl.z + (seg - nsegD) * kSegRows;
        const int s1 = min(s0 + kSegRows - 1, disk ? pl.y : pl.w);
        if (disk) {
            const real *rowO = rO + (size_t)rd * p.strideD, *rowI = rI + (size_t)ri * p.strideD, *rowP = rP + (size_t)ri * p.strideD;
            // a_d * optical + c1 * IR (+ c2 * PAH): the reference's sums in its order (absent terms are absent)
            if (p.disk_mode == 2) {
                if (cs17) build_rows<real, 2>(SD + lane, s0, s1, p.strideD, rowI, c1, rowP, c2, rowP, real(0), p.apply_igm, p.igmD, g0f, g1f, g2f);
                else build_rows<real, 1>(SD + lane, s0, s1, p.strideD, rowI, c1, rowI, real(0), rowI, real(0), p.apply_igm, p.igmD, g0f, g1f, g2f);
            } else if (p.disk_mode == 1) {
                build_rows<real, 1>(SD + lane, s0, s1, p.strideD, rowO, a_d, rowO, real(0), rowO, real(0), p.apply_igm, p.igmD, g0f, g1f, g2f);
            } else if (cs17) {
                build_rows<real, 3>(SD + lane, s0, s1, p.strideD, rowO, a_d, rowI, c1, rowP, c2, p.apply_igm, p.igmD, g0f, g1f, g2f);
            } else {
                build_rows<real, 2>(SD + lane, s0, s1, p.strideD, rowO, a_d, rowI, c1, rowI, real(0), p.apply_igm, p.igmD, g0f, g1f, g2f);
            }
        } else {
            build_rows<real, 1>(SB + lane, s0, s1, p.strideB, rB + (size_t)rb * p.strideB, real(1), rB, real(0), rB, real(0), p.apply_igm, p.igmB,
                                g0f, g1f, g2f);
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

// One persistent CTA per SM.  Dynamic shared memory: the band tables of the CTA's current group, {x, 1 / dx} of every node of
// the two component grids (GXS; read from global memory through L1 when they do not fit), and per warp two staging
// buffers for SED rows with an mbarrier each.
template <typename real, bool GXS>
__global__ void __launch_bounds__(kWalkThreads, 1) lane_walk_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_tab;
    __shared__ __align__(8) uint64_t bar_rows[2 * (kWalkThreads / 32)];
    __shared__ int s_next;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ngroups = (int)p.ngroups;
    const uint32_t tab_sa = smem_u32(smem);
    constexpr uint32_t kStageBytes = kStage * kLanes * sizeof(real);
    uint8_t *after = smem + ((p.max_blob + 127) & ~127u);
    const uint32_t stage_sa = smem_u32(after) + (uint32_t)warp * 2u * kStageBytes;
    after += (size_t)(kWalkThreads / 32) * 2 * kStageBytes;
    double2 *gxs = reinterpret_cast<double2 *>(after);
    const uint32_t gxD_sa = smem_u32(gxs), gxB_sa = gxD_sa + 16u * q.rowsD;
    if (GXS) {
        for (int j = tid; j < (int)q.rowsD; j += (int)blockDim.x) gxs[j] = q.gxD[j];
        for (int j = tid; j < (int)q.rowsB; j += (int)blockDim.x) gxs[q.rowsD + j] = q.gxB[j];
    }
    const uint32_t bar_sa = smem_u32(&bar_rows[2 * warp]);
    if (lane == 0) {
        mbar_init(&bar_rows[2 * warp], 1);
        mbar_init(&bar_rows[2 * warp + 1], 1);
    }
    if (tid == 0) {
        mbar_init(&bar_tab, 1);
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
    fence_mbar_init();
    __syncthreads();
    int grp = s_next;
    uint32_t tab_phase = 0, row_phases = 0;
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
            // Tables are stored heaviest first.  Handing out the heavy tables of every batch before the light ones keeps the
            // items that are in flight when the counter runs out short (a long item started last is the kernel's tail).
            uint32_t bl, bi;
            if (q.item_mode == 2u) { bi = it / q.nbatch; bl = it - bi * q.nbatch; }
            else if (q.item_mode == 1u && nb > 1u) {
                const uint32_t nbh = nb / 2u, first = q.nbatch * nbh;
                if (it < first) { bl = it / nbh; bi = it - bl * nbh; }
                else { const uint32_t i2 = it - first, nbl = nb - nbh; bl = i2 / nbl; bi = nbh + (i2 - bl * nbl); }
            } else { bl = it / nb; bi = it - bl * nb; }
            const uint32_t fb = gbase + bi;
            const uint2 w = __ldg(reinterpret_cast<const uint2 *>(q.win + (size_t)bl * kWinStride) + fb);
            if (!w.x) continue;
            int ng;
            const uint64_t g = lane_galaxy(q, bl, lane, ng);
            const double opz = 1.0 + (double)__ldg(p.z + g);
            const BandHeader2 &h = reinterpret_cast<const BandHeader2 *>(smem)[bi];
            BandDev B;
            B.hdr = tab_sa + bi * (uint32_t)sizeof(BandHeader2);
            B.rec = tab_sa + h.rec_off; B.hs = tab_sa + h.hs_off; B.bkt = tab_sa + h.bkt_off;
            B.bk_scale = h.bk_scale; B.bk_c0 = h.bk_c0;
            B.nbkt_m1 = (int)h.nbkt - 1;
            B.nsplit = (int)h.nsplit;
            B.nsa0 = h.nsplit > 0 ? B.rec + 32u * h.msplit[0] : 0xffffffffu;
            const bool two_probe = h.two_probe != 0;
            const real *SD = reinterpret_cast<const real *>(reinterpret_cast<const uint8_t *>(q.rows) + (size_t)bl * q.rows_stride);
            double tot[2] = {0.0, 0.0}; // bulge, disk
            // the disk on the merge_add grid, then the bulge on the optical grid
#pragma unroll 1
            for (int comp = 1; comp >= 0; --comp) {
                const uint32_t wc = comp ? w.x : w.y;
                if (!wc) continue;
                const int js = (int)(wc & 0x7fffu) & ~(kSub - 1), je = (int)(((wc >> 15) & 0x7fffu) + 1u) | (kSub - 1);
                const real *rows = comp ? SD : SD + (size_t)q.rowsD * kLanes;
                const uint32_t gx_sa = comp ? gxD_sa : gxB_sa;
                const double2 *gxg = comp ? q.gxD : q.gxB;
                tot[comp] = two_probe ? walk_comp<real, true, GXS>(B, gx_sa, gxg, opz, js, je, rows, stage_sa, bar_sa, row_phases)
                                      : walk_comp<real, false, GXS>(B, gx_sa, gxg, opz, js, je, rows, stage_sa, bar_sa, row_phases);
            }
            double *res = q.res + ((size_t)bl * q.ntab + fb) * 2 * kLanes + lane;
            __stcg(res + kLanes, tot[1]);
            __stcg(res, tot[0]);
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

// ---- node sets `sed` and `filter` of [vif] sed2flux ------------------------------------------------------------
// The reference's sources do not say on which nodes vif's sed2flux (src/egg-gencat.cpp:2197) applies the trapezoid rule
// (SURVEY.md 8c).  The contract is the merged set (the walk kernel above); the two other readings are computed here so
// that `egg-gencat-b200 check sed2flux_nodes=...` can try them against a catalog written by the real program:
//   sed    - the SED nodes strictly inside the filter span and the span's two ends, the response interpolated at them;
//   filter - the filter's own nodes, the SED interpolated at them.
// Plain two-cursor walks of the filter as given, lane = galaxy, a warp per (batch, band), everything in double: this is
// a diagnostic path, not a hot one.  Returns (1 + z) times the integral over observed wavelength of R times the rows'
// Y = lambda L, the unit of the walk kernel's sums (the finish kernel applies K / d^2).
__device__ __forceinline__ double lerp2d(double y1, double y2, double x1, double x2, double x) { return y1 + (y2 - y1) * (x - x1) / (x2 - x1); }
template <typename real>
__device__ double alt_nodes_sum(int mode, int nf, const double *__restrict__ fl, const double *__restrict__ fr, const double *__restrict__ x, int n,
                                double opz, const real *__restrict__ S) {
    auto lam = [&](int j) -> double { return __ldg(x + j) * opz; }; // lam = rlam * (1 + z), src/egg-gencat.cpp:2161
    auto sed = [&](int j) -> double { return (double)__ldcg(S + (size_t)j * kLanes); };
    if (nf < 2 || n < 2) return 0.0;
    const double f0 = __ldg(fl), xend = __ldg(fl + nf - 1);
    if (!(lam(0) <= f0 && lam(n - 1) >= xend)) return 0.0; // not covered: [vif] sed2flux is NaN, the flux 0 (:2198)
    int k; // last node at or below the filter's first wavelength
    {
        int lo = 0, hi = n;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (lam(mid) <= f0) lo = mid; else hi = mid;
        }
        k = lo;
    }
    double acc = 0.0;
    if (mode == EGGPHOT_NODES_FILTER) {
        double pg = 0.0, px = f0;
        for (int i = 0; i < nf; ++i) {
            const double fi = __ldg(fl + i);
            while (k + 2 < n && lam(k + 1) <= fi) ++k;
            const double g = __ldg(fr + i) * lerp2d(sed(k), sed(k + 1), lam(k), lam(k + 1), fi);
            if (i > 0) acc += 0.5 * (pg + g) * (fi - px);
            pg = g; px = fi;
        }
        return acc * opz;
    }
    if (k >= n - 1) k = n - 2;
    int fi = 0;
    double px = f0;
    double pg = __ldg(fr) * lerp2d(sed(k), sed(k + 1), lam(k), lam(k + 1), px);
    int si = k + 1;
    while (si < n && lam(si) <= px) ++si;
    while (px < xend) {
        const double xs = si < n ? lam(si) : __longlong_as_double(0x7ff0000000000000LL);
        const bool at_sed = xs < xend; // else the walk closes at the filter's last wavelength
        const double xx = at_sed ? xs : xend;
        while (fi + 2 < nf && __ldg(fl + fi + 1) <= xx) ++fi;
        const double r = xx >= xend ? __ldg(fr + nf - 1) : lerp2d(__ldg(fr + fi), __ldg(fr + fi + 1), __ldg(fl + fi), __ldg(fl + fi + 1), xx);
        double sv;
        if (at_sed || xs == xend) sv = sed(si);
        else {
            const int j = max(si - 1, 0), j1 = j + 1 < n ? j + 1 : j;
            sv = j1 == j ? sed(j) : lerp2d(sed(j), sed(j1), lam(j), lam(j1), xx);
        }
        const double g = r * sv;
        acc += 0.5 * (pg + g) * (xx - px);
        px = xx; pg = g;
        if (at_sed || xs == xend) ++si;
    }
    return acc * opz;
}

template <typename real>
__global__ void __launch_bounds__(128) lane_alt_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    const int lane = threadIdx.x & 31;
    const uint32_t item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t nbt = p.nband_total;
    if (item >= q.nbatch * nbt) return;
    const uint32_t bl = item / nbt, col = item - bl * nbt; // (one header-only table per band: table = column)
    const uint2 w = __ldg(reinterpret_cast<const uint2 *>(q.win + (size_t)bl * kWinStride) + col);
    int ng;
    const uint64_t g = lane_galaxy(q, bl, lane, ng);
    const double opz = 1.0 + (double)__ldg(p.z + g);
    const uint32_t o0 = __ldg(q.alt_off + col), o1 = __ldg(q.alt_off + col + 1);
    const real *SD = reinterpret_cast<const real *>(reinterpret_cast<const uint8_t *>(q.rows) + (size_t)bl * q.rows_stride) + lane;
    const real *SB = SD + (size_t)q.rowsD * kLanes;
    double td = 0.0, tb = 0.0;
    if (w.x) td = alt_nodes_sum<real>((int)q.nodes_mode, (int)(o1 - o0), q.alt_fl + o0, q.alt_fr + o0, p.xD, p.nD, opz, SD);
    if (w.y) tb = alt_nodes_sum<real>((int)q.nodes_mode, (int)(o1 - o0), q.alt_fl + o0, q.alt_fr + o0, p.xB, p.nB, opz, SB);
    double *res = q.res + ((size_t)bl * q.ntab + col) * 2 * kLanes + lane;
    __stcg(res + kLanes, td);
    __stcg(res, tb);
}

// A warp per batch: rest-frame magnitudes; table sums added per output column in table order, data conditions, totals
// (src/egg-gencat.cpp:2238-2239), catalog columns.
template <typename real>
__global__ void __launch_bounds__(kFinishThreads) lane_finish_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    // a CTA per batch: its warps share the output columns (and the rest-frame bands) of the batch's 32 galaxies
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kFinishThreads / 32;
    const uint32_t bl = blockIdx.x;
    const int nbt = (int)p.nband_total, nrf = p.do_rf ? (int)p.nrf : 0;
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
        for (int r = warp; r < nrf; r += kWarps) {
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
    for (int col = warp; col < nbt; col += kWarps) {
        double sd = 0.0, sb = 0.0;
        // the tables of the column, in table order (q.tab_col: [nband + 1] offsets, then the lists)
        for (uint32_t k = __ldg(q.tab_col + col), k1 = __ldg(q.tab_col + col + 1); k < k1; ++k) {
            const uint32_t fb = __ldg(q.tab_col + nbt + 1 + k);
            if (__ldg(win + 2 * fb) != 0u) {
                sd += __ldcg(res + (size_t)(fb * 2 + 1) * kLanes);
                sb += __ldcg(res + (size_t)(fb * 2) * kLanes);
            }
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

// shared memory of the walk kernel without / with the grids
template <typename real> size_t walk_smem_base(const LaneParams &q) {
    return ((q.fp.max_blob + 127) & ~size_t(127)) + (size_t)(kWalkThreads / 32) * 2 * kStage * kLanes * sizeof(real);
}
size_t walk_grid_bytes(const LaneParams &q) { return ((size_t)q.rowsD + q.rowsB) * 16; }

template <typename real> int launch_chunk(const LaneParams &q, int num_sms, cudaStream_t st, cudaEvent_t w0, cudaEvent_t w1) {
    if (q.nbatch == 0) return 0;
    cudaError_t e = cudaMemsetAsync(q.counters, 0, kMaxGroups * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
    lane_plan_kernel<<<(q.nbatch + 7) / 8, 256, 0, st>>>(q);
    {
        const uint32_t nseg = (q.rowsD + kSegRows - 1) / kSegRows + (q.rowsB + kSegRows - 1) / kSegRows;
        const uint32_t per_batch = (nseg + kBuildThreads / 32 - 1) / (kBuildThreads / 32);
        lane_build_kernel<real><<<q.nbatch * per_batch, kBuildThreads, 0, st>>>(q);
    }
    if (q.nodes_mode != EGGPHOT_NODES_MERGED) {
        if (q.ntab > 0) lane_alt_kernel<real><<<(q.nbatch * q.ntab + 3) / 4, 128, 0, st>>>(q);
    } else if (q.fp.ngroups > 0 && q.ntab > 0) {
        if (w0) cudaEventRecord(w0, st);
        if (q.gx_in_smem) {
            const size_t smem = walk_smem_base<real>(q) + walk_grid_bytes(q);
            e = cudaFuncSetAttribute(lane_walk_kernel<real, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            lane_walk_kernel<real, true><<<num_sms, kWalkThreads, smem, st>>>(q);
        } else {
            const size_t smem = walk_smem_base<real>(q);
            e = cudaFuncSetAttribute(lane_walk_kernel<real, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            lane_walk_kernel<real, false><<<num_sms, kWalkThreads, smem, st>>>(q);
        }
        if (w1) cudaEventRecord(w1, st);
    }
    lane_finish_kernel<real><<<q.nbatch, kFinishThreads, 0, st>>>(q);
    return (int)cudaGetLastError();
}

} // namespace

size_t lane_kernel_smem(const LaneParams &q, int precision) {
    const size_t base = precision == EGGPHOT_FP64 ? walk_smem_base<double>(q) : walk_smem_base<float>(q);
    return base + (q.gx_in_smem ? walk_grid_bytes(q) : 0);
}
int launch_lane_chunk(const LaneParams &q, int precision, int num_sms, cudaStream_t stream, cudaEvent_t w0, cudaEvent_t w1) {
    return precision == EGGPHOT_FP64 ? launch_chunk<double>(q, num_sms, stream, w0, w1) : launch_chunk<float>(q, num_sms, stream, w0, w1);
}

int launch_zsort(const float *z, uint32_t ngal, uint32_t *hist, uint32_t *perm, int num_sms, cudaStream_t st) {
    int nbuckets = kZBucketsMin;
    while (nbuckets < kZBucketsMax && (uint32_t)nbuckets * 2u <= ngal) nbuckets *= 2;
    const int nblk = nbuckets / kScanBlock;
    uint32_t *bsum = hist + kZBucketsMax;
    cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)nbuckets * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
    if (ngal == 0) return 0;
    const uint32_t want = (ngal + 255) / 256, cap = (uint32_t)num_sms * 8;
    const int blocks = (int)(want < cap ? want : cap);
    zsort_count<<<blocks, 256, 0, st>>>(z, ngal, hist, nbuckets);
    zsort_scan_blocks<<<nblk, 256, 0, st>>>(hist, bsum);
    zsort_scan_totals<<<1, 1024, 0, st>>>(bsum, nblk);
    zsort_scatter<<<blocks, 256, 0, st>>>(z, ngal, hist, bsum, perm, nbuckets);
    return (int)cudaGetLastError();
}
size_t zsort_hist_bytes() { return ((size_t)kZBucketsMax + kZBucketsMax / kScanBlock) * sizeof(uint32_t); }

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
