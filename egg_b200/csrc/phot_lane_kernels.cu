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

struct BandWin { uint32_t w0, w1, w2, w3; }; // jlo | jhi << 16; first block | blocks << 16; item base in its group | bulge << 30; first slot

// ---- the walk of one block (device form of walk_block, phot_lane.cuh: same arithmetic, scheduled by hand) ----------
// Band constants of an item, addresses as 32-bit shared addresses.
struct BandDev {
    uint32_t rec, hs, bkt, hdr; // shared addresses of the record array, the half-slope array, the bucket table, the header
    double bk_scale, bk_c0;
    int nbkt_m1, m0, mlast, nsplit;
    bool two_probe;
};
// what an interval needs from the node at its blue end
struct NodeSt { double u, Lout; int r; };

template <typename real> __device__ __forceinline__ real to_real(double v);
template <> __device__ __forceinline__ float to_real<float>(double v) { return (float)v; }
template <> __device__ __forceinline__ double to_real<double>(double v) { return v; }

// E of the interval p -> node at t with expansions (u, Lin) in cell r whose constants are (k0, hr, hs); returns E in the
// gauge of p's region and sets Eadj to the same quantity in the gauge of the node's region (differs when the interval
// crosses a split)
__device__ __forceinline__ double interval_dev(const BandDev &B, const NodeSt &p, int r, double u, double Lin, double k0, double hr,
                                               double hs, double t, double idx, double opz, double &Eadj) {
    double E;
    if (r == p.r) E = opz * fma(hs * p.u, u, fma(p.u + u, hr, k0));
    else {
        double d = Lin - p.Lout;
        if (p.r <= B.mlast && r > B.m0) { // may cross a split (rare)
            double dsum = 0.0;
            for (int s = 0; s < B.nsplit; ++s) {
                const uint32_t hb = B.hdr;
                int ms;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(ms) : "r"(hb + (uint32_t)offsetof(BandHeader2, msplit) + 4u * s));
                if (p.r <= ms && r > ms) {
                    const double C = lds_f64(hb + (uint32_t)offsetof(BandHeader2, splitC) + 8u * s);
                    const double D = lds_f64(hb + (uint32_t)offsetof(BandHeader2, splitD) + 8u * s);
                    const double F = lds_f64(hb + (uint32_t)offsetof(BandHeader2, splitF) + 8u * s);
                    d += fma(D, t - F, C);
                    dsum += D;
                }
            }
            E = d * idx;
            Eadj = E - dsum * opz;
            return E;
        }
        E = d * idx;
    }
    Eadj = E;
    return E;
}

template <typename real>
__device__ __forceinline__ void walk_dev(const BandDev &B, const GridNode *__restrict__ grid, double opz, bool bulge, int a, int b,
                                         int js, int je, const real *__restrict__ sd, const real *__restrict__ sb, real &accd_out,
                                         real &accb_out) {
    NodeSt s0{0.0, 0.0, -1}, s1{0.0, 0.0, -1}, so{0.0, 0.0, -1};
    double Eprev_d = 0.0, Eprev_b = 0.0;
    real Sprev_d = real(0), Sprev_b = real(0), accd = real(0), accb = real(0);
    const unsigned span = (unsigned)(b - a);
    // one arrival: P = state of node j - 1, C receives the state of node j
    auto step = [&](const NodeSt &P, NodeSt &C, int j) {
        const double4 ga = *reinterpret_cast<const double4 *>(grid + j); // x, idxp, idxpo, {o, po}
        const int go = __double2loint(ga.w), gpo = __double2hiint(ga.w);
        const double t = opz * ga.x; // the reference's lam = rlam * (1 + z) (src/egg-gencat.cpp:2161), in double
        const bool in_range = (unsigned)(j - a) <= span;
        real Sd = __ldcg(sd + (size_t)j * kLanes);
        Sd = in_range ? Sd : real(0);
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
        const double E_d = interval_dev(B, P, r, u, Lin, q0b.x, q0b.y, hs, t, ga.y, opz, Eadj_d);
        accd = fma(to_real<real>(E_d - Eprev_d), Sprev_d, accd);
        Eprev_d = Eadj_d;
        Sprev_d = Sd;
        // bulge: its SED intervals run between optical nodes
        if (bulge && go >= 0) {
            real Sb = __ldcg(sb + (size_t)(go & (kNextNotOptical - 1)) * kLanes);
            Sb = in_range ? Sb : real(0);
            double E_b, Eadj_b;
            if (gpo == j - 1 && j > js) { E_b = E_d; Eadj_b = Eadj_d; } // the interval is a disk interval too
            else E_b = interval_dev(B, so, r, u, Lin, q0b.x, q0b.y, hs, t, ga.z, opz, Eadj_b);
            accb = fma(to_real<real>(E_b - Eprev_b), Sprev_b, accb);
            Eprev_b = Eadj_b;
            Sprev_b = Sb;
            if (go & kNextNotOptical) so = C; // the next bulge interval does not start as a disk interval
        }
    };
    // two arrivals per trip, the two states swapping roles (no register moves); an odd count runs one node past je,
    // which weighs nothing in range (the grid has two virtual nodes behind the last real one)
    for (int j = js; j <= je; j += 2) {
        step(s0, s1, j);
        step(s1, s0, j + 1);
    }
    accd_out = accd;
    accb_out = accb;
}

// ---- the kernel ---------------------------------------------------------------------------------------
template <typename real, int NT>
__global__ void __launch_bounds__(NT, kLaneCtasPerSm) lane_flux_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ LaneScal sc;
    __shared__ __align__(8) uint64_t bar_tab;
    __shared__ int item_ctr[kMaxGroups], group_items[kMaxGroups];
    __shared__ int a2_ctr;
    __shared__ int jn[4];        // disk nodes [jn0, jn1] and optical nodes [jn2, jn3] that some band of the batch reads
    __shared__ double opzmm[2];  // smallest and largest (1+z) of the batch

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbt = (int)p.nband_total, ntab = (int)q.ntab, nrf = p.do_rf ? (int)p.nrf : 0;
    const int nD = p.nD, nB = p.nB, nline = p.nline;
    const LaneSmem L = lane_smem_layout(p.max_blob, ntab, nbt, p.nrf, nline > 0 ? nline : 0);
    const uint8_t *blob = smem + L.tab;
    double *rfres = reinterpret_cast<double *>(smem + L.rfres);     // [rf][comp][lane]
    uint32_t *lwin = reinterpret_cast<uint32_t *>(smem + L.lwin);   // [comp][line][lane]
    uint32_t *lrange = reinterpret_cast<uint32_t *>(smem + L.lrange); // [comp][line]
    BandWin *bwin = reinterpret_cast<BandWin *>(smem + L.bwin);      // [table in group order]
    double2 *cov = reinterpret_cast<double2 *>(smem + L.cov);        // [col] full span of the filter (coverage test)
    const bool cs17 = p.ir_kind == EGGPHOT_IR_CS17;
    const bool has_bulge = p.has_bulge != 0;
    const int ncomp = has_bulge ? 2 : 1;
    real *SD = reinterpret_cast<real *>(p.scratch) + (size_t)blockIdx.x * q.scratch_stride;
    real *SB = SD + (size_t)(nD + 2) * kLanes;
    real *part = SB + (size_t)(nB + 2) * kLanes; // [slot][comp][lane] block sums
    const real *rO = reinterpret_cast<const real *>(p.rowsDopt), *rI = reinterpret_cast<const real *>(p.rowsDir);
    const real *rP = reinterpret_cast<const real *>(p.rowsDpah), *rB = reinterpret_cast<const real *>(p.rowsB);
    const GridNode *grid = q.grid;
    const uint32_t tab_sa = smem_u32(smem + L.tab);

    if (tid == 0) {
        mbar_init(&bar_tab, 1);
        fence_mbar_init();
    }
    // full spans of the filters by output column: the same for every batch
    for (int fb = tid; fb < ntab; fb += NT) {
        int grp = 0;
        while (grp + 1 < (int)p.ngroups && fb >= (int)p.groups[grp + 1].base) ++grp;
        const BandHeader2 *gh = reinterpret_cast<const BandHeader2 *>(p.groups[grp].blob) + (fb - (int)p.groups[grp].base);
        cov[__ldg(&gh->out_col)] = make_double2(__ldg(&gh->full_first), __ldg(&gh->full_last));
    }
    __syncthreads();
    uint32_t tab_phase = 0;
    int loaded = -1;
    const uint64_t keep = evict_last_policy();

    const uint64_t nbatch = (p.ngal + kLanes - 1) / kLanes;
    for (uint64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const uint64_t g0 = batch * kLanes;
        const int ng = (int)min((uint64_t)kLanes, p.ngal - g0);

        // ---- phase A1: per-lane scalars, four warps (one piece each: the double-precision log / exp10 / divisions of
        //      a batch run side by side); lanes beyond the end of the catalog repeat the batch's first galaxy
        if (warp < 4) {
            const uint64_t g = q.perm[g0 + (lane < ng ? lane : 0)];
            if (warp == 0) {
                const double z = __ldcs(p.z + g), d = __ldcs(p.d + g); // galaxy columns are read once: streaming loads
                sc.opz[lane] = 1.0 + z;
                sc.iopz[lane] = 1.0 / (1.0 + z);
                sc.aflux[lane] = EGGPHOT_LSUN2UJY * (1.0 + z) / (d * d);
                sc.gal[lane] = (uint32_t)g;
                sc.valid[lane] = lane < ng;
            } else if (warp == 1) {
                // the reference passes the FLOAT difference m[i]-out.m2lcor[i] to e10 (src/egg-gencat.cpp:2106, 2114)
                sc.a_bulge[lane] = has_bulge ? exp10((double)(p.m_bulge[g] - __ldcs(p.m2lcor + g))) : 0.0;
            } else if (warp == 2) {
                sc.a_disk[lane] = p.disk_mode != 2 ? exp10((double)(p.m_disk[g] - __ldcs(p.m2lcor + g))) : 0.0;
            } else {
                double c1 = 0.0, c2 = 0.0;
                bool bad = false;
                uint64_t ri = 0, rb = 0, rd = 0;
                if (p.disk_mode != 1) {
                    ri = p.ir_sed[g];
                    if (ri >= p.nir_sed) { bad = true; ri = 0; }
                    if (cs17) {
                        const double fp = p.fpah[g], md = p.mdust[g];
                        c1 = (1.0 - fp) * md;
                        c2 = fp * md;
                    } else {
                        c1 = p.lir[g];
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
                sc.c1[lane] = c1; sc.c2[lane] = c2;
                sc.row_b[lane] = (uint32_t)rb; sc.row_d[lane] = (uint32_t)rd; sc.row_i[lane] = (uint32_t)ri;
                for (int k = 0; k < 3; ++k) sc.igm[k][lane] = p.apply_igm ? __ldcs(p.igm_abs + 3 * g + k) : 1.0f;
                sc.vdisp[lane] = nline ? (double)p.vdisp[g] : 0.0;
                if (bad) atomicExch(p.error_flag, EGGPHOT_ERR_INDEX);
            }
        } else if (warp == 4) {
            if (lane < kMaxGroups) { item_ctr[lane] = 0; group_items[lane] = 0; }
            if (lane == 16) a2_ctr = 0;
            if (lane == 17) { // the rest-frame bands read fixed node ranges
                int lo = nD, hi = -1, blo = nB, bhi = -1;
                for (int r = 0; r < nrf; ++r) {
                    if (p.rfD[r].n > 0) { lo = min(lo, (int)p.rfD[r].j0); hi = max(hi, (int)(p.rfD[r].j0 + p.rfD[r].n - 1)); }
                    if (has_bulge && p.rfB[r].n > 0) { blo = min(blo, (int)p.rfB[r].j0); bhi = max(bhi, (int)(p.rfB[r].j0 + p.rfB[r].n - 1)); }
                }
                jn[0] = lo; jn[1] = hi; jn[2] = blo; jn[3] = bhi;
            }
        }
        __syncthreads();
        if (warp == 0) { // (1+z) range of the batch
            double lo = sc.opz[lane], hi = lo;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = fmin(lo, __shfl_xor_sync(kFull, lo, o));
                hi = fmax(hi, __shfl_xor_sync(kFull, hi, o));
            }
            if (lane == 0) { opzmm[0] = lo; opzmm[1] = hi; }
        }
        // emission-line windows ([vif] bounds, src/egg-gencat.cpp:2081-2084): one thread per (component, line, lane)
        for (int idx = tid; idx < 2 * nline * kLanes; idx += NT) {
            const int ln = idx & 31, cl = idx >> 5, c = cl >= nline, l = cl - c * nline; // c: 0 bulge, 1 disk
            const float *ll = c ? p.line_lum_disk : p.line_lum_bulge;
            int b0 = 1, b1 = 0; // empty
            if ((c || has_bulge) && ll != nullptr && sc.valid[ln] && ll[(uint64_t)sc.gal[ln] * nline + p.line_col[l]] != 0.0f) {
                const double *x = c ? p.xD : p.xB;
                const int n = c ? nD : nB;
                const uint16_t *lut = c ? p.lutD : p.lutB;
                const int nlut = c ? p.lutD_n : p.lutB_n, e0 = c ? p.lutD_e0 : p.lutB_e0;
                const double l0 = (double)p.line_lambda[l], lw = l0 * (sc.vdisp[ln] / 2.998e5);
                const int lo = lut_last_le(x, n, lut, nlut, lut_pos(l0 - 10 * lw, e0), l0 - 10 * lw);
                const int hi = lut_last_le(x, n, lut, nlut, lut_pos(l0 + 10 * lw, e0), l0 + 10 * lw) + 1; // first > b, n if none
                if (!(hi == 0 || lo == n - 1)) {
                    b0 = lo < 0 ? 0 : lo;
                    b1 = hi >= n ? n - 1 : hi;
                }
            }
            lwin[idx] = (uint32_t)b0 | ((uint32_t)b1 << 16);
        }
        __syncthreads();

        // ---- node window of every band table for the whole batch: one thread per table.  The window runs from the
        //      last node at or left of the table's blue end for the largest (1+z) to the first node at or right of its
        //      red end for the smallest, with a relative margin: nodes a galaxy does not weigh get the weight 0 exactly.
        for (int fb = tid; fb < ntab; fb += NT) {
            int grp = 0;
            while (grp + 1 < (int)p.ngroups && fb >= (int)p.groups[grp + 1].base) ++grp;
            const BandHeader2 *gh = reinterpret_cast<const BandHeader2 *>(p.groups[grp].blob) + (fb - (int)p.groups[grp].base);
            const double omin = opzmm[0], omax = opzmm[1];
            const double ff = __ldg(&gh->full_first), fl = __ldg(&gh->full_last);
            uint32_t w0 = 0, w1 = 0, w2 = 0;
            // [vif] sed2flux is NaN (-> flux 0, src/egg-gencat.cpp:2198) unless the SED covers the filter: a band no
            // galaxy of the batch covers is skipped
            const bool any_cov = nD >= 2 && __ldg(p.xD) * omin <= ff && __ldg(p.xD + nD - 1) * omax >= fl;
            if (any_cov && __ldg(&gh->n) > 0) {
                int jlo = max(bis_last_le(p.xD, nD, __ldg(&gh->f_first) / omax * (1.0 - 1e-9)), 0);
                int jhi = min(bis_first_ge(p.xD, nD, __ldg(&gh->f_last) / omin * (1.0 + 1e-9)), nD - 1);
                const bool bulge_any = has_bulge && nB >= 2 && __ldg(p.xB) * omin <= ff && __ldg(p.xB + nB - 1) * omax >= fl;
                if (bulge_any) widen_window_for_bulge(grid, q.next_opt, nD, jlo, jhi);
                const int kfirst = jlo / kLaneBlock, nblk = jhi / kLaneBlock - kfirst + 1;
                w0 = (uint32_t)jlo | ((uint32_t)jhi << 16);
                w1 = (uint32_t)kfirst | ((uint32_t)nblk << 16);
                w2 = bulge_any ? (1u << 30) : 0u;
                atomicMin(&jn[0], jlo);
                atomicMax(&jn[1], jhi);
                if (bulge_any) { // optical nodes inside the window
                    const int jo0 = grid[jlo].o >= 0 ? jlo : q.next_opt[jlo], jo1 = grid[jhi].o >= 0 ? jhi : grid[jhi].po;
                    if (jo0 <= jo1 && jo1 >= 0) { atomicMin(&jn[2], grid_o(grid[jo0])); atomicMax(&jn[3], grid_o(grid[jo1])); }
                }
            }
            bwin[fb].w0 = w0; bwin[fb].w1 = w1; bwin[fb].w2 = w2; bwin[fb].w3 = 0;
        }
        // node range of every line over the batch (union of the lanes' windows): one warp per (component, line)
        for (int cl = warp; cl < 2 * nline; cl += NT / 32) {
            const uint32_t w = lwin[cl * kLanes + lane];
            int b0 = (int)(w & 0xffffu), b1 = (int)(w >> 16);
            if (b0 > b1) { b0 = 0xffff; b1 = 0; }
            b0 = __reduce_min_sync(kFull, b0);
            b1 = __reduce_max_sync(kFull, b1);
            if (lane == 0) lrange[cl] = (uint32_t)b0 | ((uint32_t)b1 << 16);
        }
        __syncthreads();
        // item numbering per group (one thread): item base inside the group, first slot of the block sums
        if (tid == 0) {
            int slot = 0;
            for (int g = 0; g < (int)p.ngroups; ++g) {
                const int base = (int)p.groups[g].base, nb = (int)p.groups[g].nb;
                int run = 0;
                for (int b = 0; b < nb; ++b) {
                    const int nk = (int)(bwin[base + b].w1 >> 16);
                    bwin[base + b].w2 |= (uint32_t)run;
                    bwin[base + b].w3 = (uint32_t)slot;
                    run += nk;
                    slot += nk;
                }
                group_items[g] = run;
            }
            if (slot > (int)q.max_slots) atomicExch(p.error_flag, EGGPHOT_ERR_CUDA);
        }
        __syncthreads();
        const int jn0 = jn[0], jn1 = jn[1], on0 = jn[2], on1 = jn[3];

        // ---- phase A2: rest-frame SEDs with their emission lines, lane = galaxy: 16 bytes of each template row per load,
        //      [node][lane] stores.  Only the nodes some band of the batch reads are built.  The bulge SED carries no
        //      mass scale (applied at the end).
        {
            constexpr int NPL = 16 / (int)sizeof(real), CH = 16; // nodes per load, nodes per task
            const int d0 = jn1 >= jn0 ? (jn0 & ~3) : 0, dn = jn1 >= jn0 ? (jn1 - d0 + CH) / CH : 0;
            const int b0 = (has_bulge && on1 >= on0) ? (on0 & ~3) : 0, bn = (has_bulge && on1 >= on0) ? (on1 - b0 + CH) / CH : 0;
            const real a_d = (real)sc.a_disk[lane], c1 = (real)sc.c1[lane], c2 = (real)sc.c2[lane];
            const real g0f = sc.igm[0][lane], g1f = sc.igm[1][lane], g2f = sc.igm[2][lane];
            const real *rowO = rO + (size_t)sc.row_d[lane] * p.strideD, *rowI = rI + (size_t)sc.row_i[lane] * p.strideD;
            const real *rowP = rP + (size_t)sc.row_i[lane] * p.strideD, *rowB = rB + (size_t)sc.row_b[lane] * p.strideB;
            const double vd = sc.vdisp[lane] / 2.998e5;
            // the bulge SED carries no mass scale, so its line term is divided by it
            const double inv_b = sc.a_bulge[lane] > 0 ? 1.0 / sc.a_bulge[lane] : 0.0;
            // emission lines (add_emission_line, src/egg-gencat.cpp:2078-2095) at nodes j0 .. j0 + NPL - 1 of component c:
            // the bin integrals of all lines whose window holds the node, summed in wavelength order
            auto add_lines = [&](int c, int j0, real *v) {
                const double *x = c ? p.xD : p.xB;
                const int n = c ? nD : nB;
                const uint32_t *wl = lwin + (size_t)c * nline * kLanes + lane;
                const float *ll = (c ? p.line_lum_disk : p.line_lum_bulge);
                for (int l = 0; l < nline; ++l) {
                    const uint32_t lr = lrange[c * nline + l];
                    if ((int)(lr & 0xffffu) > j0 + NPL - 1 || (int)(lr >> 16) < j0) continue; // no galaxy of the batch (uniform)
                    const uint32_t w = wl[l * kLanes];
                    const int w0 = (int)(w & 0xffffu), w1 = (int)(w >> 16);
                    if (w0 > j0 + NPL - 1 || w1 < j0) continue;
                    const double l0 = (double)p.line_lambda[l], lw = l0 * vd;
                    const double amp = (double)ll[(uint64_t)sc.gal[lane] * nline + p.line_col[l]] * l0;
#pragma unroll
                    for (int k = 0; k < NPL; ++k) {
                        const int j = j0 + k;
                        if (j < w0 || j > w1 || j >= n) continue;
                        const double xj = __ldg(x + j);
                        const double xm = j > 0 ? __ldg(x + j - 1) : 0.0, xp = j < n - 1 ? __ldg(x + j + 1) : 0.0;
                        const double llow = j != 0 ? 0.5 * (xm + xj) : xj - 0.5 * (xp - xj);     // :2090
                        const double lup = j != n - 1 ? 0.5 * (xj + xp) : xj + 0.5 * (xj - xm); // :2091
                        v[k] += (real)(xj * line_bin_t<real>(llow, lup, l0, lw, amp, p.line_average) * (c ? 1.0 : inv_b));
                    }
                }
            };
            for (;;) {
                int task = 0;
                if (lane == 0) task = atomicAdd(&a2_ctr, 1);
                task = __shfl_sync(kFull, task, 0);
                if (task >= dn + bn) break;
                if (task < dn) {
                    const int jt = d0 + task * CH;
#pragma unroll
                    for (int qd = 0; qd < CH / NPL; ++qd) {
                        const int j0 = jt + qd * NPL;
                        if (j0 >= p.strideD) break;
                        Vec16<real> o, i, pq;
#pragma unroll
                        for (int k = 0; k < NPL; ++k) o.v[k] = i.v[k] = pq.v[k] = real(0);
                        if (p.disk_mode != 2) o = ldg16(rowO + j0);
                        if (p.disk_mode != 1) i = ldg16(rowI + j0);
                        if (p.disk_mode != 1 && cs17) pq = ldg16(rowP + j0);
                        real v[NPL];
#pragma unroll
                        for (int k = 0; k < NPL; ++k) {
                            v[k] = a_d * o.v[k] + c1 * i.v[k]; // absent terms are zero
                            if (p.disk_mode != 1 && cs17) v[k] += c2 * pq.v[k];
                            if (p.apply_igm && j0 + k < p.igmD[3])
                                v[k] *= igm_factor<real>(j0 + k, p.igmD[0], p.igmD[1], p.igmD[2], p.igmD[3], g0f, g1f, g2f);
                        }
                        if (nline > 0 && p.line_lum_disk != nullptr) add_lines(1, j0, v);
#pragma unroll
                        for (int k = 0; k < NPL; ++k)
                            if (j0 + k <= nD + 1) st_keep(SD + (size_t)(j0 + k) * kLanes + lane, j0 + k < nD ? v[k] : real(0), keep);
                    }
                } else {
                    const int jt = b0 + (task - dn) * CH;
#pragma unroll
                    for (int qd = 0; qd < CH / NPL; ++qd) {
                        const int j0 = jt + qd * NPL;
                        if (j0 >= p.strideB) break;
                        const Vec16<real> o = ldg16(rowB + j0);
                        real v[NPL];
#pragma unroll
                        for (int k = 0; k < NPL; ++k) {
                            v[k] = o.v[k];
                            if (p.apply_igm && j0 + k < p.igmB[3])
                                v[k] *= igm_factor<real>(j0 + k, p.igmB[0], p.igmB[1], p.igmB[2], p.igmB[3], g0f, g1f, g2f);
                        }
                        if (nline > 0 && p.line_lum_bulge != nullptr) add_lines(0, j0, v);
#pragma unroll
                        for (int k = 0; k < NPL; ++k)
                            if (j0 + k <= nB + 1) st_keep(SB + (size_t)(j0 + k) * kLanes + lane, j0 + k < nB ? v[k] : real(0), keep);
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase B: band groups
        // consecutive batches walk the groups in opposite orders: the tables of the group a batch ends with are
        // still in shared memory when the next batch starts with it (one table load less per batch)
        const int ngl = max((int)p.ngroups, nrf > 0 ? 1 : 0);
        const bool rev = ((batch - blockIdx.x) / gridDim.x) & 1;
        const double opz = sc.opz[lane];
        for (int gstep = 0; gstep < ngl; ++gstep) {
            const int grp = rev ? ngl - 1 - gstep : gstep;
            const int nb = grp < (int)p.ngroups ? (int)p.groups[grp].nb : 0;
            const int nband_items = grp < (int)p.ngroups ? group_items[grp] : 0;
            if (nb > 0 && nband_items > 0 && loaded != grp) {
                if (tid == 0) {
                    fence_proxy_async(); // earlier generic-proxy reads of the table region precede the async writes
                    const uint32_t bytes = p.groups[grp].bytes;
                    mbar_expect_tx(&bar_tab, bytes);
                    for (uint32_t o = 0; o < bytes; o += 32768u)
                        bulk_g2s(smem + L.tab + o, p.groups[grp].blob + o, min(32768u, bytes - o), &bar_tab);
                }
                mbar_wait(&bar_tab, tab_phase);
                tab_phase ^= 1;
                loaded = grp;
            }
            const int gbase = grp < (int)p.ngroups ? (int)p.groups[grp].base : 0;
            const int nitems = nband_items + (grp == 0 ? nrf * ncomp : 0);
            for (;;) {
                int idx = 0;
                if (lane == 0) idx = atomicAdd(&item_ctr[grp], 1);
                idx = __shfl_sync(kFull, idx, 0);
                if (idx >= nitems) break;
                if (idx < nband_items) {
                    // (table, block): the table whose item range holds idx (tables are stored longest first)
                    int bi = 0;
                    for (int b = lane; b < nb; b += 32) {
                        const BandWin &bw = bwin[gbase + b];
                        const int ib = (int)(bw.w2 & 0x3fffffffu), nk = (int)(bw.w1 >> 16);
                        if (idx >= ib && idx < ib + nk) bi = b + 1;
                    }
                    bi = __reduce_max_sync(kFull, bi) - 1;
                    const BandWin bw = bwin[gbase + bi];
                    const int kk = idx - (int)(bw.w2 & 0x3fffffffu), k = (int)(bw.w1 & 0xffffu) + kk;
                    const int jlo = (int)(bw.w0 & 0xffffu), jhi = (int)(bw.w0 >> 16);
                    const bool bulge = (bw.w2 >> 30) & 1u;
                    const uint32_t hsa = tab_sa + (uint32_t)bi * (uint32_t)sizeof(BandHeader2);
                    const BandHeader2 &h = reinterpret_cast<const BandHeader2 *>(blob)[bi];
                    BandDev B;
                    B.hdr = hsa;
                    B.rec = tab_sa + h.rec_off; B.hs = tab_sa + h.hs_off; B.bkt = tab_sa + h.bkt_off;
                    B.bk_scale = h.bk_scale; B.bk_c0 = h.bk_c0;
                    B.nbkt_m1 = (int)h.nbkt - 1;
                    B.nsplit = (int)h.nsplit;
                    B.m0 = (int)h.msplit[0];
                    B.mlast = (int)h.msplit[B.nsplit > 0 ? B.nsplit - 1 : 0];
                    B.two_probe = h.two_probe != 0;
                    const int a = max(jlo, k * kLaneBlock), b = min(jhi, k * kLaneBlock + kLaneBlock - 1);
                    int js, je;
                    block_arrivals(grid, q.next_opt, bulge, a, b, js, je);
                    real accd, accb;
                    walk_dev<real>(B, grid, opz, bulge, a, b, js, je, SD + lane, SB + lane, accd, accb);
                    real *ps = part + (size_t)(bw.w3 + (uint32_t)kk) * 2 * kLanes + lane;
                    __stcg(ps + kLanes, accd);
                    __stcg(ps, bulge ? accb : real(0));
                } else {
                    // rest-frame magnitude: dot product with the z = 0 weights (src/egg-gencat.cpp:2150-2158)
                    const int t2 = idx - nband_items;
                    const int comp = has_bulge ? (t2 & 1) : 1, rb = has_bulge ? (t2 >> 1) : t2;
                    const RfDesc dsc = comp ? p.rfD[rb] : p.rfB[rb];
                    const real *S = (comp ? SD : SB) + lane;
                    const real *W = reinterpret_cast<const real *>(p.rfW) + dsc.off;
                    CompSum<real> acc;
                    acc.clear();
                    for (int k = 0; k < dsc.n; ++k) acc.add(__ldcg(S + (size_t)(dsc.j0 + k) * kLanes) * __ldg(W + k));
                    rfres[(rb * 2 + comp) * kLanes + lane] =
                        dsc.covered ? ((double)acc.s + (double)acc.c) * (comp ? 1.0 : sc.a_bulge[lane]) : __longlong_as_double(0x7ff8000000000000LL);
                }
            }
            __syncthreads(); // every item of the group is done; the table region may be overwritten
        }

        // ---- phase C: block sums added per output column in a fixed order (tables in group order, blocks in node
        //      order), data conditions, totals (src/egg-gencat.cpp:2238-2239), catalog columns (rows of one galaxy are
        //      contiguous: consecutive threads write consecutive bands)
        for (int i = tid; i < ng * nbt; i += NT) {
            const int ln = i / nbt, col = i - ln * nbt;
            double sd = 0.0, sb = 0.0;
            for (int fb = 0; fb < ntab; ++fb) {
                if ((int)q.tab_col[fb] != col) continue;
                const BandWin bw = bwin[fb];
                const int nk = (int)(bw.w1 >> 16);
                const real *ps = part + (size_t)bw.w3 * 2 * kLanes + ln;
                for (int kk = 0; kk < nk; ++kk, ps += 2 * kLanes) {
                    sd += (double)__ldcg(ps + kLanes);
                    sb += (double)__ldcg(ps);
                }
            }
            const double o = sc.opz[ln];
            const double2 cv = cov[col];
            // [vif] sed2flux: NaN (-> flux 0, src/egg-gencat.cpp:2198) unless the SED covers the filter
            const bool covd = nD >= 2 && __ldg(p.xD) * o <= cv.x && __ldg(p.xD + nD - 1) * o >= cv.y;
            const bool covb = covd && has_bulge && nB >= 2 && __ldg(p.xB) * o <= cv.x && __ldg(p.xB + nB - 1) * o >= cv.y;
            const double sca = sc.iopz[ln] * sc.aflux[ln];
            double fd = covd ? sd * sca : 0.0;
            double fbv = covb ? sb * sca * sc.a_bulge[ln] : 0.0;
            if (!isfinite(fd)) fd = 0.0;
            if (!isfinite(fbv)) fbv = 0.0;
            const float fdf = (float)fd, fbf = (float)fbv;
            const size_t oo = (size_t)sc.gal[ln] * nbt + col;
            // streaming stores: the catalog columns are written once and must not evict the scratch from L2
            if (p.flux_disk) __stcs(p.flux_disk + oo, fdf);
            if (p.flux_bulge) __stcs(p.flux_bulge + oo, fbf);
            if (p.flux) __stcs(p.flux + oo, fdf + fbf); // vec2f sum
            if (p.flux_disk_f64) __stcs(p.flux_disk_f64 + oo, fd);
            if (p.flux_bulge_f64) __stcs(p.flux_bulge_f64 + oo, fbv);
        }
        for (int i = tid; i < ng * nrf; i += NT) {
            const int ln = i / nrf, r = i - ln * nrf;
            // [vif] lsun2uJy(0, 1e-5, ...) then uJy2mag; non-finite -> +99 (src/egg-gencat.cpp:2152-2156)
            const double arf = EGGPHOT_LSUN2UJY / (1e-5 * 1e-5);
            double mg[2] = {0.0, 0.0}; // the bulge column stays 0 when there is no bulge pass (no_stellar)
            for (int comp = has_bulge ? 0 : 1; comp < 2; ++comp) {
                const double m = -2.5 * log10(rfres[(r * 2 + comp) * kLanes + ln] * arf) + 23.9;
                mg[comp] = isfinite(m) ? m : 99.0;
            }
            const size_t oo = (size_t)sc.gal[ln] * nrf + r;
            const float md = (float)mg[1], mb = (float)mg[0];
            if (p.rfmag_disk) __stcs(p.rfmag_disk + oo, md);
            if (p.rfmag_bulge) __stcs(p.rfmag_bulge + oo, mb);
            if (p.rfmag) // uJy2mag(mag2uJy(disk) + mag2uJy(bulge)) on the f32 columns (src/egg-gencat.cpp:2239)
                __stcs(p.rfmag + oo, (float)(-2.5 * log10(exp10(0.4 * (23.9 - (double)md)) + exp10(0.4 * (23.9 - (double)mb))) + 23.9));
            if (p.rfmag_disk_f64) __stcs(p.rfmag_disk_f64 + oo, mg[1]);
            if (p.rfmag_bulge_f64) __stcs(p.rfmag_bulge_f64 + oo, mg[0]);
        }
        __syncthreads(); // sc / windows / slots are rewritten by the next batch
    }
}

#ifndef EGG_LANE_THREADS
#define EGG_LANE_THREADS 384
#endif
constexpr int kLaneThreads = EGG_LANE_THREADS;

size_t lane_smem_bytes(const LaneParams &q) {
    const FluxParams &p = q.fp;
    return lane_smem_layout(p.max_blob, q.ntab, p.nband_total, p.nrf, p.nline > 0 ? p.nline : 0).total;
}

template <typename real> int launch_lane(const LaneParams &q, int blocks, cudaStream_t st) {
    const size_t smem = lane_smem_bytes(q);
    cudaError_t e = cudaFuncSetAttribute(lane_flux_kernel<real, kLaneThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lane_flux_kernel<real, kLaneThreads><<<blocks, kLaneThreads, smem, st>>>(q);
    return (int)cudaGetLastError();
}

template <typename real> int lane_occupancy(const LaneParams &q) {
    const size_t smem = lane_smem_bytes(q);
    int nb = 0;
    cudaError_t e = cudaFuncSetAttribute(lane_flux_kernel<real, kLaneThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lane_flux_kernel<real, kLaneThreads>, kLaneThreads, smem);
    return e == cudaSuccess ? nb : -(int)e;
}

} // namespace

size_t lane_kernel_smem(const LaneParams &q) { return lane_smem_bytes(q); }

int lane_kernel_max_blocks_per_sm(const LaneParams &q, int precision) {
    return precision == EGGPHOT_FP64 ? lane_occupancy<double>(q) : lane_occupancy<float>(q);
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

int launch_lane_flux_kernel(const LaneParams &q, int precision, int grid_blocks, cudaStream_t stream) {
    return precision == EGGPHOT_FP64 ? launch_lane<double>(q, grid_blocks, stream) : launch_lane<float>(q, grid_blocks, stream);
}

} // namespace eggphot
