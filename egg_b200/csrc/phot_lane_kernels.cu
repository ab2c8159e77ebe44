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
//   * One persistent CTA per SM takes batches of 32 galaxies.  Phase A: per-lane scalars, one node window per band for
//     the batch, rest-frame SEDs of the 32 disks (get_opt_sed / get_ir_sed / merge_add / IGM, :708-712, :2039-2057,
//     :2118-2140) and bulges into a per-CTA scratch laid out [node][lane] (coalesced 128-byte rows; stays in L2), then
//     the emission lines (:2078-2095), each lane depositing its own galaxy's lines.  Phase B: per band group, the
//     node records of the group's filters are staged in shared memory by TMA bulk copies (SASS UBLKCP + mbarrier);
//     warps pull (band, block of 64 nodes) items; every lane walks the block for its galaxy (phot_lane.cuh: two
//     expansions per node, weights from their differences, no shuffles, no per-lane windows); block sums are added
//     into the batch's result array in block order (ticket per band: bit-reproducible).  Phase C: data conditions
//     (uncovered band -> 0, :2198; non-finite magnitude -> 99, :2156), totals (:2238-2239), catalog columns.
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
    uint32_t loc[kZBuckets / 1024], s = 0;
#pragma unroll
    for (int k = 0; k < kZBuckets / 1024; ++k) { loc[k] = hist[t * (kZBuckets / 1024) + k]; s += loc[k]; }
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const uint32_t v = t >= o ? part[t - o] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - s;
#pragma unroll
    for (int k = 0; k < kZBuckets / 1024; ++k) { hist[t * (kZBuckets / 1024) + k] = run; run += loc[k]; }
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

// a lane's column of the [node][lane] scratch arrays (read through L2: the lines were written by this CTA)
template <typename real> struct LaneSed {
    const real *sd, *sb;
    __device__ __forceinline__ real disk(int j) const { return __ldcg(sd + (size_t)j * kLanes); }
    __device__ __forceinline__ real bulge(int o) const { return __ldcg(sb + (size_t)o * kLanes); }
};

// spin limit of the in-kernel waits (tickets, table barrier): a protocol error must end the kernel with an error code,
// not hang the device
constexpr uint32_t kSpinLimit = 1u << 26;

struct BandWin { uint32_t w0, w1, w2; int turn; }; // jlo | jhi << 16; first block | blocks << 16; item base | bulge << 30

// ---- the kernel ---------------------------------------------------------------------------------------
template <typename real, int NT>
__global__ void __launch_bounds__(NT, 1) lane_flux_kernel(const __grid_constant__ LaneParams q) {
    const FluxParams &p = q.fp;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ LaneScal sc;
    __shared__ __align__(8) uint64_t bar_tab;
    __shared__ int item_ctr[kMaxGroups], group_items[kMaxGroups];
    __shared__ int a2_ctr, a3_ctr, any_bulge_line;
    __shared__ int jn[4];        // disk nodes [jn0, jn1] and optical nodes [jn2, jn3] that some band of the batch reads
    __shared__ double opzmm[2];  // smallest and largest (1+z) of the batch

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const int nbt = (int)p.nband_total, nrf = p.do_rf ? (int)p.nrf : 0;
    const int nD = p.nD, nB = p.nB, nline = p.nline;
    const LaneSmem L = lane_smem_layout(p.max_blob, nbt, p.nrf, nline > 0 ? nline : 0);
    const uint8_t *blob = smem + L.tab;
    double *res = reinterpret_cast<double *>(smem + L.res);     // [col][comp][lane]
    double *rfres = reinterpret_cast<double *>(smem + L.rfres); // [rf][comp][lane]
    uint32_t *lwin = reinterpret_cast<uint32_t *>(smem + L.lwin); // [comp][line][lane]
    BandWin *bwin = reinterpret_cast<BandWin *>(smem + L.bwin);  // [band in group order]
    double2 *cov = reinterpret_cast<double2 *>(smem + L.cov);    // [col] full span of the filter (coverage test)
    const BandHeader2 *hdr = reinterpret_cast<const BandHeader2 *>(blob);
    const bool cs17 = p.ir_kind == EGGPHOT_IR_CS17;
    const bool has_bulge = p.has_bulge != 0;
    const int ncomp = has_bulge ? 2 : 1;
    real *SD = reinterpret_cast<real *>(p.scratch) + (size_t)blockIdx.x * q.scratch_stride;
    real *SB = SD + (size_t)(nD + 1) * kLanes;
    const real *rO = reinterpret_cast<const real *>(p.rowsDopt), *rI = reinterpret_cast<const real *>(p.rowsDir);
    const real *rP = reinterpret_cast<const real *>(p.rowsDpah), *rB = reinterpret_cast<const real *>(p.rowsB);
    const GridNode *grid = q.grid;

    if (tid == 0) {
        mbar_init(&bar_tab, 1);
        fence_mbar_init();
    }
    // full spans of the filters by output column: the same for every batch
    for (int fb = tid; fb < nbt; fb += NT) {
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
            if (lane == 16) { any_bulge_line = 0; a2_ctr = 0; a3_ctr = 0; }
            if (lane == 17) { // the rest-frame bands read fixed node ranges
                int lo = nD, hi = -1, blo = nB, bhi = -1;
                for (int r = 0; r < nrf; ++r) {
                    if (p.rfD[r].n > 0) { lo = min(lo, (int)p.rfD[r].j0); hi = max(hi, (int)(p.rfD[r].j0 + p.rfD[r].n - 1)); }
                    if (has_bulge && p.rfB[r].n > 0) { blo = min(blo, (int)p.rfB[r].j0); bhi = max(bhi, (int)(p.rfB[r].j0 + p.rfB[r].n - 1)); }
                }
                jn[0] = lo; jn[1] = hi; jn[2] = blo; jn[3] = bhi;
            }
        }
        for (int i = tid; i < nbt * 2 * kLanes; i += NT) res[i] = 0.0;
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
        __syncthreads();

        // ---- node window of every band for the whole batch: one thread per band.  The window runs from the last node
        //      at or left of the filter's blue end for the largest (1+z) to the first node at or right of its red end for
        //      the smallest, with a relative margin: nodes a galaxy does not weigh get the weight 0 exactly.
        for (int fb = tid; fb < nbt; fb += NT) {
            int grp = 0;
            while (grp + 1 < (int)p.ngroups && fb >= (int)p.groups[grp + 1].base) ++grp;
            const BandHeader2 *gh = reinterpret_cast<const BandHeader2 *>(p.groups[grp].blob) + (fb - (int)p.groups[grp].base);
            const double omin = opzmm[0], omax = opzmm[1];
            const double ff = __ldg(&gh->full_first), fl = __ldg(&gh->full_last);
            uint32_t w0 = 0, w1 = 0, w2 = 0;
            // [vif] sed2flux is NaN (-> flux 0, src/egg-gencat.cpp:2198) unless the SED covers the filter: a band no
            // galaxy of the batch covers is skipped
            const bool any_cov = nD >= 2 && __ldg(p.xD) * omin <= ff && __ldg(p.xD + nD - 1) * omax >= fl;
            if (any_cov && __ldg(&gh->n) >= 2) {
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
                    if (jo0 <= jo1 && jo1 >= 0) { atomicMin(&jn[2], grid[jo0].o); atomicMax(&jn[3], grid[jo1].o); }
                }
            }
            bwin[fb].w0 = w0; bwin[fb].w1 = w1; bwin[fb].w2 = w2; bwin[fb].turn = 0;
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
            if (!c && b0 <= b1) any_bulge_line = 1;
        }
        __syncthreads();
        // item numbering per group (one thread per group); the nodes the SED build has to provide
        if (tid < (int)p.ngroups) {
            const int base = (int)p.groups[tid].base, nb = (int)p.groups[tid].nb;
            int run = 0;
            for (int b = 0; b < nb; ++b) {
                bwin[base + b].w2 |= (uint32_t)run;
                run += (int)(bwin[base + b].w1 >> 16);
            }
            group_items[tid] = run;
        }
        __syncthreads();
        const int jn0 = jn[0], jn1 = jn[1], on0 = jn[2], on1 = jn[3];
        const bool bulge_lines = any_bulge_line != 0;

        // ---- phase A2: rest-frame SEDs, lane = galaxy: 16 bytes of each template row per load, [node][lane] stores.
        //      Only the nodes some band of the batch reads are built.  The bulge SED carries no mass scale (applied at
        //      the end).
        {
            constexpr int NPL = 16 / (int)sizeof(real), CH = 16; // nodes per load, nodes per task
            const int d0 = jn1 >= jn0 ? (jn0 & ~3) : 0, dn = jn1 >= jn0 ? (jn1 - d0 + CH) / CH : 0;
            const int b0 = (has_bulge && on1 >= on0) ? (on0 & ~3) : 0, bn = (has_bulge && on1 >= on0) ? (on1 - b0 + CH) / CH : 0;
            const real a_d = (real)sc.a_disk[lane], c1 = (real)sc.c1[lane], c2 = (real)sc.c2[lane];
            const real g0f = sc.igm[0][lane], g1f = sc.igm[1][lane], g2f = sc.igm[2][lane];
            const real *rowO = rO + (size_t)sc.row_d[lane] * p.strideD, *rowI = rI + (size_t)sc.row_i[lane] * p.strideD;
            const real *rowP = rP + (size_t)sc.row_i[lane] * p.strideD, *rowB = rB + (size_t)sc.row_b[lane] * p.strideB;
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
#pragma unroll
                        for (int k = 0; k < NPL; ++k) {
                            real v = a_d * o.v[k] + c1 * i.v[k]; // absent terms are zero
                            if (p.disk_mode != 1 && cs17) v += c2 * pq.v[k];
                            if (p.apply_igm && j0 + k < p.igmD[3])
                                v *= igm_factor<real>(j0 + k, p.igmD[0], p.igmD[1], p.igmD[2], p.igmD[3], g0f, g1f, g2f);
                            if (j0 + k <= nD) st_keep(SD + (size_t)(j0 + k) * kLanes + lane, j0 + k < nD ? v : real(0), keep);
                        }
                    }
                } else {
                    const int jt = b0 + (task - dn) * CH;
#pragma unroll
                    for (int qd = 0; qd < CH / NPL; ++qd) {
                        const int j0 = jt + qd * NPL;
                        if (j0 >= p.strideB) break;
                        const Vec16<real> o = ldg16(rowB + j0);
#pragma unroll
                        for (int k = 0; k < NPL; ++k) {
                            real v = o.v[k];
                            if (p.apply_igm && j0 + k < p.igmB[3])
                                v *= igm_factor<real>(j0 + k, p.igmB[0], p.igmB[1], p.igmB[2], p.igmB[3], g0f, g1f, g2f);
                            if (j0 + k <= nB) st_keep(SB + (size_t)(j0 + k) * kLanes + lane, j0 + k < nB ? v : real(0), keep);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase A3: emission lines (add_emission_line, src/egg-gencat.cpp:2078-2095).  An item = one line of one
        //      component for the 32 galaxies; lines are sorted by wavelength, so their node windows are ordered too:
        //      the first line whose window holds node j owns it and adds the contributions of all lines at j, in that
        //      order (no two items touch the same node of a galaxy; fixed sum order).
        if (nline > 0) {
            const int nitem = (bulge_lines ? 2 : 1) * nline; // disk lines first
            for (;;) {
                int it = 0;
                if (lane == 0) it = atomicAdd(&a3_ctr, 1);
                it = __shfl_sync(kFull, it, 0);
                if (it >= nitem) break;
                const int c = it < nline, l = c ? it : it - nline;
                const uint32_t *wl = lwin + (size_t)c * nline * kLanes + lane;
                const uint32_t w = wl[l * kLanes];
                const int b0 = (int)(w & 0xffffu), b1 = (int)(w >> 16);
                int prevmax = -1;
                for (int k = 0; k < l; ++k) {
                    const uint32_t wk = wl[k * kLanes];
                    if ((wk & 0xffffu) <= (wk >> 16)) prevmax = max(prevmax, (int)(wk >> 16));
                }
                // nodes no band reads need no deposit
                const int lo = c ? jn0 : on0, hi = c ? jn1 : on1;
                const int start = max(max(b0, prevmax + 1), lo), stop = min(b1, hi);
                const double *x = c ? p.xD : p.xB;
                const int n = c ? nD : nB;
                real *S = (c ? SD : SB) + lane;
                const float *ll = (c ? p.line_lum_disk : p.line_lum_bulge) + (uint64_t)sc.gal[lane] * nline;
                const double vd = sc.vdisp[lane] / 2.998e5;
                // the bulge SED carries no mass scale, so its line term is divided by it
                const double inv = c ? 1.0 : (sc.a_bulge[lane] > 0 ? 1.0 / sc.a_bulge[lane] : 0.0);
                for (int j = start; j <= stop; ++j) {
                    const double xj = __ldg(x + j);
                    const double xm = j > 0 ? __ldg(x + j - 1) : 0.0, xp = j < n - 1 ? __ldg(x + j + 1) : 0.0;
                    const double llow = j != 0 ? 0.5 * (xm + xj) : xj - 0.5 * (xp - xj);     // :2090
                    const double lup = j != n - 1 ? 0.5 * (xj + xp) : xj + 0.5 * (xj - xm); // :2091
                    double acc = 0.0;
                    for (int k = l; k < nline; ++k) {
                        const uint32_t wk = wl[k * kLanes];
                        const int k0 = (int)(wk & 0xffffu), k1 = (int)(wk >> 16);
                        if (k0 > k1) continue; // empty window
                        if (k0 > j) break;     // windows are ordered: no later line reaches back to j
                        if (j <= k1) {
                            const double l0 = (double)p.line_lambda[k], lw = l0 * vd;
                            acc += line_bin_t<real>(llow, lup, l0, lw, (double)ll[p.line_col[k]] * l0, p.line_average);
                        }
                    }
                    real *sp = S + (size_t)j * kLanes;
                    st_keep(sp, __ldcg(sp) + (real)(xj * acc * inv), keep);
                }
            }
            __syncthreads();
        }

        // ---- phase B: band groups
        // consecutive batches walk the groups in opposite orders: the tables of the group a batch ends with are
        // still in shared memory when the next batch starts with it (one table load less per batch)
        const int ngl = max((int)p.ngroups, nrf > 0 ? 1 : 0);
        const bool rev = ((batch - blockIdx.x) / gridDim.x) & 1;
        const double opz = sc.opz[lane];
        const LaneSed<real> sed{SD + lane, SB + lane};
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
                    // (band, block): the band whose item range holds idx (bands are stored longest first)
                    int bi = 0;
                    for (int b = lane; b < nb; b += 32) {
                        const BandWin &bw = bwin[gbase + b];
                        const int ib = (int)(bw.w2 & 0x3fffffffu), nk = (int)(bw.w1 >> 16);
                        if (idx >= ib && idx < ib + nk) bi = b + 1;
                    }
                    bi = __reduce_max_sync(kFull, bi) - 1;
                    BandWin &bw = bwin[gbase + bi];
                    const int kk = idx - (int)(bw.w2 & 0x3fffffffu), k = (int)(bw.w1 & 0xffffu) + kk, nk = (int)(bw.w1 >> 16);
                    const int jlo = (int)(bw.w0 & 0xffffu), jhi = (int)(bw.w0 >> 16);
                    const bool bulge = (bw.w2 >> 30) & 1u;
                    const BandHeader2 &h = hdr[bi];
                    Band2 bv;
                    make_band2(blob, h, bv);
                    const int a = max(jlo, k * kLaneBlock), b = min(jhi, k * kLaneBlock + kLaneBlock - 1);
                    int js, je;
                    block_arrivals(grid, q.next_opt, bulge, a, b, js, je);
                    real accd = real(0), accb = real(0);
                    walk_block<real>(bv, grid, opz, bulge, a, b, js, je, sed, accd, accb);
                    // block sums are added in block order (ticket per band): the result does not depend on which warp
                    // ran which block
                    if (nk > 1) {
                        volatile int *turn = &bw.turn;
                        uint32_t spins = 0;
                        while (*turn != kk) {
                            __nanosleep(20);
                            if (++spins > kSpinLimit) { atomicExch(p.error_flag, EGGPHOT_ERR_CUDA); break; }
                        }
                        __threadfence_block();
                    }
                    double *r = res + (size_t)h.out_col * 2 * kLanes + lane;
                    r[kLanes] += (double)accd;
                    if (bulge) r[0] += (double)accb;
                    if (nk > 1) {
                        __threadfence_block();
                        __syncwarp();
                        if (lane == 0) *(volatile int *)&bw.turn = kk + 1;
                    }
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

        // ---- phase C: data conditions, totals (src/egg-gencat.cpp:2238-2239), catalog columns (rows of one galaxy
        //      are contiguous: consecutive threads write consecutive bands)
        for (int i = tid; i < ng * nbt; i += NT) {
            const int ln = i / nbt, col = i - ln * nbt;
            const double o = sc.opz[ln];
            const double2 cv = cov[col];
            // [vif] sed2flux: NaN (-> flux 0, src/egg-gencat.cpp:2198) unless the SED covers the filter
            const bool covd = nD >= 2 && __ldg(p.xD) * o <= cv.x && __ldg(p.xD + nD - 1) * o >= cv.y;
            const bool covb = covd && has_bulge && nB >= 2 && __ldg(p.xB) * o <= cv.x && __ldg(p.xB + nB - 1) * o >= cv.y;
            const double sca = sc.iopz[ln] * sc.aflux[ln];
            double fd = covd ? res[(size_t)(col * 2 + 1) * kLanes + ln] * sca : 0.0;
            double fbv = covb ? res[(size_t)(col * 2) * kLanes + ln] * sca * sc.a_bulge[ln] : 0.0;
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
        __syncthreads(); // res / sc / windows are rewritten by the next batch
    }
}

template <typename real> struct LaneThreads;
template <> struct LaneThreads<float> { static constexpr int n = 512; };
template <> struct LaneThreads<double> { static constexpr int n = 512; };

template <typename real> size_t lane_smem_bytes(const LaneParams &q) {
    const FluxParams &p = q.fp;
    return lane_smem_layout(p.max_blob, p.nband_total, p.nrf, p.nline > 0 ? p.nline : 0).total;
}

template <typename real> int launch_lane(const LaneParams &q, int blocks, cudaStream_t st) {
    constexpr int NT = LaneThreads<real>::n;
    const size_t smem = lane_smem_bytes<real>(q);
    cudaError_t e = cudaFuncSetAttribute(lane_flux_kernel<real, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lane_flux_kernel<real, NT><<<blocks, NT, smem, st>>>(q);
    return (int)cudaGetLastError();
}

} // namespace

size_t lane_kernel_smem(const LaneParams &q) { return lane_smem_bytes<float>(q); }

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
