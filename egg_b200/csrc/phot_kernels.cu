// phot_kernels.cu — sm_100a kernels of the photometry path.
//
// flux_kernel<real> replaces the two get_flux passes and the totals of src/egg-gencat.cpp:2097-2239
// for one group of bands.  One persistent CTA walks galaxies g = blockIdx.x, +gridDim.x, ...;
// per galaxy it
//   0. derives the per-galaxy scalars ((1+z), K(1+z)/d^2, 10^(M - m2lcor), L_IR ...) and asks the TMA
//      engine (cp.async.bulk, SASS UBLKCP) for the galaxy's template rows;
//   1. while those land: finds, per (band, component), the SED nodes overlapping the filter span and
//      the node windows of the emission lines; then builds the rest-frame SEDs of the disk
//      (optical + IR on the merge_add grid, IGM applied) and of the bulge in shared memory
//      (get_opt_sed / get_ir_sed / merge_add / IGM block, src/egg-gencat.cpp:708-712, 2039-2057, 2118-2140);
//   2. deposits the emission lines (add_emission_line, :2078-2095) with a deterministic owner rule;
//   3. integrates: 32 lanes = 32 consecutive SED nodes, each lane finds its node's filter cell,
//      forms the cumulative quantities K0/K1 (phot_core.cuh), gets its right neighbour's by warp
//      shuffle, and adds the interval's contribution to a compensated per-lane sum; warps split the
//      galaxy's passes evenly and reduce by shuffle;  rest-frame magnitudes are dot products with
//      precomputed z = 0 weights (:2150-2158);
//   4. scales, applies the reference's data conditions (uncovered band -> 0, non-finite mag -> 99),
//      forms the totals (:2238-2239) and writes the f32 catalog columns.
// The band tables of the group (filter nodes with prefix sums, bucket index) are staged once per
// CTA into shared memory by TMA bulk copies.
#include "phot_kernels.cuh"

#include <cooperative_groups.h>
#include <cstdio>

#include "../../include/eggphot.h"

namespace eggphot {

namespace {

constexpr int kMaxItems = 2 * kMaxGroupBands;
constexpr int kLineSlots = 16;
constexpr unsigned kFull = 0xffffffffu;

// ---- PTX wrappers (mbarrier + TMA bulk copy) -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- per-galaxy scalars ----------------------------------------------------------------------------
struct GalScal {
    double opz;      // 1 + z
    double aflux;    // K (1+z) / d^2          [vif] lsun2uJy(z, d, ...)   (src/egg-gencat.cpp:2162)
    double a_bulge;  // 10^(m_bulge - m2lcor)  get_opt_sed mass scale      (:711, :2106)
    double a_disk;   // 10^(m_disk  - m2lcor)
    double c1, c2;   // IR: L_IR (generic, :2055) or Mdust(1-fPAH), Mdust fPAH (cs17, :2047-2051)
    double vdisp;
    float igm[3];
    uint32_t row_b, row_d, row_i;
};

__device__ void gal_scalars(const FluxParams &p, uint64_t g, GalScal &s) {
    const double z = p.z[g], d = p.d[g];
    s.opz = 1.0 + z;
    s.aflux = EGGPHOT_LSUN2UJY * (1.0 + z) / (d * d);
    const float m2l = p.m2lcor[g];
    // the reference passes the FLOAT difference m[i]-out.m2lcor[i] to e10 (src/egg-gencat.cpp:2106, 2114)
    s.a_bulge = p.has_bulge ? exp10((double)(p.m_bulge[g] - m2l)) : 0.0;
    s.a_disk = p.disk_mode != 2 ? exp10((double)(p.m_disk[g] - m2l)) : 0.0;
    s.c1 = s.c2 = 0.0;
    bool bad = false;
    uint64_t ri = 0, rb = 0, rd = 0;
    if (p.disk_mode != 1) {
        ri = p.ir_sed[g];
        if (ri >= p.nir_sed) { bad = true; ri = 0; }
        if (p.ir_kind == EGGPHOT_IR_CS17) {
            const double fp = p.fpah[g], md = p.mdust[g];
            s.c1 = (1.0 - fp) * md;
            s.c2 = fp * md;
        } else {
            s.c1 = p.lir[g];
        }
    }
    if (p.has_bulge) {
        rb = p.opt_sed_bulge[g];
        if (rb >= p.nopt_sed || !p.opt_use[rb]) { bad = true; rb = 0; }
    }
    if (p.disk_mode != 2) {
        rd = p.opt_sed_disk[g];
        if (rd >= p.nopt_sed || !p.opt_use[rd]) { bad = true; rd = 0; }
    }
    s.row_b = (uint32_t)rb; s.row_d = (uint32_t)rd; s.row_i = (uint32_t)ri;
    if (p.apply_igm) { s.igm[0] = p.igm_abs[3 * g]; s.igm[1] = p.igm_abs[3 * g + 1]; s.igm[2] = p.igm_abs[3 * g + 2]; }
    else s.igm[0] = s.igm[1] = s.igm[2] = 1.0f;
    s.vdisp = p.nline ? (double)p.vdisp[g] : 0.0;
    if (bad) atomicExch(p.error_flag, EGGPHOT_ERR_INDEX);
}

// ---- small helpers ---------------------------------------------------------------------------------
// last j in [0, n) with x[j] <= v, -1 if none
__device__ __forceinline__ int last_le(const double *__restrict__ x, int n, double v) {
    int lo = -1, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(x + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}
// first j in [0, n) with x[j] >= v, n if none
__device__ __forceinline__ int first_ge(const double *__restrict__ x, int n, double v) {
    int lo = -1, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(x + mid) >= v) hi = mid; else lo = mid;
    }
    return hi;
}

__device__ __forceinline__ void shfl_node(const NodeState<float> &a, NodeState<float> &r) {
    r.t = __shfl_down_sync(kFull, a.t, 1);
    r.S = __shfl_down_sync(kFull, a.S, 1);
    r.cell = __shfl_down_sync(kFull, a.cell, 1);
    r.q.k0h = __shfl_down_sync(kFull, a.q.k0h, 1);
    r.q.k0l = __shfl_down_sync(kFull, a.q.k0l, 1);
    r.q.k1h = __shfl_down_sync(kFull, a.q.k1h, 1);
    r.q.k1l = __shfl_down_sync(kFull, a.q.k1l, 1);
}
__device__ __forceinline__ void shfl_node(const NodeState<double> &a, NodeState<double> &r) {
    r.t = __shfl_down_sync(kFull, a.t, 1);
    r.S = __shfl_down_sync(kFull, a.S, 1);
    r.cell = __shfl_down_sync(kFull, a.cell, 1);
    r.q.k0 = __shfl_down_sync(kFull, a.q.k0, 1);
    r.q.k1 = __shfl_down_sync(kFull, a.q.k1, 1);
}

template <typename real> __device__ __forceinline__ CompSum<real> warp_reduce(CompSum<real> a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        CompSum<real> b;
        b.s = __shfl_down_sync(kFull, a.s, o);
        b.c = __shfl_down_sync(kFull, a.c, o);
        a.add(b);
    }
    return a;
}

// rest grid access in the kernel's precision
template <typename real> struct GridAcc;
template <> struct GridAcc<float> {
    const float *hi, *lo;
    float opz_hi, opz_lo;
    __device__ __forceinline__ float pos(int j, float fc) const {
        return node_pos(__ldg(hi + j), __ldg(lo + j), opz_hi, opz_lo, fc);
    }
};
template <> struct GridAcc<double> {
    const double *x;
    double opz;
    __device__ __forceinline__ double pos(int j, double fc) const { return node_pos(__ldg(x + j), opz, fc); }
};

struct SmemLayout {
    uint32_t blob, bufA, bufB, bufC, bufS, total;
};
__host__ __device__ inline SmemLayout smem_layout(const FluxParams &p, size_t rs) {
    SmemLayout L;
    auto up = [](size_t v) { return (uint32_t)((v + 127) & ~size_t(127)); };
    uint32_t o = 0;
    L.blob = o; o += up(p.blob_bytes);
    const uint32_t rowD = up((size_t)p.strideD * rs), rowB = up((size_t)p.strideB * rs);
    L.bufA = o; o += rowD;                                   // disk: first row, becomes the disk SED
    L.bufB = o; if (p.disk_mode == 0) o += rowD;             // disk: IR row when both are present
    L.bufC = o; if (p.disk_mode != 1 && p.ir_kind == EGGPHOT_IR_CS17) o += rowD; // cs17 PAH row
    L.bufS = o; if (p.has_bulge) o += rowB;                  // bulge row, becomes the bulge SED
    L.total = o;
    return L;
}

// ---- the kernel ------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(kThreads, 1) flux_kernel(const __grid_constant__ FluxParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ GalScal sc[2];
    __shared__ int it_jlo[kMaxItems], it_jhi[kMaxItems], it_pend[kMaxItems + 1]; // pend: inclusive scan of passes
    __shared__ CompSum<real> part[kMaxItems + kWarps];
    __shared__ int ln_b0[2][kMaxLines], ln_b1[2][kMaxLines];
    __shared__ double rfres[2][16];
    __shared__ __align__(8) uint64_t bar_tab, bar_rows;
    __shared__ int total_passes;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SmemLayout L = smem_layout(p, sizeof(real));
    const uint8_t *blob = smem + L.blob;
    real *SD = reinterpret_cast<real *>(smem + L.bufA);
    real *bufB = reinterpret_cast<real *>(smem + L.bufB);
    real *bufC = reinterpret_cast<real *>(smem + L.bufC);
    real *SB = reinterpret_cast<real *>(smem + L.bufS);
    const BandHeader *hdr = reinterpret_cast<const BandHeader *>(blob);
    const int ncomp = p.has_bulge ? 2 : 1;
    const int nitems = (int)p.nb * ncomp;
    const bool cs17 = p.ir_kind == EGGPHOT_IR_CS17;

    // stage the band tables once per CTA
    if (tid == 0) {
        mbar_init(&bar_tab, 1);
        mbar_init(&bar_rows, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar_tab, p.blob_bytes);
        for (uint32_t o = 0; o < p.blob_bytes; o += 32768u) {
            const uint32_t n = min(32768u, p.blob_bytes - o);
            bulk_g2s(smem + L.blob + o, p.blob + o, n, &bar_tab);
        }
    }
    mbar_wait(&bar_tab, 0);

    uint32_t iter = 0;
    for (uint64_t g = blockIdx.x; g < p.ngal; g += gridDim.x, ++iter) {
        const int par = iter & 1;
        GalScal &s = sc[par];
        // ---- phase 0: scalars + TMA requests for the template rows
        if (tid == 0) {
            gal_scalars(p, g, s);
            fence_proxy_async(); // order the previous galaxy's generic-proxy accesses before the async writes
            const uint32_t rbD = (uint32_t)p.strideD * sizeof(real), rbB = (uint32_t)p.strideB * sizeof(real);
            uint32_t bytes = rbD;
            if (p.disk_mode == 0) bytes += rbD;
            if (p.disk_mode != 1 && cs17) bytes += rbD;
            if (p.has_bulge) bytes += rbB;
            mbar_expect_tx(&bar_rows, bytes);
            const real *rO = reinterpret_cast<const real *>(p.rowsDopt), *rI = reinterpret_cast<const real *>(p.rowsDir);
            const real *rP = reinterpret_cast<const real *>(p.rowsDpah), *rB = reinterpret_cast<const real *>(p.rowsB);
            if (p.disk_mode == 2) bulk_g2s(SD, rI + (size_t)s.row_i * p.strideD, rbD, &bar_rows);
            else bulk_g2s(SD, rO + (size_t)s.row_d * p.strideD, rbD, &bar_rows);
            if (p.disk_mode == 0) bulk_g2s(bufB, rI + (size_t)s.row_i * p.strideD, rbD, &bar_rows);
            if (p.disk_mode != 1 && cs17) bulk_g2s(bufC, rP + (size_t)s.row_i * p.strideD, rbD, &bar_rows);
            if (p.has_bulge) bulk_g2s(SB, rB + (size_t)s.row_b * p.strideB, rbB, &bar_rows);
        }
        __syncthreads(); // #1: scalars visible

        // ---- phase 1a: node windows per (band, component); emission-line windows
        if (tid < nitems) {
            const int bi = tid / ncomp, comp = p.has_bulge ? (tid % ncomp) : 1; // 0 bulge, 1 disk
            const BandHeader &h = hdr[bi];
            const double *x = comp ? p.xD : p.xB;
            const int n = comp ? p.nD : p.nB;
            // [vif] sed2flux: NaN (-> flux 0, src/egg-gencat.cpp:2198) unless the SED covers the filter
            const bool covered = n >= 2 && __ldg(x) * s.opz <= h.full_first && __ldg(x + n - 1) * s.opz >= h.full_last;
            int jlo = 0, jhi = 0;
            if (covered && h.n >= 2) {
                jlo = max(last_le(x, n, h.eff_first / s.opz * (1.0 - 1e-15)), 0);
                jhi = min(first_ge(x, n, h.eff_last / s.opz * (1.0 + 1e-15)), n - 1);
            }
            it_jlo[tid] = jlo;
            it_jhi[tid] = covered ? jhi : -1; // -1 marks "not covered"
        } else if (tid >= 128 && tid < 128 + 2 * kMaxLines) {
            const int c = (tid - 128) / kMaxLines, l = (tid - 128) % kMaxLines; // c: 0 bulge, 1 disk
            const float *ll = c ? p.line_lum_disk : p.line_lum_bulge;
            const bool have = c ? true : (bool)p.has_bulge;
            if (l < p.nline) {
                int b0 = 1, b1 = 0; // empty
                if (have && ll != nullptr && ll[g * p.nline + l] != 0.0f) {
                    const double *x = c ? p.xD : p.xB;
                    const int n = c ? p.nD : p.nB;
                    const double l0 = (double)p.line_lambda[l], lw = l0 * (s.vdisp / 2.998e5);
                    // [vif] bounds(lam, a, b) = {last <= a, first > b}  (src/egg-gencat.cpp:2081-2084)
                    int lo = last_le(x, n, l0 - 10 * lw);
                    int hi = last_le(x, n, l0 + 10 * lw) + 1; // first > b, n if none
                    if (!(hi == 0 || lo == n - 1)) {
                        b0 = lo < 0 ? 0 : lo;
                        b1 = hi >= n ? n - 1 : hi;
                    }
                }
                ln_b0[c][l] = b0;
                ln_b1[c][l] = b1;
            }
        }
        // ---- phase 1b: rest-frame SEDs in place (rows have landed)
        mbar_wait(&bar_rows, iter & 1);
        {
            const real a_d = (real)s.a_disk, c1 = (real)s.c1, c2 = (real)s.c2;
            const real g0 = s.igm[0], g1 = s.igm[1], g2 = s.igm[2];
            for (int j = tid; j < p.nD; j += kThreads) {
                real v;
                if (p.disk_mode == 0) v = a_d * SD[j] + c1 * bufB[j];
                else if (p.disk_mode == 1) v = a_d * SD[j];
                else v = c1 * SD[j];
                if (p.disk_mode != 1 && cs17) v += c2 * bufC[j];
                if (p.apply_igm) v *= igm_factor<real>(j, p.igmD[0], p.igmD[1], p.igmD[2], p.igmD[3], g0, g1, g2);
                SD[j] = v;
            }
            if (p.has_bulge && p.apply_igm)
                for (int j = tid; j < p.igmB[3]; j += kThreads)
                    SB[j] *= igm_factor<real>(j, p.igmB[0], p.igmB[1], p.igmB[2], p.igmB[3], g0, g1, g2);
        }
        __syncthreads(); // #2

        // ---- phase 2: emission lines; pass bookkeeping
        if (p.nline) {
            for (int idx = tid; idx < 2 * p.nline * kLineSlots; idx += kThreads) {
                const int c = idx / (p.nline * kLineSlots), rem = idx % (p.nline * kLineSlots);
                const int l = rem / kLineSlots, w = rem % kLineSlots;
                const float *ll = c ? p.line_lum_disk : p.line_lum_bulge;
                if (ll == nullptr || (c == 0 && !p.has_bulge)) continue;
                const double *x = c ? p.xD : p.xB;
                const int n = c ? p.nD : p.nB;
                real *S = c ? SD : SB;
                // bulge SED carries no mass scale (applied at the end), so its line term is divided by it
                const double inv = c ? 1.0 : (s.a_bulge > 0 ? 1.0 / s.a_bulge : 0.0);
                for (int j = ln_b0[c][l] + w; j <= ln_b1[c][l]; j += kLineSlots) {
                    bool owner = true; // the first line whose window holds node j adds all of them, in order
                    for (int k = 0; k < l; ++k) owner = owner && !(ln_b0[c][k] <= j && j <= ln_b1[c][k]);
                    if (!owner) continue;
                    const double xj = __ldg(x + j);
                    const double xm = j > 0 ? __ldg(x + j - 1) : 0.0, xp = j < n - 1 ? __ldg(x + j + 1) : 0.0;
                    const double llow = j != 0 ? 0.5 * (xm + xj) : xj - 0.5 * (xp - xj);       // :2090
                    const double lup = j != n - 1 ? 0.5 * (xj + xp) : xj + 0.5 * (xj - xm);   // :2091
                    double acc = 0.0;
                    for (int k = l; k < p.nline; ++k) {
                        if (ln_b0[c][k] <= j && j <= ln_b1[c][k]) {
                            const double l0 = (double)p.line_lambda[k], lw = l0 * (s.vdisp / 2.998e5);
                            acc += line_bin(llow, lup, l0, lw, (double)ll[g * p.nline + k] * l0, p.line_average);
                        }
                    }
                    S[j] += (real)(xj * acc * inv);
                }
            }
        }
        if (warp == kWarps - 1) { // inclusive scan of the pass counts (31 intervals per pass)
            int carry = 0;
            for (int base = 0; base < nitems; base += 32) {
                const int it = base + lane;
                int np = 0;
                if (it < nitems && it_jhi[it] > it_jlo[it]) np = (it_jhi[it] - it_jlo[it] + 30) / 31;
                int v = np;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(kFull, v, o);
                    if (lane >= o) v += u;
                }
                if (it < nitems) it_pend[it] = carry + v;
                carry += __shfl_sync(kFull, v, 31);
            }
            if (lane == 0) total_passes = carry;
        }
        __syncthreads(); // #3

        // ---- phase 3: quadrature
        {
            const int P = total_passes;
            const int per = (P + kWarps - 1) / kWarps;
            int pcur = warp * per;
            const int pstop = min(pcur + per, P);
            GridAcc<real> gaB, gaD;
            if constexpr (sizeof(real) == 4) {
                const float oh = (float)s.opz, ol = (float)(s.opz - (double)oh);
                gaB.hi = p.xBhi; gaB.lo = p.xBlo; gaB.opz_hi = oh; gaB.opz_lo = ol;
                gaD.hi = p.xDhi; gaD.lo = p.xDlo; gaD.opz_hi = oh; gaD.opz_lo = ol;
            } else {
                gaB.x = p.xB; gaB.opz = s.opz;
                gaD.x = p.xD; gaD.opz = s.opz;
            }
            int it = 0;
            if (pcur < pstop) { // first item with pend > pcur
                int cnt = 0;
                for (int base = 0; base < nitems; base += 32) {
                    const int k = base + lane;
                    cnt += __popc(__ballot_sync(kFull, k < nitems && it_pend[k] <= pcur));
                }
                it = cnt;
            }
            while (pcur < pstop) {
                while (it_pend[it] <= pcur) ++it;
                const int pbeg = it ? it_pend[it - 1] : 0;
                const int pend = min(it_pend[it], pstop);
                const int bi = it / ncomp, comp = p.has_bulge ? (it % ncomp) : 1;
                BandView<real> bv;
                make_band_view(blob, hdr[bi], bv);
                const real *S = comp ? SD : SB;
                const GridAcc<real> &ga = comp ? gaD : gaB;
                const int jlo = it_jlo[it], jhi = it_jhi[it];
                CompSum<real> acc;
                acc.clear();
                for (int k = pcur - pbeg; k < pend - pbeg; ++k) {
                    int j = jlo + 31 * k + lane;
                    const bool live = lane < 31 && j < jhi;
                    j = min(j, jhi);
                    NodeState<real> nl, nr;
                    eval_node<real>(bv, ga.pos(j, bv.fc), S[j], nl);
                    shfl_node(nl, nr);
                    const real T = eval_interval<real>(bv, nl, nr);
                    if (live) acc.add(T);
                }
                acc = warp_reduce(acc);
                if (lane == 0) part[it + warp] = acc;
                pcur = pend;
            }
            // rest-frame magnitudes: dot products with the z = 0 weights
            if (p.do_rf) {
                for (int r = warp; r < (int)p.nrf * ncomp; r += kWarps) {
                    const int rb = r / ncomp, comp = p.has_bulge ? (r % ncomp) : 1;
                    const RfDesc dsc = comp ? p.rfD[rb] : p.rfB[rb];
                    const real *S = comp ? SD : SB;
                    const real *W = reinterpret_cast<const real *>(p.rfW) + dsc.off;
                    CompSum<real> acc;
                    acc.clear();
                    for (int k = lane; k < dsc.n; k += 32) acc.add(S[dsc.j0 + k] * __ldg(W + k));
                    acc = warp_reduce(acc);
                    if (lane == 0)
                        rfres[comp][rb] = dsc.covered ? (double)acc.s + (double)acc.c : __longlong_as_double(0x7ff8000000000000LL);
                }
            }
        }
        __syncthreads(); // #4

        // ---- phase 4: scale, data conditions, totals, store
        if (tid < (int)p.nb) {
            const int P = total_passes, per = max((P + kWarps - 1) / kWarps, 1);
            double f[2] = {0.0, 0.0};
            for (int comp = p.has_bulge ? 0 : 1; comp < 2; ++comp) {
                const int it = p.has_bulge ? tid * 2 + comp : tid;
                const int pbeg = it ? it_pend[it - 1] : 0, pend = it_pend[it];
                if (pend > pbeg) {
                    CompSum<real> tot;
                    tot.clear();
                    for (int w = pbeg / per; w <= (pend - 1) / per; ++w) tot.add(part[it + w]);
                    f[comp] = ((double)tot.s + (double)tot.c) * s.aflux * (comp ? 1.0 : s.a_bulge);
                }
                if (!isfinite(f[comp])) f[comp] = 0.0; // src/egg-gencat.cpp:2198
            }
            const size_t o = (size_t)g * p.nband_total + hdr[tid].out_col;
            const float fd = (float)f[1], fb = (float)f[0];
            if (p.flux_disk) p.flux_disk[o] = fd;
            if (p.flux_bulge) p.flux_bulge[o] = fb;
            if (p.flux) p.flux[o] = fd + fb; // vec2f sum (src/egg-gencat.cpp:2238)
            if (p.flux_disk_f64) p.flux_disk_f64[o] = f[1];
            if (p.flux_bulge_f64) p.flux_bulge_f64[o] = f[0];
        } else if (p.do_rf && tid >= 64 && tid < 64 + (int)p.nrf) {
            const int rb = tid - 64;
            // [vif] lsun2uJy(0, 1e-5, ...) then uJy2mag; non-finite -> +99 (src/egg-gencat.cpp:2152-2156)
            const double arf = EGGPHOT_LSUN2UJY / (1e-5 * 1e-5);
            double mg[2] = {0.0, 0.0};
            for (int comp = p.has_bulge ? 0 : 1; comp < 2; ++comp) {
                const double fl = rfres[comp][rb] * arf * (comp ? 1.0 : s.a_bulge);
                const double m = -2.5 * log10(fl) + 23.9;
                mg[comp] = isfinite(m) ? m : 99.0;
            }
            const size_t o = (size_t)g * p.nrf + rb;
            const float md = (float)mg[1], mb = (float)mg[0]; // bulge column stays 0 when there is no bulge pass
            if (p.rfmag_disk) p.rfmag_disk[o] = md;
            if (p.rfmag_bulge) p.rfmag_bulge[o] = mb;
            if (p.rfmag) // uJy2mag(mag2uJy(disk) + mag2uJy(bulge)) on the f32 columns (src/egg-gencat.cpp:2239)
                p.rfmag[o] = (float)(-2.5 * log10(exp10(0.4 * (23.9 - (double)md)) + exp10(0.4 * (23.9 - (double)mb))) + 23.9);
            if (p.rfmag_disk_f64) p.rfmag_disk_f64[o] = mg[1];
            if (p.rfmag_bulge_f64) p.rfmag_bulge_f64[o] = mg[0];
        }
        // no barrier needed here: the next iteration writes the other parity of sc[], and everything
        // else it touches before its barrier #1 (TMA into the row buffers) is ordered by barrier #4
    }
}

// ---- save_sed (src/egg-gencat.cpp:2161-2179): one CTA per galaxy, observed-frame lam and sed as f32 ----
template <typename real>
__global__ void __launch_bounds__(256) sed_kernel(const __grid_constant__ SedParams q) {
    const FluxParams &p = q.fp;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ GalScal s;
    __shared__ int b0s[kMaxLines], b1s[kMaxLines];
    const uint64_t g = q.first + blockIdx.x;
    const int comp = q.component;
    const int n = comp ? p.nD : p.nB;
    const double *x = comp ? p.xD : p.xB;
    double *S = reinterpret_cast<double *>(smem);
    if (threadIdx.x == 0) gal_scalars(p, g, s);
    __syncthreads();
    const real *rO = reinterpret_cast<const real *>(p.rowsDopt), *rI = reinterpret_cast<const real *>(p.rowsDir);
    const real *rP = reinterpret_cast<const real *>(p.rowsDpah), *rB = reinterpret_cast<const real *>(p.rowsB);
    const bool cs17 = p.ir_kind == EGGPHOT_IR_CS17;
    const int *ig = comp ? p.igmD : p.igmB;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double v;
        if (!comp) v = s.a_bulge * (double)rB[(size_t)s.row_b * p.strideB + j];
        else {
            v = 0.0;
            if (p.disk_mode != 2) v += s.a_disk * (double)rO[(size_t)s.row_d * p.strideD + j];
            if (p.disk_mode != 1) {
                v += s.c1 * (double)rI[(size_t)s.row_i * p.strideD + j];
                if (cs17) v += s.c2 * (double)rP[(size_t)s.row_i * p.strideD + j];
            }
        }
        if (p.apply_igm) v *= igm_factor<double>(j, ig[0], ig[1], ig[2], ig[3], s.igm[0], s.igm[1], s.igm[2]);
        S[j] = v;
    }
    const float *ll = comp ? p.line_lum_disk : p.line_lum_bulge;
    if ((int)threadIdx.x < p.nline) {
        const int l = threadIdx.x;
        int b0 = 1, b1 = 0;
        if (ll != nullptr && ll[g * p.nline + l] != 0.0f) {
            const double l0 = (double)p.line_lambda[l], lw = l0 * (s.vdisp / 2.998e5);
            int lo = last_le(x, n, l0 - 10 * lw), hi = last_le(x, n, l0 + 10 * lw) + 1;
            if (!(hi == 0 || lo == n - 1)) { b0 = lo < 0 ? 0 : lo; b1 = hi >= n ? n - 1 : hi; }
        }
        b0s[l] = b0; b1s[l] = b1;
    }
    __syncthreads();
    if (p.nline && ll != nullptr) {
        for (int idx = threadIdx.x; idx < p.nline * kLineSlots; idx += blockDim.x) {
            const int l = idx / kLineSlots, w = idx % kLineSlots;
            for (int j = b0s[l] + w; j <= b1s[l]; j += kLineSlots) {
                bool owner = true;
                for (int k = 0; k < l; ++k) owner = owner && !(b0s[k] <= j && j <= b1s[k]);
                if (!owner) continue;
                const double xj = x[j], xm = j > 0 ? x[j - 1] : 0.0, xp = j < n - 1 ? x[j + 1] : 0.0;
                const double llow = j != 0 ? 0.5 * (xm + xj) : xj - 0.5 * (xp - xj);
                const double lup = j != n - 1 ? 0.5 * (xj + xp) : xj + 0.5 * (xj - xm);
                double acc = 0.0;
                for (int k = l; k < p.nline; ++k)
                    if (b0s[k] <= j && j <= b1s[k]) {
                        const double l0 = (double)p.line_lambda[k], lw = l0 * (s.vdisp / 2.998e5);
                        acc += line_bin(llow, lup, l0, lw, (double)ll[g * p.nline + k] * l0, p.line_average);
                    }
                S[j] += xj * acc;
            }
        }
    }
    __syncthreads();
    float *lam = q.lam + (size_t)blockIdx.x * n, *sed = q.sed + (size_t)blockIdx.x * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        lam[j] = (float)(x[j] * s.opz);      // src/egg-gencat.cpp:2161
        sed[j] = (float)(S[j] * s.aflux);    // :2162  (S = lambda_rest * L)
    }
}

} // namespace

size_t flux_kernel_smem(const FluxParams &p, int precision) {
    return smem_layout(p, precision == EGGPHOT_FP64 ? 8 : 4).total;
}

template <typename real> static int launch_t(const FluxParams &p, size_t smem, int blocks, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(flux_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    flux_kernel<real><<<blocks, kThreads, smem, st>>>(p);
    return (int)cudaGetLastError();
}

int launch_flux_kernel(const FluxParams &p, int precision, int grid_blocks, cudaStream_t stream) {
    const size_t smem = flux_kernel_smem(p, precision);
    return precision == EGGPHOT_FP64 ? launch_t<double>(p, smem, grid_blocks, stream)
                                     : launch_t<float>(p, smem, grid_blocks, stream);
}

int flux_kernel_max_blocks_per_sm(const FluxParams &p, int precision) {
    const size_t smem = flux_kernel_smem(p, precision);
    int nb = 0;
    cudaError_t e;
    if (precision == EGGPHOT_FP64) {
        e = cudaFuncSetAttribute(flux_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, flux_kernel<double>, kThreads, smem);
    } else {
        e = cudaFuncSetAttribute(flux_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, flux_kernel<float>, kThreads, smem);
    }
    return e == cudaSuccess ? nb : -(int)e;
}

int launch_sed_kernel(const SedParams &q, int precision, cudaStream_t stream) {
    const int n = q.component ? q.fp.nD : q.fp.nB;
    const size_t smem = (size_t)n * sizeof(double);
    cudaError_t e;
    if (precision == EGGPHOT_FP64) {
        e = cudaFuncSetAttribute(sed_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sed_kernel<double><<<(unsigned)q.count, 256, smem, stream>>>(q);
    } else {
        e = cudaFuncSetAttribute(sed_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sed_kernel<float><<<(unsigned)q.count, 256, smem, stream>>>(q);
    }
    return (int)cudaGetLastError();
}

} // namespace eggphot
