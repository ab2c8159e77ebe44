// phot_kernels.cu — sm_100a kernels of the photometry path.
//
// flux_kernel<real> replaces the two get_flux passes and the totals of src/egg-gencat.cpp:2097-2239.
// One persistent CTA per SM walks batches of galaxies (32 in the fp32 mode, 16 in fp64).  Per batch it
//   A. derives the per-galaxy scalars ((1+z), K(1+z)/d^2, 10^(M - m2lcor), L_IR ...), lays out the node window
//      [jlo, jhi] and split node of every (band, galaxy) and the windows of the emission lines, and builds the
//      rest-frame SED of every disk (optical + IR on the merge_add grid, IGM applied) together with the bulge
//      SED resampled on the same grid (get_opt_sed / get_ir_sed / merge_add / IGM block,
//      src/egg-gencat.cpp:708-712, 2039-2057, 2118-2140) into a per-CTA scratch that stays in L2 - only at the
//      nodes some band of the galaxy reads - then deposits the emission lines (add_emission_line, :2078-2095)
//      with a deterministic owner rule;
//   B. for each band group: stages the group's filter tables (per filter cell: node position, {Lambda, K0},
//      R/2, {slope, width}; bucket index) into shared memory with TMA bulk copies (cp.async.bulk, SASS UBLKCP);
//      warps pull (band, galaxy) items off a shared counter; a pass of an item = 32 lanes on 32 consecutive SED
//      nodes: each lane finds its node's filter cell (exact, in double), evaluates the band's piece-wise
//      quadratic Lambda there, the lanes exchange Lambda and its first differences by shuffle and each owned
//      node's quadrature weight W_j multiplies the disk and bulge SED values (phot_core.cuh, "weights form");
//      per-lane sums, one shuffle reduction per item for both components; rest-frame magnitudes are dot
//      products with precomputed z = 0 weights (:2150-2158);
//   C. applies the reference's data conditions (uncovered band -> 0, non-finite mag -> 99), forms
//      the totals (:2238-2239) and writes the f32 catalog columns with coalesced streaming stores.
// What limits it is the SM's load/store data pipe.
#include "phot_kernels.cuh"

#include <cstdio>
#include <cstdlib>

#include "../../include/eggphot.h"
#include "phot_tables.hpp"
#include "phot_device.cuh"

namespace eggphot {

namespace {

constexpr int kLineSlots = 16;

// ---- per-galaxy scalars ----------------------------------------------------------------------------
struct alignas(16) GalScal {
    // the four scalars every (band, galaxy) item reads, adjacent and 16-byte aligned: two 16-byte loads
    double opz;      // 1 + z
    double iopz;     // 1/(1+z)
    double aflux;    // K (1+z) / d^2          [vif] lsun2uJy(z, d, ...)   (src/egg-gencat.cpp:2162)
    double a_bulge;  // 10^(m_bulge - m2lcor)  get_opt_sed mass scale      (:711, :2106)
    double a_disk;   // 10^(m_disk  - m2lcor)
    double c1, c2;   // IR: L_IR (generic, :2055) or Mdust(1-fPAH), Mdust fPAH (cs17, :2047-2051)
    double vdisp;
    float l2opz;     // log2(1+z) * kLutPerOctave
    float igm[3];
    uint32_t row_b, row_d, row_i;
};

// The scalars of one galaxy in four independent pieces (one thread each in the flux kernel, so that the
// double-precision log2 / exp10 / divisions of a batch run side by side): piece < 0 computes all of them.
__device__ void gal_scalars(const FluxParams &p, uint64_t g, GalScal &s, int piece = -1) {
    if (piece < 0 || piece == 0) {
        const double z = __ldcs(p.z + g), d = __ldcs(p.d + g); // galaxy columns are read once: streaming loads
        s.opz = 1.0 + z;
        s.iopz = 1.0 / s.opz;
        s.l2opz = (float)(log2(s.opz) * kLutPerOctave);
        s.aflux = EGGPHOT_LSUN2UJY * (1.0 + z) / (d * d);
    }
    // the reference passes the FLOAT difference m[i]-out.m2lcor[i] to e10 (src/egg-gencat.cpp:2106, 2114)
    if (piece < 0 || piece == 1) s.a_bulge = p.has_bulge ? exp10((double)(p.m_bulge[g] - __ldcs(p.m2lcor + g))) : 0.0;
    if (piece < 0 || piece == 2) s.a_disk = p.disk_mode != 2 ? exp10((double)(p.m_disk[g] - __ldcs(p.m2lcor + g))) : 0.0;
    if (piece < 0 || piece == 3) {
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
        if (p.apply_igm) { s.igm[0] = __ldcs(p.igm_abs + 3 * g); s.igm[1] = __ldcs(p.igm_abs + 3 * g + 1); s.igm[2] = __ldcs(p.igm_abs + 3 * g + 2); }
        else s.igm[0] = s.igm[1] = s.igm[2] = 1.0f;
        s.vdisp = p.nline ? (double)p.vdisp[g] : 0.0;
        if (bad) atomicExch(p.error_flag, EGGPHOT_ERR_INDEX);
    }
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
// lane sums -> one double: five shuffle-add steps in a fixed order (bit-reproducible)
__device__ __forceinline__ double warp_total(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
    return v;
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

// find_cell (phot_core.cuh) for a full warp.  When the band's buckets are narrower than its cells one comparison
// settles it; otherwise the forward scan runs as a warp-uniform loop on a vote instead of a per-lane
// data-dependent loop: the warp executes the longest lane's scan either way, and the lanes provably stay
// converged for the shuffles that follow (with a per-lane loop nvcc 12.9 emitted neither a reconvergence barrier
// after it nor a WARPSYNC before the shuffles, and lanes read stale registers).
template <typename real>
__device__ __forceinline__ int find_cell_warp(const BandView<real> &b, double t) {
    int k = (int)((float)(t - b.f_first) * b.bkt_inv_w);
    k = max(0, min(k, b.nbkt - 1));
    int c = b.bkt[k];
    bool up = t >= b.f[cell_slot(c + 1)];
    if (b.one_probe) return c + (up ? 1 : 0);
    while (__any_sync(kFull, up)) {
        c += up ? 1 : 0;
        up = t >= b.f[cell_slot(c + 1)];
    }
    return c;
}

__device__ __forceinline__ float rcp_approx(float x) { // one MUFU.RCP, no range fix-ups
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// 16 bytes of a 16-byte aligned row: four floats or two doubles
template <typename real> struct Vec16 { real v[16 / sizeof(real)]; };
__device__ __forceinline__ Vec16<float> ldg16(const float *p) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
    return Vec16<float>{{a.x, a.y, a.z, a.w}};
}
__device__ __forceinline__ Vec16<double> ldg16(const double *p) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    return Vec16<double>{{a.x, a.y}};
}
// the same with an L2 evict-last hint: the scratch is re-read and rewritten every batch and should outlive the
// streamed inputs and outputs in L2
__device__ __forceinline__ uint64_t evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st16_keep(float *p, const Vec16<float> &q, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(q.v[0]), "f"(q.v[1]), "f"(q.v[2]), "f"(q.v[3]), "l"(pol) : "memory");
}
__device__ __forceinline__ void st16_keep(double *p, const Vec16<double> &q, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(q.v[0]), "d"(q.v[1]), "l"(pol) : "memory");
}
// four consecutive elements of a 16-byte aligned row
template <typename real> struct Quad { real v[4]; };
__device__ __forceinline__ Quad<float> ldg4(const float *p) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
    return Quad<float>{{a.x, a.y, a.z, a.w}};
}
__device__ __forceinline__ Quad<double> ldg4(const double *p) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return Quad<double>{{a.x, a.y, b.x, b.y}};
}
__device__ __forceinline__ void st4(float *p, const Quad<float> &q) { *reinterpret_cast<float4 *>(p) = make_float4(q.v[0], q.v[1], q.v[2], q.v[3]); }
__device__ __forceinline__ void st4(double *p, const Quad<double> &q) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(q.v[0], q.v[1]);
    reinterpret_cast<double2 *>(p)[1] = make_double2(q.v[2], q.v[3]);
}
template <typename real> struct Pair;
template <> struct Pair<float> { typedef float2 type; };
template <> struct Pair<double> { typedef double2 type; };
__device__ __forceinline__ GridA ldg_grid(const GridA *g) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(g));
    return GridA{a.x, a.y};
}
__device__ __forceinline__ GridBD ldg_grid(const GridBD *g) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(g));
    return GridBD{a.x, a.y};
}

struct SmemLayout {
    uint32_t tab, res, rfres, lb, lt, win, total;
};
__host__ __device__ inline SmemLayout smem_layout(const FluxParams &p) {
    SmemLayout L;
    auto up = [](size_t v) { return (uint32_t)((v + 127) & ~size_t(127)); };
    const size_t nl = p.nline > 0 ? p.nline : 0;
    uint32_t o = 0;
    L.tab = o; o += up(p.max_blob);                                              // band tables of the current group
    L.res = o; o += up((size_t)p.batch * p.nband_total * 2 * sizeof(double));     // [gal][col][comp] scaled fluxes
    L.rfres = o; o += up((size_t)p.batch * p.nrf * 2 * sizeof(double));           // [gal][rf][comp] rest-frame fluxes
    L.lb = o; o += up((size_t)p.batch * 2 * nl * 2 * sizeof(int));                // line windows [b0, b1]
    L.lt = o; o += up((size_t)p.batch * 2 * nl * 2 * sizeof(int));                // line tasks {first owned node, task offset}
    L.win = o; o += up((size_t)p.batch * p.nband_total * sizeof(uint2));          // node windows {jlo | jhi << 16, split node | bulge flag << 16}
    L.total = o;
    return L;
}

// Optional per-phase cycle counters (build with -DEGG_PHASE_TIMING: thread 0 of CTAs 0 and 77 prints them)
#ifdef EGG_PHASE_TIMING
#define EGG_PHASE(tag) if (tid == 0) { const long long tn_ = clock64(); ph_t[PH_##tag] += tn_ - ph_last; ph_last = tn_; }
#else
#define EGG_PHASE(tag)
#endif

// ---- the kernel ------------------------------------------------------------------------------------
// One persistent CTA carries a batch of kBatch galaxies through phases A (rest-frame SEDs into a
// per-CTA global scratch that stays in L2), B (for each band group: tables -> shared memory by TMA,
// warps pull (galaxy, band) items off a shared counter and integrate both components) and C (totals,
// data conditions, coalesced stores).
// Scratch of one galaxy: [nD pairs {disk SED, bulge SED resampled on the disk grid}][bulge SED on its own grid].
template <typename real, int NT>
__global__ void __launch_bounds__(NT, 1) flux_kernel(const __grid_constant__ FluxParams p) {
    typedef typename Pair<real>::type real2;
    constexpr int kBatch = sizeof(real) == 4 ? kBatchF32 : kBatchF64;
    // phase A1 hands the galaxies of a batch to fixed thread ranges (one warp lane per galaxy, set-up threads 128 .. 192 + kBatch)
    static_assert(kBatch <= 32 && (kBatch & (kBatch - 1)) == 0 && 192 + kBatch <= NT, "EGG_KBATCH_* must be a power of two <= 32");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ GalScal sc[kBatch];
    __shared__ __align__(8) uint64_t bar_tab;
    __shared__ int item_ctr[kMaxGroups];
    __shared__ int a2a_ctr, a2b_ctr; // task counters of phases A2a and A2b
    __shared__ int jneed[kBatch][2]; // first and last disk-grid node that any band of the galaxy reads
    __shared__ int any_bulge_line; // some bulge of the batch has an emission line with flux (normally none)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const SmemLayout L = smem_layout(p);
    const uint8_t *blob = smem + L.tab;
    double *res = reinterpret_cast<double *>(smem + L.res);
    double *rfres = reinterpret_cast<double *>(smem + L.rfres);
    int *lb = reinterpret_cast<int *>(smem + L.lb); // [gal][comp][line][2]
    int *lt = reinterpret_cast<int *>(smem + L.lt); // [gal][comp][line][2]
    uint2 *win = reinterpret_cast<uint2 *>(smem + L.win); // [band in group order][gal]: one 8-byte record per item
    const BandHeader *hdr = reinterpret_cast<const BandHeader *>(blob);
    const int ncomp = p.has_bulge ? 2 : 1;
    const bool cs17 = p.ir_kind == EGGPHOT_IR_CS17;
    const int nbt = (int)p.nband_total, nrf = p.do_rf ? (int)p.nrf : 0;
    const int nD = p.nD, nB = p.nB, nline = p.nline;
    const int sstride = 2 * p.strideD + p.strideB;
    real *scr = reinterpret_cast<real *>(p.scratch) + (size_t)blockIdx.x * kBatch * sstride;
    const real *rO = reinterpret_cast<const real *>(p.rowsDopt), *rI = reinterpret_cast<const real *>(p.rowsDir);
    const real *rP = reinterpret_cast<const real *>(p.rowsDpah), *rB = reinterpret_cast<const real *>(p.rowsB);
    const GridA *gridA = reinterpret_cast<const GridA *>(p.gridA);

    if (tid == 0) {
        mbar_init(&bar_tab, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t tab_phase = 0;
    int loaded = -1;

#ifdef EGG_PHASE_TIMING
    enum { PH_A1, PH_A2a_win, PH_Blines, PH_A2b, PH_A3lines, PH_TMA, PH_B, PH_C, PH_N };
    long long ph_t[PH_N] = {0}, ph_last = clock64();
#endif
    const uint64_t nbatch = (p.ngal + kBatch - 1) / kBatch;
    for (uint64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const uint64_t g0 = batch * kBatch;
        const int ng = (int)min((uint64_t)kBatch, p.ngal - g0);

        // ---- phase A1: per-galaxy scalars, four threads per galaxy (one per piece), spread over four warps
        if ((tid & 31) < kBatch && (tid >> 5) < 4 && (tid & 31) < ng) gal_scalars(p, g0 + (tid & 31), sc[tid & 31], tid >> 5);
        if (tid >= 128 && tid < 128 + kMaxGroups) item_ctr[tid - 128] = 0;
        if (tid == 160) { any_bulge_line = 0; a2a_ctr = 0; a2b_ctr = 0; }
        if (tid >= 192 && tid < 192 + kBatch) { // the rest-frame bands read fixed node ranges
            int lo = nD, hi = -1;
            for (int r = 0; r < nrf; ++r) { lo = min(lo, (int)p.rfD[r].j0); hi = max(hi, (int)(p.rfD[r].j0 + p.rfD[r].n - 1)); }
            jneed[tid - 192][0] = lo; jneed[tid - 192][1] = hi;
        }
        for (int i = tid; i < ng * nbt * 2; i += NT) res[i] = 0.0;
        __syncthreads();
        EGG_PHASE(A1)

        // node windows of every (band, galaxy): one thread each
        for (int idx = tid; idx < nbt * kBatch; idx += NT) {
            const int gi = idx & (kBatch - 1), fb = idx / kBatch;
            uint32_t wv = 0;
            uint16_t wm = 0;
            uint8_t fl = 0;
            if (gi < ng) {
                int grp = 0;
                while (grp + 1 < (int)p.ngroups && fb >= (int)p.groups[grp + 1].base) ++grp;
                const BandHeader *gh = reinterpret_cast<const BandHeader *>(p.groups[grp].blob) + (fb - (int)p.groups[grp].base);
                const double full_first = __ldg(&gh->full_first), full_last = __ldg(&gh->full_last);
                const GalScal &s = sc[gi];
                const double *x = p.xD;
                const int n = nD;
                // [vif] sed2flux: NaN (-> flux 0, src/egg-gencat.cpp:2198) unless the SED covers the filter
                const bool covered = n >= 2 && __ldg(x) * s.opz <= full_first && __ldg(x + n - 1) * s.opz >= full_last;
                if (covered && __ldg(&gh->n) >= 2) {
                    // rest-frame ends of the filter span; their lookup buckets come from log2(lambda) - log2(1+z)
                    const float koff = (float)p.lutD_e0 * (float)kLutPerOctave + s.l2opz;
                    const int jlo = max(lut_last_le(x, n, p.lutD, p.lutD_n, __ldg(&gh->l2first) - koff, __ldg(&gh->eff_first) * s.iopz * (1.0 - 1e-15)), 0);
                    const int jhi = min(lut_first_ge(x, n, p.lutD, p.lutD_n, __ldg(&gh->l2last) - koff, __ldg(&gh->eff_last) * s.iopz * (1.0 + 1e-15)), n - 1);
                    if (jhi > jlo) {
                        wv = (uint32_t)jlo | ((uint32_t)jhi << 16);
                        atomicMin(&jneed[gi][0], jlo);
                        atomicMax(&jneed[gi][1], jhi);
                        // b = last node of [jlo, jhi) left of the band's split node, by the same expression phase B
                        // uses to place a node in a filter cell (the two must agree on which gauge a node is in)
                        const double fc = __ldg(&gh->fc), fs = __ldg(&gh->f_split);
                        // first guess from the wavelength lookup table, settled with the exact predicate
                        int k = (int)(__ldg(&gh->l2split) - koff);
                        k = max(0, min(k, p.lutD_n - 1));
                        int lo = max(jlo, min((int)__ldg(p.lutD + k), jhi - 1));
                        double xa = __ldg(x + lo), xb = __ldg(x + min(lo + 1, n - 1));
                        while (lo > jlo && !(fma(s.opz, xa, -fc) < fs)) { --lo; xb = xa; xa = __ldg(x + lo); }
                        while (lo + 1 < jhi && fma(s.opz, xb, -fc) < fs) { ++lo; xb = __ldg(x + min(lo + 1, n - 1)); }
                        wm = (uint16_t)lo;
                    }
                    if (p.has_bulge && nB >= 2 && __ldg(p.xB) * s.opz <= full_first && __ldg(p.xB + nB - 1) * s.opz >= full_last) fl = 1;
                }
            }
            win[idx] = make_uint2(wv, (uint32_t)wm | ((uint32_t)fl << 16));
        }
        // emission-line windows ([vif] bounds, src/egg-gencat.cpp:2081-2084): one thread per (galaxy, component, line)
        for (int idx = tid; idx < ng * 2 * nline; idx += NT) {
            const int gi = idx / (2 * nline), rem = idx - gi * 2 * nline;
            const int c = rem >= nline, l = rem - c * nline; // c: 0 bulge, 1 disk
            const GalScal &s = sc[gi];
            const float *ll = c ? p.line_lum_disk : p.line_lum_bulge;
            int b0 = 1, b1 = 0; // empty
            if ((c || p.has_bulge) && ll != nullptr && ll[(g0 + gi) * nline + p.line_col[l]] != 0.0f) {
                const double *x = c ? p.xD : p.xB;
                const int n = c ? nD : nB;
                const uint16_t *lut = c ? p.lutD : p.lutB;
                const int nlut = c ? p.lutD_n : p.lutB_n, e0 = c ? p.lutD_e0 : p.lutB_e0;
                const double l0 = (double)p.line_lambda[l], lw = l0 * (s.vdisp / 2.998e5);
                const int lo = lut_last_le(x, n, lut, nlut, lut_pos(l0 - 10 * lw, e0), l0 - 10 * lw);
                const int hi = lut_last_le(x, n, lut, nlut, lut_pos(l0 + 10 * lw, e0), l0 + 10 * lw) + 1; // first > b, n if none
                if (!(hi == 0 || lo == n - 1)) {
                    b0 = lo < 0 ? 0 : lo;
                    b1 = hi >= n ? n - 1 : hi;
                }
            }
            lb[2 * idx] = b0;
            lb[2 * idx + 1] = b1;
            if (!c && b0 <= b1) any_bulge_line = 1;
        }
        __syncthreads();
        EGG_PHASE(A2a_win)

        // ---- phase A3: emission lines (add_emission_line, src/egg-gencat.cpp:2078-2095), one warp per (galaxy,
        //      component).  Lines are sorted by wavelength, so their node windows [b0, b1] are ordered as well.
        //      The first line whose window holds node j owns it and adds the contributions of all lines at j,
        //      in that order: no two lanes touch the same node and the sum order is fixed.  The owned nodes of
        //      all lines are numbered consecutively (prefix sum) and dealt to the lanes.
        auto deposit_lines = [&](const int c) { // c: 0 bulge (before it is resampled on the disk grid), 1 disk
            const float *llp = c ? p.line_lum_disk : p.line_lum_bulge;
            // two warps per galaxy: both number the owned nodes (identical results) and take alternate 32-node slices
            for (int ht = warp; ht < 2 * ng; ht += NW) {
                const int gi = ht >> 1, half = ht & 1;
                const int task = 2 * gi + c;
                const int *wb = lb + 2 * task * nline;
                int *wt = lt + 2 * task * nline;
                int carry_b1 = -1, total = 0;
                for (int base = 0; base < nline; base += 32) {
                    const int l = base + lane;
                    int b0 = 1, b1 = 0;
                    if (l < nline) { b0 = wb[2 * l]; b1 = wb[2 * l + 1]; }
                    const bool ne = b0 <= b1;
                    int m = ne ? b1 : -1; // inclusive running maximum of b1 over the lines so far
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(kFull, m, o);
                        if (lane >= o) m = max(m, t);
                    }
                    int prev = __shfl_up_sync(kFull, m, 1);
                    prev = max(lane ? prev : -1, carry_b1);
                    const int start = max(b0, prev + 1);
                    const int cnt = ne ? max(0, b1 - start + 1) : 0;
                    int inc = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(kFull, inc, o);
                        if (lane >= o) inc += t;
                    }
                    if (l < nline) { wt[2 * l] = start; wt[2 * l + 1] = total + inc - cnt; }
                    total += __shfl_sync(kFull, inc, 31);
                    carry_b1 = max(carry_b1, __shfl_sync(kFull, m, 31));
                }
                __syncwarp();
                const GalScal &s = sc[gi];
                const float *ll = llp + (g0 + gi) * nline;
                const double *x = c ? p.xD : p.xB;
                const int n = c ? nD : nB;
                real *S = scr + (size_t)gi * sstride + (c ? 0 : 2 * p.strideD);
                const int sstep = c ? 2 : 1;
                // the bulge SED carries no mass scale (applied at the end), so its line term is divided by it
                const double inv = c ? 1.0 : (s.a_bulge > 0 ? 1.0 / s.a_bulge : 0.0);
                for (int t = lane + 32 * half; t < total; t += 64) {
                    int lo = 0, hi = nline; // last line whose task offset is <= t and that owns nodes
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (wt[2 * mid + 1] <= t) lo = mid; else hi = mid;
                    }
                    const int l = lo, j = wt[2 * l] + (t - wt[2 * l + 1]);
                    const real s_old = S[(size_t)j * sstep]; // requested with the wavelengths: one round trip, not two
                    const double xj = __ldg(x + j);
                    const double xm = j > 0 ? __ldg(x + j - 1) : 0.0, xp = j < n - 1 ? __ldg(x + j + 1) : 0.0;
                    const double llow = j != 0 ? 0.5 * (xm + xj) : xj - 0.5 * (xp - xj);       // :2090
                    const double lup = j != n - 1 ? 0.5 * (xj + xp) : xj + 0.5 * (xj - xm);   // :2091
                    double acc = 0.0;
                    for (int k = l; k < nline; ++k) {
                        const int k0 = wb[2 * k], k1 = wb[2 * k + 1];
                        if (k0 > k1) { // empty window; the past-the-end marker of the skipped red lines ends the scan
                            if (k0 >= n) break;
                            continue;
                        }
                        if (k0 > j) break;     // windows are ordered: no later line reaches back to j
                        if (j <= k1) {
                            const double l0 = (double)p.line_lambda[k], lw = l0 * (s.vdisp / 2.998e5);
                            acc += line_bin_t<real>(llow, lup, l0, lw, (double)ll[p.line_col[k]] * l0, p.line_average);
                        }
                    }
                    S[(size_t)j * sstep] = s_old + (real)(xj * acc * inv);
                }
            }
        };
        // ---- phase A2a: bulge SEDs on their own grid (get_opt_sed + IGM, src/egg-gencat.cpp:708-712, 2122-2140).
        //      Only when some bulge of the batch has an emission line (normally none): the lines are deposited on the
        //      bulge's own grid before it is resampled.  Otherwise phase A2b and the rest-frame items read the
        //      template row itself.  A warp takes 128 consecutive nodes of one galaxy at a time, four per lane.
        const bool use_sb = any_bulge_line != 0; // block-uniform: written before the barrier above
        if (use_sb) {
            const int ncB = (p.strideB + 127) >> 7; // to the stride: the spare element after the last node is zeroed
            for (;;) {
                int task = 0;
                if (lane == 0) task = atomicAdd(&a2a_ctr, 1);
                task = __shfl_sync(kFull, task, 0);
                if (task >= ng * ncB) break;
                const int gi = task / ncB, j0 = ((task - gi * ncB) << 7) + 4 * lane;
                if (j0 >= p.strideB) continue;
                const GalScal &s = sc[gi];
                real *SB = scr + (size_t)gi * sstride + 2 * p.strideD;
                Quad<real> v = ldg4(rB + (size_t)s.row_b * p.strideB + j0);
                if (p.apply_igm && j0 < p.igmB[3]) {
                    const real g0f = s.igm[0], g1f = s.igm[1], g2f = s.igm[2];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        v.v[k] *= igm_factor<real>(j0 + k, p.igmB[0], p.igmB[1], p.igmB[2], p.igmB[3], g0f, g1f, g2f);
                }
                st4(SB + j0, v); // rows are zero-padded to the stride: the tail stores zeros
            }
        }
        if (use_sb) {
            __syncthreads();
            deposit_lines(0);
            __syncthreads();
            EGG_PHASE(Blines)
        }

        // disk lines whose nodes no band of the galaxy reads (far-IR and CO lines beyond the reddest band, mostly) need
        // no deposit: their windows are emptied here, before the line phase reads them (a barrier away)
        for (int idx = tid; idx < ng * nline; idx += NT) {
            const int gi = idx / nline, l = idx - gi * nline;
            int *w = lb + 2 * ((2 * gi + 1) * nline + l);
            if (w[0] <= w[1]) {
                // empty either way; a window that starts past every node also ends the ordered line scans of the deposit
                if (w[0] > jneed[gi][1]) { w[0] = nD; w[1] = nD - 1; }
                else if (w[1] < jneed[gi][0]) { w[0] = 1; w[1] = 0; }
            }
        }

        // ---- phase A2b: disk SEDs (get_opt_sed / get_ir_sed / merge_add / IGM, src/egg-gencat.cpp:2104-2140) and the
        //      bulge SED resampled on the disk grid (the same piece-wise linear function; lets phase B serve both
        //      components from one set of weights), stored as {disk, bulge} pairs
        {
            // a task = 32 x NPL consecutive nodes (16 bytes of a row per lane: four nodes in fp32, two in fp64) for up
            // to four galaxies: the disk -> optical maps of the nodes are loaded once per task (L1 is too small to
            // keep them).  Nodes that are optical nodes themselves carry a zero fraction: fma(0, b1 - b0, b0) = b0.
            constexpr int NPL = 16 / (int)sizeof(real), LG = NPL == 4 ? 7 : 6;
            const uint64_t keep = evict_last_policy(); // scratch lines should outlive the streamed data in L2 (halves the write-backs)
            const real *ofrac = reinterpret_cast<const real *>(p.u_ofrac);
            // Only the nodes some band of the galaxy reads are built (jneed: union of its node windows).
            const int ncD = (nD + 32 * NPL - 1) >> LG, nq = (ng + 3) >> 2;
            for (;;) {
                int task = 0;
                if (lane == 0) task = atomicAdd(&a2b_ctr, 1);
                task = __shfl_sync(kFull, task, 0);
                if (task >= ncD * nq) break;
                const int q = task / ncD, c0 = (task - q * ncD) << LG, j0 = c0 + NPL * lane;
                int glo = nD, ghi = -1; // the quad's union
                for (int gi = 4 * q; gi < min(4 * q + 4, ng); ++gi) { glo = min(glo, jneed[gi][0]); ghi = max(ghi, jneed[gi][1]); }
                if (c0 > ghi || c0 + 32 * NPL <= glo || j0 >= nD) continue;
                int ol[NPL], oa[NPL];
                Vec16<real> fr;
#pragma unroll
                for (int k = 0; k < NPL; ++k) { ol[k] = -1; fr.v[k] = real(0); }
                if (p.has_bulge) {
                    if constexpr (NPL == 4) {
                        const int4 o4 = __ldg(reinterpret_cast<const int4 *>(p.u_oleft + j0));
                        ol[0] = o4.x; ol[1] = o4.y; ol[2] = o4.z; ol[3] = o4.w;
                    } else {
                        const int2 o2 = __ldg(reinterpret_cast<const int2 *>(p.u_oleft + j0));
                        ol[0] = o2.x; ol[1] = o2.y;
                    }
                    fr = ldg16(ofrac + j0);
                }
#pragma unroll
                for (int k = 0; k < NPL; ++k) oa[k] = max(ol[k], 0);
                const bool igm_here = p.apply_igm && j0 < p.igmD[3];
#pragma unroll 1
                for (int gi = 4 * q; gi < min(4 * q + 4, ng); ++gi) {
                    if (c0 > jneed[gi][1] || c0 + 32 * NPL <= jneed[gi][0]) continue; // warp-uniform
                    const GalScal &s = sc[gi];
                    real *DU = scr + (size_t)gi * sstride;
                    const real *SB = DU + 2 * p.strideD;
                    Vec16<real> o, i, pq;
#pragma unroll
                    for (int k = 0; k < NPL; ++k) o.v[k] = i.v[k] = pq.v[k] = real(0);
                    if (p.disk_mode != 2) o = ldg16(rO + (size_t)s.row_d * p.strideD + j0);
                    if (p.disk_mode != 1) i = ldg16(rI + (size_t)s.row_i * p.strideD + j0);
                    if (p.disk_mode != 1 && cs17) pq = ldg16(rP + (size_t)s.row_i * p.strideD + j0);
                    // branch-free gather of the two optical nodes around each disk node (the row has a spare element
                    // after the last node: oa + 1 is always in range, and the last node has a zero fraction)
                    real b0[NPL], b1[NPL];
#pragma unroll
                    for (int k = 0; k < NPL; ++k) b0[k] = b1[k] = real(0);
                    if (p.has_bulge) {
                        if (use_sb) {
#pragma unroll
                            for (int k = 0; k < NPL; ++k) { b0[k] = SB[oa[k]]; b1[k] = SB[oa[k] + 1]; }
                        } else { // the template row itself (get_opt_sed + IGM, src/egg-gencat.cpp:708-712, 2122-2140)
                            const real *R = rB + (size_t)s.row_b * p.strideB;
#pragma unroll
                            for (int k = 0; k < NPL; ++k) { b0[k] = __ldg(R + oa[k]); b1[k] = __ldg(R + oa[k] + 1); }
                            if (p.apply_igm && oa[0] < p.igmB[3]) {
                                const real g0f = s.igm[0], g1f = s.igm[1], g2f = s.igm[2];
#pragma unroll
                                for (int k = 0; k < NPL; ++k) {
                                    b0[k] *= igm_factor<real>(oa[k], p.igmB[0], p.igmB[1], p.igmB[2], p.igmB[3], g0f, g1f, g2f);
                                    b1[k] *= igm_factor<real>(oa[k] + 1, p.igmB[0], p.igmB[1], p.igmB[2], p.igmB[3], g0f, g1f, g2f);
                                }
                            }
                        }
                    }
                    const real a_d = (real)s.a_disk, c1 = (real)s.c1, c2 = (real)s.c2;
                    real v[NPL], w[NPL];
#pragma unroll
                    for (int k = 0; k < NPL; ++k) {
                        v[k] = a_d * o.v[k] + c1 * i.v[k]; // absent terms are zero
                        if (p.disk_mode != 1 && cs17) v[k] += c2 * pq.v[k];
                    }
                    if (igm_here) {
                        const real g0f = s.igm[0], g1f = s.igm[1], g2f = s.igm[2];
#pragma unroll
                        for (int k = 0; k < NPL; ++k)
                            v[k] *= igm_factor<real>(j0 + k, p.igmD[0], p.igmD[1], p.igmD[2], p.igmD[3], g0f, g1f, g2f);
                    }
#pragma unroll
                    for (int k = 0; k < NPL; ++k) w[k] = ol[k] < 0 ? real(0) : fma(fr.v[k], b1[k] - b0[k], b0[k]);
                    // the row stride is a multiple of four: nodes j0 .. j0 + NPL - 1 <= stride - 1 always have room
                    Vec16<real> st;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int k = 0; k < NPL / 2; ++k) { st.v[2 * k] = v[h * (NPL / 2) + k]; st.v[2 * k + 1] = w[h * (NPL / 2) + k]; }
                        st16_keep(DU + 2 * j0 + h * NPL, st, keep);
                    }
                }
            }
        }
        __syncthreads();
        EGG_PHASE(A2b)
        if (nline > 0 && p.line_lum_disk != nullptr) {
            deposit_lines(1);
            __syncthreads();
            EGG_PHASE(A3lines)
        }

        // ---- phase B: band groups
        // consecutive batches walk the groups in opposite orders: the tables of the group a batch ends with are
        // still in shared memory when the next batch starts with it (one table load less per batch)
        const int ngl = max((int)p.ngroups, nrf > 0 ? 1 : 0);
        const bool rev = ((batch - blockIdx.x) / gridDim.x) & 1;
        for (int gstep = 0; gstep < ngl; ++gstep) {
            const int grp = rev ? ngl - 1 - gstep : gstep;
            const int nb = grp < (int)p.ngroups ? (int)p.groups[grp].nb : 0;
            if (nb > 0 && loaded != grp) {
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
                EGG_PHASE(TMA)
            }
            const int nband_items = nb * kBatch;
            const int nitems = nband_items + (grp == 0 ? nrf * ncomp * kBatch : 0);
            for (;;) {
                int idx = 0;
                if (lane == 0) idx = atomicAdd(&item_ctr[grp], 1);
                idx = __shfl_sync(kFull, idx, 0);
                if (idx >= nitems) break;
                if (idx < nband_items) {
                    // (band, galaxy slot), both components: bands are stored in decreasing order of expected work
                    const int gi = idx & (kBatch - 1), bi = idx / kBatch;
                    if (gi >= ng) continue;
                    const int widx = ((int)p.groups[grp].base + bi) * kBatch + gi;
                    const uint2 wrec = win[widx];
                    const uint32_t wv = wrec.x;
                    if (wv == 0) continue; // not covered, empty response or empty window: flux stays 0
                    const int jlo = (int)(wv & 0xffffu), jhi = (int)(wv >> 16), b = (int)(wrec.y & 0xffffu);
                    const bool bulge_on = (wrec.y >> 16) != 0;
                    const BandHeader &h = hdr[bi];
                    const GalScal &s = sc[gi];
                    BandView<real> bv;
                    make_band_view<real>(blob, h, bv);
                    const real2 *DU = reinterpret_cast<const real2 *>(scr + (size_t)gi * sstride);
                    const double opz = s.opz;
                    const real opzr = (real)opz;
                    // the SED interval [b, b+1] holds the band's split node (phot_core.cuh, two-sided gauge)
                    const double tot0z = bv.tot0 * opz;
                    const double asym_b1 = fma(bv.tot0, fma(opz, __ldg(&gridA[b + 1].x), -bv.fc), bv.asym0);
                    real accd = real(0), accb = real(0);
                    // lane l of a pass sits on node jb + l; lanes 1..30 own their node, lanes 0 and 31 only supply
                    // their neighbours' Lambda / E
                    // the node data of the next pass are requested before this pass's arithmetic: they come from L2
                    // (grid arrays, scratch) and their latency would otherwise head every pass's dependency chain
                    struct NodeIn { GridA g; real2 su; real dxo_or_dx, dxo; };
                    auto fetch = [&](int jb_, NodeIn &n) {
                        const int jc_ = min(max(jb_ + lane, jlo), jhi);
                        n.g = ldg_grid(gridA + jc_);
                        n.su = DU[jc_];
                        if constexpr (sizeof(real) == 4) {
                            n.dxo = __ldg(reinterpret_cast<const float *>(p.gridB) + jc_);
                            n.dxo_or_dx = real(0);
                        } else {
                            const GridBD gb = ldg_grid(reinterpret_cast<const GridBD *>(p.gridB) + jc_);
                            n.dxo_or_dx = gb.dx; n.dxo = gb.dxo;
                        }
                    };
                    NodeIn cur, nxt;
                    // fp32 mode: lanes 1..30 of a pass own their node and take E and gamma of the left neighbour by shuffle
                    // (direct form, sum_j S_j W_j, float products: every term has the sign of the flux).  fp64 mode:
                    // summation by parts, sum_j (E_j + g_j / 2) (S_j - S_{j+1}): a lane needs its right neighbour only
                    // (two shuffles fewer, 31 owned lanes); the terms are larger than their sum, which double carries
                    // (in fp32 the float -> double conversions it needs cost what the shuffles save).
                    constexpr bool kByParts = sizeof(real) == 8;
                    constexpr int kOwned = kByParts ? 31 : 30;
                    const int jb0 = kByParts ? jlo : jlo - 1;
                    fetch(jb0, cur);
                    for (int jb = jb0; jb < jhi; jb += kOwned) {
                        const int j = jb + lane;
                        if (jb + kOwned < jhi) fetch(jb + kOwned, nxt);
                        const GridA g = cur.g;
                        const real2 su = cur.su;
                        real dt, dto, gidx; // observed widths to the next disk / optical node; 1/dx in the mode's type
                        if constexpr (sizeof(real) == 4) {
                            gidx = (float)g.idx;
                            dt = opzr * rcp_approx(gidx); // 1/idx = x[j+1] - x[j]; inf at the grid's last node (unused there)
                            dto = cur.dxo * opzr;
                        } else {
                            gidx = g.idx; dt = cur.dxo_or_dx * opzr; dto = cur.dxo * opzr;
                        }
                        const double t = fma(opz, g.x, -bv.fc);
                        const int cs = cell_slot(find_cell_warp<real>(bv, t));
                        const Pair2<real> sd = bv.sdl[cs];
                        const bool inner = j > jlo && j < jhi;
                        const double u = t - bv.f[cs];
                        const LamK0 lk = bv.lk[cs];
                        const double lam = inner ? fma(u, fma(u, (double)bv.hr[cs], lk.k0), lk.lam) : 0.0; // 0 at both window ends
                        const real ur = inner ? (real)u : real(0);
                        const real sr = sd.x, dlr = sd.y;
                        if constexpr (kByParts) {
                            const real sxn = __shfl_down_sync(kFull, su.x, 1);
                            double dlam = __shfl_down_sync(kFull, lam, 1) - lam;
                            if (j == b) dlam += asym_b1;
                            const double E = dlam * g.idx;
                            const bool own = lane <= 30 && j < jhi; // j >= jlo always
                            const real gD = gamma_term<real>(sr, dlr, ur, dt) * gidx;
                            accd = fma(E + 0.5 * gD, own ? su.x - sxn : real(0), accd);
                            // the interval that holds the split node: W_{b+1} takes E_b in the right gauge, E_b - tot0 (1+z)
                            if (j == b && own) accd = fma(sxn, tot0z, accd);
                            if (bulge_on) { // warp-uniform
                                const real syn = __shfl_down_sync(kFull, su.y, 1);
                                const real gB = gamma_term<real>(sr, dlr, ur, dto) * gidx;
                                accb = fma(E + 0.5 * gB, own ? su.y - syn : real(0), accb);
                                if (j == b && own) accb = fma(syn, tot0z, accb);
                            }
                        } else {
                            double dlam = __shfl_down_sync(kFull, lam, 1) - lam;
                            if (j == b) dlam += asym_b1;
                            const double E = dlam * g.idx; // 0 left of jlo and from jhi on: Lambda is 0 on both sides there
                            double Ep = __shfl_up_sync(kFull, E, 1);
                            if (j == b + 1) Ep -= tot0z;
                            const real W = (real)(E - Ep);
                            const bool own = lane >= 1 && lane <= 30 && j <= jhi;
                            // the in-cell terms by interval: lanes 0..29 own [j, j+1] and add g_j / 2 (S_j - S_{j+1}); the
                            // difference of neighbouring SED values is exact in float, so a sparse bulge grid under a
                            // coarse filter does not amplify rounding (1.4e-5 on a three-node 35 um band before)
                            const bool owni = lane <= 29 && j < jhi;
                            const real gD = gamma_term<real>(sr, dlr, ur, dt) * gidx;
                            const real sxn = __shfl_down_sync(kFull, su.x, 1);
                            accd = fma(su.x, own ? W : real(0), accd);
                            accd = fma(owni ? real(0.5) * gD : real(0), su.x - sxn, accd);
                            if (bulge_on) {
                                const real gB = gamma_term<real>(sr, dlr, ur, dto) * gidx;
                                const real syn = __shfl_down_sync(kFull, su.y, 1);
                                accb = fma(su.y, own ? W : real(0), accb);
                                accb = fma(owni ? real(0.5) * gB : real(0), su.y - syn, accb);
                            }
                        }
                        cur = nxt;
                    }
                    if (bulge_on) {
                        // both totals in five shuffle steps: the halves of the warp first swap what the other needs,
                        // then lanes 0-15 reduce the disk and lanes 16-31 the bulge (same pairs, same order of
                        // additions as warp_total: bit-identical)
                        const bool lo = lane < 16;
                        double v = (lo ? (double)accd : (double)accb) + __shfl_xor_sync(kFull, lo ? (double)accb : (double)accd, 16);
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
                        if ((lane & 15) == 0) {
                            double f = v * s.iopz * s.aflux * (lo ? 1.0 : s.a_bulge);
                            if (!isfinite(f)) f = 0.0; // src/egg-gencat.cpp:2198
                            res[(gi * nbt + (int)h.out_col) * 2 + (lo ? 1 : 0)] = f;
                        }
                    } else {
                        const double totd = warp_total((double)accd);
                        if (lane == 0) {
                            double f = totd * s.iopz * s.aflux;
                            if (!isfinite(f)) f = 0.0;
                            res[(gi * nbt + (int)h.out_col) * 2 + 1] = f;
                        }
                    }
                } else {
                    // rest-frame magnitude: dot product with the z = 0 weights (src/egg-gencat.cpp:2150-2158)
                    const int r = idx - nband_items;
                    const int gi = r & (kBatch - 1), t2 = r / kBatch;
                    if (gi >= ng) continue;
                    const int comp = p.has_bulge ? (t2 & 1) : 1, rb = p.has_bulge ? (t2 >> 1) : t2;
                    const RfDesc dsc = comp ? p.rfD[rb] : p.rfB[rb];
                    const real *S = scr + (size_t)gi * sstride + (comp ? 0 : 2 * p.strideD);
                    const int sstep = comp ? 2 : 1;
                    const real *W = reinterpret_cast<const real *>(p.rfW) + dsc.off;
                    CompSum<real> acc;
                    acc.clear();
                    if (comp || use_sb) {
                        for (int k = lane; k < dsc.n; k += 32) acc.add(S[(size_t)(dsc.j0 + k) * sstep] * __ldg(W + k));
                    } else { // no bulge SED in the scratch: the template row with the IGM factors
                        const GalScal &s = sc[gi];
                        const real *R = rB + (size_t)s.row_b * p.strideB;
                        const real g0f = s.igm[0], g1f = s.igm[1], g2f = s.igm[2];
                        for (int k = lane; k < dsc.n; k += 32) {
                            const int j = dsc.j0 + k;
                            real v = __ldg(R + j);
                            if (p.apply_igm && j < p.igmB[3]) v *= igm_factor<real>(j, p.igmB[0], p.igmB[1], p.igmB[2], p.igmB[3], g0f, g1f, g2f);
                            acc.add(v * __ldg(W + k));
                        }
                    }
                    acc = warp_reduce(acc);
                    if (lane == 0)
                        rfres[(gi * nrf + rb) * 2 + comp] = dsc.covered ? ((double)acc.s + (double)acc.c) * (comp ? 1.0 : sc[gi].a_bulge)
                                                                         : __longlong_as_double(0x7ff8000000000000LL);
                }
            }
            __syncthreads(); // every item of the group is done; the table region may be overwritten
            EGG_PHASE(B)
        }

        // ---- phase C: totals (src/egg-gencat.cpp:2238-2239) and coalesced stores
        for (int i = tid; i < ng * nbt; i += NT) {
            const float fb = (float)res[2 * i], fd = (float)res[2 * i + 1];
            const size_t o = (size_t)g0 * nbt + i;
            // streaming stores: the catalog columns are written once and must not evict the scratch from L2
            if (p.flux_disk) __stcs(p.flux_disk + o, fd);
            if (p.flux_bulge) __stcs(p.flux_bulge + o, fb);
            if (p.flux) __stcs(p.flux + o, fd + fb); // vec2f sum
            if (p.flux_disk_f64) __stcs(p.flux_disk_f64 + o, res[2 * i + 1]);
            if (p.flux_bulge_f64) __stcs(p.flux_bulge_f64 + o, res[2 * i]);
        }
        for (int i = tid; i < ng * nrf; i += NT) {
            // [vif] lsun2uJy(0, 1e-5, ...) then uJy2mag; non-finite -> +99 (src/egg-gencat.cpp:2152-2156)
            const double arf = EGGPHOT_LSUN2UJY / (1e-5 * 1e-5);
            double mg[2] = {0.0, 0.0}; // the bulge column stays 0 when there is no bulge pass (no_stellar)
            for (int comp = p.has_bulge ? 0 : 1; comp < 2; ++comp) {
                const double m = -2.5 * log10(rfres[2 * i + comp] * arf) + 23.9;
                mg[comp] = isfinite(m) ? m : 99.0;
            }
            const size_t o = (size_t)g0 * nrf + i;
            const float md = (float)mg[1], mb = (float)mg[0];
            if (p.rfmag_disk) __stcs(p.rfmag_disk + o, md);
            if (p.rfmag_bulge) __stcs(p.rfmag_bulge + o, mb);
            if (p.rfmag) // uJy2mag(mag2uJy(disk) + mag2uJy(bulge)) on the f32 columns (src/egg-gencat.cpp:2239)
                __stcs(p.rfmag + o, (float)(-2.5 * log10(exp10(0.4 * (23.9 - (double)md)) + exp10(0.4 * (23.9 - (double)mb))) + 23.9));
            if (p.rfmag_disk_f64) __stcs(p.rfmag_disk_f64 + o, mg[1]);
            if (p.rfmag_bulge_f64) __stcs(p.rfmag_bulge_f64 + o, mg[0]);
        }
        __syncthreads(); // res / sc are rewritten by the next batch
        EGG_PHASE(C)
    }
#ifdef EGG_PHASE_TIMING
    if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 77))
        printf("PHASES cta %d: A1 %lld win+A2a %lld bulge-lines %lld A2b %lld A3 %lld TMA %lld B %lld C %lld\n", (int)blockIdx.x, ph_t[PH_A1],
               ph_t[PH_A2a_win], ph_t[PH_Blines], ph_t[PH_A2b], ph_t[PH_A3lines], ph_t[PH_TMA], ph_t[PH_B], ph_t[PH_C]);
#endif
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
        if (ll != nullptr && ll[g * p.nline + p.line_col[l]] != 0.0f) {
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
                        acc += line_bin(llow, lup, l0, lw, (double)ll[g * p.nline + p.line_col[k]] * l0, p.line_average);
                    }
                S[j] += xj * acc;
            }
        }
    }
    __syncthreads();
    float *lam = q.lam + (size_t)blockIdx.x * q.row_stride, *sed = q.sed + (size_t)blockIdx.x * q.row_stride;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        lam[j] = (float)(x[j] * s.opz);      // src/egg-gencat.cpp:2161
        sed[j] = (float)(S[j] * s.aflux);    // :2162  (S = lambda_rest * L)
    }
}

} // namespace

size_t flux_kernel_smem(const FluxParams &p, int) { return smem_layout(p).total; }

template <typename real> struct Threads;
template <> struct Threads<float> { static constexpr int n = 1024; };  // 64 registers per thread
#ifndef EGG_FP64_THREADS
#define EGG_FP64_THREADS 768
#endif
template <> struct Threads<double> { static constexpr int n = EGG_FP64_THREADS; };  // 80 registers per thread (640 threads: -5 %, 896: -2 %)

// The thread count trades registers per thread against warps in flight; both variants of the fp32
// kernel are built and EGGPHOT_THREADS (768 | 1024) selects one for experiments (default 1024).
static int fp32_threads() {
    static int nt = [] {
        const char *e = getenv("EGGPHOT_THREADS");
        const int v = e ? atoi(e) : 0;
        return v == 768 ? 768 : 1024;
    }();
    return nt;
}

template <typename real, int NT> static int launch_nt(const FluxParams &p, size_t smem, int blocks, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(flux_kernel<real, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    flux_kernel<real, NT><<<blocks, NT, smem, st>>>(p);
    return (int)cudaGetLastError();
}

int launch_flux_kernel(const FluxParams &p, int precision, int grid_blocks, cudaStream_t stream) {
    const size_t smem = flux_kernel_smem(p, precision);
    if (precision == EGGPHOT_FP64) return launch_nt<double, Threads<double>::n>(p, smem, grid_blocks, stream);
    return fp32_threads() == 768 ? launch_nt<float, 768>(p, smem, grid_blocks, stream)
                                 : launch_nt<float, 1024>(p, smem, grid_blocks, stream);
}

template <typename real, int NT> static int occupancy_nt(size_t smem) {
    int nb = 0;
    cudaError_t e = cudaFuncSetAttribute(flux_kernel<real, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, flux_kernel<real, NT>, NT, smem);
    return e == cudaSuccess ? nb : -(int)e;
}

int flux_kernel_max_blocks_per_sm(const FluxParams &p, int precision) {
    const size_t smem = flux_kernel_smem(p, precision);
    if (precision == EGGPHOT_FP64) return occupancy_nt<double, Threads<double>::n>(smem);
    return fp32_threads() == 768 ? occupancy_nt<float, 768>(smem) : occupancy_nt<float, 1024>(smem);
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
