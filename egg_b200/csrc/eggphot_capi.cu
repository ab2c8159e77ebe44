// eggphot_capi.cu — the C ABI of include/eggphot.h: context, device tables, chunked host pipeline.
//
// Replaces the block `if (!no_flux) {...}` of src/egg-gencat.cpp:2024-2240 behind plain-C entry
// points.  No CPU compute path exists here: without a CUDA device every entry point returns
// EGGPHOT_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/eggphot.h"
#include "../../include/eggprops.h"
#include "phot_kernels.cuh"
#include "phot_tables.hpp"

using namespace eggphot;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) {
        if (p && bytes >= n) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        if (n == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    // Blocking copy on the legacy default stream.  The kernels run on non-blocking streams, which do not order
    // themselves after it, and for small pageable sources the call may return before the DMA has landed: callers
    // synchronise the device once after their last upload (eggphot_create) or use upload_async on the kernel's stream.
    cudaError_t upload(const void *src, size_t n) {
        cudaError_t e = alloc(n);
        if (e != cudaSuccess || n == 0) return e;
        return cudaMemcpy(p, src, n, cudaMemcpyHostToDevice);
    }
    cudaError_t upload_async(const void *src, size_t n, cudaStream_t st) {
        cudaError_t e = alloc(n);
        if (e != cudaSuccess || n == 0) return e;
        return cudaMemcpyAsync(p, src, n, cudaMemcpyHostToDevice, st);
    }
};

constexpr int kSlots = 4;   // staging slots of the host-memory entry point (chunks in flight: copy in / compute / copy out)
constexpr int kLanesC = 2;  // compute streams (consecutive chunks alternate: the tail of one chunk's kernels overlaps the next's)

// device staging of one chunk of galaxies for the host-memory entry point
struct Slot {
    cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr; // inputs landed / kernels done / outputs back on the host
    DevBuf z, d, m_disk, m_bulge, m2lcor, lir, fpah, mdust, igm_abs, vdisp, lld, llb, osd, osb, irs;
    DevBuf flux, flux_disk, flux_bulge, rfmag, rfmag_disk, rfmag_bulge, fd64, fb64, rd64, rb64;
};

} // namespace

struct eggphot_ctx {
    std::string err;
    Prepared P;
    int device = 0, num_sms = 0;
    bool ready = false;
    cudaStream_t stream = nullptr;
    // prepared tables on the device
    std::vector<DevBuf> blobs;
    DevBuf xB, xD, gridA, gridB, lutB, lutD, rowsB, rowsDopt, rowsDir, rowsDpah, opt_use, rfB, rfD, rfW, errflag, scratch, u_oleft, u_ofrac;
    int strideB = 0, strideD = 0;
    FluxParams base{};
    int grid_blocks = 0;
    size_t scratch_region_bytes = 0;
    uint64_t launches = 0;
    uint64_t table_bytes = 0;
    Slot slots[kSlots];
    cudaStream_t s_in = nullptr, s_out = nullptr, s_c[kLanesC] = {nullptr, nullptr};
    size_t chunk = 0;
    // the lane-per-galaxy kernel (phot_lane_kernels.cu): its own band groups, grid records, scratch and, per scratch
    // region, the redshift-bucketing buffers.  EGGPHOT_KERNEL=node selects the node-per-lane kernel of phot_kernels.cu.
    bool use_lane = true;
    std::vector<DevBuf> blobs2;
    DevBuf gridN, nextOpt, tabCol, colSpan, altFl, altFr, altOff;
    // save_sed staging: the columns of one range of galaxies and two output buffers (the copy out of one range overlaps the
    // kernel of the next), kept between calls
    DevBuf sedCols[2][15], sedOut[2], sedLamTmp;
    cudaEvent_t sedDone[2] = {nullptr, nullptr};
    int sedSlot = 0;
    // events around the walk kernel launches of the device entry point since the last eggphot_get_stats (at most 256 pairs)
    std::vector<cudaEvent_t> walkEv;
    size_t walkUsed = 0, walkCalls = 0;
    bool walk_split = false;
    DevBuf zhist[kLanesC + 1], zperm[kLanesC + 1];
    // per region: SED rows, band-table windows, build ranges, table sums and item counters of one chunk of batches
    DevBuf lrows[kLanesC + 1], lwin[kLanesC + 1], lplan[kLanesC + 1], lres[kLanesC + 1], lcnt[kLanesC + 1];
    LaneParams lbase{};
    size_t lane_cap_batches = 0; // batches per chunk (bounds the SED rows of a region)

    int fail(int code, const std::string &m) { err = m; return code; }
    int cuda_fail(cudaError_t e, const char *what) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return EGGPHOT_ERR_CUDA;
    }
};

namespace {

template <typename T> std::vector<T> pad_rows(const std::vector<double> &rows, size_t nrow, int n, int stride) {
    std::vector<T> out(nrow * (size_t)stride, T(0));
    for (size_t r = 0; r < nrow; ++r)
        for (int j = 0; j < n; ++j) out[r * stride + j] = (T)rows[r * n + j];
    return out;
}

cudaError_t upload_rows(DevBuf &b, const std::vector<double> &rows, size_t nrow, int n, int stride, int precision,
                        uint64_t &bytes) {
    if (rows.empty()) return cudaSuccess;
    cudaError_t e;
    if (precision == EGGPHOT_FP64) {
        auto v = pad_rows<double>(rows, nrow, n, stride);
        e = b.upload(v.data(), v.size() * sizeof(double));
    } else {
        auto v = pad_rows<float>(rows, nrow, n, stride);
        e = b.upload(v.data(), v.size() * sizeof(float));
    }
    bytes += b.bytes;
    return e;
}

cudaError_t upload_grid(const Grid &g, DevBuf &x, DevBuf &lut, uint64_t &bytes) {
    if (g.n() == 0) return cudaSuccess;
    cudaError_t e = x.upload(g.x.data(), g.n() * sizeof(double));
    if (e == cudaSuccess) e = lut.upload(g.lut.data(), g.lut.size() * sizeof(uint16_t));
    bytes += x.bytes + lut.bytes;
    return e;
}

int check_galaxies(eggphot_ctx *c, const eggphot_galaxies *g) {
    const Prepared &P = c->P;
    if (!g || !g->z || !g->d || !g->m2lcor) return c->fail(EGGPHOT_ERR_ARG, "galaxy columns z, d, m2lcor are required");
    if (!g->m_disk && P.disk_mode != DISK_IR) return c->fail(EGGPHOT_ERR_ARG, "m_disk is required");
    if (P.has_bulge && (!g->m_bulge || !g->opt_sed_bulge)) return c->fail(EGGPHOT_ERR_ARG, "m_bulge and opt_sed_bulge are required");
    if (P.disk_mode != DISK_IR && !g->opt_sed_disk) return c->fail(EGGPHOT_ERR_ARG, "opt_sed_disk is required");
    if (P.disk_mode != DISK_OPT) {
        if (!g->ir_sed) return c->fail(EGGPHOT_ERR_ARG, "ir_sed is required");
        if (P.ir_kind == EGGPHOT_IR_CS17 ? (!g->fpah || !g->mdust) : !g->lir)
            return c->fail(EGGPHOT_ERR_ARG, "lir (generic IR library) or fpah+mdust (cs17) are required");
    }
    if (P.apply_igm && !g->igm_abs) return c->fail(EGGPHOT_ERR_ARG, "igm_abs is required unless igm is none/constant");
    if (!P.no_nebular && (!g->vdisp || !g->line_lum_disk))
        return c->fail(EGGPHOT_ERR_ARG, "vdisp and line_lum_disk are required unless no_nebular is set");
    return EGGPHOT_OK;
}

void fill_params(const eggphot_ctx *c, const eggphot_galaxies &g, size_t ngal, const eggphot_outputs &o, FluxParams &p) {
    p = c->base;
    p.z = g.z; p.d = g.d; p.m_disk = g.m_disk; p.m_bulge = g.m_bulge; p.m2lcor = g.m2lcor;
    p.lir = g.lir; p.fpah = g.fpah; p.mdust = g.mdust; p.igm_abs = g.igm_abs; p.vdisp = g.vdisp;
    p.line_lum_disk = c->P.no_nebular ? nullptr : g.line_lum_disk;
    p.line_lum_bulge = c->P.no_nebular ? nullptr : g.line_lum_bulge;
    p.opt_sed_disk = g.opt_sed_disk; p.opt_sed_bulge = g.opt_sed_bulge; p.ir_sed = g.ir_sed;
    p.ngal = ngal;
    p.flux = o.flux; p.flux_disk = o.flux_disk; p.flux_bulge = o.flux_bulge;
    p.rfmag = o.rfmag; p.rfmag_disk = o.rfmag_disk; p.rfmag_bulge = o.rfmag_bulge;
    p.flux_disk_f64 = o.flux_disk_f64; p.flux_bulge_f64 = o.flux_bulge_f64;
    p.rfmag_disk_f64 = o.rfmag_disk_f64; p.rfmag_bulge_f64 = o.rfmag_bulge_f64;
}

int enqueue(eggphot_ctx *c, const eggphot_galaxies &g, size_t ngal, const eggphot_outputs &o, cudaStream_t st, int region) {
    if (ngal == 0) return EGGPHOT_OK;
    FluxParams p;
    fill_params(c, g, ngal, o, p);
    // kernels of different regions may overlap (the two chunk streams of eggphot_compute); each has its own scratch
    p.scratch = (char *)c->scratch.p + (size_t)region * c->scratch_region_bytes;
    p.do_rf = c->P.rfbands.size() && (o.rfmag || o.rfmag_disk || o.rfmag_bulge || o.rfmag_disk_f64 || o.rfmag_bulge_f64);
    if (p.ngroups == 0 && !p.do_rf) return EGGPHOT_OK;
    if (c->use_lane) {
        if (ngal > 0xfffffff0ull) return c->fail(EGGPHOT_ERR_ARG, "more than 2^32 galaxies in one call; split the catalog");
        cudaError_t ce = c->zperm[region].alloc(ngal * sizeof(uint32_t));
        if (ce != cudaSuccess) return c->cuda_fail(ce, "allocating the redshift order");
        LaneParams q = c->lbase;
        const GroupDesc *groups2 = c->lbase.fp.groups;
        const uint32_t ngroups2 = c->lbase.fp.ngroups, max_blob2 = c->lbase.fp.max_blob;
        q.fp = p;
        q.fp.ngroups = ngroups2; q.fp.max_blob = max_blob2;
        for (int k = 0; k < kMaxGroups; ++k) q.fp.groups[k] = groups2[k];
        q.fp.scratch = nullptr;
        q.perm = (const uint32_t *)c->zperm[region].p;
        int e = launch_zsort(g.z, (uint32_t)ngal, (uint32_t *)c->zhist[region].p, (uint32_t *)c->zperm[region].p, c->num_sms, st);
        if (e != 0) return c->cuda_fail((cudaError_t)e, "redshift bucketing launch");
        c->launches += 4;
        // the call is worked off in chunks of batches whose SED rows fit the region's buffer
        const size_t nbatch = (ngal + kLanes - 1) / kLanes, cap = std::min(nbatch, c->lane_cap_batches);
        const size_t ntab = std::max<uint32_t>(q.ntab, 1);
        ce = c->lrows[region].alloc(cap * q.rows_stride + 65536); // (+ room for the L2 prefetches that run past the last row)
        if (ce == cudaSuccess) ce = c->lwin[region].alloc(cap * 128 * sizeof(uint32_t));
        if (ce == cudaSuccess) ce = c->lplan[region].alloc(cap * sizeof(int4));
        if (ce == cudaSuccess) ce = c->lres[region].alloc(cap * ntab * 2 * kLanes * sizeof(double));
        if (ce == cudaSuccess) ce = c->lcnt[region].alloc(kMaxGroups * sizeof(uint32_t));
        if (ce != cudaSuccess) return c->cuda_fail(ce, "allocating the chunk buffers of the flux pipeline");
        q.rows = c->lrows[region].p;
        q.win = (uint32_t *)c->lwin[region].p;
        q.plan = (int4 *)c->lplan[region].p;
        q.res = (double *)c->lres[region].p;
        q.counters = (uint32_t *)c->lcnt[region].p;
        for (size_t b0 = 0; b0 < nbatch; b0 += cap) {
            q.batch0 = b0;
            q.nbatch = (uint32_t)std::min(cap, nbatch - b0);
            cudaEvent_t w0 = nullptr, w1 = nullptr;
            if (region == 0 && c->walkUsed + 2 <= 512) { // (the device entry point: one stream, events in order)
                while (c->walkEv.size() < c->walkUsed + 2) { cudaEvent_t ev; if (cudaEventCreate(&ev) != cudaSuccess) break; c->walkEv.push_back(ev); }
                if (c->walkEv.size() >= c->walkUsed + 2) { w0 = c->walkEv[c->walkUsed]; w1 = c->walkEv[c->walkUsed + 1]; c->walkUsed += 2; }
            }
            // The walk kernel is persistent, one CTA per SM.  The chunks of the host entry point run on two streams: giving
            // each walk half of the SMs lets two chunks' walks run side by side, so that the tail of one (its last items, its
            // group changes) is covered by the body of the other instead of idling the device.
            int walk_ctas = c->num_sms;
            if (region != 0 && c->walk_split) walk_ctas = std::max(1, c->num_sms / 2);
            e = launch_lane_chunk(q, c->P.precision, walk_ctas, st, w0, w1);
            if (e != 0) return c->cuda_fail((cudaError_t)e, "flux pipeline launch");
            c->launches += (q.fp.ngroups > 0 && q.ntab > 0) ? 4 : 3;
        }
        return EGGPHOT_OK;
    }
    const size_t nbatch = (ngal + p.batch - 1) / p.batch;
    const int blocks = (int)std::min<size_t>((size_t)c->grid_blocks, nbatch);
    int e = launch_flux_kernel(p, c->P.precision, blocks, st);
    if (e != 0) return c->cuda_fail((cudaError_t)e, "flux kernel launch");
    ++c->launches;
    return EGGPHOT_OK;
}

int read_error_flag(eggphot_ctx *c) {
    int flag = 0;
    cudaError_t e = cudaMemcpy(&flag, c->errflag.p, sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return c->cuda_fail(e, "reading the error flag");
    if (flag) {
        int zero = 0;
        cudaMemcpy(c->errflag.p, &zero, sizeof(int), cudaMemcpyHostToDevice);
        cudaDeviceSynchronize(); // the reset must have landed before a kernel on a non-blocking stream can set it again
        if (flag == EGGPHOT_ERR_CUDA) return c->fail(flag, "internal error: a wait inside the flux kernel ran into its spin limit");
        return c->fail(flag, "a galaxy points to a template outside the library or to an unusable optical SED (use = false)");
    }
    return EGGPHOT_OK;
}

} // namespace

extern "C" {

int eggphot_create(const eggphot_libs *libs, const eggphot_filters *bands, const eggphot_filters *rfbands,
                   const eggphot_opts *opts, eggphot_ctx **out) {
    if (!out) return EGGPHOT_ERR_ARG;
    eggphot_ctx *c = new (std::nothrow) eggphot_ctx();
    *out = c;
    if (!c) return EGGPHOT_ERR_ARG;
    if (!libs || !opts) return c->fail(EGGPHOT_ERR_ARG, "libs and opts are required");
    Status st = prepare(*libs, bands, rfbands, *opts, c->P);
    if (!st.ok()) return c->fail(st.code, st.msg);
    const Prepared &P = c->P;
    if (P.rfbands.size() > 16) return c->fail(EGGPHOT_ERR_UNSUPPORTED, "more than 16 rfbands");

    // ---- device
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return c->fail(EGGPHOT_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0") +
                                           " (this library has no CPU path)");
    c->device = opts->device;
    if ((e = cudaSetDevice(c->device)) != cudaSuccess) return c->cuda_fail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, c->device)) != cudaSuccess) return c->cuda_fail(e, "cudaGetDeviceProperties");
    if (prop.major < 10) return c->fail(EGGPHOT_ERR_CUDA, "device is not sm_100 class (kernels are built for sm_100a only)");
    c->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return c->cuda_fail(e, "cudaStreamCreate");
    for (cudaStream_t *st : {&c->s_in, &c->s_out, &c->s_c[0], &c->s_c[1]})
        if ((e = cudaStreamCreateWithFlags(st, cudaStreamNonBlocking)) != cudaSuccess) return c->cuda_fail(e, "cudaStreamCreate");
    for (auto &s : c->slots)
        for (cudaEvent_t *ev : {&s.ev_in, &s.ev_k, &s.ev_out})
            if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return c->cuda_fail(e, "cudaEventCreate");

    // ---- tables
    uint64_t tb = 0;
    const int per16 = 4; // rows are read four elements at a time in both precisions
    // one spare (zero) element after the last bulge node: phase A2b reads the right neighbour of every optical node
    c->strideB = (P.grid_bulge.n() + 1 + per16 - 1) / per16 * per16;
    c->strideD = (P.grid_disk.n() + per16 - 1) / per16 * per16;
#define CK(x) do { if ((e = (x)) != cudaSuccess) return c->cuda_fail(e, #x); } while (0)
    CK(upload_grid(P.grid_bulge, c->xB, c->lutB, tb));
    CK(upload_grid(P.grid_disk, c->xD, c->lutD, tb));
    {   // disk grid records for phase B (+ one padding record: x of node b + 1 is read)
        const Grid &g = P.grid_disk;
        const int n = g.n();
        auto dxo = [&](int j) { return j < (int)P.u_dxo.size() ? P.u_dxo[j] : 0.0; };
        auto dx = [&](int j) { return j + 1 < n ? g.x[j + 1] - g.x[j] : 0.0; };
        std::vector<GridA> ra((size_t)n + 1, GridA{0, 0});
        for (int j = 0; j < n; ++j) ra[j] = GridA{g.x[j], g.idx[j]};
        CK(c->gridA.upload(ra.data(), ra.size() * sizeof(GridA)));
        if (P.precision == EGGPHOT_FP64) {
            std::vector<GridBD> r((size_t)n + 1, GridBD{0, 0});
            for (int j = 0; j < n; ++j) r[j] = GridBD{dx(j), dxo(j)};
            CK(c->gridB.upload(r.data(), r.size() * sizeof(GridBD)));
        } else {
            std::vector<float> r((size_t)n + 1, 0.0f);
            for (int j = 0; j < n; ++j) r[j] = (float)dxo(j);
            CK(c->gridB.upload(r.data(), r.size() * sizeof(float)));
        }
        tb += c->gridA.bytes + c->gridB.bytes;
    }
    CK(upload_rows(c->rowsB, P.rows_bulge_opt, P.nopt_sed, P.grid_bulge.n(), c->strideB, P.precision, tb));
    CK(upload_rows(c->rowsDopt, P.rows_disk_opt, P.nopt_sed, P.grid_disk.n(), c->strideD, P.precision, tb));
    CK(upload_rows(c->rowsDir, P.rows_disk_ir, P.nir_sed, P.grid_disk.n(), c->strideD, P.precision, tb));
    CK(upload_rows(c->rowsDpah, P.rows_disk_pah, P.nir_sed, P.grid_disk.n(), c->strideD, P.precision, tb));
    if (!P.opt_use.empty()) CK(c->opt_use.upload(P.opt_use.data(), P.opt_use.size()));
    if (!P.u_oleft.empty()) {
        // (padded to a multiple of four entries: phase A4 reads them four at a time)
        const size_t n4 = (P.u_oleft.size() + 3) & ~size_t(3);
        std::vector<int32_t> ol(P.u_oleft); ol.resize(n4, -1);
        std::vector<double> fr(P.u_ofrac); fr.resize(n4, 0.0);
        CK(c->u_oleft.upload(ol.data(), ol.size() * sizeof(int32_t)));
        if (P.precision == EGGPHOT_FP64) CK(c->u_ofrac.upload(fr.data(), fr.size() * sizeof(double)));
        else {
            std::vector<float> f(fr.begin(), fr.end());
            CK(c->u_ofrac.upload(f.data(), f.size() * sizeof(float)));
        }
    }
    c->blobs.resize(P.groups.size());
    for (size_t g = 0; g < P.groups.size(); ++g) {
        CK(c->blobs[g].upload(P.groups[g].blob.data(), P.groups[g].blob.size()));
        tb += c->blobs[g].bytes;
    }
    // rest-frame weights
    {
        std::vector<RfDesc> dB(P.rfbands.size()), dD(P.rfbands.size());
        std::vector<double> w;
        auto add = [&](const RfWeights &r, RfDesc &d) {
            d.j0 = r.j0; d.n = (int)r.w.size(); d.off = (int)w.size(); d.covered = r.covered;
            w.insert(w.end(), r.w.begin(), r.w.end());
        };
        for (size_t b = 0; b < P.rfbands.size(); ++b) { add(P.rf_bulge[b], dB[b]); add(P.rf_disk[b], dD[b]); }
        if (!dB.empty()) {
            CK(c->rfB.upload(dB.data(), dB.size() * sizeof(RfDesc)));
            CK(c->rfD.upload(dD.data(), dD.size() * sizeof(RfDesc)));
            if (w.empty()) w.push_back(0.0);
            if (P.precision == EGGPHOT_FP64) CK(c->rfW.upload(w.data(), w.size() * sizeof(double)));
            else {
                std::vector<float> wf(w.begin(), w.end());
                CK(c->rfW.upload(wf.data(), wf.size() * sizeof(float)));
            }
            tb += c->rfW.bytes;
        }
    }
    int zero = 0;
    CK(c->errflag.upload(&zero, sizeof(int)));
    c->table_bytes = tb;

    // ---- launch-invariant parameters
    FluxParams &p = c->base;
    std::memset(&p, 0, sizeof(p));
    p.nband_total = (uint32_t)P.bands.size();
    p.nrf = (uint32_t)P.rfbands.size();
    p.nB = P.grid_bulge.n(); p.nD = P.grid_disk.n();
    p.strideB = c->strideB; p.strideD = c->strideD;
    p.batch = batch_of(P.precision);
    p.xB = (const double *)c->xB.p; p.xD = (const double *)c->xD.p;
    p.gridA = c->gridA.p; p.gridB = c->gridB.p;
    p.lutB = (const uint16_t *)c->lutB.p; p.lutD = (const uint16_t *)c->lutD.p;
    p.lutB_n = (int)P.grid_bulge.lut.size(); p.lutD_n = (int)P.grid_disk.lut.size();
    p.lutB_e0 = P.grid_bulge.lut_e0; p.lutD_e0 = P.grid_disk.lut_e0;
    const Grid &gb = P.grid_bulge, &gd = P.grid_disk;
    p.igmB[0] = gb.lz; p.igmB[1] = gb.l0; p.igmB[2] = gb.l1; p.igmB[3] = gb.l2;
    p.igmD[0] = gd.lz; p.igmD[1] = gd.l0; p.igmD[2] = gd.l1; p.igmD[3] = gd.l2;
    p.rowsB = c->rowsB.p; p.rowsDopt = c->rowsDopt.p; p.rowsDir = c->rowsDir.p; p.rowsDpah = c->rowsDpah.p;
    p.opt_use = (const uint8_t *)c->opt_use.p;
    p.u_oleft = (const int32_t *)c->u_oleft.p; p.u_ofrac = c->u_ofrac.p;
    p.nopt_sed = P.nopt_sed; p.nir_sed = P.nir_sed;
    p.ir_kind = P.ir_kind; p.disk_mode = (int)P.disk_mode; p.has_bulge = P.has_bulge; p.apply_igm = P.apply_igm;
    p.nline = P.no_nebular ? 0 : (int)P.line_lambda.size();
    p.line_average = P.line_profile == EGGPHOT_LINE_AVERAGE;
    {   // lines sorted by wavelength: their node windows are then ordered too (see phase A3 of the kernel)
        std::vector<int> ord(p.nline);
        for (int l = 0; l < p.nline; ++l) ord[l] = l;
        std::stable_sort(ord.begin(), ord.end(), [&](int i, int j) { return P.line_lambda[i] < P.line_lambda[j]; });
        for (int l = 0; l < p.nline; ++l) { p.line_lambda[l] = P.line_lambda[ord[l]]; p.line_col[l] = (uint8_t)ord[l]; }
    }
    p.rfB = (const RfDesc *)c->rfB.p; p.rfD = (const RfDesc *)c->rfD.p; p.rfW = c->rfW.p;
    p.error_flag = (int32_t *)c->errflag.p;

    // ---- band groups, occupancy, scratch
    if (P.groups.size() > (size_t)kMaxGroups) return c->fail(EGGPHOT_ERR_UNSUPPORTED, "more than 16 band groups; raise smem_table_budget");
    p.ngroups = (uint32_t)P.groups.size();
    p.max_blob = 0;
    for (size_t g = 0; g < P.groups.size(); ++g) {
        p.groups[g].blob = (const uint8_t *)c->blobs[g].p;
        p.groups[g].bytes = (uint32_t)P.groups[g].blob.size();
        p.groups[g].nb = (uint32_t)P.groups[g].cols.size();
        p.groups[g].base = g ? p.groups[g - 1].base + p.groups[g - 1].nb : 0;
        if (p.groups[g].nb > (uint32_t)kMaxGroupBands) return c->fail(EGGPHOT_ERR_UNSUPPORTED, "more than 64 bands in one group");
        p.max_blob = std::max(p.max_blob, p.groups[g].bytes);
    }
    {
        const size_t smem = flux_kernel_smem(p, P.precision);
        if (smem > (size_t)prop.sharedMemPerBlockOptin)
            return c->fail(EGGPHOT_ERR_UNSUPPORTED, "band tables need " + std::to_string(smem) + " B of shared memory, device offers " +
                                                     std::to_string(prop.sharedMemPerBlockOptin) + "; lower smem_table_budget");
        const int occ = flux_kernel_max_blocks_per_sm(p, P.precision);
        if (occ <= 0) return c->fail(EGGPHOT_ERR_CUDA, "flux kernel cannot be resident (occupancy query failed)");
        c->grid_blocks = c->num_sms * occ;
        const size_t rs = P.precision == EGGPHOT_FP64 ? 8 : 4;
        c->scratch_region_bytes = ((size_t)c->grid_blocks * p.batch * (size_t)(2 * c->strideD + c->strideB) * rs + 255) & ~size_t(255);
        const char *sel = getenv("EGGPHOT_KERNEL");
        c->use_lane = !(sel && (std::string(sel) == "node" || std::string(sel) == "1"));
        if (!c->use_lane) CK(c->scratch.alloc(c->scratch_region_bytes * (kLanesC + 1)));
        p.scratch = c->scratch.p;
    }
    // ---- the lane kernel's tables, scratch and bucketing buffers
    {
        if (P.groups2.size() > (size_t)kMaxGroups) return c->fail(EGGPHOT_ERR_UNSUPPORTED, "more than 16 band groups; raise smem_table_budget");
        {   // {x, 1 / dx} of the two component grids as the walk reads them
            std::vector<double> gx;
            for (const GridNode &g : P.grid_nodes) { gx.push_back(g.x); gx.push_back(g.idxp); }
            CK(c->gridN.upload(gx.data(), gx.size() * sizeof(double)));
            gx.clear();
            for (const GridNode &g : P.grid_nodes_b) { gx.push_back(g.x); gx.push_back(g.idxp); }
            CK(c->nextOpt.upload(gx.data(), gx.size() * sizeof(double)));
        }
        c->blobs2.resize(P.groups2.size());
        LaneParams &q = c->lbase;
        q.fp = c->base;
        q.fp.ngroups = (uint32_t)P.groups2.size();
        q.fp.max_blob = 0;
        for (size_t g = 0; g < P.groups2.size(); ++g) {
            CK(c->blobs2[g].upload(P.groups2[g].blob.data(), P.groups2[g].blob.size()));
            c->table_bytes += c->blobs2[g].bytes;
            q.fp.groups[g].blob = (const uint8_t *)c->blobs2[g].p;
            q.fp.groups[g].bytes = (uint32_t)P.groups2[g].blob.size();
            q.fp.groups[g].nb = (uint32_t)P.groups2[g].cols.size();
            q.fp.groups[g].base = g ? q.fp.groups[g - 1].base + q.fp.groups[g - 1].nb : 0;
            if (q.fp.groups[g].nb > (uint32_t)kMaxGroupBands) return c->fail(EGGPHOT_ERR_UNSUPPORTED, "more than 64 bands in one group");
            q.fp.max_blob = std::max(q.fp.max_blob, q.fp.groups[g].bytes);
        }
        q.nodes_mode = (uint32_t)P.nodes_mode;
        if (P.nodes_mode != EGGPHOT_NODES_MERGED) {
            if (!c->use_lane) return c->fail(EGGPHOT_ERR_UNSUPPORTED, "sed2flux_nodes=sed|filter needs the lane pipeline (unset EGGPHOT_KERNEL)");
            CK(c->altFl.upload(P.alt_fl.data(), P.alt_fl.size() * sizeof(double)));
            CK(c->altFr.upload(P.alt_fr.data(), P.alt_fr.size() * sizeof(double)));
            CK(c->altOff.upload(P.alt_off.data(), P.alt_off.size() * sizeof(uint32_t)));
            q.alt_fl = (const double *)c->altFl.p; q.alt_fr = (const double *)c->altFr.p; q.alt_off = (const uint32_t *)c->altOff.p;
        }
        q.gxD = (const double2 *)c->gridN.p;
        q.gxB = (const double2 *)c->nextOpt.p;
        q.item_mode = getenv("EGGPHOT_ITEM_MODE") ? (uint32_t)atoi(getenv("EGGPHOT_ITEM_MODE")) : 2u;
        q.gx_in_smem = lane_gx_in_smem(P.precision == EGGPHOT_FP64, P.grid_disk.n(), P.grid_bulge.n(), kSmemPerCta) ? 1 : 0;
        {
            std::vector<uint32_t> cols;
            for (const BandGroup &g : P.groups2) cols.insert(cols.end(), g.cols.begin(), g.cols.end());
            q.ntab = (uint32_t)cols.size();
            // the tables of every output column, in table order
            std::vector<uint32_t> lists(P.bands.size() + 1, 0);
            for (size_t col = 0; col < P.bands.size(); ++col) {
                for (size_t fb = 0; fb < cols.size(); ++fb)
                    if (cols[fb] == col) lists.push_back((uint32_t)fb);
                lists[col + 1] = (uint32_t)(lists.size() - P.bands.size() - 1);
            }
            cols = lists;
            CK(c->tabCol.upload(cols.data(), cols.size() * sizeof(uint32_t)));
            q.tab_col = (const uint32_t *)c->tabCol.p;
        }
        if (q.ntab > 64) return c->fail(EGGPHOT_ERR_UNSUPPORTED, "more than 64 band tables; raise smem_table_budget");
        if (P.grid_disk.n() > 32000) return c->fail(EGGPHOT_ERR_UNSUPPORTED, "SED wavelength grid with more than 32000 nodes");
        {
            std::vector<double> span(2 * std::max<size_t>(P.bands.size(), 1), 0.0);
            for (size_t b = 0; b < P.bands.size(); ++b) { span[2 * b] = P.bands[b].lam.front(); span[2 * b + 1] = P.bands[b].lam.back(); }
            CK(c->colSpan.upload(span.data(), span.size() * sizeof(double)));
            q.col_span = (const double2 *)c->colSpan.p;
        }
        {
            const size_t rs = P.precision == EGGPHOT_FP64 ? 8 : 4;
            q.rowsD = (uint32_t)lane_grid_rows(P.grid_disk.n());
            q.rowsB = (uint32_t)lane_grid_rows(P.grid_bulge.n());
            q.rows_stride = ((size_t)q.rowsD + q.rowsB) * kLanes * rs;
            for (int g = 0; g < kMaxGroups; ++g) q.group_cost[g] = 0.f;
            for (size_t g = 0; g < P.groups2.size(); ++g) {
                const BandHeader2 *hdr = reinterpret_cast<const BandHeader2 *>(P.groups2[g].blob.data());
                for (size_t k = 0; k < P.groups2[g].cols.size(); ++k) q.group_cost[g] += std::max(hdr[k].cost, 1.0f);
            }
        }
        if (c->use_lane) {
            const size_t smem = lane_kernel_smem(q, P.precision);
            if (smem + kLaneFixedSmem > (size_t)prop.sharedMemPerBlockOptin)
                return c->fail(EGGPHOT_ERR_UNSUPPORTED, "band tables need " + std::to_string(smem) + " B of shared memory, device offers " +
                                                         std::to_string(prop.sharedMemPerBlockOptin) + "; lower smem_table_budget");
            // SED rows of a chunk: 16 GiB per region (three regions at most; 180 GB of HBM) unless EGGPHOT_LANE_SCRATCH_MB says otherwise, 2048 batches at least
            size_t cap_bytes = size_t(16) << 30; // (every chunk pays the pipeline's fixed cost: 4 / 8 / 16 / 32 GiB give 1.303 / 1.323 / 1.331 / 1.340e9 on the 1.4e7-galaxy shard)
            if (const char *e = getenv("EGGPHOT_LANE_SCRATCH_MB")) { const long v = atol(e); if (v > 0) cap_bytes = (size_t)v << 20; }
            c->lane_cap_batches = std::max<size_t>(cap_bytes / q.rows_stride, 2048);
            for (auto &h : c->zhist) CK(h.alloc(zsort_hist_bytes()));
        }
    }
    CK(cudaDeviceSynchronize()); // every table upload has landed before a kernel on a non-blocking stream reads it
#undef CK
    c->ready = true;
    return EGGPHOT_OK;
}

void eggphot_destroy(eggphot_ctx *c) {
    if (!c) return;
    if (c->stream) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    for (cudaStream_t st : {c->s_in, c->s_out, c->s_c[0], c->s_c[1]}) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (auto &s : c->slots)
        for (cudaEvent_t ev : {s.ev_in, s.ev_k, s.ev_out}) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : c->sedDone) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : c->walkEv) cudaEventDestroy(ev);
    delete c;
}

const char *eggphot_last_error(const eggphot_ctx *c) { return c ? c->err.c_str() : "null context"; }

uint32_t eggphot_nband(const eggphot_ctx *c, int which) {
    if (!c) return 0;
    return (uint32_t)(which ? c->P.rfbands.size() : c->P.bands.size());
}

int eggphot_band_order(const eggphot_ctx *c, int which, uint32_t *order, float *lambda) {
    if (!c) return EGGPHOT_ERR_ARG;
    const auto &o = which ? c->P.rfband_order : c->P.band_order;
    const auto &l = which ? c->P.rfband_lambda : c->P.band_lambda;
    for (size_t k = 0; k < o.size(); ++k) {
        if (order) order[k] = o[k];
        if (lambda) lambda[k] = l[k];
    }
    return EGGPHOT_OK;
}

int eggphot_compute_device(eggphot_ctx *c, const eggphot_galaxies *gal, size_t ngal, const eggphot_outputs *out, void *stream) {
    if (!c || !c->ready) return c ? c->fail(EGGPHOT_ERR_ARG, "context is not ready") : EGGPHOT_ERR_ARG;
    if (!out) return c->fail(EGGPHOT_ERR_ARG, "outputs are required");
    int rc = check_galaxies(c, gal);
    if (rc) return rc;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return c->cuda_fail(e, "cudaSetDevice");
    c->launches = 0;
    ++c->walkCalls;
    return enqueue(c, *gal, ngal, *out, (cudaStream_t)stream, 0);
}

int eggphot_synchronize(eggphot_ctx *c) {
    if (!c || !c->ready) return EGGPHOT_ERR_ARG;
    cudaError_t e = cudaSetDevice(c->device);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return c->cuda_fail(e, "synchronize");
    return read_error_flag(c);
}

int eggphot_compute(eggphot_ctx *c, const eggphot_galaxies *gal, size_t ngal, const eggphot_outputs *out) {
    if (!c || !c->ready) return c ? c->fail(EGGPHOT_ERR_ARG, "context is not ready") : EGGPHOT_ERR_ARG;
    if (!out) return c->fail(EGGPHOT_ERR_ARG, "outputs are required");
    int rc = check_galaxies(c, gal);
    if (rc) return rc;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return c->cuda_fail(e, "cudaSetDevice");
    c->launches = 0;
    const Prepared &P = c->P;
    const size_t nb = P.bands.size(), nrf = P.rfbands.size(), nl = P.no_nebular ? 0 : P.line_lambda.size();
    // chunks: every chunk pays the pipeline's fixed cost once (about 0.37 ms of tails and small kernels),
    // and the first copy in and the last copy out are exposed.  Measured at 1.95e5 galaxies: three equal chunks 5.49 ms, four
    // 5.7 - 5.8, half chunks first and last around three or four whole ones 5.7 - 5.9, two 6.2, one 6.6.  So: a third of the
    // catalog, within [2^15, 2^20] galaxies.
    size_t chunk = (std::min<size_t>(std::max<size_t>((ngal + 2) / 3, 1u << 15), 1u << 20) + kLanes - 1) / kLanes * kLanes;
    if (const char *ev = getenv("EGGPHOT_CHUNK")) { // tuning knob: galaxies per chunk of the host-memory entry point
        const long long v = atoll(ev);
        if (v > 0) chunk = ((size_t)v + kLanes - 1) / kLanes * kLanes;
    }
    chunk = std::min(chunk, ngal);
    if (chunk == 0) return EGGPHOT_OK;
#define CK(x) do { if ((e = (x)) != cudaSuccess) return c->cuda_fail(e, #x); } while (0)
    using SlotBuf = DevBuf Slot::*;
    struct Col { SlotBuf buf; const void *src; size_t elem; };
    const Col ins[] = {
        {&Slot::z, gal->z, 4}, {&Slot::d, gal->d, 4}, {&Slot::m_disk, gal->m_disk, 4}, {&Slot::m_bulge, gal->m_bulge, 4},
        {&Slot::m2lcor, gal->m2lcor, 4}, {&Slot::lir, gal->lir, 4}, {&Slot::fpah, gal->fpah, 4}, {&Slot::mdust, gal->mdust, 4},
        {&Slot::igm_abs, P.apply_igm ? gal->igm_abs : nullptr, 12}, {&Slot::vdisp, nl ? gal->vdisp : nullptr, 4},
        {&Slot::lld, nl ? gal->line_lum_disk : nullptr, 4 * nl}, {&Slot::llb, nl ? gal->line_lum_bulge : nullptr, 4 * nl},
        {&Slot::osd, gal->opt_sed_disk, 8}, {&Slot::osb, gal->opt_sed_bulge, 8}, {&Slot::irs, gal->ir_sed, 8}};
    struct OCol { SlotBuf buf; void *dst; size_t elem; };
    const OCol outs[] = {
        {&Slot::flux, out->flux, 4 * nb}, {&Slot::flux_disk, out->flux_disk, 4 * nb}, {&Slot::flux_bulge, out->flux_bulge, 4 * nb},
        {&Slot::rfmag, out->rfmag, 4 * nrf}, {&Slot::rfmag_disk, out->rfmag_disk, 4 * nrf}, {&Slot::rfmag_bulge, out->rfmag_bulge, 4 * nrf},
        {&Slot::fd64, out->flux_disk_f64, 8 * nb}, {&Slot::fb64, out->flux_bulge_f64, 8 * nb},
        {&Slot::rd64, out->rfmag_disk_f64, 8 * nrf}, {&Slot::rb64, out->rfmag_bulge_f64, 8 * nrf}};
    for (auto &s : c->slots) {
        for (const Col &k : ins) if (k.src && k.elem) CK((s.*(k.buf)).alloc(chunk * k.elem));
        for (const OCol &k : outs) if (k.dst && k.elem) CK((s.*(k.buf)).alloc(chunk * k.elem));
    }
    // Three stages on their own streams: copies in (s_in), kernels (s_c, consecutive chunks alternating between two streams so
    // that the tail of one chunk's kernels overlaps the head of the next's), copies out (s_out).  Chunk k stages through slot
    // k mod 4: its inputs may be overwritten once the kernels of chunk k - 4 are done, its outputs once they are back on the
    // host.  PCIe runs in both directions under the kernels; only the first copy in and the last copy out are exposed.
    // EGGPHOT_TRACE=1: device-side timeline of the chunks (events on the chunk streams), printed to stderr after the call
    const bool trace = getenv("EGGPHOT_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        tev.push_back(ev);
    };
    std::vector<size_t> sizes;
    for (size_t left = ngal; left > 0;) {
        const size_t n = std::min(chunk, left);
        sizes.push_back(n);
        left -= n;
    }
    // kernels of consecutive chunks alternate between two streams (EGGPHOT_CSTREAMS=1: one stream, measured 6.2 ms against 5.75)
    const int ncs = getenv("EGGPHOT_CSTREAMS") && atoi(getenv("EGGPHOT_CSTREAMS")) == 1 ? 1 : 2;
    c->walk_split = ncs == 2 && sizes.size() > 1 && getenv("EGGPHOT_WALK_SPLIT") && atoi(getenv("EGGPHOT_WALK_SPLIT")) == 1;
    size_t first = 0;
    for (size_t k = 0; k < sizes.size(); first += sizes[k], ++k) {
        Slot &s = c->slots[k % kSlots];
        cudaStream_t sc = c->s_c[k % ncs];
        const size_t n = sizes[k];
        if (k >= (size_t)kSlots) CK(cudaStreamWaitEvent(c->s_in, s.ev_k, 0));
        mark(c->s_in);
        for (const Col &col : ins)
            if (col.src && col.elem)
                CK(cudaMemcpyAsync((s.*(col.buf)).p, (const char *)col.src + first * col.elem, n * col.elem,
                                   cudaMemcpyHostToDevice, c->s_in));
        eggphot_galaxies dg{};
        dg.z = (const float *)s.z.p; dg.d = (const float *)s.d.p; dg.m_disk = (const float *)s.m_disk.p;
        dg.m_bulge = (const float *)s.m_bulge.p; dg.m2lcor = (const float *)s.m2lcor.p; dg.lir = (const float *)s.lir.p;
        dg.fpah = gal->fpah ? (const float *)s.fpah.p : nullptr; dg.mdust = gal->mdust ? (const float *)s.mdust.p : nullptr;
        dg.igm_abs = (const float *)s.igm_abs.p; dg.vdisp = (const float *)s.vdisp.p;
        dg.line_lum_disk = (nl && gal->line_lum_disk) ? (const float *)s.lld.p : nullptr;
        dg.line_lum_bulge = (nl && gal->line_lum_bulge) ? (const float *)s.llb.p : nullptr;
        dg.opt_sed_disk = (const uint64_t *)s.osd.p; dg.opt_sed_bulge = (const uint64_t *)s.osb.p; dg.ir_sed = (const uint64_t *)s.irs.p;
        eggphot_outputs dout{};
        dout.flux = out->flux ? (float *)s.flux.p : nullptr;
        dout.flux_disk = out->flux_disk ? (float *)s.flux_disk.p : nullptr;
        dout.flux_bulge = out->flux_bulge ? (float *)s.flux_bulge.p : nullptr;
        dout.rfmag = out->rfmag ? (float *)s.rfmag.p : nullptr;
        dout.rfmag_disk = out->rfmag_disk ? (float *)s.rfmag_disk.p : nullptr;
        dout.rfmag_bulge = out->rfmag_bulge ? (float *)s.rfmag_bulge.p : nullptr;
        dout.flux_disk_f64 = out->flux_disk_f64 ? (double *)s.fd64.p : nullptr;
        dout.flux_bulge_f64 = out->flux_bulge_f64 ? (double *)s.fb64.p : nullptr;
        dout.rfmag_disk_f64 = out->rfmag_disk_f64 ? (double *)s.rd64.p : nullptr;
        dout.rfmag_bulge_f64 = out->rfmag_bulge_f64 ? (double *)s.rb64.p : nullptr;
        mark(c->s_in);
        CK(cudaEventRecord(s.ev_in, c->s_in));
        CK(cudaStreamWaitEvent(sc, s.ev_in, 0));
        if (k >= (size_t)kSlots) CK(cudaStreamWaitEvent(sc, s.ev_out, 0));
        rc = enqueue(c, dg, n, dout, sc, 1 + (int)(k % ncs));
        if (rc) return rc;
        mark(sc);
        CK(cudaEventRecord(s.ev_k, sc));
        CK(cudaStreamWaitEvent(c->s_out, s.ev_k, 0));
        for (const OCol &col : outs)
            if (col.dst && col.elem)
                CK(cudaMemcpyAsync((char *)col.dst + first * col.elem, (s.*(col.buf)).p, n * col.elem,
                                   cudaMemcpyDeviceToHost, c->s_out));
        mark(c->s_out);
        CK(cudaEventRecord(s.ev_out, c->s_out));
    }
    CK(cudaStreamSynchronize(c->s_out));
    for (cudaStream_t st : {c->s_in, c->s_c[0], c->s_c[1]}) CK(cudaStreamSynchronize(st));
    if (trace) {
        for (size_t i = 0; i + 3 < tev.size(); i += 4) {
            float t[4];
            for (int j = 0; j < 4; ++j) cudaEventElapsedTime(&t[j], tev[0], tev[i + j]);
            fprintf(stderr, "eggphot trace: chunk %zu  h2d %.3f-%.3f  kernels -%.3f  d2h -%.3f ms\n", i / 4, t[0], t[1], t[2], t[3]);
        }
        for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
    }
#undef CK
    return read_error_flag(c);
}

uint32_t eggphot_sed_size(const eggphot_ctx *c, int component) {
    if (!c) return 0;
    return (uint32_t)(component ? c->P.grid_disk.n() : c->P.grid_bulge.n());
}

// save_sed of galaxies [first, first + count) of one component, enqueued on the context's stream: columns in, sed_kernel,
// spectra out to lam / sed with row strides of `stride` floats.  Staging slot `slot` (0 / 1) must be free.
static int enqueue_seds(eggphot_ctx *c, const eggphot_galaxies *gal, size_t first, size_t count, int component, int slot,
                        float *lam, float *sed, bool packed) {
    cudaError_t e;
    cudaStream_t st = c->s_c[slot]; // a stream per staging slot: the copy out of one range overlaps the kernel of the next
#define CK(x) do { if ((e = (x)) != cudaSuccess) return c->cuda_fail(e, #x); } while (0)
    const Prepared &P = c->P;
    const size_t nl = P.no_nebular ? 0 : P.line_lambda.size();
    const size_t n = eggphot_sed_size(c, component);
    DevBuf *cols = c->sedCols[slot];
    const void *src[15] = {gal->z, gal->d, gal->m_disk, gal->m_bulge, gal->m2lcor, gal->lir, gal->fpah, gal->mdust,
                           P.apply_igm ? gal->igm_abs : nullptr, nl ? gal->vdisp : nullptr, nl ? gal->line_lum_disk : nullptr,
                           nl ? gal->line_lum_bulge : nullptr, gal->opt_sed_disk, gal->opt_sed_bulge, gal->ir_sed};
    const size_t elem[15] = {4, 4, 4, 4, 4, 4, 4, 4, 12, 4, 4 * nl, 4 * nl, 8, 8, 8};
    for (int k = 0; k < 15; ++k)
        if (src[k] && elem[k]) CK(cols[k].upload_async((const char *)src[k] + first * elem[k], count * elem[k], st));
    CK(c->sedOut[slot].alloc(count * 2 * n * sizeof(float)));
    eggphot_galaxies dg{};
    dg.z = (const float *)cols[0].p; dg.d = (const float *)cols[1].p; dg.m_disk = (const float *)cols[2].p;
    dg.m_bulge = (const float *)cols[3].p; dg.m2lcor = (const float *)cols[4].p; dg.lir = (const float *)cols[5].p;
    dg.fpah = src[6] ? (const float *)cols[6].p : nullptr; dg.mdust = src[7] ? (const float *)cols[7].p : nullptr;
    dg.igm_abs = (const float *)cols[8].p;
    dg.vdisp = (const float *)cols[9].p; dg.line_lum_disk = src[10] ? (const float *)cols[10].p : nullptr;
    dg.line_lum_bulge = src[11] ? (const float *)cols[11].p : nullptr;
    dg.opt_sed_disk = (const uint64_t *)cols[12].p; dg.opt_sed_bulge = (const uint64_t *)cols[13].p; dg.ir_sed = (const uint64_t *)cols[14].p;
    SedParams q{};
    eggphot_outputs none{};
    fill_params(c, dg, count, none, q.fp);
    q.first = 0; q.count = count; q.component = component;
    float *dev = (float *)c->sedOut[slot].p;
    if (packed) { q.lam = dev; q.sed = dev + n; q.row_stride = 2 * n; }               // [count][lam | sed]: the file's layout
    else { q.lam = dev; q.sed = dev + count * n; q.row_stride = n; }                 // two [count][n] arrays
    int le = launch_sed_kernel(q, P.precision, st);
    if (le) return c->cuda_fail((cudaError_t)le, "sed kernel launch");
    if (packed) CK(cudaMemcpyAsync(lam, dev, count * 2 * n * sizeof(float), cudaMemcpyDeviceToHost, st));
    else {
        CK(cudaMemcpyAsync(lam, dev, count * n * sizeof(float), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(sed, dev + count * n, count * n * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (!c->sedDone[slot]) CK(cudaEventCreateWithFlags(&c->sedDone[slot], cudaEventDisableTiming));
    CK(cudaEventRecord(c->sedDone[slot], st));
#undef CK
    return EGGPHOT_OK;
}

int eggphot_get_seds_packed_wait(eggphot_ctx *c, int which);
static int seds_args(eggphot_ctx *c, const eggphot_galaxies *gal, int component, const void *a, const void *b) {
    if (!c || !c->ready) return c ? c->fail(EGGPHOT_ERR_ARG, "context is not ready") : EGGPHOT_ERR_ARG;
    if (!a || !b) return c->fail(EGGPHOT_ERR_ARG, "output arrays are required");
    if (component == 0 && !c->P.has_bulge) return c->fail(EGGPHOT_ERR_ARG, "no bulge component (no_stellar)");
    int rc = check_galaxies(c, gal);
    if (rc) return rc;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return c->cuda_fail(e, "cudaSetDevice");
    return EGGPHOT_OK;
}

int eggphot_get_seds(eggphot_ctx *c, const eggphot_galaxies *gal, size_t first, size_t count, int component,
                     float *lam, float *sed) {
    int rc = seds_args(c, gal, component, lam, sed);
    if (rc || count == 0) return rc;
    rc = eggphot_get_seds_packed_wait(c, 1); // (ranges begun with eggphot_get_seds_packed_begin share the staging slots)
    if (rc) return rc;
    rc = enqueue_seds(c, gal, first, count, component, 0, lam, sed, false);
    if (rc) return rc;
    return eggphot_get_seds_packed_wait(c, 1);
}

int eggphot_get_seds_packed_begin(eggphot_ctx *c, const eggphot_galaxies *gal, size_t first, size_t count, int component,
                                  float *out) {
    int rc = seds_args(c, gal, component, out, out);
    if (rc || count == 0) return rc;
    const int slot = c->sedSlot;
    c->sedSlot ^= 1;
    return enqueue_seds(c, gal, first, count, component, slot, out, nullptr, true);
}

int eggphot_get_seds_packed_wait(eggphot_ctx *c, int which) {
    if (!c || !c->ready) return EGGPHOT_ERR_ARG;
    // `which` = 0: the oldest range still in flight (the one begun before the last), 1: everything
    cudaError_t e = cudaSetDevice(c->device);
    if (e == cudaSuccess) {
        if (which == 0) { if (c->sedDone[c->sedSlot]) e = cudaEventSynchronize(c->sedDone[c->sedSlot]); }
        else { e = cudaStreamSynchronize(c->s_c[0]); if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_c[1]); }
    }
    if (e != cudaSuccess) return c->cuda_fail(e, "waiting for the spectra");
    return which == 0 ? EGGPHOT_OK : read_error_flag(c);
}

int eggphot_host_register(void *ptr, size_t bytes) {
    if (!ptr || !bytes) return EGGPHOT_OK;
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return EGGPHOT_ERR_CUDA; }
    return EGGPHOT_OK;
}
int eggphot_host_unregister(void *ptr) {
    if (!ptr) return EGGPHOT_OK;
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return EGGPHOT_ERR_CUDA; }
    return EGGPHOT_OK;
}

int eggphot_get_grid(const eggphot_ctx *c, int component, double *x) {
    if (!c || !x) return EGGPHOT_ERR_ARG;
    const Grid &g = component ? c->P.grid_disk : c->P.grid_bulge;
    std::copy(g.x.begin(), g.x.end(), x);
    return EGGPHOT_OK;
}

int eggphot_measure_fma_peak(eggphot_ctx *c, double *fp32_tflops, double *fp64_tflops) {
    if (!c || !c->ready || !fp32_tflops || !fp64_tflops) return c ? c->fail(EGGPHOT_ERR_ARG, "context is not ready") : EGGPHOT_ERR_ARG;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return c->cuda_fail(e, "cudaSetDevice");
    const int rc = measure_fma_peak(c->num_sms, c->errflag.p ? (char *)c->zhist[0].p : nullptr, c->stream, fp32_tflops, fp64_tflops);
    if (rc) return c->cuda_fail((cudaError_t)rc, "FMA micro-benchmark");
    return EGGPHOT_OK;
}

int eggphot_maglim(eggphot_ctx *c, eggprops_ctx *props, const double *zb_lo, const double *zb_hi, uint32_t nz, double maglim,
                   double mmin_strict, double mmax, uint32_t ntrial, float *z_mmin) {
    if (!c || !c->ready) return c ? c->fail(EGGPHOT_ERR_ARG, "context is not ready") : EGGPHOT_ERR_ARG;
    if (!props || !zb_lo || !zb_hi || !z_mmin || ntrial < 2) return c->fail(EGGPHOT_ERR_ARG, "maglim: null argument or fewer than 2 trial masses");
    const Prepared &P = c->P;
    if (P.bands.size() != 1 || P.disk_mode != DISK_OPT || !P.no_nebular || P.apply_igm || !P.has_bulge)
        return c->fail(EGGPHOT_ERR_ARG, "maglim: the context must hold the selection band alone, with no_dust, no_nebular and igm=none "
                                        "(the estimator sees stellar light only, src/egg-gencat.cpp:943-952)");
    if (nz == 0) return EGGPHOT_OK;
    const size_t n = (size_t)nz * ntrial;
    std::vector<float> z(n), m(n), d(n), m2l(n), mb(n, -60.0f); // (a bulge of 10^-60 Msun: its flux underflows to 0)
    std::vector<uint8_t> passive(n, 0);
    std::vector<uint64_t> sed(n);
    for (uint32_t k = 0; k < nz; ++k)
        for (uint32_t i = 0; i < ntrial; ++i) {
            z[(size_t)k * ntrial + i] = (float)(0.5 * (zb_lo[k] + zb_hi[k]));                                  // bin_center, :928
            m[(size_t)k * ntrial + i] = (float)(mmin_strict + (mmax - mmin_strict) * (double)i / (ntrial - 1)); // rgen, :931
        }
    eggprops_out E{};
    E.d = d.data(); E.m2lcor = m2l.data(); E.opt_sed_disk = sed.data();
    int rc = eggprops_compute(props, z.data(), m.data(), passive.data(), n, 0, &E);
    if (rc) return c->fail(rc, std::string("maglim: ") + eggprops_last_error(props));
    eggphot_galaxies g{};
    g.z = z.data(); g.d = d.data(); g.m2lcor = m2l.data();
    g.m_disk = m.data(); g.opt_sed_disk = sed.data();   // the whole trial mass in the template of the trial colours (:945)
    g.m_bulge = mb.data(); g.opt_sed_bulge = sed.data();
    std::vector<double> flux(n);
    eggphot_outputs o{};
    o.flux_disk_f64 = flux.data();
    rc = eggphot_compute(c, &g, n, &o);
    if (rc) return rc;
    const double flim = exp10(0.4 * (23.9 - maglim)); // mag2uJy
    for (uint32_t k = 0; k < nz; ++k) {
        z_mmin[k] = (float)mmax;
        for (uint32_t i = 0; i < ntrial; ++i)
            if (flux[(size_t)k * ntrial + i] - flim > 0) { // where_first, :955
                z_mmin[k] = (float)std::max(mmin_strict, (double)m[(size_t)k * ntrial + i] - 0.2);
                break;
            }
    }
    return EGGPHOT_OK;
}

int eggphot_get_stats(const eggphot_ctx *cc, eggphot_stats *s) {
    if (!cc || !s) return EGGPHOT_ERR_ARG;
    eggphot_ctx *c = const_cast<eggphot_ctx *>(cc); // (the event pairs are consumed)
    std::memset(s, 0, sizeof(*s));
    {   // mean device time of the walk kernel per eggphot_compute_device call since the last query
        double ms = 0.0;
        for (size_t k = 0; k + 1 < c->walkUsed; k += 2) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, c->walkEv[k], c->walkEv[k + 1]) == cudaSuccess) ms += t; else cudaGetLastError();
        }
        s->last_kernel_ms = c->walkCalls ? ms / (double)c->walkCalls : 0.0;
        c->walkUsed = 0; c->walkCalls = 0;
    }
    s->kernel_launches = c->launches;
    s->ngroups = (uint32_t)(c->use_lane ? c->P.groups2.size() : c->P.groups.size());
    s->nodes_disk = (uint32_t)c->P.grid_disk.n();
    s->nodes_bulge = (uint32_t)c->P.grid_bulge.n();
    s->nodes_ir = c->P.nir;
    s->table_bytes = c->table_bytes;
    return EGGPHOT_OK;
}

} // extern "C"
