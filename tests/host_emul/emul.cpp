// emul.cpp — TEST-ONLY sequential walk of the device algorithm on the host.
//
// Not part of the product and never loaded by egg_b200/: it exists so that the table builder
// (phot_tables.cpp) and the curvature-form arithmetic (phot_core.cuh) — the exact code the CUDA
// kernels compile — can be checked against the oracle in the CPU-only test tier, in both
// arithmetic modes, before any GPU time is spent.  The warp orchestration of phot_kernels.cu
// (passes, reductions) is replaced by plain loops; everything numeric is shared.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../egg_b200/csrc/phot_tables.hpp"

using namespace eggphot;

namespace {

int last_le(const std::vector<double> &x, double v) {
    return (int)(std::upper_bound(x.begin(), x.end(), v) - x.begin()) - 1;
}

void add_lines(const Prepared &P, const Grid &g, std::vector<double> &S, const float *ll, double vdisp, double inv) {
    const int n = g.n(), nl = (int)P.line_lambda.size();
    std::vector<int> b0(nl, 1), b1(nl, 0);
    for (int l = 0; l < nl; ++l) {
        if (ll[l] == 0.0f) continue;
        const double l0 = P.line_lambda[l], lw = l0 * (vdisp / 2.998e5);
        int lo = last_le(g.x, l0 - 10 * lw), hi = last_le(g.x, l0 + 10 * lw) + 1;
        if (hi == 0 || lo == n - 1) continue;
        b0[l] = lo < 0 ? 0 : lo;
        b1[l] = hi >= n ? n - 1 : hi;
    }
    for (int l = 0; l < nl; ++l)
        for (int j = b0[l]; j <= b1[l]; ++j) {
            bool owner = true;
            for (int k = 0; k < l; ++k) owner = owner && !(b0[k] <= j && j <= b1[k]);
            if (!owner) continue;
            const double xj = g.x[j], xm = j > 0 ? g.x[j - 1] : 0.0, xp = j < n - 1 ? g.x[j + 1] : 0.0;
            const double llow = j != 0 ? 0.5 * (xm + xj) : xj - 0.5 * (xp - xj);
            const double lup = j != n - 1 ? 0.5 * (xj + xp) : xj + 0.5 * (xj - xm);
            double acc = 0;
            for (int k = l; k < nl; ++k)
                if (b0[k] <= j && j <= b1[k]) {
                    const double l0 = P.line_lambda[k], lw = l0 * (vdisp / 2.998e5);
                    acc += line_bin(llow, lup, l0, lw, (double)ll[k] * l0, P.line_profile == EGGPHOT_LINE_AVERAGE);
                }
            S[j] += xj * acc * inv;
        }
}

template <typename real>
void run(const Prepared &P, const eggphot_galaxies &G, size_t ngal, double *fd, double *fb, double *rd, double *rb) {
    const size_t nb = P.bands.size(), nrf = P.rfbands.size();
    const int nl = P.no_nebular ? 0 : (int)P.line_lambda.size();
    const int nD = P.grid_disk.n(), nB = P.grid_bulge.n();
    std::vector<double> SDd(nD), SBd(nB);
    std::vector<real> SD(nD), SB(nB);
    for (size_t g = 0; g < ngal; ++g) {
        const double z = G.z[g], d = G.d[g], opz = 1.0 + z;
        const double aflux = EGGPHOT_LSUN2UJY * (1.0 + z) / (d * d);
        const double a_b = P.has_bulge ? std::pow(10.0, (double)(G.m_bulge[g] - G.m2lcor[g])) : 0.0;
        const double a_d = P.disk_mode != DISK_IR ? std::pow(10.0, (double)(G.m_disk[g] - G.m2lcor[g])) : 0.0;
        double c1 = 0, c2 = 0;
        if (P.disk_mode != DISK_OPT) {
            if (P.ir_kind == EGGPHOT_IR_CS17) { c1 = (1.0 - (double)G.fpah[g]) * G.mdust[g]; c2 = (double)G.fpah[g] * G.mdust[g]; }
            else c1 = G.lir[g];
        }
        float ig[3] = {1, 1, 1};
        if (P.apply_igm) for (int k = 0; k < 3; ++k) ig[k] = G.igm_abs[3 * g + k];
        // SEDs in the kernel's real type, built with the kernel's operation order
        const real ad = (real)a_d, rc1 = (real)c1, rc2 = (real)c2;
        for (int j = 0; j < nD; ++j) {
            real v = 0;
            const size_t ro = P.disk_mode != DISK_IR ? (size_t)G.opt_sed_disk[g] * nD + j : 0;
            const size_t ri = P.disk_mode != DISK_OPT ? (size_t)G.ir_sed[g] * nD + j : 0;
            if (P.disk_mode == DISK_BOTH) v = ad * (real)P.rows_disk_opt[ro] + rc1 * (real)P.rows_disk_ir[ri];
            else if (P.disk_mode == DISK_OPT) v = ad * (real)P.rows_disk_opt[ro];
            else v = rc1 * (real)P.rows_disk_ir[ri];
            if (P.disk_mode != DISK_OPT && P.ir_kind == EGGPHOT_IR_CS17) v += rc2 * (real)P.rows_disk_pah[ri];
            if (P.apply_igm)
                v *= igm_factor<real>(j, P.grid_disk.lz, P.grid_disk.l0, P.grid_disk.l1, P.grid_disk.l2, ig[0], ig[1], ig[2]);
            SD[j] = v;
        }
        for (int j = 0; j < nB; ++j) {
            real v = (real)P.rows_bulge_opt[(size_t)G.opt_sed_bulge[g] * nB + j];
            if (P.apply_igm)
                v *= igm_factor<real>(j, P.grid_bulge.lz, P.grid_bulge.l0, P.grid_bulge.l1, P.grid_bulge.l2, ig[0], ig[1], ig[2]);
            SB[j] = v;
        }
        if (nl) {
            for (int j = 0; j < nD; ++j) SDd[j] = 0;
            add_lines(P, P.grid_disk, SDd, G.line_lum_disk + g * nl, G.vdisp[g], 1.0);
            for (int j = 0; j < nD; ++j) if (SDd[j] != 0) SD[j] += (real)SDd[j];
            if (P.has_bulge && G.line_lum_bulge) {
                for (int j = 0; j < nB; ++j) SBd[j] = 0;
                add_lines(P, P.grid_bulge, SBd, G.line_lum_bulge + g * nl, G.vdisp[g], a_b > 0 ? 1.0 / a_b : 0.0);
                for (int j = 0; j < nB; ++j) if (SBd[j] != 0) SB[j] += (real)SBd[j];
            }
        }
        // the bulge SED resampled on the disk grid (kernel phase A4): same piece-wise linear function
        std::vector<real> SBU(nD, real(0));
        if (P.has_bulge)
            for (int j = 0; j < nD; ++j) {
                const int o = P.u_oleft[j];
                if (o < 0) continue;
                SBU[j] = P.u_isopt[j] ? SB[o] : SB[o] + (real)P.u_ofrac[j] * (SB[o + 1] - SB[o]);
            }
        for (const BandGroup &grp : P.groups) {
            const uint8_t *blob = grp.blob.data();
            const BandHeader *hdr = reinterpret_cast<const BandHeader *>(blob);
            for (size_t bi = 0; bi < grp.cols.size(); ++bi) {
                const BandHeader &h = hdr[bi];
                const Grid &gr = P.grid_disk;
                const int n = gr.n();
                double fdv = 0, fbv = 0;
                const bool covered = n >= 2 && gr.x[0] * opz <= h.full_first && gr.x[n - 1] * opz >= h.full_last;
                const bool bulge_on = P.has_bulge && nB >= 2 && P.grid_bulge.x[0] * opz <= h.full_first &&
                                      P.grid_bulge.x[nB - 1] * opz >= h.full_last;
                if (covered && h.n >= 2) {
                    int jlo = std::max(last_le(gr.x, h.eff_first / opz * (1.0 - 1e-15)), 0);
                    int jhi = (int)(std::lower_bound(gr.x.begin(), gr.x.end(), h.eff_last / opz * (1.0 + 1e-15)) - gr.x.begin());
                    jhi = std::min(jhi, n - 1);
                    if (jhi > jlo) {
                        BandView<real> bv;
                        make_band_view<real>(blob, h, bv);
                        const real opzr = (real)opz;
                        // b = last node of [jlo, jhi) left of the band's split node
                        int b = jlo;
                        while (b + 1 < jhi && std::fma(opz, gr.x[b + 1], -bv.fc) < bv.f_split) ++b;
                        const double asym_b1 = std::fma(bv.tot0, std::fma(opz, gr.x[b + 1], -bv.fc), bv.asym0), tot0z = bv.tot0 * opz;
                        // per node: Lambda (0 at both ends of the window) and the gamma terms of both components
                        const int N = jhi - jlo + 1;
                        std::vector<double> lam(N, 0.0);
                        std::vector<real> gD(N, real(0)), gB(N, real(0));
                        for (int j = jlo + 1; j < jhi; ++j) {
                            const double t = std::fma(opz, gr.x[j], -bv.fc);
                            const int c = find_cell<real>(bv, t);
                            double u;
                            lam[j - jlo] = lambda_in_cell<real>(bv, c, t, u);
                            const real ur = (real)u, sr = bv.sdl[cell_slot(c)].x, dlr = bv.sdl[cell_slot(c)].y;
                            const real dxr = (real)(gr.x[j + 1] - gr.x[j]);
                            gD[j - jlo] = gamma_term<real>(sr, dlr, ur, dxr * opzr) * (real)gr.idx[j];
                            if (bulge_on) gB[j - jlo] = gamma_term<real>(sr, dlr, ur, (real)P.u_dxo[j] * opzr) * (real)gr.idx[j];
                        }
                        // E_j for jlo <= j < jhi (0 outside).  fp32 mode: W_j = E_j - E_{j-1} (+ the float gamma terms), one
                        // partial sum per lane in float (node j sits on lane (j - jlo) % 30 + 1 of its pass), lanes added in
                        // double.  fp64 mode: summation by parts, (E_j + g_j / 2) (S_j - S_{j+1}) per interval (lane
                        // (j - jlo) % 31 of its pass) plus S_{b+1} tot0 (1+z) in the interval that holds the split node.
                        real accd[32] = {0}, accb[32] = {0};
                        if (sizeof(real) == 8) {
                            for (int j = jlo; j < jhi; ++j) {
                                const int k = j - jlo, ln = k % 31;
                                double dlam = lam[k + 1] - lam[k];
                                if (j == b) dlam += asym_b1;
                                const double E = dlam * gr.idx[j];
                                accd[ln] = std::fma(E + 0.5 * gD[k], (real)SD[j] - (real)SD[j + 1], accd[ln]);
                                if (j == b) accd[ln] = std::fma((real)SD[j + 1], (real)tot0z, accd[ln]);
                                if (bulge_on) {
                                    accb[ln] = std::fma(E + 0.5 * gB[k], (real)SBU[j] - (real)SBU[j + 1], accb[ln]);
                                    if (j == b) accb[ln] = std::fma((real)SBU[j + 1], (real)tot0z, accb[ln]);
                                }
                            }
                        } else {
                            double Ep = 0;
                            for (int j = jlo; j <= jhi; ++j) {
                                const int k = j - jlo;
                                double E = 0;
                                if (j < jhi) {
                                    double dlam = lam[k + 1] - lam[k];
                                    if (j == b) dlam += asym_b1;
                                    E = dlam * gr.idx[j];
                                }
                                double Eprev = Ep;
                                if (j == b + 1) Eprev -= tot0z;
                                const real W = (real)(E - Eprev);
                                const int ln = k % 30 + 1;
                                accd[ln] = std::fma((real)SD[j], W, accd[ln]);
                                if (bulge_on) accb[ln] = std::fma((real)SBU[j], W, accb[ln]);
                                // the gamma terms by interval (lane ln of the same pass): g_j / 2 (S_j - S_{j+1}), the
                                // difference of neighbouring SED values being exact in float
                                if (j < jhi) {
                                    const int li = ln == 30 ? 0 : ln; // lane 30's interval belongs to the next pass (its lane 0)
                                    accd[li] = std::fma(real(0.5) * gD[k], (real)SD[j] - (real)SD[j + 1], accd[li]);
                                    if (bulge_on) accb[li] = std::fma(real(0.5) * gB[k], (real)SBU[j] - (real)SBU[j + 1], accb[li]);
                                }
                                Ep = E;
                            }
                        }
                        const double iopz = 1.0 / opz;
                        double sd = 0, sb = 0;
                        for (int l = 0; l < 32; ++l) { sd += (double)accd[l]; sb += (double)accb[l]; }
                        fdv = sd * iopz * aflux;
                        if (bulge_on) fbv = sb * iopz * aflux * a_b;
                    }
                }
                if (!std::isfinite(fdv)) fdv = 0;
                if (!std::isfinite(fbv)) fbv = 0;
                fd[g * nb + h.out_col] = fdv;
                if (P.has_bulge) fb[g * nb + h.out_col] = fbv;
            }
        }
        const double arf = EGGPHOT_LSUN2UJY / (1e-5 * 1e-5);
        for (size_t r = 0; r < nrf; ++r)
            for (int comp = P.has_bulge ? 0 : 1; comp < 2; ++comp) {
                const RfWeights &w = comp ? P.rf_disk[r] : P.rf_bulge[r];
                const std::vector<real> &S = comp ? SD : SB;
                CompSum<real> acc;
                acc.clear();
                for (size_t k = 0; k < w.w.size(); ++k) acc.add(S[w.j0 + k] * (real)w.w[k]);
                double fl = w.covered ? ((double)acc.s + (double)acc.c) * arf * (comp ? 1.0 : a_b) : NAN;
                double m = -2.5 * std::log10(fl) + 23.9;
                (comp ? rd : rb)[g * nrf + r] = std::isfinite(m) ? m : 99.0;
            }
    }
}

// ---- the lane kernel's algorithm (phot_lane.cuh): batches of 32 z-sorted galaxies, union windows per band,
//      blocks of 64 nodes walked per galaxy, block sums added in double in block order
// emission lines as the lane kernel deposits them: line by line in wavelength order, each bin integral rounded to the
// kernel's real type and added to the node
template <typename real>
void add_lines_lane(const Prepared &P, const Grid &g, std::vector<real> &S, const float *ll, double vdisp, double inv) {
    const int n = g.n(), nl = (int)P.line_lambda.size();
    std::vector<int> ord(nl);
    for (int l = 0; l < nl; ++l) ord[l] = l;
    std::stable_sort(ord.begin(), ord.end(), [&](int i, int j) { return P.line_lambda[i] < P.line_lambda[j]; });
    for (int q = 0; q < nl; ++q) {
        const int l = ord[q];
        if (ll[l] == 0.0f) continue;
        const double l0 = P.line_lambda[l], lw = l0 * (vdisp / 2.998e5);
        int lo = last_le(g.x, l0 - 10 * lw), hi = last_le(g.x, l0 + 10 * lw) + 1;
        if (hi == 0 || lo == n - 1) continue;
        const int b0 = lo < 0 ? 0 : lo, b1 = hi >= n ? n - 1 : hi;
        for (int j = b0; j <= b1; ++j) {
            const double xj = g.x[j], xm = j > 0 ? g.x[j - 1] : 0.0, xp = j < n - 1 ? g.x[j + 1] : 0.0;
            const double llow = j != 0 ? 0.5 * (xm + xj) : xj - 0.5 * (xp - xj);
            const double lup = j != n - 1 ? 0.5 * (xj + xp) : xj + 0.5 * (xj - xm);
            S[j] += (real)(xj * line_bin_t<real>(llow, lup, l0, lw, (double)ll[l] * l0, P.line_profile == EGGPHOT_LINE_AVERAGE) * inv);
        }
    }
}

template <typename real>
void run2(const Prepared &P, const eggphot_galaxies &G, size_t ngal, double *fd, double *fb, double *rd, double *rb) {
    const size_t nb = P.bands.size(), nrf = P.rfbands.size();
    const int nl = P.no_nebular ? 0 : (int)P.line_lambda.size();
    const int nD = P.grid_disk.n(), nB = P.grid_bulge.n();
    std::vector<uint32_t> perm(ngal);
    for (size_t g = 0; g < ngal; ++g) perm[g] = (uint32_t)g;
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return G.z[a] < G.z[b]; });
    std::vector<double> SDd(nD), SBd(nB);
    std::vector<std::vector<real>> SD(kLanes, std::vector<real>(nD + 2, real(0))), SB(kLanes, std::vector<real>(nB + 2, real(0)));
    for (size_t g0 = 0; g0 < ngal; g0 += kLanes) {
        const int ng = (int)std::min<size_t>(kLanes, ngal - g0);
        double opz[kLanes], aflux[kLanes], a_b[kLanes];
        double omin = 1e300, omax = 0;
        for (int l = 0; l < ng; ++l) {
            const size_t g = perm[g0 + l];
            const double z = G.z[g], d = G.d[g];
            opz[l] = 1.0 + z;
            omin = std::min(omin, opz[l]); omax = std::max(omax, opz[l]);
            aflux[l] = EGGPHOT_LSUN2UJY * (1.0 + z) / (d * d);
            a_b[l] = P.has_bulge ? std::pow(10.0, (double)(G.m_bulge[g] - G.m2lcor[g])) : 0.0;
            const double a_d = P.disk_mode != DISK_IR ? std::pow(10.0, (double)(G.m_disk[g] - G.m2lcor[g])) : 0.0;
            double c1 = 0, c2 = 0;
            if (P.disk_mode != DISK_OPT) {
                if (P.ir_kind == EGGPHOT_IR_CS17) { c1 = (1.0 - (double)G.fpah[g]) * G.mdust[g]; c2 = (double)G.fpah[g] * G.mdust[g]; }
                else c1 = G.lir[g];
            }
            float ig[3] = {1, 1, 1};
            if (P.apply_igm) for (int k = 0; k < 3; ++k) ig[k] = G.igm_abs[3 * g + k];
            const real ad = (real)a_d, rc1 = (real)c1, rc2 = (real)c2;
            for (int j = 0; j < nD; ++j) {
                real v = 0;
                const size_t ro = P.disk_mode != DISK_IR ? (size_t)G.opt_sed_disk[g] * nD + j : 0;
                const size_t ri = P.disk_mode != DISK_OPT ? (size_t)G.ir_sed[g] * nD + j : 0;
                if (P.disk_mode == DISK_BOTH) v = ad * (real)P.rows_disk_opt[ro] + rc1 * (real)P.rows_disk_ir[ri];
                else if (P.disk_mode == DISK_OPT) v = ad * (real)P.rows_disk_opt[ro];
                else v = rc1 * (real)P.rows_disk_ir[ri];
                if (P.disk_mode != DISK_OPT && P.ir_kind == EGGPHOT_IR_CS17) v += rc2 * (real)P.rows_disk_pah[ri];
                if (P.apply_igm)
                    v *= igm_factor<real>(j, P.grid_disk.lz, P.grid_disk.l0, P.grid_disk.l1, P.grid_disk.l2, ig[0], ig[1], ig[2]);
                SD[l][j] = v;
            }
            for (int j = 0; j < nB; ++j) {
                real v = (real)P.rows_bulge_opt[(size_t)G.opt_sed_bulge[g] * nB + j];
                if (P.apply_igm)
                    v *= igm_factor<real>(j, P.grid_bulge.lz, P.grid_bulge.l0, P.grid_bulge.l1, P.grid_bulge.l2, ig[0], ig[1], ig[2]);
                SB[l][j] = v;
            }
            if (nl) {
                add_lines_lane<real>(P, P.grid_disk, SD[l], G.line_lum_disk + g * nl, G.vdisp[g], 1.0);
                if (P.has_bulge && G.line_lum_bulge)
                    add_lines_lane<real>(P, P.grid_bulge, SB[l], G.line_lum_bulge + g * nl, G.vdisp[g], a_b[l] > 0 ? 1.0 / a_b[l] : 0.0);
            }
        }
        // block sums per output column (a filter cut into several tables adds them up), tables in group order
        std::vector<double> resd(nb * kLanes, 0.0), resb(nb * kLanes, 0.0);
        std::vector<double> cfirst(nb, 0.0), clast(nb, 0.0);
        for (const BandGroup &grp : P.groups2) {
            const uint8_t *blob = grp.blob.data();
            const BandHeader2 *hdr = reinterpret_cast<const BandHeader2 *>(blob);
            for (size_t bi = 0; bi < grp.cols.size(); ++bi) {
                const BandHeader2 &h = hdr[bi];
                cfirst[h.out_col] = h.full_first; clast[h.out_col] = h.full_last;
                if (h.n == 0 || nD < 2) continue;
                Band2 bv;
                make_band2(blob, h, bv);
                // windows of the batch on the two component grids, with margins (extra nodes weigh exactly nothing);
                // the walk arrives at lo .. hi + 1
                const double xlo = h.f_first / omax * (1.0 - 1e-9), xhi = h.f_last / omin * (1.0 + 1e-9);
                for (int comp = P.has_bulge && nB >= 2 ? 0 : 1; comp < 2; ++comp) {
                    const Grid &gr = comp ? P.grid_disk : P.grid_bulge;
                    const GridNode *nodes = comp ? P.grid_nodes.data() : P.grid_nodes_b.data();
                    const int n = gr.n();
                    const int lo = std::max(last_le(gr.x, xlo), 0);
                    const int hi = std::min((int)(std::lower_bound(gr.x.begin(), gr.x.end(), xhi) - gr.x.begin()), n - 1);
                    for (int l = 0; l < ng; ++l) {
                        const double t = walk_window<real>(bv, nodes, opz[l], lo, hi + 1, comp ? SD[l].data() : SB[l].data());
                        (comp ? resd : resb)[h.out_col * kLanes + l] += t;
                    }
                }
            }
        }
        for (size_t col = 0; col < nb; ++col)
            for (int l = 0; l < ng; ++l) {
                const size_t g = perm[g0 + l];
                const bool covered = nD >= 2 && P.grid_disk.x[0] * opz[l] <= cfirst[col] && P.grid_disk.x[nD - 1] * opz[l] >= clast[col];
                const bool bulge_on = covered && P.has_bulge && nB >= 2 && P.grid_bulge.x[0] * opz[l] <= cfirst[col] &&
                                      P.grid_bulge.x[nB - 1] * opz[l] >= clast[col];
                double fdv = covered ? resd[col * kLanes + l] / opz[l] * aflux[l] : 0.0;
                double fbv = bulge_on ? resb[col * kLanes + l] / opz[l] * aflux[l] * a_b[l] : 0.0;
                if (!std::isfinite(fdv)) fdv = 0;
                if (!std::isfinite(fbv)) fbv = 0;
                fd[g * nb + col] = fdv;
                if (P.has_bulge) fb[g * nb + col] = fbv;
            }
        const double arf = EGGPHOT_LSUN2UJY / (1e-5 * 1e-5);
        for (int l = 0; l < ng; ++l) {
            const size_t g = perm[g0 + l];
            for (size_t r = 0; r < nrf; ++r)
                for (int comp = P.has_bulge ? 0 : 1; comp < 2; ++comp) {
                    const RfWeights &w = comp ? P.rf_disk[r] : P.rf_bulge[r];
                    const std::vector<real> &S = comp ? SD[l] : SB[l];
                    CompSum<real> acc;
                    acc.clear();
                    for (size_t k = 0; k < w.w.size(); ++k) acc.add(S[w.j0 + k] * (real)w.w[k]);
                    double fl = w.covered ? ((double)acc.s + (double)acc.c) * arf * (comp ? 1.0 : a_b[l]) : NAN;
                    double m = -2.5 * std::log10(fl) + 23.9;
                    (comp ? rd : rb)[g * nrf + r] = std::isfinite(m) ? m : 99.0;
                }
        }
    }
}

} // namespace

extern "C" {

// Returns 0 on success, else the eggphot status; msg (>= 256 bytes) receives the error text.
// Outputs are double: component fluxes [ngal][nband] and magnitudes [ngal][nrf] in sorted band order.
int emul_compute(const eggphot_libs *libs, const eggphot_filters *bands, const eggphot_filters *rfbands,
                 const eggphot_opts *opts, const eggphot_galaxies *gal, size_t ngal, double *flux_disk,
                 double *flux_bulge, double *rfmag_disk, double *rfmag_bulge, uint32_t *order, float *lambda, char *msg) {
    Prepared P;
    Status st = prepare(*libs, bands, rfbands, *opts, P);
    if (!st.ok()) {
        std::strncpy(msg, st.msg.c_str(), 255);
        msg[255] = 0;
        return st.code;
    }
    for (size_t k = 0; k < P.band_order.size(); ++k) { if (order) order[k] = P.band_order[k]; if (lambda) lambda[k] = P.band_lambda[k]; }
    const char *algo = getenv("EGGPHOT_EMUL_ALGO"); // 1 = node-per-lane kernel (phot_kernels.cu), default: lane kernel
    if (algo && algo[0] == '1') {
        if (P.precision == EGGPHOT_FP64) run<double>(P, *gal, ngal, flux_disk, flux_bulge, rfmag_disk, rfmag_bulge);
        else run<float>(P, *gal, ngal, flux_disk, flux_bulge, rfmag_disk, rfmag_bulge);
    } else {
        if (P.precision == EGGPHOT_FP64) run2<double>(P, *gal, ngal, flux_disk, flux_bulge, rfmag_disk, rfmag_bulge);
        else run2<float>(P, *gal, ngal, flux_disk, flux_bulge, rfmag_disk, rfmag_bulge);
    }
    return 0;
}

int emul_ngroups(const eggphot_libs *libs, const eggphot_filters *bands, const eggphot_opts *opts) {
    Prepared P;
    Status st = prepare(*libs, bands, nullptr, *opts, P);
    const char *algo = getenv("EGGPHOT_EMUL_ALGO");
    return st.ok() ? (int)((algo && algo[0] == '1') ? P.groups.size() : P.groups2.size()) : -st.code;
}

// largest group blob of the lane kernel's tables (bytes)
int emul_max_blob2(const eggphot_libs *libs, const eggphot_filters *bands, const eggphot_opts *opts) {
    Prepared P;
    Status st = prepare(*libs, bands, nullptr, *opts, P);
    size_t m = 0;
    for (const BandGroup &g : P.groups2) m = std::max(m, g.blob.size());
    return st.ok() ? (int)m : -st.code;
}
}
