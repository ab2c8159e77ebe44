// phot_tables.hpp — host-side preparation of everything the kernels read.
//
// Mirrors the one-off setup of egg-gencat: filters (src/egg-gencat.cpp:316-378), SED libraries
// (:388-541), and flattens the result into device-friendly tables:
//   * wavelength grids per galaxy component (optical, IR, or their merge_add union,
//     src/egg-utils.hpp:5-63) with the IGM index thresholds of src/egg-gencat.cpp:2123-2126;
//   * template rows Y = lambda * L resampled on those grids (what get_opt_sed / get_ir_sed /
//     merge_add produce per galaxy at :708-712, :2039-2057, :2118 — linear in the per-galaxy
//     scale factors, so done once per template here);
//   * per-band filter tables with prefix sums (see phot_core.cuh), packed into band groups that
//     fit a shared-memory budget;
//   * rest-frame (z = 0) quadrature weights for the rfbands (:2150-2158).
// Pure C++17, no CUDA: the result is plain byte blobs and vectors.
#ifndef EGGPHOT_PHOT_TABLES_HPP
#define EGGPHOT_PHOT_TABLES_HPP

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/eggphot.h"
#include "phot_core.cuh"
#include "phot_lane.cuh"

namespace eggphot {

struct Status {
    int code = EGGPHOT_OK;
    std::string msg;
    bool ok() const { return code == EGGPHOT_OK; }
    static Status error(int c, std::string m) { Status s; s.code = c; s.msg = std::move(m); return s; }
};

struct Filter {
    std::vector<double> lam, res;
    double rlam = 0; // reference wavelength (gencat:356 or [vif] get_filter)
};

// A rest-frame wavelength grid shared by all templates of a component.
constexpr int kLutPerOctave = 64;

struct Grid {
    std::vector<double> x;    // um, strictly increasing
    int lz = 0, l0 = 0, l1 = 0, l2 = 0; // IGM index thresholds (last node <= 0.06/0.0912/0.1026/0.1216 um)
    // log2-spaced lookup for "last node <= v": lut[k] = last j with x[j] <= 2^(lut_e0 + k/kLutPerOctave)
    std::vector<uint16_t> lut;
    int lut_e0 = 0;
    std::vector<double> idx;  // 1/(x[j+1]-x[j]) (0 for the last node): rest-frame slopes are differences times this
    int n() const { return (int)x.size(); }
    void build_lut();         // also fills idx
};

enum DiskMode { DISK_BOTH = 0, DISK_OPT = 1, DISK_IR = 2 };

struct BandGroup {
    std::vector<uint32_t> cols;      // sorted-band columns in this group
    std::vector<uint8_t> blob;       // [BandHeader x nb][nodes, buckets ...] 16-byte aligned pieces
};

// Sparse rest-frame weights of one rfband on one grid:  flux_rf = A_rf * sum_j S_j * w[j - j0]
struct RfWeights {
    int j0 = 0;
    std::vector<double> w;
    bool covered = false;
};

struct Prepared {
    int precision = EGGPHOT_FP32C;
    bool apply_igm = false, no_stellar = false, no_dust = false, no_nebular = false;
    int line_profile = EGGPHOT_LINE_INTEGRAL;
    int nodes_mode = EGGPHOT_NODES_MERGED;       // node set of [vif] sed2flux: merged is the contract
    std::vector<float> line_lambda;

    std::vector<Filter> bands, rfbands;          // sorted
    std::vector<uint32_t> band_order, rfband_order;
    std::vector<float> band_lambda, rfband_lambda;

    // libraries after IMF / rebin
    uint32_t nopt_sed = 0, nopt = 0;
    std::vector<float> opt_lam, opt_sed;         // [nopt_sed][nopt]
    std::vector<uint8_t> opt_use;
    int ir_kind = EGGPHOT_IR_GENERIC;
    uint32_t nir_sed = 0, nir = 0;

    // component grids and template rows (double; converted to the kernel's real type at upload)
    bool has_bulge = false;
    DiskMode disk_mode = DISK_BOTH;
    Grid grid_bulge, grid_disk;
    // rows are [ntemplate][grid.n()], values Y = x * L(x) following merge_add's rules
    std::vector<double> rows_bulge_opt;  // optical templates on grid_bulge
    std::vector<double> rows_disk_opt;   // optical templates on grid_disk   (empty if DISK_IR)
    std::vector<double> rows_disk_ir;    // IR templates (generic SED / cs17 DUST) on grid_disk (empty if DISK_OPT)
    std::vector<double> rows_disk_pah;   // cs17 PAH on grid_disk

    // bulge evaluated on the disk grid (the same piece-wise linear function, so that both components share
    // one evaluation of Lambda per node): for every disk-grid node the optical node at or before it (index
    // into grid_bulge, -1 before the first) and its fractional position inside that optical interval
    std::vector<int32_t> u_oleft;
    std::vector<double> u_ofrac;
    std::vector<uint8_t> u_isopt;
    std::vector<double> u_dxo;     // rest-frame distance from an optical node to the next optical node, 0 on the
                                   // other nodes: the bulge's union cells start at optical nodes only

    std::vector<BandGroup> groups;
    std::vector<RfWeights> rf_bulge, rf_disk;    // [nrf]

    // ---- tables of the lane kernel (phot_lane.cuh): band groups with node records, the disk grid as the walk reads it
    std::vector<BandGroup> groups2;      // blobs hold BandHeader2 + records
    std::vector<GridNode> grid_nodes;    // [lane_grid_rows(nD)]: the disk grid, two virtual nodes far to the red, padding
    std::vector<GridNode> grid_nodes_b;  // [lane_grid_rows(nB)]: the optical grid alone, likewise ({x, idxp} only)
    std::vector<int32_t> next_opt;       // [lane_grid_rows(nD)]: first optical node after j (nD = the virtual node)
    std::vector<int32_t> opt_node;       // [nB]: disk-grid index of every optical node
    // node sets `sed` and `filter`: the sorted filters as they are (wavelength, response), concatenated; alt_off[nband + 1]
    std::vector<double> alt_fl, alt_fr;
    std::vector<uint32_t> alt_off;
};


Status prepare(const eggphot_libs &libs, const eggphot_filters *bands, const eggphot_filters *rfbands,
               const eggphot_opts &opts, Prepared &out);

// pieces exposed for unit tests
Status prepare_filters(const eggphot_filters *raw, const eggphot_opts &opts, std::vector<Filter> &sorted,
                       std::vector<uint32_t> &order, std::vector<float> &lambda);
Status prepare_opt_lib(const eggphot_libs &libs, const eggphot_opts &opts, uint32_t &nopt,
                       std::vector<float> &lam, std::vector<float> &sed);

} // namespace eggphot
#endif
