// gencat_main.cpp — egg-gencat-b200: the photometry stage of egg-gencat as a stand-alone program.
//
// Keeps egg-gencat's `key=value` command line (src/egg-gencat.cpp:137-149; `-key=value` is equally
// valid, a bare `key` means true, lists are `[a,b,c]`) and its column-oriented FITS layout
// (src/egg-gencat.cpp:2253-2286) for the block `if (!no_flux) {...}` (:2024-2240).  The galaxy
// generation before that block (:610-2022: stochastic draws) is not rebuilt: the per-galaxy columns
// it produces are read from an existing egg catalog, `input_props=<catalog.fits>` — every egg output
// catalog holds them next to the fluxes, so any catalog of a real egg install can be re-computed.
//
// All computation happens in libeggphot.so on the GPU(s); there is no CPU path.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <regex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/eggphot.h"
#include "../../include/eggprops.h"
#include "fitscol.hpp"

namespace {

void note(const std::string &m) { std::cout << "note: " << m << std::endl; }
void warning(const std::string &m) { std::cerr << "warning: " << m << std::endl; }
void error(const std::string &m) { std::cerr << "error: " << m << std::endl; }

#ifndef EGG_SHARE_DIR
#define EGG_SHARE_DIR ""
#endif
#ifndef FILTER_DB_DIR
#define FILTER_DB_DIR ""
#endif

std::string env_or(const char *name, const std::string &dflt) {
    const char *e = std::getenv(name);
    return e && *e ? std::string(e) : dflt;
}
std::string with_slash(std::string d) {
    if (!d.empty() && d.back() != '/') d += '/';
    return d;
}
std::string remove_extension(const std::string &f) {
    const size_t sl = f.find_last_of('/'), dot = f.find_last_of('.');
    return (dot == std::string::npos || (sl != std::string::npos && dot < sl)) ? f : f.substr(0, dot);
}
std::string basename_of(const std::string &f) {
    const size_t sl = f.find_last_of('/');
    return sl == std::string::npos ? f : f.substr(sl + 1);
}

// ---- read_args ------------------------------------------------------------------------------------
struct Args {
    std::map<std::string, std::string> kv;
    bool has(const std::string &k) const { return kv.count(k) != 0; }
    std::string str(const std::string &k, const std::string &d = "") const { return has(k) ? kv.at(k) : d; }
    bool flag(const std::string &k) const {
        if (!has(k)) return false;
        const std::string v = kv.at(k);
        return !(v == "0" || v == "false");
    }
    double num(const std::string &k, double d) const { return has(k) ? std::atof(kv.at(k).c_str()) : d; }
    std::vector<std::string> list(const std::string &k, const std::vector<std::string> &d) const {
        if (!has(k)) return d;
        std::string v = kv.at(k);
        if (!v.empty() && v.front() == '[') v = v.substr(1);
        if (!v.empty() && v.back() == ']') v.pop_back();
        std::vector<std::string> out;
        std::string item;
        for (char c : v) {
            if (c == ',') { if (!item.empty()) out.push_back(item); item.clear(); }
            else if (c != ' ' && c != '"' && c != '\'') item += c;
        }
        if (!item.empty()) out.push_back(item);
        return out;
    }
};

Args read_args(int argc, char *argv[]) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        while (!s.empty() && s[0] == '-') s = s.substr(1);
        const size_t eq = s.find('=');
        std::string key = eq == std::string::npos ? s : s.substr(0, eq);
        std::replace(key.begin(), key.end(), '-', '_'); // `filter-db=` is `filter_db=`
        a.kv[key] = eq == std::string::npos ? "1" : s.substr(eq + 1);
    }
    if (a.has("list_band")) a.kv["list_bands"] = a.kv["list_band"]; // the manual uses both spellings
    return a;
}

// ---- filters ---------------------------------------------------------------------------------------
struct Curve {
    std::string name;
    std::vector<double> lam, res;
};

double trapz(const std::vector<double> &x, const std::vector<double> &y) {
    double s = 0;
    for (size_t i = 0; i + 1 < x.size(); ++i) s += 0.5 * (y[i] + y[i + 1]) * (x[i + 1] - x[i]);
    return s;
}

bool read_curve(const std::string &path, Curve &c) {
    fitscol::Table t = fitscol::read_table(path);
    const fitscol::Column *l = t.find("LAM"), *r = t.find("RES");
    if (!l || !r) { error("filter file '" + path + "' has no LAM/RES columns"); return false; }
    c.lam = fitscol::as_double(*l);
    c.res = fitscol::as_double(*r);
    if (c.lam.size() != c.res.size() || c.lam.size() < 2) { error("filter file '" + path + "' is empty or inconsistent"); return false; }
    return true;
}

// [vif] get_filters: plain names from the data base, `name:file:path`, `name:range:lmin:lmax`
// (doc/doc-egg-gencat.tex:254-260)
bool get_filters(const std::vector<std::pair<std::string, std::string>> &db, const std::vector<std::string> &bands,
                 std::vector<Curve> &out) {
    for (const std::string &b : bands) {
        Curve c;
        std::vector<std::string> part;
        {
            std::string item;
            for (char ch : b) { if (ch == ':') { part.push_back(item); item.clear(); } else item += ch; }
            part.push_back(item);
        }
        c.name = part[0];
        if (part.size() >= 3 && part[1] == "file") {
            std::string path = part[2];
            for (size_t k = 3; k < part.size(); ++k) path += ":" + part[k];
            if (!read_curve(path, c)) return false;
        } else if (part.size() == 4 && part[1] == "range") {
            const double lo = std::atof(part[2].c_str()), hi = std::atof(part[3].c_str());
            if (!(hi > lo)) { error("band '" + b + "': bad wavelength range"); return false; }
            c.lam = {lo, hi}; // top-hat of unit integral
            c.res = {1.0 / (hi - lo), 1.0 / (hi - lo)};
        } else {
            bool found = false;
            for (const auto &e : db)
                if (e.first == b) { found = read_curve(e.second, c); if (!found) return false; break; }
            if (!found) {
                error("unknown filter '" + b + "'");
                std::string near;
                for (const auto &e : db)
                    if (e.first.find(b.substr(0, b.find('-'))) == 0) near += " " + e.first;
                if (!near.empty()) note("filters of the same instrument:" + near);
                return false;
            }
        }
        out.push_back(std::move(c));
    }
    return true;
}

int list_bands(const std::string &pattern, const std::string &db_file) {
    if (pattern == "1") std::cout << "List of available bands:" << std::endl;
    else std::cout << "List of available bands (filter: '" << pattern << "'):" << std::endl;
    std::ifstream probe(db_file);
    if (!probe) { error("could not find filter database file '" + db_file + "'"); return 1; }
    struct Row { std::string name; float rlam, width; };
    std::vector<Row> rows;
    const std::regex re(pattern == "1" ? ".*" : pattern, std::regex::extended);
    for (const auto &e : fitscol::read_filter_db(db_file)) {
        if (pattern != "1" && !std::regex_search(e.first, re)) continue;
        Curve c;
        if (!read_curve(e.second, c)) continue;
        const double mx = *std::max_element(c.res.begin(), c.res.end());
        if (mx <= 0) { warning("zero transmission for filter " + e.first); continue; }
        std::vector<double> lr(c.lam.size());
        for (size_t i = 0; i < lr.size(); ++i) lr[i] = c.lam[i] * c.res[i];
        size_t a = c.lam.size(), b = 0;
        for (size_t i = 0; i < c.res.size(); ++i)
            if (c.res[i] > 0.5 * mx) { a = std::min(a, i); b = i; }
        rows.push_back({e.first, (float)trapz(c.lam, lr), (float)(c.lam[b] - c.lam[a])}); // src/egg-gencat.cpp:195-199
    }
    std::stable_sort(rows.begin(), rows.end(), [](const Row &x, const Row &y) {
        const std::string ix = x.name.substr(0, x.name.find('-')), iy = y.name.substr(0, y.name.find('-'));
        return ix != iy ? ix < iy : x.rlam < y.rlam;
    });
    auto fmt = [](float v) { char b[32]; std::snprintf(b, sizeof b, "%g", v); return std::string(b); };
    size_t wn = 0, wl = 0;
    for (const Row &r : rows) { wn = std::max(wn, r.name.size()); wl = std::max(wl, (fmt(r.rlam) + " um, ").size()); }
    for (const Row &r : rows) {
        std::string n = r.name, l = fmt(r.rlam) + " um, ";
        n.resize(wn, ' ');
        l.resize(wl, ' ');
        std::cout << " - " << n << "  ref-lam = " << l << "FWHM = " << fmt(r.width) << " um" << std::endl;
    }
    return 0;
}

void print_help() {
    std::cout <<
        "egg-gencat-b200: the photometry stage of egg-gencat (fluxes and rest-frame magnitudes) on NVIDIA B200\n"
        "This program replaces the block `if (!no_flux)` of egg-gencat only: catalog generation options (area, mmin, zmin,\n"
        "...) are accepted and ignored, and without input_props= or input_cat= it exits with an error.\n"
        "usage: egg-gencat-b200 input_props=catalog.fits [options]\n"
        "       egg-gencat-b200 input_cat=catalog.fits [options]\n\n"
        "  input_cat=FILE     catalog holding ID RA DEC Z M PASSIVE: every other property is generated on the GPU\n"
        "                     (src/egg-gencat.cpp:1360-2022; seed=42 no_random no_passive_lir magdis_tdust ms_disp=0.3)\n"
        "  input_props=FILE   egg catalog holding the per-galaxy columns (Z D M_DISK M_BULGE M2LCOR OPT_SED_DISK\n"
        "                     OPT_SED_BULGE IR_SED LIR FPAH MDUST IGM_ABS VDISP LINE_LUM_DISK LINE_LUM_BULGE LINE_LAMBDA)\n"
        "  bands=[...]        bands for observed fluxes (default: the 23 bands of egg-gencat)\n"
        "  rfbands=[...]      bands for rest-frame absolute magnitudes (default: none)\n"
        "  filter_db=FILE     filter data base (db.dat); default $EGG_FILTERS_PATH/db.dat\n"
        "  opt_lib=FILE ir_lib=FILE opt_lib_imf=salpeter|chabrier|kroupa opt_lib_min_step=0.001\n"
        "  igm=none|constant|madau95|inoue14 (default inoue14)\n"
        "  no_stellar no_dust no_nebular no_flux save_sed\n"
        "  trim_filters filter_flambda filter_photons filter_normalize\n"
        "  out=FILE verbose help list_bands[=regex]\n"
        "  precision=fp32c|fp64   arithmetic mode (1e-5 / 1e-10 per band against the reference algorithm)\n"
        "  check [check_tol=1e-5] compare the computed FLUX* / RFMAG* with the columns already in input_props (any catalog\n"
        "                         written by egg-gencat is a test vector); exit status 2 if a band differs by more\n"
        "  sed2flux_nodes=merged|sed|filter  line_profile=integral|average   conventions of the vif library to try with check\n"
        "  gpus=N                 shard the galaxies over N GPUs of this box\n"
        "  maglim=M selection_band=B [zmin= zmax= dz= min_dz= max_dz= mmin_strict=4 mmax=12]   without an input catalog: print\n"
        "                         the redshift-dependent mass limit z_mmin of egg-gencat's magnitude-limited mode per redshift\n"
        "                         slice (src/egg-gencat.cpp:903-969) and exit\n"
        "Catalog generation options of egg-gencat (area, mmin, zmin, ...) are accepted and ignored: this program\n"
        "replaces the block `if (!no_flux)` of egg-gencat only.\n";
}

// Column order of the reference's writer (src/egg-gencat.cpp:2253-2286): the main table, then the flux, line and
// rest-frame groups; columns egg does not know keep their relative order after them.
void order_like_the_reference(fitscol::Table &t) {
    static const char *const order[] = {
        "ID", "RA", "DEC", "CLUSTERED", "Z", "D", "M", "SFR", "RSB", "M2LCOR", "DISK_ANGLE", "DISK_RADIUS", "DISK_RATIO",
        "BULGE_ANGLE", "BULGE_RADIUS", "BULGE_RATIO", "BT", "M_DISK", "M_BULGE", "AV_DISK", "AV_BULGE", "PASSIVE",
        "RFUV_BULGE", "RFUV_DISK", "RFVJ_BULGE", "RFVJ_DISK", "SFRIR", "SFRUV", "IRX", "LIR", "IR_SED", "OPT_SED_BULGE",
        "OPT_SED_DISK", "TDUST", "IR8", "FPAH", "MDUST", "IGM_ABS", "ZB", "CMD",
        "FLUX", "FLUX_DISK", "FLUX_BULGE", "BANDS", "LAMBDA",
        "VDISP", "AVLINES_DISK", "AVLINES_BULGE", "LINE_LUM", "LINE_LUM_DISK", "LINE_LUM_BULGE", "LINES", "LINE_LAMBDA",
        "RFMAG", "RFMAG_DISK", "RFMAG_BULGE", "RFBANDS", "RFLAMBDA"};
    std::vector<fitscol::Column> out;
    std::vector<bool> used(t.cols.size(), false);
    for (const char *name : order)
        for (size_t i = 0; i < t.cols.size(); ++i)
            if (!used[i] && t.cols[i].name == name) { out.push_back(std::move(t.cols[i])); used[i] = true; }
    for (size_t i = 0; i < t.cols.size(); ++i)
        if (!used[i]) out.push_back(std::move(t.cols[i]));
    t.cols = std::move(out);
}

// `maglim= selection_band=` without an input catalog: the redshift-dependent mass limit of src/egg-gencat.cpp:903-969,
// printed per redshift slice (the slices are those of :803-839; the strict lower mass, which the reference takes from its
// mass-function file, :885, is the option mmin_strict).
int run_maglim(const Args &a, const std::string &filter_db_file, const std::string &opt_lib_file, bool verbose) {
    const double maglim = a.num("maglim", 0.0), zmin = a.num("zmin", 0.05), zmax = a.num("zmax", 10.5); // :34-35
    const double dz = a.num("dz", 0.05), min_dz = a.num("min_dz", 0.001), max_dz = a.num("max_dz", 0.5); // :37-39
    const double mmin_strict = a.num("mmin_strict", 4.0), mmax = a.num("mmax", 12.0);                 // :33
    const std::string selection_band = a.str("selection_band");
    if (selection_band.empty()) { // :240-243
        error("requesting a magnitude limited catalog, please provide the selection band (selection_band=...)");
        return 1;
    }
    if (min_dz > max_dz) { error("min_dz must be smaller than max_dz"); return 1; }
    if (!(zmin > 0) || !(zmax > zmin)) { error("need 0 < zmin < zmax"); return 1; }
    // make_zbins(zmin, zmax, dz, min_dz, max_dz), :803-836
    std::vector<double> tzb;
    for (double z1 = zmin; z1 < zmax;) {
        tzb.push_back(z1);
        z1 *= 1.0 + dz;
        if (z1 - tzb.back() < min_dz) z1 = tzb.back() + min_dz;
        else if (z1 - tzb.back() > max_dz) z1 = tzb.back() + max_dz;
    }
    const double min_end_dz = std::min(0.5 * zmax * dz, min_dz);
    if (tzb.size() > 1 && zmax - tzb.back() < min_end_dz) tzb.back() = zmax; else tzb.push_back(zmax);
    const size_t nz = tzb.size() - 1;
    std::vector<double> lo(tzb.begin(), tzb.end() - 1), hi(tzb.begin() + 1, tzb.end());

    const auto db = fitscol::read_filter_db(filter_db_file);
    std::vector<Curve> fsel;
    if (!get_filters(db, {selection_band}, fsel)) return 1;
    fitscol::Table t = fitscol::read_table(opt_lib_file);
    const fitscol::Column *cl = t.find("LAM"), *cs = t.find("SED"), *cu = t.find("USE"), *bu = t.find("BUV"), *bv = t.find("BVJ"), *ca = t.find("AV");
    if (!cl || !cs || !cu || !bu || !bv || !ca || cl->dims.size() != 3 || bu->dims.size() != 2) {
        error("optical library '" + opt_lib_file + "' must hold LAM, SED [nuv][nvj][nlam], USE, BUV, BVJ and AV");
        return 1;
    }
    const std::vector<float> olam = fitscol::as_float(*cl), osed = fitscol::as_float(*cs), buv = fitscol::as_float(*bu), bvj = fitscol::as_float(*bv),
                             av = fitscol::as_float(*ca);
    const std::vector<uint8_t> ouse = fitscol::as_bool(*cu);
    eggphot_libs L{};
    L.nuv = (uint32_t)cl->dims[0]; L.nvj = (uint32_t)cl->dims[1]; L.nopt = (uint32_t)cl->dims[2];
    L.opt_lam = olam.data(); L.opt_sed = osed.data(); L.opt_use = ouse.data();
    const uint32_t nf = (uint32_t)fsel[0].lam.size();
    const double *fl = fsel[0].lam.data(), *fr = fsel[0].res.data();
    const eggphot_filters B{1, &nf, &fl, &fr};
    const std::string imf = a.str("opt_lib_imf", "salpeter");
    eggphot_opts O{};
    O.precision = a.str("precision", "fp32c") == "fp64" ? EGGPHOT_FP64 : EGGPHOT_FP32C;
    O.igm = "none"; O.no_dust = 1; O.no_nebular = 1; // the estimator sees stellar light only (:943-952)
    O.opt_lib_imf = imf.c_str();
    O.opt_lib_min_step = a.num("opt_lib_min_step", 0.001);
    O.filter_flambda = a.flag("filter_flambda"); O.filter_photons = a.flag("filter_photons");
    O.filter_normalize = a.has("filter_normalize") ? a.flag("filter_normalize") : true;
    O.trim_filters = a.flag("trim_filters");
    eggphot_ctx *pc = nullptr;
    int rc = eggphot_create(&L, &B, nullptr, &O, &pc);
    if (rc) { error(pc ? eggphot_last_error(pc) : "allocation failed"); eggphot_destroy(pc); return 1; }
    eggprops_libs PL{};
    PL.nuv = L.nuv; PL.nvj = L.nvj; PL.buv = buv.data(); PL.bvj = bvj.data(); PL.av = av.data(); PL.use = ouse.data();
    PL.ir_lib = EGGPROPS_IR_SINGLE; PL.nir_sed = 1; // (the trial galaxies have no dust emission)
    eggprops_opts PO{};
    PO.no_random = 1; PO.no_nebular = 1; PO.ms_disp = 0.3; PO.igm = EGGPROPS_IGM_NONE;
    PO.h0 = 70.0; PO.omega_m = 0.3; PO.omega_l = 0.7;
    eggprops_ctx *qc = nullptr;
    rc = eggprops_create(&PL, &PO, &qc);
    if (rc) { error(qc ? eggprops_last_error(qc) : "allocation failed"); eggprops_destroy(qc); eggphot_destroy(pc); return 1; }
    if (verbose) note("estimating redshift-dependent mass limit...");
    std::vector<float> z_mmin(nz);
    rc = eggphot_maglim(pc, qc, lo.data(), hi.data(), (uint32_t)nz, maglim, mmin_strict, mmax, 1000, z_mmin.data());
    if (rc) error(eggphot_last_error(pc));
    eggprops_destroy(qc);
    eggphot_destroy(pc);
    if (rc) return 1;
    std::cout.precision(12);
    std::cout << "# z_low z_up z_mmin   (maglim=" << maglim << " in " << selection_band << ")\n";
    for (size_t k = 0; k < nz; ++k) std::cout << lo[k] << ' ' << hi[k] << ' ' << z_mmin[k] << '\n';
    const float mmin = *std::min_element(z_mmin.begin(), z_mmin.end());
    note("will generate masses from as low as " + std::to_string(mmin) + ", up to " + std::to_string(mmax)); // :967
    return 0;
}

struct Shard {
    int device = 0;
    size_t first = 0, count = 0;
    int status = 0;
    std::string msg;
};

} // namespace

int main(int argc, char *argv[]) {
    const Args a = read_args(argc, argv);
    if (a.flag("help")) { print_help(); return 0; }
    const bool verbose = a.flag("verbose");
    const std::string share = with_slash(env_or("EGG_SHARE_DIR", EGG_SHARE_DIR));
    const std::string fdir = with_slash(env_or("EGG_FILTERS_PATH", std::string(FILTER_DB_DIR).empty() ? share + "filter-db" : FILTER_DB_DIR));
    const std::string filter_db_file = a.str("filter_db", fdir + "db.dat");
    try {
        if (a.has("list_bands")) return list_bands(a.str("list_bands"), filter_db_file);

        // ---- options (names and defaults of src/egg-gencat.cpp:22-133)
        const std::string igm = a.str("igm", "inoue14");
        if (igm != "none" && igm != "constant" && igm != "madau95" && igm != "inoue14") { // :286-289
            error("unknown IGM prescription '" + igm + "'");
            error("possible values are: 'none', 'constant', 'madau95', or 'inoue14'");
            return 1;
        }
        bool no_flux = a.flag("no_flux");
        const bool no_stellar = a.flag("no_stellar"), no_dust = a.flag("no_dust"), no_nebular = a.flag("no_nebular");
        if (no_dust && no_stellar) no_flux = true; // :247-249
        const bool save_sed = a.flag("save_sed");
        std::string out_file = a.str("out");
        if (out_file.empty()) {
            char buf[32];
            const time_t t = time(nullptr);
            strftime(buf, sizeof buf, "%Y%m%d", localtime(&t));
            out_file = std::string("egg-") + buf + ".fits"; // :259-261
        }
        if (verbose) note("will save catalog in '" + out_file + "'");
        const bool from_cat = !a.has("input_props") && a.has("input_cat");
        const std::string input = a.str("input_props", a.str("input_cat"));
        if (input.empty() && a.has("maglim")) {
            const std::string lib = a.str("opt_lib", share + (igm == "constant" ? "opt_lib_fast.fits" : "opt_lib_fast_noigm.fits")); // :296-302
            return run_maglim(a, filter_db_file, lib, verbose);
        }
        if (input.empty()) {
            error("this program replaces the photometry stage of egg-gencat only (src/egg-gencat.cpp:2024-2240)");
            error("pass input_props=<egg catalog> holding the per-galaxy columns, or input_cat=<catalog with Z, M, PASSIVE>; drawing galaxies from the mass functions (area=, mmin=, ...) is out of scope");
            return 1;
        }
        for (const char *k : {"area", "mmin", "maglim", "zmin", "zmax", "ra0", "dec0", "mass_func", "no_pos", "no_clust"})
            if (a.has(k) && verbose) note(std::string("option '") + k + "' only affects catalog generation and is ignored");
        const std::vector<std::string> dflt_bands = {"vimos-u", "hst-f435w", "hst-f606w", "hst-f775w", "hst-f814w", "hst-f850lp",
            "hst-f105w", "hst-f125w", "hst-f140w", "hst-f160w", "hawki-Ks", "spitzer-irac1", "spitzer-irac2", "spitzer-irac3",
            "spitzer-irac4", "spitzer-irs16", "spitzer-mips24", "herschel-pacs70", "herschel-pacs100", "herschel-pacs160",
            "herschel-spire250", "herschel-spire350", "herschel-spire500"}; // :125-130
        std::vector<std::string> bands = a.list("bands", dflt_bands), rfbands = a.list("rfbands", {});
        const std::string ir_lib_file = a.str("ir_lib", share + "ir_lib_cs17.fits"); // :88
        const std::string opt_lib_file = a.str("opt_lib", share + (igm == "constant" ? "opt_lib_fast.fits" : "opt_lib_fast_noigm.fits")); // :296-302
        const std::string imf = a.str("opt_lib_imf", "salpeter");
        const std::string precision = a.str("precision", "fp32c");
        if (precision != "fp32c" && precision != "fp64") { error("precision must be fp32c or fp64"); return 1; }
        const int gpus = std::max(1, (int)a.num("gpus", 1));

        // ---- catalog
        if (verbose) note("reading galaxy properties from '" + input + "'...");
        fitscol::Table cat;
        const bool is_fits = input.size() >= 5 && input.compare(input.size() - 5, 5, ".fits") == 0;
        if (from_cat && !is_fits) {
            // plain text: columns ID RA DEC Z M PASSIVE, '#' comments ([vif] ascii::read_table, src/egg-gencat.cpp:765-767)
            std::ifstream f(input);
            if (!f) { error("could not open input catalog '" + input + "'"); return 1; }
            std::vector<uint64_t> id;
            std::vector<double> ra, dec;
            std::vector<float> cz, cm;
            std::vector<uint8_t> cp;
            std::string line;
            size_t lineno = 0;
            while (std::getline(f, line)) {
                ++lineno;
                const size_t a = line.find_first_not_of(" \t\r");
                if (a == std::string::npos || line[a] == '#') continue;
                std::istringstream is(line);
                double vid, vra, vdec, vz, vm, vp;
                if (!(is >> vid >> vra >> vdec >> vz >> vm >> vp)) {
                    error("input catalog '" + input + "', line " + std::to_string(lineno) + ": expected ID RA DEC Z M PASSIVE");
                    return 1;
                }
                id.push_back((uint64_t)vid); ra.push_back(vra); dec.push_back(vdec);
                cz.push_back((float)vz); cm.push_back((float)vm); cp.push_back(vp != 0 ? 'T' : 'F');
            }
            cat.set(fitscol::make_u64("ID", id)); cat.set(fitscol::make_f64("RA", ra)); cat.set(fitscol::make_f64("DEC", dec));
            cat.set(fitscol::make_f32("Z", cz)); cat.set(fitscol::make_f32("M", cm));
            fitscol::Column pc;
            pc.name = "PASSIVE"; pc.type = 'L'; pc.dims = {(long)cp.size()}; pc.data = cp;
            cat.set(pc);
        } else cat = fitscol::read_table(input);
        if (a.flag("check")) {
            // check without bands= / rfbands=: the bands the catalog itself was written with (its BANDS / RFBANDS columns)
            auto from_column = [&](const char *col, const char *key, std::vector<std::string> &dst) {
                const fitscol::Column *c = cat.find(col);
                if (a.has(key) || !c) return;
                dst = fitscol::as_strings(*c);
                for (auto &v : dst) while (!v.empty() && v.back() == ' ') v.pop_back();
            };
            from_column("BANDS", "bands", bands);
            from_column("RFBANDS", "rfbands", rfbands);
        }
        if (from_cat) {
            // ---- galaxy properties on the GPU (src/egg-gencat.cpp:714-779 for the input, :1360-2022 for the recipes)
            for (const char *k : {"Z", "M", "PASSIVE"})
                if (!cat.find(k)) { error(std::string("input catalog has no column ") + k + " (needed: ID RA DEC Z M PASSIVE)"); return 1; }
            const std::vector<float> cz = fitscol::as_float(*cat.find("Z")), cm = fitscol::as_float(*cat.find("M"));
            const std::vector<uint8_t> cp = fitscol::as_bool(*cat.find("PASSIVE"));
            const size_t n = cz.size();
            if (cm.size() != n || cp.size() != n) { error("catalog columns have inconsistent lengths"); return 1; }
            size_t nbad = 0;
            for (size_t i = 0; i < n; ++i) nbad += !(cz[i] > 0) || !(cm[i] > 4) || cm[i] > 13;
            if (nbad) { // :769-774
                error("input catalog contains " + std::to_string(nbad) + " unphysical galax" + (nbad == 1 ? "y" : "ies"));
                note("must have z > 0 and 4 < m < 13");
                return 1;
            }
            if (verbose) note("found " + std::to_string(n) + " galaxies");
            for (const char *k : {"LIR", "TDUST", "IR8"})
                if (cat.find(k)) { error(std::string("user-provided ") + k + " values in the input catalog are not supported"); return 1; }
            eggprops_libs PL{};
            std::vector<float> buv, bvj, av;
            std::vector<uint8_t> use;
            std::vector<double> ct, cld, clp, c8d, c8p;
            if (!no_stellar) {
                fitscol::Table t = fitscol::read_table(opt_lib_file);
                const fitscol::Column *cu = t.find("BUV"), *cv = t.find("BVJ"), *ca = t.find("AV"), *cs = t.find("USE");
                if (!cu || !cv || !ca || !cs || cu->dims.size() != 2) { error("optical library '" + opt_lib_file + "' must hold BUV, BVJ [2][n], AV and USE"); return 1; }
                buv = fitscol::as_float(*cu); bvj = fitscol::as_float(*cv); av = fitscol::as_float(*ca); use = fitscol::as_bool(*cs);
                PL.nuv = (uint32_t)cu->dims[1]; PL.nvj = (uint32_t)cv->dims[1];
                PL.buv = buv.data(); PL.bvj = bvj.data(); PL.av = av.data(); PL.use = use.data();
            }
            {
                fitscol::Table t = fitscol::read_table(ir_lib_file);
                const std::string base = basename_of(ir_lib_file);
                const fitscol::Column *cs = t.find(base == "ir_lib_cs17.fits" ? "DUST" : "SED");
                if (!cs) { error("missing LAM and/or SED columns in the IR library file"); return 1; }
                PL.nir_sed = cs->dims.size() == 2 ? (uint32_t)cs->dims[0] : 1;
                if (PL.nir_sed == 1) PL.ir_lib = EGGPROPS_IR_SINGLE;                        // :1682
                else if (base == "ir_lib_ce01.fits") PL.ir_lib = EGGPROPS_IR_CE01;
                else if (base == "ir_lib_m12.fits") PL.ir_lib = EGGPROPS_IR_M12;
                else if (base == "ir_lib_cs17.fits") {
                    PL.ir_lib = EGGPROPS_IR_CS17;
                    for (const char *k : {"TDUST", "LIR_DUST", "LIR_PAH", "L8_DUST", "L8_PAH"})
                        if (!t.find(k)) { error(std::string("IR library '") + ir_lib_file + "' has no column " + k); return 1; }
                    ct = fitscol::as_double(*t.find("TDUST")); cld = fitscol::as_double(*t.find("LIR_DUST")); clp = fitscol::as_double(*t.find("LIR_PAH"));
                    c8d = fitscol::as_double(*t.find("L8_DUST")); c8p = fitscol::as_double(*t.find("L8_PAH"));
                    PL.cs17_tdust = ct.data(); PL.cs17_lir_dust = cld.data(); PL.cs17_lir_pah = clp.data(); PL.cs17_l8_dust = c8d.data(); PL.cs17_l8_pah = c8p.data();
                } else { error("no calibration code available for the IR library '" + ir_lib_file + "'"); return 1; } // :1702
            }
            eggprops_opts PO{};
            PO.seed = (uint64_t)a.num("seed", 42);
            PO.no_random = a.flag("no_random"); PO.no_stellar = no_stellar; PO.no_nebular = no_nebular;
            PO.no_passive_lir = a.flag("no_passive_lir"); PO.magdis_tdust = a.flag("magdis_tdust");
            PO.ms_disp = a.num("ms_disp", 0.3);
            PO.igm = igm == "madau95" ? EGGPROPS_IGM_MADAU95 : (igm == "inoue14" ? EGGPROPS_IGM_INOUE14 : EGGPROPS_IGM_NONE);
            PO.h0 = 70.0; PO.omega_m = 0.3; PO.omega_l = 0.7;
            if (verbose) note("generating galaxy properties on the GPU...");
            eggprops_ctx *pc = nullptr;
            int rc = eggprops_create(&PL, &PO, &pc);
            if (rc) { error(pc ? eggprops_last_error(pc) : "allocation failed"); eggprops_destroy(pc); return 1; }
            struct F { const char *name; std::vector<float> v; size_t w; };
            std::vector<F> f32 = {{"D", {}, 1}, {"BT", {}, 1}, {"M_BULGE", {}, 1}, {"M_DISK", {}, 1}, {"DISK_ANGLE", {}, 1}, {"BULGE_ANGLE", {}, 1},
                {"DISK_RATIO", {}, 1}, {"BULGE_RATIO", {}, 1}, {"DISK_RADIUS", {}, 1}, {"BULGE_RADIUS", {}, 1}, {"SFR", {}, 1}, {"RSB", {}, 1}, {"IRX", {}, 1},
                {"SFRIR", {}, 1}, {"SFRUV", {}, 1}, {"RFUV_DISK", {}, 1}, {"RFVJ_DISK", {}, 1}, {"RFUV_BULGE", {}, 1}, {"RFVJ_BULGE", {}, 1}, {"M2LCOR", {}, 1},
                {"AV_DISK", {}, 1}, {"AV_BULGE", {}, 1}, {"TDUST", {}, 1}, {"IR8", {}, 1}, {"LIR", {}, 1}, {"FPAH", {}, 1}, {"MDUST", {}, 1}, {"IGM_ABS", {}, 3},
                {"AVLINES_DISK", {}, 1}, {"AVLINES_BULGE", {}, 1}, {"VDISP", {}, 1}, {"LINE_LUM_DISK", {}, EGGPROPS_NLINE}, {"LINE_LUM_BULGE", {}, EGGPROPS_NLINE}};
            for (F &f : f32) f.v.assign(n * f.w, 0.0f);
            std::vector<uint64_t> od(n), ob(n), is(n);
            eggprops_out E{};
            float **slots[] = {&E.d, &E.bt, &E.m_bulge, &E.m_disk, &E.disk_angle, &E.bulge_angle, &E.disk_ratio, &E.bulge_ratio, &E.disk_radius, &E.bulge_radius,
                &E.sfr, &E.rsb, &E.irx, &E.sfrir, &E.sfruv, &E.rfuv_disk, &E.rfvj_disk, &E.rfuv_bulge, &E.rfvj_bulge, &E.m2lcor, &E.av_disk, &E.av_bulge,
                &E.tdust, &E.ir8, &E.lir, &E.fpah, &E.mdust, &E.igm_abs, &E.avlines_disk, &E.avlines_bulge, &E.vdisp, &E.line_lum_disk, &E.line_lum_bulge};
            for (size_t k = 0; k < f32.size(); ++k) *slots[k] = f32[k].v.data();
            E.opt_sed_disk = od.data(); E.opt_sed_bulge = ob.data(); E.ir_sed = is.data();
            rc = eggprops_compute(pc, cz.data(), cm.data(), cp.data(), n, 0, &E);
            if (rc) { error(eggprops_last_error(pc)); eggprops_destroy(pc); return 1; }
            eggprops_destroy(pc);
            for (const F &f : f32) {
                const bool line_col = std::string(f.name).rfind("LINE_", 0) == 0 || std::string(f.name).rfind("AVLINES", 0) == 0 || std::string(f.name) == "VDISP";
                if (line_col && no_nebular) continue;
                cat.set(fitscol::make_f32(f.name, f.v, f.w > 1 ? std::vector<long>{(long)n, (long)f.w} : std::vector<long>{}));
            }
            cat.set(fitscol::make_u64("OPT_SED_DISK", od)); cat.set(fitscol::make_u64("OPT_SED_BULGE", ob)); cat.set(fitscol::make_u64("IR_SED", is));
            if (!no_nebular) {
                std::vector<std::string> names;
                std::vector<float> lam(eggprops_line_lambda(), eggprops_line_lambda() + EGGPROPS_NLINE), tot(n * EGGPROPS_NLINE);
                for (uint32_t l = 0; l < EGGPROPS_NLINE; ++l) names.push_back(eggprops_line_name(l));
                const std::vector<float> &ld = f32[31].v, &lb = f32[32].v;
                for (size_t i = 0; i < tot.size(); ++i) tot[i] = ld[i] + lb[i];                 // :2024
                cat.set(fitscol::make_strings("LINES", names)); cat.set(fitscol::make_f32("LINE_LAMBDA", lam));
                cat.set(fitscol::make_f32("LINE_LUM", tot, {(long)n, (long)EGGPROPS_NLINE}));
            }
        }
        auto need = [&](const char *name) -> const fitscol::Column & {
            const fitscol::Column *c = cat.find(name);
            if (!c) throw std::runtime_error(std::string("catalog '") + input + "' has no column " + name);
            return *c;
        };
        const std::vector<float> z = fitscol::as_float(need("Z")), d = fitscol::as_float(need("D"));
        const size_t ngal = z.size();
        const std::vector<float> m2lcor = fitscol::as_float(need("M2LCOR"));
        std::vector<float> m_disk, m_bulge, lir, fpah, mdust, igm_abs, vdisp, lld, llb, line_lambda;
        std::vector<uint64_t> osd, osb, irs;
        if (!no_stellar) {
            m_disk = fitscol::as_float(need("M_DISK")); m_bulge = fitscol::as_float(need("M_BULGE"));
            osd = fitscol::as_u64(need("OPT_SED_DISK")); osb = fitscol::as_u64(need("OPT_SED_BULGE"));
        }
        const bool cs17 = basename_of(ir_lib_file) == "ir_lib_cs17.fits"; // :401
        if (!no_dust) {
            irs = fitscol::as_u64(need("IR_SED"));
            if (cs17) { fpah = fitscol::as_float(need("FPAH")); mdust = fitscol::as_float(need("MDUST")); }
            else lir = fitscol::as_float(need("LIR"));
        }
        if (igm == "madau95" || igm == "inoue14") igm_abs = fitscol::as_float(need("IGM_ABS"));
        bool nebular = !no_nebular;
        if (nebular && !cat.find("LINE_LAMBDA")) {
            error("the catalog has no emission line columns (LINE_LAMBDA, LINE_LUM_DISK, VDISP); pass no_nebular to ignore lines");
            return 1;
        }
        if (nebular) {
            line_lambda = fitscol::as_float(need("LINE_LAMBDA"));
            vdisp = fitscol::as_float(need("VDISP"));
            lld = fitscol::as_float(need("LINE_LUM_DISK"));
            if (cat.find("LINE_LUM_BULGE")) llb = fitscol::as_float(need("LINE_LUM_BULGE"));
        }
        {   // every column the kernels index by galaxy must have ngal entries (x its second dimension)
            const size_t nl = line_lambda.size();
            const struct { const char *name; size_t have, want; } chk[] = {
                {"D", d.size(), ngal}, {"M2LCOR", m2lcor.size(), ngal},
                {"M_DISK", m_disk.size(), no_stellar ? 0 : ngal}, {"M_BULGE", m_bulge.size(), no_stellar ? 0 : ngal},
                {"OPT_SED_DISK", osd.size(), no_stellar ? 0 : ngal}, {"OPT_SED_BULGE", osb.size(), no_stellar ? 0 : ngal},
                {"IR_SED", irs.size(), no_dust ? 0 : ngal}, {"LIR", lir.size(), (no_dust || cs17) ? 0 : ngal},
                {"FPAH", fpah.size(), (!no_dust && cs17) ? ngal : 0}, {"MDUST", mdust.size(), (!no_dust && cs17) ? ngal : 0},
                {"IGM_ABS", igm_abs.size(), (igm == "madau95" || igm == "inoue14") ? 3 * ngal : 0},
                {"VDISP", vdisp.size(), nebular ? ngal : 0}, {"LINE_LUM_DISK", lld.size(), nebular ? ngal * nl : 0},
                {"LINE_LUM_BULGE", llb.size(), llb.empty() ? 0 : ngal * nl}};
            for (const auto &c : chk)
                if (c.have != c.want) {
                    error("catalog columns have inconsistent lengths (" + std::string(c.name) + " holds " + std::to_string(c.have) +
                          " values, expected " + std::to_string(c.want) + ")");
                    return 1;
                }
        }
        if (ngal == 0) {
            if (verbose) note("the catalog holds no galaxy: writing it unchanged");
            fitscol::write_table(out_file, cat);
            return 0;
        }

        if (no_flux) {
            if (verbose) note("no_flux: writing the catalog unchanged");
            fitscol::write_table(out_file, cat);
            return 0;
        }

        // ---- filters and libraries (src/egg-gencat.cpp:316-327, 401-458)
        if (verbose) note("initializing filters...");
        const auto db = fitscol::read_filter_db(filter_db_file);
        std::vector<Curve> fb, fr;
        if (!get_filters(db, bands, fb) || !get_filters(db, rfbands, fr)) return 1;
        if (verbose) note("initializing SED libraries...");
        eggphot_libs L{};
        std::vector<float> olam, osed;
        std::vector<uint8_t> ouse;
        std::vector<double> ilam, ised, ipah;
        if (!no_stellar) {
            fitscol::Table t = fitscol::read_table(opt_lib_file);
            const fitscol::Column *cl = t.find("LAM"), *cs = t.find("SED"), *cu = t.find("USE");
            if (!cl || !cs || !cu || cl->dims.size() != 3) { error("optical library '" + opt_lib_file + "' must hold LAM, SED [nuv][nvj][nlam] and USE"); return 1; }
            olam = fitscol::as_float(*cl); osed = fitscol::as_float(*cs); ouse = fitscol::as_bool(*cu);
            L.nuv = (uint32_t)cl->dims[0]; L.nvj = (uint32_t)cl->dims[1]; L.nopt = (uint32_t)cl->dims[2];
            L.opt_lam = olam.data(); L.opt_sed = osed.data(); L.opt_use = ouse.data();
        }
        if (!no_dust) {
            fitscol::Table t = fitscol::read_table(ir_lib_file);
            const fitscol::Column *cl = t.find("LAM"), *cs = t.find(cs17 ? "DUST" : "SED"), *cp = t.find("PAH");
            if (!cl || !cs || (cs17 && !cp)) { error("missing LAM and/or SED columns in the IR library file"); return 1; } // :438-441
            if (cs->dims.size() > 2) { error("IR library must contain either 1D or 2D columns LAM and SED"); return 1; }
            ilam = fitscol::as_double(*cl); ised = fitscol::as_double(*cs);
            if (cs17) ipah = fitscol::as_double(*cp);
            L.ir_kind = cs17 ? EGGPHOT_IR_CS17 : EGGPHOT_IR_GENERIC;
            L.nir_sed = cs->dims.size() == 2 ? (uint32_t)cs->dims[0] : 1;
            L.nir = (uint32_t)cs->dims.back();
            L.ir_lam = ilam.data(); L.ir_sed = ised.data(); L.ir_pah = cs17 ? ipah.data() : nullptr;
        }
        auto pack = [](const std::vector<Curve> &f, std::vector<uint32_t> &n, std::vector<const double *> &l, std::vector<const double *> &r) {
            for (const Curve &c : f) { n.push_back((uint32_t)c.lam.size()); l.push_back(c.lam.data()); r.push_back(c.res.data()); }
            return eggphot_filters{(uint32_t)f.size(), n.data(), l.data(), r.data()};
        };
        std::vector<uint32_t> nb_, nr_;
        std::vector<const double *> lb_, rb_, lr_, rr_;
        const eggphot_filters B = pack(fb, nb_, lb_, rb_), R = pack(fr, nr_, lr_, rr_);
        eggphot_opts O{};
        O.precision = precision == "fp64" ? EGGPHOT_FP64 : EGGPHOT_FP32C;
        O.igm = igm.c_str();
        O.no_stellar = no_stellar; O.no_dust = no_dust; O.no_nebular = !nebular;
        O.opt_lib_imf = imf.c_str();
        O.opt_lib_min_step = a.num("opt_lib_min_step", 0.001);
        O.filter_flambda = a.flag("filter_flambda"); O.filter_photons = a.flag("filter_photons");
        O.filter_normalize = a.has("filter_normalize") ? a.flag("filter_normalize") : true; // default true (src/egg-gencat.cpp:98)
        O.trim_filters = a.flag("trim_filters");
        O.nline = (uint32_t)line_lambda.size(); O.line_lambda = line_lambda.data();
        {   // the two [vif] conventions the reference's own sources do not settle: switchable, for `check`
            const std::string nodes = a.str("sed2flux_nodes", "merged"), prof = a.str("line_profile", "integral");
            if (nodes != "merged" && nodes != "sed" && nodes != "filter") { error("sed2flux_nodes must be merged, sed or filter"); return 1; }
            if (prof != "integral" && prof != "average") { error("line_profile must be integral or average"); return 1; }
            O.sed2flux_nodes = nodes == "sed" ? EGGPHOT_NODES_SED : (nodes == "filter" ? EGGPHOT_NODES_FILTER : EGGPHOT_NODES_MERGED);
            O.line_profile = prof == "average" ? EGGPHOT_LINE_AVERAGE : EGGPHOT_LINE_INTEGRAL;
        }

        // ---- compute: one context (and host thread) per GPU, contiguous shards, no exchange between them
        const size_t nband = fb.size(), nrf = fr.size();
        std::vector<float> flux(ngal * nband), flux_disk(ngal * nband), flux_bulge(ngal * nband);
        std::vector<float> rfmag(ngal * nrf), rfmag_disk(ngal * nrf), rfmag_bulge(ngal * nrf);
        std::vector<uint32_t> order(nband), rforder(nrf);
        std::vector<float> lambda(nband), rflambda(nrf);
        if (verbose) note("computing fluxes" + std::string(nrf ? " and absolute magnitudes" : "") + " on " + std::to_string(gpus) + " GPU(s)...");
        std::vector<Shard> shards(gpus);
        std::vector<eggphot_ctx *> ctxs(gpus, nullptr);
        auto galaxies = [&](size_t first) {
            eggphot_galaxies G{};
            auto off = [&](const std::vector<float> &v, size_t stride) { return v.empty() ? nullptr : v.data() + first * stride; };
            auto offu = [&](const std::vector<uint64_t> &v) { return v.empty() ? nullptr : v.data() + first; };
            G.z = off(z, 1); G.d = off(d, 1); G.m_disk = off(m_disk, 1); G.m_bulge = off(m_bulge, 1); G.m2lcor = off(m2lcor, 1);
            G.opt_sed_disk = offu(osd); G.opt_sed_bulge = offu(osb); G.ir_sed = offu(irs);
            G.lir = off(lir, 1); G.fpah = off(fpah, 1); G.mdust = off(mdust, 1); G.igm_abs = off(igm_abs, 3);
            G.vdisp = off(vdisp, 1); G.line_lum_disk = off(lld, line_lambda.size()); G.line_lum_bulge = off(llb, line_lambda.size());
            return G;
        };
        auto work = [&](int g) {
            Shard &s = shards[g];
            s.device = g;
            s.first = ngal * (size_t)g / gpus;
            s.count = ngal * (size_t)(g + 1) / gpus - s.first;
            eggphot_opts Og = O;
            Og.device = g;
            s.status = eggphot_create(&L, &B, nrf ? &R : nullptr, &Og, &ctxs[g]);
            if (s.status) { s.msg = ctxs[g] ? eggphot_last_error(ctxs[g]) : "allocation failed"; return; }
            const eggphot_galaxies G = galaxies(s.first);
            eggphot_outputs F{};
            F.flux = flux.data() + s.first * nband; F.flux_disk = flux_disk.data() + s.first * nband; F.flux_bulge = flux_bulge.data() + s.first * nband;
            if (nrf) { F.rfmag = rfmag.data() + s.first * nrf; F.rfmag_disk = rfmag_disk.data() + s.first * nrf; F.rfmag_bulge = rfmag_bulge.data() + s.first * nrf; }
            s.status = eggphot_compute(ctxs[g], &G, s.count, &F);
            if (s.status) s.msg = eggphot_last_error(ctxs[g]);
        };
        {
            std::vector<std::thread> th;
            for (int g = 1; g < gpus; ++g) th.emplace_back(work, g);
            work(0);
            for (auto &t : th) t.join();
        }
        for (const Shard &s : shards)
            if (s.status) {
                error("GPU " + std::to_string(s.device) + ": " + s.msg);
                for (auto *c : ctxs) eggphot_destroy(c);
                return 1;
            }
        eggphot_band_order(ctxs[0], 0, order.data(), lambda.data());
        if (nrf) eggphot_band_order(ctxs[0], 1, rforder.data(), rflambda.data());

        // ---- save_sed (src/egg-gencat.cpp:2059-2072, 2164-2236): all bulges, then all disks; lam then sed, f32
        if (save_sed) {
            const std::string seds_file = remove_extension(out_file) + "-seds.dat";
            if (verbose) note("will save spectra in '" + seds_file + "'");
            std::ofstream f(seds_file, std::ios::binary);
            if (!f) { error("could not write '" + seds_file + "'"); return 1; }
            std::vector<uint64_t> start[2] = {std::vector<uint64_t>(ngal, 0), std::vector<uint64_t>(ngal, 0)};
            std::vector<uint64_t> nbyte[2] = {std::vector<uint64_t>(ngal, 0), std::vector<uint64_t>(ngal, 0)};
            const eggphot_galaxies G = galaxies(0);
            uint64_t pos = 0;
            // Ranges of 4096 galaxies, two in flight: while the GPU builds range k + 1 and copies it out, this thread writes
            // range k, which arrives in the file's own layout (per galaxy lam then sed) in one of two pinned buffers.
            for (int comp = no_stellar ? 1 : 0; comp < 2; ++comp) { // bulge pass first (:2206-2212)
                const size_t n = eggphot_sed_size(ctxs[0], comp), chunk = 4096;
                std::vector<float> buf[2] = {std::vector<float>(chunk * 2 * n), std::vector<float>(chunk * 2 * n)};
                for (auto &b : buf) eggphot_host_register(b.data(), b.size() * sizeof(float));
                auto fail = [&]() { error(eggphot_last_error(ctxs[0])); for (auto &b : buf) eggphot_host_unregister(b.data()); return 1; };
                const size_t nrange = (ngal + chunk - 1) / chunk;
                for (size_t r = 0; r <= nrange; ++r) {
                    if (r < nrange) {
                        const size_t first = r * chunk, cnt = std::min(chunk, ngal - first);
                        if (eggphot_get_seds_packed_begin(ctxs[0], &G, first, cnt, comp, buf[r & 1].data())) return fail();
                    }
                    if (r == 0) continue;
                    // range r - 1 has landed once the range begun before the last is done (all of them after the last begin)
                    if (eggphot_get_seds_packed_wait(ctxs[0], r < nrange ? 0 : 1)) return fail();
                    const size_t first = (r - 1) * chunk, cnt = std::min(chunk, ngal - first);
                    f.write(reinterpret_cast<const char *>(buf[(r - 1) & 1].data()), (std::streamsize)(cnt * 2 * n * sizeof(float)));
                    for (size_t k = 0; k < cnt; ++k) {
                        start[comp][first + k] = pos;
                        nbyte[comp][first + k] = 2 * n * sizeof(float);
                        pos += 2 * n * sizeof(float);
                    }
                }
                for (auto &b : buf) eggphot_host_unregister(b.data());
            }
            fitscol::Table lk;
            std::vector<uint64_t> id(ngal);
            if (const fitscol::Column *c = cat.find("ID")) id = fitscol::as_u64(*c);
            else for (size_t i = 0; i < ngal; ++i) id[i] = i;
            lk.set(fitscol::make_u64("ID", id));
            lk.set(fitscol::make_u64("BULGE_START", start[0])); lk.set(fitscol::make_u64("BULGE_NBYTE", nbyte[0]));
            lk.set(fitscol::make_u64("DISK_START", start[1])); lk.set(fitscol::make_u64("DISK_NBYTE", nbyte[1]));
            lk.set(fitscol::make_u64("ELEM_SIZE", {sizeof(float)}));
            fitscol::write_table(remove_extension(seds_file) + "-lookup.fits", lk);
        }
        for (auto *c : ctxs) eggphot_destroy(c);

        // ---- check: an egg output catalog holds every input column of this stage next to its
        //      outputs, so any catalog of a real egg install is a test vector.  Compare what was just computed with the
        //      FLUX* / RFMAG* columns found in the input, band by band (matched by name), before they are replaced.
        int check_status = 0;
        if (a.flag("check")) {
            const double tol = a.num("check_tol", 1e-5);
            auto trim = [](std::string v) { while (!v.empty() && v.back() == ' ') v.pop_back(); return v; };
            auto compare = [&](const char *col, const char *names_col, const std::vector<float> &mine, size_t nmine,
                               const std::vector<std::string> &names_mine, bool is_mag) {
                const fitscol::Column *oc = cat.find(col), *nc = cat.find(names_col);
                if (!oc || !nc) { warning(std::string("check: the catalog has no ") + col + " / " + names_col + " columns"); check_status = std::max(check_status, 1); return; }
                const std::vector<float> old = fitscol::as_float(*oc);
                std::vector<std::string> onames = fitscol::as_strings(*nc);
                for (auto &v : onames) v = trim(v);
                const size_t nold = onames.size();
                if (old.size() != ngal * nold) { warning(std::string("check: ") + col + " is not [ngal, nband]"); check_status = std::max(check_status, 1); return; }
                for (size_t k = 0; k < nmine; ++k) {
                    const auto it = std::find(onames.begin(), onames.end(), names_mine[k]);
                    if (it == onames.end()) { note(std::string("check: band ") + names_mine[k] + " is not in the catalog's " + names_col + ": skipped"); continue; }
                    const size_t ko = (size_t)(it - onames.begin());
                    std::vector<double> rel;
                    size_t above = 0, zero_mismatch = 0;
                    for (size_t g = 0; g < ngal; ++g) {
                        const double x = mine[g * nmine + k], y = old[g * nold + ko];
                        double e;
                        if (is_mag) e = std::fabs(x - y);
                        else if (y == 0.0) { zero_mismatch += x != 0.0; continue; }
                        else e = std::fabs(x - y) / std::fabs(y);
                        rel.push_back(e);
                        above += e > tol;
                    }
                    double mx = 0, med = 0;
                    if (!rel.empty()) {
                        mx = *std::max_element(rel.begin(), rel.end());
                        std::nth_element(rel.begin(), rel.begin() + rel.size() / 2, rel.end());
                        med = rel[rel.size() / 2];
                    }
                    char line[256];
                    std::snprintf(line, sizeof line, "check: %-12s %-22s max %s %.3e  median %.3e  above %.0e: %zu of %zu%s", col, names_mine[k].c_str(),
                                  is_mag ? "abs" : "rel", mx, med, tol, above, rel.size(),
                                  zero_mismatch ? (", non-zero where the catalog has 0: " + std::to_string(zero_mismatch)).c_str() : "");
                    std::cout << line << std::endl;
                    if (above || zero_mismatch) check_status = 2;
                }
            };
            std::vector<std::string> bn, rn;
            for (uint32_t k : order) bn.push_back(fb[k].name);
            for (uint32_t k : rforder) rn.push_back(fr[k].name);
            if (nband) {
                compare("FLUX_DISK", "BANDS", flux_disk, nband, bn, false);
                compare("FLUX_BULGE", "BANDS", flux_bulge, nband, bn, false);
                compare("FLUX", "BANDS", flux, nband, bn, false);
            }
            if (nrf) {
                compare("RFMAG_DISK", "RFBANDS", rfmag_disk, nrf, rn, true);
                compare("RFMAG_BULGE", "RFBANDS", rfmag_bulge, nrf, rn, true);
                compare("RFMAG", "RFBANDS", rfmag, nrf, rn, true);
            }
            std::cout << "check: " << (check_status == 0 ? "every compared band is within tolerance"
                                       : check_status == 1 ? "nothing to compare against"
                                                           : "DIFFERENCES above tolerance (try sed2flux_nodes= and line_profile=)")
                      << " [sed2flux_nodes=" << a.str("sed2flux_nodes", "merged") << " line_profile=" << a.str("line_profile", "integral")
                      << " precision=" << precision << "]" << std::endl;
            if (!a.has("out")) return check_status; // nothing is written unless an output file is named
        }

        // ---- output (src/egg-gencat.cpp:2249-2286): pass the catalog through, (re)write the flux groups
        if (verbose) note("saving catalog...");
        std::string cmd;
        for (int i = 0; i < argc; ++i) cmd += std::string(i ? " " : "") + argv[i];
        cat.set(fitscol::make_string("CMD", cmd));
        const std::vector<long> d2 = {(long)ngal, (long)nband}, r2 = {(long)ngal, (long)nrf};
        if (nband) {
            std::vector<std::string> names;
            for (uint32_t k : order) names.push_back(fb[k].name);
            cat.set(fitscol::make_f32("FLUX", flux, d2)); cat.set(fitscol::make_f32("FLUX_DISK", flux_disk, d2));
            cat.set(fitscol::make_f32("FLUX_BULGE", flux_bulge, d2));
            cat.set(fitscol::make_strings("BANDS", names)); cat.set(fitscol::make_f32("LAMBDA", lambda));
        }
        if (nrf) {
            std::vector<std::string> names;
            for (uint32_t k : rforder) names.push_back(fr[k].name);
            cat.set(fitscol::make_f32("RFMAG", rfmag, r2)); cat.set(fitscol::make_f32("RFMAG_DISK", rfmag_disk, r2));
            cat.set(fitscol::make_f32("RFMAG_BULGE", rfmag_bulge, r2));
            cat.set(fitscol::make_strings("RFBANDS", names)); cat.set(fitscol::make_f32("RFLAMBDA", rflambda));
        }
        if (from_cat && !cat.find("CLUSTERED")) { // src/egg-gencat.cpp:1356: no clustering without generated positions
            fitscol::Column cc;
            cc.name = "CLUSTERED"; cc.type = 'L'; cc.dims = {(long)ngal}; cc.data.assign(ngal, 'F');
            cat.set(cc);
        }
        order_like_the_reference(cat);
        fitscol::write_table(out_file, cat);
        return check_status;
    } catch (const std::exception &e) {
        error(e.what());
        return 1;
    }
}
