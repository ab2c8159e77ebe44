// fitscol.hpp — column-oriented FITS tables without cfitsio.
//
// Every table egg reads or writes is a primary HDU with NAXIS = 0 followed by one BINTABLE with a
// single row whose columns are whole vectors (TFORMn = '<count><type>', TDIMn = '(fastest,...,slowest)';
// doc/doc-egg-generic.tex:67-69).  This is the C++ side used by the egg-gencat-b200 front-end for the
// filter curves (src/egg-gencat.cpp:317-327), the SED libraries (:401-458), the input catalog and the
// output column groups (:2253-2286).
#ifndef EGGPHOT_FITSCOL_HPP
#define EGGPHOT_FITSCOL_HPP

#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace fitscol {

struct Column {
    std::string name;        // upper case
    char type = 'E';         // FITS letter: L B I J K E D A
    std::vector<long> dims;  // C order (slowest first); for 'A' the last dim is the string width
    std::vector<uint8_t> data; // native byte order
    size_t count() const;    // number of elements (bytes for 'A' and 'L')
    size_t elem_size() const;
};

struct Table {
    std::vector<Column> cols;
    const Column *find(const std::string &name) const; // case-insensitive
    Column *find(const std::string &name);
    void set(Column c); // replace or append
};

Table read_table(const std::string &path);
void write_table(const std::string &path, const Table &t);

// typed access with promotion (E/D/J/K/I/B -> requested)
std::vector<double> as_double(const Column &c);
std::vector<float> as_float(const Column &c);
std::vector<uint64_t> as_u64(const Column &c);
std::vector<uint8_t> as_bool(const Column &c);
std::vector<std::string> as_strings(const Column &c);

Column make_f32(const std::string &name, const std::vector<float> &v, std::vector<long> dims = {});
Column make_f64(const std::string &name, const std::vector<double> &v, std::vector<long> dims = {});
Column make_u64(const std::string &name, const std::vector<uint64_t> &v, std::vector<long> dims = {});
Column make_strings(const std::string &name, const std::vector<std::string> &v);
Column make_string(const std::string &name, const std::string &v);

// filter data base: "name=relative/path.fits" lines (src/egg-gencat.cpp:317)
std::vector<std::pair<std::string, std::string>> read_filter_db(const std::string &db_file);

} // namespace fitscol
#endif
