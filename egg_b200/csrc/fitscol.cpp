// fitscol.cpp — see fitscol.hpp.
#include "fitscol.hpp"

#include <algorithm>
#include <cctype>
#include <cstring>
#include <fstream>
#include <sstream>

namespace fitscol {

namespace {

constexpr size_t kBlock = 2880;

size_t type_size(char t) {
    switch (t) {
    case 'L': case 'B': case 'A': return 1;
    case 'I': return 2;
    case 'J': case 'E': return 4;
    case 'K': case 'D': return 8;
    }
    throw std::runtime_error(std::string("unsupported FITS column type '") + t + "'");
}

std::string upper(std::string s) {
    for (char &c : s) c = (char)std::toupper((unsigned char)c);
    return s;
}

std::string trim(const std::string &s) {
    size_t a = s.find_first_not_of(' '), b = s.find_last_not_of(' ');
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

// parses one header unit; returns keyword -> raw value string (quotes removed for strings)
std::map<std::string, std::string> parse_header(const std::vector<char> &buf, size_t &off) {
    std::map<std::string, std::string> h;
    for (;;) {
        if (off + kBlock > buf.size()) throw std::runtime_error("truncated FITS header");
        bool done = false;
        for (size_t i = 0; i < kBlock; i += 80) {
            const std::string card(buf.data() + off + i, 80);
            const std::string key = trim(card.substr(0, 8));
            if (key == "END") { done = true; break; }
            if (card.compare(8, 2, "= ") != 0) continue;
            std::string val = card.substr(10);
            const size_t q = val.find('\'');
            if (q != std::string::npos && trim(val.substr(0, q)).empty()) {
                std::string s;
                size_t p = q + 1;
                while (p < val.size()) {
                    if (val[p] == '\'') {
                        if (p + 1 < val.size() && val[p + 1] == '\'') { s += '\''; p += 2; continue; }
                        break;
                    }
                    s += val[p++];
                }
                while (!s.empty() && s.back() == ' ') s.pop_back();
                h[key] = s;
            } else {
                const size_t sl = val.find('/');
                h[key] = trim(sl == std::string::npos ? val : val.substr(0, sl));
            }
        }
        off += kBlock;
        if (done) return h;
    }
}

template <typename T> void swap_copy(const char *src, uint8_t *dst, size_t n) {
    for (size_t i = 0; i < n; ++i)
        for (size_t b = 0; b < sizeof(T); ++b) dst[i * sizeof(T) + b] = (uint8_t)src[i * sizeof(T) + sizeof(T) - 1 - b];
}

void to_native(char type, const char *src, uint8_t *dst, size_t n) {
    switch (type_size(type)) {
    case 1: std::memcpy(dst, src, n); break;
    case 2: swap_copy<uint16_t>(src, dst, n); break;
    case 4: swap_copy<uint32_t>(src, dst, n); break;
    case 8: swap_copy<uint64_t>(src, dst, n); break;
    }
}

std::string card(const std::string &key, const std::string &value) {
    std::string c = key;
    c.resize(8, ' ');
    c += "= ";
    std::string v = value;
    if (v.size() < 20 && (v.empty() || v[0] != '\'')) v = std::string(20 - v.size(), ' ') + v;
    c += v;
    c.resize(80, ' ');
    return c;
}
std::string qstr(const std::string &s) {
    std::string o = "'";
    for (char ch : s) { o += ch; if (ch == '\'') o += '\''; }
    while (o.size() < 9) o += ' ';
    return o + "'";
}
void pad_block(std::string &s, char fill) { s.resize((s.size() + kBlock - 1) / kBlock * kBlock, fill); }

template <typename T, typename F> std::vector<T> convert(const Column &c, F f) {
    std::vector<T> out(c.count());
    for (size_t i = 0; i < out.size(); ++i) out[i] = f(i);
    return out;
}

template <typename T> T at(const Column &c, size_t i) {
    T v;
    std::memcpy(&v, c.data.data() + i * sizeof(T), sizeof(T));
    return v;
}

double get_num(const Column &c, size_t i) {
    switch (c.type) {
    case 'E': return at<float>(c, i);
    case 'D': return at<double>(c, i);
    case 'J': return at<int32_t>(c, i);
    case 'K': return (double)at<int64_t>(c, i);
    case 'I': return at<int16_t>(c, i);
    case 'B': return at<uint8_t>(c, i);
    case 'L': return c.data[i] == 'T';
    }
    throw std::runtime_error("column " + c.name + " is not numeric");
}

} // namespace

size_t Column::elem_size() const { return type_size(type); }
size_t Column::count() const { return data.size() / elem_size(); }

const Column *Table::find(const std::string &name) const {
    const std::string u = upper(name);
    for (const Column &c : cols) if (c.name == u) return &c;
    return nullptr;
}
Column *Table::find(const std::string &name) { return const_cast<Column *>(static_cast<const Table *>(this)->find(name)); }
void Table::set(Column c) {
    c.name = upper(c.name);
    if (Column *e = find(c.name)) *e = std::move(c);
    else cols.push_back(std::move(c));
}

Table read_table(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("could not open '" + path + "'");
    std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t off = 0;
    auto h = parse_header(buf, off);
    const int naxis = h.count("NAXIS") ? std::stoi(h["NAXIS"]) : 0;
    if (naxis > 0) {
        size_t n = (size_t)std::abs(std::stoi(h["BITPIX"])) / 8;
        for (int k = 1; k <= naxis; ++k) n *= (size_t)std::stol(h["NAXIS" + std::to_string(k)]);
        off += (n + kBlock - 1) / kBlock * kBlock;
    }
    h = parse_header(buf, off);
    if (trim(h["XTENSION"]) != "BINTABLE") throw std::runtime_error(path + ": first extension is not a BINTABLE");
    if (std::stol(h["NAXIS2"]) != 1) throw std::runtime_error(path + ": not a column-oriented table (NAXIS2 != 1)");
    const int nf = std::stoi(h["TFIELDS"]);
    Table t;
    size_t pos = off;
    for (int c = 1; c <= nf; ++c) {
        const std::string tform = trim(h["TFORM" + std::to_string(c)]);
        size_t p = 0;
        while (p < tform.size() && std::isdigit((unsigned char)tform[p])) ++p;
        const size_t count = p ? (size_t)std::stoull(tform.substr(0, p)) : 1;
        if (p >= tform.size()) throw std::runtime_error(path + ": bad TFORM '" + tform + "'");
        Column col;
        col.type = tform[p];
        col.name = upper(trim(h["TTYPE" + std::to_string(c)]));
        const size_t es = type_size(col.type);
        if (pos + count * es > buf.size()) throw std::runtime_error(path + ": truncated table data");
        col.data.resize(count * es);
        to_native(col.type, buf.data() + pos, col.data.data(), count);
        pos += count * es;
        const auto td = h.find("TDIM" + std::to_string(c));
        if (td != h.end()) {
            std::string s = td->second;
            s.erase(std::remove_if(s.begin(), s.end(), [](char ch) { return ch == '(' || ch == ')' || ch == ' '; }), s.end());
            std::stringstream ss(s);
            std::string item;
            while (std::getline(ss, item, ',')) if (!item.empty()) col.dims.push_back(std::stol(item));
            std::reverse(col.dims.begin(), col.dims.end());
        } else {
            col.dims = {(long)count};
        }
        t.cols.push_back(std::move(col));
    }
    return t;
}

void write_table(const std::string &path, const Table &t) {
    std::string prim = card("SIMPLE", "T") + card("BITPIX", "8") + card("NAXIS", "0") + card("EXTEND", "T") + "END";
    pad_block(prim, ' ');
    size_t nbytes = 0;
    for (const Column &c : t.cols) nbytes += c.data.size();
    std::string ext = card("XTENSION", qstr("BINTABLE")) + card("BITPIX", "8") + card("NAXIS", "2") +
                      card("NAXIS1", std::to_string(nbytes)) + card("NAXIS2", "1") + card("PCOUNT", "0") +
                      card("GCOUNT", "1") + card("TFIELDS", std::to_string(t.cols.size()));
    int k = 0;
    for (const Column &c : t.cols) {
        ++k;
        ext += card("TTYPE" + std::to_string(k), qstr(c.name));
        ext += card("TFORM" + std::to_string(k), qstr(std::to_string(c.count()) + c.type));
        if (c.dims.size() > 1 || c.type == 'A') {
            std::string d = "(";
            for (size_t i = c.dims.size(); i-- > 0;) d += std::to_string(c.dims[i]) + (i ? "," : "");
            ext += card("TDIM" + std::to_string(k), qstr(d + ")"));
        }
    }
    ext += "END";
    pad_block(ext, ' ');
    std::string body;
    body.reserve(nbytes + kBlock);
    for (const Column &c : t.cols) {
        const size_t es = c.elem_size(), n = c.count();
        const size_t o = body.size();
        body.resize(o + c.data.size());
        // big-endian on disk: the byte swap is its own inverse
        to_native(c.type, reinterpret_cast<const char *>(c.data.data()), reinterpret_cast<uint8_t *>(&body[o]), n);
        (void)es;
    }
    pad_block(body, '\0');
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("could not write '" + path + "'");
    f.write(prim.data(), (std::streamsize)prim.size());
    f.write(ext.data(), (std::streamsize)ext.size());
    f.write(body.data(), (std::streamsize)body.size());
}

std::vector<double> as_double(const Column &c) { return convert<double>(c, [&](size_t i) { return get_num(c, i); }); }
std::vector<float> as_float(const Column &c) { return convert<float>(c, [&](size_t i) { return (float)get_num(c, i); }); }
std::vector<uint64_t> as_u64(const Column &c) {
    return convert<uint64_t>(c, [&](size_t i) { return c.type == 'K' ? (uint64_t)at<int64_t>(c, i) : (uint64_t)get_num(c, i); });
}
std::vector<uint8_t> as_bool(const Column &c) {
    return convert<uint8_t>(c, [&](size_t i) { return (uint8_t)(c.type == 'L' ? (c.data[i] == 'T') : (get_num(c, i) != 0)); });
}
std::vector<std::string> as_strings(const Column &c) {
    if (c.type != 'A') throw std::runtime_error("column " + c.name + " is not a string column");
    const size_t w = c.dims.empty() ? c.data.size() : (size_t)c.dims.back();
    std::vector<std::string> out;
    for (size_t i = 0; w && i + w <= c.data.size(); i += w) {
        std::string s(reinterpret_cast<const char *>(c.data.data()) + i, w);
        while (!s.empty() && (s.back() == ' ' || s.back() == '\0')) s.pop_back();
        out.push_back(s);
    }
    return out;
}

template <typename T> static Column make_num(const std::string &name, char type, const std::vector<T> &v, std::vector<long> dims) {
    Column c;
    c.name = upper(name);
    c.type = type;
    c.dims = dims.empty() ? std::vector<long>{(long)v.size()} : dims;
    c.data.resize(v.size() * sizeof(T));
    if (!v.empty()) std::memcpy(c.data.data(), v.data(), c.data.size());
    return c;
}
Column make_f32(const std::string &n, const std::vector<float> &v, std::vector<long> d) { return make_num(n, 'E', v, std::move(d)); }
Column make_f64(const std::string &n, const std::vector<double> &v, std::vector<long> d) { return make_num(n, 'D', v, std::move(d)); }
Column make_u64(const std::string &n, const std::vector<uint64_t> &v, std::vector<long> d) { return make_num(n, 'K', v, std::move(d)); }
Column make_strings(const std::string &name, const std::vector<std::string> &v) {
    Column c;
    c.name = upper(name);
    c.type = 'A';
    size_t w = 1;
    for (const auto &s : v) w = std::max(w, s.size());
    c.dims = {(long)v.size(), (long)w};
    c.data.assign(v.size() * w, ' ');
    for (size_t i = 0; i < v.size(); ++i) std::memcpy(c.data.data() + i * w, v[i].data(), v[i].size());
    return c;
}
Column make_string(const std::string &name, const std::string &v) {
    Column c = make_strings(name, {v});
    c.dims = {(long)std::max<size_t>(v.size(), 1)};
    return c;
}

std::vector<std::pair<std::string, std::string>> read_filter_db(const std::string &db_file) {
    std::ifstream f(db_file);
    if (!f) throw std::runtime_error("could not find filter database file '" + db_file + "'");
    std::string dir = db_file;
    const size_t sl = dir.find_last_of('/');
    dir = sl == std::string::npos ? "" : dir.substr(0, sl + 1);
    std::vector<std::pair<std::string, std::string>> out;
    std::string line;
    while (std::getline(f, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#') continue;
        const size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        const std::string rel = trim(line.substr(eq + 1));
        out.emplace_back(trim(line.substr(0, eq)), (!rel.empty() && rel[0] == '/') ? rel : dir + rel);
    }
    return out;
}

} // namespace fitscol
