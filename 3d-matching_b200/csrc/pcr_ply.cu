// pcr_ply.cu — host-only part of the C ABI: native PLY reader / writer.
//
// Replaces o3d.io.read_point_cloud (src/ply/ply.py:80) and o3d.io.write_point_cloud (trim_ply.py:40,
// convert_stl-ply.py:8) for the vertex data the path needs: x y z (any scalar type), optional nx ny nz, optional
// red green blue on write.  ASCII (what the reference's own converter emits), binary_little_endian and
// binary_big_endian.  The file is mapped, the vertex records are decoded by a few host threads straight into the
// packed float4 layout the device kernels read (pinned memory if the caller passes it), so a cloud goes
// file -> one cudaMemcpy -> kernels without an intermediate (n,3) fp64 array or a pack kernel.
// Decimal text is converted to fp64 (correctly rounded, std::from_chars) and then rounded once to fp32 — the same
// two roundings the fp64 -> fp32 quantisation of rule D1 applies to an Open3D cloud.
//
// No CUDA in this file; it is a .cu only so that the one Makefile rule builds it.
#include "../../include/pcr.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct PlyError {
    int code;
    std::string msg;
};

[[noreturn]] void fail(int code, const std::string &m) { throw PlyError{code, m}; }

struct Mapped {
    const char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    explicit Mapped(const char *path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) fail(PCR_ERR_IO, std::string("cannot open ") + path);
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
            ::close(fd);
            fail(PCR_ERR_IO, std::string("not a regular file: ") + path);
        }
        n = (size_t)st.st_size;
        if (n) {
            void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) {
                ::close(fd);
                fail(PCR_ERR_IO, std::string("cannot map ") + path);
            }
            p = (const char *)m;
            madvise(m, n, MADV_SEQUENTIAL);
        }
    }
    ~Mapped() {
        if (p) munmap((void *)p, n);
        if (fd >= 0) ::close(fd);
    }
    Mapped(const Mapped &) = delete;
    Mapped &operator=(const Mapped &) = delete;
};

enum ScalarType { T_I8, T_U8, T_I16, T_U16, T_I32, T_U32, T_F32, T_F64, T_BAD };

ScalarType scalar_type(const std::string &s) {
    if (s == "char" || s == "int8") return T_I8;
    if (s == "uchar" || s == "uint8") return T_U8;
    if (s == "short" || s == "int16") return T_I16;
    if (s == "ushort" || s == "uint16") return T_U16;
    if (s == "int" || s == "int32") return T_I32;
    if (s == "uint" || s == "uint32") return T_U32;
    if (s == "float" || s == "float32") return T_F32;
    if (s == "double" || s == "float64") return T_F64;
    return T_BAD;
}

int type_size(ScalarType t) {
    switch (t) {
        case T_I8: case T_U8: return 1;
        case T_I16: case T_U16: return 2;
        case T_I32: case T_U32: case T_F32: return 4;
        case T_F64: return 8;
        default: return 0;
    }
}

struct Prop {
    std::string name;
    ScalarType type = T_BAD;
    bool is_list = false;
    ScalarType count_type = T_BAD;
    int offset = 0; // byte offset inside a binary record (scalar-only elements)
};

struct Element {
    std::string name;
    int64_t count = 0;
    std::vector<Prop> props;
    bool has_list = false;
    int stride = 0;
};

struct Header {
    int format = -1; // 0 ascii, 1 little, 2 big
    std::vector<Element> elements;
    int vertex = -1;
    size_t data_offset = 0;
    int ix[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1}; // x y z nx ny nz red green blue
};

std::vector<std::string> split_ws(const char *b, const char *e) {
    std::vector<std::string> out;
    while (b < e) {
        while (b < e && (*b == ' ' || *b == '\t' || *b == '\r')) ++b;
        const char *s = b;
        while (b < e && !(*b == ' ' || *b == '\t' || *b == '\r')) ++b;
        if (b > s) out.emplace_back(s, b);
    }
    return out;
}

int64_t parse_count(const std::string &s) {
    int64_t v = -1;
    auto r = std::from_chars(s.data(), s.data() + s.size(), v);
    if (r.ec != std::errc() || r.ptr != s.data() + s.size() || v < 0) fail(PCR_ERR_INVALID, "bad element count '" + s + "'");
    return v;
}

Header parse_header(const Mapped &f) {
    Header h;
    const char *p = f.p, *end = f.p + f.n;
    bool first = true, done = false;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        auto tok = split_ws(p, le);
        p = nl ? nl + 1 : end;
        if (first) {
            if (tok.size() != 1 || tok[0] != "ply") fail(PCR_ERR_INVALID, "not a PLY file (magic line missing)");
            first = false;
            continue;
        }
        if (tok.empty() || tok[0] == "comment" || tok[0] == "obj_info") continue;
        if (tok[0] == "format") {
            if (tok.size() < 2) fail(PCR_ERR_INVALID, "malformed format line");
            if (tok[1] == "ascii") h.format = 0;
            else if (tok[1] == "binary_little_endian") h.format = 1;
            else if (tok[1] == "binary_big_endian") h.format = 2;
            else fail(PCR_ERR_INVALID, "unsupported PLY format " + tok[1]);
        } else if (tok[0] == "element") {
            if (tok.size() < 3) fail(PCR_ERR_INVALID, "malformed element line");
            Element e;
            e.name = tok[1];
            e.count = parse_count(tok[2]);
            if (e.name == "vertex" && h.vertex < 0) h.vertex = (int)h.elements.size();
            h.elements.push_back(std::move(e));
        } else if (tok[0] == "property") {
            if (h.elements.empty()) fail(PCR_ERR_INVALID, "property before any element");
            Element &e = h.elements.back();
            Prop pr;
            if (tok.size() >= 5 && tok[1] == "list") {
                pr.is_list = true;
                pr.count_type = scalar_type(tok[2]);
                pr.type = scalar_type(tok[3]);
                pr.name = tok[4];
                if (pr.count_type == T_BAD || pr.count_type == T_F32 || pr.count_type == T_F64 || pr.type == T_BAD)
                    fail(PCR_ERR_INVALID, "bad list property type");
                e.has_list = true;
            } else if (tok.size() >= 3) {
                pr.type = scalar_type(tok[1]);
                pr.name = tok[2];
                if (pr.type == T_BAD) fail(PCR_ERR_INVALID, "unknown property type " + tok[1]);
                pr.offset = e.stride;
                e.stride += type_size(pr.type);
            } else {
                fail(PCR_ERR_INVALID, "malformed property line");
            }
            e.props.push_back(std::move(pr));
        } else if (tok[0] == "end_header") {
            done = true;
            break;
        } else {
            fail(PCR_ERR_INVALID, "unexpected header keyword " + tok[0]);
        }
    }
    if (first) fail(PCR_ERR_INVALID, "not a PLY file (empty)");
    if (!done) fail(PCR_ERR_INVALID, "unexpected end of PLY header");
    if (h.format < 0) fail(PCR_ERR_INVALID, "PLY header has no format line");
    h.data_offset = (size_t)(p - f.p);
    if (h.vertex < 0) {
        // a PLY without a vertex element is an empty cloud (Open3D: "Read PLY failed: number of vertex <= 0")
        Element e;
        e.name = "vertex";
        h.vertex = (int)h.elements.size();
        h.elements.push_back(e);
        return h;
    }
    const Element &v = h.elements[h.vertex];
    if (v.has_list) fail(PCR_ERR_INVALID, "list properties on vertices are not supported");
    static const char *names[9] = {"x", "y", "z", "nx", "ny", "nz", "red", "green", "blue"};
    for (size_t i = 0; i < v.props.size(); ++i)
        for (int k = 0; k < 9; ++k)
            if (h.ix[k] < 0 && v.props[i].name == names[k]) h.ix[k] = (int)i;
    if (v.count > 0 && (h.ix[0] < 0 || h.ix[1] < 0 || h.ix[2] < 0)) fail(PCR_ERR_INVALID, "PLY vertex element lacks x/y/z");
    return h;
}

template <typename T>
inline T load_raw(const char *p, bool swap) {
    unsigned char b[sizeof(T)];
    memcpy(b, p, sizeof(T));
    if (swap) std::reverse(b, b + sizeof(T));
    T v;
    memcpy(&v, b, sizeof(T));
    return v;
}

inline double load_scalar(const char *p, ScalarType t, bool swap) {
    switch (t) {
        case T_I8: return (double)*(const signed char *)p;
        case T_U8: return (double)*(const unsigned char *)p;
        case T_I16: return (double)load_raw<int16_t>(p, swap);
        case T_U16: return (double)load_raw<uint16_t>(p, swap);
        case T_I32: return (double)load_raw<int32_t>(p, swap);
        case T_U32: return (double)load_raw<uint32_t>(p, swap);
        case T_F32: return (double)load_raw<float>(p, swap);
        case T_F64: return load_raw<double>(p, swap);
        default: return 0.0;
    }
}

inline int64_t load_int(const char *p, ScalarType t, bool swap) { return (int64_t)load_scalar(p, t, swap); }

struct Sinks {
    float *xyzw;
    float *nrm;
    double *xyz64;
    bool want_nrm;
};

inline void store_vertex(const Sinks &s, int64_t i, const double *v /* x y z nx ny nz */) {
    if (s.xyzw) {
        float *o = s.xyzw + 4 * i;
        o[0] = (float)v[0]; o[1] = (float)v[1]; o[2] = (float)v[2]; o[3] = 0.0f;
    }
    if (s.xyz64) {
        double *o = s.xyz64 + 3 * i;
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    }
    if (s.want_nrm) {
        float *o = s.nrm + 4 * i;
        o[0] = (float)v[3]; o[1] = (float)v[4]; o[2] = (float)v[5]; o[3] = 0.0f;
    }
}

int pick_threads(int requested, int64_t n) {
    if (n < 32768) return 1;
    int hw = (int)std::thread::hardware_concurrency();
    if (hw <= 0) hw = 4;
    int t = requested > 0 ? requested : std::min(hw, 16);
    return (int)std::max<int64_t>(1, std::min<int64_t>(t, n / 16384));
}

template <typename F>
void run_parallel(int threads, F &&body) {
    if (threads <= 1) {
        body(0);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(threads - 1);
    for (int t = 1; t < threads; ++t) pool.emplace_back([&body, t] { body(t); });
    body(0);
    for (auto &th : pool) th.join();
}

// offset of the vertex element's first byte for binary files (elements in front of it are skipped)
size_t binary_vertex_offset(const Mapped &f, const Header &h) {
    const bool swap = h.format == 2;
    size_t off = h.data_offset;
    for (int e = 0; e < h.vertex; ++e) {
        const Element &el = h.elements[e];
        if (!el.has_list) {
            if (off > f.n || (el.stride > 0 && (size_t)el.count > (f.n - off) / (size_t)el.stride))
                fail(PCR_ERR_INVALID, "PLY file is truncated");
            off += (size_t)el.stride * (size_t)el.count;
            continue;
        }
        for (int64_t i = 0; i < el.count; ++i)
            for (const Prop &pr : el.props) {
                if (!pr.is_list) {
                    off += type_size(pr.type);
                    continue;
                }
                if (off + type_size(pr.count_type) > f.n) fail(PCR_ERR_INVALID, "PLY file is truncated");
                int64_t c = load_int(f.p + off, pr.count_type, swap);
                if (c < 0) fail(PCR_ERR_INVALID, "negative list length");
                off += type_size(pr.count_type) + (size_t)c * type_size(pr.type);
            }
    }
    return off;
}

void read_binary(const Mapped &f, const Header &h, const Sinks &s, int threads) {
    const Element &v = h.elements[h.vertex];
    const bool swap = h.format == 2; // host is little-endian (x86-64 / aarch64 Linux)
    const size_t off = binary_vertex_offset(f, h);
    const size_t stride = (size_t)v.stride;
    if (off > f.n || (f.n - off) / std::max<size_t>(stride, 1) < (size_t)v.count) fail(PCR_ERR_INVALID, "PLY file is truncated");
    int po[6];
    ScalarType pt[6];
    for (int k = 0; k < 6; ++k) {
        po[k] = h.ix[k] >= 0 ? v.props[h.ix[k]].offset : -1;
        pt[k] = h.ix[k] >= 0 ? v.props[h.ix[k]].type : T_BAD;
    }
    const int nk = s.want_nrm ? 6 : 3;
    const char *base = f.p + off;
    const int64_t n = v.count;
    const int T = pick_threads(threads, n);
    run_parallel(T, [&](int t) {
        int64_t b = n * t / T, e = n * (t + 1) / T;
        double val[6] = {0, 0, 0, 0, 0, 0};
        for (int64_t i = b; i < e; ++i) {
            const char *r = base + (size_t)i * stride;
            for (int k = 0; k < nk; ++k) val[k] = load_scalar(r + po[k], pt[k], swap);
            store_vertex(s, i, val);
        }
    });
}

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// Exact fast path (Clinger 1990): a decimal with at most 15 significant-or-leading digits and |exponent| <= 22 is
// m * 10^k or m / 10^k with m < 2^53 and 10^k exactly representable, i.e. ONE correctly rounded IEEE operation —
// the same double std::from_chars / strtod return.  Anything else (long mantissas, big exponents, inf, nan)
// returns nullptr and takes the general routine.  Three to four times faster than from_chars on "%.9g" text.
inline const char *parse_number_fast(const char *p, const char *e, double &out) {
    static const double P10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char *q = p;
    bool neg = false;
    if (q < e && *q == '-') {
        neg = true;
        ++q;
    }
    uint64_t m = 0;
    int digits = 0, frac = 0;
    while (q < e && (unsigned)(*q - '0') < 10u) {
        m = m * 10 + (unsigned)(*q - '0');
        ++digits;
        ++q;
    }
    if (q < e && *q == '.') {
        ++q;
        while (q < e && (unsigned)(*q - '0') < 10u) {
            m = m * 10 + (unsigned)(*q - '0');
            ++digits;
            ++frac;
            ++q;
        }
    }
    if (digits == 0 || digits > 15) return nullptr;  // 15 digits: m < 10^15 < 2^53, no overflow above
    int ex = 0;
    if (q < e && (*q == 'e' || *q == 'E')) {
        const char *r = q + 1;
        bool eneg = false;
        if (r < e && (*r == '-' || *r == '+')) {
            eneg = *r == '-';
            ++r;
        }
        int ed = 0;
        while (r < e && (unsigned)(*r - '0') < 10u && ed < 4) {
            ex = ex * 10 + (*r - '0');
            ++ed;
            ++r;
        }
        if (ed == 0 || (r < e && (unsigned)(*r - '0') < 10u)) return nullptr;
        if (eneg) ex = -ex;
        q = r;
    }
    const int k = ex - frac;
    if (k < -22 || k > 22) return nullptr;
    const double v = k >= 0 ? (double)m * P10[k] : (double)m / P10[-k];
    out = neg ? -v : v;
    return q;
}

// one number; accepts what strtod accepts in PLY files: optional sign, decimal / exponent forms, inf, nan
inline const char *parse_number(const char *p, const char *e, double &out) {
    const char *q = p;
    if (q < e && *q == '+') ++q;
    if (const char *f = parse_number_fast(q, e, out)) return f;
    auto r = std::from_chars(q, e, out);
    if (r.ec == std::errc::result_out_of_range) {
        // from_chars leaves `out` untouched: take strtod's answer (±HUGE_VAL or a denormal / 0)
        std::string tmp(q, r.ptr);
        out = strtod(tmp.c_str(), nullptr);
        return r.ptr;
    }
    if (r.ec != std::errc()) return nullptr;
    return r.ptr;
}

// token-stream parse (newlines are plain white space, as in rply): vertex records [first, first + count)
// starting at p; returns the position after the last token
const char *parse_ascii_stream(const char *p, const char *e, int nprops, const int *slot_of_prop, const Sinks &s,
                               int64_t first, int64_t count) {
    for (int64_t i = first; i < first + count; ++i) {
        double val[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < nprops; ++k) {
            while (p < e && is_ws(*p)) ++p;
            if (p >= e) fail(PCR_ERR_INVALID, "PLY file is truncated (vertex " + std::to_string(i) + ")");
            double x;
            const char *q = parse_number(p, e, x);
            if (!q || (q < e && !is_ws(*q))) fail(PCR_ERR_INVALID, "malformed number in vertex " + std::to_string(i));
            if (slot_of_prop[k] >= 0) val[slot_of_prop[k]] = x;
            p = q;
        }
        store_vertex(s, i, val);
    }
    return p;
}

// one vertex per line, exactly nprops tokens: returns false when the line does not have that shape
inline bool parse_ascii_line(const char *p, const char *le, int nprops, const int *slot_of_prop, double *val) {
    for (int k = 0; k < nprops; ++k) {
        while (p < le && (*p == ' ' || *p == '\t')) ++p;
        if (p >= le) return false;
        double x;
        const char *q = parse_number(p, le, x);
        if (!q || (q < le && !(*q == ' ' || *q == '\t' || *q == '\r'))) return false;
        if (slot_of_prop[k] >= 0) val[slot_of_prop[k]] = x;
        p = q;
    }
    while (p < le && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
    return p == le;
}

void read_ascii(const Mapped &f, const Header &h, const Sinks &s, int threads) {
    const Element &v = h.elements[h.vertex];
    const int nprops = (int)v.props.size();
    std::vector<int> slot(nprops, -1);
    for (int k = 0; k < (s.want_nrm ? 6 : 3); ++k) slot[h.ix[k]] = k;
    const char *p = f.p + h.data_offset, *end = f.p + f.n;
    // elements in front of the vertex element: skip their tokens (scalar-only) or their lines (lists)
    for (int e = 0; e < h.vertex; ++e) {
        const Element &el = h.elements[e];
        for (int64_t i = 0; i < el.count; ++i) {
            while (p < end && is_ws(*p)) ++p;
            const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
            if (!nl && p >= end) fail(PCR_ERR_INVALID, "PLY file is truncated");
            p = nl ? nl + 1 : end;
        }
    }
    const int64_t n = v.count;
    const int T = pick_threads(threads, n);
    if (T > 1) {
        // line-parallel fast path: chunk the body at byte boundaries, count the newlines per chunk, and let every
        // thread decode the complete lines that START in its chunk.  Any line that is not "nprops numbers" makes the
        // whole read fall back to the sequential token stream (blank lines, several vertices per line, ...).
        const size_t body = (size_t)(end - p);
        std::vector<int64_t> lines(T + 1, 0);
        std::vector<size_t> cut(T + 1);
        for (int t = 0; t <= T; ++t) cut[t] = body * (size_t)t / (size_t)T;
        run_parallel(T, [&](int t) {
            int64_t c = 0;
            const char *b = p + cut[t], *e = p + cut[t + 1];
            while (b < e) {
                const char *nl = (const char *)memchr(b, '\n', (size_t)(e - b));
                if (!nl) break;
                ++c;
                b = nl + 1;
            }
            lines[t + 1] = c;
        });
        for (int t = 0; t < T; ++t) lines[t + 1] += lines[t];
        std::atomic<bool> ok{true};
        std::atomic<int64_t> decoded{0};
        run_parallel(T, [&](int t) {
            // first line starting in this chunk = the byte after the first newline at or after cut[t]-1
            const char *b = p + cut[t];
            int64_t line = lines[t];
            if (t > 0) {
                if (b[-1] != '\n') {
                    const char *nl = (const char *)memchr(b, '\n', (size_t)(end - b));
                    if (!nl) return;
                    b = nl + 1;
                    line += 1; // the newline that ends the straddling line lies in this chunk
                }
            }
            const char *lim = p + cut[t + 1];
            int64_t mine = 0;
            while (b < lim && line < n && ok.load(std::memory_order_relaxed)) {
                const char *nl = (const char *)memchr(b, '\n', (size_t)(end - b));
                const char *le = nl ? nl : end;
                double val[6] = {0, 0, 0, 0, 0, 0};
                if (!parse_ascii_line(b, le, nprops, slot.data(), val)) {
                    ok.store(false);
                    return;
                }
                store_vertex(s, line, val);
                ++line;
                ++mine;
                b = nl ? nl + 1 : end;
            }
            decoded.fetch_add(mine);
        });
        if (ok.load() && decoded.load() == n) return;
    }
    parse_ascii_stream(p, end, nprops, slot.data(), s, 0, n);
}

void fill_info(const Header &h, pcr_ply_info *info) {
    memset(info, 0, sizeof(*info));
    const Element &v = h.elements[h.vertex];
    info->n_vertex = v.count;
    info->data_offset = (int64_t)h.data_offset;
    info->format = h.format;
    info->has_normals = (h.ix[3] >= 0 && h.ix[4] >= 0 && h.ix[5] >= 0) ? 1 : 0;
    info->has_colors = (h.ix[6] >= 0 && h.ix[7] >= 0 && h.ix[8] >= 0) ? 1 : 0;
    info->n_props = (int32_t)v.props.size();
    info->vertex_stride = h.format == 0 ? 0 : v.stride;
}

int report(const PlyError &e, char *err, int cap) {
    if (err && cap > 0) snprintf(err, (size_t)cap, "%s", e.msg.c_str());
    return e.code;
}

template <typename F>
int guarded(char *err, int cap, F &&f) {
    if (err && cap > 0) err[0] = 0;
    try {
        f();
        return PCR_OK;
    } catch (const PlyError &e) {
        return report(e, err, cap);
    } catch (const std::bad_alloc &) {
        return report(PlyError{PCR_ERR_OOM, "out of host memory"}, err, cap);
    } catch (const std::exception &e) {
        return report(PlyError{PCR_ERR_IO, e.what()}, err, cap);
    } catch (...) {
        return report(PlyError{PCR_ERR_IO, "unknown failure"}, err, cap);
    }
}

struct File {
    FILE *f;
    explicit File(const char *path) : f(fopen(path, "wb")) {
        if (!f) fail(PCR_ERR_IO, std::string("cannot create ") + path);
    }
    ~File() {
        if (f) fclose(f);
    }
    void write(const void *p, size_t n) {
        if (n && fwrite(p, 1, n, f) != n) fail(PCR_ERR_IO, "short write");
    }
    void close() {
        FILE *g = f;
        f = nullptr;
        if (fclose(g) != 0) fail(PCR_ERR_IO, "close failed");
    }
};

inline char *put_float(char *p, char *e, float v) {
    if (std::isnan(v)) {
        memcpy(p, "nan", 3);
        return p + 3;
    }
    if (std::isinf(v)) {
        const char *s = v < 0 ? "-inf" : "inf";
        size_t n = strlen(s);
        memcpy(p, s, n);
        return p + n;
    }
    auto r = std::to_chars(p, e, v); // shortest text that reads back to the same fp32
    return r.ptr;
}

} // namespace

extern "C" {

int pcr_ply_probe(const char *path, pcr_ply_info *info, char *err, int err_cap) {
    return guarded(err, err_cap, [&] {
        if (!path || !info) fail(PCR_ERR_INVALID, "null argument");
        Mapped f(path);
        Header h = parse_header(f);
        // A header may claim any vertex count: callers size (and pin) their buffers from it, so a count the file cannot
        // hold is rejected here, before anything is allocated.  Binary: stride bytes per vertex; ASCII: at
        // least two bytes ("0\n") per property of a vertex line would be generous — one byte per vertex is the safe floor.
        const Element &v = h.elements[h.vertex];
        const size_t avail = f.n > h.data_offset ? f.n - h.data_offset : 0;
        const double need = h.format == 0 ? (double)v.count : (double)v.count * (double)v.stride;
        if (need > (double)avail)
            fail(PCR_ERR_INVALID, "header declares " + std::to_string(v.count) + " vertices but only " + std::to_string(avail) +
                                      " bytes follow it (truncated file)");
        fill_info(h, info);
    });
}

int pcr_ply_read(const char *path, int64_t n_cap, float *xyzw_host, float *normals_xyzw_host, double *xyz64_host,
                 int threads, pcr_ply_info *info, char *err, int err_cap) {
    return guarded(err, err_cap, [&] {
        if (!path) fail(PCR_ERR_INVALID, "null path");
        Mapped f(path);
        Header h = parse_header(f);
        pcr_ply_info local;
        fill_info(h, &local);
        if (info) *info = local;
        if (local.n_vertex > n_cap) fail(PCR_ERR_INVALID, "buffer holds " + std::to_string(n_cap) + " vertices, file has " + std::to_string(local.n_vertex));
        if (local.n_vertex == 0) return;
        Sinks s{xyzw_host, normals_xyzw_host, xyz64_host, normals_xyzw_host != nullptr && local.has_normals != 0};
        if (h.format == 0) read_ascii(f, h, s, threads);
        else read_binary(f, h, s, threads);
    });
}

int pcr_ply_write(const char *path, const float *xyzw_host, int64_t n, const float *normals_xyzw_host,
                  const unsigned char *rgb_host, int binary, char *err, int err_cap) {
    return guarded(err, err_cap, [&] {
        if (!path || n < 0 || (n > 0 && !xyzw_host)) fail(PCR_ERR_INVALID, "bad argument");
        File out(path);
        std::string hdr = "ply\n";
        hdr += binary ? "format binary_little_endian 1.0\n" : "format ascii 1.0\n";
        hdr += "comment written by libpcr_b200\n";
        hdr += "element vertex " + std::to_string(n) + "\n";
        hdr += "property float x\nproperty float y\nproperty float z\n";
        if (normals_xyzw_host) hdr += "property float nx\nproperty float ny\nproperty float nz\n";
        if (rgb_host) hdr += "property uchar red\nproperty uchar green\nproperty uchar blue\n";
        hdr += "end_header\n";
        out.write(hdr.data(), hdr.size());
        const int64_t chunk = 1 << 16;
        std::vector<char> buf;
        const size_t rec_bin = 12 + (normals_xyzw_host ? 12 : 0) + (rgb_host ? 3 : 0);
        buf.resize((size_t)chunk * (binary ? rec_bin : (size_t)(6 * 20 + 3 * 4 + 2)));
        for (int64_t b = 0; b < n; b += chunk) {
            const int64_t e = std::min(n, b + chunk);
            char *p = buf.data(), *lim = buf.data() + buf.size();
            for (int64_t i = b; i < e; ++i) {
                if (binary) {
                    memcpy(p, xyzw_host + 4 * i, 12);
                    p += 12;
                    if (normals_xyzw_host) {
                        memcpy(p, normals_xyzw_host + 4 * i, 12);
                        p += 12;
                    }
                    if (rgb_host) {
                        memcpy(p, rgb_host + 3 * i, 3);
                        p += 3;
                    }
                } else {
                    for (int k = 0; k < 3; ++k) {
                        if (k) *p++ = ' ';
                        p = put_float(p, lim, xyzw_host[4 * i + k]);
                    }
                    if (normals_xyzw_host)
                        for (int k = 0; k < 3; ++k) {
                            *p++ = ' ';
                            p = put_float(p, lim, normals_xyzw_host[4 * i + k]);
                        }
                    if (rgb_host)
                        for (int k = 0; k < 3; ++k) {
                            *p++ = ' ';
                            auto r = std::to_chars(p, lim, (unsigned)rgb_host[3 * i + k]);
                            p = r.ptr;
                        }
                    *p++ = '\n';
                }
            }
            out.write(buf.data(), (size_t)(p - buf.data()));
        }
        out.close();
    });
}

} // extern "C"
