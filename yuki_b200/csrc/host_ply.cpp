// PLY mesh loader for the host side of the B200 backend: the step before the hot path for config 3
// ("BVH-intersections debug integrator on a ~1M-triangle PLY mesh"). Restates what yuki/src/scene/ply.rs:19-156,
// 242-284 takes from a file — through the `ply-rs` parser (git dependency, yuki/Cargo.toml:28, not vendored) —
// with the same acceptance rules:
//   * element `vertex` must have x, y, z; nx/ny/nz and u/v are optional; only `float` properties are consumed
//     (ply.rs:253-272 matches Property::Float), anything else is skipped;
//   * element `face` must have `vertex_index` or `vertex_indices` as a list of int / uint items (ply.rs:287-301);
//   * faces are fan-triangulated (v0, v_i, v_i+1) (ply.rs:81-92).
// ASCII, binary_little_endian and binary_big_endian payloads are read; elements other than vertex/face are skipped in
// file order. The fit-to-unit transform and the Scene::ply defaults live above the ABI (yuki_b200/scenes.py::ply).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "yuki_gpu.h"
#include "yk_guard.h"

int yk_set_error(int code, const std::string& msg);  // host_scene.cpp

struct yk_ply {
    std::vector<float> points, normals, uvs;
    std::vector<uint32_t> indices;
};

namespace {

enum class PlyType { I8, U8, I16, U16, I32, U32, F32, F64, Unknown };

PlyType parse_type(const std::string& s) {
    if (s == "char" || s == "int8") return PlyType::I8;
    if (s == "uchar" || s == "uint8") return PlyType::U8;
    if (s == "short" || s == "int16") return PlyType::I16;
    if (s == "ushort" || s == "uint16") return PlyType::U16;
    if (s == "int" || s == "int32") return PlyType::I32;
    if (s == "uint" || s == "uint32") return PlyType::U32;
    if (s == "float" || s == "float32") return PlyType::F32;
    if (s == "double" || s == "float64") return PlyType::F64;
    return PlyType::Unknown;
}
size_t type_size(PlyType t) {
    switch (t) {
        case PlyType::I8: case PlyType::U8: return 1;
        case PlyType::I16: case PlyType::U16: return 2;
        case PlyType::I32: case PlyType::U32: case PlyType::F32: return 4;
        case PlyType::F64: return 8;
        default: return 0;
    }
}

struct Property {
    std::string name;
    bool is_list = false;
    PlyType type = PlyType::Unknown;        // scalar type, or list item type
    PlyType count_type = PlyType::Unknown;  // lists only
};
struct Element {
    std::string name;
    uint64_t count = 0;
    std::vector<Property> props;
};
enum class Format { Ascii, LittleEndian, BigEndian };

// One scalar as a double (exact for every PLY type except 64-bit ints, which PLY does not have).
struct Reader {
    std::istream& in;
    Format fmt;
    bool ok = true;
    double scalar(PlyType t) {
        if (fmt == Format::Ascii) {
            double v = 0;
            if (!(in >> v)) ok = false;
            return v;
        }
        unsigned char b[8] = {0};
        const size_t n = type_size(t);
        if (!in.read(reinterpret_cast<char*>(b), (std::streamsize)n)) { ok = false; return 0; }
        if (fmt == Format::BigEndian)
            for (size_t i = 0; i < n / 2; ++i) std::swap(b[i], b[n - 1 - i]);
        switch (t) {
            case PlyType::I8: { int8_t v; std::memcpy(&v, b, 1); return v; }
            case PlyType::U8: return b[0];
            case PlyType::I16: { int16_t v; std::memcpy(&v, b, 2); return v; }
            case PlyType::U16: { uint16_t v; std::memcpy(&v, b, 2); return v; }
            case PlyType::I32: { int32_t v; std::memcpy(&v, b, 4); return v; }
            case PlyType::U32: { uint32_t v; std::memcpy(&v, b, 4); return v; }
            case PlyType::F32: { float v; std::memcpy(&v, b, 4); return v; }
            case PlyType::F64: { double v; std::memcpy(&v, b, 8); return v; }
            default: ok = false; return 0;
        }
    }
    // f32 payloads must keep their bits (a double round trip does, ASCII goes through strtof-equivalent rounding)
    float f32(PlyType t) {
        if (fmt == Format::Ascii) {
            std::string tok;
            if (!(in >> tok)) { ok = false; return 0.0f; }
            return std::strtof(tok.c_str(), nullptr);
        }
        return (float)scalar(t);
    }
};

}  // namespace

extern "C" {

int yk_ply_load(const char* path, yk_ply** out) {
    return yk_guard("yk_ply_load", [&]() -> int {
    if (!path || !out) return yk_set_error(YK_ERR_INVALID, "yk_ply_load: null argument");
    std::ifstream in(path, std::ios::binary);
    if (!in) return yk_set_error(YK_ERR_INVALID, std::string("Could not open '") + path + "'");  // ply.rs:27-30
    std::string line;
    if (!std::getline(in, line) || line.substr(0, 3) != "ply") return yk_set_error(YK_ERR_INVALID, "PLY: missing magic");
    Format fmt = Format::Ascii;
    bool have_format = false, ended = false;
    std::vector<Element> elements;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string kw;
        ls >> kw;
        if (kw == "format") {
            std::string f;
            ls >> f;
            if (f == "ascii") fmt = Format::Ascii;
            else if (f == "binary_little_endian") fmt = Format::LittleEndian;
            else if (f == "binary_big_endian") fmt = Format::BigEndian;
            else return yk_set_error(YK_ERR_INVALID, "PLY: unknown format '" + f + "'");
            have_format = true;
        } else if (kw == "element") {
            Element e;
            ls >> e.name >> e.count;
            elements.push_back(e);
        } else if (kw == "property") {
            if (elements.empty()) return yk_set_error(YK_ERR_INVALID, "PLY: property before any element");
            Property p;
            std::string t;
            ls >> t;
            if (t == "list") {
                std::string ct, it;
                ls >> ct >> it >> p.name;
                p.is_list = true;
                p.count_type = parse_type(ct);
                p.type = parse_type(it);
                if (p.count_type == PlyType::Unknown) return yk_set_error(YK_ERR_INVALID, "PLY: unknown list count type");
            } else {
                p.type = parse_type(t);
                ls >> p.name;
            }
            if (p.type == PlyType::Unknown) return yk_set_error(YK_ERR_INVALID, "PLY: unknown property type in '" + line + "'");
            elements.back().props.push_back(p);
        } else if (kw == "end_header") {
            ended = true;
            break;
        }  // comment / obj_info / blank: ignored
    }
    if (!ended || !have_format) return yk_set_error(YK_ERR_INVALID, "PLY: truncated header");

    // is_valid, ply.rs:169-240
    const Element* vert = nullptr;
    const Element* face = nullptr;
    for (const Element& e : elements) {
        if (e.name == "vertex") vert = &e;
        else if (e.name == "face") face = &e;
    }
    auto has = [](const Element* e, const char* name) {
        for (const Property& p : e->props)
            if (p.name == name) return true;
        return false;
    };
    if (!vert) return yk_set_error(YK_ERR_INVALID, "PLY: Missing element 'vertex'");
    for (const char* need : {"x", "y", "z"})
        if (!has(vert, need)) return yk_set_error(YK_ERR_INVALID, std::string("PLY: Element 'vertex' missing property '") + need + "'");
    if (!face) return yk_set_error(YK_ERR_INVALID, "PLY: Missing element 'face'");
    if (!has(face, "vertex_index") && !has(face, "vertex_indices"))
        return yk_set_error(YK_ERR_INVALID, "PLY: Elemnent 'face' should have either 'vertex_index' or 'vertex_indices'");
    // The reference's Vertex::set_property unwraps the normal / uv it creates on `nx` / `u` (ply.rs:262-271): a file that
    // lists ny, nz or v first (or without nx / u) panics there. Same inputs are rejected here.
    {
        int i_nx = -1, i_ny = -1, i_nz = -1, i_u = -1, i_v = -1;
        for (size_t i = 0; i < vert->props.size(); ++i) {
            const Property& p = vert->props[i];
            if (p.is_list || p.type != PlyType::F32) continue;
            if (p.name == "nx") i_nx = (int)i;
            if (p.name == "ny") i_ny = (int)i;
            if (p.name == "nz") i_nz = (int)i;
            if (p.name == "u") i_u = (int)i;
            if (p.name == "v") i_v = (int)i;
        }
        if ((i_ny >= 0 && (i_nx < 0 || i_ny < i_nx)) || (i_nz >= 0 && (i_nx < 0 || i_nz < i_nx)) || (i_v >= 0 && (i_u < 0 || i_v < i_u)))
            return yk_set_error(YK_ERR_INVALID, "PLY: ny/nz before nx or v before u (the reference panics on this layout)");
    }

    // Element counts come straight from the header: bound what is reserved up front by what the payload can hold (a row
    // takes at least one byte per property in every format), so that a hostile `element vertex 9e18` is a parse error
    // ("truncated payload") and not a length_error / bad_alloc.
    uint64_t payload_bytes = 0;
    {
        const std::streampos here = in.tellg();
        in.seekg(0, std::ios::end);
        const std::streampos end = in.tellg();
        in.seekg(here);
        if (here >= 0 && end >= here) payload_bytes = (uint64_t)(end - here);
    }
    auto ply = std::make_unique<yk_ply>();
    Reader rd{in, fmt};
    for (const Element& e : elements) {
        const bool is_vert = &e == vert, is_face = &e == face;
        bool e_has_normal = false, e_has_uv = false;
        if (is_vert) {
            for (const Property& p : e.props) {
                if (!p.is_list && p.type == PlyType::F32 && p.name == "nx") e_has_normal = true;
                if (!p.is_list && p.type == PlyType::F32 && p.name == "u") e_has_uv = true;
            }
            const uint64_t rows = std::min<uint64_t>(e.count, payload_bytes / std::max<size_t>(e.props.size(), 1));
            ply->points.reserve((size_t)rows * 3);
            if (e_has_normal) ply->normals.reserve((size_t)rows * 3);
            if (e_has_uv) ply->uvs.reserve((size_t)rows * 2);
        }
        std::vector<uint32_t> poly;
        for (uint64_t row = 0; row < e.count; ++row) {
            float pt[3] = {0, 0, 0}, nn[3] = {0, 0, 0}, uv[2] = {0, 0};
            poly.clear();
            for (const Property& p : e.props) {
                if (p.is_list) {
                    const double cnt = rd.scalar(p.count_type);
                    if (!rd.ok || cnt < 0) return yk_set_error(YK_ERR_INVALID, "PLY: truncated payload");
                    const bool take = is_face && (p.name == "vertex_index" || p.name == "vertex_indices") &&
                                      (p.type == PlyType::I32 || p.type == PlyType::U32);  // ListInt / ListUInt only
                    if (take) poly.clear();
                    for (uint64_t k = 0; k < (uint64_t)cnt; ++k) {
                        const double v = rd.scalar(p.type);
                        if (take) {
                            if (v < 0) return yk_set_error(YK_ERR_INVALID, "Negative PLY index");  // ply.rs:292
                            poly.push_back((uint32_t)v);
                        }
                    }
                } else if (is_vert && p.type == PlyType::F32) {
                    const float v = rd.f32(p.type);
                    if (p.name == "x") pt[0] = v;
                    else if (p.name == "y") pt[1] = v;
                    else if (p.name == "z") pt[2] = v;
                    else if (p.name == "nx") nn[0] = v;
                    else if (p.name == "ny") nn[1] = v;
                    else if (p.name == "nz") nn[2] = v;
                    else if (p.name == "u") uv[0] = v;
                    else if (p.name == "v") uv[1] = v;
                } else {
                    (void)rd.scalar(p.type);
                }
                if (!rd.ok) return yk_set_error(YK_ERR_INVALID, "PLY: truncated payload");
            }
            if (is_vert) {
                ply->points.insert(ply->points.end(), pt, pt + 3);
                if (e_has_normal) ply->normals.insert(ply->normals.end(), nn, nn + 3);
                if (e_has_uv) ply->uvs.insert(ply->uvs.end(), uv, uv + 2);
            } else if (is_face && poly.size() >= 3) {
                for (size_t k = 1; k + 1 < poly.size(); ++k) {  // ply.rs:81-92
                    ply->indices.push_back(poly[0]);
                    ply->indices.push_back(poly[k]);
                    ply->indices.push_back(poly[k + 1]);
                }
            }
        }
    }
    const size_t n_points = ply->points.size() / 3;
    for (uint32_t i : ply->indices)
        if (i >= n_points) return yk_set_error(YK_ERR_INVALID, "PLY: face index past the vertex list");
    *out = ply.release();
    return YK_OK;
    });
}

void yk_ply_view(const yk_ply* p, yk_ply_data* out) {
    std::memset(out, 0, sizeof(*out));
    if (!p) return;
    out->n_points = (uint32_t)(p->points.size() / 3);
    out->points = p->points.data();
    out->normals = p->normals.empty() ? nullptr : p->normals.data();
    out->uvs = p->uvs.empty() ? nullptr : p->uvs.data();
    out->n_indices = (uint32_t)p->indices.size();
    out->indices = p->indices.data();
}

void yk_ply_destroy(yk_ply* p) { delete p; }

}  // extern "C"
