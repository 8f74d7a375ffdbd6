// Minimal OpenEXR 2.0 writer for the film: single-part scanline image, three 32-bit float channels (B, G, R in the
// file's alphabetical channel order), no compression, increasing-y line order. Stands in for
// `exr::prelude::write_rgb_file` as called by yuki/src/app/util.rs:89-110 (the `exr` crate is a crates.io dependency,
// yuki/Cargo.toml, not vendored): same pixels, same channel names and type, readable by any EXR reader.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "yuki_gpu.h"
#include "yk_guard.h"

int yk_set_error(int code, const std::string& msg);  // host_scene.cpp

namespace {

void put(std::vector<uint8_t>& b, const void* p, size_t n) { b.insert(b.end(), (const uint8_t*)p, (const uint8_t*)p + n); }
void put_u32(std::vector<uint8_t>& b, uint32_t v) { put(b, &v, 4); }
void put_i32(std::vector<uint8_t>& b, int32_t v) { put(b, &v, 4); }
void put_u64(std::vector<uint8_t>& b, uint64_t v) { put(b, &v, 8); }
void put_f32(std::vector<uint8_t>& b, float v) { put(b, &v, 4); }
void put_str(std::vector<uint8_t>& b, const char* s) { put(b, s, std::strlen(s) + 1); }
void attr(std::vector<uint8_t>& b, const char* name, const char* type, const std::vector<uint8_t>& value) {
    put_str(b, name);
    put_str(b, type);
    put_u32(b, (uint32_t)value.size());
    put(b, value.data(), value.size());
}

}  // namespace

extern "C" int yk_write_exr(const char* path, uint32_t width, uint32_t height, const float* rgb) {
    return yk_guard("yk_write_exr", [&]() -> int {
    if (!path || !rgb || !width || !height) return yk_set_error(YK_ERR_INVALID, "yk_write_exr: null / empty argument");
    std::vector<uint8_t> head;
    put_u32(head, 20000630u);  // magic
    put_u32(head, 2u);         // version 2, single-part scanline, no long names
    {
        std::vector<uint8_t> ch;
        for (const char* name : {"B", "G", "R"}) {
            put_str(ch, name);
            put_i32(ch, 2);  // FLOAT
            ch.push_back(0);  // pLinear
            ch.push_back(0); ch.push_back(0); ch.push_back(0);
            put_i32(ch, 1);  // xSampling
            put_i32(ch, 1);  // ySampling
        }
        ch.push_back(0);
        attr(head, "channels", "chlist", ch);
    }
    { std::vector<uint8_t> v{0}; attr(head, "compression", "compression", v); }  // NO_COMPRESSION
    {
        std::vector<uint8_t> v;
        put_i32(v, 0); put_i32(v, 0); put_i32(v, (int32_t)width - 1); put_i32(v, (int32_t)height - 1);
        attr(head, "dataWindow", "box2i", v);
        attr(head, "displayWindow", "box2i", v);
    }
    { std::vector<uint8_t> v{0}; attr(head, "lineOrder", "lineOrder", v); }  // INCREASING_Y
    { std::vector<uint8_t> v; put_f32(v, 1.0f); attr(head, "pixelAspectRatio", "float", v); }
    { std::vector<uint8_t> v; put_f32(v, 0.0f); put_f32(v, 0.0f); attr(head, "screenWindowCenter", "v2f", v); }
    { std::vector<uint8_t> v; put_f32(v, 1.0f); attr(head, "screenWindowWidth", "float", v); }
    head.push_back(0);  // end of header

    FILE* f = std::fopen(path, "wb");
    if (!f) return yk_set_error(YK_ERR_INVALID, std::string("Error writing EXR to '") + path + "'");
    const uint64_t line_bytes = (uint64_t)width * 3 * 4;
    const uint64_t table_pos = head.size();
    const uint64_t first_line = table_pos + (uint64_t)height * 8;
    std::vector<uint8_t> table;
    for (uint32_t y = 0; y < height; ++y) put_u64(table, first_line + (uint64_t)y * (8 + line_bytes));
    bool ok = std::fwrite(head.data(), 1, head.size(), f) == head.size() && std::fwrite(table.data(), 1, table.size(), f) == table.size();
    std::vector<float> line((size_t)width * 3);
    for (uint32_t y = 0; y < height && ok; ++y) {
        const float* row = rgb + (size_t)y * width * 3;
        for (uint32_t x = 0; x < width; ++x) {  // planar per scanline, channels in file order B, G, R
            line[x] = row[3 * x + 2];
            line[width + x] = row[3 * x + 1];
            line[2 * (size_t)width + x] = row[3 * x];
        }
        const int32_t yy = (int32_t)y;
        const uint32_t nbytes = (uint32_t)line_bytes;
        ok = std::fwrite(&yy, 4, 1, f) == 1 && std::fwrite(&nbytes, 4, 1, f) == 1 && std::fwrite(line.data(), 4, line.size(), f) == line.size();
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) return yk_set_error(YK_ERR_INVALID, std::string("Error writing EXR to '") + path + "'");
    return YK_OK;
    });
}
