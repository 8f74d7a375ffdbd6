// Exception barrier of the C ABI: nothing may unwind through an `extern "C"` entry point (a ctypes / Rust FFI caller cannot
// catch it and the process would end in std::terminate). Every fallible exported function runs its body through yk_guard,
// which turns an exception into the ABI's error convention (negative yk_status + yk_last_error()), the counterpart of the
// reference's `Result::Err` on malformed scene files (scene/ply.rs, scene/pbrt/mod.rs, scene/mitsuba/mod.rs).
#pragma once
#include <exception>
#include <new>
#include <string>

#include "yuki_gpu.h"

int yk_set_error(int code, const std::string& msg);  // host_scene.cpp

template <class F>
int yk_guard(const char* who, F&& body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        try { return yk_set_error(YK_ERR_NOMEM, std::string(who) + ": out of host memory"); } catch (...) { return YK_ERR_NOMEM; }
    } catch (const std::exception& e) {
        try { return yk_set_error(YK_ERR_INVALID, std::string(who) + ": " + e.what()); } catch (...) { return YK_ERR_INVALID; }
    } catch (...) {
        try { return yk_set_error(YK_ERR_INVALID, std::string(who) + ": unknown exception"); } catch (...) { return YK_ERR_INVALID; }
    }
}
