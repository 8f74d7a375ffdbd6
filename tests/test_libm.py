"""yk_libm.h restates glibc's sinf / cosf / atanf / atan2f / acosf / logf; on this host it must be bit-identical to the libm the
oracle calls."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstring>
#include "yk_libm.h"
int main() {
    uint64_t bad = 0, n = 0;
    for (uint32_t u = 0x30000000u; u < 0x42f00000u; u += 61) for (int sg = 0; sg < 2; ++sg) {
        uint32_t b = u | (sg ? 0x80000000u : 0u); float x; memcpy(&x, &b, 4);
        float a = sinf(x), c = cosf(x), a2 = yklibm::sinf_glibc(x), c2 = yklibm::cosf_glibc(x);
        bad += memcmp(&a, &a2, 4) != 0; bad += memcmp(&c, &c2, 4) != 0; ++n;
    }
    float specials[] = {0.0f, -0.0f, 1e-30f, 0.78539816f, 0.78539822f, 1.5707964f, 3.1415927f, 6.2831855f, 119.99f};
    for (float x : specials) { float a = sinf(x), a2 = yklibm::sinf_glibc(x), c = cosf(x), c2 = yklibm::cosf_glibc(x);
        bad += memcmp(&a, &a2, 4) != 0; bad += memcmp(&c, &c2, 4) != 0; }
    printf("%llu %llu\n", (unsigned long long)n, (unsigned long long)bad);
    return bad != 0;
}
'''


def test_sinf_cosf_bit_identical_to_host_libm(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "yuki_b200", "csrc"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    n, bad = out.stdout.split()
    assert out.returncode == 0 and int(bad) == 0 and int(n) > 10_000_000, out.stdout


SRC_INV = r'''
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstring>
#include "yk_libm.h"
static uint32_t rng = 12345;
static uint32_t nx() { rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5; return rng; }
static bool differ(float a, float b) { return memcmp(&a, &b, 4) != 0 && !(a != a && b != b); }
int main() {
    unsigned long long n = 0, bad = 0;
    for (uint32_t u = 0; u <= 0x7f800000u; u += 37) for (int sg = 0; sg < 2; ++sg) {  // every binade, both signs, infinities
        uint32_t b = u | (sg ? 0x80000000u : 0u); float x; memcpy(&x, &b, 4);
        bad += differ(atanf(x), yklibm::atanf_glibc(x)); ++n;
    }
    for (uint32_t u = 0; u <= 0x3f800010u; u += 11) for (int sg = 0; sg < 2; ++sg) {  // [-1, 1] and just outside (NaN)
        uint32_t b = u | (sg ? 0x80000000u : 0u); float x; memcpy(&x, &b, 4);
        bad += differ(acosf(x), yklibm::acosf_glibc(x)); ++n;
    }
    for (long i = 0; i < 40000000L; ++i) {
        uint32_t a = nx(), b = nx();
        if (i & 1) {  // magnitudes a sphere's hit point has; the other half sweeps every exponent pair
            a = (a & 0x807fffffu) | ((100 + (a >> 23) % 56) << 23);
            b = (b & 0x807fffffu) | ((100 + (b >> 23) % 56) << 23);
        }
        float y, x; memcpy(&y, &a, 4); memcpy(&x, &b, 4);
        bad += differ(atan2f(y, x), yklibm::atan2f_glibc(y, x)); ++n;
    }
    const float sp[] = {0.0f, -0.0f, 1.0f, -1.0f, INFINITY, -INFINITY, 1e-30f, -1e-30f, 1e30f, NAN, 0.5f, -0.5f};
    for (float y : sp) for (float x : sp) {
        bad += differ(atan2f(y, x), yklibm::atan2f_glibc(y, x));
        bad += differ(acosf(x), yklibm::acosf_glibc(x)); bad += differ(atanf(x), yklibm::atanf_glibc(x)); n += 3;
    }
    printf("%llu %llu\n", n, bad);
    return bad != 0;
}
'''


def test_atanf_atan2f_acosf_bit_identical_to_host_libm(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC_INV)
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "yuki_b200", "csrc"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    n, bad = out.stdout.split()
    assert out.returncode == 0 and int(bad) == 0 and int(n) > 100_000_000, out.stdout


SRC_LOG = r'''
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstring>
#include "yk_libm.h"
int main() {
    unsigned long long n = 0, bad = 0;
    for (uint64_t u = 0; u <= 0x7f800000ull; u += 3) {   // zero, subnormals, every binade, infinity
        uint32_t b = (uint32_t)u; float x; memcpy(&x, &b, 4);
        float a = logf(x), c = yklibm::logf_glibc(x);
        bad += memcmp(&a, &c, 4) != 0; ++n;
    }
    for (uint32_t b = 0x3a000000u; b < 0x3f800000u; ++b) {   // every float of roughness_to_alpha's domain [1e-3, 1)
        float x; memcpy(&x, &b, 4);
        float a = logf(x), c = yklibm::logf_glibc(x);
        bad += memcmp(&a, &c, 4) != 0; ++n;
    }
    float neg = yklibm::logf_glibc(-1.0f), nan_in = yklibm::logf_glibc(NAN);
    bad += !(neg != neg) + !(nan_in != nan_in);
    printf("%llu %llu\n", n, bad);
    return bad != 0;
}
'''


def test_logf_bit_identical_to_host_libm(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC_LOG)
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "yuki_b200", "csrc"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    n, bad = out.stdout.split()
    assert out.returncode == 0 and int(bad) == 0 and int(n) > 700_000_000, out.stdout
