// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load anything under oracle/.  PARITY UNPINNED: the reference (vxlerian/cpu-raymarcher) ships no
// tests, golden vectors or fixtures (package.json:7), and cannot be executed in this container
// (no JS engine, gl-matrix 3.4.4 not vendored).  This is a restatement of its TypeScript plus the
// ECMAScript / V8 / gl-matrix numeric semantics it relies on.
//
// jsnum.hpp — the JavaScript numeric model:
//   * every scalar is an IEEE-754 double, arithmetic is never fused (-ffp-contract=off);
//   * every store into a gl-matrix vec3/mat3/mat4 (Float32Array) rounds to float32 (RNE);
//   * Math.min / Math.max propagate NaN and order -0 < +0   (ECMA-262 21.3.2.24/25);
//   * Math.hypot is V8's max-scaled, Kahan-compensated version (v8/src/builtins/math.tq MathHypot);
//   * Uint8ClampedArray stores use ToUint8Clamp (NaN->0, clamp, round-half-to-even);
//   * Uint16Array stores wrap modulo 65536.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace js {

constexpr double kInf = std::numeric_limits<double>::infinity();
constexpr double kNaN = std::numeric_limits<double>::quiet_NaN();

// Math.min(a, b)
inline double min2(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return kNaN;
    if (a < b) return a;
    if (b < a) return b;
    // equal (or +-0): -0 wins
    if (a == 0.0 && b == 0.0) return std::signbit(a) ? a : b;
    return a;
}
// Math.max(a, b)
inline double max2(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return kNaN;
    if (a > b) return a;
    if (b > a) return b;
    if (a == 0.0 && b == 0.0) return std::signbit(a) ? b : a;
    return a;
}
inline double min3(double a, double b, double c) { return min2(min2(a, b), c); }
inline double max3(double a, double b, double c) { return max2(max2(a, b), c); }

// V8 MathHypot for three arguments.
inline double hypot3(double a, double b, double c) {
    double v[3] = {a, b, c};
    double absv[3] = {0, 0, 0};
    bool one_nan = false;
    double mx = 0;
    for (int i = 0; i < 3; ++i) {
        if (std::isnan(v[i])) {
            one_nan = true;
        } else {
            double av = std::fabs(v[i]);
            absv[i] = av;
            if (av > mx) mx = av;
        }
    }
    if (mx == kInf) return kInf;
    if (one_nan) return kNaN;
    if (mx == 0) return 0;
    double sum = 0, comp = 0;
    for (int i = 0; i < 3; ++i) {
        double n = absv[i] / mx;
        double summand = (n * n) - comp;
        double prelim = sum + summand;
        comp = (prelim - sum) - summand;
        sum = prelim;
    }
    return std::sqrt(sum) * mx;
}

// ToUint8Clamp (ECMA-262 7.1.12): NaN -> 0, clamp to [0,255], round half to even.
inline uint8_t to_u8_clamp(double x) {
    if (std::isnan(x)) return 0;
    if (x <= 0) return 0;
    if (x >= 255) return 255;
    double f = std::floor(x);
    if (f + 0.5 < x) return (uint8_t)(f + 1);
    if (x < f + 0.5) return (uint8_t)f;
    // exactly half: even
    uint32_t fi = (uint32_t)f;
    return (uint8_t)((fi % 2 == 0) ? fi : fi + 1);
}

// Math.round: nearest integer, ties toward +Infinity, keeps -0 (V8 Float64Round: ceil, then step down).
inline double round(double x) {
    double r = std::ceil(x);
    return (r - 0.5 > x) ? r - 1.0 : r;
}
// Math.sin / Math.cos.  Current V8 builds route these to the glibc dbl-64 routines it vendors
// (v8_use_libm_trig_functions, third_party/glibc); this container's libm is the same family, so the
// oracle calls it directly.  (Older V8: fdlibm, < 1 ulp as well — a last-bit difference survives the
// float32 store that follows every use in the reference with probability ~2^-29.)
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
// Math.atan2 / asin / pow / log (mandelbulb.ts): V8 uses its fdlibm port (ieee754::atan2 ...), < 1 ulp; the oracle calls
// this container's libm.  The Mandelbulb's iteration stores z in a Float32Array every pass, which absorbs last-bit
// differences with probability 1 - 2^-29 each; parity for that primitive is therefore stated as a pixel-agreement bar.
inline double atan2(double y, double x) { return std::atan2(y, x); }
inline double asin(double x) { return std::asin(x); }
inline double pow(double x, double y) { return std::pow(x, y); }
inline double log(double x) { return std::log(x); }

// float32 store
inline float f32(double x) { return (float)x; }

}  // namespace js
