// rm_numeric.cuh — numeric models of the render kernels.
//
// CONTROL arithmetic (ray set-up, march distance t, sample point p = f32(o + d*t), BVH slab
// intervals, octree skips, quantisation of depth) is ALWAYS done the way the reference's JavaScript
// does it: IEEE doubles, never fused, every gl-matrix vec3 store rounded to float32 (SURVEY.md
// Appendix A).  Both translation units are compiled with -fmad=false so nvcc cannot contract it.
// Keeping control exact in the fast build too is what makes its branch decisions (box containment,
// interval cursor, BVH full-scene fallback) follow the reference's; those decisions are chaotic in
// the last ulp, the SDF values are not.
//
// FIELD arithmetic (the primitive SDF evaluations — >99% of the FLOPs) comes in two models:
//   NumJS   F = double, JS-exact (Math.hypot Kahan sum, f32 stores inside box.ts, NaN-propagating min)
//           -> the fp64 validation build: hit masks, counters, depth bit-exact against the oracle;
//   NumFast F = float, explicit fmaf + MUFU sqrt -> the fp32 fast path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rm {

#define RM_DEV __device__ __forceinline__

// Float32Array store: round-to-nearest-even to f32.
RM_DEV float f32r(double x) { return __double2float_rn(x); }
RM_DEV double d_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
RM_DEV double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// Math.min / Math.max: NaN-propagating (ECMA-262).  The -0 < +0 ordering is not reproduced: no
// consumer on the path can observe the sign of a zero (nothing divides by, or copies the sign of,
// these values).
RM_DEV double jsmin(double a, double b) {
    if (a != a || b != b) return d_nan();
    return a < b ? a : b;
}
RM_DEV double jsmax(double a, double b) {
    if (a != a || b != b) return d_nan();
    return a > b ? a : b;
}

// ToUint8Clamp: NaN -> 0, clamp to [0,255], round half to even (cvt.rni).
RM_DEV unsigned to_u8_clamp(double x) {
    if (!(x > 0.0)) return 0u;  // also NaN
    if (x >= 255.0) return 255u;
    return (unsigned)__double2int_rn(x);
}
RM_DEV unsigned to_u8_clamp(float x) {
    if (!(x > 0.f)) return 0u;
    if (x >= 255.f) return 255u;
    return (unsigned)__float2int_rn(x);
}

// V8 Math.hypot(a,b,c): max-scaled, Kahan-compensated (v8/src/builtins/math.tq MathHypot).
RM_DEV double v8_hypot3(double a, double b, double c) {
    bool one_nan = (a != a) || (b != b) || (c != c);
    double aa = (a != a) ? 0.0 : fabs(a), ab = (b != b) ? 0.0 : fabs(b), ac = (c != c) ? 0.0 : fabs(c);
    double mx = 0.0;
    if (aa > mx) mx = aa;
    if (ab > mx) mx = ab;
    if (ac > mx) mx = ac;
    if (mx == d_inf()) return mx;
    if (one_nan) return d_nan();
    if (mx == 0.0) return 0.0;
    double sum = 0.0, comp = 0.0;
    {
        double n = aa / mx;
        double summand = (n * n) - comp;
        double prelim = sum + summand;
        comp = (prelim - sum) - summand;
        sum = prelim;
    }
    {
        double n = ab / mx;
        double summand = (n * n) - comp;
        double prelim = sum + summand;
        comp = (prelim - sum) - summand;
        sum = prelim;
    }
    {
        double n = ac / mx;
        double summand = (n * n) - comp;
        double prelim = sum + summand;
        comp = (prelim - sum) - summand;
        sum = prelim;
    }
    return sqrt(sum) * mx;
}

struct NumJS {
    typedef double F;
    static constexpr bool kExact = true;
    RM_DEV static F fmin_(F a, F b) { return jsmin(a, b); }  // Math.min(primitive.sdf(p), closest)
    RM_DEV static F pow32(F x) { return pow(x, 32.0); }      // Math.pow(x, 32) (phongModel.ts:57-60)
    RM_DEV static F st(F x) { return (double)f32r(x); }       // vec3 store inside shading / normals
    RM_DEV static F rsqrt_(F x) { return 1.0 / sqrt(x); }
};

struct NumFast {
    typedef float F;
    static constexpr bool kExact = false;
    RM_DEV static F fmin_(F a, F b) { return fminf(a, b); }
    RM_DEV static F pow32(F x) {
        F a = x * x;
        a = a * a;
        a = a * a;
        a = a * a;
        return a * a;
    }
    RM_DEV static F st(F x) { return x; }
    RM_DEV static F rsqrt_(F x) { return rsqrtf(x); }
    RM_DEV static F sqrt_(F x) {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
};

}  // namespace rm
