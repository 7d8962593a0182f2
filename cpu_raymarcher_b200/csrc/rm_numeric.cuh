// rm_numeric.cuh — the two numeric models the kernels are instantiated with.
//
//  NumJS   (validation build, compiled with -fmad=false):  reproduces the reference's arithmetic
//          bit for bit — JS Numbers are IEEE doubles, never fused; every gl-matrix vec3/mat store
//          rounds to float32.  Used to prove hit masks and per-pixel
//          SDF-call / iteration counters bit-exact against the oracle.
//  NumFast (fp32 fast path): everything in float with FMA contraction and MUFU sqrt.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rm {

#define RM_DEV __device__ __forceinline__

struct NumJS {
    typedef double S;
    static constexpr bool kExact = true;
    // Float32Array store: round-to-nearest-even to f32, then promote back exactly.
    RM_DEV static S vst(S x) { return (double)__double2float_rn(x); }
    RM_DEV static S sqrt_(S x) { return sqrt(x); }  // IEEE-754 correctly rounded
    RM_DEV static S abs_(S x) { return fabs(x); }
    RM_DEV static S nan_() { return __longlong_as_double(0x7ff8000000000000LL); }
    RM_DEV static S inf_() { return __longlong_as_double(0x7ff0000000000000LL); }
    RM_DEV static bool isnan_(S x) { return x != x; }
    RM_DEV static int to_int_rn(S x) { return __double2int_rn(x); }
    RM_DEV static S pow32(S x) { return pow(x, 32.0); }  // Math.pow(x, 32) (phongModel.ts:57-60)
};

struct NumFast {
    typedef float S;
    static constexpr bool kExact = false;
    RM_DEV static S vst(S x) { return x; }
    RM_DEV static S sqrt_(S x) {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    RM_DEV static S abs_(S x) { return fabsf(x); }
    RM_DEV static S nan_() { return __int_as_float(0x7fc00000); }
    RM_DEV static S inf_() { return __int_as_float(0x7f800000); }
    RM_DEV static bool isnan_(S x) { return x != x; }
    RM_DEV static int to_int_rn(S x) { return __float2int_rn(x); }
    RM_DEV static S pow32(S x) {
        S a = x * x;  // x^2
        a = a * a;    // 4
        a = a * a;    // 8
        a = a * a;    // 16
        return a * a;  // 32
    }
};

// Math.min / Math.max: NaN-propagating (ECMA-262).  The -0 < +0 ordering is not reproduced: no
// consumer on the path can observe the sign of a zero (no division by, or copysign of, these values).
template <class NP>
RM_DEV typename NP::S jsmin(typename NP::S a, typename NP::S b) {
    if (NP::isnan_(a) || NP::isnan_(b)) return NP::nan_();
    return a < b ? a : b;
}
template <class NP>
RM_DEV typename NP::S jsmax(typename NP::S a, typename NP::S b) {
    if (NP::isnan_(a) || NP::isnan_(b)) return NP::nan_();
    return a > b ? a : b;
}

// ToUint8Clamp: NaN -> 0, clamp to [0,255], round half to even (cvt.rni).
template <class NP>
RM_DEV uint8_t to_u8_clamp(typename NP::S x) {
    if (!(x > (typename NP::S)0)) return 0;  // also NaN
    if (x >= (typename NP::S)255) return 255;
    return (uint8_t)NP::to_int_rn(x);
}

// V8 Math.hypot(a,b,c): max-scaled, Kahan-compensated (v8/src/builtins/math.tq).
RM_DEV double v8_hypot3(double a, double b, double c) {
    bool one_nan = (a != a) || (b != b) || (c != c);
    double aa = (a != a) ? 0.0 : fabs(a), ab = (b != b) ? 0.0 : fabs(b), ac = (c != c) ? 0.0 : fabs(c);
    double mx = 0.0;
    if (aa > mx) mx = aa;
    if (ab > mx) mx = ab;
    if (ac > mx) mx = ac;
    if (mx == NumJS::inf_()) return mx;
    if (one_nan) return NumJS::nan_();
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

}  // namespace rm
