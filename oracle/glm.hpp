// ORACLE — TEST INFRASTRUCTURE ONLY (see jsnum.hpp header).  PARITY UNPINNED.
//
// glm.hpp — restatement of the gl-matrix 3.4.4 functions that sit on the reference's hot path.
// gl-matrix is a third-party npm dependency (package.json:27, package-lock.json:1258-1263) that is
// NOT present under /root/reference; these bodies restate its published 3.x algorithm
// (ARRAY_TYPE = Float32Array, column-major mat4).  Call sites in the reference:
//   vec3.fromValues      raymarcher.ts:84,99,128-132  box.ts:14,20
//   vec3.transformMat3   raymarcher.ts:87
//   vec3.normalize       raymarcher.ts:88,133  phongModel.ts:16,47,55
//   vec3.scaleAndAdd     raymarcher.ts:95  sphereTracer.ts:45 (every algorithm)  octree.ts:255
//   vec3.transformMat4   primitive.ts:35
//   vec3.length          sphere.ts:13  box.ts:26,33
//   mat3.fromMat4        raymarcher.ts:64
//   mat4.invert          primitive.ts:22  boundingBox.ts:140  sceneManager.ts:35
//   mat4.rotateX/Y/Z, fromTranslation, fromRotationTranslationScale   sceneManager.ts:26-31
//   mat4.rotateY/rotateX/translate                                    camera.ts:83-87
// Open question carried from SURVEY.md Appendix B: whether 3.4.4's vec3.length is Math.hypot or
// Math.sqrt(x*x+y*y+z*z).  g_length_uses_hypot selects (default: hypot, the 3.x behaviour).
#pragma once
#include "jsnum.hpp"

namespace glm {

inline bool& length_uses_hypot() {
    static bool v = true;
    return v;
}

// NOTE: there is deliberately no non-const operator[] returning float&: `v[i] + 1` would then be
// evaluated in float arithmetic by C++'s promotion rules, whereas JS reads a Float32Array element as
// a double.  Reads always go through the double-returning accessor; stores go through `.e[i] =`.
struct vec3 {
    float e[3];
    double operator[](int i) const { return (double)e[i]; }  // reads promote exactly
};
struct mat3 {
    float e[9];
};
struct mat4 {
    float e[16];
    double operator[](int i) const { return (double)e[i]; }
};

inline vec3 v3_create() { return vec3{{0.f, 0.f, 0.f}}; }
inline vec3 v3_from(double x, double y, double z) { return vec3{{js::f32(x), js::f32(y), js::f32(z)}}; }

inline double v3_length(const vec3& a) {
    double x = a[0], y = a[1], z = a[2];
    if (length_uses_hypot()) return js::hypot3(x, y, z);
    return std::sqrt(x * x + y * y + z * z);
}
inline void v3_normalize(vec3& out, const vec3& a) {
    double x = a[0], y = a[1], z = a[2];
    double len = x * x + y * y + z * z;
    if (len > 0) len = 1 / std::sqrt(len);
    out.e[0] = js::f32(a[0] * len);
    out.e[1] = js::f32(a[1] * len);
    out.e[2] = js::f32(a[2] * len);
}
inline void v3_scale_and_add(vec3& out, const vec3& a, const vec3& b, double s) {
    double r0 = a[0] + b[0] * s, r1 = a[1] + b[1] * s, r2 = a[2] + b[2] * s;
    out.e[0] = js::f32(r0);
    out.e[1] = js::f32(r1);
    out.e[2] = js::f32(r2);
}
inline double v3_dot(const vec3& a, const vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void v3_scale(vec3& out, const vec3& a, double s) {
    double r0 = a[0] * s, r1 = a[1] * s, r2 = a[2] * s;
    out.e[0] = js::f32(r0);
    out.e[1] = js::f32(r1);
    out.e[2] = js::f32(r2);
}
inline void v3_subtract(vec3& out, const vec3& a, const vec3& b) {
    double r0 = a[0] - b[0], r1 = a[1] - b[1], r2 = a[2] - b[2];
    out.e[0] = js::f32(r0);
    out.e[1] = js::f32(r1);
    out.e[2] = js::f32(r2);
}
// gl-matrix 3.4.4 vec3.distance: Math.hypot of the component differences (same caveat as vec3.length)
inline double v3_distance(const vec3& a, const vec3& b) {
    double x = b[0] - a[0], y = b[1] - a[1], z = b[2] - a[2];
    if (length_uses_hypot()) return js::hypot3(x, y, z);
    return std::sqrt(x * x + y * y + z * z);
}
inline void v3_transform_mat3(vec3& out, const vec3& a, const mat3& m) {
    double x = a[0], y = a[1], z = a[2];
    double r0 = x * (double)m.e[0] + y * (double)m.e[3] + z * (double)m.e[6];
    double r1 = x * (double)m.e[1] + y * (double)m.e[4] + z * (double)m.e[7];
    double r2 = x * (double)m.e[2] + y * (double)m.e[5] + z * (double)m.e[8];
    out.e[0] = js::f32(r0);
    out.e[1] = js::f32(r1);
    out.e[2] = js::f32(r2);
}
inline void v3_transform_mat4(vec3& out, const vec3& a, const mat4& m) {
    double x = a[0], y = a[1], z = a[2];
    double w = m[3] * x + m[7] * y + m[11] * z + m[15];
    if (w == 0.0 || std::isnan(w)) w = 1.0;  // w = w || 1.0
    double r0 = (m[0] * x + m[4] * y + m[8] * z + m[12]) / w;
    double r1 = (m[1] * x + m[5] * y + m[9] * z + m[13]) / w;
    double r2 = (m[2] * x + m[6] * y + m[10] * z + m[14]) / w;
    out.e[0] = js::f32(r0);
    out.e[1] = js::f32(r1);
    out.e[2] = js::f32(r2);
}

inline void m4_scale(mat4& out, const mat4& a, const double v[3]);
inline mat4 m4_create() {
    mat4 m{};
    m.e[0] = m.e[5] = m.e[10] = m.e[15] = 1.f;
    return m;
}
inline void m3_from_mat4(mat3& out, const mat4& a) {
    out.e[0] = a.e[0]; out.e[1] = a.e[1]; out.e[2] = a.e[2];
    out.e[3] = a.e[4]; out.e[4] = a.e[5]; out.e[5] = a.e[6];
    out.e[6] = a.e[8]; out.e[7] = a.e[9]; out.e[8] = a.e[10];
}

// mat4.invert: returns false when det == 0 (gl-matrix returns null and leaves `out` untouched).
inline bool m4_invert(mat4& out, const mat4& a) {
    double a00 = a[0], a01 = a[1], a02 = a[2], a03 = a[3];
    double a10 = a[4], a11 = a[5], a12 = a[6], a13 = a[7];
    double a20 = a[8], a21 = a[9], a22 = a[10], a23 = a[11];
    double a30 = a[12], a31 = a[13], a32 = a[14], a33 = a[15];
    double b00 = a00 * a11 - a01 * a10;
    double b01 = a00 * a12 - a02 * a10;
    double b02 = a00 * a13 - a03 * a10;
    double b03 = a01 * a12 - a02 * a11;
    double b04 = a01 * a13 - a03 * a11;
    double b05 = a02 * a13 - a03 * a12;
    double b06 = a20 * a31 - a21 * a30;
    double b07 = a20 * a32 - a22 * a30;
    double b08 = a20 * a33 - a23 * a30;
    double b09 = a21 * a32 - a22 * a31;
    double b10 = a21 * a33 - a23 * a31;
    double b11 = a22 * a33 - a23 * a32;
    double det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06;
    if (det == 0.0 || std::isnan(det)) return false;  // if (!det) return null
    det = 1.0 / det;
    float o[16];
    o[0] = js::f32((a11 * b11 - a12 * b10 + a13 * b09) * det);
    o[1] = js::f32((a02 * b10 - a01 * b11 - a03 * b09) * det);
    o[2] = js::f32((a31 * b05 - a32 * b04 + a33 * b03) * det);
    o[3] = js::f32((a22 * b04 - a21 * b05 - a23 * b03) * det);
    o[4] = js::f32((a12 * b08 - a10 * b11 - a13 * b07) * det);
    o[5] = js::f32((a00 * b11 - a02 * b08 + a03 * b07) * det);
    o[6] = js::f32((a32 * b02 - a30 * b05 - a33 * b01) * det);
    o[7] = js::f32((a20 * b05 - a22 * b02 + a23 * b01) * det);
    o[8] = js::f32((a10 * b10 - a11 * b08 + a13 * b06) * det);
    o[9] = js::f32((a01 * b08 - a00 * b10 - a03 * b06) * det);
    o[10] = js::f32((a30 * b04 - a31 * b02 + a33 * b00) * det);
    o[11] = js::f32((a21 * b02 - a20 * b04 - a23 * b00) * det);
    o[12] = js::f32((a11 * b07 - a10 * b09 - a12 * b06) * det);
    o[13] = js::f32((a00 * b09 - a01 * b07 + a02 * b06) * det);
    o[14] = js::f32((a31 * b01 - a30 * b03 - a32 * b00) * det);
    o[15] = js::f32((a20 * b03 - a21 * b01 + a22 * b00) * det);
    for (int i = 0; i < 16; ++i) out.e[i] = o[i];
    return true;
}

// mat4.rotateX/Y/Z(out, a, rad) — `a` is copied first so out may alias a.
inline void m4_rotate_x(mat4& out, const mat4& a_in, double rad) {
    mat4 a = a_in;
    double s = std::sin(rad), c = std::cos(rad);
    double a10 = a[4], a11 = a[5], a12 = a[6], a13 = a[7];
    double a20 = a[8], a21 = a[9], a22 = a[10], a23 = a[11];
    out = a;
    out.e[4] = js::f32(a10 * c + a20 * s);
    out.e[5] = js::f32(a11 * c + a21 * s);
    out.e[6] = js::f32(a12 * c + a22 * s);
    out.e[7] = js::f32(a13 * c + a23 * s);
    out.e[8] = js::f32(a20 * c - a10 * s);
    out.e[9] = js::f32(a21 * c - a11 * s);
    out.e[10] = js::f32(a22 * c - a12 * s);
    out.e[11] = js::f32(a23 * c - a13 * s);
}
inline void m4_rotate_y(mat4& out, const mat4& a_in, double rad) {
    mat4 a = a_in;
    double s = std::sin(rad), c = std::cos(rad);
    double a00 = a[0], a01 = a[1], a02 = a[2], a03 = a[3];
    double a20 = a[8], a21 = a[9], a22 = a[10], a23 = a[11];
    out = a;
    out.e[0] = js::f32(a00 * c - a20 * s);
    out.e[1] = js::f32(a01 * c - a21 * s);
    out.e[2] = js::f32(a02 * c - a22 * s);
    out.e[3] = js::f32(a03 * c - a23 * s);
    out.e[8] = js::f32(a00 * s + a20 * c);
    out.e[9] = js::f32(a01 * s + a21 * c);
    out.e[10] = js::f32(a02 * s + a22 * c);
    out.e[11] = js::f32(a03 * s + a23 * c);
}
inline void m4_rotate_z(mat4& out, const mat4& a_in, double rad) {
    mat4 a = a_in;
    double s = std::sin(rad), c = std::cos(rad);
    double a00 = a[0], a01 = a[1], a02 = a[2], a03 = a[3];
    double a10 = a[4], a11 = a[5], a12 = a[6], a13 = a[7];
    out = a;
    out.e[0] = js::f32(a00 * c + a10 * s);
    out.e[1] = js::f32(a01 * c + a11 * s);
    out.e[2] = js::f32(a02 * c + a12 * s);
    out.e[3] = js::f32(a03 * c + a13 * s);
    out.e[4] = js::f32(a10 * c - a00 * s);
    out.e[5] = js::f32(a11 * c - a01 * s);
    out.e[6] = js::f32(a12 * c - a02 * s);
    out.e[7] = js::f32(a13 * c - a03 * s);
}
// mat4.translate(out, a, v)
inline void m4_translate(mat4& out, const mat4& a_in, const vec3& v) {
    mat4 a = a_in;
    double x = v[0], y = v[1], z = v[2];
    for (int i = 0; i < 12; ++i) out.e[i] = a.e[i];
    out.e[12] = js::f32(a[0] * x + a[4] * y + a[8] * z + a[12]);
    out.e[13] = js::f32(a[1] * x + a[5] * y + a[9] * z + a[13]);
    out.e[14] = js::f32(a[2] * x + a[6] * y + a[10] * z + a[14]);
    out.e[15] = js::f32(a[3] * x + a[7] * y + a[11] * z + a[15]);
}
// mat4.fromTranslation(out, [x,y,z])  (plain JS array of doubles)
inline void m4_from_translation(mat4& out, double x, double y, double z) {
    out = m4_create();
    out.e[12] = js::f32(x);
    out.e[13] = js::f32(y);
    out.e[14] = js::f32(z);
}
// mat4.fromRotationTranslationScale(out, q, v, s) with plain JS arrays.
inline void m4_from_rts(mat4& out, const double q[4], const double v[3], const double s[3]) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double x2 = x + x, y2 = y + y, z2 = z + z;
    double xx = x * x2, xy = x * y2, xz = x * z2;
    double yy = y * y2, yz = y * z2, zz = z * z2;
    double wx = w * x2, wy = w * y2, wz = w * z2;
    double sx = s[0], sy = s[1], sz = s[2];
    out.e[0] = js::f32((1 - (yy + zz)) * sx);
    out.e[1] = js::f32((xy + wz) * sx);
    out.e[2] = js::f32((xz - wy) * sx);
    out.e[3] = 0;
    out.e[4] = js::f32((xy - wz) * sy);
    out.e[5] = js::f32((1 - (xx + zz)) * sy);
    out.e[6] = js::f32((yz + wx) * sy);
    out.e[7] = 0;
    out.e[8] = js::f32((xz + wy) * sz);
    out.e[9] = js::f32((yz - wx) * sz);
    out.e[10] = js::f32((1 - (xx + yy)) * sz);
    out.e[11] = 0;
    out.e[12] = js::f32(v[0]);
    out.e[13] = js::f32(v[1]);
    out.e[14] = js::f32(v[2]);
    out.e[15] = 1;
}

// mat4.scale(out, a, v): columns 0..2 scaled, column 3 copied (gl-matrix 3.4.4)
inline void m4_scale(mat4& out, const mat4& a, const double v[3]) {
    mat4 r;
    for (int c = 0; c < 3; ++c)
        for (int k = 0; k < 4; ++k) r.e[4 * c + k] = js::f32(a[4 * c + k] * v[c]);
    for (int k = 12; k < 16; ++k) r.e[k] = a.e[k];
    out = r;
}

}  // namespace glm
