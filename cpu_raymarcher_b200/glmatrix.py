"""Host-side restatement of the gl-matrix 3.4.4 functions the reference's *host* code uses to build
scene and camera matrices (Float32Array storage, double arithmetic, column-major mat4).

In the reference these run in TypeScript and stay there (src/util/sceneManager.ts:21-37,
src/util/camera.ts:81-88); this module exists because no JS runtime is available here and the
Python host must hand the library the same f32 matrices.  gl-matrix is an npm dependency that is not
vendored in the reference (package-lock.json:1258-1263); bodies restate its published 3.x algorithm.
"""
from __future__ import annotations

import math

import numpy as np


def f32(x: float) -> float:
    """Float32Array store: round to nearest-even float32, read back as a double."""
    return float(np.float32(x))


def mat4_create() -> list[float]:
    return [1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0]


def mat4_from_translation(x: float, y: float, z: float) -> list[float]:
    m = mat4_create()
    m[12], m[13], m[14] = f32(x), f32(y), f32(z)
    return m


def mat4_from_rotation_translation_scale(q, v, s) -> list[float]:
    x, y, z, w = q
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz = x * x2, x * y2, x * z2
    yy, yz, zz = y * y2, y * z2, z * z2
    wx, wy, wz = w * x2, w * y2, w * z2
    sx, sy, sz = s
    return [
        f32((1 - (yy + zz)) * sx), f32((xy + wz) * sx), f32((xz - wy) * sx), 0.0,
        f32((xy - wz) * sy), f32((1 - (xx + zz)) * sy), f32((yz + wx) * sy), 0.0,
        f32((xz + wy) * sz), f32((yz - wx) * sz), f32((1 - (xx + yy)) * sz), 0.0,
        f32(v[0]), f32(v[1]), f32(v[2]), 1.0,
    ]


def mat4_rotate_x(a, rad: float) -> list[float]:
    s, c = math.sin(rad), math.cos(rad)
    o = list(a)
    a10, a11, a12, a13 = a[4], a[5], a[6], a[7]
    a20, a21, a22, a23 = a[8], a[9], a[10], a[11]
    o[4], o[5], o[6], o[7] = f32(a10 * c + a20 * s), f32(a11 * c + a21 * s), f32(a12 * c + a22 * s), f32(a13 * c + a23 * s)
    o[8], o[9], o[10], o[11] = f32(a20 * c - a10 * s), f32(a21 * c - a11 * s), f32(a22 * c - a12 * s), f32(a23 * c - a13 * s)
    return o


def mat4_rotate_y(a, rad: float) -> list[float]:
    s, c = math.sin(rad), math.cos(rad)
    o = list(a)
    a00, a01, a02, a03 = a[0], a[1], a[2], a[3]
    a20, a21, a22, a23 = a[8], a[9], a[10], a[11]
    o[0], o[1], o[2], o[3] = f32(a00 * c - a20 * s), f32(a01 * c - a21 * s), f32(a02 * c - a22 * s), f32(a03 * c - a23 * s)
    o[8], o[9], o[10], o[11] = f32(a00 * s + a20 * c), f32(a01 * s + a21 * c), f32(a02 * s + a22 * c), f32(a03 * s + a23 * c)
    return o


def mat4_rotate_z(a, rad: float) -> list[float]:
    s, c = math.sin(rad), math.cos(rad)
    o = list(a)
    a00, a01, a02, a03 = a[0], a[1], a[2], a[3]
    a10, a11, a12, a13 = a[4], a[5], a[6], a[7]
    o[0], o[1], o[2], o[3] = f32(a00 * c + a10 * s), f32(a01 * c + a11 * s), f32(a02 * c + a12 * s), f32(a03 * c + a13 * s)
    o[4], o[5], o[6], o[7] = f32(a10 * c - a00 * s), f32(a11 * c - a01 * s), f32(a12 * c - a02 * s), f32(a13 * c - a03 * s)
    return o


def mat4_translate(a, v) -> list[float]:
    x, y, z = v
    o = list(a)
    o[12] = f32(a[0] * x + a[4] * y + a[8] * z + a[12])
    o[13] = f32(a[1] * x + a[5] * y + a[9] * z + a[13])
    o[14] = f32(a[2] * x + a[6] * y + a[10] * z + a[14])
    o[15] = f32(a[3] * x + a[7] * y + a[11] * z + a[15])
    return o


def mat4_scale(a, v) -> list[float]:
    x, y, z = v
    o = list(a)
    for k in range(4):
        o[k] = f32(a[k] * x)
        o[4 + k] = f32(a[4 + k] * y)
        o[8 + k] = f32(a[8 + k] * z)
    return o


def mat4_invert(a):
    """Returns the inverse, or None when det == 0 (gl-matrix returns null)."""
    a00, a01, a02, a03, a10, a11, a12, a13, a20, a21, a22, a23, a30, a31, a32, a33 = a
    b00 = a00 * a11 - a01 * a10
    b01 = a00 * a12 - a02 * a10
    b02 = a00 * a13 - a03 * a10
    b03 = a01 * a12 - a02 * a11
    b04 = a01 * a13 - a03 * a11
    b05 = a02 * a13 - a03 * a12
    b06 = a20 * a31 - a21 * a30
    b07 = a20 * a32 - a22 * a30
    b08 = a20 * a33 - a23 * a30
    b09 = a21 * a32 - a22 * a31
    b10 = a21 * a33 - a23 * a31
    b11 = a22 * a33 - a23 * a32
    det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06
    if not det:  # 0 or NaN
        return None
    det = 1.0 / det
    return [
        f32((a11 * b11 - a12 * b10 + a13 * b09) * det),
        f32((a02 * b10 - a01 * b11 - a03 * b09) * det),
        f32((a31 * b05 - a32 * b04 + a33 * b03) * det),
        f32((a22 * b04 - a21 * b05 - a23 * b03) * det),
        f32((a12 * b08 - a10 * b11 - a13 * b07) * det),
        f32((a00 * b11 - a02 * b08 + a03 * b07) * det),
        f32((a32 * b02 - a30 * b05 - a33 * b01) * det),
        f32((a20 * b05 - a22 * b02 + a23 * b01) * det),
        f32((a10 * b10 - a11 * b08 + a13 * b06) * det),
        f32((a01 * b08 - a00 * b10 - a03 * b06) * det),
        f32((a30 * b04 - a31 * b02 + a33 * b00) * det),
        f32((a21 * b02 - a20 * b04 - a23 * b00) * det),
        f32((a11 * b07 - a10 * b09 - a12 * b06) * det),
        f32((a00 * b09 - a01 * b07 + a02 * b06) * det),
        f32((a31 * b01 - a30 * b03 - a32 * b00) * det),
        f32((a20 * b03 - a21 * b01 + a22 * b00) * det),
    ]
