"""Scene presets — host-side mirror of src/util/sceneManager.ts restricted to the hot path's
primitives (sphere / box / torus, hard-min union): presets 0-5 and 7-9.  Presets built from SDF
operator trees or the mandelbulb (6, 10-18) are outside the path and raise
UnsupportedPreset — there is no CPU fallback to route them to."""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import glmatrix as gm

SPHERE, BOX, TORUS = 0, 1, 2

PRESET_NAMES = [
    "Sphere", "Random Spheres", "Grid of Spheres", "Dense Sphere Grid", "Atom", "Torus", "Rounded Box", "Cube",
    "Sphere and Cube", "Pyramid of Boxes", "Smooth Union", "Smooth Subtraction", "Smooth Union [A]", "Mandelbulb [A]",
    "Twisted Torus", "Infinite Spheres", "Screw", "Chicken", "67",
]
SUPPORTED_PRESETS = (0, 1, 2, 3, 4, 5, 7, 8, 9)


class UnsupportedPreset(ValueError):
    pass


@dataclass
class PrimitiveList:
    """Scene.objectSDFs flattened to the SoA the C ABI takes (include/rm.h rm_scene)."""
    type: list = field(default_factory=list)
    world_to_local: list = field(default_factory=list)
    params: list = field(default_factory=list)

    def __len__(self):
        return len(self.type)

    def arrays(self):
        n = len(self.type)
        return (np.array(self.type, np.uint8).reshape(n), np.array(self.world_to_local, np.float32).reshape(n, 16),
                np.array(self.params, np.float64).reshape(n, 4))


def get_transform(x, y, z, rotation=None):  # sceneManager.ts:21-37
    if rotation is not None:
        model = gm.mat4_from_translation(x, y, z)
        model = gm.mat4_rotate_x(model, gm.f32(rotation[0]))  # rotation is a vec3 (f32)
        model = gm.mat4_rotate_y(model, gm.f32(rotation[1]))
        model = gm.mat4_rotate_z(model, gm.f32(rotation[2]))
    else:
        model = gm.mat4_from_rotation_translation_scale((0, 0, 0, 1), (x, y, z), (1, 1, 1))
    inv = gm.mat4_invert(model)
    return inv if inv is not None else gm.mat4_create()


def add_sphere(pl: PrimitiveList, x, y, z, radius, rotation=None):  # sceneManager.ts:39-41
    pl.type.append(SPHERE)
    pl.world_to_local.append(get_transform(x, y, z, rotation))
    pl.params.append([float(radius), 0.0, 0.0, 0.0])


def add_box(pl: PrimitiveList, x, y, z, half_size, rotation=None):  # sceneManager.ts:43-45 (vec3 -> f32)
    pl.type.append(BOX)
    pl.world_to_local.append(get_transform(x, y, z, rotation))
    pl.params.append([gm.f32(half_size[0]), gm.f32(half_size[1]), gm.f32(half_size[2]), 0.0])


def add_torus(pl: PrimitiveList, x, y, z, radius, rotation=None):  # sceneManager.ts:47-49
    pl.type.append(TORUS)
    pl.world_to_local.append(get_transform(x, y, z, rotation))
    pl.params.append([float(radius), float(radius) / 4, 0.0, 0.0])


def get_preset(index: int) -> PrimitiveList:
    """SceneManager.getPreset(index).objects (sceneManager.ts:102-207,359-361)."""
    pl = PrimitiveList()
    if index == 0:
        add_sphere(pl, 0, 0, 0, 1.5)
    elif index == 1:
        for c in ((0.8, -0.3, 0.2, 0.4), (-0.5, 0.9, -0.1, 0.5), (0.2, 0.1, 0.8, 0.3), (-0.9, -0.4, -0.6, 0.6),
                  (0.4, -0.8, 0.5, 0.35), (-0.2, 0.6, -0.9, 0.4), (0.7, 0.3, -0.4, 0.25)):
            add_sphere(pl, *c)
    elif index == 2:
        for y in (-1, 0, 1):
            for x in (-1, 0, 1):
                add_sphere(pl, x, y, 0, 0.3)
    elif index == 3:
        grid, spacing = 5, 0.6
        offset = (grid - 1) * spacing / 2
        for x in range(grid):
            for y in range(grid):
                for z in range(grid):
                    add_sphere(pl, x * spacing - offset, y * spacing - offset, z * spacing - offset, 0.15)
    elif index == 4:
        for c in ((0, 0, 0, 0.5), (1.2, 0, 0, 0.3), (-1.2, 0, 0, 0.3), (0, 1.2, 0, 0.3), (0, -1.2, 0, 0.3),
                  (0, 0, 1.2, 0.3), (0, 0, -1.2, 0.3)):
            add_sphere(pl, *c)
    elif index == 5:
        add_torus(pl, 0, 0, 0, 1.3, (-math.pi / 2, 0, 0))
    elif index == 7:
        add_box(pl, 0, 0, 0, (1, 1, 1))
    elif index == 8:
        add_sphere(pl, -0.7, 0, 0, 0.5)
        add_box(pl, 1, 0, 0, (0.5, 0.5, 0.5))
    elif index == 9:
        add_box(pl, 0, 0.5, 0, (0.9, 0.25, 0.9))
        add_box(pl, 0, 0, 0, (0.6, 0.25, 0.6))
        add_box(pl, 0, -0.5, 0, (0.3, 0.25, 0.3))
    elif 0 <= index < len(PRESET_NAMES):
        raise UnsupportedPreset(f"preset {index} ({PRESET_NAMES[index]!r}) uses SDF operators / mandelbulb, which are "
                                "outside the B200 hot path (sphere/box/torus unions only)")
    else:
        raise IndexError(index)
    return pl


def get_preset_count() -> int:
    return len(PRESET_NAMES)


def mulberry32(seed: int):
    """The config-4 generator: 32-bit integer ops only, identical in JS / C++ / Python."""
    a = seed & 0xFFFFFFFF

    def nxt() -> float:
        nonlocal a
        a = (a + 0x6D2B79F5) & 0xFFFFFFFF
        t = ((a ^ (a >> 15)) * (1 | a)) & 0xFFFFFFFF
        t = ((t + (((t ^ (t >> 7)) * (61 | t)) & 0xFFFFFFFF)) & 0xFFFFFFFF) ^ t
        return ((t ^ (t >> 14)) & 0xFFFFFFFF) / 4294967296.0

    return nxt


def synthetic_spheres(n: int, seed: int = 0x5EED0001) -> PrimitiveList:
    """"Random Spheres" (preset 1) scaled to n primitives: prims 0-6 are the reference's, the rest seeded."""
    pl = get_preset(1)
    if n < len(pl):
        pl.type, pl.world_to_local, pl.params = pl.type[:n], pl.world_to_local[:n], pl.params[:n]
    extra = n - len(pl)
    if extra <= 0:
        return pl
    # mulberry32, vectorised: the state after k draws is seed + k*0x6D2B79F5 (mod 2^32)
    k = np.arange(1, 4 * extra + 1, dtype=np.uint64)
    a = ((np.uint64(seed & 0xFFFFFFFF) + k * np.uint64(0x6D2B79F5)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    with np.errstate(over="ignore"):
        t = ((a ^ (a >> np.uint32(15))) * (np.uint32(1) | a)).astype(np.uint32)
        t = ((t + ((t ^ (t >> np.uint32(7))) * (np.uint32(61) | t)).astype(np.uint32)).astype(np.uint32)) ^ t
        u = (t ^ (t >> np.uint32(14))).astype(np.float64) / 4294967296.0
    u = u.reshape(extra, 4)
    centres = -2.5 + 5 * u[:, :3]
    radii = 0.02 + 0.03 * u[:, 3]
    w2l = get_transform_batch(centres)
    pl.type.extend([SPHERE] * extra)
    pl.world_to_local.extend(w2l.tolist())
    pl.params.extend(np.stack([radii, np.zeros(extra), np.zeros(extra), np.zeros(extra)], 1).tolist())
    return pl


def get_transform_batch(xyz: np.ndarray) -> np.ndarray:
    """get_transform(x, y, z) (no rotation) for many primitives at once: the same
    fromRotationTranslationScale + mat4.invert arithmetic, vectorised (float64 math, float32 stores)."""
    n = xyz.shape[0]
    a = np.zeros((n, 16), np.float64)
    a[:, 0] = a[:, 5] = a[:, 10] = a[:, 15] = 1.0
    a[:, 12:15] = xyz.astype(np.float32).astype(np.float64)
    a00, a01, a02, a03, a10, a11, a12, a13, a20, a21, a22, a23, a30, a31, a32, a33 = (a[:, i] for i in range(16))
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
    det = 1.0 / det
    o = np.stack([
        (a11 * b11 - a12 * b10 + a13 * b09) * det, (a02 * b10 - a01 * b11 - a03 * b09) * det,
        (a31 * b05 - a32 * b04 + a33 * b03) * det, (a22 * b04 - a21 * b05 - a23 * b03) * det,
        (a12 * b08 - a10 * b11 - a13 * b07) * det, (a00 * b11 - a02 * b08 + a03 * b07) * det,
        (a32 * b02 - a30 * b05 - a33 * b01) * det, (a20 * b05 - a22 * b02 + a23 * b01) * det,
        (a10 * b10 - a11 * b08 + a13 * b06) * det, (a01 * b08 - a00 * b10 - a03 * b06) * det,
        (a30 * b04 - a31 * b02 + a33 * b00) * det, (a21 * b02 - a20 * b04 - a23 * b00) * det,
        (a11 * b07 - a10 * b09 - a12 * b06) * det, (a00 * b09 - a01 * b07 + a02 * b06) * det,
        (a31 * b01 - a30 * b03 - a32 * b00) * det, (a20 * b03 - a21 * b01 + a22 * b00) * det,
    ], axis=1)
    return o.astype(np.float32)
