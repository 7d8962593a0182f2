"""Scene presets — host-side mirror of src/util/sceneManager.ts: sphere / box / torus / mandelbulb primitives, the
six SDF operators of src/util/primitive_operations/ (Round, Twist, SmoothUnion, SmoothSubtraction, Repetition,
AnimatedTranslate) and all 19 presets."""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import glmatrix as gm

SPHERE, BOX, TORUS, MANDELBULB = 0, 1, 2, 3

PRESET_NAMES = [
    "Sphere", "Random Spheres", "Grid of Spheres", "Dense Sphere Grid", "Atom", "Torus", "Rounded Box", "Cube",
    "Sphere and Cube", "Pyramid of Boxes", "Smooth Union", "Smooth Subtraction", "Smooth Union [A]", "Mandelbulb [A]",
    "Twisted Torus", "Infinite Spheres", "Screw", "Chicken", "67",
]
SUPPORTED_PRESETS = tuple(range(19))
OPERATOR_PRESETS = (6, 10, 11, 12, 14, 15, 16, 17, 18)
NODE_PRESETS = OPERATOR_PRESETS + (13,)  # built from Node objects (operators, mandelbulb)

# numpy view of rm_op_node (include/rm.h)
OP_NODE_DTYPE = np.dtype([("kind", np.int32), ("child", np.int32, 2), ("prim", np.int32), ("p", np.float64, 4),
                          ("dir", np.float32, 4), ("transform", np.float32, 16)])
NODE_KINDS = {"primitive": 0, "round": 1, "twist": 2, "smooth-union": 3, "smooth-subtraction": 4, "repetition": 5,
              "animated-translate": 6}


class UnsupportedPreset(ValueError):
    pass


@dataclass
class PrimitiveList:
    """Scene.objectSDFs flattened to the SoA the C ABI takes (include/rm.h rm_scene)."""
    type: list = field(default_factory=list)
    world_to_local: list = field(default_factory=list)
    params: list = field(default_factory=list)
    # operator trees (rm_scene.op_nodes / object_root); None = every primitive is a scene object
    op_nodes: np.ndarray | None = None
    object_root: np.ndarray | None = None

    def __len__(self):
        return len(self.type)

    @property
    def n_objects(self) -> int:
        return len(self.object_root) if self.object_root is not None else len(self.type)

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


# ---- operator trees: one Node per reference Primitive object --------------------------------------------
@dataclass
class Node:
    kind: str                 # key of NODE_KINDS
    transform: list           # Primitive.transform (primitive.ts:10)
    children: tuple = ()
    type: int = -1            # leaves: SPHERE / BOX / TORUS
    params: tuple = (0.0, 0.0, 0.0, 0.0)
    p: tuple = (0.0, 0.0, 0.0, 0.0)
    dir: tuple = (0.0, 0.0, 0.0)


def create_sphere(x, y, z, radius, rotation=None) -> Node:  # sceneManager.ts:39-41
    return Node("primitive", get_transform(x, y, z, rotation), type=SPHERE, params=(float(radius), 0.0, 0.0, 0.0))


def create_box(x, y, z, half_size, rotation=None) -> Node:  # sceneManager.ts:43-45
    return Node("primitive", get_transform(x, y, z, rotation), type=BOX,
                params=(gm.f32(half_size[0]), gm.f32(half_size[1]), gm.f32(half_size[2]), 0.0))


def create_torus(x, y, z, radius, rotation=None) -> Node:  # sceneManager.ts:47-49
    return Node("primitive", get_transform(x, y, z, rotation), type=TORUS, params=(float(radius), float(radius) / 4, 0.0, 0.0))


def create_mandelbulb(x, y, z, power=8.0, iterations=9, enable_animation=True, animation_speed=0.05, rotation=None) -> Node:
    """sceneManager.ts:51-71: getTransform, then mat4.scale(transform, transform, [0.5, 0.5, 0.5])."""
    transform = gm.mat4_scale(get_transform(x, y, z, rotation), (0.5, 0.5, 0.5))
    return Node("primitive", transform, type=MANDELBULB,
                params=(float(power), float(iterations), 1.0 if enable_animation else 0.0, float(animation_speed)))


def create_smooth_union(prim1: Node, prim2: Node, k: float) -> Node:  # smoothUnion.ts:9-15: super(mat4.create())
    return Node("smooth-union", gm.mat4_create(), (prim1, prim2), p=(float(k), 0.0, 0.0, 0.0))


def create_smooth_subtract(prim1: Node, prim2: Node, k: float) -> Node:  # smoothSubstraction.ts:9-14
    return Node("smooth-subtraction", gm.mat4_create(), (prim1, prim2), p=(float(k), 0.0, 0.0, 0.0))


def create_twist(prim: Node, twist_amount: float = 10.0) -> Node:  # twist.ts:8-12: super(primitive.transform)
    return Node("twist", prim.transform, (prim,), p=(float(twist_amount), 0.0, 0.0, 0.0))


def create_round(prim: Node, radius: float) -> Node:  # round.ts:8-12
    return Node("round", prim.transform, (prim,), p=(float(radius), 0.0, 0.0, 0.0))


def create_repetition(prim: Node, spacing) -> Node:  # repetition.ts:8-12 (spacing is a vec3: f32)
    return Node("repetition", prim.transform, (prim,), p=(gm.f32(spacing[0]), gm.f32(spacing[1]), gm.f32(spacing[2]), 0.0))


def create_animated_translate(prim: Node, direction=(1, 0, 0), amplitude: float = 2.0, speed: float = 0.5) -> Node:
    """animatedTranslate.ts:15-28: direction = vec3.normalize(direction) (double math, f32 stores)."""
    x, y, z = (gm.f32(direction[0]), gm.f32(direction[1]), gm.f32(direction[2]))
    ln = x * x + y * y + z * z
    if ln > 0:
        ln = 1 / math.sqrt(ln)
    return Node("animated-translate", prim.transform, (prim,), p=(float(amplitude), float(speed), 0.0, 0.0),
                dir=(gm.f32(x * ln), gm.f32(y * ln), gm.f32(z * ln)))


def flatten(objects) -> PrimitiveList:
    """Scene.objectSDFs -> the arrays of rm_scene.  Nodes are laid out in pre-order per object and leaves are
    numbered in encounter order.  A list of plain primitives needs no node array."""
    pl = PrimitiveList()
    nodes = []

    def visit(n: Node) -> int:
        me = len(nodes)
        rec = dict(kind=NODE_KINDS[n.kind], child=[-1, -1], prim=-1, p=list(n.p), dir=list(n.dir) + [0.0],
                   transform=list(n.transform))
        nodes.append(rec)
        if n.kind == "primitive":
            rec["prim"] = len(pl.type)
            pl.type.append(n.type)
            pl.world_to_local.append(list(n.transform))
            pl.params.append(list(n.params))
            rec["p"] = [0.0] * 4
        else:
            for k, c in enumerate(n.children):
                rec["child"][k] = visit(c)
        return me

    roots = [visit(o) for o in objects]
    if any(o.kind != "primitive" for o in objects):
        arr = np.zeros(len(nodes), OP_NODE_DTYPE)
        for i, rec in enumerate(nodes):
            arr[i] = (rec["kind"], rec["child"], rec["prim"], rec["p"], rec["dir"], rec["transform"])
        pl.op_nodes = arr
        pl.object_root = np.array(roots, np.int32)
    return pl


_CHICKEN_BOXES = ((0, 0, 0, 0.6, 0.6, 0.8), (0, -0.2, 0, 0.8, 0.4, 0.6), (0, -0.8, 0.8, 0.4, 0.6, 0.3),
                  (0, -0.8, 1.2, 0.4, 0.2, 0.2), (0, -0.4, 1.0, 0.2, 0.2, 0.2), (0.3, 1, 0, 0.1, 0.6, 0.01),
                  (-0.3, 1, 0, 0.1, 0.6, 0.01), (0, 1.6, 0.2, 0.6, 0.01, 0.2), (0.3, 1.6, 0.5, 0.1, 0.01, 0.1),
                  (-0.3, 1.6, 0.5, 0.1, 0.01, 0.1))


def _operator_preset(index: int):
    """sceneManager.ts:177-186, 209-246, 254-356."""
    if index == 6:
        return [create_round(create_box(0, 0, 0, (0.4, 0.4, 0.4)), 0.3)]
    if index == 10:
        return [create_smooth_union(create_sphere(0, 0, 0, 0.5), create_box(0, 0.5, 0, (1, 0.2, 1)), 0.2)]
    if index == 11:
        return [create_smooth_subtract(create_round(create_box(0, 0, 0, (1, 1, 1), (0, math.pi / 4, 0)), 0.1),
                                       create_sphere(0, 0, 0, 0.9), 0.2)]
    if index == 12:
        return [create_smooth_union(create_animated_translate(create_sphere(0, 0, 0, 1), (1, 0, 0), 3.0, 0.005),
                                    create_sphere(0, 0, 0, 1), 0.2)]
    if index == 13:
        return [create_mandelbulb(0, 0, 0, 8, 80, True, -0.0001)]
    if index == 14:
        return [create_twist(create_torus(0, 0, 0, 1.3, (-math.pi / 2, 0, 0)), 3)]
    if index == 15:
        return [create_repetition(create_sphere(0, 0, 0, 0.3), (1.5, 1.5, 1.5))]
    if index == 16:
        return [create_round(create_twist(create_box(0, 0, 0, (0.4, 1.5, 0.4)), 4.0), 0.1)]
    if index == 17:
        acc = create_box(*_CHICKEN_BOXES[0][:3], _CHICKEN_BOXES[0][3:])
        for b in _CHICKEN_BOXES[1:]:
            acc = create_smooth_union(acc, create_box(*b[:3], b[3:]), 0.0001)
        return [acc]
    if index == 18:
        return [
            create_smooth_union(create_round(create_box(-1.25, -0.8, 0, (0.05, 0.7, 0.05), (0, 0, math.pi / 5)), 0.20),
                                create_round(create_torus(-1.25, 0.5, 0, 0.8, (-math.pi / 2, 0, 0)), 0.05), 0.0001),
            create_smooth_union(create_round(create_box(1.35, 0, 0, (0.05, 1.5, 0.05), (0, 0, math.pi / 7)), 0.20),
                                create_round(create_box(1.25, -1.4, 0, (0.05, 0.8, 0.05), (0, 0, math.pi / 2)), 0.20), 0.0001),
        ]
    raise IndexError(index)


def get_preset(index: int) -> PrimitiveList:
    """SceneManager.getPreset(index).objects (sceneManager.ts:102-356,359-361)."""
    if index in NODE_PRESETS:
        return flatten(_operator_preset(index))
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
