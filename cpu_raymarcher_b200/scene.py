"""Scene — host-side mirror of src/util/scene.ts as the worker uses it: a primitive list, a camera
and the name of the acceleration structure.  Distance evaluation itself lives on the GPU."""
from __future__ import annotations

from . import scene_manager as sm
from .camera import Camera


class Scene:
    def __init__(self, acceleration_structure: str = "None"):  # scene.ts:24-29
        self.camera = Camera()
        self.acceleration_structure = acceleration_structure if acceleration_structure in ("Octree", "BVH") else "None"
        self.current_preset_index = 0
        self.primitives = sm.PrimitiveList()
        self.synthetic = None
        self.load_preset(0)

    def load_preset(self, index: int):  # scene.ts:38-59 (index is clamped like Math.max(0, Math.min(...)))
        self.current_preset_index = max(0, min(int(index), sm.get_preset_count() - 1))
        self.primitives = sm.get_preset(self.current_preset_index)
        self.synthetic = None

    def load_synthetic(self, n: int, seed: int = 0x5EED0001):
        """Config 4 of BASELINE.json: 'Random Spheres' scaled to n seeded primitives."""
        self.primitives = sm.synthetic_spheres(n, seed)
        self.current_preset_index = 1
        self.synthetic = (n, seed)

    def key(self):
        return (self.current_preset_index, self.synthetic, self.acceleration_structure)

    def get_current_preset_info(self):  # scene.ts:126-133
        return {"index": self.current_preset_index, "name": sm.PRESET_NAMES[self.current_preset_index],
                "total": sm.get_preset_count()}
