"""Orbit camera — host-side mirror of src/util/camera.ts (the reference keeps it on the host too;
the device only ever receives the 3x3 rotation and the position, raymarcher.ts:62-67)."""
from __future__ import annotations

import math

import numpy as np

from . import glmatrix as gm


class Camera:
    def __init__(self):  # camera.ts:13-20
        self.camera_distance = 3.0
        self.pitch = 0.0
        self.yaw = 0.0
        self.camera_transform = gm.mat4_create()
        self._update()

    def rotate_camera(self, pitch: float, yaw: float):  # camera.ts:26-31
        self.pitch = min(max(self.pitch + pitch, -math.pi / 2), math.pi / 2)
        self.yaw += yaw
        self._update()

    def get_angles(self):  # camera.ts:34-36
        return (self.pitch, self.yaw)

    def set_angles(self, pitch: float, yaw: float):  # camera.ts:58-62
        self.pitch = min(max(pitch, -math.pi / 2), math.pi / 2)
        self.yaw = yaw
        self._update()

    def get_rotation_matrix3(self) -> np.ndarray:
        """mat3.fromMat4(camera.getRotationMatrix()) as raymarcher.ts:62-64 builds it (column-major)."""
        t = self.camera_transform
        return np.array([t[0], t[1], t[2], t[4], t[5], t[6], t[8], t[9], t[10]], np.float32)

    def get_position(self) -> np.ndarray:  # camera.ts:64-69
        t = self.camera_transform
        return np.array([t[12], t[13], t[14]], np.float32)

    def _update(self):  # camera.ts:81-88
        temp = gm.mat4_rotate_y(gm.mat4_create(), self.yaw)
        orbit = gm.mat4_rotate_x(temp, self.pitch)
        self.camera_transform = gm.mat4_translate(orbit, (0.0, 0.0, gm.f32(abs(self.camera_distance))))
