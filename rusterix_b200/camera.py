"""Cameras that produce the view / projection matrices handed to Rasterizer.setup.
Reference: src/camera/d3orbit.rs:18-56,188-195; src/camera/d3firstp.rs:16-42; src/camera/d3iso.rs:102-120.
Host-side input producers only; not part of the device path."""
import math

import numpy as np

from . import vekmath

f32 = np.float32


class D3OrbitCamera:
    def __init__(self):
        self.center = np.zeros(3, dtype=np.float32)
        self.distance = 20.0
        self.azimuth = math.pi / 2.0
        self.elevation = 0.698
        self.up = np.array([0.0, 1.0, 0.0], dtype=np.float32)
        self.fov = 75.0
        self.near = 0.01
        self.far = 100.0

    @staticmethod
    def new():
        return D3OrbitCamera()

    def id(self):
        return "orbit"

    def set_parameter_f32(self, key, value):
        if key == "distance":
            self.distance = float(value)

    def set_parameter_vec2(self, key, value):
        if key == "from_normalized":
            self.azimuth = math.pi * value[0]
            self.elevation = math.pi * (value[1] - 0.5)

    def eye_position(self):
        d, az, el = f32(self.distance), f32(self.azimuth), f32(self.elevation)
        x = d * f32(math.cos(az)) * f32(math.cos(el))
        y = d * f32(math.sin(el))
        z = d * f32(math.sin(az)) * f32(math.cos(el))
        return np.array([x, y, z], dtype=np.float32) + self.center

    position = eye_position

    def view_matrix(self):
        return vekmath.look_at_rh(self.eye_position(), self.center, self.up)

    def projection_matrix(self, width, height):
        return vekmath.perspective_fov_rh_zo(math.radians(self.fov), width, height, self.near, self.far)


class D3FirstPCamera:
    def __init__(self):
        self.position = np.zeros(3, dtype=np.float32)
        self.center = np.zeros(3, dtype=np.float32)
        self.fov = 75.0
        self.near = 0.01
        self.far = 100.0

    @staticmethod
    def new():
        return D3FirstPCamera()

    def id(self):
        return "firstp"

    def set_parameter_f32(self, key, value):
        if key in ("fov", "near", "far"):
            setattr(self, key, float(value))

    def set_parameter_vec3(self, key, value):
        if key == "position":
            self.position = np.asarray(value, dtype=np.float32)
        elif key == "center":
            self.center = np.asarray(value, dtype=np.float32)

    def view_matrix(self):
        return vekmath.look_at_rh(self.position, self.center, np.array([0.0, 1.0, 0.0], dtype=np.float32))

    def projection_matrix(self, width, height):
        return vekmath.perspective_fov_rh_zo(math.radians(self.fov), width, height, self.near, self.far)


class D3IsoCamera:
    def __init__(self):
        self.center = np.zeros(3, dtype=np.float32)
        self.azimuth = math.radians(135.0)
        self.elevation = math.radians(35.264)
        self.distance = 20.0
        self.scale = 4.0
        self.near = 0.1
        self.far = 100.0

    @staticmethod
    def new():
        return D3IsoCamera()

    def id(self):
        return "iso"

    def eye_position(self):
        x = self.distance * math.cos(self.elevation) * math.sin(self.azimuth)
        y = self.distance * math.sin(self.elevation)
        z = self.distance * math.cos(self.elevation) * math.cos(self.azimuth)
        return self.center + np.array([x, y, z], dtype=np.float32)

    def view_matrix(self):
        return vekmath.look_at_rh(self.eye_position(), self.center, np.array([0.0, 1.0, 0.0], dtype=np.float32))

    def projection_matrix(self, width, height):
        half_h = self.scale
        half_w = half_h * (width / height)
        return vekmath.orthographic_rh_no(-half_w, half_w, -half_h, half_h, self.near, self.far)
