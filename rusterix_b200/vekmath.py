"""Host-side float32 matrix helpers standing in for the `vek` calls the reference's callers make
(camera matrices, Mat4::inverted, Mat4::scaling_3d).  These only PRODUCE inputs for the path; the
path itself receives matrices through rxc_frame.  Matrices are numpy float32 arrays indexed
[row, col]; `to_cols` flattens them column-major the way vek stores `Mat4.cols`."""
import math

import numpy as np

f32 = np.float32


def identity4():
    return np.eye(4, dtype=np.float32)


def identity3():
    return np.eye(3, dtype=np.float32)


def to_cols(m):
    """Row/col indexed matrix -> flat column-major list (m[c*n+r])."""
    m = np.asarray(m, dtype=np.float32)
    return np.ascontiguousarray(m.T).reshape(-1)


def _normalized(v):
    v = np.asarray(v, dtype=np.float32)
    mag = f32(math.sqrt(f32(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])))
    return (v / mag).astype(np.float32)


def look_at_rh(eye, target, up):
    """vek Mat4::look_at_rh (reference src/camera/d3orbit.rs:49-52, d3firstp.rs:36-38)."""
    eye = np.asarray(eye, dtype=np.float32)
    target = np.asarray(target, dtype=np.float32)
    up = np.asarray(up, dtype=np.float32)
    f = _normalized(target - eye)
    s = _normalized(np.cross(f, up).astype(np.float32))
    u = np.cross(s, f).astype(np.float32)
    m = np.array(
        [
            [s[0], s[1], s[2], -np.dot(s, eye)],
            [u[0], u[1], u[2], -np.dot(u, eye)],
            [-f[0], -f[1], -f[2], np.dot(f, eye)],
            [0.0, 0.0, 0.0, 1.0],
        ],
        dtype=np.float32,
    )
    return m


def perspective_fov_rh_zo(fov_y_radians, width, height, near, far):
    """vek Mat4::perspective_fov_rh_zo (reference src/camera/d3orbit.rs:54-56)."""
    rad = f32(fov_y_radians)
    h = f32(math.cos(rad / 2.0) / math.sin(rad / 2.0))
    w = f32(h * f32(height) / f32(width))
    near = f32(near)
    far = f32(far)
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0] = w
    m[1, 1] = h
    m[2, 2] = far / (near - far)
    m[2, 3] = -(far * near) / (far - near)
    m[3, 2] = -1.0
    return m


def orthographic_rh_no(left, right, bottom, top, near, far):
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0] = 2.0 / (right - left)
    m[1, 1] = 2.0 / (top - bottom)
    m[2, 2] = -2.0 / (far - near)
    m[0, 3] = -(right + left) / (right - left)
    m[1, 3] = -(top + bottom) / (top - bottom)
    m[2, 3] = -(far + near) / (far - near)
    m[3, 3] = 1.0
    return m


def scaling_3d(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[1, 1], m[2, 2] = x, y, z
    return m


def translation_3d(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[0, 3], m[1, 3], m[2, 3] = x, y, z
    return m


def rotation_y(angle):
    c, s = math.cos(angle), math.sin(angle)
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


def inverted(m):
    """Stand-in for vek Mat4::inverted (reference src/rasterizer.rs:97,116): float64 inverse
    rounded to float32.  A Rust host passes vek's own result through rxc_frame instead."""
    return np.linalg.inv(np.asarray(m, dtype=np.float64)).astype(np.float32)
