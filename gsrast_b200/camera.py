"""Camera matrices exactly as GSRast hands them to the splat draw path.

Mirrors (file:line relative to /root/reference):
  FirstPersonCamera::update          FirstPersonCamera.cpp:28-38   glm::lookAt + glm::perspective
  GSGaussians::draw matrix set-up    apps/gsrast/GSGaussians.cpp:155-176
  initial pose                       apps/gsrast/GSRastWindow.cpp:17-37 (eye (0,0,-5), invertUp)
  DEFAULT_FOV / NEAR / FAR           Config.hpp:21-23

All matrices are float32, column-major flat[16] (flat[col*4+row]) — glm's memory layout,
which is what the rasterizer receives as `viewmatrix` / `projmatrix`.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

DEFAULT_FOV = math.radians(45.0)  # Config.hpp:23
DEFAULT_NEAR = 0.001              # Config.hpp:21
DEFAULT_FAR = 100.0               # Config.hpp:22


def _normalize(v):
    v = np.asarray(v, dtype=np.float32)
    return v / np.float32(np.sqrt(np.dot(v, v)))


def look_at(eye, center, up) -> np.ndarray:
    """glm::lookAt (right-handed). Returns M[col][row] as a (4,4) array indexed [col, row]."""
    eye = np.asarray(eye, dtype=np.float32)
    f = _normalize(np.asarray(center, dtype=np.float32) - eye)
    s = _normalize(np.cross(f, np.asarray(up, dtype=np.float32)))
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[1, 0], m[2, 0] = s
    m[0, 1], m[1, 1], m[2, 1] = u
    m[0, 2], m[1, 2], m[2, 2] = -f
    m[3, 0] = -np.dot(s, eye)
    m[3, 1] = -np.dot(u, eye)
    m[3, 2] = np.dot(f, eye)
    return m


def perspective(fovy: float, aspect: float, near: float, far: float) -> np.ndarray:
    """glm::perspective (right-handed, depth -1..1). (4,4) indexed [col, row]."""
    t = np.float32(math.tan(fovy / 2.0))
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0] = np.float32(1.0) / (np.float32(aspect) * t)
    m[1, 1] = np.float32(1.0) / t
    m[2, 2] = -np.float32(far + near) / np.float32(far - near)
    m[2, 3] = -1.0
    m[3, 2] = -(np.float32(2.0) * np.float32(far) * np.float32(near)) / np.float32(far - near)
    return m


def _matmul_cm(a, b):
    """glm a*b for [col,row]-indexed arrays."""
    # result[c][r] = sum_k a[k][r] * b[c][k]
    return np.einsum("kr,ck->cr", a, b).astype(np.float32)


@dataclasses.dataclass
class Camera:
    """What GSGaussians::draw uploads per frame (GSGaussians.cpp:157-176)."""

    viewmatrix: np.ndarray  # float32[16], column-major, row 2 negated (+z forward)
    projmatrix: np.ndarray  # float32[16], perspective * view (un-flipped view)
    cam_pos: np.ndarray     # float32[3]
    tan_fovx: float
    tan_fovy: float
    width: int
    height: int

    def packed(self) -> np.ndarray:
        """view[16] | proj[16] | cam_pos[3] | pad — 36 floats, the per-frame H2D payload."""
        out = np.zeros(36, dtype=np.float32)
        out[0:16] = self.viewmatrix
        out[16:32] = self.projmatrix
        out[32:35] = self.cam_pos
        return out


def make_camera(eye, center, width: int, height: int, fovy: float = DEFAULT_FOV, up=(0.0, -1.0, 0.0),
                near: float = DEFAULT_NEAR, far: float = DEFAULT_FAR) -> Camera:
    """Build the matrices the way GSRast does: invertUp camera, view row 2 negated,
    proj = perspective * (un-negated) view, tanFOVx = tanFOVy * W / H."""
    view = look_at(eye, center, up)
    persp = perspective(fovy, float(width) / float(height), near, far)
    proj = _matmul_cm(persp, view)
    view = view.copy()
    view[:, 2] *= np.float32(-1.0)  # row 2 of every column (GSGaussians.cpp:160-169)
    tan_fovy = float(np.float32(math.tan(fovy * 0.5)))
    tan_fovx = float(np.float32(tan_fovy) * np.float32(np.float32(width) / np.float32(height)))
    return Camera(view.reshape(16).copy(), proj.reshape(16).copy(), np.asarray(eye, dtype=np.float32).copy(),
                  tan_fovx, tan_fovy, width, height)


def default_camera(width: int, height: int, span: float = 7.0) -> Camera:
    """GSRastWindow's initial pose: eye (0,0,-5) looking at the origin, near/far from the
    scene span (GSRastWindow.cpp:29-37)."""
    return make_camera((0.0, 0.0, -5.0), (0.0, 0.0, 0.0), width, height, near=0.001 * span, far=span)


def orbit_cameras(n: int, width: int, height: int, seed: int = 2, span: float = 7.0) -> list[Camera]:
    """C4 pose set: seeded orbit + jitter, radius U(4,6), looking at the origin."""
    rng = np.random.default_rng(seed + 1000)
    cams = []
    for i in range(n):
        ang = 2.0 * math.pi * i / n + rng.normal(0.0, 0.02)
        rad = rng.uniform(4.0, 6.0)
        elev = rng.normal(0.0, 0.15)
        eye = (rad * math.sin(ang) * math.cos(elev), rad * math.sin(elev), -rad * math.cos(ang) * math.cos(elev))
        cams.append(make_camera(eye, (0.0, 0.0, 0.0), width, height, near=0.001 * span, far=2.0 * span))
    return cams
