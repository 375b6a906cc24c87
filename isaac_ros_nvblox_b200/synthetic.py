"""Synthetic depth + pose sequences (input generation only; numpy, host side).

Restates what the reference's test scenes produce so that benchmark inputs have
the shape SURVEY.md section 8(d) names (C1..C5):
  * primitives::Scene::generateDepthImageFromScene
    (nvblox/include/nvblox/primitives/internal/impl/scene_impl.h:147-185):
    one ray per pixel centre, nearest analytic hit within max_dist, stored
    value = z of the hit in the camera frame, 0 where nothing is hit.
  * test_utils::getSphereInBox (nvblox/tests/lib/integrator_utils.cpp:33-46).
  * the 80-pose circular trajectory of TsdfIntegratorTestFixture.SphereSceneTest
    (nvblox/tests/test_tsdf_integrator.cpp:191-243).
The generator is an input source for BOTH the CUDA path and the oracle, so its
own arithmetic (float64) is not part of any parity claim.
"""
import math

import numpy as np


class Scene:
    """Axis-aligned planes, spheres and axis-aligned boxes."""

    def __init__(self):
        self.planes = []   # (axis, offset): plane {p[axis] == offset}
        self.spheres = []  # (center(3), radius)
        self.boxes = []    # (min(3), max(3))

    def add_plane(self, axis, offset):
        self.planes.append((int(axis), float(offset)))
        return self

    def add_sphere(self, center, radius):
        self.spheres.append((np.asarray(center, dtype=np.float64), float(radius)))
        return self

    def add_box(self, bmin, bmax):
        self.boxes.append((np.asarray(bmin, dtype=np.float64), np.asarray(bmax, dtype=np.float64)))
        return self

    def ray_distance(self, origin, dirs, max_dist):
        """Nearest hit distance along unit rays dirs (...,3) from origin; inf if none."""
        o = np.asarray(origin, dtype=np.float64)
        best = np.full(dirs.shape[:-1], np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            for axis, off in self.planes:
                t = (off - o[axis]) / dirs[..., axis]
                t = np.where(np.isfinite(t) & (t > 0), t, np.inf)
                best = np.minimum(best, t)
            for c, r in self.spheres:
                oc = o - c
                b = dirs @ oc
                cc = oc @ oc - r * r
                disc = b * b - cc
                sq = np.sqrt(np.maximum(disc, 0.0))
                t0, t1 = -b - sq, -b + sq
                t = np.where(t0 > 0, t0, t1)
                t = np.where((disc >= 0) & (t > 0), t, np.inf)
                best = np.minimum(best, t)
            for bmin, bmax in self.boxes:
                inv = 1.0 / dirs
                ta = (bmin - o) * inv
                tb = (bmax - o) * inv
                tmin = np.nanmax(np.minimum(ta, tb), axis=-1)
                tmax = np.nanmin(np.maximum(ta, tb), axis=-1)
                t = np.where(tmin > 0, tmin, tmax)
                t = np.where((tmax >= tmin) & (t > 0), t, np.inf)
                best = np.minimum(best, t)
        return np.where(best <= max_dist, best, np.inf)

    def distance(self, points):
        """Distance of points (...,3) to the nearest primitive: unsigned for planes, signed (negative inside)
        for spheres and boxes (Scene::getSignedDistanceToPoint, primitives/scene.h, restricted to the room's
        interior where the reference's inward-facing plane normals give the same value)."""
        p = np.asarray(points, dtype=np.float64)
        best = np.full(p.shape[:-1], np.inf)
        for axis, off in self.planes:
            best = np.minimum(best, np.abs(p[..., axis] - off))
        for c, r in self.spheres:
            best = np.minimum(best, np.linalg.norm(p - c, axis=-1) - r)
        for bmin, bmax in self.boxes:
            q = np.maximum(bmin - p, p - bmax)
            outside = np.linalg.norm(np.maximum(q, 0.0), axis=-1)
            inside = np.minimum(np.max(q, axis=-1), 0.0)
            best = np.minimum(best, outside + inside)
        return best


def sphere_in_box():
    """getSphereInBox: ground z=0, ceiling z=5, walls at +-5, sphere r=2 at (0,0,2)."""
    s = Scene()
    s.add_plane(2, 0.0).add_plane(2, 5.0)
    s.add_plane(0, -5.0).add_plane(0, 5.0).add_plane(1, -5.0).add_plane(1, 5.0)
    s.add_sphere((0.0, 0.0, 2.0), 2.0)
    return s


def box_with_cube():
    """Obstacle::kBoxWithCube shape: the same room with a 2 m cube at (0,0,2)."""
    s = Scene()
    s.add_plane(2, 0.0).add_plane(2, 5.0)
    s.add_plane(0, -5.0).add_plane(0, 5.0).add_plane(1, -5.0).add_plane(1, 5.0)
    s.add_box((-1.0, -1.0, 1.0), (1.0, 1.0, 3.0))
    return s


def plane_scene(z=5.0):
    """A single plane facing an identity-pose camera (C1)."""
    return Scene().add_plane(2, z)


class PinholeCamera:
    def __init__(self, fu=300.0, fv=300.0, cu=320.0, cv=240.0, width=640, height=480):
        self.fu, self.fv, self.cu, self.cv = float(fu), float(fv), float(cu), float(cv)
        self.width, self.height = int(width), int(height)


def render_depth(scene, cam, T_S_C, max_dist=10.0, invalid_depth=0.0):
    """Depth image (rows, cols) float32 of `scene` seen from T_S_C (4x4)."""
    T = np.asarray(T_S_C, dtype=np.float64)
    R, t = T[:3, :3], T[:3, 3]
    cols = (np.arange(cam.width, dtype=np.float64) + 0.5 - cam.cu) / cam.fu
    rows = (np.arange(cam.height, dtype=np.float64) + 0.5 - cam.cv) / cam.fv
    vx, vy = np.meshgrid(cols, rows)
    v = np.stack([vx, vy, np.ones_like(vx)], axis=-1)
    norm = np.linalg.norm(v, axis=-1, keepdims=True)
    d_C = v / norm
    d_S = d_C @ R.T
    dist = scene.ray_distance(t, d_S, max_dist)
    depth = dist * d_C[..., 2]
    return np.where(np.isfinite(depth), depth, invalid_depth).astype(np.float32)


def _quat_to_rot(w, x, y, z):
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def circle_pose(theta, radius=4.0, height=2.0, yaw_offset=0.0):
    """T_S_C on the test circle: camera z axis points at the vertical axis."""
    base = _quat_to_rot(0.5, 0.5, 0.5, 0.5)
    a = math.pi + theta + yaw_offset
    rz = np.array([[math.cos(a), -math.sin(a), 0.0], [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]])
    T = np.eye(4)
    T[:3, :3] = rz @ base
    T[:3, 3] = [radius * math.cos(theta), radius * math.sin(theta), height]
    return T.astype(np.float32)


def circle_trajectory(num_poses=80, radius=4.0, height=2.0, yaw_offset=0.0):
    inc = 2.0 * math.pi / num_poses
    return [circle_pose(inc * i, radius, height, yaw_offset) for i in range(num_poses)]


def make_sequence(scene, cam, poses, max_dist=10.0, noise_sigma_rel=0.0, dropout=0.0, seed=0):
    """List of (depth float32 (H,W), T_L_C float32 4x4). Optional N(0,(s*z)^2) noise and
    Bernoulli zero-dropout (3DMatch-like 13 % invalid pixels), both seeded."""
    rng = np.random.default_rng(seed)
    out = []
    for T in poses:
        d = render_depth(scene, cam, T, max_dist)
        if noise_sigma_rel > 0:
            d = (d + rng.normal(0.0, 1.0, d.shape).astype(np.float32) * noise_sigma_rel * d).astype(np.float32)
        if dropout > 0:
            d = np.where(rng.random(d.shape) < dropout, np.float32(0), d).astype(np.float32)
        out.append((np.ascontiguousarray(d), np.asarray(T, dtype=np.float32)))
    return out


def moving_sphere_sequence(cam, poses, step_m=0.05, max_dist=10.0):
    """C5: sphere_in_box plus a small sphere translating step_m per frame, with its
    image-space mask (uint8, 255 where the moving sphere is the nearest hit)."""
    out = []
    for i, T in enumerate(poses):
        static = sphere_in_box()
        mover_c = np.array([-3.0 + step_m * i, 2.5, 1.0])
        d_static = render_depth(static, cam, T, max_dist)
        d_mover = render_depth(Scene().add_sphere(mover_c, 0.5), cam, T, max_dist)
        hit = (d_mover > 0) & ((d_static <= 0) | (d_mover < d_static))
        depth = np.where(hit, d_mover, d_static).astype(np.float32)
        mask = np.where(hit, 255, 0).astype(np.uint8)
        out.append((np.ascontiguousarray(depth), np.asarray(T, dtype=np.float32), np.ascontiguousarray(mask)))
    return out
