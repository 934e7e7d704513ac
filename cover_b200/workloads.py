"""Synthetic costmaps, footprints and pose sets of the shapes BASELINE.json names.

Everything is generated on the HOST with numpy `default_rng(seed)` and handed identically to the CUDA
library and to the CPU oracle; trigonometry is evaluated here (numpy / glibc), never on the device.
"""
from __future__ import annotations

import math

import numpy as np

LETHAL, NO_INFORMATION, INSCRIBED = 254, 255, 253

# 1.0 x 0.6 m rectangular footprint (BASELINE.json north_star), [2, V]
RECT = np.array([[0.5, 0.5, -0.5, -0.5], [0.3, -0.3, -0.3, 0.3]], dtype=np.float64)


def potato(n_vertices: int = 48) -> np.ndarray:
    """Non-convex closed curve r(phi) = 0.55 + 0.12 cos 2phi + 0.08 sin 3phi + 0.05 cos 5phi, scaled (1.0, 0.7)."""
    phi = 2 * math.pi * np.arange(n_vertices) / n_vertices
    r = 0.55 + 0.12 * np.cos(2 * phi) + 0.08 * np.sin(3 * phi) + 0.05 * np.cos(5 * phi)
    return np.stack([1.0 * r * np.cos(phi), 0.7 * r * np.sin(phi)])


def scaled(polygon: np.ndarray, s: float) -> np.ndarray:
    c = polygon.mean(axis=1, keepdims=True)
    return (polygon - c) * s + c


def random_costmap(size_x: int, size_y: int, seed: int, lethal_p: float = 1e-3) -> np.ndarray:
    """Uniform costs 0..252 with each cell lethal w.p. lethal_p (C1 'sparse-lethal'; lethal_p=0 -> 'free')."""
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 253, (size_y, size_x), dtype=np.uint8)
    if lethal_p > 0:
        data[rng.random((size_y, size_x)) < lethal_p] = LETHAL
    return data


def inflated_costmap(size_x: int, size_y: int, resolution: float, seed: int, obstacle_p: float = 2e-4,
                     inscribed_radius: float = 0.3, inflation_radius: float = 1.0, cost_scaling: float = 10.0) -> np.ndarray:
    """Point obstacles (254) inflated like costmap_2d::InflationLayer::computeCost: 253 within the inscribed
    radius, (uint8)(252 * exp(-scaling * (d - inscribed))) out to the inflation radius, else 0.  The cost is
    monotone in the distance, so stamping the per-obstacle kernel with `maximum` equals the exact Euclidean
    distance transform formulation."""
    rng = np.random.default_rng(seed)
    n_obs = rng.binomial(size_x * size_y, obstacle_p)
    ox = rng.integers(0, size_x, n_obs)
    oy = rng.integers(0, size_y, n_obs)
    rad = int(math.ceil(inflation_radius / resolution))
    yy, xx = np.mgrid[-rad:rad + 1, -rad:rad + 1]
    d = np.hypot(xx, yy) * resolution
    kernel = np.zeros(d.shape, dtype=np.uint8)
    band = (d > inscribed_radius) & (d <= inflation_radius)
    kernel[band] = (252.0 * np.exp(-cost_scaling * (d[band] - inscribed_radius))).astype(np.uint8)
    kernel[d <= inscribed_radius] = INSCRIBED
    kernel[rad, rad] = LETHAL
    data = np.zeros((size_y, size_x), dtype=np.uint8)
    for x, y in zip(ox.tolist(), oy.tolist()):
        x0, x1 = max(0, x - rad), min(size_x, x + rad + 1)
        y0, y1 = max(0, y - rad), min(size_y, y + rad + 1)
        view = data[y0:y1, x0:x1]
        np.maximum(view, kernel[y0 - (y - rad):y1 - (y - rad), x0 - (x - rad):x1 - (x - rad)], out=view)
    return data


def inflation_lut(resolution: float, inscribed_radius: float = 0.3, inflation_radius: float = 1.0, cost_scaling: float = 10.0):
    """(radius_cells, cost_by_d2) for Costmap.inflate: InflationLayer::computeCost evaluated on the host for
    every squared cell distance 0..radius_cells^2 (0 beyond the inflation radius)."""
    rad = int(math.ceil(inflation_radius / resolution))
    d = np.sqrt(np.arange(rad * rad + 1, dtype=np.float64)) * resolution
    lut = np.zeros(rad * rad + 1, dtype=np.uint8)
    band = (d > inscribed_radius) & (d <= inflation_radius)
    lut[band] = (252.0 * np.exp(-cost_scaling * (d[band] - inscribed_radius))).astype(np.uint8)
    lut[d <= inscribed_radius] = INSCRIBED
    lut[0] = LETHAL
    return rad, lut


def random_poses(n: int, seed: int, lo: float, hi: float) -> np.ndarray:
    """[n, 4] = (cos, sin, tx, ty): position uniform in [lo, hi]^2, heading uniform in [-pi, pi)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(lo, hi, n)
    y = rng.uniform(lo, hi, n)
    th = rng.uniform(-math.pi, math.pi, n)
    return np.stack([np.cos(th), np.sin(th), x, y], axis=1)


def heading_table(n_theta: int) -> np.ndarray:
    th = -math.pi + 2 * math.pi * np.arange(n_theta) / n_theta
    return np.stack([np.cos(th), np.sin(th)], axis=1)


def dwa_lattice(nx: int = 1180, ny: int = 1180, n_theta: int = 72, x0: float = 20.0, y0: float = 20.0, pitch: float = 0.05):
    """C4: (x, y, theta) lattice at cell pitch; 1180 x 1180 x 72 = 100 252 800 poses by default."""
    return dict(x0=x0, dx=pitch, nx=nx, y0=y0, dy=pitch, ny=ny, cs=heading_table(n_theta))


def rotation(theta: float) -> np.ndarray:
    return np.array([[math.cos(theta), -math.sin(theta)], [math.sin(theta), math.cos(theta)]])
