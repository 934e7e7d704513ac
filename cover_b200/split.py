"""Grid split of a footprint into skeleton-outline cells and remainder-area cells.

The reference splits the CONTINUOUS polygon with boost::geometry (erode by r, dilate the eroded ring by r,
subtract: src/cover/footprints.cpp:79-125) and then discretises per heading (generator_footprint::init /
discrete_footprint::init, footprints.cpp:198-238).  Boost is not available here and the reference's own
tests only pin properties of the split, so this module builds the per-heading `discrete_footprint`
directly ON THE GRID, from the exact `to_area` cells the CUDA library produces (`Costmap.area_cells`):

    F  = to_area(resolution, R(theta) * polygon)                   (exact, device-side expand)
    D  = Euclidean distance (metres, cell centres) to the nearest cell outside F
    E  = {c in F : D(c) > r}                                        the always-inscribed core
    dE = cells of E with a 4-neighbour outside E                    -> OUTLINE cells (checked for 253 v 254)
    A  = {c in F : dist(c, dE) > r}                                 -> AREA cells (checked for 254)

Safety argument (the reference's regression test, test/footprints.cpp:205-242): a lethal cell p inside F is
either in A (found by the area scan) or within r of some q in dE; an inflation layer whose inscribed radius
is >= r then marks q with 253 or 254, so the outline scan finds it — no miss by construction.  No false
alarm from outside: every q in dE is farther than r from every cell outside F.

This is a host-side precompute ("the polygon split precomputed once on the host per footprint"); it is
NOT bit-parity with boost's polygons and does not claim to be.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np


def split_cells(area_cells: np.ndarray, resolution: float, radius: float) -> Tuple[np.ndarray, np.ndarray]:
    """Splits one footprint's cell set F ([N, 2] int32, any order) into (outline_cells, area_cells)."""
    if len(area_cells) == 0:
        return np.zeros((0, 2), np.int32), np.zeros((0, 2), np.int32)
    if not (radius > 0):
        raise ValueError("Radius must be positive")  # footprints.cpp:86-87
    cells = np.unique(np.asarray(area_cells, dtype=np.int64), axis=0)
    pad = int(math.ceil(radius / resolution)) + 2
    x0, y0 = cells[:, 0].min() - pad, cells[:, 1].min() - pad
    w, h = cells[:, 0].max() - x0 + pad + 1, cells[:, 1].max() - y0 + pad + 1
    inside = np.zeros((h, w), dtype=bool)
    inside[cells[:, 1] - y0, cells[:, 0] - x0] = True

    def dist_to(mask_true: np.ndarray) -> np.ndarray:
        """Exact Euclidean distance (in cells) from every grid cell to the nearest True cell of `mask_true`."""
        ty, tx = np.nonzero(mask_true)
        if len(ty) == 0:
            return np.full(mask_true.shape, np.inf)
        yy, xx = np.mgrid[0:h, 0:w]
        best = np.full((h, w), np.inf)
        for s in range(0, len(ty), 256):  # chunked brute force: footprints are at most a few 10^4 cells
            d2 = (yy[..., None] - ty[s:s + 256]) ** 2 + (xx[..., None] - tx[s:s + 256]) ** 2
            best = np.minimum(best, d2.min(axis=-1))
        return np.sqrt(best)

    d_out = dist_to(~inside) * resolution
    core = inside & (d_out > radius)
    if not core.any():  # radius too large: the whole footprint is scanned (footprints test "too_large_radius")
        return np.zeros((0, 2), np.int32), cells.astype(np.int32)
    nb = np.zeros_like(core)
    nb[1:, :] |= ~core[:-1, :]
    nb[:-1, :] |= ~core[1:, :]
    nb[:, 1:] |= ~core[:, :-1]
    nb[:, :-1] |= ~core[:, 1:]
    ring = core & nb
    d_ring = dist_to(ring) * resolution
    area = inside & (d_ring > radius)
    oy, ox = np.nonzero(ring)
    ay, ax = np.nonzero(area)
    return (np.stack([ox + x0, oy + y0], 1).astype(np.int32), np.stack([ax + x0, ay + y0], 1).astype(np.int32))


def heading_bank(cmap, polygon, radius: float, n_headings: int) -> Tuple[List[Tuple[np.ndarray, np.ndarray]], np.ndarray]:
    """Per-heading split footprints for `FootprintBank`: heading k = -pi + 2*pi*k/n.  The footprint cells of
    every heading come from the device (`cmap.area_cells`, exact to_area semantics) with the polygon rotated
    about the map origin cell, i.e. cell offsets are relative to the cell of the robot's reference point.
    Returns (bank, headings[n, 2] = (cos, sin))."""
    th = -math.pi + 2 * math.pi * np.arange(n_headings) / n_headings
    cs = np.stack([np.cos(th), np.sin(th)], 1)
    poses = np.concatenate([cs, np.full((n_headings, 1), cmap.origin_x), np.full((n_headings, 1), cmap.origin_y)], axis=1)
    cells, starts, status = cmap.area_cells(polygon, poses)
    if (status != 0).any():
        raise RuntimeError("Even number of line-pairs expected")  # generators.cpp:246-249
    bank = []
    for k in range(n_headings):
        bank.append(split_cells(cells[starts[k]:starts[k + 1]], cmap.resolution, radius))
    return bank, cs
