"""Shared test data.

`FOOTPRINTS` are the reference's own test polygons (test/cover_test_utils.hpp:17-48): all x
coordinates followed by all y coordinates.  They are test VECTORS (numbers), reused so that the
oracle is pinned on the same inputs as the reference's gtest suite.
"""
import numpy as np

_RAW = {
    "S": [-2.5, -1, 1, 2.5, 1, -1, 1, -0.5, 1.5, -1, 0.5, -1.5],
    "W": [-2, -2, -1, -1, 1, 1, 2, 2, 2, -2, -2, -1, -1, -2, -2, 2],
    "M": [2, 2, 0, -2, -2, -1.5, -1.5, 0, 1.5, 1.5, -2, 2, 0, 2, -2, -2, 1, -0.5, 1, -2],
    "Spiral": [0, 0.05, 0, -0.2, 0, 0.5, 0, -1.3, 0, 3.4, 0, 6.8, 0, -2.6, 0, 1, 0, -0.4, 0, 0.15,
               0, 0.05, 0.1, 0, -0.3, 0, 0.8, 0, -2.1, 0, 5.5, 0, -4.2, 0, 1.6, 0, -0.6, 0, 0.2, 0.05],
    "SOTO": [0.41, 1.19, 1.19, 1.19, 1.19, 0.41, -0.35, -1.13, -1.16, -1.22, -1.27, -1.33, -1.33, -1.33, -1.33,
             -1.33, -1.33, -1.27, -1.22, -1.16, -1.13, -0.35, -0.65, -0.65, -0.21, 0.21, 0.65, 0.65, 0.65, 0.65,
             0.58, 0.58, 0.58, 0.58, 0.35, 0.11, -0.11, -0.35, -0.58, -0.58, -0.58, -0.58, -0.65, -0.65],
    "TORU": [0.84, 0.84, 0.84, 0.83, 0.81, 0.76, 0.72, 0.66, 0.59, 0.51, -0.20, -0.29, -0.38, -0.45, -0.51, -0.52,
             -0.52, -0.52, -0.51, -0.45, -0.38, -0.29, -0.20, 0.51, 0.59, 0.66, 0.72, 0.76, 0.81, 0.83, 0.84, 0.84,
             0.00, 0.00, 0.03, 0.04, 0.13, 0.21, 0.26, 0.29, 0.32, 0.34, 0.34, 0.32, 0.27, 0.19, 0.10, 0.03, 0.00,
             -0.03, -0.10, -0.19, -0.27, -0.32, -0.34, -0.34, -0.32, -0.29, -0.26, -0.21, -0.13, -0.04, -0.03, -0.00],
    "L": [0.51, -0.51, -0.41, -0.31, -0.31, 0.51, 0.11, -0.11, -0.81, -0.76, -0.26, -0.11],
}


def make_footprint(coords):
    """test/cover_test_utils.hpp:58-70: first half = x row, second half = y row -> [2, V]."""
    a = np.asarray(coords, dtype=np.float64)
    return a.reshape(2, -1).copy()


FOOTPRINT_NAMES = ["S", "W", "M", "Spiral", "SOTO", "TORU", "L"]
FOOTPRINTS = [make_footprint(_RAW[k]) for k in FOOTPRINT_NAMES]

# test/expand.cpp:138-143 ("simple_footprints")
SIMPLE_FOOTPRINTS = [
    make_footprint([1, -1, -1, 0, 1, -1]),
    make_footprint([1, 1, -1, -1, -1, 1, 1, -1]),
    make_footprint([1, 1, 0, -1, -1, -1, 1, 0.5, 1, -1]),
    FOOTPRINTS[4],
    FOOTPRINTS[5],
]


def head_polygons():
    """test/expand.cpp:241-279 ("create_head"): a face with nested holes, shifted to the positive quadrant."""
    head = np.array([[-1, 1, 1.5, 1.5, -1.5, -1.5], [-2, -2, 0, 2, 2, 0]], dtype=np.float64)
    eye_left = np.array([[-1, -0.2, -0.2, -1], [0.6, 0.6, 1, 1]], dtype=np.float64)
    eye_right = np.array([[1, 0.2, 0.2, 1], [0.6, 0.6, 1, 1]], dtype=np.float64)
    ball_left = np.array([[-0.6, -0.5, -0.6, -0.7], [0.7, 0.8, 0.9, 0.8]], dtype=np.float64)
    ball_right = np.array([[0.6, 0.5, 0.6, 0.7], [0.7, 0.8, 0.9, 0.8]], dtype=np.float64)
    nose = np.array([[0, -0.25, 0.25], [0.1, -0.65, -0.65]], dtype=np.float64)
    mouth = np.array([[-0.75, -0.5, 0.5, 0.75], [-1, -1.5, -1.5, -1]], dtype=np.float64)
    polys = [head, eye_left, eye_right, ball_left, ball_right, nose, mouth]
    orig = head.min(axis=1, keepdims=True)
    return [p - orig for p in polys]


RECT = np.array([[0.5, 0.5, -0.5, -0.5], [0.3, -0.3, -0.3, 0.3]], dtype=np.float64)  # 1.0 x 0.6 m
