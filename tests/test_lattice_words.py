"""GPU parity of the bit-parallel lattice verdict kernels (cover_b200/csrc/lattice_words.cuh: k_lw_classes,
k_lw_shapes, k_lw_sweep) against the CPU oracle, bit for bit: every reference footprint as area and as outline,
axis-aligned headings (vertices on cell boundaries: several shape classes per heading), lattices that hang over
the map border, sub-ranges that start and end inside a row, byte / bit-packed / device-resident outputs, all
strip widths, statuses.  The shape-cache kernel it replaces on this path stays covered through COVER_B200_NO_WORDS=1.
"""
import math

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.fixtures import FOOTPRINTS, FOOTPRINT_NAMES, RECT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import cover_b200
    return cover_b200


def _ctx(cb, monkeypatch, **env):
    monkeypatch.setenv("COVER_B200_PLANES_MIN_BATCH", "0")  # small test lattices take the plane-based kernels too
    for k, v in env.items():
        monkeypatch.setenv(k, str(v))
    return cb.Context(0)


def mixed_map(rng, sx, sy, lethal=0.002):
    data = rng.integers(0, 253, (sy, sx), dtype=np.uint8)
    r = rng.random((sy, sx))
    data[r < lethal] = 254
    data[(r >= lethal) & (r < 1.5 * lethal)] = 255
    data[(r >= 1.5 * lethal) & (r < 2.5 * lethal)] = 253
    return data


def headings(n, extra=(0.0, math.pi / 2, math.pi, -math.pi / 2)):
    th = np.concatenate([-math.pi + 2 * math.pi * np.arange(n) / n, np.asarray(extra, dtype=np.float64)])
    return np.stack([np.cos(th), np.sin(th)], 1)


def check(cmap, omap, cb, poly, desc, shape, what, first=0, count=None):
    fn = cmap.area_is_free if shape == po.SHAPE_AREA else cmap.outline_is_free
    gpu = fn(poly, cb.Lattice(**desc), first=first, count=count)
    cpu = po.polygon_batch(omap, poly, po.Lattice(**desc), shape=shape, first=first, count=count, threads=8)
    np.testing.assert_array_equal(gpu.status, cpu.status, err_msg=f"status {what}")
    np.testing.assert_array_equal(gpu.is_free, cpu.is_free, err_msg=f"is_free {what}")
    return cpu


@pytest.mark.parametrize("strip", [2, 4, 8])
def test_rect_and_reference_footprints(cb, monkeypatch, strip):
    """Rectangle + the reference's seven test footprints (test/cover_test_utils.hpp:17-48) at 5 cm: up to ~230 scan
    rows and 16 ranges per row (Spiral), areas and outlines, on a lattice that hangs over the map on every side."""
    ctx = _ctx(cb, monkeypatch, COVER_B200_WORDS_STRIP=strip)
    launches0 = ctx.launch_count
    rng = np.random.default_rng(501)
    data = mixed_map(rng, 400, 360, lethal=0.0008)
    cmap = cb.Costmap(ctx, data, 0.05, -1.0, 0.5)
    omap = po.Costmap(data, 0.05, -1.0, 0.5)
    cs = headings(5)
    polys = [("rect", RECT)] + list(zip(FOOTPRINT_NAMES, FOOTPRINTS))
    for name, poly in polys:
        big = name in ("W", "M", "S", "Spiral")
        desc = dict(x0=-1.6, dx=0.05, nx=70 if big else 300, y0=0.1, dy=0.05, ny=33 if big else 90, cs=cs)
        for shape in (po.SHAPE_AREA, po.SHAPE_OUTLINE):
            cpu = check(cmap, omap, cb, poly, desc, shape, f"{name} shape {shape} strip {strip}")
        assert cpu.is_free.size >= 4096
    assert ctx.launch_count > launches0
    ctx.close()


def test_axis_aligned_headings_many_classes(cb, monkeypatch):
    """Vertices exactly on cell boundaries: rounding noise in x0 + ix*dx flips their cells from column to column, so a
    heading has several x and y classes and a 32-pose word mixes shapes (masked walks)."""
    ctx = _ctx(cb, monkeypatch)
    rng = np.random.default_rng(502)
    data = mixed_map(rng, 420, 380, lethal=0.002)
    th = np.array([0.0, math.pi / 2, math.pi, -math.pi / 2, math.atan2(3, 4), 1e-9, -1e-12])
    cs = np.stack([np.cos(th), np.sin(th)], 1)
    for origin in ((0.0, 0.0), (-1.3, 0.65), (0.1, -0.2)):
        for x0 in (-1.0, 0.3, 0.7000000000000001):
            desc = dict(x0=x0 + origin[0], dx=0.05, nx=380, y0=origin[1] - 0.95, dy=0.05, ny=150, cs=cs)
            cmap = cb.Costmap(ctx, data, 0.05, *origin)
            omap = po.Costmap(data, 0.05, *origin)
            cpu = check(cmap, omap, cb, RECT, desc, po.SHAPE_AREA, f"origin {origin} x0 {x0}")
            check(cmap, omap, cb, RECT, desc, po.SHAPE_OUTLINE, f"outline origin {origin} x0 {x0}")
            assert 0.02 < cpu.is_free.mean() < 0.98
    ctx.close()


def test_sub_ranges_bits_and_device_outputs(cb, monkeypatch):
    """first / count anywhere inside rows (pose sharding), nx not a multiple of 4 (byte stores), bit-packed verdicts
    with the not-OK counter, device-resident byte outputs at an odd address."""
    import torch
    ctx = _ctx(cb, monkeypatch)
    rng = np.random.default_rng(503)
    data = mixed_map(rng, 512, 300, lethal=0.003)
    cmap = cb.Costmap(ctx, data, 0.05)
    omap = po.Costmap(data, 0.05)
    cs = headings(7)
    for nx in (331, 332, 64, 129):
        desc = dict(x0=0.4, dx=0.05, nx=nx, y0=0.5, dy=0.05, ny=77, cs=cs)
        total = nx * 77 * cs.shape[0]
        full = po.polygon_batch(omap, RECT, po.Lattice(**desc), threads=8)
        for first, count in ((0, total), (12_345, min(50_001, total - 12_345)), (nx * 3 + 1, 4096), (total - 5000, 5000), (32 * 400, 32 * 1000)):
            gpu = cmap.area_is_free(RECT, cb.Lattice(**desc), first=first, count=count)
            np.testing.assert_array_equal(gpu.is_free, full.is_free[first:first + count], err_msg=f"nx {nx} first {first} count {count}")
            np.testing.assert_array_equal(gpu.status, full.status[first:first + count])
            out = cb.Result(free_bits=np.zeros((count + 31) // 32, dtype=np.uint32), not_ok=np.zeros(1, dtype=np.int64))
            cmap.area_is_free(RECT, cb.Lattice(**desc), first=first, count=count, out=out)
            np.testing.assert_array_equal(out.unpack_bits(count), full.is_free[first:first + count], err_msg=f"bits nx {nx} first {first}")
            assert int(out.not_ok[0]) == 0
        # device-resident outputs, byte arrays starting at an odd address
        buf_f = torch.full((total + 3,), 7, dtype=torch.uint8, device="cuda")
        buf_s = torch.full((total + 3,), 7, dtype=torch.uint8, device="cuda")
        dev = cb.Result(buf_f[1:1 + total], None, buf_s[3:3 + total])
        cmap.area_is_free(RECT, cb.Lattice(**desc), out=dev)
        ctx.synchronize()
        np.testing.assert_array_equal(buf_f[1:1 + total].cpu().numpy(), full.is_free)
        np.testing.assert_array_equal(buf_s[3:3 + total].cpu().numpy(), full.status)
        assert int(buf_f[0]) == 7 and int(buf_f[total + 1]) == 7 and int(buf_s[2]) == 7
    ctx.close()


def test_statuses_on_a_lattice(cb, monkeypatch):
    """Open outlines (std::logic_error upstream), one-cell polygons, coordinates out of range."""
    ctx = _ctx(cb, monkeypatch)
    rng = np.random.default_rng(504)
    data = mixed_map(rng, 256, 256, lethal=0.004)
    cmap = cb.Costmap(ctx, data, 0.05)
    omap = po.Costmap(data, 0.05)
    cs = headings(8)
    desc = dict(x0=0.2, dx=0.05, nx=230, y0=0.3, dy=0.05, ny=100, cs=cs)
    polys = [np.array([[0.4, -0.3, -0.3], [0.0, 0.25, -0.25]]), np.array([[0.01], [0.01]]), np.array([[0.0, 0.3], [0.0, 0.45]]),
             np.array([[0.6, 0.6, -0.6, -0.6], [0.05, -0.05, -0.05, 0.05]]), np.array([[0.0, 0.4, 0.8], [0.0, 0.0, 0.0]])]
    seen = set()
    for poly in polys:
        for shape in (po.SHAPE_AREA, po.SHAPE_OUTLINE):
            cpu = check(cmap, omap, cb, poly, desc, shape, f"{poly.tolist()} shape {shape}")
            seen |= set(np.unique(cpu.status).tolist())
            out = cb.Result(free_bits=np.zeros((cpu.is_free.size + 31) // 32, dtype=np.uint32), not_ok=np.zeros(1, dtype=np.int64))
            (cmap.area_is_free if shape == po.SHAPE_AREA else cmap.outline_is_free)(poly, cb.Lattice(**desc), out=out)
            np.testing.assert_array_equal(out.unpack_bits(cpu.is_free.size), cpu.is_free)
            assert int(out.not_ok[0]) == int((cpu.status != 0).sum())
    assert cb.STATUS_LOGIC_ERROR in seen
    far = dict(x0=2.0e5, dx=0.05, nx=128, y0=0.3, dy=0.05, ny=40, cs=cs)  # x0 / res = 4e6 < 2^22 < (x0 + 6.4) / res for some columns
    far["x0"] = 4194304 * 0.05 - 3.0
    cpu = check(cmap, omap, cb, RECT, far, po.SHAPE_AREA, "coordinate range")
    assert cb.STATUS_COORD_RANGE in set(np.unique(cpu.status).tolist()) and cb.STATUS_OK in set(np.unique(cpu.status).tolist())
    ctx.close()


def test_non_unit_y_pitch_and_resolutions(cb, monkeypatch):
    """The x pitch must be one cell; the y pitch is free (rows carry their own anchor).  Other resolutions."""
    ctx = _ctx(cb, monkeypatch)
    rng = np.random.default_rng(505)
    cs = headings(6)
    for res, dy in ((0.05, 0.05), (0.05, 0.0173), (0.05, 0.11), (0.05, -0.05), (0.025, 0.025), (0.1, 0.1), (0.02, 0.02)):
        data = mixed_map(rng, 380, 330, lethal=0.002)
        cmap = cb.Costmap(ctx, data, res, -0.2, 0.1)
        omap = po.Costmap(data, res, -0.2, 0.1)
        desc = dict(x0=0.3, dx=res, nx=300, y0=(0.4 if dy > 0 else 300 * res), dy=dy, ny=64, cs=cs)
        for poly in (RECT, FOOTPRINTS[6], FOOTPRINTS[5]):
            check(cmap, omap, cb, poly, desc, po.SHAPE_AREA, f"res {res} dy {dy}")
    ctx.close()


def test_shape_cache_kernel_still_exact(cb, monkeypatch):
    """COVER_B200_NO_WORDS=1: the same sweeps through k_area_lattice (still the path of the value folds and of
    lattices whose x pitch is not one cell)."""
    ctx = _ctx(cb, monkeypatch, COVER_B200_NO_WORDS=1)
    rng = np.random.default_rng(506)
    data = mixed_map(rng, 400, 360, lethal=0.002)
    cmap = cb.Costmap(ctx, data, 0.05, -1.0, 0.5)
    omap = po.Costmap(data, 0.05, -1.0, 0.5)
    cs = headings(5)
    desc = dict(x0=-1.6, dx=0.05, nx=300, y0=0.1, dy=0.05, ny=90, cs=cs)
    for poly in (RECT, FOOTPRINTS[6], FOOTPRINTS[4]):
        check(cmap, omap, cb, poly, desc, po.SHAPE_AREA, "shape-cache kernel")
        check(cmap, omap, cb, poly, desc, po.SHAPE_OUTLINE, "shape-cache kernel, outline")
    ctx.close()


@pytest.mark.parametrize("env", [{}, {"COVER_B200_NO_WORDS": 1}])
def test_deferred_pose_list_overflow(cb, monkeypatch, env):
    """COVER_B200_TODO_CAP=64: every deferred-pose list overflows, so the drain kernels take the whole batch - concave
    footprints under random poses (the per-pose fast passes defer most of them, twice), a lattice over the map border
    (the shape-cache kernel defers the border poses) and outlines on it."""
    ctx = _ctx(cb, monkeypatch, COVER_B200_TODO_CAP=64, **env)
    rng = np.random.default_rng(507)
    data = mixed_map(rng, 300, 280, lethal=0.002)
    cmap = cb.Costmap(ctx, data, 0.05, -0.5, 0.25)
    omap = po.Costmap(data, 0.05, -0.5, 0.25)
    th = rng.uniform(-math.pi, math.pi, 30_000)
    poses = np.stack([np.cos(th), np.sin(th), rng.uniform(-1.0, 15.0, th.size), rng.uniform(-1.0, 14.0, th.size)], 1)
    for poly in (RECT, FOOTPRINTS[6], FOOTPRINTS[0] * 0.3, FOOTPRINTS[2] * 0.3):
        for name, op, cost in (("area_is_free", po.OP_IS_FREE, False), ("area_accumulate_cost", po.OP_ACCUMULATE, True)):
            gpu = getattr(cmap, name)(poly, poses)
            cpu = po.polygon_batch(omap, poly, poses, op=op, threads=8)
            np.testing.assert_array_equal(gpu.status, cpu.status)
            np.testing.assert_array_equal(gpu.is_free, cpu.is_free)
            if cost:
                np.testing.assert_array_equal(gpu.cost, cpu.cost)
    # the not-OK counter when the drain takes every pose: failures the fast pass already counted are not counted twice
    far = poses.copy()
    far[::7, 2] = 1e12  # coordinates out of range
    for poly in (FOOTPRINTS[0] * 0.3, FOOTPRINTS[3] * 0.2):
        out = cb.Result(free_bits=np.zeros((len(far) + 31) // 32, dtype=np.uint32), not_ok=np.zeros(1, dtype=np.int64))
        cmap.area_is_free(poly, far, out=out)
        cpu = po.polygon_batch(omap, poly, far, threads=8)
        assert int(out.not_ok[0]) == int((cpu.status != 0).sum()) > 0
        np.testing.assert_array_equal(out.unpack_bits(len(far)), cpu.is_free & (cpu.status == 0))
    desc = dict(x0=-1.2, dx=0.05, nx=330, y0=-0.6, dy=0.05, ny=150, cs=headings(5))
    for poly in (RECT, FOOTPRINTS[6]):
        check(cmap, omap, cb, poly, desc, po.SHAPE_AREA, f"overflowing list, area {env}")
        check(cmap, omap, cb, poly, desc, po.SHAPE_OUTLINE, f"overflowing list, outline {env}")
    ctx.close()


def test_table_cache_across_calls(cb, monkeypatch):
    """The class / shape tables are kept from one sweep to the next while polygon, lattice range, offset and map geometry
    stay the same - and only then: same sweep twice, a new costmap image in between (the tables do not depend on the cells),
    another polygon / lattice / sub-range / outline / offset, then the first again."""
    ctx = _ctx(cb, monkeypatch)
    rng = np.random.default_rng(508)
    images = [mixed_map(rng, 400, 360, lethal=0.002), mixed_map(rng, 400, 360, lethal=0.004)]
    cmap = cb.Costmap(ctx, images[0], 0.05, -1.0, 0.5)
    cs = headings(6)
    d1 = dict(x0=-0.3, dx=0.05, nx=300, y0=0.7, dy=0.05, ny=90, cs=cs)
    d2 = dict(x0=-0.3, dx=0.05, nx=300, y0=0.7000001, dy=0.05, ny=90, cs=cs)
    calls = [(RECT, d1, po.SHAPE_AREA, 0, None, None), (RECT, d1, po.SHAPE_AREA, 0, None, None), (RECT, d1, po.SHAPE_OUTLINE, 0, None, None),
             (FOOTPRINTS[6], d1, po.SHAPE_AREA, 0, None, None), (RECT, d2, po.SHAPE_AREA, 0, None, None), (RECT, d1, po.SHAPE_AREA, 27000 * 2 + 5, 60000, None),
             (RECT, d1, po.SHAPE_AREA, 0, None, (3, -2)), (RECT, d1, po.SHAPE_AREA, 0, None, None)]
    for rep in range(2):
        for k, (poly, desc, shape, first, count, offset) in enumerate(calls):
            image = images[(k + rep) % 2]
            cmap.upload(image)
            omap = po.Costmap(image, 0.05, -1.0, 0.5)
            fn = cmap.area_is_free if shape == po.SHAPE_AREA else cmap.outline_is_free
            launches = ctx.launch_count
            gpu = fn(poly, cb.Lattice(**desc), first=first, count=count, offset=offset)
            cpu = po.polygon_batch(omap, poly, po.Lattice(**desc), shape=shape, first=first, count=count, threads=8, offset=offset)
            np.testing.assert_array_equal(gpu.is_free, cpu.is_free, err_msg=f"call {k} rep {rep}")
            np.testing.assert_array_equal(gpu.status, cpu.status)
            if rep == 0 and k == 0:
                first_launches = ctx.launch_count - launches
            if rep == 0 and k == 1:  # the very same sweep again: no k_lw_prepare launch
                assert ctx.launch_count - launches == first_launches - 1
    ctx.close()


@pytest.mark.parametrize("env", [{}, {"COVER_B200_NO_WORDS": 1}, {"COVER_B200_TODO_CAP": 64}])
def test_packed_verdicts_do_not_depend_on_what_the_words_held(cb, monkeypatch, env):
    """Device-resident packed verdicts and not-OK counter pre-filled with ones: a sweep leaves exactly the verdicts of its
    poses (words shared by two tasks, rows that start inside a word, deferred and failed poses, the unused high bits of
    the last word), and nothing outside its words."""
    import torch
    ctx = _ctx(cb, monkeypatch, **env)
    rng = np.random.default_rng(509)
    data = mixed_map(rng, 300, 260, lethal=0.003)
    cmap = cb.Costmap(ctx, data, 0.05, -0.4, 0.3)
    omap = po.Costmap(data, 0.05, -0.4, 0.3)
    cs = headings(5)
    for poly, shape in ((RECT, po.SHAPE_AREA), (FOOTPRINTS[1], po.SHAPE_AREA), (RECT, po.SHAPE_OUTLINE), (np.array([[0.0, 1.0], [0.0, 0.4]]), po.SHAPE_AREA)):
        desc = dict(x0=-1.2, dx=0.05, nx=337, y0=-0.2, dy=0.05, ny=70, cs=cs)  # hangs over the map: COORD_RANGE never, deferred walks at the border
        total = 337 * 70 * cs.shape[0]
        full = po.polygon_batch(omap, poly, po.Lattice(**desc), shape=shape, threads=8)
        for first, count in ((0, total), (1000, 50_001), (337 * 5 + 17, 4099)):
            n_words = (count + 31) // 32
            bits = torch.full((n_words + 2,), -1, dtype=torch.int32, device="cuda")
            not_ok = torch.full((1,), 12345, dtype=torch.int64, device="cuda")
            out = cb.Result(free_bits=bits[1:1 + n_words], not_ok=not_ok)
            fn = cmap.area_is_free if shape == po.SHAPE_AREA else cmap.outline_is_free
            fn(poly, cb.Lattice(**desc), first=first, count=count, out=out)
            ctx.synchronize()
            words = bits.cpu().numpy().view(np.uint32)
            assert words[0] == 0xFFFFFFFF and words[-1] == 0xFFFFFFFF
            got = np.unpackbits(words[1:1 + n_words].view(np.uint8), bitorder="little")
            want = full.is_free[first:first + count] & (full.status[first:first + count] == 0)
            np.testing.assert_array_equal(got[:count], want, err_msg=f"first {first} count {count} env {env}")
            assert not got[count:].any()  # the unused bits of the last word are zero
            assert int(not_ok.item()) == int((full.status[first:first + count] != 0).sum())
    ctx.close()


@pytest.mark.parametrize("env", [{}, {"COVER_B200_TODO_CAP": 64}])
def test_value_folds_on_a_lattice(cb, monkeypatch, env):
    """accumulate_cost / max_cost of areas over a lattice (k_lw_values: prefix words, sliding-maximum planes) against the
    oracle: every reference footprint, a lattice that hangs over the map (deferred border poses), axis-aligned headings
    (several classes per row of poses), sentinel order (LETHAL before / after NO_INFORMATION), sub-ranges, a costmap
    update in between."""
    ctx = _ctx(cb, monkeypatch, **env)
    launches0 = ctx.launch_count
    rng = np.random.default_rng(511)
    data = mixed_map(rng, 420, 380, lethal=0.004)
    cmap = cb.Costmap(ctx, data, 0.05, -1.0, 0.5)
    omap = po.Costmap(data, 0.05, -1.0, 0.5)
    cs = headings(5)
    polys = [("rect", RECT)] + list(zip(FOOTPRINT_NAMES, FOOTPRINTS))
    for name, poly in polys:
        big = name in ("W", "M", "S", "Spiral")
        desc = dict(x0=-1.6, dx=0.05, nx=75 if big else 310, y0=0.1, dy=0.05, ny=33 if big else 90, cs=cs)
        total = desc["nx"] * desc["ny"] * cs.shape[0]
        for op_name, op in (("area_accumulate_cost", po.OP_ACCUMULATE), ("area_max_cost", po.OP_MAX)):
            for first, count in ((0, total), (desc["nx"] * 3 + 7, min(total - desc["nx"] * 3 - 7, 9001))):
                gpu = getattr(cmap, op_name)(poly, cb.Lattice(**desc), first=first, count=count)
                cpu = po.polygon_batch(omap, poly, po.Lattice(**desc), op=op, first=first, count=count, threads=8)
                np.testing.assert_array_equal(gpu.status, cpu.status, err_msg=f"{name} {op_name} {env}")
                np.testing.assert_array_equal(gpu.cost, cpu.cost, err_msg=f"{name} {op_name} {env}")
                np.testing.assert_array_equal(gpu.is_free, cpu.is_free, err_msg=f"{name} {op_name} {env}")
    # a new image: the prefix words and maximum planes are rebuilt
    data2 = mixed_map(rng, 420, 380, lethal=0.01)
    cmap.upload(data2)
    omap2 = po.Costmap(data2, 0.05, -1.0, 0.5)
    desc = dict(x0=-0.5, dx=0.05, nx=300, y0=0.6, dy=0.05, ny=120, cs=headings(3))
    for op_name, op in (("area_accumulate_cost", po.OP_ACCUMULATE), ("area_max_cost", po.OP_MAX)):
        gpu = getattr(cmap, op_name)(RECT, cb.Lattice(**desc))
        cpu = po.polygon_batch(omap2, RECT, po.Lattice(**desc), op=op, threads=8)
        np.testing.assert_array_equal(gpu.cost, cpu.cost)
        np.testing.assert_array_equal(gpu.status, cpu.status)
        assert (cpu.cost < 0).any() and (cpu.cost > 0).any()
    # a window update repairs them in place
    patch = mixed_map(rng, 64, 50, lethal=0.05)
    data2[90:140, 100:164] = patch
    cmap.upload_window(data2, 100, 90, 64, 50)
    omap3 = po.Costmap(data2, 0.05, -1.0, 0.5)
    for op_name, op in (("area_accumulate_cost", po.OP_ACCUMULATE), ("area_max_cost", po.OP_MAX)):
        gpu = getattr(cmap, op_name)(RECT, cb.Lattice(**desc))
        cpu = po.polygon_batch(omap3, RECT, po.Lattice(**desc), op=op, threads=8)
        np.testing.assert_array_equal(gpu.cost, cpu.cost)
    assert ctx.launch_count > launches0
    ctx.close()


@pytest.mark.parametrize("size", [(400, 360), (413, 300), (1040, 200)])
def test_device_resident_images(cb, monkeypatch, size):
    """cover_map_upload from DEVICE memory (k_copy16 for big 16-byte-aligned images, a driver copy otherwise) invalidates
    the planes: sweeps after such uploads - over the same rows, over other rows, after a second new image -
    agree with the oracle on the image they were given."""
    import torch
    ctx = _ctx(cb, monkeypatch)
    rng = np.random.default_rng(520)
    sx, sy = size
    images = [mixed_map(rng, sx, sy, lethal=l) for l in (0.002, 0.006, 0.0005)]
    cmap = cb.Costmap(ctx, images[0], 0.05)
    cs = headings(3)
    windows = [dict(x0=1.0, dx=0.05, nx=min(300, sx - 60), y0=2.0, dy=0.05, ny=100, cs=cs),
               dict(x0=0.4, dx=0.05, nx=min(300, sx - 60), y0=6.0, dy=0.05, ny=120, cs=cs)]
    check(cmap, po.Costmap(images[0], 0.05), cb, RECT, windows[0], po.SHAPE_AREA, "first image")
    for k, (img, win) in enumerate(((images[1], windows[0]), (images[2], windows[1]), (images[0], windows[1]), (images[1], windows[0]))):
        dev = torch.from_numpy(img).cuda()  # (kept alive until the checks below have returned their results)
        cmap.upload(dev)
        check(cmap, po.Costmap(img, 0.05), cb, RECT, win, po.SHAPE_AREA, f"device image {k}")
        check(cmap, po.Costmap(img, 0.05), cb, FOOTPRINTS[6], win, po.SHAPE_OUTLINE, f"device image {k}, outline")
    ctx.close()


@pytest.mark.parametrize("chunks", [2, 3, 4])
def test_chunked_download_of_a_lattice_sweep(cb, monkeypatch, chunks):
    """The pipelined path of big lattice sweeps with host results (chunk k's results travel back while chunk k + 1 is
    computed; default from 2^28 poses) forced onto a small lattice: byte, bit-packed and status outputs and the not-OK
    counter across chunk borders, for verdicts and for a value fold."""
    ctx = _ctx(cb, monkeypatch, COVER_B200_PIPELINE_MIN=20000, COVER_B200_CHUNKS=chunks)
    rng = np.random.default_rng(530 + chunks)
    data = mixed_map(rng, 400, 300, lethal=0.003)
    cmap, omap = cb.Costmap(ctx, data, 0.05, -0.5, 0.2), po.Costmap(data, 0.05, -0.5, 0.2)
    desc = dict(x0=-1.0, dx=0.05, nx=333, y0=-0.3, dy=0.05, ny=97, cs=headings(4))
    total = 333 * 97 * 8
    for poly, shape in ((RECT, po.SHAPE_AREA), (np.array([[0.0, 1.0], [0.0, 0.4]]), po.SHAPE_AREA), (FOOTPRINTS[6], po.SHAPE_OUTLINE)):
        cpu = po.polygon_batch(omap, poly, po.Lattice(**desc), shape=shape, threads=8)
        fn = cmap.area_is_free if shape == po.SHAPE_AREA else cmap.outline_is_free
        gpu = fn(poly, cb.Lattice(**desc))
        np.testing.assert_array_equal(gpu.status, cpu.status)
        np.testing.assert_array_equal(gpu.is_free, cpu.is_free)
        out = cb.Result(free_bits=np.zeros((total + 31) // 32, dtype=np.uint32), not_ok=np.zeros(1, dtype=np.int64))
        fn(poly, cb.Lattice(**desc), out=out)
        np.testing.assert_array_equal(out.unpack_bits(total), cpu.is_free & (cpu.status == 0))
        assert int(out.not_ok[0]) == int((cpu.status != 0).sum())
        first, count = 333 * 5 + 11, 150_001  # a sub-range whose chunks start inside rows
        sub = fn(poly, cb.Lattice(**desc), first=first, count=count)
        np.testing.assert_array_equal(sub.is_free, cpu.is_free[first:first + count])
    acc = cmap.area_accumulate_cost(RECT, cb.Lattice(**desc))
    cpu = po.polygon_batch(omap, RECT, po.Lattice(**desc), op=po.OP_ACCUMULATE, threads=8)
    np.testing.assert_array_equal(acc.cost, cpu.cost)
    ctx.close()


def test_async_host_results(cb, monkeypatch):
    """COVER_MEM_HOST_ASYNC: results into page-locked host arrays without waiting - two contexts kept in flight by one
    thread (banded upload + sweep each), valid after Context.synchronize(); verdict bits, byte outputs and the counter."""
    import torch
    rng = np.random.default_rng(540)
    data = [mixed_map(rng, 500, 400, lethal=l) for l in (0.002, 0.005)]
    desc = dict(x0=0.5, dx=0.05, nx=410, y0=0.7, dy=0.05, ny=130, cs=headings(4))
    total = 410 * 130 * 8
    n_words = (total + 31) // 32
    slots = []
    for img in data:
        ctx = _ctx(cb, monkeypatch)
        pinned_img = torch.from_numpy(img).pin_memory()
        bits = torch.full((n_words,), -1, dtype=torch.int32).pin_memory()
        nok = torch.full((1,), 77, dtype=torch.int64).pin_memory()
        free = torch.full((total,), 9, dtype=torch.uint8).pin_memory()
        status = torch.full((total,), 9, dtype=torch.uint8).pin_memory()
        slots.append(dict(ctx=ctx, cmap=cb.Costmap(ctx, np.zeros_like(img), 0.05), img=pinned_img, bits=bits, nok=nok, free=free, status=status))
    for rep in range(3):
        for s in slots:  # queue both cycles, then wait
            s["cmap"].upload_async(s["img"].numpy())
            s["cmap"].area_is_free(RECT, cb.Lattice(**desc), out=cb.Result(free_bits=s["bits"].numpy().view(np.uint32), not_ok=s["nok"].numpy(), host_async=True))
        for s in slots:
            s["ctx"].synchronize()
    for s in slots:
        s["cmap"].outline_is_free(FOOTPRINTS[6], cb.Lattice(**desc), out=cb.Result(s["free"].numpy(), None, s["status"].numpy(), host_async=True))
    for s, img in zip(slots, data):
        s["ctx"].synchronize()
        omap = po.Costmap(img, 0.05)
        cpu = po.polygon_batch(omap, RECT, po.Lattice(**desc), threads=8)
        got = np.unpackbits(s["bits"].numpy().view(np.uint8), bitorder="little")[:total]
        np.testing.assert_array_equal(got, cpu.is_free)
        assert int(s["nok"][0]) == 0
        cpu_o = po.polygon_batch(omap, FOOTPRINTS[6], po.Lattice(**desc), shape=po.SHAPE_OUTLINE, threads=8)
        np.testing.assert_array_equal(s["free"].numpy(), cpu_o.is_free)
        np.testing.assert_array_equal(s["status"].numpy(), cpu_o.status)
        s["ctx"].close()
