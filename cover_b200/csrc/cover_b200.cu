// cover_b200.cu — kernels and the C ABI (include/cover_b200.h) of the batched footprint checker.
//
// The costmap (<= 64 MB) lives in the 126 MB L2; HBM only sees the first touch.  Kernels:
//   k_area_lattice<OP>   pose-lattice checks: warp = 64 x-adjacent poses, separable exact vertex profiles,
//                        per-CTA shape cache, verdicts from precomputed lethal bit planes (k_lethal_planes), value
//                        folds from per-row prefix / sliding-max tables (the headline path)
//   k_area_pair<OP>      area checks, arbitrary poses, rows with <= 2 ranges: thread = pose, the warp's 32 outline
//                        walks fused in lock-step, 64-bit pair table in shared memory
//   k_area_quad<OP>      the same walk for non-convex polygons up to 250 cells wide: four 8+8-bit ranges per (row, thread)
//                        in the same 64-bit word, one pass
//   k_area_slots<OP,4>   second per-thread pass over the poses k_area_pair deferred: 4 ranges/row (COVER_B200_NO_QUAD=1, A/B only)
//   k_area_rows<OP>      area checks, any polygon <= 256 vertices: warp = pose, lane = scan row (area_rows.cuh);
//                        also drains the poses the two kernels above defer;  k_area_rows_multi<OP>: holes
//   k_area_list<OP>      area checks for > 256 vertices / many short edges: per-thread sorted lists in global scratch
//   k_area_packed<OP>    first-generation per-thread kernel (COVER_B200_OLD_PACKED=1, A/B measurements only)
//   k_outline<OP>        outline_generator checks   (a5 + a10/a13)
//   k_rays<OP>           ray_generator checks       (a4)
//   k_cells_free / k_cells<OP>   cached discrete_polygon x offsets (a11): row spans / ordered list
//   k_footprints         discrete/generator footprint bank x offsets (a15/a16): host-precomputed row spans
//   k_split_lockstep     continuous_footprint::is_free (a14): thread = pose, the warp's walks fused in lock-step
//   k_split              the same with independent per-thread walks (what the lock-step pass defers)
//   k_area_expand, k_outline_expand   to_area / to_outline, batched
//   k_trajectories       trajectory scoring (8f-4);  k_inflate: inflation layer (8f-3)
// There is no CPU path in this file: without a CUDA device every entry returns COVER_E_CUDA.
#include "../../include/cover_b200.h"
#include "area_rows.cuh"
#include "device_core.cuh"
#include "lattice_words.cuh"
#include "polygon_split.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace cvr;

namespace {

constexpr int kBlock = 128;
constexpr int kPackedMaxRows = 96;      // 96 rows x 128 threads x 4 B = 48 KB of shared memory
constexpr int kPackedMaxWidth = 32000;  // 15-bit relative x

// ===================================================================================================
// Kernels
// ===================================================================================================

/// Copies the polygon (column-major xy) into shared memory; returns the pointer.
__device__ __forceinline__ double* stage_polygon(const double* __restrict__ poly, int nv, double* smem) {
  for (int i = threadIdx.x; i < 2 * nv; i += blockDim.x) smem[i] = poly[i];
  __syncthreads();
  return smem;
}

template <int OP>
struct CellSink {
  Fold<OP> fold;
  const MapView& m;
  int ox, oy;
  __device__ __forceinline__ CellSink(const MapView& mv, int offx = 0, int offy = 0) : m(mv), ox(offx), oy(offy) {}
  __device__ __forceinline__ bool cell(int x, int y) { return fold.cell(m, x + ox, y + oy); }
};

template <int OP>
__device__ __forceinline__ void outline_pose(const MapView& m, const double* __restrict__ poly, int nv, const PoseSource& src, long long i,
                                             const OutView& out) {
  const Pose T = load_pose(src, i);
  int x0, y0, x1, y1;
  bool ok;
  CellSink<OP> sink(m);
  if (src.format == POSE_NONE) {
    ok = polygon_bbox<XF_NONE>(poly, nv, T, m, x0, y0, x1, y1);
    if (ok) walk_outline<XF_NONE>(poly, nv, T, m, sink);
  } else {
    ok = polygon_bbox<XF_PATH_A>(poly, nv, T, m, x0, y0, x1, y1);
    if (ok) walk_outline<XF_PATH_A>(poly, nv, T, m, sink);
  }
  if (!ok) {
    store_failure(out, i, ST_COORD_RANGE);
    return;
  }
  sink.fold.store(out, i);
  if (out.status) out.status[i] = ST_OK;
}

template <int OP>
__global__ void __launch_bounds__(kBlock) k_outline(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src,
                                                   long long count, OutView out) {
  extern __shared__ double smem_d[];
  const double* poly = stage_polygon(poly_g, nv, smem_d);
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count) outline_pose<OP>(m, poly, nv, src, i, out);
}

/// The same over the poses an outline sweep of the lattice kernel deferred (persistent grid over the todo list).
template <int OP>
__global__ void __launch_bounds__(kBlock) k_outline_drain(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src, long long count,
                                                         OutView out, const unsigned int* __restrict__ todo,
                                                         const unsigned int* __restrict__ todo_count) {
  extern __shared__ double smem_d[];
  const double* poly = stage_polygon(poly_g, nv, smem_d);
  bool listed;
  const long long n = todo_items(todo_count, count, listed), stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long j = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += stride)
    outline_pose<OP>(m, poly, nv, src, listed ? static_cast<long long>(todo[j]) : j, out);
}

template <int OP>
__global__ void __launch_bounds__(kBlock) k_area_packed(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src,
                                                       long long count, OutView out, int rows_cap, unsigned int* __restrict__ todo,
                                                       unsigned int* __restrict__ todo_count) {
  extern __shared__ double smem_d[];
  const double* poly = stage_polygon(poly_g, nv, smem_d);
  uint32_t* table = reinterpret_cast<uint32_t*>(smem_d + 2 * nv);
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const Pose T = load_pose(src, i);
  const bool none = src.format == POSE_NONE;
  int x0, y0, x1, y1;
  const bool ok = none ? polygon_bbox<XF_NONE>(poly, nv, T, m, x0, y0, x1, y1) : polygon_bbox<XF_PATH_A>(poly, nv, T, m, x0, y0, x1, y1);
  if (!ok) {
    store_failure(out, i, ST_COORD_RANGE);
    return;
  }
  if (nv == 0) {  // empty generator: nothing is lethal, the sum is 0
    Fold<OP> fold;
    fold.store(out, i);
    if (out.status) out.status[i] = ST_OK;
    return;
  }
  const int rows = y1 - y0 + 1;
  bool defer = rows > rows_cap || (x1 - x0) > kPackedMaxWidth;
  PackedStore store{table + threadIdx.x, static_cast<int>(blockDim.x), x0, y0};
  if (!defer) {
    store.clear(rows);
    RunDetector<PackedStore> runs(store, 0, x0, x1);
    if (none) walk_outline<XF_NONE>(poly, nv, T, m, runs);
    else walk_outline<XF_PATH_A>(poly, nv, T, m, runs);
    runs.finish();
    defer = store.overflow;
  }
  if (defer) {  // more than two ranges on some row: the list kernel finishes this pose
    push_todo(todo, todo_count, i);
    return;
  }
  bool odd = false;
  for (int r = 0; r < rows; ++r) odd |= (store.tab[r * store.stride] >> 30) == 1u;
  if (odd) {  // std::logic_error("Even number of line-pairs expected")
    store_failure(out, i, ST_LOGIC_ERROR);
    return;
  }
  Fold<OP> fold;
  for (int r = 0; r < rows; ++r) {
    const uint32_t w = store.tab[r * store.stride];
    if ((w >> 30) != 2u) continue;
    if (!fold.span(m, y0 + r, x0 + static_cast<int>(w & 0x7FFFu), x0 + static_cast<int>((w >> 15) & 0x7FFFu))) break;
  }
  fold.store(out, i);
  if (out.status) out.status[i] = ST_OK;
}


// ---------------------------------------------------------------------------------------------------
// k_area_pair<OP>: the per-thread fast path, second generation (replaces k_area_packed on the hot path).
// Same semantics — a pose is finished here iff every scan row gets at most two ranges — but written as one
// flat loop nest so that the compiler keeps the walker state in registers:
//   * the Bresenham loop of a ray is specialised by major axis (in a y-major ray every cell closes a run);
//   * a closed run is APPENDED to its row (one 64-bit shared-memory word per (row, thread): two 16+16-bit
//     ranges in insertion order, 0xFFFFFFFF = empty); the pair is merged once, when the row is scanned:
//     sorted by `min`, insertion order breaking ties, result [first.min, second.max] (generators.cpp:236-255);
//   * the run containing outline cell 0 is closed last but ranks first: it is inserted in FRONT.
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kPairEmpty = 0xFFFFFFFFu;
constexpr unsigned kFullMask = 0xFFFFFFFFu;

struct PairWalker {
  uint2* tab;   // &table[0][threadIdx.x], stride = blockDim.x
  int stride;
  int xb, yb;   // bounding-box origin: ranges are stored relative to xb
  int y_run, x_first, x_prev, y_before;
  int x0, y0, f_xlast, f_ynext;
  bool in_first, overflow;

  __device__ __forceinline__ void append(int y, int a, int b, int yp, int yn) {
    const uint32_t e = static_cast<uint32_t>(min(a, b) - xb) | (static_cast<uint32_t>(max(a, b) - xb) << 16);
    uint2& w = tab[(y - yb) * stride];
    uint2 v = w;
    const bool cusp = (yp - y) * (y - yn) < 0;  // generators.cpp:100-103: counted twice
    if (v.x == kPairEmpty) {
      v.x = e;
      if (cusp) v.y = e;
    } else if (v.y == kPairEmpty && !cusp) {
      v.y = e;
    } else {
      overflow = true;
    }
    w = v;
  }
  /// the run that contains outline cell 0: inserted BEFORE whatever the row already holds
  __device__ __forceinline__ void prepend(int y, int a, int b, int yp, int yn) {
    const uint32_t e = static_cast<uint32_t>(min(a, b) - xb) | (static_cast<uint32_t>(max(a, b) - xb) << 16);
    uint2& w = tab[(y - yb) * stride];
    uint2 v = w;
    const bool cusp = (yp - y) * (y - yn) < 0;
    if (v.y != kPairEmpty || (cusp && v.x != kPairEmpty)) {
      overflow = true;
    } else {
      v.y = cusp ? e : v.x;
      v.x = e;
    }
    w = v;
  }
  /// the outline moves on to row `y` at cell x: close the run that ended at x_prev
  __device__ __forceinline__ void row_change(int x, int y) {
    if (in_first) {
      f_xlast = x_prev, f_ynext = y;
      in_first = false;
    } else {
      append(y_run, x_first, x_prev, y_before, y);
    }
    y_before = y_run;
    y_run = y;
    x_first = x;
  }
  /// One outline cell: a row change closes the run that ended at x_prev.
  __device__ __forceinline__ void visit(int x, int y) {
    if (y != y_run) row_change(x, y);
    x_prev = x;
  }
};

/// Row table of the second per-thread pass: SLOTS (4) 32-bit ranges per (row, thread) in insertion order,
/// laid out [row][slot][thread] so that a warp's accesses never conflict.  0xFFFFFFFF = empty.
template <int SLOTS>
struct SlotWalker {
  uint32_t* tab;  // &table[0][0][threadIdx.x]
  int stride;     // threads per block
  int xb, yb;     // bounding-box origin: ranges are stored relative to xb
  int y_run, x_first, x_prev, y_before;
  int x0, y0, f_xlast, f_ynext;
  bool in_first, overflow;

  __device__ __forceinline__ uint32_t& slot(int row, int k) { return tab[(row * SLOTS + k) * stride]; }
  __device__ __forceinline__ int count(int row) {
    int n = 0;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) n += slot(row, k) != kPairEmpty;
    return n;
  }
  __device__ __forceinline__ void append(int y, int a, int b, int yp, int yn) {
    const uint32_t e = static_cast<uint32_t>(min(a, b) - xb) | (static_cast<uint32_t>(max(a, b) - xb) << 16);
    const int row = y - yb;
    const int need = ((yp - y) * (y - yn) < 0) ? 2 : 1;  // a cusp counts twice (generators.cpp:100-103)
    const int n = count(row);
    if (n + need > SLOTS) {
      overflow = true;
      return;
    }
    slot(row, n) = e;
    if (need == 2) slot(row, n + 1) = e;
  }
  /// the run that contains outline cell 0: inserted BEFORE whatever the row already holds
  __device__ __forceinline__ void prepend(int y, int a, int b, int yp, int yn) {
    const uint32_t e = static_cast<uint32_t>(min(a, b) - xb) | (static_cast<uint32_t>(max(a, b) - xb) << 16);
    const int row = y - yb;
    const int need = ((yp - y) * (y - yn) < 0) ? 2 : 1;
    const int n = count(row);
    if (n + need > SLOTS) {
      overflow = true;
      return;
    }
    for (int k = n - 1; k >= 0; --k) slot(row, k + need) = slot(row, k);
    slot(row, 0) = e;
    if (need == 2) slot(row, 1) = e;
  }
  /// the outline moves on to row `y` at cell x: close the run that ended at x_prev
  __device__ __forceinline__ void row_change(int x, int y) {
    if (in_first) {
      f_xlast = x_prev, f_ynext = y;
      in_first = false;
    } else {
      append(y_run, x_first, x_prev, y_before, y);
    }
    y_before = y_run;
    y_run = y;
    x_first = x;
  }
  /// One outline cell: a row change closes the run that ended at x_prev.
  __device__ __forceinline__ void visit(int x, int y) {
    if (y != y_run) row_change(x, y);
    x_prev = x;
  }
};

/// Per-lane Bresenham ray (generators.cpp:30-52) in step-vector form, so that x-major and y-major rays of
/// different lanes run through the SAME instructions.
struct LaneRay {
  int x, y, size, add, num;
  int major_x, major_y, minor_x, minor_y;
  __device__ __forceinline__ void set(int bx, int by, int ex, int ey) {
    const int dx = ex - bx, dy = ey - by, adx = abs(dx), ady = abs(dy);
    const int sx = (dx > 0) - (dx < 0), sy = (dy > 0) - (dy < 0);
    const bool xm = adx >= ady;  // ties: x is the major axis
    x = bx, y = by;
    size = xm ? adx : ady, add = xm ? ady : adx, num = size >> 1;
    major_x = xm ? sx : 0, major_y = xm ? 0 : sy, minor_x = xm ? 0 : sx, minor_y = xm ? sy : 0;
  }
  __device__ __forceinline__ void step() {
    num += add;
    if (num >= size) {
      num -= size;
      x += minor_x, y += minor_y;
    }
    x += major_x, y += major_y;
  }
};

/// Walks the rays of 32 poses in lock-step: ray e of every lane runs for max-over-lanes(size_e) iterations
/// (random headings make the per-lane sizes differ; one fused, predicated loop avoids the SIMT divergence of
/// 32 independently nested loops).  All 32 lanes must call this together.
template <int XF, int OP>
__device__ __forceinline__ void area_pair_pose(const double* __restrict__ poly, int nv, const Pose& T, const MapView& m, uint2* table, int rows_cap,
                                               const OutView& out, long long i, bool valid, unsigned int* __restrict__ todo,
                                               unsigned int* __restrict__ todo_count) {
  int x0, y0, x1, y1;
  const bool in_range = polygon_bbox<XF>(poly, nv, T, m, x0, y0, x1, y1);
  if (nv == 0) {  // empty generator: nothing is lethal, the sum is 0 (uniform branch: nv is a kernel argument)
    if (valid) {
      Fold<OP> fold;
      fold.store(out, i);
      if (out.status) out.status[i] = ST_OK;
    }
    return;
  }
  const int rows = y1 - y0 + 1;
  const bool too_big = rows > rows_cap || (x1 - x0) > 65000;
  const bool walk = valid && in_range && !too_big;
  PairWalker w;
  w.tab = table + threadIdx.x, w.stride = static_cast<int>(blockDim.x), w.xb = x0, w.yb = y0;
  const int rows_max = __reduce_max_sync(kFullMask, walk ? rows : 0);
  for (int r = 0; r < rows_max; ++r)
    if (walk && r < rows) w.tab[r * w.stride] = make_uint2(kPairEmpty, kPairEmpty);
  int fx, fy;
  to_cell<XF>(T, poly[0], poly[1], m, fx, fy);
  w.y_run = w.y0 = fy, w.x_first = w.x0 = w.x_prev = fx, w.y_before = fy;
  w.in_first = true, w.overflow = false;
  w.f_xlast = fx, w.f_ynext = fy;
  int px = fx, py = fy, rays = 0;
  LaneRay ray;
  for (int v = 1; v <= nv; ++v) {  // v == nv: the closing ray, or the one-cell dummy ray of an open outline
    int cx, cy;
    if (v < nv) {
      to_cell<XF>(T, poly[2 * v], poly[2 * v + 1], m, cx, cy);
    } else {
      cx = fx, cy = fy;
    }
    ray.set(px, py, cx, cy);
    if (v < nv) {
      rays += ray.size != 0;  // degenerate rays are skipped (generators.cpp:86-93)
    } else if (rays <= 1) {
      ray.size = 1, ray.add = 0, ray.num = 0;  // dummy ray: the last point itself (generators.cpp:72-78)
      ray.major_x = ray.major_y = ray.minor_x = ray.minor_y = 0;
    }
    if (!walk) ray.size = 0;
    const int n_iter = __reduce_max_sync(kFullMask, ray.size);
    for (int k = 0; k < n_iter; ++k) {
      if (k < ray.size) {
        w.visit(ray.x, ray.y);
        ray.step();
      }
    }
    px = cx, py = cy;
  }
  if (walk) {
    if (w.in_first) {  // every cell on one row: ranges (min,min),(max,max) (generators.cpp:197-204)
      w.tab[(w.y0 - y0) * w.stride] = make_uint2(0u, static_cast<uint32_t>(x1 - x0) | (static_cast<uint32_t>(x1 - x0) << 16));
    } else if (w.y_run == w.y0) {  // the tail run continues into cell 0
      w.prepend(w.y0, w.x_first, w.f_xlast, w.y_before, w.f_ynext);
    } else {
      w.append(w.y_run, w.x_first, w.x_prev, w.y_before, w.y0);
      w.prepend(w.y0, w.x0, w.f_xlast, w.y_run, w.f_ynext);
    }
  }
  // statuses and deferrals
  bool scan = walk && !w.overflow;
  if (valid && !in_range) store_failure(out, i, ST_COORD_RANGE);
  if (valid && in_range && (too_big || w.overflow)) push_todo(todo, todo_count, i);  // > 2 ranges on a row
  bool odd = false;
  for (int r = 0; r < rows_max; ++r) {
    if (scan && r < rows) {
      const uint2 v = w.tab[r * w.stride];
      odd |= v.x != kPairEmpty && v.y == kPairEmpty;
    }
  }
  if (scan && odd) store_failure(out, i, ST_LOGIC_ERROR);  // std::logic_error("Even number of line-pairs expected")
  scan = scan && !odd;
  Fold<OP> fold;
  bool more = scan;
  for (int r = 0; r < rows_max; ++r) {
    if (__all_sync(kFullMask, !more)) break;
    if (more && r < rows) {
      const uint2 v = w.tab[r * w.stride];
      if (v.x != kPairEmpty) {
        const int lo1 = v.x & 0xFFFF, hi1 = v.x >> 16, lo2 = v.y & 0xFFFF, hi2 = v.y >> 16;
        const int lo = min(lo1, lo2), hi = lo2 < lo1 ? hi1 : hi2;  // stable sort by min, then [first.min, second.max]
        more = fold.span(m, y0 + r, x0 + lo, x0 + hi);
      }
    }
  }
  if (scan) {
    fold.store(out, i);
    if (out.status) out.status[i] = ST_OK;
  }
}

template <int OP>
__global__ void __launch_bounds__(kBlock) k_area_pair(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src, long long count,
                                                     OutView out, int rows_cap, unsigned int* __restrict__ todo,
                                                     unsigned int* __restrict__ todo_count) {
  extern __shared__ double smem_d[];
  const double* poly = stage_polygon(poly_g, nv, smem_d);
  uint2* table = reinterpret_cast<uint2*>(smem_d + 2 * nv);
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < count;  // no early return: the 32 lanes of a warp walk their outlines in lock-step
  const Pose T = load_pose(src, valid ? i : 0);
  if (src.format == POSE_NONE) area_pair_pose<XF_NONE, OP>(poly, nv, T, m, table, rows_cap, out, i, valid, todo, todo_count);
  else area_pair_pose<XF_PATH_A, OP>(poly, nv, T, m, table, rows_cap, out, i, valid, todo, todo_count);
}

/// Second per-thread pass: the poses k_area_pair deferred (more than two ranges on a row) are walked again with
/// SLOTS ranges per row — an L, a notch or a kidney rarely needs more than four; what still overflows goes to the
/// general kernels.  Same lock-step walk as area_pair_pose.
template <int XF, int OP, int SLOTS>
__device__ __forceinline__ void area_slot_pose(const double* __restrict__ poly, int nv, const Pose& T, const MapView& m, uint32_t* table,
                                               int rows_cap, const OutView& out, long long i, bool valid, unsigned int* __restrict__ todo,
                                               unsigned int* __restrict__ todo_count) {
  int x0, y0, x1, y1;
  const bool in_range = polygon_bbox<XF>(poly, nv, T, m, x0, y0, x1, y1);
  if (nv == 0) {  // empty generator: nothing is lethal, the sum is 0 (uniform branch: nv is a kernel argument)
    if (valid) {
      Fold<OP> fold;
      fold.store(out, i);
      if (out.status) out.status[i] = ST_OK;
    }
    return;
  }
  const int rows = y1 - y0 + 1;
  const bool too_big = rows > rows_cap || (x1 - x0) > 65000;
  const bool walk = valid && in_range && !too_big;
  SlotWalker<SLOTS> w;
  w.tab = table + threadIdx.x, w.stride = static_cast<int>(blockDim.x), w.xb = x0, w.yb = y0;
  const int rows_max = __reduce_max_sync(kFullMask, walk ? rows : 0);
  for (int r = 0; r < rows_max; ++r)
    if (walk && r < rows) {
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) w.slot(r, k) = kPairEmpty;
    }
  int fx, fy;
  to_cell<XF>(T, poly[0], poly[1], m, fx, fy);
  w.y_run = w.y0 = fy, w.x_first = w.x0 = w.x_prev = fx, w.y_before = fy;
  w.in_first = true, w.overflow = false;
  w.f_xlast = fx, w.f_ynext = fy;
  int px = fx, py = fy, rays = 0;
  LaneRay ray;
  for (int v = 1; v <= nv; ++v) {  // v == nv: the closing ray, or the one-cell dummy ray of an open outline
    int cx, cy;
    if (v < nv) {
      to_cell<XF>(T, poly[2 * v], poly[2 * v + 1], m, cx, cy);
    } else {
      cx = fx, cy = fy;
    }
    ray.set(px, py, cx, cy);
    if (v < nv) {
      rays += ray.size != 0;  // degenerate rays are skipped (generators.cpp:86-93)
    } else if (rays <= 1) {
      ray.size = 1, ray.add = 0, ray.num = 0;  // dummy ray: the last point itself (generators.cpp:72-78)
      ray.major_x = ray.major_y = ray.minor_x = ray.minor_y = 0;
    }
    if (!walk) ray.size = 0;
    const int n_iter = __reduce_max_sync(kFullMask, ray.size);
    for (int k = 0; k < n_iter; ++k) {
      if (k < ray.size) {
        w.visit(ray.x, ray.y);
        ray.step();
      }
    }
    px = cx, py = cy;
  }
  if (walk) {
    if (w.in_first) {  // every cell on one row: ranges (min,min),(max,max) (generators.cpp:197-204)
      const uint32_t wd = static_cast<uint32_t>(x1 - x0);
      w.slot(w.y0 - y0, 0) = 0u;
      w.slot(w.y0 - y0, 1) = wd | (wd << 16);
    } else if (w.y_run == w.y0) {  // the tail run continues into cell 0
      w.prepend(w.y0, w.x_first, w.f_xlast, w.y_before, w.f_ynext);
    } else {
      w.append(w.y_run, w.x_first, w.x_prev, w.y_before, w.y0);
      w.prepend(w.y0, w.x0, w.f_xlast, w.y_run, w.f_ynext);
    }
  }
  // statuses and deferrals
  bool scan = walk && !w.overflow;
  if (valid && !in_range) store_failure(out, i, ST_COORD_RANGE);
  if (valid && in_range && (too_big || w.overflow)) push_todo(todo, todo_count, i);  // too many ranges on a row
  bool odd = false;
  for (int r = 0; r < rows_max; ++r)
    if (scan && r < rows) odd |= (w.count(r) & 1) != 0;
  if (scan && odd) store_failure(out, i, ST_LOGIC_ERROR);  // std::logic_error("Even number of line-pairs expected")
  scan = scan && !odd;
  Fold<OP> fold;
  bool more = scan;
  for (int r = 0; r < rows_max; ++r) {
    if (__all_sync(kFullMask, !more)) break;
    if (more && r < rows) {
      uint32_t e[SLOTS];
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) e[k] = w.slot(r, k);
      if (SLOTS == 4 && e[2] != kPairEmpty) {
        // four ranges: stable insertion sort by `min` (low 16 bits), then [r0.min, r1.max], [r2.min, r3.max]
#pragma unroll
        for (int a = 1; a < 4; ++a) {
#pragma unroll
          for (int b = a; b > 0; --b) {
            if ((e[b] & 0xFFFFu) < (e[b - 1] & 0xFFFFu)) {
              const uint32_t t = e[b];
              e[b] = e[b - 1], e[b - 1] = t;
            }
          }
        }
        more = fold.span(m, y0 + r, x0 + static_cast<int>(e[0] & 0xFFFFu), x0 + static_cast<int>(e[1] >> 16));
        if (more) more = fold.span(m, y0 + r, x0 + static_cast<int>(e[2] & 0xFFFFu), x0 + static_cast<int>(e[3] >> 16));
      } else if (e[0] != kPairEmpty) {
        const int lo1 = e[0] & 0xFFFF, hi1 = e[0] >> 16, lo2 = e[1] & 0xFFFF, hi2 = e[1] >> 16;
        const int lo = min(lo1, lo2), hi = lo2 < lo1 ? hi1 : hi2;  // stable sort by min, then [first.min, second.max]
        more = fold.span(m, y0 + r, x0 + lo, x0 + hi);
      }
    }
  }
  if (scan) {
    fold.store(out, i);
    if (out.status) out.status[i] = ST_OK;
  }
}

template <int OP, int SLOTS>
__global__ void __launch_bounds__(kBlock) k_area_slots(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src, long long count, OutView out,
                                                      int rows_cap, const unsigned int* __restrict__ in_list,
                                                      const unsigned int* __restrict__ in_count, unsigned int* __restrict__ todo,
                                                      unsigned int* __restrict__ todo_count) {
  extern __shared__ double smem_d[];
  const double* poly = stage_polygon(poly_g, nv, smem_d);
  uint32_t* table = reinterpret_cast<uint32_t*>(smem_d + 2 * nv);
  bool listed;
  const long long n_items = todo_items(in_count, count, listed);
  const long long n_threads = static_cast<long long>(gridDim.x) * blockDim.x;
  // persistent grid over the deferred poses; whole warps iterate together (lock-step walk)
  for (long long base = static_cast<long long>(blockIdx.x) * blockDim.x; base < n_items; base += n_threads) {
    const long long item = base + threadIdx.x;
    const bool valid = item < n_items;
    const long long i = valid ? (listed ? static_cast<long long>(in_list[item]) : item) : 0;
    const Pose T = load_pose(src, i);
    if (src.format == POSE_NONE) area_slot_pose<XF_NONE, OP, SLOTS>(poly, nv, T, m, table, rows_cap, out, i, valid, todo, todo_count);
    else area_slot_pose<XF_PATH_A, OP, SLOTS>(poly, nv, T, m, table, rows_cap, out, i, valid, todo, todo_count);
  }
}

// ---------------------------------------------------------------------------------------------------
// k_area_quad<OP>: the per-thread pass of NON-CONVEX polygons whose bounding box is at most 250 cells wide.  Same
// lock-step walk as k_area_pair, but a (row, thread) word of the shared-memory table holds FOUR ranges of 8+8 bits
// (0xFFFF = empty, filled from the low end in insertion order) - an L, an S, a comb rarely puts more than four on a
// row - so one pass finishes what k_area_pair + k_area_slots needed two for, at the pair kernel's occupancy.
// ---------------------------------------------------------------------------------------------------
constexpr int kQuadMaxWidth = 250;  // relative x in 8 bits, and no entry is ever 0xFFFF
constexpr int kQuadMaxRows = 96;    // taller polygons (four-metre combs at 5 cm) do better with a warp per pose (k_area_rows)

struct QuadWalker {
  uint2* tab;  // &table[0][threadIdx.x], stride = blockDim.x; .x = entries 0 and 1, .y = entries 2 and 3
  int stride;
  int xb, yb;
  int y_run, x_first, x_prev, y_before;
  int x0, y0, f_xlast, f_ynext;
  bool in_first, overflow;

  /// entries fill from the low end, so "entry k present" reads off two words: 0: x != ~0, 1: x < 0xFFFF0000, 2, 3: same of y
  __device__ __forceinline__ static bool odd(uint2 v) { return (v.x != kPairEmpty && v.x >= 0xFFFF0000u) || (v.y != kPairEmpty && v.y >= 0xFFFF0000u); }
  __device__ __forceinline__ uint32_t entry(int a, int b) const {
    return static_cast<uint32_t>(min(a, b) - xb) | (static_cast<uint32_t>(max(a, b) - xb) << 8);
  }
  __device__ __forceinline__ void append(int y, int a, int b, int yp, int yn) {
    const uint32_t e = entry(a, b);
    uint2& w = tab[(y - yb) * stride];
    uint2 v = w;
    const bool f0 = v.x != kPairEmpty, f1 = v.x < 0xFFFF0000u, f2 = v.y != kPairEmpty, f3 = v.y < 0xFFFF0000u;
    if ((yp - y) * (y - yn) < 0) {  // a cusp counts twice (generators.cpp:100-103)
      const uint32_t ee = e | (e << 16);
      if (!f0) {
        v.x = ee;
      } else if (!f1) {
        v.x = __byte_perm(v.x, e, 0x5410), v.y = __byte_perm(v.y, e, 0x3254);
      } else if (!f2) {
        v.y = ee;
      } else {
        overflow = true;
      }
    } else {  // the first free 16-bit slot takes the entry
      v.x = __byte_perm(v.x, e, !f0 ? 0x3254 : (!f1 ? 0x5410 : 0x3210));
      v.y = __byte_perm(v.y, e, (f1 && !f2) ? 0x3254 : ((f2 && !f3) ? 0x5410 : 0x3210));
      overflow |= f3;
    }
    w = v;
  }
  /// the run that contains outline cell 0: inserted BEFORE whatever the row already holds
  __device__ __forceinline__ void prepend(int y, int a, int b, int yp, int yn) {
    const uint32_t e = entry(a, b);
    uint2& w = tab[(y - yb) * stride];
    const uint2 v = w;
    const bool f1 = v.x < 0xFFFF0000u, f2 = v.y != kPairEmpty, f3 = v.y < 0xFFFF0000u;
    if ((yp - y) * (y - yn) < 0) {
      if (f2) overflow = true;
      else w = make_uint2(e | (e << 16), v.x);
    } else {
      if (f3) overflow = true;
      else w = make_uint2((v.x << 16) | e, __byte_perm(v.x, v.y, 0x5432));
    }
    (void)f1;
  }
  __device__ __forceinline__ void row_change(int x, int y) {
    if (in_first) {
      f_xlast = x_prev, f_ynext = y;
      in_first = false;
    } else {
      append(y_run, x_first, x_prev, y_before, y);
    }
    y_before = y_run;
    y_run = y;
    x_first = x;
  }
  __device__ __forceinline__ void visit(int x, int y) {
    if (y != y_run) row_change(x, y);
    x_prev = x;
  }
};

/// The walk of one polygon per lane, in lock-step (all 32 lanes call this together; `active` lanes have a pose whose
/// vertices [x0, x1] x [y0, y1] are in range).  Returns 0: the area was scanned into `fold`; 1: left to the general
/// kernels (taller / wider than the table, more than four ranges on a row); 2: an odd number of ranges on a row
/// (std::logic_error upstream); -1: inactive lane.
template <int XF, int OP>
__device__ __forceinline__ int area_quad_core(const double* __restrict__ poly, int nv, const Pose& T, const MapView& m, uint2* table, int rows_cap,
                                              bool active, int x0, int y0, int x1, int y1, Fold<OP>& fold) {
  const int rows = y1 - y0 + 1;
  const bool too_big = rows > rows_cap || (x1 - x0) > kQuadMaxWidth;
  const bool walk = active && !too_big;
  QuadWalker w;
  w.tab = table + threadIdx.x, w.stride = static_cast<int>(blockDim.x), w.xb = x0, w.yb = y0;
  const int rows_max = __reduce_max_sync(kFullMask, walk ? rows : 0);
  for (int r = 0; r < rows_max; ++r)
    if (walk && r < rows) w.tab[r * w.stride] = make_uint2(kPairEmpty, kPairEmpty);
  int fx, fy;
  to_cell<XF>(T, poly[0], poly[1], m, fx, fy);
  w.y_run = w.y0 = fy, w.x_first = w.x0 = w.x_prev = fx, w.y_before = fy;
  w.in_first = true, w.overflow = false;
  w.f_xlast = fx, w.f_ynext = fy;
  int px = fx, py = fy, rays = 0;
  LaneRay ray;
  for (int v = 1; v <= nv; ++v) {  // v == nv: the closing ray, or the one-cell dummy ray of an open outline
    int cx, cy;
    if (v < nv) {
      to_cell<XF>(T, poly[2 * v], poly[2 * v + 1], m, cx, cy);
    } else {
      cx = fx, cy = fy;
    }
    ray.set(px, py, cx, cy);
    if (v < nv) {
      rays += ray.size != 0;  // degenerate rays are skipped (generators.cpp:86-93)
    } else if (rays <= 1) {
      ray.size = 1, ray.add = 0, ray.num = 0;  // dummy ray: the last point itself (generators.cpp:72-78)
      ray.major_x = ray.major_y = ray.minor_x = ray.minor_y = 0;
    }
    if (!walk) ray.size = 0;
    const int n_iter = __reduce_max_sync(kFullMask, ray.size);
    for (int k = 0; k < n_iter; ++k) {
      if (k < ray.size) {
        w.visit(ray.x, ray.y);
        ray.step();
      }
    }
    px = cx, py = cy;
  }
  if (walk) {
    if (w.in_first) {  // every cell on one row: ranges (min,min),(max,max) (generators.cpp:197-204)
      const uint32_t wd = static_cast<uint32_t>(x1 - x0);
      w.tab[(w.y0 - y0) * w.stride] = make_uint2((wd | (wd << 8)) << 16, kPairEmpty);
    } else if (w.y_run == w.y0) {  // the tail run continues into cell 0
      w.prepend(w.y0, w.x_first, w.f_xlast, w.y_before, w.f_ynext);
    } else {
      w.append(w.y_run, w.x_first, w.x_prev, w.y_before, w.y0);
      w.prepend(w.y0, w.x0, w.f_xlast, w.y_run, w.f_ynext);
    }
  }
  bool scan = walk && !w.overflow;
  bool odd = false;
  for (int r = 0; r < rows_max; ++r)
    if (scan && r < rows) odd |= QuadWalker::odd(w.tab[r * w.stride]);
  scan = scan && !odd;
  bool more = scan;
  for (int r = 0; r < rows_max; ++r) {
    if (__all_sync(kFullMask, !more)) break;
    if (more && r < rows) {
      const uint2 v = w.tab[r * w.stride];
      const uint32_t e0 = v.x & 0xFFFFu, e1 = v.x >> 16, e2 = v.y & 0xFFFFu, e3 = v.y >> 16;
      if (e2 != 0xFFFFu) {
        // four ranges: stable insertion sort by `min` (low byte), then [r0.min, r1.max], [r2.min, r3.max] (generators.cpp:236-255)
        uint32_t e[4] = {e0, e1, e2, e3};
#pragma unroll
        for (int a = 1; a < 4; ++a) {
#pragma unroll
          for (int b = a; b > 0; --b) {
            if ((e[b] & 0xFFu) < (e[b - 1] & 0xFFu)) {
              const uint32_t t = e[b];
              e[b] = e[b - 1], e[b - 1] = t;
            }
          }
        }
        more = fold.span(m, y0 + r, x0 + static_cast<int>(e[0] & 0xFFu), x0 + static_cast<int>(e[1] >> 8));
        if (more) more = fold.span(m, y0 + r, x0 + static_cast<int>(e[2] & 0xFFu), x0 + static_cast<int>(e[3] >> 8));
      } else if (e0 != 0xFFFFu) {
        const int lo1 = e0 & 0xFF, hi1 = e0 >> 8, lo2 = e1 & 0xFF, hi2 = e1 >> 8;
        const int lo = min(lo1, lo2), hi = lo2 < lo1 ? hi1 : hi2;  // stable sort by min, then [first.min, second.max]
        more = fold.span(m, y0 + r, x0 + lo, x0 + hi);
      }
    }
  }
  if (!active) return -1;
  if (too_big || w.overflow) return 1;
  return odd ? 2 : 0;
}

template <int XF, int OP>
__device__ __forceinline__ void area_quad_pose(const double* __restrict__ poly, int nv, const Pose& T, const MapView& m, uint2* table,
                                               int rows_cap, const OutView& out, long long i, bool valid, unsigned int* __restrict__ todo,
                                               unsigned int* __restrict__ todo_count) {
  int x0, y0, x1, y1;
  const bool in_range = polygon_bbox<XF>(poly, nv, T, m, x0, y0, x1, y1);
  Fold<OP> fold;
  const int rc = area_quad_core<XF, OP>(poly, nv, T, m, table, rows_cap, valid && in_range, x0, y0, x1, y1, fold);
  if (valid && !in_range) store_failure(out, i, ST_COORD_RANGE);
  if (rc == 1) push_todo(todo, todo_count, i);
  if (rc == 2) store_failure(out, i, ST_LOGIC_ERROR);  // std::logic_error("Even number of line-pairs expected")
  if (rc == 0) {
    fold.store(out, i);
    if (out.status) out.status[i] = ST_OK;
  }
}

template <int OP>
__global__ void __launch_bounds__(kBlock) k_area_quad(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src, long long count,
                                                     OutView out, int rows_cap, unsigned int* __restrict__ todo,
                                                     unsigned int* __restrict__ todo_count) {
  extern __shared__ double smem_d[];
  const double* poly = stage_polygon(poly_g, nv, smem_d);
  uint2* table = reinterpret_cast<uint2*>(smem_d + 2 * nv);
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < count;  // no early return: the 32 lanes of a warp walk their outlines in lock-step
  const Pose T = load_pose(src, valid ? i : 0);
  if (src.format == POSE_NONE) area_quad_pose<XF_NONE, OP>(poly, nv, T, m, table, rows_cap, out, i, valid, todo, todo_count);
  else area_quad_pose<XF_PATH_A, OP>(poly, nv, T, m, table, rows_cap, out, i, valid, todo, todo_count);
}

// ---------------------------------------------------------------------------------------------------
// Lattice kernel (is_free on a POSE_LATTICE batch): the DWA-style headline path.
//
// Two facts about an (x, y, heading) lattice make the work shareable without giving up per-pose exactness:
//  * the pose transform separates: the x cells of the vertices depend only on (heading, ix), the y cells
//    only on (heading, iy) — a unit of 64 columns x 8 rows needs 64 x-profiles + 8 y-profiles, not 512
//    vertex transforms;
//  * almost every pose of one heading has the SAME discretised polygon up to a translation: the vertices'
//    cell offsets relative to vertex 0 only change where a coordinate sits within rounding noise of a
//    cell boundary.
// A persistent CTA walks a contiguous range of units; each warp owns one unit at a time (lane = two
// columns, ix and ix+32; lane r also owns lattice row r of the unit) and
//   1. computes the exact x-profile of its columns and the y-profile of its row (key = the packed vertex
//      offsets themselves for <= 4 vertices, else a 64-bit hash that is always verified against the actual
//      offsets) and notes which rows repeat the previous row's offsets,
//   2. forms column groups ONCE per unit: columns with identical x offsets at most 64 cells apart,
//   3. per row and group looks the shape (x key, y key) up in a shared-memory cache that persists across the
//      CTA's units - or carries the slot over from the previous row while the y offsets repeat; the lane that
//      inserts a NEW shape rasterises it on the spot (the same run detector / packed row table as
//      k_area_packed, so the scan-line semantics are literally the same code) and publishes it; warps that
//      meet a shape still being rasterised wait for its status word - no CTA-wide barrier,
//   4. tests the group's (up to 64) poses at once.  Verdicts: lane j = scan row j reads the two power-of-two
//      windows of its row from the lethal bit planes, a redux.or folds the rows, every pose picks the bit at
//      its own x offset.  Value folds / map border: the row segment is loaded one byte per lane (three
//      coalesced 32-byte pieces) and turned into a per-row table (prefix sums, sliding maxima) or a ballot
//      bit window.
// Poses whose shape needs more than two ranges on a row, is wider than 32 cells, or finds the cache
// full go to the todo list and are finished by the general kernels.
// ---------------------------------------------------------------------------------------------------
#ifndef LAT_MIN_BLOCKS
#define LAT_MIN_BLOCKS 4
#endif
#ifndef LAT_LEAN_BLOCKS
#define LAT_LEAN_BLOCKS 4
#endif
constexpr int kLatSegX = 64, kLatWarps = 8, kLatThreads = kLatWarps * 32, kLatSlotsMax = 64;
constexpr int kLatMaxVertices = 64;
constexpr int kLatMaxWidth = 32;   // widest span (cells): 64 pose offsets + 32 cells = one 96-cell window; wider pairs are split
constexpr int kLatMaxRowsCap = 72; // tallest footprint (scan rows, incl. the 3 rows of slack of rows_bound) the shape cache takes
constexpr int kLatPending = -2;    // shape inserted, not rasterised yet
constexpr int kLatUnfit = -1;      // shape not representable by this kernel
constexpr unsigned kFull = 0xFFFFFFFFu;

struct LatShape {
  int xb, x1, yb, rows, status;  // bbox of the relative vertices (x: [xb, x1], y: [yb, yb+rows-1])
  int n_spans;                   // entries of the span table (one per merged range pair, in generator order)
};

struct LatGeom {
  long long row_first;  // first lattice row (it*ny + iy) touched by this batch
  int n_rows, segs;     // lattice rows touched; 64-column segments per row
  long long n_units;    // segs * ceil(n_rows / rows_per_unit)
  int rows_per_unit;    // 8, 16 or 32 lattice rows per unit (lane r owns row r): as many as still fill the grid
  int n_slots;          // shape-cache slots per CTA (power of two <= 64: fewer when a shape's span table is big)
  int static_rounds;    // longest stretch of consecutive rounds a CTA draws at once (guided hand-out)
  long long tickets_a, tickets_b;  // draws of `static_rounds` rounds, then of half that, then single rounds
  int outline;          // 0: areas; 1: outlines as cell sets (verdict, maximum); 2: outlines in traversal order (accumulate_cost)
};

/// Second level of the shape cache, shared by all CTAs of a launch (global memory, L2-resident): the CTA that meets a
/// shape first rasterises it and publishes the result here, every other CTA copies it into its shared-memory cache
/// instead of rasterising it again (a single-lane walk of ~20 us).  This is what lets the units be handed out
/// dynamically, one round at a time, whatever heading they belong to.
struct LatGlobal {
  unsigned long long* key;   // [capacity] 0 = empty
  int* status;               // [capacity] 0 = pending, else shape status + 16
  int* data;                 // [capacity][entry_words]: LatShape (6 ints), canon (2 nv ints), span table (2 span_cap ints)
  int entry_words, capacity; // capacity is a power of two
  unsigned long long* next_round;  // dynamic work distribution: the next unit index to hand out
};
constexpr int kLatGlobalProbes = 16;

__device__ __forceinline__ unsigned long long lat_mix(unsigned long long h, int v) {
  return (h ^ static_cast<unsigned int>(v)) * 0x100000001B3ull;  // FNV-1a step
}

/// x cell of one vertex: the x half of to_cell<XF_PATH_A> (same operations in the same order).
__device__ __forceinline__ bool lat_cell_x(double c, double s, double tx, double px, double py, const MapView& m, int& cx) {
  const double q = __dsub_rn(__dadd_rn(tx, __dadd_rn(__dmul_rn(c, px), __dmul_rn(-s, py))), m.origin_x);
  const double sc = __dmul_rn(q, m.inv_res);
  cx = __double2int_rz(sc);
  return fabs(sc) < kCoordLimit;
}
/// y cell of one vertex: the y half of to_cell<XF_PATH_A>.
__device__ __forceinline__ bool lat_cell_y(double c, double s, double ty, double px, double py, const MapView& m, int& cy) {
  const double q = __dsub_rn(__dadd_rn(ty, __dadd_rn(__dmul_rn(s, px), __dmul_rn(c, py))), m.origin_y);
  const double sc = __dmul_rn(q, m.inv_res);
  cy = __double2int_rz(sc);
  return fabs(sc) < kCoordLimit;
}

/// One coordinate's profile of a column (x) or row (y): cell of vertex 0 + key of the offsets of the others.
struct LatProfile {
  int c0;                  // cell coordinate of vertex 0
  unsigned long long key;  // exact mode: offsets packed 8 bits each; hash mode: FNV of the offsets
  bool ok, packable;       // all coordinates in range; all offsets fit 8 bits (exact mode)
};

template <bool IS_X>
__device__ __forceinline__ LatProfile lat_profile(const double* __restrict__ poly, int nv, double c, double s, double t, const MapView& m,
                                                  bool exact_key) {
  LatProfile p{0, exact_key ? 0ull : 0xCBF29CE484222325ull, true, true};
  for (int v = 0; v < nv; ++v) {
    int cell;
    p.ok &= IS_X ? lat_cell_x(c, s, t, poly[2 * v], poly[2 * v + 1], m, cell) : lat_cell_y(c, s, t, poly[2 * v], poly[2 * v + 1], m, cell);
    if (v == 0) {
      p.c0 = cell;
    } else if (exact_key) {
      const int d = cell - p.c0;
      p.packable &= d >= -128 && d <= 127;
      p.key = (p.key << 8) | static_cast<unsigned long long>(d & 0xFF);
    } else {
      p.key = lat_mix(p.key, cell - p.c0);
    }
  }
  return p;
}

/// Span table of a cached shape: entry k = one merged range pair of the area, in generator order (rows ascending,
/// x ascending): `span_a[k]` = dy * pitch + (first cell, relative to vertex 0) - the byte offset from (vertex 0's
/// cell, first row of the shape) - and `span_m[k]` = width | descending << 7 | dy << 8 (width <= 32).
__device__ __forceinline__ int span_width(uint32_t m) { return static_cast<int>(m & 0x7Fu); }
/// (outline spans of accumulate_cost only) the traversal walks the span towards smaller x
__device__ __forceinline__ bool span_descending(uint32_t m) { return (m & 0x80u) != 0u; }
__device__ __forceinline__ int span_dy(uint32_t m) { return static_cast<int>(m >> 8); }

/// Byte walk of one group with per-cell bounds tests (a pose of the group touches the map border, or the bit planes
/// are unavailable).  Each lane loads one byte of each of three coalesced 32-cell pieces per span, `__ballot_sync`
/// turns "lethal or outside the map" into a 96-cell bit window and every pose ANDs its own sub-window.
__device__ __noinline__ void lattice_rows_bytes(const MapView& m, const int* __restrict__ span_a, const uint32_t* __restrict__ span_m, int n_spans,
                                                int x_base, int y_base, int lane, int off_a, int off_b, uint32_t& hit_a, uint32_t& hit_b) {
  // lane constants: which two of the three 32-cell pieces a pose's window starts in, and its shift
  const bool hi_a = off_a >= 32, hi_b = off_b >= 32;
  const int sh_a = off_a & 31, sh_b = off_b & 31;
  for (int r = 0; r < n_spans; ++r) {
    const uint32_t sm = span_m[r];
    const int dy = span_dy(sm), width = span_width(sm);
    const int y = y_base + dy;
    const bool y_in = static_cast<unsigned>(y) < static_cast<unsigned>(m.size_y);
    const uint8_t* __restrict__ rowp = m.data + static_cast<size_t>(y_in ? y : 0) * m.pitch;
    const int lo = span_a[r] - dy * m.pitch;
    uint32_t piece[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int x = x_base + lo + lane + 32 * j;
      bool lethal = true;  // outside the map counts as lethal (costmap.hpp:81-91)
      if (y_in && static_cast<unsigned>(x) < static_cast<unsigned>(m.size_x)) lethal = __ldg(rowp + x) == kLethal;
      piece[j] = __ballot_sync(kFull, lethal);
    }
    const uint32_t wm = width >= 32 ? 0xFFFFFFFFu : ((1u << width) - 1u);
    hit_a |= __funnelshift_r(hi_a ? piece[1] : piece[0], hi_a ? piece[2] : piece[1], sh_a) & wm;
    hit_b |= __funnelshift_r(hi_b ? piece[1] : piece[0], hi_b ? piece[2] : piece[1], sh_b) & wm;
  }
}

/// Lethal bit planes of a costmap (k_lethal_planes): plane k, row y, bit x = "one of the cells x .. x + 2^k - 1
/// of row y is LETHAL" (k = 0..5; bits past the map edge are 0).  Any window of w <= 32 cells is then the OR of
/// two overlapping power-of-two windows - two bits - instead of w bytes.
struct PlaneView {
  const uint32_t* __restrict__ bits;  // nullptr: planes disabled, the byte walk is used
  int ypitch;                         // COLUMN-major planes: word (column c = x / 32, row y) lives at c * ypitch + y, so the
                                      // lanes of a warp (= consecutive scan rows, nearly the same x) read consecutive words
  long long plane_words;              // words per plane = (ceil(size_x / 32) + 3 slack columns) * ypitch
  // accumulate_cost sweeps: per-row exclusive prefix sums of the cost bytes (k_cost_prefix), row-major, ppitch words
  // per row: word x = (NO_INFORMATION cells before x, mod 256) << 24 | (LETHAL cells before x, mod 256) << 16 | (sum of the
  // costs before x, mod 2^16).  A window of up to 32 cells is two loads; every field is exact as a difference.  nullptr: not built.
  const uint32_t* __restrict__ prefix;
  int ppitch;
};
constexpr int kPlaneCount = 6;

/// accumulate_cost of one group over the prefix rows (every needed cell inside the map): the spans are walked in
/// generator order (rows ascending, x ascending); a pose's window [off, off + width) of a span costs two loads -
/// consecutive poses read consecutive words.  A window that holds a LETHAL / NO_INFORMATION cell is walked byte by
/// byte for the first one (the reference's "first negative value in generator order", impl/costmap.hpp:50-63).
__device__ __forceinline__ void lattice_spans_accumulate(const MapView& m, const PlaneView& pv, const int* __restrict__ span_a,
                                                         const uint32_t* __restrict__ span_m, int n_spans, int x_base, int y_base, int off_a,
                                                         int off_b, Fold<OP_ACCUMULATE>& fa, Fold<OP_ACCUMULATE>& fb) {
  uint32_t sum_a = 0u, sum_b = 0u;
  bool more_a = true, more_b = true;
  for (int k = 0; k < n_spans; ++k) {
    const uint32_t sm = span_m[k];
    const int width = span_width(sm), dy = span_dy(sm);
    const int xs = x_base + (span_a[k] - dy * m.pitch);
    const uint32_t* __restrict__ row = pv.prefix + static_cast<long long>(y_base + dy) * pv.ppitch + xs;
    const uint32_t a0 = __ldg(row + off_a), a1 = __ldg(row + off_a + width);
    const uint32_t b0 = __ldg(row + off_b), b1 = __ldg(row + off_b + width);
    const uint32_t cnt_a = ((a1 >> 16) - (a0 >> 16)) & 0xFFFFu, cnt_b = ((b1 >> 16) - (b0 >> 16)) & 0xFFFFu;  // (both sentinel counts: zero iff none)
    sum_a += (a1 - a0) & 0xFFFFu;  // (meaningless once a sentinel was met: `early` decides then)
    sum_b += (b1 - b0) & 0xFFFFu;
    if ((cnt_a != 0u && more_a) || (cnt_b != 0u && more_b)) {
      const uint8_t* __restrict__ cells = m.data + static_cast<size_t>(y_base + dy) * m.pitch + xs;
#pragma unroll
      for (int plane = 0; plane < 2; ++plane) {
        const bool hit = plane ? (cnt_b != 0u && more_b) : (cnt_a != 0u && more_a);
        if (!hit) continue;
        const int off = plane ? off_b : off_a;
        int early = 0;
        for (int t = 0; t < width && early == 0; ++t) {  // in traversal order (areas: ascending x)
          const int c = __ldg(cells + off + (span_descending(sm) ? width - 1 - t : t));
          if (c >= kLethal) early = c == kLethal ? -1 : -2;
        }
        if (plane) fb.early = early, more_b = false;
        else fa.early = early, more_a = false;
      }
    }
  }
  fa.sum += sum_a, fb.sum += sum_b;
}

/// Builds rows [row_lo, row_hi] of the prefix rows: one warp per row, 32 cells per step, the running total carried in
/// the warp.
__global__ void __launch_bounds__(256) k_cost_prefix(MapView m, uint32_t* __restrict__ prefix, int ppitch, int row_lo, int row_hi) {
  const int lane = threadIdx.x & 31;
  const int y = row_lo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (y > row_hi) return;
  const uint8_t* __restrict__ rowp = m.data + static_cast<size_t>(y) * m.pitch;
  uint32_t* __restrict__ dst = prefix + static_cast<long long>(y) * ppitch;
  uint32_t carry_sum = 0u, carry_l = 0u, carry_n = 0u;  // running totals of the row so far (masked to their fields when stored)
  for (int x0 = 0; x0 <= m.size_x; x0 += 32) {  // (<=: the entry at x = size_x closes the last window)
    const int x = x0 + lane;
    const uint32_t c = x < m.size_x ? __ldg(rowp + x) : 0u;
    const uint32_t is_l = c == static_cast<uint32_t>(kLethal) ? 1u : 0u, is_n = c == static_cast<uint32_t>(kNoInformation) ? 1u : 0u;
    uint32_t s = c, v = is_l | (is_n << 8);  // (a warp holds at most 32 of either: the two 8-bit counts cannot carry into each other)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t us = __shfl_up_sync(kFull, s, d), uv = __shfl_up_sync(kFull, v, d);
      if (lane >= d) s += us, v += uv;
    }
    const uint32_t ex_s = carry_sum + s - c, ex_l = carry_l + (v & 0xFFu) - is_l, ex_n = carry_n + (v >> 8) - is_n;  // exclusive
    if (x <= m.size_x) dst[x] = ((ex_n & 0xFFu) << 24) | ((ex_l & 0xFFu) << 16) | (ex_s & 0xFFFFu);
    const uint32_t ts = __shfl_sync(kFull, s, 31), tv = __shfl_sync(kFull, v, 31);
    carry_sum += ts, carry_l += tv & 0xFFu, carry_n += tv >> 8;
  }
}

/// is_free walk of one group over the bit planes (every needed cell inside the map): LANE j owns span j of the
/// shape.  For the group's 64 candidate x offsets it fetches the 64-bit window of plane log2(p) that starts at the
/// row's first cell, and the one that starts `width - p` cells later (p = largest power of two <= width); OR-ing the
/// lanes' results (one `redux.or` per half) leaves bit `off` = "the pose at offset `off` hits a lethal cell", which
/// every pose then reads at its own offset.  ~30 warp instructions per 64 poses, whatever the number of rows.
__device__ __forceinline__ void lattice_rows_planes(const PlaneView& pv, const int* __restrict__ span_a, const uint32_t* __restrict__ span_m, int n_spans,
                                                    int pitch, int x_base, int y_base, int lane, bool third, int off_a, int off_b, uint32_t& hit_a,
                                                    uint32_t& hit_b) {
  uint32_t r_lo = 0u, r_hi = 0u;
  for (int j = lane; j < n_spans; j += 32) {
    const uint32_t sm = span_m[j];
    const int width = span_width(sm), dy = span_dy(sm);
    if (width == 0) continue;
    const int k = 31 - __clz(width), tail = width - (1 << k);
    const int xs = x_base + (span_a[j] - dy * pitch);
    const uint32_t* __restrict__ rowp = pv.bits + k * pv.plane_words + (y_base + dy);
    {
      const uint32_t* __restrict__ q = rowp + static_cast<long long>(xs >> 5) * pv.ypitch;
      const uint32_t w0 = __ldg(q), w1 = __ldg(q + pv.ypitch), w2 = third ? __ldg(q + 2 * pv.ypitch) : 0u;
      r_lo |= __funnelshift_r(w0, w1, xs & 31);
      r_hi |= __funnelshift_r(w1, w2, xs & 31);
    }
    if (tail) {
      const int xt = xs + tail;
      const uint32_t* __restrict__ q = rowp + static_cast<long long>(xt >> 5) * pv.ypitch;
      const uint32_t w0 = __ldg(q), w1 = __ldg(q + pv.ypitch), w2 = third ? __ldg(q + 2 * pv.ypitch) : 0u;
      r_lo |= __funnelshift_r(w0, w1, xt & 31);
      r_hi |= __funnelshift_r(w1, w2, xt & 31);
    }
  }
  r_lo = __reduce_or_sync(kFull, r_lo);
  r_hi = third ? __reduce_or_sync(kFull, r_hi) : 0u;
  hit_a = ((off_a < 32 ? r_lo : r_hi) >> (off_a & 31)) & 1u;
  hit_b = ((off_b < 32 ? r_lo : r_hi) >> (off_b & 31)) & 1u;
}

/// 32 cost bytes (two 16-byte vectors) -> one word of the marked-cell bitmap: bit i = byte i is LETHAL (or INSCRIBED, if
/// `inflated`).  Per 4 cells: a per-byte compare (0xFF / 0x00), one bit of each byte picked out, a multiply gathers the four.
__device__ __forceinline__ uint32_t mark_bits32(const uint4& a, const uint4& b, bool inflated) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t word = 0u;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t mk = __vcmpeq4(w[i], 0xFEFEFEFEu);
    if (inflated) mk |= __vcmpeq4(w[i], 0xFDFDFDFDu);
    word |= (((mk & 0x08040201u) * 0x01010101u) >> 24) << (4 * i);
  }
  return word;
}

/// Builds rows [row_lo, row_hi] of the lethal bit planes.  One warp per (32 map rows, kPlaneColsPerWarp word columns):
/// LANE = map row, so that the plane words of a column - consecutive rows - leave as one coalesced store (the planes
/// are column-major).  Every lane reads its row 32 bytes at a time and keeps the next word in hand: the doubling
/// V |= V >> 2^k on the 64-bit value L[c] | L[c + 1] << 32 yields plane k + 1 in its low word.
/// Start-of-call housekeeping a kernel can do on the side: empty the deferred-pose list {pushed, capacity}, zero the
/// caller's not-OK counter (header == nullptr: nothing to do).
struct CallReset {
  unsigned int* header;
  unsigned int capacity;
  unsigned long long* not_ok;
};

template <int kPlaneColsPerWarp>
__global__ void __launch_bounds__(256) k_lethal_planes(MapView m, uint32_t* __restrict__ planes, int ypitch, long long plane_words, int row_lo,
                                                        int row_hi, bool inflated, CallReset reset) {
  if (reset.header && blockIdx.x == 0 && threadIdx.x == 0) {  // (the sweep that reads both comes after this kernel in stream order)
    reset.header[0] = 0u, reset.header[1] = reset.capacity;
    if (reset.not_ok) *reset.not_ok = 0ull;
  }
  const int lane = threadIdx.x & 31;
  const int n_cols = (m.size_x + 31) / 32;
  const int col_groups = (n_cols + kPlaneColsPerWarp - 1) / kPlaneColsPerWarp;
  const long long n_tasks = static_cast<long long>((row_hi - row_lo) / 32 + 1) * col_groups;
  const int vec_per_row = m.pitch >> 4;  // 16-byte vectors per map row (the pitch is a multiple of 16)
  for (long long task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); task < n_tasks; task += static_cast<long long>(gridDim.x) * (blockDim.x >> 5)) {
    const int y = row_lo + static_cast<int>(task / col_groups) * 32 + lane;
    const int c0 = static_cast<int>(task % col_groups) * kPlaneColsPerWarp;
    const bool row_ok = y <= row_hi;
    const uint4* __restrict__ rowv = reinterpret_cast<const uint4*>(m.data + static_cast<size_t>(row_ok ? y : row_lo) * m.pitch);
    auto word_at = [&](int c) {  // marked cells 32 c .. 32 c + 31 of the row (cells past the map edge: 0)
      const int v = 2 * c;
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      const uint4 a = v < vec_per_row ? __ldg(rowv + v) : zero, b = v + 1 < vec_per_row ? __ldg(rowv + v + 1) : zero;
      uint32_t word = mark_bits32(a, b, inflated);
      const int in_map = m.size_x - 32 * c;
      if (in_map < 32) word &= in_map > 0 ? (1u << in_map) - 1u : 0u;
      return word;
    };
    uint32_t cur = word_at(c0);
#pragma unroll
    for (int cc = 0; cc < kPlaneColsPerWarp; ++cc) {
      const int c = c0 + cc;
      if (c >= n_cols) break;
      const uint32_t next = word_at(c + 1);
      unsigned long long v = static_cast<unsigned long long>(cur) | (static_cast<unsigned long long>(next) << 32);
      uint32_t* __restrict__ dst = planes + static_cast<long long>(c) * ypitch + y;
      if (row_ok) {
#pragma unroll
        for (int k = 0; k < kPlaneCount; ++k) {
          dst[k * plane_words] = static_cast<uint32_t>(v);
          v |= v >> (1 << k);
        }
      }
      cur = next;
    }
  }
}

/// Sink that feeds outline cells into a SlotWalker (single-lane walk).
struct LatSlotSink {
  SlotWalker<4>& w;
  __device__ __forceinline__ bool cell(int x, int y) {
    w.visit(x, y);
    return true;
  }
};

/// Sink of the OUTLINE rasteriser: maximal runs of outline cells that stay on one row become spans [min x, max x]
/// (Bresenham cells are 8-connected, so a same-row run is contiguous in x).  A verdict / maximum over the outline is
/// an any-reduction: order and duplicates do not matter.
struct OutlineSpanSink {
  uint32_t* t4;
  int cap, x0, y0;  // table capacity (spans), bounding-box origin
  bool ordered;     // accumulate_cost: spans are monotone pieces of the traversal, in traversal order, never merged
  int n = 0, y_run = 0, lo = 0, hi = 0, x_prev = 0, dir = 0;
  bool open = false, overflow = false;
  __device__ __forceinline__ void put(int a, int b, bool desc) {
    if (n >= cap) {
      overflow = true;
      return;
    }
    t4[n++] = static_cast<uint32_t>(a - x0) | (static_cast<uint32_t>(b - a + 1) << 12) | (static_cast<uint32_t>(y_run - y0) << 20) | (desc ? 1u << 28 : 0u);
  }
  __device__ __forceinline__ void close() {
    if (!open) return;
    if (ordered && dir < 0) {
      for (int x = hi; x >= lo; x -= kLatMaxWidth) put(max(lo, x - kLatMaxWidth + 1), x, true);  // chunks in traversal order
    } else {
      for (int x = lo; x <= hi; x += kLatMaxWidth) put(x, min(hi, x + kLatMaxWidth - 1), false);
    }
  }
  __device__ __forceinline__ bool cell(int x, int y) {
    bool extend = open && y == y_run;
    if (extend && ordered) {
      const int step = x - x_prev;
      extend = (step == 1 && dir >= 0) || (step == -1 && dir <= 0);
      if (extend) dir = step;
    }
    if (extend) {
      lo = min(lo, x), hi = max(hi, x);
    } else {
      close();
      open = true, y_run = y, lo = hi = x, dir = 0;
    }
    x_prev = x;
    return true;
  }
};

/// Rasterises the relative shape `verts` into the slot's span table `tbl` (span_a = tbl[0 .. cap), span_m =
/// tbl[cap .. 2 cap)) and returns its descriptor.  Runs in ONE lane.  First attempt: at most two ranges per scan row
/// (every convex shape), packed one word per row.  Second attempt (cap >= 2 * rows_cap, i.e. the polygon is not
/// convex): up to four ranges per row = two merged pairs, with the slot's whole table as the [row][4] scratch.
__device__ __noinline__ LatShape lattice_rasterise(const int2* __restrict__ verts, int nv, int rows_cap, int cap, int pitch, int* __restrict__ tbl,
                                                   int outline) {
  int* __restrict__ sa = tbl;
  uint32_t* __restrict__ sm = reinterpret_cast<uint32_t*>(tbl + cap);
  int x0 = 0x7FFFFFFF, y0 = 0x7FFFFFFF, x1 = -0x7FFFFFFF - 1, y1 = -0x7FFFFFFF - 1;
  for (int v = 0; v < nv; ++v) {
    x0 = min(x0, verts[v].x), x1 = max(x1, verts[v].x);
    y0 = min(y0, verts[v].y), y1 = max(y1, verts[v].y);
  }
  LatShape sh{x0, x1, y0, y1 - y0 + 1, ST_OK, 0};
  if (sh.rows > rows_cap || (x1 - x0) > 4000 || x0 < -30000 || x1 > 30000) {
    sh.status = kLatUnfit;
    return sh;
  }
  uint32_t* __restrict__ t4 = reinterpret_cast<uint32_t*>(tbl);
  // A merged pair [lo, hi] of row r becomes ceil(width / 32) spans, packed one word each (lo 12 bits | width 8 | row 8)
  // at t4[n++]; false when the table is full (`limit`: first index that must not be written yet).
  auto emit_pair = [&](int& n, int limit, int r, int lo, int hi) {
    for (int x = lo; x <= hi; x += kLatMaxWidth) {
      if (n >= limit) return false;
      t4[n++] = static_cast<uint32_t>(x) | (static_cast<uint32_t>(min(kLatMaxWidth, hi - x + 1)) << 12) | (static_cast<uint32_t>(r) << 20);
    }
    return true;
  };
  // Expands the n packed words downwards into span_a / span_m (n <= cap, so span_m[k] never lands on an unread word).
  auto expand = [&](int n) {
    for (int k = n - 1; k >= 0; --k) {
      const uint32_t pk = t4[k];
      const int lo = pk & 0xFFFu, width = (pk >> 12) & 0xFFu, r = (pk >> 20) & 0xFFu;
      sm[k] = static_cast<uint32_t>(width) | ((pk >> 28) << 7) | (static_cast<uint32_t>(r) << 8);
      sa[k] = r * pitch + x0 + lo;  // byte offset of the span's first cell from (vertex 0, row yb)
    }
    sh.n_spans = n;
  };
  if (outline) {  // outline_generator cells (generators.cpp:54-93) as same-row runs (2: monotone pieces in traversal order)
    OutlineSpanSink sink{t4, cap, x0, y0, outline == 2};
    walk_outline_cells(verts, nv, sink);
    sink.close();
    if (sink.overflow) {
      sh.status = kLatUnfit;
      return sh;
    }
    expand(sink.n);
    return sh;
  }
  {
    PackedStore store{sm, 1, x0, y0};  // two ranges per row, merged on arrival: one packed word per row in span_m's half
    store.clear(sh.rows);
    RunDetector<PackedStore> runs(store, 0, x0, x1);
    walk_outline_cells(verts, nv, runs);
    runs.finish();
    if (!store.overflow) {
      int n = 0;
      for (int r = 0; r < sh.rows; ++r) {
        const uint32_t w = sm[r];
        const uint32_t cnt = w >> 30;
        if (cnt == 1u) {  // odd number of ranges: std::logic_error upstream
          sh.status = ST_LOGIC_ERROR;
          return sh;
        }
        if (cnt == 2u && !emit_pair(n, cap, r, static_cast<int>(w & 0x7FFFu), static_cast<int>((w >> 15) & 0x7FFFu))) {
          sh.status = kLatUnfit;
          return sh;
        }
      }
      expand(n);
      return sh;
    }
  }
  if (cap < 2 * rows_cap) {
    sh.status = kLatUnfit;
    return sh;
  }
  // ---- four ranges per row
  SlotWalker<4> w;
  w.tab = t4, w.stride = 1, w.xb = x0, w.yb = y0;
  for (int r = 0; r < sh.rows; ++r)
    for (int q = 0; q < 4; ++q) w.slot(r, q) = kPairEmpty;
  const int2 f = verts[0];
  w.y_run = w.y0 = f.y, w.x_first = w.x0 = w.x_prev = f.x, w.y_before = f.y;
  w.in_first = true, w.overflow = false, w.f_xlast = f.x, w.f_ynext = f.y;
  LatSlotSink sink{w};
  walk_outline_cells(verts, nv, sink);
  if (w.in_first) {  // every cell on one row: ranges (min,min),(max,max) (generators.cpp:197-204)
    const uint32_t wd = static_cast<uint32_t>(x1 - x0);
    w.slot(w.y0 - y0, 0) = 0u;
    w.slot(w.y0 - y0, 1) = wd | (wd << 16);
  } else if (w.y_run == w.y0) {  // the tail run continues into cell 0
    w.prepend(w.y0, w.x_first, w.f_xlast, w.y_before, w.f_ynext);
  } else {
    w.append(w.y_run, w.x_first, w.x_prev, w.y_before, w.y0);
    w.prepend(w.y0, w.x0, w.f_xlast, w.y_run, w.f_ynext);
  }
  if (w.overflow) {  // more than four ranges on a row: the general kernel's sorted lists
    sh.status = kLatUnfit;
    return sh;
  }
  // pass 1: the merged pairs as packed spans, written over the scratch of rows already read (never past row r's own)
  int n = 0;
  for (int r = 0; r < sh.rows; ++r) {
    uint32_t e[4];
    for (int q = 0; q < 4; ++q) e[q] = w.slot(r, q);
    const int cnt = (e[0] != kPairEmpty) + (e[1] != kPairEmpty) + (e[2] != kPairEmpty) + (e[3] != kPairEmpty);
    if (cnt & 1) {
      sh.status = ST_LOGIC_ERROR;  // std::logic_error("Even number of line-pairs expected")
      return sh;
    }
    int lo[2], hi[2], pairs = 0;
    if (cnt == 4) {  // stable insertion sort by `min` (low 16 bits), then [r0.min, r1.max], [r2.min, r3.max]
      for (int a2 = 1; a2 < 4; ++a2)
        for (int b2 = a2; b2 > 0; --b2)
          if ((e[b2] & 0xFFFFu) < (e[b2 - 1] & 0xFFFFu)) {
            const uint32_t t = e[b2];
            e[b2] = e[b2 - 1], e[b2 - 1] = t;
          }
      lo[0] = e[0] & 0xFFFFu, hi[0] = e[1] >> 16, lo[1] = e[2] & 0xFFFFu, hi[1] = e[3] >> 16, pairs = 2;
    } else if (cnt == 2) {
      const int lo1 = e[0] & 0xFFFFu, hi1 = e[0] >> 16, lo2 = e[1] & 0xFFFFu, hi2 = e[1] >> 16;
      lo[0] = min(lo1, lo2), hi[0] = lo2 < lo1 ? hi1 : hi2, pairs = 1;  // stable sort by min, then [first.min, second.max]
    }
    for (int q = 0; q < pairs; ++q)
      if (!emit_pair(n, min(cap, 4 * (r + 1)), r, lo[q], hi[q])) {
        sh.status = kLatUnfit;
        return sh;
      }
  }
  expand(n);
  return sh;
}


/// accumulate_cost / max_cost for one group of the lattice kernel (all needed cells inside the map).  Per scan
/// row the warp loads the 96-byte row segment (one byte per lane and piece, coalesced) and turns it into a
/// structure every pose can query in O(1) for ITS window [off, off + width):
///   accumulate_cost: exclusive prefix sums of (cost, #cells >= 254) packed in one word (warp scans) — a window
///                    without LETHAL / NO_INFORMATION cells costs two shared-memory reads and a subtraction; a
///                    window with one walks its bytes in x order (rows ascend, x ascends: the reference's "first
///                    negative value in generator order" rule);
///   max_cost:        sliding-window maxima by doubling (level k holds max over 2^k cells, built with shuffles),
///                    window max = max(L_k[off], L_k[off + width - 2^k]) with 2^k <= width.
template <int OP>
__device__ __forceinline__ void lattice_rows_fold(const MapView& m, const int* __restrict__ row_a, const uint32_t* __restrict__ row_m, int rows,
                                                  int x_base, int y_base, int lane, bool third, int off_a, int off_b, uint8_t* rowbuf,
                                                  Fold<OP>& fa, Fold<OP>& fb) {
  uint32_t* table = reinterpret_cast<uint32_t*>(rowbuf);  // 2 x 100 words: prefix tables (accumulate) / window-max level (max)
  if constexpr (OP == OP_ACCUMULATE) {
    bool more_a = true, more_b = true;
    // lane l owns the three consecutive cells 3l .. 3l+2 of the segment: one warp scan of the per-lane totals
    // gives all 97 exclusive prefixes.  Packed value: low 16 bits cost (<= 96 * 255 in total), high bits the
    // number of cells >= 254.
    const uint8_t* __restrict__ p3 = m.data + static_cast<size_t>(y_base) * m.pitch + (x_base + 3 * lane);
    const int n_cells = third ? 96 : 64;  // cells 64.. are never part of a window when no offset reaches 32
    uint32_t sum_a = 0, sum_b = 0;
    // two scan rows per trip: their load -> scan -> table chains are independent and overlap
    for (int r = 0; r < rows; r += 2) {
      int width[2];
      uint32_t v0[2], v1[2], v2[2], s[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const bool live = r + u < rows;
        width[u] = live ? span_width(row_m[r + u]) : 0;
        const uint8_t* __restrict__ q = p3 + (live ? row_a[r + u] : 0);
        v0[u] = v1[u] = v2[u] = 0;
        if (live && 3 * lane < n_cells) v0[u] = __ldg(q), v1[u] = __ldg(q + 1), v2[u] = __ldg(q + 2);  // 64 = 3 * 21 + 1: lane 21 reads two cells more, still inside the segment
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        v0[u] |= ((v0[u] + 2u) << 8) & 0x10000u;
        v1[u] |= ((v1[u] + 2u) << 8) & 0x10000u;
        v2[u] |= ((v2[u] + 2u) << 8) & 0x10000u;
        s[u] = v0[u] + v1[u] + v2[u];
      }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t ua = __shfl_up_sync(kFull, s[0], d), ub = __shfl_up_sync(kFull, s[1], d);
        if (lane >= d) s[0] += ua, s[1] += ub;
      }
      __syncwarp();  // the previous trip's queries are done
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t* t = table + u * 100;
        const uint32_t e = s[u] - (v0[u] + v1[u] + v2[u]);
        t[3 * lane] = e, t[3 * lane + 1] = e + v0[u], t[3 * lane + 2] = e + v0[u] + v1[u];
        if (lane == 31) t[96] = s[u];
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t* t = table + u * 100;
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
          const int off = plane ? off_b : off_a;
          Fold<OP>& f = plane ? fb : fa;
          bool& more = plane ? more_b : more_a;
          uint32_t& sum = plane ? sum_b : sum_a;
          const uint32_t d = t[off + width[u]] - t[off];
          sum += d;  // (meaningless once a sentinel was met: `early` decides then)
          if ((d >> 16) != 0u && more) {  // the window holds a LETHAL / NO_INFORMATION cell: the first one in x order ends the walk
            const uint8_t* __restrict__ w = m.data + static_cast<size_t>(y_base) * m.pitch + x_base + row_a[r + u];
            const bool desc = span_descending(row_m[r + u]);
            for (int t = 0; t < width[u]; ++t) {  // in traversal order (areas: ascending x)
              const int c = __ldg(w + off + (desc ? width[u] - 1 - t : t));
              if (c >= kLethal) {
                f.early = c == kLethal ? -1 : -2;
                break;
              }
            }
            more = false;
          }
        }
      }
    }
    fa.sum += sum_a, fb.sum += sum_b;
  } else {  // OP_MAX
    const uint8_t* __restrict__ p = m.data + static_cast<size_t>(y_base) * m.pitch + (x_base + lane);
    for (int r = 0; r < rows; r += 2) {  // two scan rows per trip (independent chains)
      int width[2], k[2];
      uint32_t l0[2], l1[2], l2[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const bool live = r + u < rows;
        width[u] = live ? span_width(row_m[r + u]) : 0;
        k[u] = 31 - __clz(width[u] | 1);  // 2^k <= width
        const uint8_t* __restrict__ q = p + (live ? row_a[r + u] : 0);
        l0[u] = __ldg(q), l1[u] = __ldg(q + 32), l2[u] = third ? __ldg(q + 64) : 0u;
      }
      const int k_max = max(k[0], k[1]);
      for (int j = 0; j < k_max; ++j) {  // level j+1: max over 2^(j+1) cells starting at each position
        const int d = 1 << j;
        const bool wraps = lane + d >= 32;  // the partner cell lies in the next 32-cell piece
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t n0 = __shfl_down_sync(kFull, l0[u], d), n1 = __shfl_down_sync(kFull, l1[u], d), n2 = __shfl_down_sync(kFull, l2[u], d);
          const uint32_t w1 = __shfl_sync(kFull, l1[u], (lane + d) & 31), w2 = __shfl_sync(kFull, l2[u], (lane + d) & 31);
          if (j < k[u]) {
            l0[u] = max(l0[u], wraps ? w1 : n0);
            l1[u] = max(l1[u], wraps ? w2 : n1);
            l2[u] = max(l2[u], wraps ? 0u : n2);
          }
        }
      }
      __syncwarp();  // the previous trip's queries are done
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t* t = table + u * 100;
        t[lane] = l0[u], t[32 + lane] = l1[u], t[64 + lane] = l2[u];
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (width[u] == 0) continue;
        const uint32_t* t = table + u * 100;
        const int tail = width[u] - (1 << k[u]);
        fa.best4 = max(fa.best4, max(t[off_a], t[off_a + tail]));
        fb.best4 = max(fb.best4, max(t[off_b], t[off_b + tail]));
      }
    }
  }
}

/// Bit-packed verdicts of one lattice row of a unit: lane l holds poses `base + l` (plane A) and `base + 32 + l`
/// (plane B), so the 64 verdicts are two ballot words, shifted to the row's bit offset and OR-ed into three output
/// words - by lanes 0..2, one word each (the words are shared with neighbouring units and groups, hence atomic).
__device__ __forceinline__ void lat_store_bits(const OutView& out, uint32_t free_a, uint32_t free_b, long long base, int lane) {
  const int sh = static_cast<int>(base & 31);
  const uint32_t word = lane == 0 ? free_a << sh : (lane == 1 ? __funnelshift_l(free_a, free_b, sh) : __funnelshift_l(free_b, 0u, sh));
  // (arithmetic shift: base may be negative for a row that straddles `first` - its bits below pose 0 are zero)
  if (lane < 3 && word != 0u) atomicOr(out.free_bits + (base >> 5) + lane, word);
}

/// LEAN = verdicts of a polygon with at most 4 vertices over the bit planes (the headline case): the shape key is
/// exact, scan rows are decoded once per slot, and everything unusual (map border, shapes taller than 32 rows) goes
/// to the todo list - no hashed keys, no byte walk, no out-of-line calls on the hot path.
template <int OP, bool LEAN>
__global__ void __launch_bounds__(kLatThreads, LEAN ? LAT_LEAN_BLOCKS : LAT_MIN_BLOCKS) k_area_lattice(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src,
                                                                     long long count, OutView out, int rows_cap, int span_cap,
                                                                     unsigned int* __restrict__ todo, unsigned int* __restrict__ todo_count,
                                                                     LatGeom g, PlaneView pv, LatGlobal gl) {
  extern __shared__ double smem_d[];
  __shared__ unsigned long long s_round;
  __shared__ int s_used;  // occupied slots of the shared-memory cache
  const double* poly = stage_polygon(poly_g, nv, smem_d);
  unsigned long long* slot_key = reinterpret_cast<unsigned long long*>(smem_d + 2 * nv);  // 0 = empty
  const int kLatSlots = LEAN ? kLatSlotsMax : g.n_slots;  // (lean shapes are small: always 64 slots, a compile-time constant)
  int2* canon = reinterpret_cast<int2*>(slot_key + kLatSlots);                          // [slot][vertex] relative cells
  int* span_tbl = reinterpret_cast<int*>(canon + kLatSlots * nv);                       // [slot][2 * span_cap]: span_a, then span_m
  LatShape* meta = reinterpret_cast<LatShape*>(span_tbl + kLatSlots * 2 * span_cap);    // [slot]
  uint8_t* rowbufs = reinterpret_cast<uint8_t*>(meta + kLatSlots);                      // [warp][800]: per-row query tables of the value folds
  if (threadIdx.x < kLatSlots) {
    slot_key[threadIdx.x] = 0ull;
    meta[threadIdx.x].status = kLatPending;
  }
  if (threadIdx.x == 0) s_used = 0;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* rowbuf = rowbufs + warp * 800;
  const bool exact_key = LEAN || nv <= 4;  // 2 x 3 offsets of 8 bits: the key IS the shape, no verification pass
  const int kLatRows = g.rows_per_unit;
  const unsigned kLatRowMask = kLatRows >= 32 ? 0xFFFFFFFFu : (1u << kLatRows) - 1u;
  const long long unit_end = g.n_units;
  int chunk_left = 0;  // (thread 0) rounds left in the stretch drawn last

  // Rounds (one unit per warp) are handed out dynamically: CTAs that draw expensive units - axis-aligned headings,
  // where rounding noise splits a unit into several column groups - simply draw fewer of them.
  for (;;) {
    __syncthreads();  // everybody is done with the previous round (s_round, and the cache if it is cleared below)
    const bool start_over = s_used > (kLatSlots * 3) / 4;  // the shape cache is nearly full (CTA-uniform)
    if (threadIdx.x == 0) {
      // Guided hand-out: a CTA draws a stretch of consecutive rounds (same heading: the shapes stay in its cache, the
      // planes in its L1) whose length shrinks from `static_rounds` to one as the batch runs out, which evens out the
      // finish however uneven the units are.
      if (chunk_left == 0) {
        // ticket t -> stretch: the first tickets_a tickets are stretches of `static_rounds` rounds, the next tickets_b
        // of half that, the rest single rounds (one atomicAdd per draw: a compare-and-swap loop on one counter
        // serialises 444 CTAs)
        const long long t = static_cast<long long>(atomicAdd(gl.next_round, 1ull));
        const long long la = g.static_rounds, lb = max(1, g.static_rounds / 2);
        long long first_round, len;
        if (t < g.tickets_a) first_round = t * la, len = la;
        else if (t < g.tickets_a + g.tickets_b) first_round = g.tickets_a * la + (t - g.tickets_a) * lb, len = lb;
        else first_round = g.tickets_a * la + g.tickets_b * lb + (t - g.tickets_a - g.tickets_b), len = 1;
        s_round = static_cast<unsigned long long>(first_round) * kLatWarps, chunk_left = static_cast<int>(len);
      } else {
        s_round += kLatWarps;
      }
      --chunk_left;
    }
    if (start_over) {
      __syncthreads();
      if (threadIdx.x < kLatSlots) {
        slot_key[threadIdx.x] = 0ull;
        meta[threadIdx.x].status = kLatPending;
      }
      if (threadIdx.x == 0) s_used = 0;
    }
    __syncthreads();
    const long long round = static_cast<long long>(s_round);
    if (round >= unit_end) break;
    const long long unit = round + warp;
    if (unit >= unit_end) continue;
    const int group = static_cast<int>(unit / g.segs), seg = static_cast<int>(unit - static_cast<long long>(group) * g.segs);
    const int ix_a = seg * kLatSegX + lane, ix_b = ix_a + 32;
    auto column_tx = [&src](int ix) { return __dadd_rn(src.x0, __dmul_rn(static_cast<double>(ix), src.dx)); };
    auto heading_cs = [&src](int it) { return __ldg(reinterpret_cast<const double2*>(src.theta_cs) + it); };

    // y-profiles of the unit's 8 rows: lane r owns row r (lanes 8.. mirror rows 0..7)
    int my_it = 0, my_iy = 0;
    LatProfile my_y{0, 0ull, false, false};
    bool my_valid = false;
    {
      const int row_rel = group * kLatRows + (lane & (kLatRows - 1));
      const long long row = g.row_first + row_rel;
      if (row_rel < g.n_rows) {
        my_valid = true;
        my_it = static_cast<int>(row / src.ny);
        my_iy = static_cast<int>(row - static_cast<long long>(my_it) * src.ny);
        const double2 cs = heading_cs(my_it);
        my_y = lat_profile<false>(poly, nv, cs.x, cs.y, __dadd_rn(src.y0, __dmul_rn(static_cast<double>(my_iy), src.dy)), m, exact_key);
      }
    }
    // "row r has exactly the y offsets of row r - 1" (same heading): the shape slot of a column group carries over
    bool my_same_prev;
    {
      const int prev_it = __shfl_up_sync(kFull, my_it, 1);
      const unsigned long long prev_key = __shfl_up_sync(kFull, my_y.key, 1);
      const bool prev_good = __shfl_up_sync(kFull, static_cast<int>(my_valid && my_y.ok && my_y.packable), 1) != 0;
      my_same_prev = (lane & (kLatRows - 1)) != 0 && my_valid && my_y.ok && my_y.packable && prev_good && prev_it == my_it && prev_key == my_y.key;
      if (!exact_key) {  // the key is only a hash of the offsets: compare them
        const double2 cs = heading_cs(my_it);
        const double ty = __dadd_rn(src.y0, __dmul_rn(static_cast<double>(my_iy), src.dy));
        for (int v = 1; v < nv; ++v) {
          int cy;
          lat_cell_y(cs.x, cs.y, ty, poly[2 * v], poly[2 * v + 1], m, cy);
          const int d = cy - my_y.c0, prev_d = __shfl_up_sync(kFull, d, 1);
          my_same_prev &= d == prev_d;
        }
      }
    }
    const unsigned same_rows = __ballot_sync(kFull, my_same_prev) & kLatRowMask;  // bit r: row r repeats row r - 1
    const int rows_here = min(kLatRows, g.n_rows - group * kLatRows);
    const long long i_base = (g.row_first + static_cast<long long>(group) * kLatRows) * src.nx - src.first;  // pose index of (row 0, column 0)
    const bool col_a = ix_a < src.nx, col_b = ix_b < src.nx;
    // the usual unit lies inside the batch [0, count): no per-pose range tests
    const bool whole_unit = i_base >= 0 && i_base + static_cast<long long>(rows_here) * src.nx <= count;

    // runs of rows with the same heading (one per unit; two where the unit crosses a heading boundary)
    for (int r0 = 0; r0 < rows_here;) {
      const int it = __shfl_sync(kFull, my_it, r0);
      // rows r0 .. r1 - 1 share the heading
      const unsigned it_rows = (__ballot_sync(kFull, my_it == it) & kLatRowMask) >> r0;
      const int run = __ffs(~it_rows);  // (0: every row up to bit 31 shares the heading)
      const int r1 = min(rows_here, r0 + (run ? run - 1 : 32));
      const double2 cs = heading_cs(it);
      // x-profiles depend on the heading and the column only
      const LatProfile xa = lat_profile<true>(poly, nv, cs.x, cs.y, column_tx(ix_a), m, exact_key);
      const LatProfile xb = lat_profile<true>(poly, nv, cs.x, cs.y, column_tx(ix_b), m, exact_key);
      const bool good_a = col_a && xa.ok && xa.packable, good_b = col_b && xb.ok && xb.packable;
      const bool row_good = my_valid && my_y.ok && my_y.packable;  // (lane r: row r)
      const unsigned bad_rows = __ballot_sync(kFull, my_valid && !row_good) & kLatRowMask;

      // ---- rare: coordinates out of range (status), offsets that do not fit the packed key (general kernel) ----
      if (bad_rows != 0u || __any_sync(kFull, (col_a && !good_a) || (col_b && !good_b))) {
        for (int r = r0; r < r1; ++r) {
          const bool y_ok = __shfl_sync(kFull, static_cast<int>(my_y.ok), r) != 0, y_good = ((bad_rows >> r) & 1u) == 0u;
          const long long i_a = i_base + static_cast<long long>(r) * src.nx + ix_a, i_b = i_a + 32;
          if (col_a && i_a >= 0 && i_a < count && !(good_a && y_good)) {
            if (!(xa.ok && y_ok)) store_failure(out, i_a, ST_COORD_RANGE);
            else push_todo(todo, todo_count, i_a);
          }
          if (col_b && i_b >= 0 && i_b < count && !(good_b && y_good)) {
            if (!(xb.ok && y_ok)) store_failure(out, i_b, ST_COORD_RANGE);
            else push_todo(todo, todo_count, i_b);
          }
        }
      }

      // ---- column groups: columns with the same x offsets, at most 64 cells apart ------------------------
      unsigned pend_a = __ballot_sync(kFull, good_a), pend_b = __ballot_sync(kFull, good_b);
      while (pend_a | pend_b) {
        const bool from_a = pend_a != 0u;
        const int leader = __ffs(from_a ? pend_a : pend_b) - 1;
        const unsigned long long kx = __shfl_sync(kFull, from_a ? xa.key : xb.key, leader);
        bool same_a = ((pend_a >> lane) & 1u) && xa.key == kx, same_b = ((pend_b >> lane) & 1u) && xb.key == kx;
        const double tx_lead = __shfl_sync(kFull, column_tx(from_a ? ix_a : ix_b), leader);
        const int c0_lead = __shfl_sync(kFull, from_a ? xa.c0 : xb.c0, leader);
        if (!exact_key) {  // equal hashes only propose the group: compare the offsets with the leader's
          const double tx_a = column_tx(ix_a), tx_b = column_tx(ix_b);
          for (int v = 1; v < nv; ++v) {
            int cl, ca, cb;
            lat_cell_x(cs.x, cs.y, tx_lead, poly[2 * v], poly[2 * v + 1], m, cl);
            lat_cell_x(cs.x, cs.y, tx_a, poly[2 * v], poly[2 * v + 1], m, ca);
            lat_cell_x(cs.x, cs.y, tx_b, poly[2 * v], poly[2 * v + 1], m, cb);
            same_a &= ca - xa.c0 == cl - c0_lead;
            same_b &= cb - xb.c0 == cl - c0_lead;
          }
        }
        const int xmin = __reduce_min_sync(kFull, min(same_a ? xa.c0 : 0x7FFFFFFF, same_b ? xb.c0 : 0x7FFFFFFF));
        const int off_a = xa.c0 - xmin, off_b = xb.c0 - xmin;
        const bool fits_a = same_a && off_a < 64, fits_b = same_b && off_b < 64;  // the column holding xmin always fits
        const unsigned grp_a = __ballot_sync(kFull, fits_a), grp_b = __ballot_sync(kFull, fits_b);
        const int maxoff = __reduce_max_sync(kFull, max(fits_a ? off_a : 0, fits_b ? off_b : 0));
        const int oa = fits_a ? off_a : 0, ob = fits_b ? off_b : 0;
        pend_a &= ~grp_a;
        pend_b &= ~grp_b;

        int slot = -1, st = ST_OK;  // the group's shape slot: carried from row to row while the y offsets repeat
        int xb_s = 0, x1_s = 0, rows = 0, yb_s = 0, n_spans = 0;  // its bounding box relative to vertex 0, its span count
        const int* span_a = span_tbl;
        const uint32_t* span_m = reinterpret_cast<const uint32_t*>(span_tbl);
        // OP_IS_FREE over the bit planes, lane j = scan row j of the slot's shape, decoded once per slot: the two
        // windows' word pointers for map row 0 (the lattice row only adds y_base) and bit shifts
        const uint32_t* win_s = nullptr;
        const uint32_t* win_t = nullptr;
        int sh_s = 0, sh_t = 0;
        bool lean = false, inside_x = false;
        long long i_a = i_base + static_cast<long long>(r0) * src.nx + ix_a;
        for (int r = r0; r < r1; ++r, i_a += src.nx) {
          if ((bad_rows >> r) & 1u) {
            slot = -1;
            continue;
          }
          const long long i_b = i_a + 32;
          bool act_a = fits_a, act_b = fits_b;
          if (!whole_unit) {  // (a batch may start / end inside the row)
            act_a = fits_a && i_a >= 0 && i_a < count, act_b = fits_b && i_b >= 0 && i_b < count;
            if (!__any_sync(kFull, act_a || act_b)) {
              slot = -1;
              continue;
            }
          }
          const int vy0 = __shfl_sync(kFull, my_y.c0, r);
          if (!(r > r0 && slot >= 0 && ((same_rows >> r) & 1u) != 0u)) {
            // ---- shape cache: find, or insert + rasterise ----------------------------------------------
            const unsigned long long ykey = __shfl_sync(kFull, my_y.key, r);
            const unsigned long long kl = exact_key ? ((1ull << 63) | (kx << 24) | ykey) : ((lat_mix(kx, static_cast<int>(ykey)) ^ (ykey >> 32)) | 1ull);
            const int iy = __shfl_sync(kFull, my_iy, r);
            const double ty = __dadd_rn(src.y0, __dmul_rn(static_cast<double>(iy), src.dy));
            int sfound = -1, fill = 0, gidx = -1;  // fill: 0 = nothing to do, 1 = rasterise (and publish at gidx), 2 = copy from gidx
            if (lane == leader) {
              const unsigned long long hash = kl * 0x9E3779B97F4A7C15ull;
              const int home = static_cast<int>(hash >> 58) & (kLatSlots - 1);
              for (int j = 0; j < kLatSlots && sfound < 0; ++j) {
                const int idx = (home + j) & (kLatSlots - 1);
                unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(&slot_key[idx]);
                if (old == 0ull) old = atomicCAS(&slot_key[idx], 0ull, kl);
                if (old == 0ull) {  // we own the slot: fill it from the launch-wide table, or rasterise
                  sfound = idx, fill = 1;
                  atomicAdd(&s_used, 1);
                  const int ghome = static_cast<int>(hash >> 32);
                  for (int q = 0; q < kLatGlobalProbes; ++q) {
                    const int gi = (ghome + q) & (gl.capacity - 1);
                    unsigned long long gold = *reinterpret_cast<volatile unsigned long long*>(&gl.key[gi]);
                    if (gold == 0ull) gold = atomicCAS(&gl.key[gi], 0ull, kl);
                    if (gold == 0ull) {  // first CTA to meet this shape: rasterise and publish
                      gidx = gi;
                      break;
                    }
                    if (gold == kl) {  // somebody has it (or is making it): copy
                      gidx = gi, fill = 2;
                      break;
                    }
                  }
                } else if (old == kl) {
                  sfound = idx;
                }
              }
            }
            sfound = __shfl_sync(kFull, sfound, leader);
            fill = __shfl_sync(kFull, fill, leader);
            gidx = __shfl_sync(kFull, gidx, leader);
            if (fill != 0) {
              int* const s_tbl = span_tbl + sfound * 2 * span_cap;
              int* const g_entry = gidx >= 0 ? gl.data + static_cast<long long>(gidx) * gl.entry_words : nullptr;
              int shape_status = ST_OK;
              if (fill == 2) {  // wait for the publisher, then copy entry -> slot with the whole warp
                int gs = 0;
                if (lane == 0) {
                  while ((gs = *reinterpret_cast<volatile int*>(&gl.status[gidx])) == 0) __nanosleep(200);
                }
                gs = __shfl_sync(kFull, gs, 0);
                __threadfence();
                shape_status = gs - 16;
                if (lane < 6 && lane != 4) reinterpret_cast<int*>(&meta[sfound])[lane] = __ldcg(g_entry + lane);  // (4 = status: published last)
                for (int w = lane; w < 2 * nv; w += 32) reinterpret_cast<int*>(canon + sfound * nv)[w] = __ldcg(g_entry + 6 + w);
                for (int w = lane; w < 2 * span_cap; w += 32) s_tbl[w] = __ldcg(g_entry + 6 + 2 * nv + w);
              } else {
                if (lane == leader) {
                  int ax = 0, ay = 0;
                  for (int v = 0; v < nv; ++v) {
                    int cx, cy;
                    lat_cell_x(cs.x, cs.y, tx_lead, poly[2 * v], poly[2 * v + 1], m, cx);
                    lat_cell_y(cs.x, cs.y, ty, poly[2 * v], poly[2 * v + 1], m, cy);
                    if (v == 0) ax = cx, ay = cy;
                    canon[sfound * nv + v] = make_int2(cx - ax, cy - ay);
                  }
                  const LatShape sh = lattice_rasterise(canon + sfound * nv, nv, rows_cap, span_cap, m.pitch, s_tbl, g.outline);
                  meta[sfound].xb = sh.xb, meta[sfound].x1 = sh.x1, meta[sfound].yb = sh.yb, meta[sfound].rows = sh.rows, meta[sfound].n_spans = sh.n_spans;
                  shape_status = sh.status;
                }
                shape_status = __shfl_sync(kFull, shape_status, leader);
                __syncwarp();
                if (g_entry) {  // publish: slot -> entry with the whole warp
                  if (lane < 6) g_entry[lane] = lane == 4 ? shape_status : reinterpret_cast<const int*>(&meta[sfound])[lane];
                  for (int w = lane; w < 2 * nv; w += 32) g_entry[6 + w] = reinterpret_cast<const int*>(canon + sfound * nv)[w];
                  for (int w = lane; w < 2 * span_cap; w += 32) g_entry[6 + 2 * nv + w] = s_tbl[w];
                  __threadfence();
                  __syncwarp();
                  if (lane == 0) *reinterpret_cast<volatile int*>(&gl.status[gidx]) = shape_status + 16;
                }
              }
              __threadfence_block();
              __syncwarp();
              if (lane == 0) *reinterpret_cast<volatile int*>(&meta[sfound].status) = shape_status;
            }
            slot = sfound;
            st = ST_OK;
            lean = false;
            if (slot >= 0) {  // wait for a shape another warp is still rasterising
              while ((st = *reinterpret_cast<volatile int*>(&meta[slot].status)) == kLatPending) {
                __nanosleep(100);  // (leave the issue slots to the lane that is rasterising the shape)
              }
              __threadfence_block();
              xb_s = meta[slot].xb, x1_s = meta[slot].x1, rows = meta[slot].rows, yb_s = meta[slot].yb, n_spans = meta[slot].n_spans;
              span_a = span_tbl + slot * 2 * span_cap, span_m = reinterpret_cast<const uint32_t*>(span_a + span_cap);
              if constexpr (OP == OP_IS_FREE) {
                lean = pv.bits != nullptr && n_spans <= 32 && st == ST_OK;  // lane j = span j
                // every cell any pose of the group needs lies in x: [xmin + xb, xmin + maxoff + x1]
                inside_x = xmin + xb_s >= 0 && xmin + maxoff + x1_s < m.size_x;
                if (lean) {
                  const uint32_t sm = lane < n_spans ? span_m[lane] : 0u;
                  const int width = span_width(sm), dy = span_dy(sm);
                  win_s = win_t = nullptr;
                  if (width > 0 && inside_x) {
                    const int k = 31 - __clz(width);
                    const int xs = xmin + (span_a[lane] - dy * m.pitch), xt = xs + (width - (1 << k));
                    const uint32_t* plane_row = pv.bits + k * pv.plane_words + dy;
                    win_s = plane_row + static_cast<long long>(xs >> 5) * pv.ypitch, sh_s = xs & 31;
                    if (xt != xs) win_t = plane_row + static_cast<long long>(xt >> 5) * pv.ypitch, sh_t = xt & 31;
                  }
                }
              }
              if (!exact_key) {  // the hash only proposed the shape: compare the actual relative vertices
                bool equal = true;
                for (int v = lane; v < nv; v += 32) {
                  int cx, cy;
                  lat_cell_x(cs.x, cs.y, tx_lead, poly[2 * v], poly[2 * v + 1], m, cx);
                  lat_cell_y(cs.x, cs.y, ty, poly[2 * v], poly[2 * v + 1], m, cy);
                  const int2 q = canon[slot * nv + v];
                  equal &= cx - c0_lead == q.x && cy - vy0 == q.y;
                }
                if (!__all_sync(kFull, equal)) slot = -1;
              }
            }
          }
          if (slot < 0 || st == kLatUnfit) {  // cache full, hash clash or a shape the packed table cannot hold
            if (act_a) push_todo(todo, todo_count, i_a);
            if (act_b) push_todo(todo, todo_count, i_b);
            continue;
          }
          if (st != ST_OK) {  // open outline: std::logic_error in the reference
            if (act_a) store_failure(out, i_a, st);
            if (act_b) store_failure(out, i_b, st);
            continue;
          }
          if constexpr (OP == OP_IS_FREE) {
            // ---- the usual case, as a tight loop over the rows that keep this slot: whole unit inside the batch, rows
            //      that repeat the shape, every needed cell inside the map
            if (lean && whole_unit && inside_x) {
              const unsigned chain = r < 31 ? (same_rows & ~bad_rows) >> (r + 1) : 0u;  // bit q: row r + 1 + q continues with this slot
              const int stop = __ffs(~chain);
              const int r2 = min(r1, r + (stop ? stop : 32));
              const bool third = maxoff >= 32;
              const bool store_free = out.is_free != nullptr, store_status = out.status != nullptr, store_bits = out.free_bits != nullptr;
              int q = r;
              for (; q < r2; ++q, i_a += src.nx) {
                const int y_base = __shfl_sync(kFull, my_y.c0, q) + yb_s;
                if (y_base < 0 || y_base + rows > m.size_y) break;  // map border: back to the general path for this row
                // lane j ORs the (up to) two power-of-two windows of scan row j; one redux per 32 offsets folds the rows
                uint32_t r_lo = 0u, r_hi = 0u;
                if (win_s) {
                  const uint32_t* __restrict__ p = win_s + y_base;
                  const uint32_t w0 = __ldg(p), w1 = __ldg(p + pv.ypitch), w2 = third ? __ldg(p + 2 * pv.ypitch) : 0u;
                  r_lo = __funnelshift_r(w0, w1, sh_s), r_hi = __funnelshift_r(w1, w2, sh_s);
                }
                if (win_t) {
                  const uint32_t* __restrict__ p = win_t + y_base;
                  const uint32_t w0 = __ldg(p), w1 = __ldg(p + pv.ypitch), w2 = third ? __ldg(p + 2 * pv.ypitch) : 0u;
                  r_lo |= __funnelshift_r(w0, w1, sh_t), r_hi |= __funnelshift_r(w1, w2, sh_t);
                }
                r_lo = __reduce_or_sync(kFull, r_lo);
                r_hi = third ? __reduce_or_sync(kFull, r_hi) : 0u;
                const uint32_t hit_a = ((oa < 32 ? r_lo : r_hi) >> (oa & 31)) & 1u, hit_b = ((ob < 32 ? r_lo : r_hi) >> (ob & 31)) & 1u;
                if (store_free) {
                  uint8_t* __restrict__ pf = out.is_free + i_a;
                  if (fits_a) pf[0] = static_cast<uint8_t>(hit_a ^ 1u);
                  if (fits_b) pf[32] = static_cast<uint8_t>(hit_b ^ 1u);
                }
                if (store_status) {
                  uint8_t* __restrict__ ps = out.status + i_a;
                  if (fits_a) ps[0] = ST_OK;
                  if (fits_b) ps[32] = ST_OK;
                }
                if (store_bits) lat_store_bits(out, __ballot_sync(kFull, fits_a && hit_a == 0u), __ballot_sync(kFull, fits_b && hit_b == 0u), i_a - lane, lane);
              }
              if (q > r) {  // rows r .. q - 1 are done; the for loop's increment moves on to row q
                r = q - 1, i_a -= src.nx;
                continue;
              }
            }
          }
          // ---- up to 64 x-adjacent poses per 96-cell bit window ----------------------------------------
          const int y_base = vy0 + yb_s;
          // every cell any pose of the group needs lies in x: [xmin + xb, xmin + maxoff + x1]
          const bool inside = y_base >= 0 && y_base + rows <= m.size_y && (LEAN ? inside_x : (xmin + xb_s >= 0 && xmin + maxoff + x1_s < m.size_x));
          if constexpr (OP == OP_IS_FREE) {
            uint32_t hit_a = 0, hit_b = 0;
            if (lean && inside) {
              // lane j ORs the (up to) two power-of-two windows of scan row j; one redux per 32 offsets folds the rows
              uint32_t r_lo = 0u, r_hi = 0u;
              const bool third = maxoff >= 32;
              if (win_s) {
                const uint32_t* __restrict__ q = win_s + y_base;
                const uint32_t w0 = __ldg(q), w1 = __ldg(q + pv.ypitch), w2 = third ? __ldg(q + 2 * pv.ypitch) : 0u;
                r_lo = __funnelshift_r(w0, w1, sh_s), r_hi = __funnelshift_r(w1, w2, sh_s);
              }
              if (win_t) {
                const uint32_t* __restrict__ q = win_t + y_base;
                const uint32_t w0 = __ldg(q), w1 = __ldg(q + pv.ypitch), w2 = third ? __ldg(q + 2 * pv.ypitch) : 0u;
                r_lo |= __funnelshift_r(w0, w1, sh_t), r_hi |= __funnelshift_r(w1, w2, sh_t);
              }
              r_lo = __reduce_or_sync(kFull, r_lo);
              r_hi = third ? __reduce_or_sync(kFull, r_hi) : 0u;
              hit_a = ((oa < 32 ? r_lo : r_hi) >> (oa & 31)) & 1u;
              hit_b = ((ob < 32 ? r_lo : r_hi) >> (ob & 31)) & 1u;
            } else if constexpr (LEAN) {
              if (act_a) push_todo(todo, todo_count, i_a);
              if (act_b) push_todo(todo, todo_count, i_b);
              continue;
            } else if (inside && pv.bits) lattice_rows_planes(pv, span_a, span_m, n_spans, m.pitch, xmin, y_base, lane, maxoff >= 32, oa, ob, hit_a, hit_b);
            else lattice_rows_bytes(m, span_a, span_m, n_spans, xmin, y_base, lane, oa, ob, hit_a, hit_b);
            if (act_a) {
              if (out.is_free) out.is_free[i_a] = hit_a ? 0 : 1;
              if (out.status) out.status[i_a] = ST_OK;
            }
            if (act_b) {
              if (out.is_free) out.is_free[i_b] = hit_b ? 0 : 1;
              if (out.status) out.status[i_b] = ST_OK;
            }
            if (out.free_bits) {
              lat_store_bits(out, __ballot_sync(kFull, act_a && hit_a == 0u), __ballot_sync(kFull, act_b && hit_b == 0u), i_a - lane, lane);
            }
          } else if (inside) {  // (discarded for OP_IS_FREE: the whole else-chain hangs off an `if constexpr`)
            Fold<OP> fa, fb;
            if constexpr (OP == OP_ACCUMULATE) {
              if (pv.prefix) lattice_spans_accumulate(m, pv, span_a, span_m, n_spans, xmin, y_base, oa, ob, fa, fb);
              else lattice_rows_fold<OP>(m, span_a, span_m, n_spans, xmin, y_base, lane, maxoff >= 32, oa, ob, rowbuf, fa, fb);
            } else {
              lattice_rows_fold<OP>(m, span_a, span_m, n_spans, xmin, y_base, lane, maxoff >= 32, oa, ob, rowbuf, fa, fb);
            }
            if (act_a) {
              fa.store(out, i_a);
              if (out.status) out.status[i_a] = ST_OK;
            }
            if (act_b) {
              fb.store(out, i_b);
              if (out.status) out.status[i_b] = ST_OK;
            }
          } else {  // a pose of the group touches the map border: the general kernel resolves out-of-map cells in order
            if (act_a) push_todo(todo, todo_count, i_a);
            if (act_b) push_todo(todo, todo_count, i_b);
          }
        }
      }
      r0 = r1;
    }
  }
}

struct ListScratch {
  int* cnt;
  int2* rng;
  int* seq;
  int rows_cap;
};

/// Validates the per-row lists and folds the merged pairs in generator order (rows ascending, pairs in
/// sorted order).  Returns the status.
template <class FoldT>
__device__ __forceinline__ int fold_list_rows(const ListStore& store, int rows, const MapView& m, FoldT& fold) {
  if (store.out_of_rows) return ST_COORD_RANGE;
  bool odd = false;
  for (int r = 0; r < rows; ++r) odd |= (store.cnt[static_cast<long long>(r) * store.nthreads + store.me] & 1) != 0;
  if (odd) return ST_LOGIC_ERROR;
  if (store.overflow) return ST_ROW_COMPLEXITY;
  for (int r = 0; r < rows; ++r) {
    const int n = store.cnt[static_cast<long long>(r) * store.nthreads + store.me];
    for (int k = 0; k < n; k += 2) {
      const int lo = store.rng[store.at(r, k)].x, hi = store.rng[store.at(r, k + 1)].y;
      if (!fold.span(m, store.yb + r, lo, hi)) return ST_OK;
    }
  }
  return ST_OK;
}

template <int OP>
__global__ void __launch_bounds__(kBlock) k_area_list(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src,
                                                     long long count, OutView out, ListScratch scratch,
                                                     const unsigned int* __restrict__ todo,
                                                     const unsigned int* __restrict__ todo_count) {
  extern __shared__ double smem_d[];
  const double* poly = stage_polygon(poly_g, nv, smem_d);
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long me = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  bool listed = false;
  const long long n_items = todo ? todo_items(todo_count, count, listed) : count;
  const bool none = src.format == POSE_NONE;
  for (long long item = me; item < n_items; item += nthreads) {
    const long long i = (todo && listed) ? static_cast<long long>(todo[item]) : item;
    const Pose T = load_pose(src, i);
    int x0, y0, x1, y1;
    const bool ok = none ? polygon_bbox<XF_NONE>(poly, nv, T, m, x0, y0, x1, y1) : polygon_bbox<XF_PATH_A>(poly, nv, T, m, x0, y0, x1, y1);
    if (!ok) {
      store_failure(out, i, ST_COORD_RANGE);
      continue;
    }
    Fold<OP> fold;
    int status = ST_OK;
    if (nv > 0) {
      const int rows = y1 - y0 + 1;
      if (rows > scratch.rows_cap) {
        status = ST_COORD_RANGE;
      } else {
        ListStore store{scratch.cnt, scratch.rng, scratch.seq, nthreads, me, y0, scratch.rows_cap};
        store.clear(rows);
        RunDetector<ListStore> runs(store, 0, x0, x1);
        if (none) walk_outline<XF_NONE>(poly, nv, T, m, runs);
        else walk_outline<XF_PATH_A>(poly, nv, T, m, runs);
        runs.finish();
        status = fold_list_rows(store, rows, m, fold);
      }
    }
    if (status != ST_OK) {
      store_failure(out, i, status);
    } else {
      fold.store(out, i);
      if (out.status) out.status[i] = ST_OK;
    }
  }
}


/// Area checks for any polygon, one WARP per pose (area_rows.cuh); also drains the poses the packed and
/// lattice kernels deferred (`todo`).  Persistent grid.
constexpr int kRowsWarps = 4;
template <int OP>
__global__ void __launch_bounds__(kRowsWarps * 32) k_area_rows(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src,
                                                              long long count, OutView out, const unsigned int* __restrict__ todo,
                                                              const unsigned int* __restrict__ todo_count) {
  extern __shared__ double smem_d[];
  const double* poly = stage_polygon(poly_g, nv, smem_d);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  RayRec* rays = reinterpret_cast<RayRec*>(smem_d + 2 * nv) + warp * (nv + 1);
  int2* verts = reinterpret_cast<int2*>(reinterpret_cast<RayRec*>(smem_d + 2 * nv) + kRowsWarps * (nv + 1)) + warp * nv;
  const long long n_warps = static_cast<long long>(gridDim.x) * kRowsWarps;
  bool listed = false;
  const long long n_items = todo ? todo_items(todo_count, count, listed) : count;
  const bool none = src.format == POSE_NONE;
  for (long long item = static_cast<long long>(blockIdx.x) * kRowsWarps + warp; item < n_items; item += n_warps) {
    const long long i = (todo && listed) ? static_cast<long long>(todo[item]) : item;
    const Pose T = load_pose(src, i);
    WarpFold<OP> wf;
    const int status = none ? area_rows_pose<OP, XF_NONE>(poly, nv, T, m, verts, rays, lane, wf)
                            : area_rows_pose<OP, XF_PATH_A>(poly, nv, T, m, verts, rays, lane, wf);
    if (lane == 0) {
      if (status != ST_OK) {
        store_failure(out, i, status);
      } else {
        wf.store(out, i);
        if (out.status) out.status[i] = ST_OK;
      }
    }
    __syncwarp();
  }
}


/// expand(area_generator(res, pose*polygon - origin)) — to_area (expand.hpp:75-80).  One warp per pose.
/// Pass 1 (cells == nullptr) fills counts[]; pass 2 writes the cells of pose i at cells + starts[i].
__global__ void __launch_bounds__(kRowsWarps * 32) k_area_expand(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src,
                                                                long long count, long long* __restrict__ counts, uint8_t* __restrict__ status,
                                                                const long long* __restrict__ starts, int2* __restrict__ cells) {
  extern __shared__ double smem_d[];
  const double* poly = stage_polygon(poly_g, nv, smem_d);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  RayRec* rays = reinterpret_cast<RayRec*>(smem_d + 2 * nv) + warp * (nv + 1);
  int2* verts = reinterpret_cast<int2*>(reinterpret_cast<RayRec*>(smem_d + 2 * nv) + kRowsWarps * (nv + 1)) + warp * nv;
  const long long n_warps = static_cast<long long>(gridDim.x) * kRowsWarps;
  const bool none = src.format == POSE_NONE;
  for (long long i = static_cast<long long>(blockIdx.x) * kRowsWarps + warp; i < count; i += n_warps) {
    const Pose T = load_pose(src, i);
    long long total = 0;
    int2* dst = cells ? cells + starts[i] : nullptr;
    if (cells && starts[i + 1] == starts[i]) dst = nullptr;  // nothing to write (empty or failed pose)
    const int st = (cells && !dst) ? ST_OK
                   : none          ? area_rows_expand<XF_NONE>(poly, nv, T, m, verts, rays, lane, dst, total)
                                   : area_rows_expand<XF_PATH_A>(poly, nv, T, m, verts, rays, lane, dst, total);
    if (!cells && lane == 0) {
      counts[i] = st == ST_OK ? total : 0;
      if (status) status[i] = static_cast<uint8_t>(st);
    }
    __syncwarp();
  }
}

/// Sink that stores the outline cells in order.
struct CellWriter {
  int2* __restrict__ dst;
  long long n = 0;
  __device__ __forceinline__ bool cell(int x, int y) {
    if (dst) dst[n] = make_int2(x, y);
    ++n;
    return true;
  }
};

/// expand(outline_generator(res, pose*polygon - origin)) — to_outline (expand.hpp:68-72).  One thread per pose.
__global__ void __launch_bounds__(kBlock) k_outline_expand(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src, long long count,
                                                          long long* __restrict__ counts, uint8_t* __restrict__ status,
                                                          const long long* __restrict__ starts, int2* __restrict__ cells) {
  extern __shared__ double smem_d[];
  const double* poly = stage_polygon(poly_g, nv, smem_d);
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const Pose T = load_pose(src, i);
  int x0, y0, x1, y1;
  const bool none = src.format == POSE_NONE;
  const bool ok = none ? polygon_bbox<XF_NONE>(poly, nv, T, m, x0, y0, x1, y1) : polygon_bbox<XF_PATH_A>(poly, nv, T, m, x0, y0, x1, y1);
  CellWriter w{cells ? cells + starts[i] : nullptr};
  if (ok && !(cells && starts[i + 1] == starts[i])) {
    if (none) walk_outline<XF_NONE>(poly, nv, T, m, w);
    else walk_outline<XF_PATH_A>(poly, nv, T, m, w);
  }
  if (!cells) {
    counts[i] = ok ? w.n : 0;
    if (status) status[i] = ok ? ST_OK : ST_COORD_RANGE;
  }
}


/// Trajectory scoring (the DWA-style caller right after the pose checks).  Poses are grouped as
/// n_trajectories x steps (pose index = trajectory * steps + step).  One warp per trajectory:
///   first_blocked[t] = first step whose verdict is 0 (`steps` if the whole trajectory is free)
///   score[t]         = accumulate_cost semantics lifted to the trajectory: the first negative per-pose cost
///                      in step order if there is one, else the sum of the per-pose costs (exact: integers).
__global__ void __launch_bounds__(kBlock) k_trajectories(const uint8_t* __restrict__ is_free, const double* __restrict__ cost, long long n_traj,
                                                        int steps, int* __restrict__ first_blocked, double* __restrict__ score) {
  const int lane = threadIdx.x & 31;
  const long long t = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (t >= n_traj) return;
  const long long base = t * steps;
  int blocked = steps, neg_step = steps;
  double sum = 0.0, neg = 0.0;
  for (int s0 = 0; s0 < steps; s0 += 32) {
    const int s = s0 + lane;
    const bool in = s < steps;
    if (is_free) {
      const unsigned b = __ballot_sync(0xFFFFFFFFu, in && is_free[base + s] == 0);
      if (b && blocked == steps) blocked = s0 + __ffs(b) - 1;
    }
    if (cost) {
      const double c = in ? cost[base + s] : 0.0;
      const unsigned b = __ballot_sync(0xFFFFFFFFu, in && !(c >= 0.0));  // negative sentinel or NaN (failed pose)
      if (b && neg_step == steps) {
        neg_step = s0 + __ffs(b) - 1;
        neg = __shfl_sync(0xFFFFFFFFu, c, __ffs(b) - 1);
      }
      double part = (in && c >= 0.0) ? c : 0.0;
      for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, d);
      sum += part;
    }
  }
  if (lane == 0) {
    if (first_blocked) first_blocked[t] = blocked;
    if (score) score[t] = neg_step < steps ? neg : sum;
  }
}


/// Inflation layer (the step BEFORE the path, costmap_2d::InflationLayer restated): every cell
/// within `radius` cells of a LETHAL cell gets lut[d^2] (d = Euclidean cell distance to the nearest lethal
/// cell; the LUT is built on the host so that no device exp() is involved), combined with the old cost like
/// InflationLayer::updateCosts: NO_INFORMATION stays unless the new cost is >= INSCRIBED, else max(old, new).
/// One CTA per 32 x 32 output tile; the (32 + 2R)^2 neighbourhood is staged in shared memory.
__global__ void __launch_bounds__(256) k_inflate(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int size_x, int size_y, int pitch,
                                                int radius, const uint8_t* __restrict__ lut) {
  extern __shared__ double smem_d[];
  uint8_t* tile = reinterpret_cast<uint8_t*>(smem_d);
  const int span = 32 + 2 * radius, tx0 = blockIdx.x * 32 - radius, ty0 = blockIdx.y * 32 - radius;
  __shared__ int any_lethal;
  if (threadIdx.x == 0) any_lethal = 0;
  __syncthreads();
  bool mine = false;
  for (int idx = threadIdx.x; idx < span * span; idx += blockDim.x) {
    const int ly = idx / span, lx = idx - ly * span, x = tx0 + lx, y = ty0 + ly;
    uint8_t v = 0;
    if (x >= 0 && y >= 0 && x < size_x && y < size_y) v = src[static_cast<size_t>(y) * pitch + x];
    tile[idx] = v;
    mine |= v == kLethal;
  }
  if (mine) any_lethal = 1;
  __syncthreads();
  const bool work = any_lethal != 0;
  const int r2max = radius * radius;
  for (int c = threadIdx.x; c < 32 * 32; c += blockDim.x) {
    const int cy = c >> 5, cx = c & 31, x = blockIdx.x * 32 + cx, y = blockIdx.y * 32 + cy;
    if (x >= size_x || y >= size_y) continue;
    const uint8_t old = tile[(cy + radius) * span + cx + radius];
    uint8_t out = old;
    if (work) {
      int best = 0x7FFFFFFF;
      for (int dy = -radius; dy <= radius; ++dy) {
        const uint8_t* row = tile + (cy + radius + dy) * span + cx;
        const int dy2 = dy * dy;
        for (int dx = -radius; dx <= radius; ++dx)
          if (row[dx + radius] == kLethal) best = min(best, dy2 + dx * dx);
      }
      if (best <= r2max) {
        const uint8_t cost = lut[best];
        if (old == kNoInformation) out = cost >= kInscribed ? cost : old;
        else out = max(old, cost);
      }
    }
    dst[static_cast<size_t>(y) * pitch + x] = out;
  }
}


/// area checks over SEVERAL outlines (holes): batched is_free / accumulate_cost / max over
/// area_generator(resolution, polygon_vec) (generators.cpp:126-135).  One warp per pose.
/// Inflation layer, radius <= 31 cells: exact Euclidean distance to the nearest LETHAL cell, separably.  One CTA per
/// 32 x 32 tile: (1) every row the tile's cells can reach becomes three ballot words of "cell is LETHAL" (the tile's 32
/// columns and 32 on either side); (2) per (row, column) the horizontal distance g to the nearest lethal cell of that row
/// from two funnel-shifted words (count-leading-zeros to the left, find-first-set to the right); (3) per cell
/// d^2 = min over dy of g(y + dy)^2 + dy^2, walking outwards from dy = 0 and stopping once dy^2 reaches the best so far.
/// IN PLACE: a cell's new cost never makes or unmakes a LETHAL cell (the caller checks cost_by_d2[k > 0] != 254), so tiles
/// may read neighbours that other tiles have already rewritten; tiles without a lethal cell in reach write nothing.
constexpr int kInflateFastRadius = 31;  // (three 32-cell ballot words per row: the horizontal search reaches 31 cells either way)
__global__ void __launch_bounds__(256) k_inflate_separable(uint8_t* __restrict__ map, int size_x, int size_y, int pitch, int radius,
                                                            const uint8_t* __restrict__ lut) {
  __shared__ uint32_t hbits[32 + 2 * kInflateFastRadius][3];
  __shared__ uint8_t gbuf[32 + 2 * kInflateFastRadius][32];
  __shared__ int any_lethal;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * 32 - radius, rows = 32 + 2 * radius;
  if (threadIdx.x == 0) any_lethal = 0;
  __syncthreads();
  bool seen = false;
  for (int r = warp; r < rows; r += 8) {
    const int y = ty0 + r;
    const bool y_in = y >= 0 && y < size_y;
    const uint8_t* __restrict__ rowp = map + static_cast<size_t>(y_in ? y : 0) * pitch;
    uint32_t w[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int x = tx0 - 32 + 32 * j + lane;
      const bool lethal = y_in && x >= 0 && x < size_x && rowp[x] == kLethal;
      w[j] = __ballot_sync(0xFFFFFFFFu, lethal);
    }
    if (lane < 3) hbits[r][lane] = lane == 0 ? w[0] : (lane == 1 ? w[1] : w[2]);
    seen |= (w[0] | w[1] | w[2]) != 0u;
  }
  if (seen && lane == 0) any_lethal = 1;
  __syncthreads();
  if (!any_lethal) return;
  for (int idx = threadIdx.x; idx < rows * 32; idx += blockDim.x) {
    const int r = idx >> 5, cx = idx & 31;
    const uint32_t w0 = hbits[r][0], w1 = hbits[r][1], w2 = hbits[r][2];
    const uint32_t right = __funnelshift_r(w1, w2, cx);          // bit k: the cell k columns to the right (k = 0: the cell itself)
    const uint32_t left = __funnelshift_rc(w0, w1, cx + 1);      // bit 31 - k: the cell k columns to the left
    const int g = min(right ? __ffs(right) - 1 : 64, left ? __clz(left) : 64);
    gbuf[r][cx] = static_cast<uint8_t>(g <= radius ? g : 255);
  }
  __syncthreads();
  const int r2max = radius * radius;
  for (int c = threadIdx.x; c < 32 * 32; c += blockDim.x) {
    const int cy = c >> 5, cx = c & 31, x = tx0 + cx, y = blockIdx.y * 32 + cy;
    if (x >= size_x || y >= size_y) continue;
    const int rc = cy + radius;
    int best = gbuf[rc][cx] == 255 ? 0x3FFFFFFF : gbuf[rc][cx] * gbuf[rc][cx];
    for (int dy = 1; dy <= radius && dy * dy < best; ++dy) {
      const int a = gbuf[rc - dy][cx], b = gbuf[rc + dy][cx], m = min(a, b);
      if (m != 255) best = min(best, m * m + dy * dy);
    }
    if (best > r2max) continue;
    uint8_t* cell = map + static_cast<size_t>(y) * pitch + x;
    const uint8_t old = *cell, cost = lut[best];
    uint8_t out;
    if (old == kNoInformation) out = cost >= kInscribed ? cost : old;
    else out = max(old, cost);
    if (out != old) *cell = out;
  }
}

template <int OP>
__global__ void __launch_bounds__(kRowsWarps * 32) k_area_rows_multi(MapView m, const double* __restrict__ poly_g, const int* __restrict__ starts_g,
                                                                    int n_outlines, int nv_total, int nv_max, PoseSource src, long long count,
                                                                    OutView out) {
  extern __shared__ double smem_d[];
  const double* poly = stage_polygon(poly_g, nv_total, smem_d);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rays_per_warp = nv_total + n_outlines;
  RayRec* rays = reinterpret_cast<RayRec*>(smem_d + 2 * nv_total) + warp * rays_per_warp;
  int2* verts = reinterpret_cast<int2*>(reinterpret_cast<RayRec*>(smem_d + 2 * nv_total) + kRowsWarps * rays_per_warp) + warp * nv_max;
  OutlineRec* recs = reinterpret_cast<OutlineRec*>(reinterpret_cast<int2*>(reinterpret_cast<RayRec*>(smem_d + 2 * nv_total) + kRowsWarps * rays_per_warp) +
                                                   kRowsWarps * nv_max) + warp * kRowsMaxOutlines;
  __shared__ int starts[kRowsMaxOutlines + 1];
  if (threadIdx.x <= n_outlines) starts[threadIdx.x] = starts_g[threadIdx.x];
  __syncthreads();
  const long long n_warps = static_cast<long long>(gridDim.x) * kRowsWarps;
  const bool none = src.format == POSE_NONE;
  for (long long i = static_cast<long long>(blockIdx.x) * kRowsWarps + warp; i < count; i += n_warps) {
    const Pose T = load_pose(src, i);
    WarpFold<OP> wf;
    const int status = none ? area_rows_multi_pose<OP, XF_NONE>(poly, starts, n_outlines, T, m, verts, rays, recs, lane, wf)
                            : area_rows_multi_pose<OP, XF_PATH_A>(poly, starts, n_outlines, T, m, verts, rays, recs, lane, wf);
    if (lane == 0) {
      if (status != ST_OK) {
        store_failure(out, i, status);
      } else {
        wf.store(out, i);
        if (out.status) out.status[i] = ST_OK;
      }
    }
    __syncwarp();
  }
}


/// Streaming 16-byte-vector read of a buffer (bandwidth probe: 16 MB stays in L2, 1 GB streams from HBM).
__global__ void __launch_bounds__(256) k_probe_read(const uint4* __restrict__ buf, long long n_vec, int reps, unsigned int* __restrict__ sink) {
  unsigned int acc = 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
      const uint4 v = __ldcg(buf + i);  // cache at L2 only: every pass really crosses the L2 -> SM path
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x9E3779B9u) *sink = acc;  // keeps the loads alive
}

template <int OP>
__global__ void __launch_bounds__(kBlock) k_rays(MapView m, const int4* __restrict__ segments, long long count, int offx, int offy,
                                                OutView out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int4 s = __ldg(segments + i);
  CellSink<OP> sink(m, offx, offy);
  walk_ray(s.x, s.y, s.z, s.w, sink);
  sink.fold.store(out, i);
  if (out.status) out.status[i] = ST_OK;
}

template <int OP>
__global__ void __launch_bounds__(kBlock) k_cells(MapView m, const int2* __restrict__ cells, long long n_cells,
                                                 const int2* __restrict__ offsets, long long count, OutView out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int2 off = __ldg(offsets + i);
  Fold<OP> fold;
  bool more = true;
  for (long long j0 = 0; j0 < n_cells && more; j0 += 8) {  // 8 independent loads, consumed in list order
    int v[8];
    bool in[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int2 c = j0 + j < n_cells ? __ldg(cells + j0 + j) : make_int2(0, 0);
      const int x = c.x + off.x, y = c.y + off.y;
      in[j] = static_cast<unsigned>(x) < static_cast<unsigned>(m.size_x) && static_cast<unsigned>(y) < static_cast<unsigned>(m.size_y);
      v[j] = (j0 + j < n_cells && in[j]) ? __ldg(m.data + static_cast<size_t>(y) * m.pitch + x) : 0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (more && j0 + j < n_cells) more = fold.consume(in[j], v[j]);
  }
  fold.store(out, i);
  if (out.status) out.status[i] = ST_OK;
}

/// is_free(discrete_polygon, offset, map) over the host-precomputed row spans of the cell SET (an
/// any-reduction does not care about order or duplicates).
__global__ void __launch_bounds__(kBlock) k_cells_free(MapView m, const int4* __restrict__ spans, int n_spans, const int2* __restrict__ offsets,
                                                      long long count, OutView out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int2 off = __ldg(offsets + i);
  bool hit = false;
  for (int j = 0; j < n_spans && !hit; ++j) {
    const int4 sp = __ldg(spans + j);
    const int y = sp.x + off.y, lo = sp.y + off.x, hi = sp.z + off.x;
    hit = static_cast<unsigned>(y) >= static_cast<unsigned>(m.size_y) || lo < 0 || hi >= m.size_x ||
          (m.planes ? span_any_bits(m.planes, m.ypitch, m.plane_words, y, lo, hi)
                    : row_any_eq(m.data + static_cast<size_t>(y) * m.pitch, lo, hi, 0xFEFEFEFEu));
  }
  emit_free(out, i, !hit);
  if (out.status) out.status[i] = ST_OK;
}

/// discrete_footprint::is_free(offset, map): outline cells collide on 253/254/outside, area cells on
/// 254/outside (footprints.cpp:240-247).  The verdict is an any-reduction, so the order of the cells is
/// irrelevant: at upload the host turned each footprint's cell SET into row spans {y, x_lo, x_hi}
/// ("the polygon split precomputed once on the host per footprint"), which are scanned 4 cells per word.
__global__ void __launch_bounds__(kBlock) k_footprints(MapView m, const int4* __restrict__ spans, const int2* __restrict__ ranges,
                                                      int n_footprints, const int2* __restrict__ offsets, const int* __restrict__ which,
                                                      long long count, OutView out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int2 off = __ldg(offsets + i);
  const int k = which ? __ldg(which + i) : 0;
  if (static_cast<unsigned>(k) >= static_cast<unsigned>(n_footprints)) {  // (device-resident `which` cannot be validated on the host)
    store_failure(out, i, ST_COORD_RANGE);
    return;
  }
  const int2 o_rng = __ldg(ranges + 2 * k), a_rng = __ldg(ranges + 2 * k + 1);  // [begin, end) of outline / area spans
  bool hit = false;
  for (int j = o_rng.x; j < o_rng.y && !hit; ++j) {
    const int4 sp = __ldg(spans + j);
    const int y = sp.x + off.y, lo = sp.y + off.x, hi = sp.z + off.x;
    hit = static_cast<unsigned>(y) >= static_cast<unsigned>(m.size_y) || lo < 0 || hi >= m.size_x ||
          (m.planes_inf ? span_any_bits(m.planes_inf, m.ypitch, m.plane_words, y, lo, hi)
                        : row_any_inflated(m.data + static_cast<size_t>(y) * m.pitch, lo, hi));
  }
  for (int j = a_rng.x; j < a_rng.y && !hit; ++j) {
    const int4 sp = __ldg(spans + j);
    const int y = sp.x + off.y, lo = sp.y + off.x, hi = sp.z + off.x;
    hit = static_cast<unsigned>(y) >= static_cast<unsigned>(m.size_y) || lo < 0 || hi >= m.size_x ||
          (m.planes ? span_any_bits(m.planes, m.ypitch, m.plane_words, y, lo, hi)
                    : row_any_eq(m.data + static_cast<size_t>(y) * m.pitch, lo, hi, 0xFEFEFEFEu));
  }
  emit_free(out, i, !hit);
  if (out.status) out.status[i] = ST_OK;
}

/// Sink of continuous_footprint's outline loop: raw getCost, collide on 253 or 254 (footprints.cpp:166-173).
struct InflatedSink {
  const MapView& m;
  bool hit = false;
  __device__ __forceinline__ explicit InflatedSink(const MapView& mv) : m(mv) {}
  __device__ __forceinline__ bool cell(int x, int y) {
    const int v = __ldg(m.data + static_cast<size_t>(y) * m.pitch + x);
    hit = v == kInscribed || v == kLethal;
    return !hit;
  }
};

struct SplitView {
  const double* xy;     // all vertices, outlines first, then areas (column-major)
  const int* starts;    // n_outlines + n_areas + 1 vertex offsets
  int n_outlines, n_areas;
};

/// continuous_footprint::is_free(pose, map) — footprints.cpp:130-196.
/// Sink that feeds outline cells into a SlotWalker (per-thread walk, no lock-step).
struct SlotSink {
  SlotWalker<4>& w;
  __device__ __forceinline__ bool cell(int x, int y) {
    w.visit(x, y);
    return true;
  }
};

__global__ void __launch_bounds__(kBlock) k_split(MapView m, SplitView fp, int n_vertices_total, PoseSource src, long long count,
                                                 OutView out, ListScratch scratch, int slot_rows, const unsigned int* __restrict__ in_list,
                                                 const unsigned int* __restrict__ in_count) {
  extern __shared__ double smem_d[];
  const double* xy = stage_polygon(fp.xy, n_vertices_total, smem_d);
  uint32_t* slot_table = reinterpret_cast<uint32_t*>(smem_d + 2 * n_vertices_total);  // [slot_rows][4][blockDim.x], 0 rows = disabled
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long me = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  bool listed = false;  // (with a list: only the poses k_split_lockstep deferred - or all of them, if the list overflowed)
  const long long n_items = in_count ? todo_items(in_count, count, listed) : count;
  for (long long item = me; item < n_items; item += nthreads) {
    const long long i = listed ? static_cast<long long>(in_list[item]) : item;
    Pose T = load_pose(src, i);
    T.tx = __dadd_rn(T.tx, -m.origin_x);  // Translation2d(-origin) * pose: translation += (-origin)
    T.ty = __dadd_rn(T.ty, -m.origin_y);
    const int n_polys = fp.n_outlines + fp.n_areas;
    bool in_range = true;
    for (int k = 0; k < n_polys; ++k) {
      int a, b, c, d;
      in_range &= polygon_bbox<XF_PATH_B>(xy + 2 * fp.starts[k], fp.starts[k + 1] - fp.starts[k], T, m, a, b, c, d);
    }
    if (!in_range) {
      store_failure(out, i, ST_COORD_RANGE);
      continue;
    }
    bool free_ = true;
    int status = ST_OK;
    for (int k = 0; k < n_polys && free_ && status == ST_OK; ++k) {
      const double* poly = xy + 2 * fp.starts[k];
      const int nv = fp.starts[k + 1] - fp.starts[k];
      if (nv == 0) continue;
      int x0, y0, x1, y1;
      polygon_bbox<XF_PATH_B>(poly, nv, T, m, x0, y0, x1, y1);
      if (x0 < 0 || y0 < 0 || x1 >= m.size_x || y1 >= m.size_y) {  // footprints.cpp:139-155
        free_ = false;
        break;
      }
      if (k < fp.n_outlines) {
        InflatedSink sink(m);
        walk_outline<XF_PATH_B>(poly, nv, T, m, sink);
        free_ = !sink.hit;
      } else {
        const int rows = y1 - y0 + 1;
        if (rows > scratch.rows_cap) {
          status = ST_COORD_RANGE;
          break;
        }
        if (rows <= slot_rows && (x1 - x0) <= 65000) {
          // fast attempt: four ranges per row in shared memory (most remainder areas need no more); on overflow the
          // sub-polygon is redone below with the sorted lists in global scratch
          SlotWalker<4> w;
          w.tab = slot_table + threadIdx.x, w.stride = static_cast<int>(blockDim.x), w.xb = x0, w.yb = y0;
          for (int r = 0; r < rows; ++r)
            for (int q = 0; q < 4; ++q) w.slot(r, q) = kPairEmpty;
          int fx, fy;
          to_cell<XF_PATH_B>(T, poly[0], poly[1], m, fx, fy);
          w.y_run = w.y0 = fy, w.x_first = w.x0 = w.x_prev = fx, w.y_before = fy;
          w.in_first = true, w.overflow = false, w.f_xlast = fx, w.f_ynext = fy;
          SlotSink sink{w};
          walk_outline<XF_PATH_B>(poly, nv, T, m, sink);
          if (w.in_first) {  // every cell on one row
            const uint32_t wd = static_cast<uint32_t>(x1 - x0);
            w.slot(w.y0 - y0, 0) = 0u, w.slot(w.y0 - y0, 1) = wd | (wd << 16);
          } else if (w.y_run == w.y0) {
            w.prepend(w.y0, w.x_first, w.f_xlast, w.y_before, w.f_ynext);
          } else {
            w.append(w.y_run, w.x_first, w.x_prev, w.y_before, w.y0);
            w.prepend(w.y0, w.x0, w.f_xlast, w.y_run, w.f_ynext);
          }
          if (!w.overflow) {
            bool odd = false;
            for (int r = 0; r < rows; ++r) odd |= (w.count(r) & 1) != 0;
            if (odd) {
              status = ST_LOGIC_ERROR;
              break;
            }
            Fold<OP_IS_FREE> fold;
            bool more = true;
            for (int r = 0; r < rows && more; ++r) {
              uint32_t e[4];
              for (int q = 0; q < 4; ++q) e[q] = w.slot(r, q);
              if (e[2] != kPairEmpty) {
#pragma unroll
                for (int a = 1; a < 4; ++a)
#pragma unroll
                  for (int b = a; b > 0; --b)
                    if ((e[b] & 0xFFFFu) < (e[b - 1] & 0xFFFFu)) {
                      const uint32_t t = e[b];
                      e[b] = e[b - 1], e[b - 1] = t;
                    }
                more = fold.span(m, y0 + r, x0 + static_cast<int>(e[0] & 0xFFFFu), x0 + static_cast<int>(e[1] >> 16));
                if (more) more = fold.span(m, y0 + r, x0 + static_cast<int>(e[2] & 0xFFFFu), x0 + static_cast<int>(e[3] >> 16));
              } else if (e[0] != kPairEmpty) {
                const int lo1 = e[0] & 0xFFFF, hi1 = e[0] >> 16, lo2 = e[1] & 0xFFFF, hi2 = e[1] >> 16;
                more = fold.span(m, y0 + r, x0 + min(lo1, lo2), x0 + (lo2 < lo1 ? hi1 : hi2));
              }
            }
            free_ = !fold.hit;
            continue;
          }
        }
        ListStore store{scratch.cnt, scratch.rng, scratch.seq, nthreads, me, y0, scratch.rows_cap};
        store.clear(rows);
        RunDetector<ListStore> runs(store, 0, x0, x1);
        walk_outline<XF_PATH_B>(poly, nv, T, m, runs);
        runs.finish();
        Fold<OP_IS_FREE> fold;
        status = fold_list_rows(store, rows, m, fold);
        free_ = !fold.hit;
      }
    }
    if (status != ST_OK) {
      store_failure(out, i, status);
    } else {
      emit_free(out, i, free_);
      if (out.status) out.status[i] = ST_OK;
    }
  }
}

/// continuous_footprint::is_free(pose, map) with the warp's 32 walks fused in lock-step: every pose of a batch has the
/// same footprint, so all lanes are at the same sub-polygon and the same ray - outline cells are tested as they are
/// generated (INSCRIBED or LETHAL, footprints.cpp:157-170), areas go through the four-ranges-per-row table of
/// k_area_quad (area_quad_core).  Same order of decisions as k_split: coordinates of ALL sub-polygons in range, then per
/// sub-polygon the vertex bounding box inside the map (footprints.cpp:139-155), then its cells; the first sub-polygon
/// that is not free (or fails) ends the pose.  Poses with an area beyond the table are left to k_split (todo list).
__global__ void __launch_bounds__(kBlock) k_split_lockstep(MapView m, SplitView fp, int n_vertices_total, PoseSource src, long long count,
                                                          OutView out, int rows_cap, unsigned int* __restrict__ todo,
                                                          unsigned int* __restrict__ todo_count) {
  extern __shared__ double smem_d[];
  const double* xy = stage_polygon(fp.xy, n_vertices_total, smem_d);
  uint2* table = reinterpret_cast<uint2*>(smem_d + 2 * n_vertices_total);
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const int n_polys = fp.n_outlines + fp.n_areas;
  for (long long base = static_cast<long long>(blockIdx.x) * blockDim.x; base < count; base += nthreads) {  // whole warps iterate together
    const long long i = base + threadIdx.x;
    const bool valid = i < count;
    Pose T = load_pose(src, valid ? i : 0);
    T.tx = __dadd_rn(T.tx, -m.origin_x);  // Translation2d(-origin) * pose: translation += (-origin)
    T.ty = __dadd_rn(T.ty, -m.origin_y);
    bool in_range = true;
    for (int k = 0; k < n_polys; ++k) {
      int a, b, c, d;
      in_range &= polygon_bbox<XF_PATH_B>(xy + 2 * fp.starts[k], fp.starts[k + 1] - fp.starts[k], T, m, a, b, c, d);
    }
    if (valid && !in_range) store_failure(out, i, ST_COORD_RANGE);
    bool alive = valid && in_range;  // free and OK so far
    bool free_ = true, deferred = false;
    int status = ST_OK;
    for (int k = 0; k < n_polys; ++k) {
      if (!__any_sync(kFullMask, alive)) break;
      const double* poly = xy + 2 * fp.starts[k];
      const int nv = fp.starts[k + 1] - fp.starts[k];
      if (nv == 0) continue;
      int x0, y0, x1, y1;
      polygon_bbox<XF_PATH_B>(poly, nv, T, m, x0, y0, x1, y1);
      if (alive && (x0 < 0 || y0 < 0 || x1 >= m.size_x || y1 >= m.size_y)) free_ = false, alive = false;  // footprints.cpp:139-155
      if (k < fp.n_outlines) {
        int fx, fy;
        to_cell<XF_PATH_B>(T, poly[0], poly[1], m, fx, fy);
        int px = fx, py = fy, rays = 0;
        bool hit = false;
        LaneRay ray;
        for (int v = 1; v <= nv; ++v) {  // v == nv: the closing ray, or the one-cell dummy ray of an open outline
          int cx, cy;
          if (v < nv) {
            to_cell<XF_PATH_B>(T, poly[2 * v], poly[2 * v + 1], m, cx, cy);
          } else {
            cx = fx, cy = fy;
          }
          ray.set(px, py, cx, cy);
          if (v < nv) {
            rays += ray.size != 0;  // degenerate rays are skipped (generators.cpp:86-93)
          } else if (rays <= 1) {
            ray.size = 1, ray.add = 0, ray.num = 0;  // dummy ray: the last point itself (generators.cpp:72-78)
            ray.major_x = ray.major_y = ray.minor_x = ray.minor_y = 0;
          }
          if (!alive || hit) ray.size = 0;
          const int n_iter = __reduce_max_sync(kFullMask, ray.size);
          for (int c = 0; c < n_iter; ++c) {
            if (c < ray.size) {
              const int v8 = __ldg(m.data + static_cast<size_t>(ray.y) * m.pitch + ray.x);
              if (v8 == kInscribed || v8 == kLethal) hit = true, ray.size = 0;
              ray.step();
            }
          }
          px = cx, py = cy;
        }
        if (alive && hit) free_ = false, alive = false;
      } else {
        Fold<OP_IS_FREE> fold;
        const int rc = area_quad_core<XF_PATH_B, OP_IS_FREE>(poly, nv, T, m, table, rows_cap, alive, x0, y0, x1, y1, fold);
        if (rc == 1) deferred = true, alive = false;
        else if (rc == 2) status = ST_LOGIC_ERROR, alive = false;
        else if (rc == 0 && fold.hit) free_ = false, alive = false;
      }
    }
    if (valid && in_range) {
      if (deferred) {
        push_todo(todo, todo_count, i);
      } else if (status != ST_OK) {
        store_failure(out, i, status);
      } else {
        emit_free(out, i, free_);
        if (out.status) out.status[i] = ST_OK;
      }
    }
  }
}

// ===================================================================================================
// Host side
// ===================================================================================================

struct DeviceBuffer {
  void* ptr = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    const cudaError_t e = cudaMalloc(&ptr, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(ptr);
  }
};

}  // namespace

struct cover_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // banded costmap uploads (cover_map_upload_async) overlap the kernels of `stream`
  cudaEvent_t fence = nullptr;         // "everything queued on `stream` so far": what a banded upload has to wait for
  cudaStream_t d2h_stream = nullptr;   // result chunks of a big lattice sweep travel back while the next chunk is computed
  cudaStream_t aux_stream = nullptr;   // k_lw_prepare (needs no map data) runs beside the plane build / the costmap upload
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr, aux_planes = nullptr, aux_done = nullptr, aux_zero = nullptr;
  bool zero_on_aux = false;             // this call's packed verdicts are being zeroed on the auxiliary stream (aux_zero marks the end)
  cudaEvent_t chunk_done = nullptr;
  std::string error;
  int64_t launches = 0;
  unsigned int todo_header[2] = {0u, 0u}, todo2_header[2] = {0u, 0u};  // {pushed = 0, capacity} staged for the deferred-pose lists
  bool disable_lattice_kernel = false;
  bool disable_words = false;           // COVER_B200_NO_WORDS=1: lattice verdicts through the shape-cache kernel (A/B measurements, its tests)
  // class / shape tables of the last bit-parallel lattice sweep (lw_tables): they depend on the polygon, the lattice and the
  // map's GEOMETRY only - not on its cells - and are reused while those stay the same (COVER_B200_NO_TABLE_CACHE=1: never)
  std::string theta_host;                // bytes of the heading table currently in `theta` (re-uploaded only when they change)
  const void* theta_dev = nullptr;
  std::string lw_key_base, lw_key;
  cvr::LwTables lw_cached_t{};
  cvr::LwGeom lw_cached_g{};
  bool lw_cache_valid = false, lw_cache_disabled = false;
  int64_t todo_cap_limit = 0;           // COVER_B200_TODO_CAP: caps every deferred-pose list (tests of the overflow path)
  int words_strip = 4;                  // COVER_B200_WORDS_STRIP: 32-pose words per thread of k_lw_sweep (2, 4 or 8)
  bool disable_planes = false;          // COVER_B200_NO_PLANES=1: lattice verdicts walk the cost bytes (A/B measurements)
  int64_t planes_min_batch = 32768;     // COVER_B200_PLANES_MIN_BATCH: smallest batch that triggers a bit-plane build (tests: 0)
  int max_stretch = 4;                  // COVER_B200_MAX_STRETCH: longest run of consecutive rounds a lattice CTA draws at once
  int64_t pipeline_min = 0;             // COVER_B200_PIPELINE_MIN: smallest lattice sweep that is downloaded in chunks (tests; default 2^28 poses)
  int pipeline_chunks = 0;              // COVER_B200_CHUNKS: download chunks of a big lattice sweep (0 = by size)
  bool old_packed_kernel = false;       // COVER_B200_OLD_PACKED=1: first-generation per-thread kernel (A/B measurements)
  bool force_list_kernel = false;
  CallReset pending_reset{nullptr, 0u, nullptr};  // handed to the next k_lethal_planes launch of this call (ensure_plane_set)
  bool lazy_out = false;                // this call's launch_polygon zeroes the packed verdict words / the not-OK counter itself (or has no need to)
  void* lazy_not_ok = nullptr;
  int plane_cols = 0;                   // COVER_B200_PLANE_COLS: word columns per warp of k_lethal_planes (2, 4, 8; 0 = by size)
  bool timeline = false;                // COVER_B200_TIMELINE=1: events between the operations of a call, printed at cover_ctx_synchronize (bring-up)
  std::vector<std::pair<const char*, cudaEvent_t>> timeline_events;
  int quad_max_rows = kQuadMaxRows;     // COVER_B200_QUAD_MAX_ROWS: tallest polygon (scan rows) k_area_quad takes (measurements)
  bool disable_split_lockstep = false;  // COVER_B200_NO_SPLIT_LOCKSTEP=1: continuous_footprint checks through k_split alone (A/B measurements)
  bool disable_values = false;          // COVER_B200_NO_VALUES=1: lattice accumulate_cost / max_cost through the shape-cache kernel (A/B measurements)
  bool disable_quad = false;            // COVER_B200_NO_QUAD=1: non-convex polygons through k_area_pair + k_area_slots (A/B measurements)
  bool always_try_packed = true;        // COVER_B200_PACKED_ONLY_CONVEX=1 restricts the packed pass to convex polygons       // COVER_B200_FORCE_LIST_KERNEL=1: A/B measurements and tests of k_area_list  // COVER_B200_NO_LATTICE_KERNEL=1: A/B measurements only
  DeviceBuffer poses, theta, poly, in_a, in_b, out_free, out_cost, out_status, out_bits, out_not_ok, todo, todo_count, todo2, todo2_count, list_cnt, list_rng, list_seq, lat_global, lw_tables;
};

constexpr int kMapBands = 32;

struct cover_map {
  cover_ctx* ctx;
  MapView view;
  uint8_t* dev;
  double resolution;
  // cover_map_upload_async: the image travels as kMapBands row bands on ctx->copy_stream, one event per band, the
  // bands the last lattice sweep read first.  While `bands_pending`, every reader of the map is first ordered (on
  // the device) after the bands it can touch: a lattice sweep after the bands under its pose window, anything else
  // after the whole image.
  cudaEvent_t band_done[kMapBands] = {};
  int band_rows = 0;
  int issue_pos[kMapBands] = {};  // copy group of each band in the upload in flight = index of its event in band_done
  int last_band = kMapBands - 1;  // the band copied last
  mutable bool bands_pending = false;
  mutable int hint_lo = 0, hint_hi = -1;  // bands the last lattice sweep needed
  // lethal bit planes for lattice verdict sweeps (PlaneView): built lazily for the map rows a sweep can touch,
  // invalidated by everything that changes the map
  mutable uint32_t* planes = nullptr;      // LETHAL planes
  mutable uint32_t* planes_inf = nullptr;  // INSCRIBED-or-LETHAL planes (outline cells of split footprints)
  mutable int planes_lo = 0, planes_hi = -1;  // rows whose planes match the current image
  mutable int planes_inf_lo = 0, planes_inf_hi = -1;
  mutable uint32_t* prefix = nullptr;  // per-row prefix sums of the costs (accumulate_cost lattice sweeps)
  mutable int prefix_lo = 0, prefix_hi = -1;
  mutable uint8_t* maxp = nullptr;  // sliding-maximum planes 0..5 of the costs (max_cost lattice sweeps, k_max_planes)
  mutable int maxp_lo = 0, maxp_hi = -1;
  int ppitch = 0;
  int ypitch = 0;
  long long plane_words = 0;
};

struct cover_cells {
  cover_ctx* ctx;
  int2* dev;     // the cells in list order (accumulate_cost / max_cost)
  int64_t n;
  int4* spans;   // {y, x_lo, x_hi, 0} row spans of the cell set (is_free)
  int n_spans;
};

struct cover_footprints {
  cover_ctx* ctx;
  int4* spans;   // {y, x_lo, x_hi, 0} row spans of every footprint's outline and area cell sets
  int2* ranges;  // per footprint: [2k] = outline span range, [2k+1] = area span range
  int32_t n;
};

struct cover_split {
  cover_ctx* ctx;
  double* xy;
  int* starts;
  int32_t n_outlines, n_areas, n_vertices;
  double max_area_diameter;
};

namespace {

/// The auxiliary stream runs short kernels BESIDE the main stream's big one (the general half of a lattice sweep beside
/// the simple half): at the highest priority its blocks are placed first, instead of waiting for the big grid to drain.
cudaError_t create_priority_stream(cudaStream_t* s) {
  int least = 0, greatest = 0;
  if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) {
    cudaGetLastError();
    return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
  }
  return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, greatest);
}

/// COVER_B200_TIMELINE=1: an event on the main stream, labelled (printed at the next cover_ctx_synchronize)
void tl(cover_ctx* ctx, const char* label) {
  if (!ctx->timeline) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, ctx->stream);
  ctx->timeline_events.emplace_back(label, e);
}

int fail(cover_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->error = msg;
  return code;
}

int cuda_fail(cover_ctx* ctx, cudaError_t e, const char* what) {
  return fail(ctx, e == cudaErrorMemoryAllocation ? COVER_E_NOMEM : COVER_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(call)                                                 \
  do {                                                           \
    const cudaError_t e_ = (call);                               \
    if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);     \
  } while (0)

struct Bind {  // makes `device` current for the duration of a call
  int prev = -1;
  explicit Bind(int device) {
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    else prev = -1;
  }
  ~Bind() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

/// Results staged on the device when the caller's arrays are host memory.
struct OutStage {
  OutView dev{nullptr, nullptr, nullptr, nullptr, nullptr};
  const cover_results* user = nullptr;
  bool host = false;
  bool async = false;  // COVER_MEM_HOST_ASYNC: the downloads are queued, the call does not wait for them
};

int stage_out(cover_ctx* ctx, const cover_results* out, int64_t count, OutStage& st, bool lazy = false) {
  st.user = out;
  ctx->lazy_out = lazy, ctx->lazy_not_ok = nullptr;
  if (!out) return fail(ctx, COVER_E_INVALID_ARG, "results must not be NULL");
  st.host = out->mem == COVER_MEM_HOST || out->mem == COVER_MEM_HOST_ASYNC;
  st.async = out->mem == COVER_MEM_HOST_ASYNC;
  const size_t n = static_cast<size_t>(std::max<int64_t>(count, 1));
  const size_t bit_bytes = ((n + 31) / 32) * sizeof(uint32_t);
  if (!st.host) {
    st.dev = OutView{out->is_free, out->cost, out->status, out->free_bits, reinterpret_cast<unsigned long long*>(out->not_ok)};
    if (lazy) {
      ctx->lazy_not_ok = out->not_ok;
      return COVER_OK;
    }
    if (out->free_bits) CU(cudaMemsetAsync(out->free_bits, 0, bit_bytes, ctx->stream));
    if (out->not_ok) CU(cudaMemsetAsync(out->not_ok, 0, sizeof(int64_t), ctx->stream));
    return COVER_OK;
  }
  if (out->is_free) {
    CU(ctx->out_free.ensure(n));
    st.dev.is_free = ctx->out_free.as<uint8_t>();
  }
  if (out->cost) {
    CU(ctx->out_cost.ensure(n * sizeof(double)));
    st.dev.cost = ctx->out_cost.as<double>();
  }
  if (out->status) {
    CU(ctx->out_status.ensure(n));
    st.dev.status = ctx->out_status.as<uint8_t>();
  }
  if (out->free_bits) {
    CU(ctx->out_bits.ensure(bit_bytes));
    if (!lazy) CU(cudaMemsetAsync(ctx->out_bits.ptr, 0, bit_bytes, ctx->stream));
    st.dev.free_bits = ctx->out_bits.as<uint32_t>();
  }
  if (out->not_ok) {
    CU(ctx->out_not_ok.ensure(sizeof(unsigned long long)));
    if (lazy) ctx->lazy_not_ok = ctx->out_not_ok.ptr;
    else CU(cudaMemsetAsync(ctx->out_not_ok.ptr, 0, sizeof(unsigned long long), ctx->stream));
    st.dev.not_ok = ctx->out_not_ok.as<unsigned long long>();
  }
  return COVER_OK;
}

__global__ void __launch_bounds__(256) k_copy16(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16, uint8_t* __restrict__ dst8,
                                                const uint8_t* __restrict__ src8, size_t bytes) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {  // four independent 16-byte loads in flight per thread
    const uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
    dst[i] = a, dst[i + stride] = b, dst[i + 2 * stride] = c, dst[i + 3 * stride] = d;
  }
  for (; i < n16; i += stride) dst[i] = __ldcs(src + i);
  if (blockIdx.x == 0 && threadIdx.x < (bytes & 15u)) dst8[(bytes & ~static_cast<size_t>(15)) + threadIdx.x] = src8[(bytes & ~static_cast<size_t>(15)) + threadIdx.x];
}

/// `rows` rows of `width` bytes, source unpadded, destination pitched: ONE linear copy when the pitch equals the width
/// (a pitched copy of 4000-byte rows runs at about half the PCIe rate of a linear one on the B200 boxes measured).
cudaError_t copy_rows(uint8_t* dst, size_t dpitch, const uint8_t* src, size_t width, size_t rows, cudaMemcpyKind kind, cudaStream_t stream,
                      int64_t* launches = nullptr) {
  if (dpitch == width && kind == cudaMemcpyDeviceToDevice && width * rows >= (1u << 20) &&
      ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0) {
    // a costmap that is already on the device: 16-byte loads and stores from every SM (a 16 MB image in ~7 us; the
    // driver's device-to-device copy took 14 us on the B200 boxes measured)
    const size_t bytes = width * rows;
    k_copy16<<<148 * 8, 256, 0, stream>>>(reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), bytes / 16, dst, src, bytes);
    if (launches) ++*launches;
    return cudaGetLastError();
  }
  if (dpitch == width) return cudaMemcpyAsync(dst, src, width * rows, kind, stream);
  return cudaMemcpy2DAsync(dst, dpitch, src, width, width, rows, kind, stream);
}

/// Orders `ctx->stream` after an upload that is still in flight on the copy stream (all bands).
cudaError_t map_ready(cover_ctx* ctx, const cover_map* map) {
  if (!map->bands_pending) return cudaSuccess;
  map->bands_pending = false;
  return cudaStreamWaitEvent(ctx->stream, map->band_done[map->issue_pos[map->last_band]], 0);
}

/// Same for a reader that touches map rows [row_lo, row_hi] only (rows outside the map are never read).
cudaError_t map_rows_ready(cover_ctx* ctx, const cover_map* map, long long row_lo, long long row_hi) {
  if (!map->bands_pending) return cudaSuccess;
  row_lo = std::max<long long>(row_lo, 0), row_hi = std::min<long long>(row_hi, map->view.size_y - 1);
  if (row_lo > row_hi) return cudaSuccess;
  const int b_lo = static_cast<int>(row_lo / map->band_rows), b_hi = static_cast<int>(row_hi / map->band_rows);
  int latest = b_lo;  // the copy stream is in order: the band issued last among the needed ones covers them all
  for (int b = b_lo; b <= b_hi; ++b)
    if (map->issue_pos[b] > map->issue_pos[latest]) latest = b;
  map->hint_lo = b_lo, map->hint_hi = b_hi;
  if (map->issue_pos[latest] == map->issue_pos[map->last_band]) map->bands_pending = false;
  return cudaStreamWaitEvent(ctx->stream, map->band_done[map->issue_pos[latest]], 0);
}

/// Map rows a polygon check over a pose lattice can read: the lattice's y-extent +- the footprint's reach,
/// clamped to the map (lo > hi: none).  False when the extent is not finite (the caller assumes every row).
bool lattice_row_range(const cover_map* map, const double* polygon_xy, int nv, const cover_poses* poses, int& row_lo, int& row_hi) {
  const cover_lattice& L = poses->lattice;
  double radius = 0, scale = 0;
  for (int v = 0; v < nv; ++v) radius = std::max(radius, std::hypot(polygon_xy[2 * v], polygon_xy[2 * v + 1]));
  for (int t = 0; t < L.ntheta; ++t) scale = std::max(scale, std::hypot(L.cs[2 * t], L.cs[2 * t + 1]));  // |(c,s)| need not be 1
  const double reach = radius * scale;
  const double ya = L.y0, yb = L.y0 + static_cast<double>(L.ny - 1) * L.dy;
  const double lo = (std::min(ya, yb) - reach - map->view.origin_y) * map->view.inv_res - 2.0 + poses->cell_offset[1];  // (the offset moves every cell)
  const double hi = (std::max(ya, yb) + reach - map->view.origin_y) * map->view.inv_res + 2.0 + poses->cell_offset[1];
  row_lo = 0, row_hi = map->view.size_y - 1;
  if (!(std::isfinite(lo) && std::isfinite(hi))) return false;
  if (lo > row_lo) row_lo = lo > row_hi ? row_hi + 1 : static_cast<int>(std::floor(lo));
  if (hi < row_hi) row_hi = hi < 0 ? -1 : static_cast<int>(std::ceil(hi));
  return true;
}

/// Grid of k_lethal_planes for rows [row_lo, row_hi]: one warp per (32 rows, cols_per_warp word columns).
unsigned int planes_grid(const cover_ctx* ctx, int size_x, int row_lo, int row_hi, int cols_per_warp) {
  const long long tasks = static_cast<long long>((row_hi - row_lo) / 32 + 1) * (((size_x + 31) / 32 + cols_per_warp - 1) / cols_per_warp);
  return static_cast<unsigned int>(std::max<long long>(1, std::min<long long>((tasks + 7) / 8, static_cast<long long>(ctx->sm_count) * 16)));
}

/// Launches k_lethal_planes for rows [row_lo, row_hi]: few word columns per warp when the rows are few (a lattice window:
/// more warps in flight), eight when the whole of a big map is built.
void launch_planes(cover_ctx* ctx, const MapView& v, uint32_t* planes, int ypitch, long long plane_words, int row_lo, int row_hi, bool inflated,
                   const CallReset& reset) {
  const long long warps8 = static_cast<long long>((row_hi - row_lo) / 32 + 1) * (((v.size_x + 31) / 32 + 7) / 8);
  int cols = warps8 >= static_cast<long long>(ctx->sm_count) * 48 ? 8 : (warps8 >= static_cast<long long>(ctx->sm_count) * 24 ? 4 : 2);
  if (ctx->plane_cols > 0) cols = ctx->plane_cols;
  const unsigned int grid = planes_grid(ctx, v.size_x, row_lo, row_hi, cols);
  if (cols == 8) k_lethal_planes<8><<<grid, 256, 0, ctx->stream>>>(v, planes, ypitch, plane_words, row_lo, row_hi, inflated, reset);
  else if (cols == 4) k_lethal_planes<4><<<grid, 256, 0, ctx->stream>>>(v, planes, ypitch, plane_words, row_lo, row_hi, inflated, reset);
  else k_lethal_planes<2><<<grid, 256, 0, ctx->stream>>>(v, planes, ypitch, plane_words, row_lo, row_hi, inflated, reset);
}

/// Bit planes of rows [row_lo, row_hi] (built on ctx->stream when the current ones do not cover them): the LETHAL
/// set, or the INSCRIBED-or-LETHAL set.  Returns nullptr in `bits` when planes are disabled or cannot be allocated.
int ensure_plane_set(cover_ctx* ctx, const cover_map* map, bool inflated, int row_lo, int row_hi, const uint32_t*& bits) {
  bits = nullptr;
  if (ctx->disable_planes || row_lo > row_hi) return COVER_OK;
  uint32_t*& store = inflated ? map->planes_inf : map->planes;
  int& lo = inflated ? map->planes_inf_lo : map->planes_lo;
  int& hi = inflated ? map->planes_inf_hi : map->planes_hi;
  if (!store) {
    const size_t bytes = static_cast<size_t>(map->plane_words) * kPlaneCount * sizeof(uint32_t);
    uint32_t* p = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&p), bytes) != cudaSuccess) {
      cudaGetLastError();  // no memory for the acceleration structure: the byte walks still work
      return COVER_OK;
    }
    CU(cudaMemsetAsync(p, 0, bytes, ctx->stream));
    store = p;
    lo = 0, hi = -1;
  }
  if (row_lo < lo || row_hi > hi) {
    launch_planes(ctx, map->view, store, map->ypitch, map->plane_words, row_lo, row_hi, inflated, ctx->pending_reset);
    ctx->pending_reset = CallReset{nullptr, 0u, nullptr};
    ++ctx->launches;
    lo = row_lo, hi = row_hi;
  }
  bits = store;
  return COVER_OK;
}

/// Building planes costs a pass over the map rows: worth it for big batches, or when they are there already.
bool planes_pay_off(const cover_ctx* ctx, const cover_map* map, bool inflated, int row_lo, int row_hi, int64_t count) {
  if (ctx->disable_planes) return false;
  const bool ready = map->planes && row_lo >= map->planes_lo && row_hi <= map->planes_hi &&
                     (!inflated || (map->planes_inf && row_lo >= map->planes_inf_lo && row_hi <= map->planes_inf_hi));
  return ready || count >= ctx->planes_min_batch;
}

int ensure_planes(cover_ctx* ctx, const cover_map* map, int row_lo, int row_hi, PlaneView& pv) {
  const uint32_t* bits = nullptr;
  tl(ctx, "before planes");
  const int rc = ensure_plane_set(ctx, map, false, row_lo, row_hi, bits);
  tl(ctx, "planes ensured");
  pv = PlaneView{bits, map->ypitch, map->plane_words, nullptr, 0};
  return rc;
}

/// Prefix rows [row_lo, row_hi] for accumulate_cost lattice sweeps (built on ctx->stream when the current ones do not
/// cover them); nullptr when disabled or out of memory (4 bytes per cell).
int ensure_prefix(cover_ctx* ctx, const cover_map* map, int row_lo, int row_hi, const uint32_t*& prefix) {
  prefix = nullptr;
  if (ctx->disable_planes || row_lo > row_hi) return COVER_OK;
  if (!map->prefix) {
    uint32_t* p = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&p), static_cast<size_t>(map->ppitch) * map->view.size_y * sizeof(uint32_t)) != cudaSuccess) {
      cudaGetLastError();
      return COVER_OK;
    }
    map->prefix = p;
    map->prefix_lo = 0, map->prefix_hi = -1;
  }
  if (row_lo < map->prefix_lo || row_hi > map->prefix_hi) {
    const int rows = row_hi - row_lo + 1;
    k_cost_prefix<<<(rows + 7) / 8, 256, 0, ctx->stream>>>(map->view, map->prefix, map->ppitch, row_lo, row_hi);
    ++ctx->launches;
    map->prefix_lo = row_lo, map->prefix_hi = row_hi;
  }
  prefix = map->prefix;
  return COVER_OK;
}

void launch_max_planes(cover_ctx* ctx, const cover_map* map, int row_lo, int row_hi) {
  const MapView& v = map->view;
  const dim3 grid(static_cast<unsigned int>(((v.pitch >> 2) + 255) / 256), static_cast<unsigned int>(row_hi - row_lo + 1));
  k_max_planes<<<grid, 256, 0, ctx->stream>>>(v, map->maxp, static_cast<long long>(v.pitch) * v.size_y, row_lo);
  ++ctx->launches;
}

/// Sliding-maximum planes of rows [row_lo, row_hi] for max_cost lattice sweeps (built on ctx->stream when the current
/// ones do not cover them); nullptr when disabled, out of memory (6 bytes per cell) or beyond 32-bit offsets.
int ensure_max_planes(cover_ctx* ctx, const cover_map* map, int row_lo, int row_hi, const uint8_t*& maxp) {
  maxp = nullptr;
  if (ctx->disable_planes || row_lo > row_hi) return COVER_OK;
  if (static_cast<long long>(map->view.pitch) * map->view.size_y * 6 >= (1ll << 31)) return COVER_OK;  // (k_lw_values indexes the planes with 32 bits)
  if (!map->maxp) {
    uint8_t* p = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&p), static_cast<size_t>(map->view.pitch) * map->view.size_y * 6 + 64) != cudaSuccess) {
      cudaGetLastError();
      return COVER_OK;
    }
    map->maxp = p;
    map->maxp_lo = 0, map->maxp_hi = -1;
  }
  if (row_lo < map->maxp_lo || row_hi > map->maxp_hi) {
    launch_max_planes(ctx, map, row_lo, row_hi);
    map->maxp_lo = row_lo, map->maxp_hi = row_hi;
  }
  maxp = map->maxp;
  return COVER_OK;
}

/// The map as the span-scanning kernels see it for a verdict: with the bit planes of rows [row_lo, row_hi] attached
/// (`inflated`: both sets).  Only for launches whose poses cannot touch other rows of the map.
int view_with_planes(cover_ctx* ctx, const cover_map* map, bool inflated, int row_lo, int row_hi, MapView& view) {
  view = map->view;
  view.ypitch = map->ypitch, view.plane_words = map->plane_words;
  const uint32_t *lethal = nullptr, *inscribed = nullptr;
  int rc = ensure_plane_set(ctx, map, false, row_lo, row_hi, lethal);
  if (rc == COVER_OK && inflated) rc = ensure_plane_set(ctx, map, true, row_lo, row_hi, inscribed);
  view.planes = lethal, view.planes_inf = inflated ? inscribed : nullptr;
  if (inflated && !inscribed) view.planes = nullptr;  // (both sets or none: keep the two scans of a footprint consistent)
  return rc;
}

int finish_out(cover_ctx* ctx, const OutStage& st, int64_t count) {
  CU(cudaGetLastError());
  if (!st.host) return COVER_OK;
  const size_t n = static_cast<size_t>(count);
  if (n) {
    if (st.user->is_free) CU(cudaMemcpyAsync(st.user->is_free, st.dev.is_free, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (st.user->cost) CU(cudaMemcpyAsync(st.user->cost, st.dev.cost, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (st.user->status) CU(cudaMemcpyAsync(st.user->status, st.dev.status, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (st.user->free_bits) CU(cudaMemcpyAsync(st.user->free_bits, st.dev.free_bits, ((n + 31) / 32) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (st.user->not_ok) CU(cudaMemcpyAsync(st.user->not_ok, st.dev.not_ok, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (st.async) return COVER_OK;  // (the caller waits with cover_ctx_synchronize, which also lets a banded upload land)
  CU(cudaStreamSynchronize(ctx->stream));
  // cover_map_upload_async's contract: the caller may reuse its image buffer once a call has returned results to the
  // host.  A windowed sweep only waited (on the device) for the bands under its window - let the rest land first.
  CU(cudaStreamSynchronize(ctx->copy_stream));
  return COVER_OK;
}

int stage_poses(cover_ctx* ctx, const cover_poses* poses, PoseSource& src) {
  if (!poses) return fail(ctx, COVER_E_INVALID_ARG, "poses must not be NULL");
  if (poses->count < 0) return fail(ctx, COVER_E_INVALID_ARG, "negative pose count");
  std::memset(&src, 0, sizeof(src));
  src.format = poses->format;
  src.off_x = poses->cell_offset[0], src.off_y = poses->cell_offset[1];
  switch (poses->format) {
    case COVER_POSE_CS: {
      if (poses->count && !poses->data) return fail(ctx, COVER_E_INVALID_ARG, "POSE_CS needs data");
      if (poses->mem == COVER_MEM_DEVICE) {
        src.cs = static_cast<const double*>(poses->data);
      } else {
        const size_t bytes = static_cast<size_t>(poses->count) * 4 * sizeof(double);
        CU(ctx->poses.ensure(std::max<size_t>(bytes, 32)));
        if (bytes) CU(cudaMemcpyAsync(ctx->poses.ptr, poses->data, bytes, cudaMemcpyHostToDevice, ctx->stream));
        src.cs = ctx->poses.as<double>();
      }
      return COVER_OK;
    }
    case COVER_POSE_LATTICE: {
      const cover_lattice& L = poses->lattice;
      if (L.nx <= 0 || L.ny <= 0 || L.ntheta <= 0 || !L.cs) return fail(ctx, COVER_E_INVALID_ARG, "bad lattice descriptor");
      const int64_t total = static_cast<int64_t>(L.nx) * L.ny * L.ntheta;
      if (poses->first < 0 || poses->first + poses->count > total) return fail(ctx, COVER_E_INVALID_ARG, "lattice range out of bounds");
      const size_t bytes = static_cast<size_t>(L.ntheta) * 2 * sizeof(double);
      CU(ctx->theta.ensure(bytes));
      // (a planner sweeps the same heading table every cycle: the device copy is kept while the bytes are the same)
      if (ctx->theta_host.size() != bytes || ctx->theta_dev != ctx->theta.ptr || std::memcmp(ctx->theta_host.data(), L.cs, bytes) != 0) {
        CU(cudaMemcpyAsync(ctx->theta.ptr, L.cs, bytes, cudaMemcpyHostToDevice, ctx->stream));
        ctx->theta_host.assign(reinterpret_cast<const char*>(L.cs), bytes), ctx->theta_dev = ctx->theta.ptr;
      }
      src.theta_cs = ctx->theta.as<double>();
      src.first = poses->first;
      src.x0 = L.x0, src.dx = L.dx, src.y0 = L.y0, src.dy = L.dy;
      src.nx = L.nx, src.ny = L.ny;
      return COVER_OK;
    }
    case COVER_POSE_NONE:
      if (poses->count != 1) return fail(ctx, COVER_E_INVALID_ARG, "POSE_NONE is a single check (count must be 1)");
      return COVER_OK;
    default:
      return fail(ctx, COVER_E_INVALID_ARG, "unknown pose format");
  }
}

double polygon_diameter(const double* xy, int n) {
  double best = 0;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) best = std::max(best, std::hypot(xy[2 * i] - xy[2 * j], xy[2 * i + 1] - xy[2 * j + 1]));
  return best;
}

/// Upper bound of the number of scan rows any rigid pose of a polygon with this diameter can touch.
int rows_bound(double diameter, double inv_res) { return static_cast<int>(std::ceil(diameter * inv_res * 1.0001)) + 4; }

bool is_convex(const double* xy, int n) {
  if (n < 3) return true;
  int sign = 0;
  for (int i = 0; i < n; ++i) {
    const double ax = xy[2 * ((i + 1) % n)] - xy[2 * i], ay = xy[2 * ((i + 1) % n) + 1] - xy[2 * i + 1];
    const double bx = xy[2 * ((i + 2) % n)] - xy[2 * ((i + 1) % n)], by = xy[2 * ((i + 2) % n) + 1] - xy[2 * ((i + 1) % n) + 1];
    const double cr = ax * by - ay * bx;
    if (cr > 0) {
      if (sign < 0) return false;
      sign = 1;
    } else if (cr < 0) {
      if (sign > 0) return false;
      sign = -1;
    }
  }
  return true;
}

/// Dynamic shared memory of k_area_rows: polygon + per-warp ray records and vertices.
size_t rows_smem(int nv) {
  return static_cast<size_t>(nv) * 16 + kRowsWarps * (static_cast<size_t>(nv + 1) * sizeof(RayRec) + static_cast<size_t>(nv) * sizeof(int2));
}

unsigned int grid_for(int64_t count) { return static_cast<unsigned int>((count + kBlock - 1) / kBlock); }

/// Global scratch of the list store for `nthreads` resident threads.
int ensure_list_scratch(cover_ctx* ctx, int rows_cap, long long nthreads, ListScratch& s) {
  const size_t rows = static_cast<size_t>(rows_cap), nt = static_cast<size_t>(nthreads);
  CU(ctx->list_cnt.ensure(rows * nt * sizeof(int)));
  CU(ctx->list_rng.ensure(rows * kMaxRanges * nt * sizeof(int2)));
  CU(ctx->list_seq.ensure(rows * kMaxRanges * nt * sizeof(int)));
  s = ListScratch{ctx->list_cnt.as<int>(), ctx->list_rng.as<int2>(), ctx->list_seq.as<int>(), rows_cap};
  return COVER_OK;
}

/// Persistent grid for the list kernels: at most `blocks_per_sm` blocks per SM, capped so that the
/// scratch stays below ~2 GB.
unsigned int list_grid(const cover_ctx* ctx, int64_t items, int rows_cap, int blocks_per_sm) {
  const size_t per_thread = static_cast<size_t>(rows_cap) * (kMaxRanges * 12 + 4);
  int64_t blocks = static_cast<int64_t>(ctx->sm_count) * blocks_per_sm;
  const int64_t cap = std::max<int64_t>(1, static_cast<int64_t>((2ull << 30) / (per_thread * kBlock)));
  blocks = std::min(blocks, cap);
  blocks = std::min<int64_t>(blocks, std::max<int64_t>(1, (items + kBlock - 1) / kBlock));
  return static_cast<unsigned int>(blocks);
}

enum Shape { SHAPE_OUTLINE = 1, SHAPE_AREA = 2 };

constexpr int kMaxVertices = 4096;                                          // 64 KB of staged polygon
constexpr int kMaxDynamicSmem = kMaxVertices * 16 + kPackedMaxRows * kBlock * 8;  // 160 KB

/// Threads per CTA of k_area_quad: the table takes rows_cap x threads x 8 B of shared memory, so a tall polygon runs in
/// smaller CTAs to keep the warps per SM up (0: the table does not fit at all).
int quad_block_threads(size_t poly_bytes, int rows_cap) {
  int best = 0, best_warps = 0;
  for (int t : {128, 64, 32}) {
    const size_t smem = poly_bytes + static_cast<size_t>(rows_cap) * t * sizeof(unsigned long long);
    if (smem > static_cast<size_t>(kMaxDynamicSmem)) continue;
    const int ctas = static_cast<int>(std::min<size_t>({size_t{32}, (size_t{227} << 10) / (smem + 1024), static_cast<size_t>(1536 / t)}));
    const int warps = ctas * t / 32;
    if (warps * 10 > best_warps * 11) best = t, best_warps = warps;  // smaller CTAs only for > 10 % more warps
  }
  return best;
}

/// Opt every kernel that stages a polygon into its dynamic shared-memory budget (once per device).
cudaError_t configure_kernels() {
  cudaError_t e = cudaSuccess;
  auto set = [&e](const void* f) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynamicSmem);
  };
  set(reinterpret_cast<const void*>(&k_outline<OP_IS_FREE>));
  set(reinterpret_cast<const void*>(&k_outline<OP_ACCUMULATE>));
  set(reinterpret_cast<const void*>(&k_outline<OP_MAX>));
  set(reinterpret_cast<const void*>(&k_area_packed<OP_IS_FREE>));
  set(reinterpret_cast<const void*>(&k_area_packed<OP_ACCUMULATE>));
  set(reinterpret_cast<const void*>(&k_area_packed<OP_MAX>));
  set(reinterpret_cast<const void*>(&k_area_pair<OP_IS_FREE>));
  set(reinterpret_cast<const void*>(&k_area_pair<OP_ACCUMULATE>));
  set(reinterpret_cast<const void*>(&k_area_pair<OP_MAX>));
  set(reinterpret_cast<const void*>(&k_area_quad<OP_IS_FREE>));
  set(reinterpret_cast<const void*>(&k_area_quad<OP_ACCUMULATE>));
  set(reinterpret_cast<const void*>(&k_area_quad<OP_MAX>));
  set(reinterpret_cast<const void*>(&k_area_slots<OP_IS_FREE, 4>));
  set(reinterpret_cast<const void*>(&k_area_slots<OP_ACCUMULATE, 4>));
  set(reinterpret_cast<const void*>(&k_area_slots<OP_MAX, 4>));
  set(reinterpret_cast<const void*>(&k_area_list<OP_IS_FREE>));
  set(reinterpret_cast<const void*>(&k_area_list<OP_ACCUMULATE>));
  set(reinterpret_cast<const void*>(&k_area_list<OP_MAX>));
  set(reinterpret_cast<const void*>(&k_split));
  set(reinterpret_cast<const void*>(&k_split_lockstep));
  set(reinterpret_cast<const void*>(&k_area_rows<OP_IS_FREE>));
  set(reinterpret_cast<const void*>(&k_area_rows<OP_ACCUMULATE>));
  set(reinterpret_cast<const void*>(&k_area_rows<OP_MAX>));
  set(reinterpret_cast<const void*>(&k_area_lattice<OP_IS_FREE, false>));
  set(reinterpret_cast<const void*>(&k_area_lattice<OP_IS_FREE, true>));
  set(reinterpret_cast<const void*>(&k_area_lattice<OP_ACCUMULATE, false>));
  set(reinterpret_cast<const void*>(&k_area_lattice<OP_MAX, false>));
  set(reinterpret_cast<const void*>(&k_area_expand));
  set(reinterpret_cast<const void*>(&k_outline_expand));
  set(reinterpret_cast<const void*>(&k_inflate));
  set(reinterpret_cast<const void*>(&k_lethal_planes<2>));
  set(reinterpret_cast<const void*>(&k_lethal_planes<4>));
  set(reinterpret_cast<const void*>(&k_lethal_planes<8>));
  set(reinterpret_cast<const void*>(&k_area_rows_multi<OP_IS_FREE>));
  set(reinterpret_cast<const void*>(&k_area_rows_multi<OP_ACCUMULATE>));
  set(reinterpret_cast<const void*>(&k_area_rows_multi<OP_MAX>));
  set(reinterpret_cast<const void*>(&k_lw_prepare));
  set(reinterpret_cast<const void*>(&k_lw_sweep_simple<2>));
  set(reinterpret_cast<const void*>(&k_lw_sweep_simple<4>));
  set(reinterpret_cast<const void*>(&k_lw_sweep_simple<8>));
  return e;
}


/// Dynamic shared memory of k_lw_prepare: polygon, class tables of both axes, anchors, the shape warps' ray tables.
size_t lw_prepare_smem(int nv, int shape_warps) {
  return static_cast<size_t>(nv) * 16 + 2 * static_cast<size_t>(kLwClasses) * nv * sizeof(int) + kLwClasses * sizeof(int) +
         shape_warps * (static_cast<size_t>(nv + 1) * sizeof(RayRec) + static_cast<size_t>(nv) * sizeof(int2));
}

/// Start of a word sweep: the deferred-pose list {pushed, capacity} is emptied and the caller's not-OK counter zeroed
/// (one launch instead of a copy and a memset between the kernels).
__global__ void k_call_reset(unsigned int* __restrict__ header, unsigned int capacity, unsigned long long* __restrict__ not_ok) {
  if (threadIdx.x == 0) {
    header[0] = 0u, header[1] = capacity;
    if (not_ok) *not_ok = 0ull;
  }
}

/// Before a kernel that drains a deferred-pose list: if the list overflowed, that kernel takes EVERY pose of the batch
/// again (todo_items) - the failures the earlier passes counted would be counted twice, so the counter starts over.
__global__ void k_overflow_recount(const unsigned int* __restrict__ header, unsigned long long* __restrict__ not_ok) {
  if (threadIdx.x == 0 && header[0] > header[1]) *not_ok = 0ull;
}

struct LwPlan {
  LwTables t{};
  LwGeom g{};
  int nk = kPlaneCount;  // planes the simple sweep stages in shared memory: 5 while no span can be 32 cells wide, else all 6
};

/// First half of a bit-parallel lattice sweep (lattice_words.cuh): the class tables of every heading of the batch and one
/// rasterisation per distinct shape (k_lw_prepare).  None of it depends on the costmap's cells, so it runs on the
/// context's auxiliary stream, beside whatever the main stream still has to do to the map (upload, plane build).
int lw_prepare_async(cover_ctx* ctx, const cover_map* map, const double* poly_dev, int nv, bool outline, const PoseSource& src, int64_t count,
                     const OutView& out, LwPlan& plan) {
  const long long per_heading = static_cast<long long>(src.nx) * src.ny;
  LwGeom& g = plan.g;
  g.it_lo = static_cast<int>(src.first / per_heading);
  g.n_head = static_cast<int>((src.first + count - 1) / per_heading) - g.it_lo + 1;
  g.nb = (src.ny + 31) / 32;
  g.nbg = (g.nb + kLwWarps - 1) / kLwWarps;
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; };
  g.aligned4 = src.nx % 4 == 0 && src.first % 4 == 0 && aligned(out.is_free) && aligned(out.status);
  // the tables of the last sweep, if they were made from the same polygon, lattice range and map geometry
  std::string key = ctx->lw_key_base;
  key.append(reinterpret_cast<const char*>(&g.it_lo), sizeof g.it_lo).append(reinterpret_cast<const char*>(&g.n_head), sizeof g.n_head);
  if (!ctx->lw_cache_disabled && ctx->lw_cache_valid && !ctx->lw_key_base.empty() && key == ctx->lw_key) {
    const int aligned4 = g.aligned4;
    plan.t = ctx->lw_cached_t, plan.g = ctx->lw_cached_g;
    plan.g.aligned4 = aligned4;
    return COVER_OK;
  }
  ctx->lw_cache_valid = false;
  // one allocation: [counters | per-heading ints | anchors | shapes | spans | row tables | class bytes]
  LwTables& t = plan.t;
  t.max_shapes = std::min(4096, g.n_head * 64);
  t.span_pool = 1 << 19;
  const size_t nh = static_cast<size_t>(g.n_head);
  size_t off = 0;
  auto carve = [&off](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~static_cast<size_t>(255);
    return at;
  };
  const size_t o_counters = carve(16), o_nx = carve(nh * 4), o_ny = carve(nh * 4), o_base = carve(nh * 4), o_simple = carve(nh * 4), o_order = carve(nh * 4), o_heads = carve(nh * sizeof(LwHead));
  const size_t o_anchor = carve(nh * kLwClasses * 4), o_shapes = carve(static_cast<size_t>(t.max_shapes) * sizeof(LwShape));
  const size_t o_spans = carve(static_cast<size_t>(t.span_pool) * sizeof(LwSpan));
  const size_t o_yc0 = carve(nh * src.ny * 4), o_xcls = carve(nh * src.nx), o_ycls = carve(nh * src.ny);
  CU(ctx->lw_tables.ensure(off));
  char* base = ctx->lw_tables.as<char>();
  t.counters = reinterpret_cast<unsigned int*>(base + o_counters);
  t.n_xcls = reinterpret_cast<int*>(base + o_nx), t.n_ycls = reinterpret_cast<int*>(base + o_ny);
  t.shape_base = reinterpret_cast<int*>(base + o_base), t.simple = reinterpret_cast<int*>(base + o_simple), t.order = reinterpret_cast<int*>(base + o_order), t.heads = reinterpret_cast<LwHead*>(base + o_heads);
  t.shapes = reinterpret_cast<LwShape*>(base + o_shapes), t.spans = reinterpret_cast<LwSpan*>(base + o_spans);
  t.xanchor = reinterpret_cast<int*>(base + o_anchor), t.yc0 = reinterpret_cast<int*>(base + o_yc0);
  t.xcls = reinterpret_cast<uint8_t*>(base + o_xcls), t.ycls = reinterpret_cast<uint8_t*>(base + o_ycls);
  // fork: everything queued on the main stream so far (polygon / heading table uploads, the previous sweep that still
  // reads these tables) comes first
  CU(cudaEventRecord(ctx->aux_fork, ctx->stream));
  CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_fork, 0));
  CU(cudaMemsetAsync(base + o_counters, 0, 16, ctx->aux_stream));
  const size_t per_warp = static_cast<size_t>(nv + 1) * sizeof(RayRec) + static_cast<size_t>(nv) * sizeof(int2);
  const int shape_warps = static_cast<int>(std::max<size_t>(1, std::min<size_t>(16, (static_cast<size_t>(kMaxDynamicSmem) - lw_prepare_smem(nv, 0)) / per_warp)));
  k_lw_prepare<<<g.n_head, kLwPrepThreads, lw_prepare_smem(nv, shape_warps), ctx->aux_stream>>>(map->view, poly_dev, nv, src, g.it_lo, t, outline ? 1 : 0,
                                                                                              map->plane_words, shape_warps,
                                                                                              32 * (ctx->words_strip == 8 || ctx->words_strip == 2 ? ctx->words_strip : 4));
  ++ctx->launches;
  CU(cudaEventRecord(ctx->aux_join, ctx->aux_stream));
  if (!ctx->lw_key_base.empty()) ctx->lw_key = key, ctx->lw_cached_t = plan.t, ctx->lw_cached_g = plan.g, ctx->lw_cache_valid = true;
  return COVER_OK;
}

template <int W>
void launch_lw_sweeps(cover_ctx* ctx, const cover_map* map, const uint32_t* bits, const PoseSource& src, int64_t count, const OutView& out,
                      const LwTables& t, LwGeom g, int nk, bool simple, cudaStream_t stream) {
  g.ns = (src.nx + 32 * W - 1) / (32 * W);
  const dim3 grid(static_cast<unsigned int>(g.nbg), static_cast<unsigned int>(g.ns), static_cast<unsigned int>(g.n_head));
  if (simple) {
    // shared-memory tile of the planes (+ the transposition buffer of the byte outputs)
    const size_t smem = static_cast<size_t>(nk) * (W + kLwTileColsExtra) * kLwTileRows * sizeof(uint32_t) +
                        ((out.is_free || out.status) ? static_cast<size_t>(kLwWarps) * 32 * 2 * W * sizeof(uint32_t) : 0);
    k_lw_sweep_simple<W><<<grid, kLwWarps * 32, smem, stream>>>(map->view, bits, map->ypitch, map->plane_words, src, count, out, t, g, nk);
  } else {
    k_lw_sweep_general<W><<<grid, kLwWarps * 32, 0, stream>>>(map->view, bits, map->ypitch, map->plane_words, src, count, out, t, g,
                                                              ctx->todo.as<unsigned int>(), ctx->todo_count.as<unsigned int>());
  }
}

/// Second half: the sweep over the lethal bit planes `bits`, two launches side by side - the tasks of single-shape
/// headings inside the map on the main stream, the rest (few tasks, several walks each: latency-bound on its own) on the
/// auxiliary stream, where k_lw_prepare just finished.  Poses neither can finish land on ctx->todo (already reset);
/// `drain_on(stream)` launches the kernel that takes them.
template <class Drain>
int lw_sweep(cover_ctx* ctx, const cover_map* map, const PoseSource& src, int64_t count, const OutView& out, const LwPlan& plan, const uint32_t* bits,
             Drain&& drain_on) {
  auto launch = [&](bool simple, cudaStream_t stream) {
    if (ctx->words_strip == 8) launch_lw_sweeps<8>(ctx, map, bits, src, count, out, plan.t, plan.g, plan.nk, simple, stream);
    else if (ctx->words_strip == 2) launch_lw_sweeps<2>(ctx, map, bits, src, count, out, plan.t, plan.g, plan.nk, simple, stream);
    else launch_lw_sweeps<4>(ctx, map, bits, src, count, out, plan.t, plan.g, plan.nk, simple, stream);
  };
  // the auxiliary stream needs what the main stream did so far: planes, zeroed outputs and counters
  tl(ctx, "sweep: before fork");
  CU(cudaEventRecord(ctx->aux_planes, ctx->stream));
  CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_planes, 0));
  launch(false, ctx->aux_stream);
  {  // only the general half defers poses: their drain follows it on the auxiliary stream, beside the simple half
    const int rc = drain_on(ctx->aux_stream);
    if (rc) return rc;
  }
  CU(cudaEventRecord(ctx->aux_done, ctx->aux_stream));
  CU(cudaStreamWaitEvent(ctx->stream, ctx->aux_join, 0));  // the tables of k_lw_prepare
  if (ctx->zero_on_aux && out.free_bits) CU(cudaStreamWaitEvent(ctx->stream, ctx->aux_zero, 0));
  tl(ctx, "sweep: before simple");
  launch(true, ctx->stream);
  tl(ctx, "sweep: after simple");
  CU(cudaStreamWaitEvent(ctx->stream, ctx->aux_done, 0));
  tl(ctx, "sweep: joined general");
  ctx->launches += 2;
  if (std::getenv("COVER_B200_DEBUG_SYNC")) {  // (bring-up: a fault surfaces here, not at the next call)
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "k_lw_sweep");
  }
  return COVER_OK;
}

/// Capacity of the deferred-pose list of a lattice sweep (it defers next to nothing: poses at the map border, shapes
/// beyond the tables; if the list ever overflows the drain kernel takes every pose).
int64_t words_todo_cap(const cover_ctx* ctx, int64_t count) {
  int64_t cap = std::min<int64_t>(count, int64_t{4} << 20);
  if (ctx->todo_cap_limit > 0) cap = std::min(cap, ctx->todo_cap_limit);
  return cap;
}

template <int OP>
int launch_polygon(cover_ctx* ctx, const cover_map* map, const double* poly_dev, const double* poly_host, int nv, int shape,
                   const PoseSource& src, int64_t count, const OutView& out, int row_lo, int row_hi) {
  const size_t poly_bytes = static_cast<size_t>(nv) * 2 * sizeof(double);
  if (count == 0) return COVER_OK;
  const int rows_cap = rows_bound(polygon_diameter(poly_host, nv), map->view.inv_res);
  // Outlines over a pose lattice go through the shape-cache kernel with the outline's same-row runs as spans: merged
  // runs for a verdict or a maximum (any-reductions over the cell set), monotone pieces in traversal order for
  // accumulate_cost (first negative value in generator order, a cell counted once per visit).
  // Verdict sweeps of a lattice at cell pitch in x (any polygon up to 256 vertices, areas and outlines): the bit-parallel
  // kernels of lattice_words.cuh, when the lethal bit planes of the rows under the window are (or are worth being) built.
  const uint32_t* word_bits = nullptr;
  bool reset_done = false;
  ctx->zero_on_aux = false;
  LwPlan lw_plan;
  if (OP == OP_IS_FREE && src.format == POSE_LATTICE && nv >= 1 && nv <= kRowsMaxVertices && !ctx->disable_lattice_kernel &&
      !ctx->disable_words && !ctx->disable_planes && std::fabs(src.dx * map->view.inv_res - 1.0) < 1e-9 && count >= 4096 &&
      planes_pay_off(ctx, map, false, row_lo, row_hi, count) && map->plane_words * kPlaneCount < (1ll << 31) &&
      (src.first + count - 1) / (static_cast<long long>(src.nx) * src.ny) - src.first / (static_cast<long long>(src.nx) * src.ny) < 65535) {
    int rc = lw_prepare_async(ctx, map, poly_dev, nv, shape == SHAPE_OUTLINE, src, count, out, lw_plan);  // (beside the plane build)
    if (rc) return rc;
    lw_plan.nk = rows_cap - 4 < 32 ? kPlaneCount - 1 : kPlaneCount;  // (rows_cap bounds the bounding box's extent in cells, + 4)
    if (ctx->lazy_out && out.free_bits) {
      // the packed verdicts are zeroed on the auxiliary stream, beside the plane build (kernels OR their bits into shared words)
      CU(cudaEventRecord(ctx->aux_fork, ctx->stream));
      CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_fork, 0));
      CU(cudaMemsetAsync(out.free_bits, 0, static_cast<size_t>((count + 31) / 32) * sizeof(uint32_t), ctx->aux_stream));
      CU(cudaEventRecord(ctx->aux_zero, ctx->aux_stream));
      ctx->zero_on_aux = true;
    }
    if (ctx->lazy_out) {  // the plane build, if this call needs one, also resets the deferred-pose list and the not-OK counter
      CU(ctx->todo.ensure(static_cast<size_t>(words_todo_cap(ctx, count)) * sizeof(unsigned int)));
      CU(ctx->todo_count.ensure(2 * sizeof(unsigned int)));
      ctx->pending_reset = CallReset{ctx->todo_count.as<unsigned int>(), static_cast<unsigned int>(words_todo_cap(ctx, count)),
                                     static_cast<unsigned long long*>(ctx->lazy_not_ok)};
    }
    PlaneView pvw{nullptr, 0, 0, nullptr, 0};
    rc = ensure_planes(ctx, map, row_lo, row_hi, pvw);
    reset_done = ctx->lazy_out && ctx->pending_reset.header == nullptr;
    ctx->pending_reset = CallReset{nullptr, 0u, nullptr};
    if (reset_done) ctx->lazy_not_ok = nullptr;
    if (rc) return rc;
    word_bits = pvw.bits;  // (nullptr: no memory for the planes - the other kernels walk bytes then)
    if (!word_bits) CU(cudaStreamWaitEvent(ctx->stream, ctx->aux_join, 0));
  }
  const bool words = word_bits != nullptr;
  // accumulate_cost / max_cost of an area over such a lattice: the same class / shape tables, values from per-row prefix
  // words / sliding-maximum planes (k_lw_values)
  bool values = false;
  LwValueTables value_tables{nullptr, 0, nullptr, 0};
  if (OP != OP_IS_FREE && shape == SHAPE_AREA && src.format == POSE_LATTICE && nv >= 1 && nv <= kRowsMaxVertices && !ctx->disable_lattice_kernel &&
      !ctx->disable_words && !ctx->disable_values && !ctx->disable_planes && std::fabs(src.dx * map->view.inv_res - 1.0) < 1e-9 &&
      count >= std::max<int64_t>(4096, ctx->planes_min_batch) && static_cast<long long>(map->ppitch) * map->view.size_y < (1ll << 31) &&
      (src.first + count - 1) / (static_cast<long long>(src.nx) * src.ny) - src.first / (static_cast<long long>(src.nx) * src.ny) < 65535) {
    int rc = lw_prepare_async(ctx, map, poly_dev, nv, false, src, count, out, lw_plan);
    if (rc) return rc;
    if (OP == OP_ACCUMULATE) {
      const uint32_t* prefix = nullptr;
      rc = ensure_prefix(ctx, map, row_lo, row_hi, prefix);
      value_tables.prefix = prefix, value_tables.ppitch = map->ppitch;
      values = value_tables.prefix != nullptr;
    } else {
      const uint8_t* maxp = nullptr;
      rc = ensure_max_planes(ctx, map, row_lo, row_hi, maxp);
      value_tables.maxp = maxp;
      value_tables.max_plane_bytes = static_cast<long long>(map->view.pitch) * map->view.size_y;
      values = value_tables.maxp != nullptr;
    }
    if (rc) return rc;
    CU(cudaStreamWaitEvent(ctx->stream, ctx->aux_join, 0));  // the tables of k_lw_prepare
  }
  bool reset_by_kernel = false;
  if (ctx->lazy_out) {
    // every kernel ORs its bits into zeroed words of the packed verdicts (the word sweeps store the words they own entirely)
    if (out.free_bits && ctx->zero_on_aux && !words) CU(cudaStreamWaitEvent(ctx->stream, ctx->aux_zero, 0));  // (lw_sweep waits for it itself)
    if (out.free_bits && !ctx->zero_on_aux) CU(cudaMemsetAsync(out.free_bits, 0, static_cast<size_t>((count + 31) / 32) * sizeof(uint32_t), ctx->stream));
    if (words) {
      reset_by_kernel = true;  // (one launch zeroes the counter and resets the deferred-pose list)
    } else if (ctx->lazy_not_ok) {
      CU(cudaMemsetAsync(ctx->lazy_not_ok, 0, sizeof(unsigned long long), ctx->stream));
      ctx->lazy_not_ok = nullptr;
    }
  }
  const bool outline_lattice = shape == SHAPE_OUTLINE && src.format == POSE_LATTICE && nv >= 1 && !ctx->disable_lattice_kernel && count >= 4096 &&
                               (words || (nv <= kLatMaxVertices && rows_cap <= kLatMaxRowsCap && src.off_x == 0 && src.off_y == 0));
  if (shape == SHAPE_OUTLINE && !outline_lattice) {
    k_outline<OP><<<grid_for(count), kBlock, poly_bytes, ctx->stream>>>(map->view, poly_dev, nv, src, count, out);
    ++ctx->launches;
    return COVER_OK;
  }
  // verdicts scan row spans: attach the lethal bit planes of the rows this batch can touch (two loads per span)
  MapView area_view = map->view;
  if (OP == OP_IS_FREE && planes_pay_off(ctx, map, false, row_lo, row_hi, count)) {
    const int rc = view_with_planes(ctx, map, false, row_lo, row_hi, area_view);
    if (rc) return rc;
  }
  // 1. fast pass: poses whose scan rows all have <= 2 ranges (every pose of a convex polygon, many poses of a
  //    mildly concave one) are finished by the packed / lattice kernel; the rest lands on the todo list.
  //    Polygons that are far from convex skip it (the pass would defer almost everything).
  //    A non-convex polygon of moderate size goes through k_area_quad instead (four ranges per row in the same table).
  const int quad_threads = (!outline_lattice && !words && !values && !ctx->old_packed_kernel && !ctx->disable_quad && rows_cap <= ctx->quad_max_rows && !is_convex(poly_host, nv))
                               ? quad_block_threads(poly_bytes, rows_cap) : 0;
  const bool fast = outline_lattice || words || values || quad_threads > 0 || (rows_cap <= kPackedMaxRows && (is_convex(poly_host, nv) || ctx->always_try_packed));
  // 2. general pass over everything (or the todo list): one warp per pose when the outline has few rays
  //    relative to its rows, else one thread per pose with per-row lists in global scratch.
  const bool rows_kernel = nv <= kRowsMaxVertices && rows_cap <= 30000 && !ctx->force_list_kernel && (nv <= 16 || rows_cap > 48);
  const unsigned int* todo = nullptr;
  const unsigned int* todo_count = nullptr;
  bool second_list = false;
  // Lattice batches of small footprints: shape-deduplicating warp-cooperative kernel.
  const bool lattice = fast && src.format == POSE_LATTICE && nv >= 1 && nv <= kLatMaxVertices && src.off_x == 0 && src.off_y == 0 &&
                       rows_cap <= kLatMaxRowsCap && !ctx->disable_lattice_kernel;  // (the shape-cache kernel knows no cell offset)
  // The kernel that takes the deferred poses (or, without a fast pass, every pose), on `stream`.
  bool drained = false;
  auto drain = [&](cudaStream_t stream) -> int {
    if (todo_count && out.not_ok) {
      k_overflow_recount<<<1, 32, 0, stream>>>(todo_count, out.not_ok);
      ++ctx->launches;
    }
    if (outline_lattice) {  // one thread per deferred pose (map border for max_cost, shapes the cache cannot hold)
      const unsigned int grid = static_cast<unsigned int>(std::max<long long>(1, std::min<long long>(static_cast<long long>(ctx->sm_count) * 8, (count + kBlock - 1) / kBlock)));
      k_outline_drain<OP><<<grid, kBlock, poly_bytes, stream>>>(map->view, poly_dev, nv, src, count, out, todo, todo_count);
      ++ctx->launches;
      return COVER_OK;
    }
    // A convex polygon defers (almost) nothing: a small persistent grid is enough to drain the list.
    const bool small_drain = words || values || (fast && is_convex(poly_host, nv));  // (the word sweep defers next to nothing, whatever the polygon)
    if (rows_kernel) {
      const long long blocks = small_drain ? ctx->sm_count : static_cast<long long>(ctx->sm_count) * 12;
      const unsigned int grid = static_cast<unsigned int>(std::max<long long>(1, std::min<long long>(blocks, (count + kRowsWarps - 1) / kRowsWarps)));
      k_area_rows<OP><<<grid, kRowsWarps * 32, rows_smem(nv), stream>>>(area_view, poly_dev, nv, src, count, out, todo, todo_count);
    } else {
      ListScratch scratch{};
      const unsigned int grid = list_grid(ctx, count, rows_cap, small_drain ? 1 : 4);
      const int rc = ensure_list_scratch(ctx, rows_cap, static_cast<long long>(grid) * kBlock, scratch);
      if (rc) return rc;
      k_area_list<OP><<<grid, kBlock, poly_bytes, stream>>>(area_view, poly_dev, nv, src, count, out, scratch, todo, todo_count);
    }
    ++ctx->launches;
    return COVER_OK;
  };
  if (fast) {
    // deferred-pose list: {pushed, capacity} + the entries.  A lattice sweep defers next to nothing (poses at the map
    // border, shapes beyond the tables): a small list, and if it ever overflows the drain kernel takes every pose.  The
    // per-pose fast passes may defer most of a batch (concave polygons): room for all of it.
    int64_t todo_cap = (words || values || (src.format == POSE_LATTICE && !ctx->disable_lattice_kernel)) ? words_todo_cap(ctx, count) : count;
    if (ctx->todo_cap_limit > 0) todo_cap = std::min(todo_cap, ctx->todo_cap_limit);
    CU(ctx->todo.ensure(static_cast<size_t>(todo_cap) * sizeof(unsigned int)));
    CU(ctx->todo_count.ensure(2 * sizeof(unsigned int)));
    ctx->todo_header[0] = 0u, ctx->todo_header[1] = static_cast<unsigned int>(todo_cap);
    if (reset_by_kernel && reset_done) {
      // (k_lethal_planes did it on the way)
    } else if (reset_by_kernel) {
      // on the auxiliary stream, in front of the general half and the drain - the only kernels of a word sweep that push
      // poses or count failures (the simple half on the main stream does neither), and off the main stream's critical path
      k_call_reset<<<1, 32, 0, ctx->aux_stream>>>(ctx->todo_count.as<unsigned int>(), ctx->todo_header[1], static_cast<unsigned long long*>(ctx->lazy_not_ok));
      ctx->lazy_not_ok = nullptr;
    } else {
      CU(cudaMemcpyAsync(ctx->todo_count.ptr, ctx->todo_header, 2 * sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
    }
    tl(ctx, "todo header reset");
    if (words) {
      todo = ctx->todo.as<unsigned int>(), todo_count = ctx->todo_count.as<unsigned int>();
      const int rc = lw_sweep(ctx, map, src, count, out, lw_plan, word_bits, drain);
      if (rc) return rc;
      drained = true;
      --ctx->launches;  // (counted once more below)
    } else if (values) {
      const dim3 grid(static_cast<unsigned int>((src.ny + 7) / 8), static_cast<unsigned int>((src.nx + 63) / 64), static_cast<unsigned int>(lw_plan.g.n_head));
      k_lw_values<OP == OP_IS_FREE ? OP_MAX : OP><<<grid, 256, 0, ctx->stream>>>(map->view, value_tables, src, count, out, lw_plan.t, lw_plan.g,
                                                                              ctx->todo.as<unsigned int>(), ctx->todo_count.as<unsigned int>());
    } else if (lattice) {
      LatGeom g{};
      g.row_first = src.first / src.nx;
      const long long row_last = (src.first + count - 1) / src.nx;
      g.n_rows = static_cast<int>(row_last - g.row_first + 1);
      g.segs = (src.nx + kLatSegX - 1) / kLatSegX;
      g.rows_per_unit = 32;  // more rows per unit amortise the per-unit column work; fewer keep a small batch parallel
      const long long want_units = 4ll * ctx->sm_count * LAT_MIN_BLOCKS * kLatWarps;
      while (g.rows_per_unit > 8 && static_cast<long long>((g.n_rows + g.rows_per_unit - 1) / g.rows_per_unit) * g.segs < want_units) g.rows_per_unit /= 2;
      g.n_units = static_cast<long long>((g.n_rows + g.rows_per_unit - 1) / g.rows_per_unit) * g.segs;
      // span table per cached shape: one span per scan row for a convex polygon, up to two (and the [row][4] scratch
      // of the four-range rasteriser) otherwise
      const int span_cap = outline_lattice ? 2 * rows_cap + 8 : ((is_convex(poly_host, nv) && rows_cap - 3 <= kLatMaxWidth) ? rows_cap : 2 * rows_cap);
      g.outline = outline_lattice ? (OP == OP_ACCUMULATE ? 2 : 1) : 0;
      const size_t per_slot = sizeof(unsigned long long) + static_cast<size_t>(nv) * sizeof(int2) + static_cast<size_t>(span_cap) * 2 * sizeof(uint32_t) + sizeof(LatShape);
      g.n_slots = kLatSlotsMax;  // (a heading has a handful of distinct shapes: 16 slots are plenty for the big ones)
      while (g.n_slots > 16 && g.n_slots * per_slot > (28u << 10)) g.n_slots /= 2;
      const size_t smem = poly_bytes + g.n_slots * per_slot +
                          (OP == OP_IS_FREE ? 0 : kLatWarps * 800);  // the value folds' per-warp query tables
      // (the lean verdict kernel - at most 4 vertices, planes on, at most 32 spans per shape - runs more CTAs per SM)
      const bool lean_kernel = OP == OP_IS_FREE && !ctx->disable_planes && nv <= 4 && span_cap <= 32;
      const unsigned int grid = static_cast<unsigned int>(std::min<long long>((g.n_units + kLatWarps - 1) / kLatWarps,
                                                                              static_cast<long long>(ctx->sm_count) * (lean_kernel ? LAT_LEAN_BLOCKS : LAT_MIN_BLOCKS)));
      // launch-wide shape table + the round counter of the dynamic hand-out: [counter | keys | statuses | entries]
      constexpr int kGlobalShapes = 2048;
      LatGlobal gl{};
      gl.capacity = kGlobalShapes, gl.entry_words = 6 + 2 * nv + 2 * span_cap;
      const size_t head_bytes = 16 + static_cast<size_t>(kGlobalShapes) * (sizeof(unsigned long long) + sizeof(int));
      CU(ctx->lat_global.ensure(head_bytes + static_cast<size_t>(kGlobalShapes) * gl.entry_words * sizeof(int)));
      CU(cudaMemsetAsync(ctx->lat_global.ptr, 0, head_bytes, ctx->stream));
      gl.next_round = ctx->lat_global.as<unsigned long long>();
      gl.key = gl.next_round + 2;
      gl.status = reinterpret_cast<int*>(gl.key + kGlobalShapes);
      gl.data = gl.status + kGlobalShapes;
      {
        const long long total_rounds = (g.n_units + kLatWarps - 1) / kLatWarps;
        g.static_rounds = static_cast<int>(std::max<long long>(1, std::min<long long>(ctx->max_stretch, total_rounds / (2ll * grid))));
        g.tickets_a = total_rounds / 2 / g.static_rounds;                                   // half of the batch in long stretches,
        g.tickets_b = total_rounds / 4 / std::max(1, g.static_rounds / 2);                  // a quarter in half-length ones
      }
      PlaneView pv{nullptr, 0, 0, nullptr, 0};
      if (OP == OP_IS_FREE) {  // verdicts read the lethal bit planes of the rows under the window (built on first use)
        const int rc = ensure_planes(ctx, map, row_lo, row_hi, pv);
        if (rc) return rc;
      } else if (OP == OP_ACCUMULATE && count >= ctx->planes_min_batch) {  // sums read the prefix rows
        const uint32_t* prefix = nullptr;
        const int rc = ensure_prefix(ctx, map, row_lo, row_hi, prefix);
        if (rc) return rc;
        pv.prefix = prefix, pv.ppitch = map->ppitch;
      }
      if (lean_kernel && pv.bits)  // (the lean kernel defers shapes with more than 32 spans)
        k_area_lattice<OP_IS_FREE, true><<<grid, kLatThreads, smem, ctx->stream>>>(map->view, poly_dev, nv, src, count, out, rows_cap, span_cap,
                                                                                   ctx->todo.as<unsigned int>(), ctx->todo_count.as<unsigned int>(), g, pv, gl);
      else
        k_area_lattice<OP, false><<<grid, kLatThreads, smem, ctx->stream>>>(map->view, poly_dev, nv, src, count, out, rows_cap, span_cap,
                                                                            ctx->todo.as<unsigned int>(), ctx->todo_count.as<unsigned int>(), g, pv, gl);
    } else {
      if (quad_threads > 0) {
        const size_t smem = poly_bytes + static_cast<size_t>(rows_cap) * quad_threads * sizeof(unsigned long long);
        const unsigned int grid = static_cast<unsigned int>((count + quad_threads - 1) / quad_threads);
        k_area_quad<OP><<<grid, quad_threads, smem, ctx->stream>>>(area_view, poly_dev, nv, src, count, out, rows_cap, ctx->todo.as<unsigned int>(),
                                                                   ctx->todo_count.as<unsigned int>());
      } else if (ctx->old_packed_kernel) {
        const size_t smem = poly_bytes + static_cast<size_t>(rows_cap) * kBlock * sizeof(uint32_t);
        k_area_packed<OP><<<grid_for(count), kBlock, smem, ctx->stream>>>(map->view, poly_dev, nv, src, count, out, rows_cap,
                                                                          ctx->todo.as<unsigned int>(), ctx->todo_count.as<unsigned int>());
      } else {
        const size_t smem = poly_bytes + static_cast<size_t>(rows_cap) * kBlock * sizeof(uint2);
        k_area_pair<OP><<<grid_for(count), kBlock, smem, ctx->stream>>>(area_view, poly_dev, nv, src, count, out, rows_cap,
                                                                        ctx->todo.as<unsigned int>(), ctx->todo_count.as<unsigned int>());
        if (!is_convex(poly_host, nv) && rows_cap <= 64) {
          // second per-thread pass over the deferred poses with four ranges per row; its own overflow list feeds the
          // general kernel below
          CU(ctx->todo2.ensure(static_cast<size_t>(count) * sizeof(unsigned int)));  // (a concave polygon may defer most of the batch twice)
          CU(ctx->todo2_count.ensure(2 * sizeof(unsigned int)));
          ctx->todo2_header[0] = 0u, ctx->todo2_header[1] = static_cast<unsigned int>(ctx->todo_cap_limit > 0 ? std::min<int64_t>(count, ctx->todo_cap_limit) : count);
          CU(cudaMemcpyAsync(ctx->todo2_count.ptr, ctx->todo2_header, 2 * sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
          const size_t smem4 = poly_bytes + static_cast<size_t>(rows_cap) * kBlock * 4 * sizeof(uint32_t);
          const unsigned int grid4 = static_cast<unsigned int>(std::max<long long>(1, std::min<long long>(static_cast<long long>(ctx->sm_count) * 4, (count + kBlock - 1) / kBlock)));
          if (out.not_ok) {
            k_overflow_recount<<<1, 32, 0, ctx->stream>>>(ctx->todo_count.as<unsigned int>(), out.not_ok);
            ++ctx->launches;
          }
          k_area_slots<OP, 4><<<grid4, kBlock, smem4, ctx->stream>>>(area_view, poly_dev, nv, src, count, out, rows_cap, ctx->todo.as<unsigned int>(),
                                                                    ctx->todo_count.as<unsigned int>(), ctx->todo2.as<unsigned int>(),
                                                                    ctx->todo2_count.as<unsigned int>());
          ++ctx->launches;
          second_list = true;
        }
      }
    }
    ++ctx->launches;
    todo = second_list ? ctx->todo2.as<unsigned int>() : ctx->todo.as<unsigned int>();
    todo_count = second_list ? ctx->todo2_count.as<unsigned int>() : ctx->todo_count.as<unsigned int>();
  }
  if (drained) return COVER_OK;
  return drain(ctx->stream);
}

int polygon_check(cover_ctx* ctx, const cover_map* map, const double* polygon_xy, int32_t nv, int shape, int op,
                  const cover_poses* poses, const cover_results* out) {
  if (!ctx) return COVER_E_INVALID_ARG;
  if (!map || map->ctx != ctx) return fail(ctx, COVER_E_INVALID_ARG, "map does not belong to this context");
  if (nv < 0 || (nv > 0 && !polygon_xy)) return fail(ctx, COVER_E_INVALID_ARG, "bad polygon");
  if (nv > kMaxVertices) return fail(ctx, COVER_E_UNSUPPORTED, "polygons with more than 4096 vertices are not supported");
  if (poses && poses->count > 0xFFFFFFFFll) return fail(ctx, COVER_E_UNSUPPORTED, "more than 2^32 - 1 poses in one call (deferred poses are indexed with 32 bits): split the batch");
  Bind bind(ctx->device);
  tl(ctx, "polygon_check: entry");
  PoseSource src;
  int rc = stage_poses(ctx, poses, src);
  if (rc) return rc;
  tl(ctx, "poses staged");
  OutStage st;
  // (a verdict sweep of a lattice writes whole words of the packed verdicts: launch_polygon decides what needs zeroing)
  rc = stage_out(ctx, out, poses->count, st, op == OP_IS_FREE && poses->format == COVER_POSE_LATTICE && poses->count > 0);
  if (rc) return rc;
  tl(ctx, "outputs staged");
  const size_t poly_bytes = static_cast<size_t>(nv) * 2 * sizeof(double);
  CU(ctx->poly.ensure(std::max<size_t>(poly_bytes, 16)));
  if (poly_bytes) CU(cudaMemcpyAsync(ctx->poly.ptr, polygon_xy, poly_bytes, cudaMemcpyHostToDevice, ctx->stream));
  tl(ctx, "polygon uploaded");
  const double* pd = ctx->poly.as<double>();
  ctx->lw_key_base.clear();
  if (poses->format == COVER_POSE_LATTICE && poses->lattice.cs) {  // everything the class / shape tables of a lattice sweep depend on
    auto put = [ctx](const void* p, size_t n) { ctx->lw_key_base.append(static_cast<const char*>(p), n); };
    const cover_lattice& L = poses->lattice;
    const MapView& v = map->view;
    put(&nv, sizeof nv), put(&shape, sizeof shape), put(polygon_xy, poly_bytes);
    put(&L.x0, sizeof L.x0), put(&L.dx, sizeof L.dx), put(&L.y0, sizeof L.y0), put(&L.dy, sizeof L.dy);
    put(&L.nx, sizeof L.nx), put(&L.ny, sizeof L.ny), put(&L.ntheta, sizeof L.ntheta), put(L.cs, static_cast<size_t>(L.ntheta) * 2 * sizeof(double));
    put(poses->cell_offset, sizeof poses->cell_offset);
    put(&v.size_x, sizeof v.size_x), put(&v.size_y, sizeof v.size_y), put(&v.inv_res, sizeof v.inv_res), put(&v.origin_x, sizeof v.origin_x), put(&v.origin_y, sizeof v.origin_y);
    put(&map->plane_words, sizeof map->plane_words);
  }
  // costmap upload still in flight: a lattice sweep starts as soon as the rows under its window have landed
  int row_lo = 0, row_hi = map->view.size_y - 1;  // map rows the check can read
  if (poses->format == COVER_POSE_LATTICE && poses->lattice.cs && lattice_row_range(map, polygon_xy, nv, poses, row_lo, row_hi))
    CU(map_rows_ready(ctx, map, row_lo, row_hi));
  else
    CU(map_ready(ctx, map));
  auto launch = [&](const PoseSource& s, int64_t n, const OutView& o) {
    switch (op) {
      case OP_IS_FREE: return launch_polygon<OP_IS_FREE>(ctx, map, pd, polygon_xy, nv, shape, s, n, o, row_lo, row_hi);
      case OP_ACCUMULATE: return launch_polygon<OP_ACCUMULATE>(ctx, map, pd, polygon_xy, nv, shape, s, n, o, row_lo, row_hi);
      default: return launch_polygon<OP_MAX>(ctx, map, pd, polygon_xy, nv, shape, s, n, o, row_lo, row_hi);
    }
  };
  // A big lattice sweep whose results go to the host runs as 2..4 contiguous index chunks: chunk k's results travel
  // back on a second stream while chunk k + 1 is computed (chunk borders are multiples of 32 poses = whole bit words).
  // (a 1e8-pose verdict sweep computes in ~0.1 ms, its 12.5 MB of bits travel 0.22 ms: nothing to hide behind)
  const int64_t kPipelineMin = ctx->pipeline_min > 0 ? ctx->pipeline_min : (256ll << 20);
  // (two chunks up to ~1.5e8 poses: every chunk costs ~50 us of launch overhead and kernel tail, a chunk's download ~0.1 ms)
  int kPipelineChunks = static_cast<int>(std::min<int64_t>(4, std::max<int64_t>(2, poses->count / (48ll << 20))));
  if (ctx->pipeline_chunks > 0) kPipelineChunks = ctx->pipeline_chunks;  // (measurements)
  if (!(st.host && poses->format == COVER_POSE_LATTICE && poses->count >= kPipelineMin)) {
    rc = launch(src, poses->count, st.dev);
    if (rc) return rc;
    return finish_out(ctx, st, poses->count);
  }
  const int64_t per = ((poses->count + kPipelineChunks - 1) / kPipelineChunks + 31) & ~static_cast<int64_t>(31);
  for (int64_t off = 0; off < poses->count; off += per) {
    const int64_t n = std::min<int64_t>(per, poses->count - off);
    PoseSource s = src;
    s.first = src.first + off;
    OutView o = st.dev;
    if (o.is_free) o.is_free += off;
    if (o.cost) o.cost += off;
    if (o.status) o.status += off;
    if (o.free_bits) o.free_bits += off / 32;
    rc = launch(s, n, o);
    if (rc) return rc;
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->chunk_done, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->d2h_stream, ctx->chunk_done, 0));
    const size_t m = static_cast<size_t>(n);
    if (out->is_free) CU(cudaMemcpyAsync(out->is_free + off, o.is_free, m, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    if (out->cost) CU(cudaMemcpyAsync(out->cost + off, o.cost, m * sizeof(double), cudaMemcpyDeviceToHost, ctx->d2h_stream));
    if (out->status) CU(cudaMemcpyAsync(out->status + off, o.status, m, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    if (out->free_bits) CU(cudaMemcpyAsync(out->free_bits + off / 32, o.free_bits, ((m + 31) / 32) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->d2h_stream));
  }
  if (out->not_ok) CU(cudaMemcpyAsync(out->not_ok, st.dev.not_ok, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaStreamSynchronize(ctx->d2h_stream));
  CU(cudaStreamSynchronize(ctx->copy_stream));  // (see finish_out)
  return COVER_OK;
}

/// Copies a host or device int array into a context buffer when needed; returns the device pointer.
template <class T>
int stage_array(cover_ctx* ctx, DeviceBuffer& buf, const T* data, size_t n, int mem, const T** dev) {
  if (mem == COVER_MEM_DEVICE) {
    *dev = data;
    return COVER_OK;
  }
  CU(buf.ensure(std::max<size_t>(n * sizeof(T), 16)));
  if (n) CU(cudaMemcpyAsync(buf.ptr, data, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *dev = buf.as<T>();
  return COVER_OK;
}

int rays_check(cover_ctx* ctx, const cover_map* map, const int32_t* segments, int32_t mem, int64_t count, const int32_t* offset,
               int op, const cover_results* out) {
  if (!ctx) return COVER_E_INVALID_ARG;
  if (!map || map->ctx != ctx) return fail(ctx, COVER_E_INVALID_ARG, "map does not belong to this context");
  if (count < 0 || (count && !segments)) return fail(ctx, COVER_E_INVALID_ARG, "bad segments");
  Bind bind(ctx->device);
  OutStage st;
  int rc = stage_out(ctx, out, count, st);
  if (rc) return rc;
  const int32_t* dev = nullptr;
  rc = stage_array(ctx, ctx->in_a, segments, static_cast<size_t>(count) * 4, mem, &dev);
  if (rc) return rc;
  const int ox = offset ? offset[0] : 0, oy = offset ? offset[1] : 0;
  if (count) {
    const int4* seg = reinterpret_cast<const int4*>(dev);
    CU(map_ready(ctx, map));
    if (op == OP_IS_FREE) k_rays<OP_IS_FREE><<<grid_for(count), kBlock, 0, ctx->stream>>>(map->view, seg, count, ox, oy, st.dev);
    else if (op == OP_ACCUMULATE) k_rays<OP_ACCUMULATE><<<grid_for(count), kBlock, 0, ctx->stream>>>(map->view, seg, count, ox, oy, st.dev);
    else k_rays<OP_MAX><<<grid_for(count), kBlock, 0, ctx->stream>>>(map->view, seg, count, ox, oy, st.dev);
    ++ctx->launches;
  }
  return finish_out(ctx, st, count);
}

int cells_check(cover_ctx* ctx, const cover_map* map, const cover_cells* cells, const int32_t* offsets, int32_t mem, int64_t count,
                int op, const cover_results* out) {
  if (!ctx) return COVER_E_INVALID_ARG;
  if (!map || map->ctx != ctx || !cells || cells->ctx != ctx) return fail(ctx, COVER_E_INVALID_ARG, "handle does not belong to this context");
  if (count < 0 || (count && !offsets)) return fail(ctx, COVER_E_INVALID_ARG, "bad offsets");
  Bind bind(ctx->device);
  OutStage st;
  int rc = stage_out(ctx, out, count, st);
  if (rc) return rc;
  const int32_t* dev = nullptr;
  rc = stage_array(ctx, ctx->in_a, offsets, static_cast<size_t>(count) * 2, mem, &dev);
  if (rc) return rc;
  if (count) {
    const int2* off = reinterpret_cast<const int2*>(dev);
    CU(map_ready(ctx, map));
    if (op == OP_IS_FREE) {
      MapView view = map->view;
      if (planes_pay_off(ctx, map, false, 0, view.size_y - 1, count)) {
        rc = view_with_planes(ctx, map, false, 0, view.size_y - 1, view);
        if (rc) return rc;
      }
      k_cells_free<<<grid_for(count), kBlock, 0, ctx->stream>>>(view, cells->spans, cells->n_spans, off, count, st.dev);
    }
    else if (op == OP_ACCUMULATE) k_cells<OP_ACCUMULATE><<<grid_for(count), kBlock, 0, ctx->stream>>>(map->view, cells->dev, cells->n, off, count, st.dev);
    else k_cells<OP_MAX><<<grid_for(count), kBlock, 0, ctx->stream>>>(map->view, cells->dev, cells->n, off, count, st.dev);
    ++ctx->launches;
  }
  return finish_out(ctx, st, count);
}

template <class T>
int upload_new(cover_ctx* ctx, const T* host, size_t n, T** dev) {
  *dev = nullptr;
  CU(cudaMalloc(reinterpret_cast<void**>(dev), std::max<size_t>(n * sizeof(T), 16)));
  if (n) CU(cudaMemcpyAsync(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return COVER_OK;
}

}  // namespace

// ===================================================================================================
// C ABI
// ===================================================================================================


namespace {
/// Shared implementation of cover_area_expand / cover_outline_expand (two kernel passes + a host prefix sum).
int expand_cells(cover_ctx* ctx, const cover_map* map, const double* polygon_xy, int32_t nv, bool area, const cover_poses* poses,
                 int32_t* cells, int64_t capacity, int64_t* starts, uint8_t* status, int32_t mem) {
  if (!ctx) return COVER_E_INVALID_ARG;
  if (!map || map->ctx != ctx) return fail(ctx, COVER_E_INVALID_ARG, "map does not belong to this context");
  if (nv < 0 || (nv > 0 && !polygon_xy) || !starts) return fail(ctx, COVER_E_INVALID_ARG, "bad arguments");
  if (nv > kRowsMaxVertices) return fail(ctx, COVER_E_UNSUPPORTED, "expand supports at most 256 vertices");
  if (mem != COVER_MEM_HOST) return fail(ctx, COVER_E_UNSUPPORTED, "expand returns host arrays (it is a per-footprint precompute)");
  Bind bind(ctx->device);
  PoseSource src;
  int rc = stage_poses(ctx, poses, src);
  if (rc) return rc;
  const int64_t n = poses->count;
  const size_t poly_bytes = static_cast<size_t>(nv) * 2 * sizeof(double);
  CU(ctx->poly.ensure(std::max<size_t>(poly_bytes, 16)));
  if (poly_bytes) CU(cudaMemcpyAsync(ctx->poly.ptr, polygon_xy, poly_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx->in_a.ensure(static_cast<size_t>(n + 1) * sizeof(long long)));  // counts, then starts
  CU(ctx->out_status.ensure(static_cast<size_t>(std::max<int64_t>(n, 1))));
  long long* d_counts = ctx->in_a.as<long long>();
  uint8_t* d_status = ctx->out_status.as<uint8_t>();
  const unsigned int rows_grid = static_cast<unsigned int>(std::max<long long>(1, std::min<long long>(static_cast<long long>(ctx->sm_count) * 12, (n + kRowsWarps - 1) / kRowsWarps)));
  auto launch = [&](const long long* d_starts, int2* d_cells) {
    if (n == 0) return;
    if (area) k_area_expand<<<rows_grid, kRowsWarps * 32, rows_smem(nv), ctx->stream>>>(map->view, ctx->poly.as<double>(), nv, src, n, d_counts, d_status, d_starts, d_cells);
    else k_outline_expand<<<grid_for(n), kBlock, poly_bytes, ctx->stream>>>(map->view, ctx->poly.as<double>(), nv, src, n, d_counts, d_status, d_starts, d_cells);
    ++ctx->launches;
  };
  launch(nullptr, nullptr);
  CU(cudaGetLastError());
  std::vector<long long> counts(static_cast<size_t>(n));
  if (n) CU(cudaMemcpyAsync(counts.data(), d_counts, static_cast<size_t>(n) * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  if (status && n) CU(cudaMemcpyAsync(status, d_status, static_cast<size_t>(n), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  std::vector<long long> st(static_cast<size_t>(n) + 1, 0);
  for (int64_t i = 0; i < n; ++i) st[i + 1] = st[i] + counts[i];
  for (int64_t i = 0; i <= n; ++i) starts[i] = st[i];
  if (!cells) return COVER_OK;  // size query
  if (st[n] > capacity) return fail(ctx, COVER_E_INVALID_ARG, "cells buffer too small (call with cells = NULL to get the size)");
  if (st[n] == 0) return COVER_OK;
  CU(ctx->in_b.ensure(static_cast<size_t>(st[n]) * sizeof(int2)));
  CU(cudaMemcpyAsync(d_counts, st.data(), st.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  launch(d_counts, ctx->in_b.as<int2>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(cells, ctx->in_b.ptr, static_cast<size_t>(st[n]) * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return COVER_OK;
}
}  // namespace

namespace {
/// Cells [b, e) of a 2 x N list -> sorted, merged row spans {y, x_lo, x_hi, 0} appended to `spans`.
void append_row_spans(std::vector<int4>& spans, const int32_t* cells, int64_t b, int64_t e) {
  std::vector<std::pair<int, int>> yx;
  yx.reserve(static_cast<size_t>(e - b));
  for (int64_t i = b; i < e; ++i) yx.emplace_back(cells[2 * i + 1], cells[2 * i]);
  std::sort(yx.begin(), yx.end());
  yx.erase(std::unique(yx.begin(), yx.end()), yx.end());
  for (size_t i = 0; i < yx.size();) {
    size_t j = i;
    while (j + 1 < yx.size() && yx[j + 1].first == yx[i].first && yx[j + 1].second == yx[j].second + 1) ++j;
    spans.push_back(make_int4(yx[i].first, yx[i].second, yx[j].second, 0));
    i = j + 1;
  }
}
}  // namespace

extern "C" {

int cover_abi_version(void) { return COVER_B200_ABI_VERSION; }

int cover_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int cover_ctx_create(int device, cover_ctx** out) {
  if (!out) return COVER_E_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return COVER_E_CUDA;  // no CPU fallback
  cover_ctx* ctx = new (std::nothrow) cover_ctx();
  if (!ctx) return COVER_E_NOMEM;
  ctx->device = device;
  Bind bind(device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || configure_kernels() != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->fence, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->chunk_done, cudaEventDisableTiming) != cudaSuccess ||
      create_priority_stream(&ctx->aux_stream) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->aux_planes, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->aux_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->aux_zero, cudaEventDisableTiming) != cudaSuccess) {
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
    if (ctx->aux_join) cudaEventDestroy(ctx->aux_join);
    if (ctx->aux_planes) cudaEventDestroy(ctx->aux_planes);
    if (ctx->aux_done) cudaEventDestroy(ctx->aux_done);
    if (ctx->aux_zero) cudaEventDestroy(ctx->aux_zero);
    if (ctx->chunk_done) cudaEventDestroy(ctx->chunk_done);
    if (ctx->fence) cudaEventDestroy(ctx->fence);
    delete ctx;
    return COVER_E_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  if (const char* e = std::getenv("COVER_B200_NO_LATTICE_KERNEL")) ctx->disable_lattice_kernel = e[0] == '1';
  if (const char* e = std::getenv("COVER_B200_NO_PLANES")) ctx->disable_planes = e[0] == '1';
  if (const char* e = std::getenv("COVER_B200_NO_WORDS")) ctx->disable_words = e[0] == '1';
  if (const char* e = std::getenv("COVER_B200_WORDS_STRIP")) ctx->words_strip = std::atoi(e);
  if (const char* e = std::getenv("COVER_B200_TODO_CAP")) ctx->todo_cap_limit = std::atoll(e);
  if (const char* e = std::getenv("COVER_B200_NO_TABLE_CACHE")) ctx->lw_cache_disabled = e[0] == '1';
  if (const char* e = std::getenv("COVER_B200_PLANES_MIN_BATCH")) ctx->planes_min_batch = std::atoll(e);
  if (const char* e = std::getenv("COVER_B200_MAX_STRETCH")) ctx->max_stretch = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("COVER_B200_PIPELINE_MIN")) ctx->pipeline_min = std::max<long long>(0, std::atoll(e));
  if (const char* e = std::getenv("COVER_B200_CHUNKS")) ctx->pipeline_chunks = std::max(0, std::atoi(e));
  if (const char* e = std::getenv("COVER_B200_PACKED_ONLY_CONVEX")) ctx->always_try_packed = e[0] != '1';
  if (const char* e = std::getenv("COVER_B200_NO_QUAD")) ctx->disable_quad = e[0] == '1';
  if (const char* e = std::getenv("COVER_B200_NO_VALUES")) ctx->disable_values = e[0] == '1';
  if (const char* e = std::getenv("COVER_B200_QUAD_MAX_ROWS")) ctx->quad_max_rows = std::min(250, std::max(1, std::atoi(e)));
  if (const char* e = std::getenv("COVER_B200_NO_SPLIT_LOCKSTEP")) ctx->disable_split_lockstep = e[0] == '1';
  if (const char* e = std::getenv("COVER_B200_TIMELINE")) ctx->timeline = e[0] == '1';
  if (const char* e = std::getenv("COVER_B200_PLANE_COLS")) ctx->plane_cols = std::atoi(e);
  if (const char* e = std::getenv("COVER_B200_OLD_PACKED")) ctx->old_packed_kernel = e[0] == '1';
  if (const char* e = std::getenv("COVER_B200_FORCE_LIST_KERNEL")) ctx->force_list_kernel = e[0] == '1';
  *out = ctx;
  return COVER_OK;
}

void cover_ctx_destroy(cover_ctx* ctx) {
  if (!ctx) return;
  Bind bind(ctx->device);
  cudaStreamSynchronize(ctx->copy_stream);
  cudaStreamSynchronize(ctx->stream);
  for (DeviceBuffer* b : {&ctx->poses, &ctx->theta, &ctx->poly, &ctx->in_a, &ctx->in_b, &ctx->out_free, &ctx->out_cost, &ctx->out_status, &ctx->out_bits, &ctx->out_not_ok,
                          &ctx->todo, &ctx->todo_count, &ctx->todo2, &ctx->todo2_count, &ctx->list_cnt, &ctx->list_rng, &ctx->list_seq, &ctx->lat_global, &ctx->lw_tables})
    b->release();
  cudaEventDestroy(ctx->fence);
  cudaEventDestroy(ctx->chunk_done);
  cudaEventDestroy(ctx->aux_fork);
  cudaEventDestroy(ctx->aux_join);
  cudaEventDestroy(ctx->aux_planes);
  cudaEventDestroy(ctx->aux_done);
  cudaEventDestroy(ctx->aux_zero);
  cudaStreamSynchronize(ctx->aux_stream);
  cudaStreamDestroy(ctx->aux_stream);
  cudaStreamDestroy(ctx->d2h_stream);
  cudaStreamDestroy(ctx->copy_stream);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* cover_last_error(const cover_ctx* ctx) { return ctx ? ctx->error.c_str() : "null context"; }
void* cover_ctx_stream(cover_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }
int64_t cover_ctx_launch_count(const cover_ctx* ctx) { return ctx ? ctx->launches : 0; }

int cover_ctx_synchronize(cover_ctx* ctx) {
  if (!ctx) return COVER_E_INVALID_ARG;
  Bind bind(ctx->device);
  CU(cudaStreamSynchronize(ctx->copy_stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaStreamSynchronize(ctx->d2h_stream));
  CU(cudaStreamSynchronize(ctx->aux_stream));
  if (ctx->timeline) {
    for (size_t k = 1; k < ctx->timeline_events.size(); ++k) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->timeline_events[k - 1].second, ctx->timeline_events[k].second);
      std::fprintf(stderr, "[timeline] %-28s -> %-28s %8.2f us\n", ctx->timeline_events[k - 1].first, ctx->timeline_events[k].first, ms * 1e3f);
    }
    for (auto& e : ctx->timeline_events) cudaEventDestroy(e.second);
    ctx->timeline_events.clear();
  }
  return COVER_OK;
}

int cover_ctx_bind(cover_ctx* ctx) {
  if (!ctx) return COVER_E_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  return COVER_OK;
}

int cover_map_create(cover_ctx* ctx, int32_t size_x, int32_t size_y, double resolution, double origin_x, double origin_y,
                     cover_map** out) {
  if (!ctx || !out) return COVER_E_INVALID_ARG;
  *out = nullptr;
  if (!(resolution > 0)) return fail(ctx, COVER_E_INVALID_ARG, "The resolution must be positive");
  if (size_x <= 0 || size_y <= 0) return fail(ctx, COVER_E_INVALID_ARG, "map size must be positive");
  Bind bind(ctx->device);
  const int pitch = (size_x + 15) & ~15;
  uint8_t* dev = nullptr;
  // + 128 bytes: k_area_lattice's unchecked loop reads (and masks off) up to 95 bytes past the last needed cell
  CU(cudaMalloc(reinterpret_cast<void**>(&dev), static_cast<size_t>(pitch) * size_y + 128));
  CU(cudaMemsetAsync(dev, 0, static_cast<size_t>(pitch) * size_y + 128, ctx->stream));
  cover_map* m = new (std::nothrow) cover_map();
  if (!m) {
    cudaFree(dev);
    return fail(ctx, COVER_E_NOMEM, "out of host memory");
  }
  m->ctx = ctx;
  m->dev = dev;
  m->resolution = resolution;
  m->band_rows = (size_y + kMapBands - 1) / kMapBands;
  m->ypitch = (size_y + 31) & ~31;
  m->ppitch = (size_x + 1 + 31) & ~31;
  m->plane_words = static_cast<long long>((size_x + 31) / 32 + 3) * m->ypitch;
  for (cudaEvent_t& e : m->band_done) {
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
      for (cudaEvent_t& d : m->band_done)
        if (d) cudaEventDestroy(d);
      cudaFree(dev);
      delete m;
      return fail(ctx, COVER_E_CUDA, "cudaEventCreate failed");
    }
  }
  m->view = MapView{dev, size_x, size_y, pitch, 1. / resolution, origin_x, origin_y, nullptr, nullptr, 0, 0};  // base.hpp:56: x * (1./res)
  *out = m;
  return COVER_OK;
}

int cover_map_upload_async(cover_map* map, const uint8_t* host_data) {
  if (!map || !host_data) return COVER_E_INVALID_ARG;
  cover_ctx* ctx = map->ctx;
  Bind bind(ctx->device);
  const MapView& v = map->view;
  // the copy may not overtake kernels already queued that still read (or an earlier call that writes) the map
  CU(cudaEventRecord(ctx->fence, ctx->stream));
  CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->fence, 0));
  // Copy order: the bands the last lattice sweep needed first, as ONE copy (a planner sweeps the same window every
  // cycle), then the rows below and above them; without a hint four copies of eight bands.  Few calls matter: the
  // host issues the sweep only after it has queued the whole upload.
  int first[4], last[4], groups = 0;  // band ranges [first, last] in copy order
  const bool hinted = map->hint_lo >= 0 && map->hint_hi < kMapBands && map->hint_lo <= map->hint_hi;
  if (hinted) {
    first[groups] = map->hint_lo, last[groups++] = map->hint_hi;
    if (map->hint_lo > 0) first[groups] = 0, last[groups++] = map->hint_lo - 1;
    if (map->hint_hi < kMapBands - 1) first[groups] = map->hint_hi + 1, last[groups++] = kMapBands - 1;
  } else {
    for (int q = 0; q < 4; ++q) first[groups] = q * (kMapBands / 4), last[groups++] = (q + 1) * (kMapBands / 4) - 1;
  }
  for (int k = 0; k < groups; ++k) {
    const int y0 = static_cast<int>(std::min<long long>(v.size_y, static_cast<long long>(first[k]) * map->band_rows));
    const int y1 = static_cast<int>(std::min<long long>(v.size_y, static_cast<long long>(last[k] + 1) * map->band_rows));
    if (y1 > y0) CU(copy_rows(map->dev + static_cast<size_t>(y0) * v.pitch, v.pitch, host_data + static_cast<size_t>(y0) * v.size_x, v.size_x, y1 - y0,
                              cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaEventRecord(map->band_done[k], ctx->copy_stream));
    for (int bb = first[k]; bb <= last[k]; ++bb) map->issue_pos[bb] = k;  // (= index of the band's event)
  }
  map->last_band = last[groups - 1];
  map->bands_pending = true;
  map->planes_hi = map->planes_inf_hi = map->prefix_hi = map->maxp_hi = -1;  // the bit planes describe the old image
  return COVER_OK;
}

int cover_map_upload(cover_map* map, const uint8_t* data, int32_t mem) {
  if (!map || !data) return COVER_E_INVALID_ARG;
  cover_ctx* ctx = map->ctx;
  Bind bind(ctx->device);
  const MapView& v = map->view;
  CU(map_ready(ctx, map));
  map->planes_hi = map->planes_inf_hi = map->prefix_hi = map->maxp_hi = -1;
  tl(ctx, "upload: entry");
  CU(copy_rows(map->dev, v.pitch, data, v.size_x, v.size_y, mem == COVER_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream,
               &ctx->launches));
  tl(ctx, "upload: copied");
  if (mem == COVER_MEM_HOST) CU(cudaStreamSynchronize(ctx->stream));
  return COVER_OK;
}

int cover_map_upload_window(cover_map* map, const uint8_t* full_image, int32_t x0, int32_t y0, int32_t w, int32_t h) {
  if (!map || !full_image) return COVER_E_INVALID_ARG;
  cover_ctx* ctx = map->ctx;
  const MapView& v = map->view;
  if (x0 < 0 || y0 < 0 || w < 0 || h < 0 || x0 + w > v.size_x || y0 + h > v.size_y) return fail(ctx, COVER_E_INVALID_ARG, "window outside the map");
  if (w == 0 || h == 0) return COVER_OK;
  Bind bind(ctx->device);
  CU(map_ready(ctx, map));
  CU(cudaMemcpy2DAsync(map->dev + static_cast<size_t>(y0) * v.pitch + x0, v.pitch, full_image + static_cast<size_t>(y0) * v.size_x + x0,
                       v.size_x, w, h, cudaMemcpyHostToDevice, ctx->stream));
  // The acceleration structures are per map row: repair the rows of the window that they currently cover instead of
  // throwing everything away (a planner's costmap usually changes locally).
  const int r_lo = y0, r_hi = y0 + h - 1;
  auto repair = [&](uint32_t* planes, int lo, int hi, bool inflated) {
    const int a = std::max(lo, r_lo), b = std::min(hi, r_hi);
    if (!planes || a > b) return;
    launch_planes(ctx, v, planes, map->ypitch, map->plane_words, a, b, inflated, CallReset{nullptr, 0u, nullptr});
    ++ctx->launches;
  };
  repair(map->planes, map->planes_lo, map->planes_hi, false);
  repair(map->planes_inf, map->planes_inf_lo, map->planes_inf_hi, true);
  {
    const int a = std::max(map->prefix_lo, r_lo), b = std::min(map->prefix_hi, r_hi);
    if (map->prefix && a <= b) {
      k_cost_prefix<<<(b - a + 1 + 7) / 8, 256, 0, ctx->stream>>>(v, map->prefix, map->ppitch, a, b);
      ++ctx->launches;
    }
  }
  {
    const int a = std::max(map->maxp_lo, r_lo), b = std::min(map->maxp_hi, r_hi);
    if (map->maxp && a <= b) launch_max_planes(ctx, map, a, b);
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return COVER_OK;
}

int cover_probe_read_bandwidth(cover_ctx* ctx, int64_t bytes, int32_t reps, double* gb_per_s) {
  if (!ctx || !gb_per_s || bytes < 4096 || reps < 1) return COVER_E_INVALID_ARG;
  Bind bind(ctx->device);
  const long long n_vec = bytes / 16;
  CU(ctx->list_rng.ensure(static_cast<size_t>(n_vec) * 16 + 16));
  CU(cudaMemsetAsync(ctx->list_rng.ptr, 1, static_cast<size_t>(n_vec) * 16, ctx->stream));
  unsigned int* sink = reinterpret_cast<unsigned int*>(static_cast<char*>(ctx->list_rng.ptr) + n_vec * 16);
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  const unsigned int grid = static_cast<unsigned int>(ctx->sm_count * 8);
  k_probe_read<<<grid, 256, 0, ctx->stream>>>(ctx->list_rng.as<uint4>(), n_vec, 2, sink);  // warm-up (fills the L2)
  double best = 0;
  for (int t = 0; t < 10; ++t) {  // best of 10
    cudaEventRecord(e0, ctx->stream);
    k_probe_read<<<grid, 256, 0, ctx->stream>>>(ctx->list_rng.as<uint4>(), n_vec, reps, sink);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms > 0) best = std::max(best, static_cast<double>(n_vec) * 16 * reps / (ms * 1e-3) / 1e9);
  }
  ctx->launches += 11;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CU(cudaGetLastError());
  *gb_per_s = best;
  return COVER_OK;
}

int cover_map_download(const cover_map* map, uint8_t* data) {
  if (!map || !data) return COVER_E_INVALID_ARG;
  cover_ctx* ctx = map->ctx;
  Bind bind(ctx->device);
  const MapView& v = map->view;
  CU(map_ready(ctx, map));
  CU(cudaMemcpy2DAsync(data, v.size_x, map->dev, v.pitch, v.size_x, v.size_y, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return COVER_OK;
}

int cover_map_inflate(cover_map* map, int32_t radius_cells, const uint8_t* cost_by_d2) {
  if (!map || !cost_by_d2) return COVER_E_INVALID_ARG;
  cover_ctx* ctx = map->ctx;
  if (radius_cells < 0 || radius_cells > 96) return fail(ctx, COVER_E_INVALID_ARG, "inflation radius must be 0..96 cells");
  Bind bind(ctx->device);
  const MapView& v = map->view;
  const size_t bytes = static_cast<size_t>(v.pitch) * v.size_y;
  CU(map_ready(ctx, map));
  map->planes_hi = map->planes_inf_hi = map->prefix_hi = map->maxp_hi = -1;
  const size_t lut_bytes = static_cast<size_t>(radius_cells) * radius_cells + 1;
  CU(ctx->in_b.ensure(lut_bytes));
  CU(cudaMemcpyAsync(ctx->in_b.ptr, cost_by_d2, lut_bytes, cudaMemcpyHostToDevice, ctx->stream));
  const dim3 grid((v.size_x + 31) / 32, (v.size_y + 31) / 32);
  bool makes_lethal = false;  // a table that turns neighbours LETHAL would change what other tiles read: snapshot path
  for (size_t k = 1; k < lut_bytes; ++k) makes_lethal = makes_lethal || cost_by_d2[k] == kLethal;
  if (radius_cells <= kInflateFastRadius && !makes_lethal) {
    k_inflate_separable<<<grid, 256, 0, ctx->stream>>>(map->dev, v.size_x, v.size_y, v.pitch, radius_cells, ctx->in_b.as<uint8_t>());
  } else {
    CU(ctx->in_a.ensure(bytes));  // snapshot of the un-inflated map: the kernel reads it and writes the map in place
    CU(cudaMemcpyAsync(ctx->in_a.ptr, map->dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    const int span = 32 + 2 * radius_cells;
    k_inflate<<<grid, 256, static_cast<size_t>(span) * span, ctx->stream>>>(ctx->in_a.as<uint8_t>(), map->dev, v.size_x, v.size_y, v.pitch,
                                                                           radius_cells, ctx->in_b.as<uint8_t>());
  }
  ++ctx->launches;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return COVER_OK;
}

void cover_map_destroy(cover_map* map) {
  if (!map) return;
  Bind bind(map->ctx->device);
  cudaStreamSynchronize(map->ctx->copy_stream);
  cudaStreamSynchronize(map->ctx->stream);
  for (cudaEvent_t e : map->band_done) cudaEventDestroy(e);
  cudaFree(map->planes);
  cudaFree(map->planes_inf);
  cudaFree(map->prefix);
  cudaFree(map->maxp);
  cudaFree(map->dev);
  delete map;
}

int cover_area_is_free(cover_ctx* ctx, const cover_map* map, const double* xy, int32_t nv, const cover_poses* poses, const cover_results* out) {
  return polygon_check(ctx, map, xy, nv, SHAPE_AREA, OP_IS_FREE, poses, out);
}
int cover_area_accumulate_cost(cover_ctx* ctx, const cover_map* map, const double* xy, int32_t nv, const cover_poses* poses, const cover_results* out) {
  return polygon_check(ctx, map, xy, nv, SHAPE_AREA, OP_ACCUMULATE, poses, out);
}
int cover_area_max_cost(cover_ctx* ctx, const cover_map* map, const double* xy, int32_t nv, const cover_poses* poses, const cover_results* out) {
  return polygon_check(ctx, map, xy, nv, SHAPE_AREA, OP_MAX, poses, out);
}
int cover_area_multi_check(cover_ctx* ctx, const cover_map* map, const double* polygons_xy, const int32_t* starts, int32_t n_outlines, int32_t op,
                           const cover_poses* poses, const cover_results* out) {
  if (!ctx) return COVER_E_INVALID_ARG;
  if (!map || map->ctx != ctx) return fail(ctx, COVER_E_INVALID_ARG, "map does not belong to this context");
  if (n_outlines < 0 || (n_outlines && !starts) || op < OP_IS_FREE || op > OP_MAX) return fail(ctx, COVER_E_INVALID_ARG, "bad arguments");
  if (n_outlines > kRowsMaxOutlines) return fail(ctx, COVER_E_UNSUPPORTED, "at most 16 outlines");
  int nv_total = 0, nv_max = 0;
  for (int o = 0; o < n_outlines; ++o) {
    const int n = starts[o + 1] - starts[o];
    if (n < 0 || starts[0] != 0) return fail(ctx, COVER_E_INVALID_ARG, "starts must begin at 0 and be non-decreasing");
    nv_total += n;
    nv_max = std::max(nv_max, n);
  }
  if (nv_total > kRowsMaxVertices) return fail(ctx, COVER_E_UNSUPPORTED, "at most 256 vertices in total");
  if (nv_total && !polygons_xy) return fail(ctx, COVER_E_INVALID_ARG, "missing vertices");
  Bind bind(ctx->device);
  PoseSource src;
  int rc = stage_poses(ctx, poses, src);
  if (rc) return rc;
  OutStage st;
  rc = stage_out(ctx, out, poses->count, st);
  if (rc) return rc;
  const size_t poly_bytes = static_cast<size_t>(nv_total) * 2 * sizeof(double);
  CU(ctx->poly.ensure(std::max<size_t>(poly_bytes, 16)));
  if (poly_bytes) CU(cudaMemcpyAsync(ctx->poly.ptr, polygons_xy, poly_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx->in_b.ensure((kRowsMaxOutlines + 1) * sizeof(int)));
  std::vector<int> st_host(starts, starts + n_outlines + 1);
  if (n_outlines == 0) st_host.assign(1, 0);
  CU(cudaMemcpyAsync(ctx->in_b.ptr, st_host.data(), st_host.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  const int64_t count = poses->count;
  if (count) {
    const size_t smem = poly_bytes + kRowsWarps * (static_cast<size_t>(nv_total + n_outlines) * sizeof(RayRec) + static_cast<size_t>(nv_max) * sizeof(int2) +
                                                   kRowsMaxOutlines * sizeof(OutlineRec));
    const unsigned int grid = static_cast<unsigned int>(std::max<long long>(1, std::min<long long>(static_cast<long long>(ctx->sm_count) * 12, (count + kRowsWarps - 1) / kRowsWarps)));
    const double* pd = ctx->poly.as<double>();
    const int* sd = ctx->in_b.as<int>();
    CU(map_ready(ctx, map));
    if (op == OP_IS_FREE) k_area_rows_multi<OP_IS_FREE><<<grid, kRowsWarps * 32, smem, ctx->stream>>>(map->view, pd, sd, n_outlines, nv_total, nv_max, src, count, st.dev);
    else if (op == OP_ACCUMULATE) k_area_rows_multi<OP_ACCUMULATE><<<grid, kRowsWarps * 32, smem, ctx->stream>>>(map->view, pd, sd, n_outlines, nv_total, nv_max, src, count, st.dev);
    else k_area_rows_multi<OP_MAX><<<grid, kRowsWarps * 32, smem, ctx->stream>>>(map->view, pd, sd, n_outlines, nv_total, nv_max, src, count, st.dev);
    ++ctx->launches;
  }
  return finish_out(ctx, st, count);
}

int cover_outline_is_free(cover_ctx* ctx, const cover_map* map, const double* xy, int32_t nv, const cover_poses* poses, const cover_results* out) {
  return polygon_check(ctx, map, xy, nv, SHAPE_OUTLINE, OP_IS_FREE, poses, out);
}
int cover_outline_accumulate_cost(cover_ctx* ctx, const cover_map* map, const double* xy, int32_t nv, const cover_poses* poses, const cover_results* out) {
  return polygon_check(ctx, map, xy, nv, SHAPE_OUTLINE, OP_ACCUMULATE, poses, out);
}
int cover_outline_max_cost(cover_ctx* ctx, const cover_map* map, const double* xy, int32_t nv, const cover_poses* poses, const cover_results* out) {
  return polygon_check(ctx, map, xy, nv, SHAPE_OUTLINE, OP_MAX, poses, out);
}

int cover_ray_is_free(cover_ctx* ctx, const cover_map* map, const int32_t* seg, int32_t mem, int64_t count, const int32_t* offset, const cover_results* out) {
  return rays_check(ctx, map, seg, mem, count, offset, OP_IS_FREE, out);
}
int cover_ray_accumulate_cost(cover_ctx* ctx, const cover_map* map, const int32_t* seg, int32_t mem, int64_t count, const int32_t* offset, const cover_results* out) {
  return rays_check(ctx, map, seg, mem, count, offset, OP_ACCUMULATE, out);
}
int cover_ray_max_cost(cover_ctx* ctx, const cover_map* map, const int32_t* seg, int32_t mem, int64_t count, const int32_t* offset, const cover_results* out) {
  return rays_check(ctx, map, seg, mem, count, offset, OP_MAX, out);
}

int cover_area_expand(cover_ctx* ctx, const cover_map* map, const double* xy, int32_t nv, const cover_poses* poses, int32_t* cells,
                      int64_t capacity, int64_t* starts, uint8_t* status, int32_t mem) {
  return expand_cells(ctx, map, xy, nv, true, poses, cells, capacity, starts, status, mem);
}
int cover_outline_expand(cover_ctx* ctx, const cover_map* map, const double* xy, int32_t nv, const cover_poses* poses, int32_t* cells,
                         int64_t capacity, int64_t* starts, uint8_t* status, int32_t mem) {
  return expand_cells(ctx, map, xy, nv, false, poses, cells, capacity, starts, status, mem);
}

double cover_get_radius(const double* polygon_xy, int32_t n_vertices) {
  if (n_vertices <= 0 || !polygon_xy) return 0;
  return cvr_split::get_radius(polygon_xy, n_vertices);
}

int cover_split_polygon(double radius, const double* polygon_xy, int32_t n_vertices, double* outline_xy, int32_t* outline_starts,
                        int32_t* n_outlines, double* area_xy, int32_t* area_starts, int32_t* n_areas, int32_t max_vertices,
                        int32_t max_polygons) {
  return cover_split_polygon_slack(radius, 0.0, polygon_xy, n_vertices, outline_xy, outline_starts, n_outlines, area_xy, area_starts, n_areas,
                                   max_vertices, max_polygons);
}

int cover_split_polygon_slack(double radius, double slack, const double* polygon_xy, int32_t n_vertices, double* outline_xy,
                              int32_t* outline_starts, int32_t* n_outlines, double* area_xy, int32_t* area_starts, int32_t* n_areas,
                              int32_t max_vertices, int32_t max_polygons) {
  if (n_vertices < 0 || (n_vertices > 0 && !polygon_xy) || !outline_starts || !n_outlines || !area_starts || !n_areas) return COVER_E_INVALID_ARG;
  cvr_split::Split sp;
  try {
    sp = cvr_split::split(radius, polygon_xy, n_vertices, slack);
  } catch (const std::invalid_argument&) {
    return COVER_E_INVALID_ARG;
  } catch (const std::bad_alloc&) {
    return COVER_E_NOMEM;
  }
  auto pack = [&](const std::vector<cvr_split::Ring>& rings, double* xy, int32_t* starts, int32_t* count) {
    size_t total = 0;
    for (const cvr_split::Ring& r : rings) total += r.size();
    if (rings.size() > static_cast<size_t>(max_polygons) || total > static_cast<size_t>(max_vertices) || (total && !xy)) return false;
    starts[0] = 0;
    int32_t at = 0;
    for (size_t k = 0; k < rings.size(); ++k) {
      for (const cvr_split::Pt& p : rings[k]) xy[2 * at] = p.x, xy[2 * at + 1] = p.y, ++at;
      starts[k + 1] = at;
    }
    *count = static_cast<int32_t>(rings.size());
    return true;
  };
  if (!pack(sp.outlines, outline_xy, outline_starts, n_outlines) || !pack(sp.areas, area_xy, area_starts, n_areas)) return COVER_E_NOMEM;
  return COVER_OK;
}

int cover_trajectory_reduce(cover_ctx* ctx, const uint8_t* is_free, const double* cost, int32_t mem, int64_t n_trajectories, int32_t steps,
                            int32_t* first_blocked, double* score) {
  if (!ctx) return COVER_E_INVALID_ARG;
  if (n_trajectories < 0 || steps <= 0 || (!is_free && !cost) || (first_blocked && !is_free) || (score && !cost))
    return fail(ctx, COVER_E_INVALID_ARG, "bad trajectory arguments");
  if (n_trajectories == 0) return COVER_OK;
  Bind bind(ctx->device);
  const size_t n = static_cast<size_t>(n_trajectories) * static_cast<size_t>(steps);
  const uint8_t* d_free = is_free;
  const double* d_cost = cost;
  int32_t* d_first = first_blocked;
  double* d_score = score;
  if (mem == COVER_MEM_HOST) {
    if (is_free) {
      CU(ctx->out_free.ensure(n));
      CU(cudaMemcpyAsync(ctx->out_free.ptr, is_free, n, cudaMemcpyHostToDevice, ctx->stream));
      d_free = ctx->out_free.as<uint8_t>();
    }
    if (cost) {
      CU(ctx->out_cost.ensure(n * sizeof(double)));
      CU(cudaMemcpyAsync(ctx->out_cost.ptr, cost, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      d_cost = ctx->out_cost.as<double>();
    }
    if (first_blocked) {
      CU(ctx->in_a.ensure(static_cast<size_t>(n_trajectories) * sizeof(int32_t)));
      d_first = ctx->in_a.as<int32_t>();
    }
    if (score) {
      CU(ctx->in_b.ensure(static_cast<size_t>(n_trajectories) * sizeof(double)));
      d_score = ctx->in_b.as<double>();
    }
  }
  const unsigned int grid = static_cast<unsigned int>((n_trajectories * 32 + kBlock - 1) / kBlock);
  k_trajectories<<<grid, kBlock, 0, ctx->stream>>>(d_free, d_cost, n_trajectories, steps, d_first, d_score);
  ++ctx->launches;
  CU(cudaGetLastError());
  if (mem == COVER_MEM_HOST) {
    if (first_blocked) CU(cudaMemcpyAsync(first_blocked, d_first, static_cast<size_t>(n_trajectories) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (score) CU(cudaMemcpyAsync(score, d_score, static_cast<size_t>(n_trajectories) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return COVER_OK;
}

int cover_cells_create(cover_ctx* ctx, const int32_t* cells, int64_t n_cells, cover_cells** out) {
  if (!ctx || !out) return COVER_E_INVALID_ARG;
  *out = nullptr;
  if (n_cells < 0 || (n_cells && !cells)) return fail(ctx, COVER_E_INVALID_ARG, "bad cell list");
  Bind bind(ctx->device);
  int2* dev = nullptr;
  int rc = upload_new(ctx, reinterpret_cast<const int2*>(cells), static_cast<size_t>(n_cells), &dev);
  if (rc) return rc;
  std::vector<int4> spans;
  append_row_spans(spans, cells, 0, n_cells);
  int4* dev_spans = nullptr;
  rc = upload_new(ctx, spans.data(), spans.size(), &dev_spans);
  if (rc) {
    cudaFree(dev);
    return rc;
  }
  *out = new cover_cells{ctx, dev, n_cells, dev_spans, static_cast<int>(spans.size())};
  return COVER_OK;
}

void cover_cells_destroy(cover_cells* cells) {
  if (!cells) return;
  Bind bind(cells->ctx->device);
  cudaStreamSynchronize(cells->ctx->stream);
  cudaFree(cells->dev);
  cudaFree(cells->spans);
  delete cells;
}

int cover_cells_is_free(cover_ctx* ctx, const cover_map* map, const cover_cells* cells, const int32_t* offsets, int32_t mem, int64_t count, const cover_results* out) {
  return cells_check(ctx, map, cells, offsets, mem, count, OP_IS_FREE, out);
}
int cover_cells_accumulate_cost(cover_ctx* ctx, const cover_map* map, const cover_cells* cells, const int32_t* offsets, int32_t mem, int64_t count, const cover_results* out) {
  return cells_check(ctx, map, cells, offsets, mem, count, OP_ACCUMULATE, out);
}
int cover_cells_max_cost(cover_ctx* ctx, const cover_map* map, const cover_cells* cells, const int32_t* offsets, int32_t mem, int64_t count, const cover_results* out) {
  return cells_check(ctx, map, cells, offsets, mem, count, OP_MAX, out);
}

int cover_footprints_create(cover_ctx* ctx, const int32_t* outline_cells, const int64_t* o_starts, const int32_t* area_cells,
                            const int64_t* a_starts, int32_t n, cover_footprints** out) {
  if (!ctx || !out) return COVER_E_INVALID_ARG;
  *out = nullptr;
  if (n <= 0 || !o_starts || !a_starts) return fail(ctx, COVER_E_INVALID_ARG, "bad footprint bank");
  if (o_starts[0] != 0 || a_starts[0] != 0) return fail(ctx, COVER_E_INVALID_ARG, "starts must begin at 0");
  for (int k = 0; k < n; ++k)
    if (o_starts[k + 1] < o_starts[k] || a_starts[k + 1] < a_starts[k]) return fail(ctx, COVER_E_INVALID_ARG, "starts must be non-decreasing");
  if ((o_starts[n] && !outline_cells) || (a_starts[n] && !area_cells)) return fail(ctx, COVER_E_INVALID_ARG, "missing cells");
  // cell set -> sorted, merged row spans (exact for the any-reduction; duplicates and order do not matter)
  std::vector<int4> spans;
  std::vector<int2> ranges;
  auto append_spans = [&spans](const int32_t* cells, int64_t b, int64_t e) { append_row_spans(spans, cells, b, e); };
  for (int k = 0; k < n; ++k) {
    const int ob = static_cast<int>(spans.size());
    append_spans(outline_cells, o_starts[k], o_starts[k + 1]);
    const int ab = static_cast<int>(spans.size());
    append_spans(area_cells, a_starts[k], a_starts[k + 1]);
    ranges.push_back(make_int2(ob, ab));
    ranges.push_back(make_int2(ab, static_cast<int>(spans.size())));
  }
  Bind bind(ctx->device);
  cover_footprints* f = new cover_footprints{ctx, nullptr, nullptr, n};
  int rc = upload_new(ctx, spans.data(), spans.size(), &f->spans);
  if (!rc) rc = upload_new(ctx, ranges.data(), ranges.size(), &f->ranges);
  if (rc) {
    cover_footprints_destroy(f);
    return rc;
  }
  *out = f;
  return COVER_OK;
}

void cover_footprints_destroy(cover_footprints* f) {
  if (!f) return;
  Bind bind(f->ctx->device);
  cudaStreamSynchronize(f->ctx->stream);
  cudaFree(f->spans);
  cudaFree(f->ranges);
  delete f;
}

int cover_footprints_is_free(cover_ctx* ctx, const cover_map* map, const cover_footprints* fps, const int32_t* offsets,
                             const int32_t* which, int32_t mem, int64_t count, const cover_results* out) {
  if (!ctx) return COVER_E_INVALID_ARG;
  if (!map || map->ctx != ctx || !fps || fps->ctx != ctx) return fail(ctx, COVER_E_INVALID_ARG, "handle does not belong to this context");
  if (count < 0 || (count && !offsets)) return fail(ctx, COVER_E_INVALID_ARG, "bad offsets");
  if (mem == COVER_MEM_HOST && which)
    for (int64_t i = 0; i < count; ++i)
      if (which[i] < 0 || which[i] >= fps->n) return fail(ctx, COVER_E_INVALID_ARG, "footprint index out of range");
  Bind bind(ctx->device);
  OutStage st;
  int rc = stage_out(ctx, out, count, st);
  if (rc) return rc;
  const int32_t *off_dev = nullptr, *which_dev = nullptr;
  rc = stage_array(ctx, ctx->in_a, offsets, static_cast<size_t>(count) * 2, mem, &off_dev);
  if (rc) return rc;
  if (which) {
    rc = stage_array(ctx, ctx->in_b, which, static_cast<size_t>(count), mem, &which_dev);
    if (rc) return rc;
  }
  if (count) {
    CU(map_ready(ctx, map));
    MapView mview = map->view;
    if (planes_pay_off(ctx, map, true, 0, mview.size_y - 1, count)) {  // outline spans: 253 | 254 planes, area spans: 254 planes
      rc = view_with_planes(ctx, map, true, 0, mview.size_y - 1, mview);
      if (rc) return rc;
    }
    k_footprints<<<grid_for(count), kBlock, 0, ctx->stream>>>(mview, fps->spans, fps->ranges, fps->n, reinterpret_cast<const int2*>(off_dev),
                                                             which_dev, count, st.dev);
    ++ctx->launches;
  }
  return finish_out(ctx, st, count);
}

int cover_split_create(cover_ctx* ctx, const double* outline_xy, const int32_t* outline_starts, int32_t n_outlines,
                       const double* area_xy, const int32_t* area_starts, int32_t n_areas, cover_split** out) {
  if (!ctx || !out) return COVER_E_INVALID_ARG;
  *out = nullptr;
  if (n_outlines < 0 || n_areas < 0 || (n_outlines && !outline_starts) || (n_areas && !area_starts)) return fail(ctx, COVER_E_INVALID_ARG, "bad split");
  std::vector<double> xy;
  std::vector<int> starts{0};
  double diam = 0;
  for (int k = 0; k < n_outlines; ++k) {
    const int b = outline_starts[k], e = outline_starts[k + 1];
    if (e < b) return fail(ctx, COVER_E_INVALID_ARG, "starts must be non-decreasing");
    xy.insert(xy.end(), outline_xy + 2 * b, outline_xy + 2 * e);
    starts.push_back(starts.back() + (e - b));
  }
  for (int k = 0; k < n_areas; ++k) {
    const int b = area_starts[k], e = area_starts[k + 1];
    if (e < b) return fail(ctx, COVER_E_INVALID_ARG, "starts must be non-decreasing");
    diam = std::max(diam, polygon_diameter(area_xy + 2 * b, e - b));
    xy.insert(xy.end(), area_xy + 2 * b, area_xy + 2 * e);
    starts.push_back(starts.back() + (e - b));
  }
  if (starts.back() > kMaxVertices) return fail(ctx, COVER_E_UNSUPPORTED, "split footprints with more than 4096 vertices are not supported");
  Bind bind(ctx->device);
  cover_split* s = new cover_split{ctx, nullptr, nullptr, n_outlines, n_areas, starts.back(), diam};
  int rc = upload_new(ctx, xy.data(), xy.size(), &s->xy);
  if (!rc) rc = upload_new(ctx, starts.data(), starts.size(), &s->starts);
  if (rc) {
    cover_split_destroy(s);
    return rc;
  }
  *out = s;
  return COVER_OK;
}

void cover_split_destroy(cover_split* s) {
  if (!s) return;
  Bind bind(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  cudaFree(s->xy);
  cudaFree(s->starts);
  delete s;
}

int cover_split_is_free(cover_ctx* ctx, const cover_map* map, const cover_split* split, const cover_poses* poses, const cover_results* out) {
  if (!ctx) return COVER_E_INVALID_ARG;
  if (!map || map->ctx != ctx || !split || split->ctx != ctx) return fail(ctx, COVER_E_INVALID_ARG, "handle does not belong to this context");
  if (poses && poses->format == COVER_POSE_NONE) return fail(ctx, COVER_E_INVALID_ARG, "continuous_footprint::is_free needs a pose");
  Bind bind(ctx->device);
  PoseSource src;
  int rc = stage_poses(ctx, poses, src);
  if (rc) return rc;
  OutStage st;
  rc = stage_out(ctx, out, poses->count, st);
  if (rc) return rc;
  if (poses->count) {
    const int rows_cap = rows_bound(split->max_area_diameter, map->view.inv_res);
    const unsigned int grid = list_grid(ctx, poses->count, rows_cap, 4);
    ListScratch scratch{};
    rc = ensure_list_scratch(ctx, rows_cap, static_cast<long long>(grid) * kBlock, scratch);
    if (rc) return rc;
    const SplitView view{split->xy, split->starts, split->n_outlines, split->n_areas};
    const int slot_rows = rows_cap <= 48 ? rows_cap : 0;  // shared-memory 4-range table: rows x 4 x 128 threads x 4 B
    const size_t smem = static_cast<size_t>(split->n_vertices) * 2 * sizeof(double) + static_cast<size_t>(slot_rows) * 4 * kBlock * sizeof(uint32_t);
    CU(map_ready(ctx, map));
    MapView mview = map->view;
    if (planes_pay_off(ctx, map, false, 0, mview.size_y - 1, poses->count)) {  // the area parts scan row spans
      rc = view_with_planes(ctx, map, false, 0, mview.size_y - 1, mview);
      if (rc) return rc;
    }
    // first pass in lock-step (thread = pose, the warp at the same ray); what its four-ranges-per-row table cannot hold
    // is left to k_split over a todo list
    const size_t poly_bytes = static_cast<size_t>(split->n_vertices) * 2 * sizeof(double);
    const int lock_threads = (!ctx->disable_split_lockstep && rows_cap <= kQuadMaxRows) ? quad_block_threads(poly_bytes, rows_cap) : 0;
    const unsigned int* todo = nullptr;
    const unsigned int* todo_count = nullptr;
    if (lock_threads > 0) {
      int64_t todo_cap = poses->count;
      if (ctx->todo_cap_limit > 0) todo_cap = std::min(todo_cap, ctx->todo_cap_limit);
      CU(ctx->todo.ensure(static_cast<size_t>(todo_cap) * sizeof(unsigned int)));
      CU(ctx->todo_count.ensure(2 * sizeof(unsigned int)));
      ctx->todo_header[0] = 0u, ctx->todo_header[1] = static_cast<unsigned int>(todo_cap);
      CU(cudaMemcpyAsync(ctx->todo_count.ptr, ctx->todo_header, 2 * sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
      const size_t smem_l = poly_bytes + static_cast<size_t>(rows_cap) * lock_threads * sizeof(uint2);
      const unsigned int grid_l = static_cast<unsigned int>(std::max<long long>(1, std::min<long long>((poses->count + lock_threads - 1) / lock_threads,
                                                                                                     static_cast<long long>(ctx->sm_count) * 32)));
      k_split_lockstep<<<grid_l, lock_threads, smem_l, ctx->stream>>>(mview, view, split->n_vertices, src, poses->count, st.dev, rows_cap,
                                                                      ctx->todo.as<unsigned int>(), ctx->todo_count.as<unsigned int>());
      ++ctx->launches;
      todo = ctx->todo.as<unsigned int>(), todo_count = ctx->todo_count.as<unsigned int>();
      if (st.dev.not_ok) {
        k_overflow_recount<<<1, 32, 0, ctx->stream>>>(todo_count, st.dev.not_ok);
        ++ctx->launches;
      }
    }
    k_split<<<lock_threads > 0 ? std::min<unsigned int>(grid, static_cast<unsigned int>(ctx->sm_count) * 2) : grid, kBlock, smem, ctx->stream>>>(
        mview, view, split->n_vertices, src, poses->count, st.dev, scratch, slot_rows, todo, todo_count);
    ++ctx->launches;
  }
  return finish_out(ctx, st, poses->count);
}

}  // extern "C"
