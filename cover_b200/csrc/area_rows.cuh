// area_rows.cuh — warp-per-pose area checks for ARBITRARY polygons (non-convex: cusps, ties, up to 16
// ranges per scan row), the general counterpart of the packed two-range path.
//
// One warp owns one pose; inside a chunk of 32 scan rows every LANE owns one row.  Nothing is walked cell
// by cell: for each Bresenham ray of the outline a lane computes, in closed form, the stretch of the ray's
// cells that lies on its row (cell k of a ray sits at begin + k*major + floor((size/2 + k*add)/size)*minor,
// src/cover/generators.hpp:117-128), and a small state machine glues stretches of consecutive rays into
// the reference's RUNS (maximal cyclic stretches of outline cells on one row, generators.cpp:189-232):
// a run contributes the range (x of its first cell, x of its last cell), twice when y turns around at the
// run (cusp); the run containing outline cell 0 may start at the tail of the outline and is inserted with
// the outline index of its END, like every other run, which is exactly the reference's insertion order.
// Each lane keeps its row's ranges sorted by (min, outline index) — what libstdc++'s std::sort yields for
// <= 16 elements (generators.cpp:236-241) — pairs them [r[2k].min, r[2k+1].max] (generators.cpp:253-254)
// and folds the map bytes of its spans; lanes are then combined in row order with warp shuffles/votes.
// All state lives in registers, per-warp shared memory (vertices + ray records) and 192 B of per-lane
// local memory: no global scratch.
#pragma once

#include "device_core.cuh"

namespace cvr {

constexpr int kRowsMaxVertices = 256;

struct RayRec {
  int bx, by;     // first cell
  int size, add;  // Bresenham parameters (generators.cpp:38-41)
  int sx, sy;     // signs of dx, dy
  int start;      // outline index of the first cell
  int ylast;      // y of the last cell (k = size - 1)
  int xmajor;     // |dx| >= |dy| (ties: x is the major axis)
  double rcp_size, rcp_add;  // 1/size, 1/add (0 when add == 0): seeds of the exact divisions
};

/// floor(a / b) for 0 <= a < 2^31, 0 < b: one FP64 multiply by a precomputed 1/b plus an exact integer fix-up
/// (the product is within one unit of the true quotient), instead of the ~20-instruction integer division.
__device__ __forceinline__ int floor_div_pos(int a, int b, double rcp_b) {
  int q = __double2int_rd(__dmul_rn(static_cast<double>(a), rcp_b));
  const int r = a - q * b;
  q += (r >= b) - (r < 0);
  return q;
}
/// ceil(a / b) for b > 0 and any a (|a| < 2^31 - b).
__device__ __forceinline__ int ceil_div(int a, int b, double rcp_b) {
  return a > 0 ? floor_div_pos(a + b - 1, b, rcp_b) : -floor_div_pos(-a, b, rcp_b);
}

__device__ __forceinline__ RayRec make_ray_rec(int2 b, int2 e) {
  const int dx = e.x - b.x, dy = e.y - b.y, adx = abs(dx), ady = abs(dy);
  RayRec r;
  r.bx = b.x, r.by = b.y;
  r.xmajor = adx >= ady;
  r.size = r.xmajor ? adx : ady;
  r.add = r.xmajor ? ady : adx;
  r.sx = (dx > 0) - (dx < 0);
  r.sy = (dy > 0) - (dy < 0);
  r.start = 0;
  r.rcp_size = 1.0 / static_cast<double>(r.size);
  r.rcp_add = r.add ? 1.0 / static_cast<double>(r.add) : 0.0;
  const int k = r.size - 1;
  r.ylast = r.xmajor ? b.y + r.sy * floor_div_pos((r.size >> 1) + k * r.add, r.size, r.rcp_size) : b.y + r.sy * k;
  return r;
}

/// The one-cell dummy ray that keeps the last point of an open outline (generators.cpp:72-78).
__device__ __forceinline__ RayRec dummy_ray_rec(int2 p) { return RayRec{p.x, p.y, 1, 0, 0, 0, 0, p.y, 1, 1.0, 0.0}; }

/// Cells of ray r on scan row y: indices [k_lo, k_hi] and the x of both ends.  false = none.
__device__ __forceinline__ bool ray_row_interval(const RayRec& r, int y, int& k_lo, int& k_hi, int& x_lo, int& x_hi) {
  const int h = r.size >> 1;
  if (r.xmajor) {
    const int m = r.sy == 0 ? (y == r.by ? 0 : -1) : (y - r.by) * r.sy;  // steps along the minor (y) axis
    if (m < 0) return false;
    if (r.add == 0) {
      if (m != 0) return false;
      k_lo = 0, k_hi = r.size - 1;
    } else {
      // minor offset of cell k is floor((h + k*add) / size) == m  <=>  m*size <= h + k*add < (m+1)*size
      k_lo = m == 0 ? 0 : ceil_div(m * r.size - h, r.add, r.rcp_add);
      if (k_lo >= r.size) return false;
      k_hi = min(r.size - 1, ceil_div((m + 1) * r.size - h, r.add, r.rcp_add) - 1);
    }
    x_lo = r.bx + r.sx * k_lo;
    x_hi = r.bx + r.sx * k_hi;
  } else {
    const int k = (y - r.by) * r.sy;  // y-major rays have sy != 0 and exactly one cell per row
    if (k < 0 || k >= r.size) return false;
    k_lo = k_hi = k;
    x_lo = x_hi = r.bx + r.sx * floor_div_pos(h + k * r.add, r.size, r.rcp_size);
  }
  return true;
}

/// A scan row's ranges, sorted by (min, outline index of the run's end).
struct RowList {
  int lo[kMaxRanges], hi[kMaxRanges], sq[kMaxRanges];
  int n = 0;
  bool overflow = false;
  __device__ __forceinline__ void add(int l, int h, int s) {
    if (n >= kMaxRanges) {  // keep counting: the parity still decides std::logic_error
      overflow = true;
      ++n;
      return;
    }
    int k = n;
    while (k > 0 && (lo[k - 1] > l || (lo[k - 1] == l && sq[k - 1] > s))) {
      lo[k] = lo[k - 1], hi[k] = hi[k - 1], sq[k] = sq[k - 1];
      --k;
    }
    lo[k] = l, hi[k] = h, sq[k] = s;
    ++n;
  }
  /// One run: range of its end cells, counted twice at a cusp (generators.cpp:95-103,226-229).
  __device__ __forceinline__ void run(int y, int x_first, int x_last, int y_prev, int y_next, int s) {
    const int l = min(x_first, x_last), h = max(x_first, x_last);
    add(l, h, s);
    if ((y_prev - y) * (y - y_next) < 0) add(l, h, s);
  }
};

/// Builds the scan row `y` of the outline described by `rays[0..n_rays)` (cyclic).
__device__ __forceinline__ void build_row(const RayRec* __restrict__ rays, int n_rays, int y, RowList& L) {
  bool open = false;         // a run reaches the end of the previous ray and continues into this one
  bool run_first = false;    // the current run contains outline cell 0
  int run_x = 0, run_yprev = 0;
  bool f_valid = false;      // the run containing outline cell 0, closed; emitted last
  int f_xlast = 0, f_ynext = 0, f_seq = 0;
  bool tail_open = false;    // the last ray's run continues (cyclically) into cell 0
  int tail_x = 0, tail_yprev = 0;
  for (int e = 0; e < n_rays; ++e) {
    const RayRec r = rays[e];
    int k_lo, k_hi, x_lo, x_hi;
    if (!ray_row_interval(r, y, k_lo, k_hi, x_lo, x_hi)) continue;
    if (!(open && k_lo == 0)) {  // a new run starts inside this ray (or at its first cell)
      run_x = x_lo;
      run_yprev = k_lo == 0 ? rays[e == 0 ? n_rays - 1 : e - 1].ylast : y - r.sy;
      run_first = e == 0 && k_lo == 0;
    }
    open = false;
    const bool at_end = k_hi == r.size - 1;
    const int next_by = rays[e + 1 == n_rays ? 0 : e + 1].by;  // the cell after the ray's last cell, cyclically
    if (at_end && next_by == y) {
      if (e + 1 == n_rays) {
        tail_open = true, tail_x = run_x, tail_yprev = run_yprev;
      } else {
        open = true;
      }
      continue;
    }
    const int y_next = at_end ? next_by : y + r.sy;
    if (run_first) {
      f_valid = true, f_xlast = x_hi, f_ynext = y_next, f_seq = r.start + k_hi;
      run_first = false;
    } else {
      L.run(y, run_x, x_hi, run_yprev, y_next, r.start + k_hi);
    }
  }
  if (f_valid) {
    if (tail_open) L.run(y, tail_x, f_xlast, tail_yprev, f_ynext, f_seq);
    else L.run(y, rays[0].bx, f_xlast, rays[n_rays - 1].ylast, f_ynext, f_seq);
  }
}

/// Warp-cooperative: ray table (shared memory) of the outline through the discretised vertices `verts[0..nv)`:
/// rays v_e -> v_{e+1} with degenerate ones skipped, the closing ray iff more than one ray so far, else the one-cell
/// dummy ray holding the last point (generators.cpp:54-93).  Returns the number of rays (nv >= 1).
__device__ __forceinline__ int build_rays_from_verts(const int2* __restrict__ verts, int nv, RayRec* rays, int lane) {
  int n = 0;
  for (int e0 = 0; e0 < nv - 1; e0 += 32) {  // rays v_e -> v_{e+1}, degenerate ones skipped, order kept
    const int e = e0 + lane;
    bool nd = false;
    int2 b = make_int2(0, 0), c = b;
    if (e < nv - 1) {
      b = verts[e], c = verts[e + 1];
      nd = b.x != c.x || b.y != c.y;
    }
    const unsigned msk = __ballot_sync(0xFFFFFFFFu, nd);
    if (nd) rays[n + __popc(msk & ((1u << lane) - 1u))] = make_ray_rec(b, c);
    n += __popc(msk);
  }
  __syncwarp();
  if (lane == 0) {
    const int2 last = verts[nv - 1], first = verts[0];
    if (n > 1) {  // close the outline (generators.cpp:69-71)
      if (last.x != first.x || last.y != first.y) rays[n++] = make_ray_rec(last, first);
    } else {
      rays[n++] = dummy_ray_rec(last);
    }
  }
  n = __shfl_sync(0xFFFFFFFFu, n, 0);
  __syncwarp();
  int run = 0;
  for (int e0 = 0; e0 < n; e0 += 32) {  // outline index of each ray's first cell = exclusive prefix sum of sizes
    const int e = e0 + lane;
    const int sz = e < n ? rays[e].size : 0;
    int incl = sz;
    for (int d = 1; d < 32; d <<= 1) {
      const int up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += up;
    }
    if (e < n) rays[e].start = run + incl - sz;
    run += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  __syncwarp();
  return n;
}

/// Warp-cooperative: discretises the polygon, builds the ray table in shared memory and returns the
/// number of rays (0 for an empty polygon); `ok` = all coordinates in range; bbox in x0..y1.
template <int XF>
__device__ __forceinline__ int build_rays(const double* __restrict__ poly, int nv, const Pose& T, const MapView& m, int2* verts,
                                          RayRec* rays, int lane, bool& ok, int& x0, int& y0, int& x1, int& y1) {
  bool good = true;
  int lx0 = 0x7FFFFFFF, ly0 = 0x7FFFFFFF, lx1 = -0x7FFFFFFF - 1, ly1 = -0x7FFFFFFF - 1;
  for (int v = lane; v < nv; v += 32) {
    int cx, cy;
    good &= to_cell<XF>(T, poly[2 * v], poly[2 * v + 1], m, cx, cy);
    verts[v] = make_int2(cx, cy);
    lx0 = min(lx0, cx), lx1 = max(lx1, cx), ly0 = min(ly0, cy), ly1 = max(ly1, cy);
  }
  ok = __all_sync(0xFFFFFFFFu, good);
  x0 = __reduce_min_sync(0xFFFFFFFFu, lx0), x1 = __reduce_max_sync(0xFFFFFFFFu, lx1);
  y0 = __reduce_min_sync(0xFFFFFFFFu, ly0), y1 = __reduce_max_sync(0xFFFFFFFFu, ly1);
  __syncwarp();
  if (nv == 0) return 0;
  return build_rays_from_verts(verts, nv, rays, lane);
}

/// Per-pose result of the warp; valid in every lane after `finish`.
template <int OP>
struct WarpFold;

template <>
struct WarpFold<OP_IS_FREE> {
  bool hit = false;
  __device__ __forceinline__ bool done() const { return hit; }
  __device__ __forceinline__ void merge(const Fold<OP_IS_FREE>& f) { hit |= __any_sync(0xFFFFFFFFu, f.hit) != 0; }
  __device__ __forceinline__ void store(const OutView& o, long long i) const { emit_free(o, i, !hit); }
};

template <>
struct WarpFold<OP_ACCUMULATE> {
  unsigned long long sum = 0;
  int early = 0;
  __device__ __forceinline__ bool done() const { return early != 0; }
  __device__ __forceinline__ void merge(const Fold<OP_ACCUMULATE>& f) {  // lanes = rows in ascending order
    const unsigned stopped = __ballot_sync(0xFFFFFFFFu, f.early != 0);
    if (stopped) {
      early = __shfl_sync(0xFFFFFFFFu, f.early, __ffs(stopped) - 1);  // the first row with a negative value wins
      return;
    }
    unsigned long long s = f.sum;
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    sum += s;
  }
  __device__ __forceinline__ void store(const OutView& o, long long i) const {
    if (o.cost) o.cost[i] = early ? static_cast<double>(early) : static_cast<double>(sum);
    emit_free(o, i, early == 0);
  }
};

template <>
struct WarpFold<OP_MAX> {
  int best = 0;
  bool outside = false;
  __device__ __forceinline__ bool done() const { return outside; }
  __device__ __forceinline__ void merge(const Fold<OP_MAX>& f) {
    outside |= __any_sync(0xFFFFFFFFu, f.outside) != 0;
    best = max(best, __reduce_max_sync(0xFFFFFFFFu, f.value()));
  }
  __device__ __forceinline__ void store(const OutView& o, long long i) const {
    if (o.cost) o.cost[i] = outside ? -3.0 : static_cast<double>(best);
    emit_free(o, i, !outside && best != kLethal);
  }
};

/// One pose, one warp.  Returns the status; the result is in `wf`.
template <int OP, int XF>
__device__ __forceinline__ int area_rows_pose(const double* __restrict__ poly, int nv, const Pose& T, const MapView& m, int2* verts,
                                              RayRec* rays, int lane, WarpFold<OP>& wf) {
  bool ok;
  int x0, y0, x1, y1;
  const int n_rays = build_rays<XF>(poly, nv, T, m, verts, rays, lane, ok, x0, y0, x1, y1);
  if (!ok) return ST_COORD_RANGE;
  if (n_rays == 0) return ST_OK;  // empty generator
  if (y0 == y1) {  // every cell on one row: ranges (min,min),(max,max) -> [min, max]  (generators.cpp:197-204)
    Fold<OP> f;
    if (lane == 0) f.span(m, y0, x0, x1);
    wf.merge(f);
    return ST_OK;
  }
  bool odd = false, overflow = false;
  for (int yb = y0; yb <= y1; yb += 32) {
    const int y = yb + lane;
    RowList L;
    if (y <= y1) build_row(rays, n_rays, y, L);
    odd |= (L.n & 1) != 0;
    overflow |= L.overflow;
    if (!wf.done()) {  // later chunks cannot change a result that is already decided, only the status
      Fold<OP> f;
      if (!(L.n & 1) && !L.overflow)
        for (int k = 0; k + 1 < L.n; k += 2)
          if (!f.span(m, y, L.lo[k], L.hi[k + 1])) break;
      wf.merge(f);
    }
  }
  if (__any_sync(0xFFFFFFFFu, odd)) return ST_LOGIC_ERROR;  // std::logic_error("Even number of line-pairs expected")
  if (__any_sync(0xFFFFFFFFu, overflow)) return ST_ROW_COMPLEXITY;
  return ST_OK;
}


/// expand(area_generator): counts (cells == nullptr) or writes the area cells of one pose in generator
/// order — rows ascending, merged ranges in sorted order, x ascending (generators.hpp:417-423,
/// impl/expand.hpp:23-39).  One warp; returns the status, `total` = number of cells.
template <int XF>
__device__ __forceinline__ int area_rows_expand(const double* __restrict__ poly, int nv, const Pose& T, const MapView& m, int2* verts,
                                                RayRec* rays, int lane, int2* __restrict__ cells, long long& total) {
  total = 0;
  bool ok;
  int x0, y0, x1, y1;
  const int n_rays = build_rays<XF>(poly, nv, T, m, verts, rays, lane, ok, x0, y0, x1, y1);
  if (!ok) return ST_COORD_RANGE;
  if (n_rays == 0) return ST_OK;
  if (y0 == y1) {  // single row: [min, max]
    total = x1 - x0 + 1;
    if (cells)
      for (int x = x0 + lane; x <= x1; x += 32) cells[x - x0] = make_int2(x, y0);
    return ST_OK;
  }
  bool odd = false, overflow = false;
  long long base = 0;
  for (int yb = y0; yb <= y1; yb += 32) {
    const int y = yb + lane;
    RowList L;
    if (y <= y1) build_row(rays, n_rays, y, L);
    odd |= (L.n & 1) != 0;
    overflow |= L.overflow;
    int mine = 0;
    if (!(L.n & 1) && !L.overflow)
      for (int k = 0; k + 1 < L.n; k += 2) mine += L.hi[k + 1] - L.lo[k] + 1;
    int incl = mine;
    for (int d = 1; d < 32; d <<= 1) {
      const int up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += up;
    }
    if (cells && !(L.n & 1) && !L.overflow) {
      long long at = base + incl - mine;
      for (int k = 0; k + 1 < L.n; k += 2)
        for (int x = L.lo[k]; x <= L.hi[k + 1]; ++x) cells[at++] = make_int2(x, y);
    }
    base += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  if (__any_sync(0xFFFFFFFFu, odd)) return ST_LOGIC_ERROR;
  if (__any_sync(0xFFFFFFFFu, overflow)) return ST_ROW_COMPLEXITY;
  total = base;
  return ST_OK;
}

// ---------------------------------------------------------------------------------------------------
// Several outlines: area_generator(resolution, polygon_vec) — nested outlines give holes (generalised
// even-odd rule, src/cover/generators.cpp:126-135,181-233).  Every outline contributes its own cyclic ray
// list; all ranges of a row land in ONE list, ordered by (min, outline order, outline index).
// ---------------------------------------------------------------------------------------------------
constexpr int kRowsMaxOutlines = 16;

struct OutlineRec {
  int ray_begin, n_rays;  // slice of the warp's ray table
  int single_row;         // every cell on one row: ranges (min,min),(max,max) (generators.cpp:197-204)
  int y, min_x, max_x, seq;
};

template <int OP, int XF>
__device__ __forceinline__ int area_rows_multi_pose(const double* __restrict__ poly, const int* __restrict__ starts, int n_outlines,
                                                    const Pose& T, const MapView& m, int2* verts, RayRec* rays, OutlineRec* recs, int lane,
                                                    WarpFold<OP>& wf) {
  bool ok_all = true;
  int y_lo = 0x7FFFFFFF, y_hi = -0x7FFFFFFF - 1, ray_base = 0, seq_base = 0;
  for (int o = 0; o < n_outlines; ++o) {
    const int nv = starts[o + 1] - starts[o];
    OutlineRec rec{ray_base, 0, 0, 0, 0, 0, seq_base};
    if (nv > 0) {  // (an empty outline is skipped; the reference reads out of bounds there)
      bool ok;
      int x0, y0, x1, y1;
      const int n = build_rays<XF>(poly + 2 * starts[o], nv, T, m, verts, rays + ray_base, lane, ok, x0, y0, x1, y1);
      ok_all &= ok;
      for (int e = lane; e < n; e += 32) rays[ray_base + e].start += seq_base;
      __syncwarp();
      const RayRec last = rays[ray_base + n - 1];
      rec = OutlineRec{ray_base, n, y0 == y1, y0, x0, x1, seq_base};
      y_lo = min(y_lo, y0), y_hi = max(y_hi, y1);
      ray_base += n;
      seq_base = last.start + last.size;
    }
    if (lane == 0) recs[o] = rec;
    __syncwarp();
  }
  if (!ok_all) return ST_COORD_RANGE;
  if (ray_base == 0) return ST_OK;  // no cells at all
  bool odd = false, overflow = false;
  for (int yb = y_lo; yb <= y_hi; yb += 32) {
    const int y = yb + lane;
    RowList L;
    if (y <= y_hi) {
      for (int o = 0; o < n_outlines; ++o) {
        const OutlineRec rec = recs[o];
        if (rec.n_rays == 0) continue;
        if (rec.single_row) {
          if (y == rec.y) {
            L.add(rec.min_x, rec.min_x, rec.seq);
            L.add(rec.max_x, rec.max_x, rec.seq);
          }
        } else {
          build_row(rays + rec.ray_begin, rec.n_rays, y, L);
        }
      }
    }
    odd |= (L.n & 1) != 0;
    overflow |= L.overflow;
    if (!wf.done()) {
      Fold<OP> f;
      if (!(L.n & 1) && !L.overflow)
        for (int k = 0; k + 1 < L.n; k += 2)
          if (!f.span(m, y, L.lo[k], L.hi[k + 1])) break;
      wf.merge(f);
    }
  }
  if (__any_sync(0xFFFFFFFFu, odd)) return ST_LOGIC_ERROR;
  if (__any_sync(0xFFFFFFFFu, overflow)) return ST_ROW_COMPLEXITY;
  return ST_OK;
}

}  // namespace cvr
