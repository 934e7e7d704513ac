// lattice_words.cuh — sweeps over an (x, y, heading) pose lattice at cell pitch: verdicts bit-parallel in x, value folds
// from per-row tables.
//
// is_free(polygon, pose, map) (src/cover/costmap.cpp:78-85) and is_free(outline_generator(...), map)
// (src/cover/impl/costmap.hpp:23-27) for every pose of a cover_lattice whose x pitch is one cell; accumulate_cost
// (impl/costmap.hpp:50-63) and the max_cost extension of areas over the same lattices.  Kernels:
//
//  k_lw_prepare        one CTA per heading.  The pose transform separates - the x cells of the vertices depend on
//                      (heading, ix) only, the y cells on (heading, iy) only - so a heading needs nx + ny vertex profiles,
//                      not nx * ny transforms (the same __dmul_rn/__dadd_rn/__dsub_rn/__double2int_rz sequence as
//                      to_cell, per coordinate).  Columns are sorted into CLASSES of identical (vertex offsets relative to
//                      vertex 0, anchor = cell of vertex 0 minus ix), rows into classes of identical offsets; nearly every
//                      heading has one class per axis, the axis-aligned headings (vertices within rounding noise of a cell
//                      boundary) a handful.  The comparison is exact (no hashing).  Then one warp per (x class, y class):
//                      the discretised polygon relative to vertex 0 is rasterised once with the closed-form scan-row
//                      construction of area_rows.cuh (runs, cusps, stable sort by `min`, pair merge -
//                      generators.cpp:189-255) into a list of spans (dx, dy, width <= 32) in a global pool.
//  k_lw_sweep_simple   one warp per (heading, 32 lattice rows, 32 * W columns) of a single-shape heading inside the map:
//                      LANE = lattice row, every thread keeps the verdicts of 32 * W x-adjacent poses as W words.  "Some
//                      cell of a w-cell window is lethal" is the OR of two overlapping power-of-two windows of the lethal
//                      bit planes (k_lethal_planes), and for 32 x-adjacent poses those are 32 consecutive plane bits: per
//                      span W + 2 word loads (from a shared-memory tile of the planes) and 2 W funnel shifts decide 32 W
//                      poses.
//  k_lw_sweep_general  the other tasks (several classes per heading: columns of another x class are masked out and walked
//                      with their own shape, every lane walks the shape of its row's y class; map border; statuses).
//  k_lw_values         accumulate_cost / max_cost: warp = 64 x-adjacent poses of one lattice row, two table reads per span
//                      and pose (prefix words / sliding-maximum planes, k_max_planes).
// Nothing here waits for another warp: the sweeps only ever read finished tables.
#pragma once

#include "area_rows.cuh"

#include <cstdio>

namespace cvr {

constexpr int kLwClasses = 32;                                   // distinct profiles per axis and heading
constexpr int kLwBad = 0xFF, kLwOverflow = 0xFE, kLwNone = 0xFD;  // class codes: coordinates out of range / table full / no such column
constexpr int kLwUnfit = -1;                                     // shape status besides cvr::Status: left to the general kernels
constexpr int kLwWarps = 8;
constexpr int kLwPrepThreads = 512;                              // k_lw_prepare: threads per CTA (up to 16 of its warps rasterise shapes)
constexpr int kLwSpanCache = 128;                                // spans of the current shape kept in shared memory per warp

/// One span of a rasterised shape, decoded for the walk: .x = first cell relative to vertex 0, .y = word offset
/// k * plane_words + dy (k = floor(log2(width)): the plane; dy = row relative to vertex 0), .z = width - 2^k (start of
/// the second power-of-two window), .w = (dy + 32768) | k << 16 (for the bounds-checked walk at the map border).
typedef int4 LwSpan;

struct LwShape {
  int status, n_spans, span_first;
  int xb, x1, yb, rows;  // bounding box of the vertices relative to vertex 0: x in [xb, x1], y in [yb, yb + rows - 1]
  int pad;               // the largest plane index k = floor(log2(width)) of its spans
};

/// Everything k_lw_sweep_simple needs to know about a heading, in one record (one load instead of a chain of three):
/// `simple` as in LwTables::simple, the anchor of x class 0 and the heading's one shape.
struct LwHead {
  int simple, anchor0, span_first, n_spans;
  int xb, x1, yb, rows;
  int kmax, pad0, pad1, pad2;
};

struct LwTables {
  int* n_xcls;         // [n_head]
  int* n_ycls;         // [n_head]
  int* shape_base;     // [n_head] shape (xc, yc) of the heading lives at shape_base + xc * n_ycls + yc; -1: pool exhausted
  int* simple;         // [n_head] 1: one x class, one y class, every column and row in range, the shape rasterised fine
  LwHead* heads;       // [n_head]
  int* order;          // [n_head] the headings, those that are not simple first (the general sweep's CTAs of the few
                       //   headings with real work start together instead of one heading after the other)
  uint8_t* xcls;       // [n_head][nx]
  uint8_t* ycls;       // [n_head][ny]
  int* yc0;            // [n_head][ny] map row of vertex 0
  int* xanchor;        // [n_head][kLwClasses] map column of vertex 0 minus ix
  LwShape* shapes;     // [max_shapes]
  LwSpan* spans;       // [span_pool]
  unsigned int* counters;  // [0] shapes handed out, [1] spans handed out, [2] / [3] headings placed at the front / back of `order`
  int max_shapes, span_pool;
};

/// x cell of one vertex: the x half of to_cell<XF_PATH_A> (same operations in the same order).
__device__ __forceinline__ bool lw_cell_x(double c, double s, double tx, double px, double py, const MapView& m, int& cx) {
  const double q = __dsub_rn(__dadd_rn(tx, __dadd_rn(__dmul_rn(c, px), __dmul_rn(-s, py))), m.origin_x);
  const double sc = __dmul_rn(q, m.inv_res);
  cx = __double2int_rz(sc);
  return fabs(sc) < kCoordLimit;
}
/// y cell of one vertex: the y half of to_cell<XF_PATH_A>.
__device__ __forceinline__ bool lw_cell_y(double c, double s, double ty, double px, double py, const MapView& m, int& cy) {
  const double q = __dsub_rn(__dadd_rn(ty, __dadd_rn(__dmul_rn(s, px), __dmul_rn(c, py))), m.origin_y);
  const double sc = __dmul_rn(q, m.inv_res);
  cy = __double2int_rz(sc);
  return fabs(sc) < kCoordLimit;
}

/// Spans of scan row `y` of the shape (area: merged range pairs; outline: the rays' cells on the row), each cut into
/// pieces of at most 32 cells: calls emit(index, x_first, width) per piece and returns their number.
template <class Emit>
__device__ __forceinline__ int lw_row_spans(const RayRec* __restrict__ rays, int n_rays, int y, bool outline, bool single_row, int x0, int x1,
                                            bool& odd, bool& overflow, Emit emit) {
  int n = 0;
  auto piece = [&](int lo, int hi) {
    for (int x = lo; x <= hi; x += 32) {
      emit(n, x, min(32, hi - x + 1));
      ++n;
    }
  };
  if (outline) {
    for (int e = 0; e < n_rays; ++e) {
      int k_lo, k_hi, xa, xb;
      if (ray_row_interval(rays[e], y, k_lo, k_hi, xa, xb)) piece(min(xa, xb), max(xa, xb));
    }
  } else if (single_row) {  // every cell on one row: ranges (min,min),(max,max) -> [min, max] (generators.cpp:197-204)
    piece(x0, x1);
  } else {
    RowList L;
    build_row(rays, n_rays, y, L);
    odd |= (L.n & 1) != 0;
    overflow |= L.overflow;
    if (!(L.n & 1) && !L.overflow)
      for (int k = 0; k + 1 < L.n; k += 2) piece(L.lo[k], L.hi[k + 1]);  // [r[2k].min, r[2k+1].max] (generators.cpp:253-254)
  }
  return n;
}

/// One warp: rasterises the shape with vertices `verts` (relative to vertex 0) into the span pool; returns its record.
__device__ __forceinline__ LwShape lw_rasterise(int2* verts, int nv, RayRec* rays, int lane, int outline, long long plane_words, const LwTables& t) {
  int lx0 = 0x7FFFFFFF, ly0 = 0x7FFFFFFF, lx1 = -0x7FFFFFFF - 1, ly1 = -0x7FFFFFFF - 1;
  for (int v = lane; v < nv; v += 32) {
    const int2 c = verts[v];
    lx0 = min(lx0, c.x), lx1 = max(lx1, c.x), ly0 = min(ly0, c.y), ly1 = max(ly1, c.y);
  }
  const int x0 = __reduce_min_sync(0xFFFFFFFFu, lx0), x1 = __reduce_max_sync(0xFFFFFFFFu, lx1);
  const int y0 = __reduce_min_sync(0xFFFFFFFFu, ly0), y1 = __reduce_max_sync(0xFFFFFFFFu, ly1);
  const int n_rays = build_rays_from_verts(verts, nv, rays, lane);
  const bool single_row = y0 == y1;
  int status = ST_OK, n_total = 0, span_first = 0, kmax = 0;
  if (y1 - y0 > 30000 || x1 - x0 > 1000000) status = kLwUnfit;  // (dy travels in 16 bits)
  for (int pass = 0; pass < 2 && status == ST_OK; ++pass) {
    int base = 0;
    bool odd = false, overflow = false;
    for (int yb = y0; yb <= y1; yb += 32) {
      const int y = yb + lane;
      const bool live = y <= y1 && !(single_row && !outline && lane != 0);
      int mine = 0;
      if (live) mine = lw_row_spans(rays, n_rays, y, outline != 0, single_row, x0, x1, odd, overflow, [](int, int, int) {});
      int incl = mine;
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += up;
      }
      if (pass == 1 && live) {
        LwSpan* dst = t.spans + span_first + base + incl - mine;
        bool o2 = false, v2 = false;
        lw_row_spans(rays, n_rays, y, outline != 0, single_row, x0, x1, o2, v2, [&](int k, int x, int width) {
          const int lg = 31 - __clz(width);
          kmax = max(kmax, lg);
          dst[k] = make_int4(x, static_cast<int>(lg * plane_words) + y, width - (1 << lg), (y + 32768) | (lg << 16));
        });
      }
      base += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (pass == 0) {
      if (__any_sync(0xFFFFFFFFu, odd)) status = ST_LOGIC_ERROR;  // std::logic_error("Even number of line-pairs expected")
      else if (__any_sync(0xFFFFFFFFu, overflow)) status = ST_ROW_COMPLEXITY;
      n_total = base;
      if (status == ST_OK) {
        unsigned int first = 0;
        if (lane == 0) first = atomicAdd(&t.counters[1], static_cast<unsigned int>(n_total));
        first = __shfl_sync(0xFFFFFFFFu, first, 0);
        if (first + static_cast<unsigned int>(n_total) > static_cast<unsigned int>(t.span_pool)) status = kLwUnfit;
        span_first = static_cast<int>(first);
      }
    }
  }
  kmax = __reduce_max_sync(0xFFFFFFFFu, kmax);
  return LwShape{status, n_total, span_first, x0, x1, y0, y1 - y0 + 1, kmax};
}

/// Per heading (one CTA): column and row classes, then one rasterisation per (x class, y class).
/// Dynamic shared memory: polygon | x offsets [K][nv] | y offsets [K][nv] | anchors [K] | per shape warp: rays, vertices.
__global__ void __launch_bounds__(kLwPrepThreads) k_lw_prepare(MapView m, const double* __restrict__ poly_g, int nv, PoseSource src, int it_lo,
                                                              LwTables t, int outline, long long plane_words, int shape_warps, int strip_cols) {
  extern __shared__ double smem_d[];
  __shared__ int s_n[2], s_elect, s_base, s_flaw, s_bad_shape, s_ymin, s_ymax;
  const int tid = threadIdx.x;
  double* poly = smem_d;
  int* xoff = reinterpret_cast<int*>(smem_d + 2 * nv);
  int* yoff = xoff + kLwClasses * nv;
  int* anchor_tab = yoff + kLwClasses * nv;
  for (int i = tid; i < 2 * nv; i += blockDim.x) poly[i] = poly_g[i];
  const int h = blockIdx.x, it = it_lo + h;
  const double2 cs = __ldg(reinterpret_cast<const double2*>(src.theta_cs) + it);
  if (tid == 0) s_n[0] = s_n[1] = 0, s_elect = 0x7FFFFFFF, s_flaw = 0, s_bad_shape = 0, s_ymin = 0x7FFFFFFF, s_ymax = -0x7FFFFFFF - 1;
  __syncthreads();
#pragma unroll 1
  for (int axis = 0; axis < 2; ++axis) {
    const bool is_x = axis == 0;
    const int n = is_x ? src.nx : src.ny;
    int* off_tab = is_x ? xoff : yoff;
    uint8_t* cls_out = is_x ? t.xcls + static_cast<long long>(h) * src.nx : t.ycls + static_cast<long long>(h) * src.ny;
    for (int base = 0; base < n; base += blockDim.x) {
      const int idx = base + tid;
      const bool valid = idx < n;
      // t = x0 + i * dx: multiply, then add, each rounded (the ABI's stated rule, cover_b200.h)
      const double tr = is_x ? __dadd_rn(src.x0, __dmul_rn(static_cast<double>(idx), src.dx)) : __dadd_rn(src.y0, __dmul_rn(static_cast<double>(idx), src.dy));
      auto cell_of = [&](int v, int& cell) {
        return is_x ? lw_cell_x(cs.x, cs.y, tr, poly[2 * v], poly[2 * v + 1], m, cell) : lw_cell_y(cs.x, cs.y, tr, poly[2 * v], poly[2 * v + 1], m, cell);
      };
      int cell0 = 0;
      bool ok = true;
      for (int v = 0; v < nv; ++v) {
        int c;
        ok &= cell_of(v, c);
        if (v == 0) cell0 = c;
      }
      const int raw0 = cell0;  // vertex 0's cell without the offset: what the other vertices' cells are relative to
      cell0 += is_x ? src.off_x : src.off_y;  // the offset overloads: added after discretisation, so only vertex 0's cell moves
      const int anchor = is_x ? cell0 - idx : 0;
      bool pending = valid && ok;
      int my = valid ? (ok ? kLwOverflow : kLwBad) : kLwNone;
      int tried = 0;
      for (;;) {
        const int ncur = s_n[axis];  // (only changes between the barriers below)
        if (pending) {
          for (int k = tried; k < ncur && pending; ++k) {
            if (is_x && anchor_tab[k] != anchor) continue;
            bool same = true;
            for (int v = 1; v < nv && same; ++v) {
              int c;
              cell_of(v, c);
              same = c - raw0 == off_tab[k * nv + v];
            }
            if (same) my = k, pending = false;
          }
          if (pending && ncur >= kLwClasses) pending = false;  // table full: `my` stays kLwOverflow
        }
        tried = ncur;
        if (!__syncthreads_or(pending)) break;
        if (pending) atomicMin(&s_elect, tid);
        __syncthreads();
        if (tid == s_elect) {  // the first unmatched column / row founds the next class
          for (int v = 0; v < nv; ++v) {
            int c;
            cell_of(v, c);
            off_tab[ncur * nv + v] = c - raw0;
          }
          if (is_x) {
            anchor_tab[ncur] = anchor;
            t.xanchor[h * kLwClasses + ncur] = anchor;
          }
          my = ncur, pending = false;
          s_n[axis] = ncur + 1;
        }
        __syncthreads();
        if (tid == 0) s_elect = 0x7FFFFFFF;
        __syncthreads();
      }
      if (valid) {
        cls_out[idx] = static_cast<uint8_t>(my);
        if (!is_x) {
          t.yc0[static_cast<long long>(h) * src.ny + idx] = cell0;
          atomicMin(&s_ymin, cell0), atomicMax(&s_ymax, cell0);
        }
        if (my >= kLwClasses) s_flaw = 1;  // (benign race: everybody writes 1)
      }
    }
    __syncthreads();
  }
  // shape slots of this heading: one per (x class, y class)
  const int nxc = s_n[0], nyc = s_n[1], n_shapes = nxc * nyc;
  if (tid == 0) {
    t.n_xcls[h] = nxc, t.n_ycls[h] = nyc;
    const unsigned int b = atomicAdd(&t.counters[0], static_cast<unsigned int>(n_shapes));
    s_base = b + static_cast<unsigned int>(n_shapes) <= static_cast<unsigned int>(t.max_shapes) ? static_cast<int>(b) : -1;
    t.shape_base[h] = s_base;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  if (s_base >= 0 && warp < shape_warps) {
    RayRec* rays = reinterpret_cast<RayRec*>(anchor_tab + kLwClasses) + warp * (nv + 1);
    int2* verts = reinterpret_cast<int2*>(reinterpret_cast<RayRec*>(anchor_tab + kLwClasses) + shape_warps * (nv + 1)) + warp * nv;
    for (int i = warp; i < n_shapes; i += shape_warps) {
      const int xc = i / nyc, yc = i - xc * nyc;
      __syncwarp();
      for (int v = lane; v < nv; v += 32) verts[v] = make_int2(xoff[xc * nv + v], yoff[yc * nv + v]);
      __syncwarp();
      const LwShape sh = lw_rasterise(verts, nv, rays, lane, outline, plane_words, t);
      if (lane == 0) {
        t.shapes[s_base + i] = sh;
        if (sh.status != ST_OK) s_bad_shape = 1;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    int simple = (s_base >= 0 && n_shapes == 1 && !s_flaw && !s_bad_shape) ? 1 : 0;
    if (simple) {  // 2: and every pose of the heading's lattice keeps its shape inside the map - the general sweep has nothing to do here
      const LwShape& sh = t.shapes[s_base];
      const int xa = anchor_tab[0];
      const int last_col = (src.nx + strip_cols - 1) / strip_cols * strip_cols - 1;  // (lw_task_simple tests whole strips of 32 W columns)
      if (xa + sh.xb >= 0 && xa + last_col + sh.x1 < m.size_x && s_ymin + sh.yb >= 0 && s_ymax + sh.yb + sh.rows <= m.size_y) simple = 2;
    }
    t.simple[h] = simple;
    LwHead hd{simple, anchor_tab[0], 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (simple) {
      const LwShape& sh = t.shapes[s_base];
      hd.span_first = sh.span_first, hd.n_spans = sh.n_spans, hd.xb = sh.xb, hd.x1 = sh.x1, hd.yb = sh.yb, hd.rows = sh.rows, hd.kmax = sh.pad;
    }
    t.heads[h] = hd;
    if (simple) t.order[gridDim.x - 1u - atomicAdd(&t.counters[3], 1u)] = h;
    else t.order[atomicAdd(&t.counters[2], 1u)] = h;
  }
}

/// ORs, for the 32 * W x-adjacent poses whose vertex 0 sits at columns x_base .. x_base + 32 W - 1 of map row y_lane,
/// "some cell of the span is lethal" over all spans of ONE shape into acc (every needed word inside the planes).
template <int W>
__device__ __forceinline__ void lw_walk(const uint32_t* __restrict__ bits, int ypitch, const LwSpan* __restrict__ spans, int n_spans, int x_base,
                                        int y_lane, uint32_t (&acc)[W]) {
  const long long col = ypitch;  // (64-bit: the column offsets j * col below are loop constants)
#pragma unroll 2
  for (int s = 0; s < n_spans; ++s) {
    const LwSpan sp = spans[s];
    const int X = x_base + sp.x;
    const int c = X >> 5, sh = X & 31;
    uint32_t L[W + 2];
    const uint32_t* __restrict__ p = bits + static_cast<unsigned int>(c * ypitch + sp.y + y_lane);
#pragma unroll
    for (int j = 0; j < W + 2; ++j) L[j] = __ldg(p + j * col);
    // second power-of-two window: `tail` cells later (the same window again when the width is a power of two)
    const int sh2 = sh + sp.z;
    if (sh2 < 32) {
#pragma unroll
      for (int w = 0; w < W; ++w) acc[w] |= __funnelshift_r(L[w], L[w + 1], sh) | __funnelshift_r(L[w], L[w + 1], sh2);
    } else {
#pragma unroll
      for (int w = 0; w < W; ++w) acc[w] |= __funnelshift_r(L[w], L[w + 1], sh) | __funnelshift_r(L[w + 1], L[w + 2], sh2 - 32);
    }
  }
}

/// The same walk where every LANE has its own shape (rows of different y classes side by side in one warp): lane-private
/// span lists of n_sp spans (0: nothing to walk), read from global memory; n_max = the longest list of the warp.
/// GUARD: words outside the planes read as zero, and a pose whose span leaves the map counts as hit - cells outside the
/// map are lethal (costmap.hpp:81-91).  This is tested per SPAN, not per bounding box: the reference's ranges use the
/// runs' end points only, so the outermost cell of a cusp may belong to no range (generators.cpp:211-232).
template <int W, bool GUARD>
__device__ __forceinline__ void lw_walk_lanes(const uint32_t* __restrict__ bits, int ypitch, long long plane_words, int wcols, int size_x, int size_y,
                                              const LwSpan* __restrict__ spans, int n_sp, int n_max, int x_base, int y_lane, uint32_t (&acc)[W]) {
  const long long col = ypitch;
  for (int s = 0; s < n_max; ++s) {
    const bool live = s < n_sp;
    LwSpan sp = make_int4(0, 0, 0, 32768);
    if (live) sp = __ldg(spans + s);
    const int X = x_base + sp.x;
    const int c = X >> 5, sh = X & 31;
    uint32_t L[W + 2];
    if (!GUARD) {
      const uint32_t* __restrict__ p = bits + static_cast<unsigned int>(c * ypitch + sp.y + y_lane);
#pragma unroll
      for (int j = 0; j < W + 2; ++j) L[j] = live ? __ldg(p + j * col) : 0u;
    } else {
      const int y = y_lane + (sp.w & 0xFFFF) - 32768, k = sp.w >> 16;
      const bool y_ok = static_cast<unsigned>(y) < static_cast<unsigned>(size_y);
#pragma unroll
      for (int j = 0; j < W + 2; ++j) {
        const int cc = c + j;
        L[j] = (live && y_ok && static_cast<unsigned>(cc) < static_cast<unsigned>(wcols)) ? __ldg(bits + k * plane_words + cc * col + y) : 0u;
      }
      if (live) {
        const int width = (1 << k) + sp.z;
#pragma unroll
        for (int w = 0; w < W; ++w) {
          // pose j of word w has the span's cells at columns X + 32 w + j .. + width - 1: inside iff lo_ok <= j < hi_ok
          const int lo_ok = -(X + 32 * w), hi_ok = size_x - (width - 1) - (X + 32 * w);
          uint32_t in = y_ok ? 0xFFFFFFFFu : 0u;
          if (lo_ok > 0) in &= lo_ok >= 32 ? 0u : (0xFFFFFFFFu << lo_ok);
          if (hi_ok < 32) in &= hi_ok <= 0 ? 0u : ((1u << hi_ok) - 1u);
          acc[w] |= ~in;
        }
      }
    }
    const int sh2 = sh + sp.z;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const uint32_t second = sh2 < 32 ? __funnelshift_r(L[w], L[w + 1], sh2) : __funnelshift_r(L[w + 1], L[w + 2], sh2 - 32);
      acc[w] |= __funnelshift_r(L[w], L[w + 1], sh) | second;
    }
  }
}

struct LwGeom {
  int it_lo, n_head;
  int nb, ns;    // 32-row blocks per heading, 32 W-column strips per row
  int nbg;       // groups of kLwWarps row blocks: grid = (nbg, ns, n_head)
  int aligned4;  // pose index of (row, strip start) is a multiple of 4 for every row: verdict bytes go out as words
};

/// Results of one warp task: lane = lattice row, fr / valid = verdict and ownership bits of its 32 W poses starting at pose
/// index i0.  `plain`: no pose of the task was failed or deferred (statuses and deferred poses are written elsewhere).
template <int W>
__device__ __forceinline__ void lw_store(const OutView& out, const uint32_t (&fr)[W], const uint32_t (&valid)[W], const uint32_t (&skip)[W],
                                         long long i0, bool any_valid, unsigned live_rows, bool plain, int aligned4, int warp, int lane,
                                         uint32_t (*s_bits)[2 * W]) {
  if (out.free_bits && any_valid) {
    // the row's 32 W verdicts start at bit i0 of the packed array: W + 1 words, the inner ones owned by this thread
    const long long wi = i0 >> 5;  // (arithmetic: i0 may be negative for the batch's first row; its valid bits are not)
    const int sh = static_cast<int>(i0 & 31);
#pragma unroll
    for (int j = 0; j <= W; ++j) {
      const uint32_t lo_f = j > 0 ? fr[j - 1] : 0u, hi_f = j < W ? fr[j] : 0u;
      const uint32_t lo_v = j > 0 ? valid[j - 1] : 0u, hi_v = j < W ? valid[j] : 0u;
      const uint32_t word = __funnelshift_l(lo_f, hi_f, sh), owned = __funnelshift_l(lo_v, hi_v, sh);
      if (owned == 0xFFFFFFFFu) out.free_bits[wi + j] = word;
      else if (word != 0u) atomicOr(out.free_bits + (wi + j), word);  // (a word shared with the neighbouring task: zeroed before the sweep)
    }
  }
  if (out.is_free || out.status) {
    if (aligned4 && plain) {
      // transpose through shared memory so that a warp writes one lattice row at a time: lane l owns poses 4 l .. 4 l + 3
      __syncwarp();
#pragma unroll
      for (int w = 0; w < W; ++w) s_bits[lane][w] = fr[w], s_bits[lane][W + w] = valid[w];
      __syncwarp();
      unsigned rows_left = live_rows;
      while (rows_left) {
        const int r = __ffs(rows_left) - 1;
        rows_left &= rows_left - 1u;
        const long long ir = __shfl_sync(0xFFFFFFFFu, i0, r);
        for (int gq = lane; gq < 8 * W; gq += 32) {
          const int sh4 = (gq & 7) * 4;
          const uint32_t nib = (s_bits[r][gq >> 3] >> sh4) & 0xFu, vn = (s_bits[r][W + (gq >> 3)] >> sh4) & 0xFu;
          const long long i = ir + 4 * gq;
          if (vn == 0xFu) {
            if (out.is_free) *reinterpret_cast<uint32_t*>(out.is_free + i) = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
            if (out.status) *reinterpret_cast<uint32_t*>(out.status + i) = 0u;  // ST_OK
          } else if (vn) {
            for (int b = 0; b < 4; ++b)
              if ((vn >> b) & 1u) {
                if (out.is_free) out.is_free[i + b] = static_cast<uint8_t>((nib >> b) & 1u);
                if (out.status) out.status[i + b] = ST_OK;
              }
          }
        }
      }
    } else {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        uint32_t b = valid[w] & ~skip[w];
        while (b) {
          const int j = __ffs(b) - 1;
          b &= b - 1u;
          const long long i = i0 + 32 * w + j;
          if (out.is_free) out.is_free[i] = static_cast<uint8_t>((fr[w] >> j) & 1u);
          if (out.status) out.status[i] = ST_OK;
        }
      }
    }
  }
}

/// One warp's task: heading h, lattice rows rb * 32 + lane, columns ix0 .. ix0 + 32 W - 1.
template <int W>
struct LwTask {
  int h, it, iy, ix0, y0c;
  bool row_ok, any_valid;
  long long i0;          // pose index of (iy, ix0) in this batch (may lie outside [0, count) for the batch's first / last rows)
  uint32_t valid[W];     // poses of this lane's row that belong to the batch
  unsigned live_rows;    // lanes with at least one such pose
};

/// Fills the task of this warp; false when it has no pose of the batch (the whole warp leaves).
template <int W>
__device__ __forceinline__ bool lw_task_setup(const PoseSource& src, long long count, const LwTables& t, const LwGeom& g, int rb, int sb, int h,
                                              int lane, LwTask<W>& k) {
  k.h = h, k.it = g.it_lo + h;
  k.iy = rb * 32 + lane, k.ix0 = sb * 32 * W;
  k.row_ok = rb < g.nb && k.iy < src.ny;
  k.i0 = (static_cast<long long>(k.it) * src.ny + k.iy) * src.nx + k.ix0 - src.first;
  k.any_valid = false;
  const bool whole = k.i0 >= 0 && k.i0 + 32 * W <= count;  // (the usual row: inside the batch)
#pragma unroll
  for (int w = 0; w < W; ++w) {
    uint32_t v = 0u;
    if (k.row_ok) {
      const int in_row = src.nx - (k.ix0 + 32 * w);  // columns of this word inside the lattice row
      if (whole) {
        v = in_row >= 32 ? 0xFFFFFFFFu : (in_row > 0 ? (1u << in_row) - 1u : 0u);
      } else {  // bits j with ix0 + 32 w + j < nx and 0 <= i0 + 32 w + j < count
        const long long lo = max(0ll, -(k.i0 + 32 * w)), hi = min(min(32ll, static_cast<long long>(in_row)), count - (k.i0 + 32 * w));
        if (hi > lo) v = (hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & (0xFFFFFFFFu << lo);
      }
    }
    k.valid[w] = v;
    k.any_valid |= v != 0u;
  }
  k.live_rows = __ballot_sync(0xFFFFFFFFu, k.any_valid);
  k.y0c = k.row_ok ? t.yc0[static_cast<long long>(h) * src.ny + k.iy] : 0;
  return k.live_rows != 0u;
}

/// Is this a SIMPLE task - a heading with a single shape, every pose of the task inside the map?  (Evaluated identically
/// by both sweep kernels: exactly one of them takes the task.)  x_base = map column of vertex 0 for the task's first pose.
template <int W>
__device__ __forceinline__ bool lw_task_simple(const LwTask<W>& k, const MapView& m, const LwTables& t, int& x_base) {
  x_base = 0;
  if (t.simple[k.h] == 0) return false;
  const LwShape* shp = t.shapes + t.shape_base[k.h];
  x_base = k.ix0 + t.xanchor[k.h * kLwClasses];
  const bool x_in = x_base + shp->xb >= 0 && x_base + 32 * W - 1 + shp->x1 < m.size_x;
  const bool y_in = k.y0c + shp->yb >= 0 && k.y0c + shp->yb + shp->rows <= m.size_y;
  return x_in && __all_sync(0xFFFFFFFFu, y_in || !k.any_valid);
}

// Shared-memory tile of the bit planes for one CTA of the simple sweep: [plane k][column][row], rows fastest, so that the
// lanes of a warp (consecutive map rows) read consecutive words and the W + 2 words of a window are a fixed distance apart.
constexpr int kLwTileRows = 288;                // 8 row blocks of 32 + the shape's rows + alignment slack
constexpr int kLwTileColsExtra = 3;             // columns beyond W: the shape's x extent (< 32 cells + the window's two extra words)

/// The tasks of headings with a single shape whose poses all lie inside the map (nearly all of them): the walk and nothing
/// else.  The eight warps of a CTA sweep eight vertically adjacent row blocks of one strip; the plane words they need -
/// (256 + shape rows) map rows x (W + 3) word columns x NK planes - are first copied into shared memory (cp.async,
/// 16 bytes each), so that the walk's loads are LDS with immediate offsets: no address arithmetic per word, no L1 tag
/// traffic, ~25 cycles instead of an L1 / L2 round trip.  Falls back to loading from the planes in global memory when the
/// shape does not fit the tile (wider than 32 cells, more than `nk` planes, more than 128 spans, rows not at cell pitch).
template <int W>
__global__ void __launch_bounds__(kLwWarps * 32) k_lw_sweep_simple(MapView m, const uint32_t* __restrict__ bits, int ypitch, long long plane_words,
                                                                  PoseSource src, long long count, OutView out, LwTables t, LwGeom g, int nk) {
  extern __shared__ uint4 lw_smem[];
  __shared__ LwSpan s_local[kLwSpanCache];
  __shared__ int s_ymin, s_ymax, s_active, s_n_near, s_n_far;
  constexpr int NC = W + kLwTileColsExtra, TH = kLwTileRows;
  uint32_t* tile = reinterpret_cast<uint32_t*>(lw_smem);                  // [nk][NC][TH]
  uint32_t(*s_bits)[2 * W] = reinterpret_cast<uint32_t(*)[2 * W]>(tile + nk * NC * TH);  // [warp * 32 + lane][2 W] (byte outputs only)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sb = blockIdx.y, rb = blockIdx.x * kLwWarps + warp;  // grid: (row block groups, strips, headings)
  const int h = blockIdx.z;
  if (threadIdx.x == 0) s_ymin = 0x7FFFFFFF, s_ymax = -0x7FFFFFFF - 1, s_active = 0, s_n_near = 0, s_n_far = 0;
  __syncthreads();
  const LwHead hd = t.heads[h];  // (CTA-uniform)
  if (hd.simple == 0) return;
  LwTask<W> k;
  bool active = lw_task_setup<W>(src, count, t, g, rb, sb, h, lane, k);
  const int x_base = k.ix0 + hd.anchor0;
  if (active && hd.simple != 2) {  // the same test as lw_task_simple (2: the heading's whole lattice passes it)
    const bool x_in = x_base + hd.xb >= 0 && x_base + 32 * W - 1 + hd.x1 < m.size_x;
    const bool y_in = k.y0c + hd.yb >= 0 && k.y0c + hd.yb + hd.rows <= m.size_y;
    active = x_in && __all_sync(0xFFFFFFFFu, y_in || !k.any_valid);
  }
  if (active) {
    const int lo = __reduce_min_sync(0xFFFFFFFFu, k.any_valid ? k.y0c : 0x7FFFFFFF), hi = __reduce_max_sync(0xFFFFFFFFu, k.any_valid ? k.y0c : -0x7FFFFFFF - 1);
    if (lane == 0) atomicMin(&s_ymin, lo), atomicMax(&s_ymax, hi), s_active = 1;
  }
  __syncthreads();
  if (!s_active) return;
  // (CTA-uniform from here: the heading's one shape, the strip's x_base)
  const int n_spans = hd.n_spans, xb = hd.xb, x1 = hd.x1, yb = hd.yb, rows = hd.rows, kmax = hd.kmax;
  const LwSpan* spans = t.spans + hd.span_first;
  const int xs_base = sb * 32 * W + hd.anchor0 + xb;  // map column of the shape's leftmost cell for the strip's first pose
  const int c_min = xs_base >> 5, xr = xs_base & 31;
  const int y_t0 = (s_ymin + yb) & ~3, rows_needed = s_ymax + yb + rows - y_t0;
  const bool tile_ok = n_spans <= kLwSpanCache && kmax < nk && rows_needed <= TH && ((xr + (x1 - xb)) >> 5) + W + 2 <= NC;
  if (tile_ok) {
    // one (plane, column) segment = rows_needed consecutive words in global memory: a warp copies it 16 bytes per lane
    const int chunks = (rows_needed + 3) >> 2, n_cols = ((xr + (x1 - xb)) >> 5) + W + 2;
    for (int kk = 0; kk <= kmax; ++kk) {
      for (int col = warp; col < n_cols; col += kLwWarps) {
        const uint32_t* gsrc = bits + kk * plane_words + static_cast<long long>(c_min + col) * ypitch + y_t0;
        const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(tile + (kk * NC + col) * TH));
        for (int ch = lane; ch < chunks; ch += 32) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * ch), "l"(gsrc + 4 * ch));
      }
    }
    // the spans, decoded for this strip (everything but the lane's row is CTA-uniform): word offset into the tile, bit
    // offset of the first window, bit offset of the second.  Spans whose second window starts in the same word as the
    // first ("near") go to the front of the list, the others ("far": it starts in the next word) to the back, so that
    // each kind gets a loop without a branch (the verdict is an OR: the order of the spans does not matter).
    for (int sidx = threadIdx.x; sidx < n_spans; sidx += blockDim.x) {
      const LwSpan sp = spans[sidx];
      const int kk = sp.w >> 16, dy = (sp.w & 0xFFFF) - 32768;
      const int tx = xr + sp.x - xb, sh = tx & 31, sh2 = sh + sp.z;
      const int off = (kk * NC + (tx >> 5)) * TH + (dy - yb);
      const bool near = sh2 < 32;
      const int pos = near ? atomicAdd(&s_n_near, 1) : n_spans - 1 - atomicAdd(&s_n_far, 1);
      s_local[pos] = make_int4(off, sh, near ? sh2 : sh2 - 32, 0);
    }
    asm volatile("cp.async.commit_group;" ::);
    asm volatile("cp.async.wait_group 0;" ::);
  }
  __syncthreads();
  if (!active) return;

  uint32_t hit[W];
#pragma unroll
  for (int w = 0; w < W; ++w) hit[w] = 0u;
  const int y_first = __shfl_sync(0xFFFFFFFFu, k.y0c, __ffs(k.live_rows) - 1);  // (every lane takes part: never inside a conditional)
  const int y_lane = k.any_valid ? k.y0c : y_first;  // idle lanes read where a live one reads
  if (tile_ok) {
    const uint32_t* __restrict__ base = tile + (y_lane + yb - y_t0);
    const int n_near = s_n_near;
#pragma unroll 4
    for (int s = 0; s < n_near; ++s) {  // both windows start in the same plane word: W + 1 words
      const LwSpan sp = s_local[s];
      const uint32_t* __restrict__ p = base + sp.x;
      uint32_t L[W + 1];
#pragma unroll
      for (int j = 0; j < W + 1; ++j) L[j] = p[j * TH];
#pragma unroll
      for (int w = 0; w < W; ++w) hit[w] |= __funnelshift_r(L[w], L[w + 1], sp.y) | __funnelshift_r(L[w], L[w + 1], sp.z);
    }
#pragma unroll 4
    for (int s = n_near; s < n_spans; ++s) {  // the second power-of-two window starts in the next word: W + 2 words
      const LwSpan sp = s_local[s];
      const uint32_t* __restrict__ p = base + sp.x;
      uint32_t L[W + 2];
#pragma unroll
      for (int j = 0; j < W + 2; ++j) L[j] = p[j * TH];
#pragma unroll
      for (int w = 0; w < W; ++w) hit[w] |= __funnelshift_r(L[w], L[w + 1], sp.y) | __funnelshift_r(L[w + 1], L[w + 2], sp.z);
    }
  } else {
    lw_walk<W>(bits, ypitch, spans, n_spans, x_base, y_lane, hit);
  }
  uint32_t fr[W], none[W];
#pragma unroll
  for (int w = 0; w < W; ++w) fr[w] = ~hit[w] & k.valid[w], none[w] = 0u;
  lw_store<W>(out, fr, k.valid, none, k.i0, k.any_valid, k.live_rows, true, g.aligned4, warp, lane, s_bits + warp * 32);
}

/// The remaining tasks: several classes per heading (masked walks, every lane the shape of its own row), the map border
/// (bounds-checked walk, a pose whose span leaves the map is not free), statuses, deferrals.
template <int W>
__global__ void __launch_bounds__(kLwWarps * 32) k_lw_sweep_general(MapView m, const uint32_t* __restrict__ bits, int ypitch, long long plane_words,
                                                                   PoseSource src, long long count, OutView out, LwTables t, LwGeom g,
                                                                   unsigned int* __restrict__ todo, unsigned int* __restrict__ todo_count) {
  __shared__ uint32_t s_bits[kLwWarps][32][2 * W];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sb = blockIdx.y, rb = blockIdx.x * kLwWarps + warp;  // grid: (row block groups, strips, headings)
  const int h = t.order[blockIdx.z];
  if (t.simple[h] == 2) return;  // (every task of the heading is k_lw_sweep_simple's)
  LwTask<W> k;
  if (!lw_task_setup<W>(src, count, t, g, rb, sb, h, lane, k)) return;
  int x_simple;
  if (lw_task_simple<W>(k, m, t, x_simple)) return;  // (k_lw_sweep_simple's)
  const int iy = k.iy, ix0 = k.ix0, y0c = k.y0c;
  const bool row_ok = k.row_ok, any_valid = k.any_valid;
  const long long i0 = k.i0;
  const int base = t.shape_base[h];
  const int wcols = (m.size_x + 31) / 32 + 3;
  uint32_t hit[W];
#pragma unroll
  for (int w = 0; w < W; ++w) hit[w] = 0u;
  uint32_t (&valid)[W] = k.valid;
  {
    const int yc = row_ok ? t.ycls[static_cast<long long>(h) * src.ny + iy] : kLwNone;
    int cw[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int ix = ix0 + 32 * w + lane;
      cw[w] = ix < src.nx ? t.xcls[static_cast<long long>(h) * src.nx + ix] : kLwNone;
    }
    const int n_x = t.n_xcls[h], n_y = t.n_ycls[h];
    uint32_t skip[W];  // poses whose result is written elsewhere: status other than OK (stored right away), or left to the general kernel
#pragma unroll
    for (int w = 0; w < W; ++w) skip[w] = 0u;

    auto fail_poses = [&](const uint32_t (&mask)[W], int status) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        uint32_t b = mask[w] & valid[w] & ~skip[w];
        skip[w] |= b;
        while (b) {
          const int j = __ffs(b) - 1;
          b &= b - 1u;
          store_failure(out, i0 + 32 * w + j, status);
        }
      }
    };
    auto defer_poses = [&](const uint32_t (&mask)[W]) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        uint32_t b = mask[w] & valid[w] & ~skip[w];
        skip[w] |= b;
        while (b) {
          const int j = __ffs(b) - 1;
          b &= b - 1u;
          push_todo(todo, todo_count, i0 + 32 * w + j);
        }
      }
    };

    {  // coordinates out of range (status), class tables full (general kernel)
      uint32_t bad[W], over[W];
      bool any = false;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        bad[w] = __ballot_sync(0xFFFFFFFFu, cw[w] == kLwBad) | (yc == kLwBad ? 0xFFFFFFFFu : 0u);
        over[w] = __ballot_sync(0xFFFFFFFFu, cw[w] == kLwOverflow) | ((yc == kLwOverflow || base < 0) ? 0xFFFFFFFFu : 0u);
        any |= (bad[w] | over[w]) != 0u;
      }
      if (any) {
        fail_poses(bad, ST_COORD_RANGE);
        defer_poses(over);
      }
    }

    if (base >= 0) {
      const bool has = any_valid && yc < kLwClasses;  // this lane's row has a shape per x class
      for (int q = 0; q < n_x; ++q) {
        uint32_t M[W];
        uint32_t any_m = 0u;
#pragma unroll
        for (int w = 0; w < W; ++w) {
          M[w] = __ballot_sync(0xFFFFFFFFu, cw[w] == q);
          any_m |= M[w];
        }
        if (any_m == 0u) continue;
        const int xq = ix0 + t.xanchor[h * kLwClasses + q];
        // every lane walks the shape of ITS row's y class
        const LwShape* shp = t.shapes + base + q * n_y + (has ? yc : 0);
        const int status = has ? shp->status : ST_OK;
        if (has && status == kLwUnfit) defer_poses(M);
        else if (has && status != ST_OK) fail_poses(M, status);  // open outline (std::logic_error upstream) / more than 16 ranges on a row
        const bool walk = has && status == ST_OK;
        const int n_sp = walk ? shp->n_spans : 0;
        const int n_max = __reduce_max_sync(0xFFFFFFFFu, n_sp);
        const LwSpan* spans = t.spans + (walk ? shp->span_first : 0);
        // every needed word inside the planes?  (bounding box of the vertices: sufficient, the spans lie inside it)
        const bool inside = !walk || (xq + shp->xb >= 0 && xq + 32 * W - 1 + shp->x1 < m.size_x && y0c + shp->yb >= 0 && y0c + shp->yb + shp->rows <= m.size_y);
        uint32_t acc[W];
#pragma unroll
        for (int w = 0; w < W; ++w) acc[w] = 0u;
        if (__all_sync(0xFFFFFFFFu, inside)) lw_walk_lanes<W, false>(bits, ypitch, plane_words, wcols, m.size_x, m.size_y, spans, n_sp, n_max, xq, y0c, acc);
        else lw_walk_lanes<W, true>(bits, ypitch, plane_words, wcols, m.size_x, m.size_y, spans, n_sp, n_max, xq, y0c, acc);
#pragma unroll
        for (int w = 0; w < W; ++w) hit[w] |= acc[w] & M[w];
      }
    }
    uint32_t fr[W];
    bool special = false;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      fr[w] = ~hit[w] & valid[w] & ~skip[w];
      special |= skip[w] != 0u;
    }
    const bool plain = !__any_sync(0xFFFFFFFFu, special);
    lw_store<W>(out, fr, valid, skip, i0, any_valid, k.live_rows, plain, g.aligned4, warp, lane, s_bits[warp]);
  }
}

}  // namespace cvr

namespace cvr {

/// Tables the value folds of a lattice sweep read: accumulate_cost the per-row exclusive prefix words of k_cost_prefix
/// (word x = NO_INFORMATION cells before x, mod 256, << 24 | LETHAL cells before x, mod 256, << 16 | sum of the costs
/// before x, mod 2^16: a window of up to 32 cells is two loads, every field exact as a difference); max_cost the sliding-maximum byte planes of k_max_planes
/// (plane k, byte x of row y = max of the cells x .. x + 2^k - 1; plane 0 is a copy of the costmap's rows, so that a span's
/// plane is an offset, not a choice of pointers).
struct LwValueTables {
  const uint32_t* __restrict__ prefix;
  int ppitch;
  const uint8_t* __restrict__ maxp;  // planes 0..5, plane k at maxp + k * max_plane_bytes, row pitch = the map's (plane 0: a copy of the rows)
  long long max_plane_bytes;
};

/// accumulate_cost / max_cost(area_generator(res, pose * polygon - origin), map) over a lattice at cell pitch, from the
/// class / shape tables of k_lw_prepare: WARP = 64 x-adjacent poses of one lattice row (two per lane), the spans of their
/// shape are uniform loads, a span costs every pose two table reads - consecutive poses read consecutive words / bytes.
/// Sums: the first window (the spans are in generator order: rows ascending, x ascending) that holds a LETHAL or a
/// NO_INFORMATION cell decides (the reference's "first negative value in generator order", impl/costmap.hpp:50-63) - from
/// the two counts alone unless it holds both kinds, then it is walked byte by byte for the first one.  Poses whose bounding box leaves the map, shapes beyond the tables: deferred to the per-pose kernels.
/// -1 / -2: the first of `width` cells that is LETHAL / NO_INFORMATION (the window holds both kinds)
__device__ __noinline__ int lw_first_sentinel(const uint8_t* __restrict__ cells, int width) {
  for (int c = 0; c < width; ++c) {
    const int v = __ldg(cells + c);
    if (v >= kLethal) return v == kLethal ? -1 : -2;
  }
  return 0;
}

template <int OP>
__global__ void __launch_bounds__(256) k_lw_values(MapView m, LwValueTables vt, PoseSource src, long long count, OutView out, LwTables t, LwGeom g,
                                                  unsigned int* __restrict__ todo, unsigned int* __restrict__ todo_count) {
  constexpr unsigned kAll = 0xFFFFFFFFu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.z, it = g.it_lo + h;
  const int iy = blockIdx.x * 8 + warp;
  if (iy >= src.ny) return;
  const long long row_i = (static_cast<long long>(it) * src.ny + iy) * src.nx - src.first;  // pose index of column 0 in this batch
  const int ix_a = blockIdx.y * 64 + lane, ix_b = ix_a + 32;
  const long long i_a = row_i + ix_a, i_b = row_i + ix_b;
  const bool valid_a = ix_a < src.nx && i_a >= 0 && i_a < count, valid_b = ix_b < src.nx && i_b >= 0 && i_b < count;
  if (!__any_sync(kAll, valid_a || valid_b)) return;
  const int yc = t.ycls[static_cast<long long>(h) * src.ny + iy], y0c = t.yc0[static_cast<long long>(h) * src.ny + iy];
  const int xc_a = valid_a ? t.xcls[static_cast<long long>(h) * src.nx + ix_a] : kLwNone;
  const int xc_b = valid_b ? t.xcls[static_cast<long long>(h) * src.nx + ix_b] : kLwNone;
  const int base = t.shape_base[h], n_y = t.n_ycls[h];
  // coordinates out of range (status); class tables full or shape pool exhausted (per-pose kernels)
  if (valid_a && (xc_a == kLwBad || yc == kLwBad)) store_failure(out, i_a, ST_COORD_RANGE);
  else if (valid_a && (xc_a == kLwOverflow || yc == kLwOverflow || base < 0)) push_todo(todo, todo_count, i_a);
  if (valid_b && (xc_b == kLwBad || yc == kLwBad)) store_failure(out, i_b, ST_COORD_RANGE);
  else if (valid_b && (xc_b == kLwOverflow || yc == kLwOverflow || base < 0)) push_todo(todo, todo_count, i_b);
  const bool live_a = valid_a && xc_a < kLwClasses && yc < kLwClasses && base >= 0;
  const bool live_b = valid_b && xc_b < kLwClasses && yc < kLwClasses && base >= 0;
  unsigned pend_a = __ballot_sync(kAll, live_a), pend_b = __ballot_sync(kAll, live_b);
  while (pend_a | pend_b) {  // one pass per x class present among the 64 poses (nearly always one)
    const int leader = pend_a ? __ffs(pend_a) - 1 : __ffs(pend_b) - 1;
    const int qa = __shfl_sync(kAll, xc_a, leader), qb = __shfl_sync(kAll, xc_b, leader);
    const int q = pend_a ? qa : qb;
    const bool in_a = live_a && xc_a == q, in_b = live_b && xc_b == q;
    pend_a &= ~__ballot_sync(kAll, in_a), pend_b &= ~__ballot_sync(kAll, in_b);
    const LwShape* shp = t.shapes + base + q * n_y + yc;
    const int status = shp->status;
    if (status == kLwUnfit) {
      if (in_a) push_todo(todo, todo_count, i_a);
      if (in_b) push_todo(todo, todo_count, i_b);
      continue;
    }
    if (status != ST_OK) {  // std::logic_error upstream / more than 16 ranges on a row
      if (in_a) store_failure(out, i_a, status);
      if (in_b) store_failure(out, i_b, status);
      continue;
    }
    const int X_a = ix_a + t.xanchor[h * kLwClasses + q], X_b = X_a + 32;  // map column of vertex 0
    const bool y_in = y0c + shp->yb >= 0 && y0c + shp->yb + shp->rows <= m.size_y;
    const bool go_a = in_a && y_in && X_a + shp->xb >= 0 && X_a + shp->x1 < m.size_x;
    const bool go_b = in_b && y_in && X_b + shp->xb >= 0 && X_b + shp->x1 < m.size_x;
    if (in_a && !go_a) push_todo(todo, todo_count, i_a);  // at the map border: the per-pose kernels know the rules
    if (in_b && !go_b) push_todo(todo, todo_count, i_b);
    const LwSpan* __restrict__ spans = t.spans + shp->span_first;
    const int n_spans = shp->n_spans;
    // lanes without a pose in this pass read what a lane with one reads (in range; their results are dropped)
    const unsigned ga = __ballot_sync(kAll, go_a), gb = __ballot_sync(kAll, go_b);
    if ((ga | gb) == 0u) continue;
    const int donor = ga ? __ffs(ga) - 1 : __ffs(gb) - 1;
    const int donor_a = __shfl_sync(kAll, X_a, donor), donor_b = __shfl_sync(kAll, X_b, donor);
    const int Xd = ga ? donor_a : donor_b;
    const int Xa = go_a ? X_a : Xd, Xb = go_b ? X_b : Xd;
    if (OP == OP_ACCUMULATE) {
      // (32-bit indices into the tables: the host takes this path only while rows x pitch stays below 2^31)
      uint32_t sum_a = 0u, sum_b = 0u;
      int early_a = 0, early_b = 0;
      for (int s = 0; s < n_spans; ++s) {
        const LwSpan sp = __ldg(spans + s);
        const int dy = (sp.w & 0xFFFF) - 32768, width = sp.z + (1 << (sp.w >> 16));
        const int o = (y0c + dy) * vt.ppitch + sp.x;
        const int oa = o + Xa, ob = o + Xb;
        const uint32_t a0 = __ldg(vt.prefix + oa), a1 = __ldg(vt.prefix + oa + width);
        const uint32_t b0 = __ldg(vt.prefix + ob), b1 = __ldg(vt.prefix + ob + width);
        sum_a += (a1 - a0) & 0xFFFFu, sum_b += (b1 - b0) & 0xFFFFu;  // (meaningless once a sentinel was met: `early` decides then)
        // (the two sentinel counts live in the upper halves: a window of <= 32 cells holds a sentinel iff they differ)
        const bool sa = ((a1 ^ a0) >> 16) != 0u && early_a == 0 && go_a, sb = ((b1 ^ b0) >> 16) != 0u && early_b == 0 && go_b;
        if (sa || sb) {
          if (sa) {
            const uint32_t n_l = ((a1 >> 16) - (a0 >> 16)) & 0xFFu, n_n = ((a1 >> 24) - (a0 >> 24)) & 0xFFu;
            early_a = n_n == 0u ? -1 : (n_l == 0u ? -2 : lw_first_sentinel(m.data + static_cast<size_t>(y0c + dy) * m.pitch + sp.x + X_a, width));
          }
          if (sb) {
            const uint32_t n_l = ((b1 >> 16) - (b0 >> 16)) & 0xFFu, n_n = ((b1 >> 24) - (b0 >> 24)) & 0xFFu;
            early_b = n_n == 0u ? -1 : (n_l == 0u ? -2 : lw_first_sentinel(m.data + static_cast<size_t>(y0c + dy) * m.pitch + sp.x + X_b, width));
          }
        }
      }
      Fold<OP_ACCUMULATE> fa, fb;
      fa.sum = sum_a, fa.early = early_a, fb.sum = sum_b, fb.early = early_b;
      if (go_a) {
        fa.store(out, i_a);
        if (out.status) out.status[i_a] = ST_OK;
      }
      if (go_b) {
        fb.store(out, i_b);
        if (out.status) out.status[i_b] = ST_OK;
      }
    } else {
      uint32_t best_a = 0u, best_b = 0u;
      const int plane_bytes = static_cast<int>(vt.max_plane_bytes);
      for (int s = 0; s < n_spans; ++s) {
        const LwSpan sp = __ldg(spans + s);
        const int dy = (sp.w & 0xFFFF) - 32768, lg = sp.w >> 16;
        const int o = lg * plane_bytes + (y0c + dy) * m.pitch + sp.x;  // (32-bit: the host takes this path while 6 planes stay below 2^31 bytes)
        const int oa = o + Xa, ob = o + Xb;
        const uint32_t va = max(static_cast<uint32_t>(__ldg(vt.maxp + oa)), static_cast<uint32_t>(__ldg(vt.maxp + oa + sp.z)));
        const uint32_t vb = max(static_cast<uint32_t>(__ldg(vt.maxp + ob)), static_cast<uint32_t>(__ldg(vt.maxp + ob + sp.z)));
        best_a = max(best_a, va), best_b = max(best_b, vb);
      }
      Fold<OP_MAX> fa, fb;
      fa.best4 = best_a, fb.best4 = best_b;
      if (go_a) {
        fa.store(out, i_a);
        if (out.status) out.status[i_a] = ST_OK;
      }
      if (go_b) {
        fb.store(out, i_b);
        if (out.status) out.status[i_b] = ST_OK;
      }
    }
  }
}

/// Sliding-maximum planes 0..5 of rows [row_lo, row_hi] (grid.y = row, one thread per 4 cells): the 4 + 32 cells after a
/// word's first cell are doubled in registers, M_{k+1}[x] = max(M_k[x], M_k[x + 2^k]) byte-wise (__vmaxu4).
__global__ void __launch_bounds__(256) k_max_planes(MapView m, uint8_t* __restrict__ maxp, long long plane_bytes, int row_lo) {
  const int words = m.pitch >> 2;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= words) return;
  const int y = row_lo + blockIdx.y;
  const uint32_t* __restrict__ row = reinterpret_cast<const uint32_t*>(m.data + static_cast<size_t>(y) * m.pitch);
  uint32_t R[10];
#pragma unroll
  for (int d = 0; d < 9; ++d) R[d] = j + d < words ? __ldg(row + j + d) : 0u;
  R[9] = 0u;
  uint32_t* __restrict__ dst = reinterpret_cast<uint32_t*>(maxp + static_cast<size_t>(y) * m.pitch) + j;
  const long long plane_words = plane_bytes >> 2;
  dst[0] = R[0];
#pragma unroll
  for (int d = 0; d < 9; ++d) R[d] = __vmaxu4(R[d], __funnelshift_r(R[d], R[d + 1], 8));
  dst[plane_words] = R[0];
#pragma unroll
  for (int d = 0; d < 8; ++d) R[d] = __vmaxu4(R[d], __funnelshift_r(R[d], R[d + 1], 16));
  dst[2 * plane_words] = R[0];
#pragma unroll
  for (int d = 0; d < 7; ++d) R[d] = __vmaxu4(R[d], R[d + 1]);
  dst[3 * plane_words] = R[0];
#pragma unroll
  for (int d = 0; d < 5; ++d) R[d] = __vmaxu4(R[d], R[d + 2]);
  dst[4 * plane_words] = R[0];
  dst[5 * plane_words] = __vmaxu4(R[0], R[4]);
}

}  // namespace cvr
