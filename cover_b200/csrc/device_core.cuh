// device_core.cuh — device-side building blocks of the batched footprint checker (sm_100a).
//
// Everything here is integer/byte work plus a handful of FP64 operations per vertex; there is no
// contraction, so no tensor cores.  The pieces mirror the reference's FUNCTIONS (not its code
// structure): a pose transform + discretisation, a Bresenham outline walker that calls a sink per cell,
// a run detector that turns the cell stream into the reference's scan-line ranges, two row stores
// (a packed 2-range table in shared memory for y-monotone outlines, sorted lists in global scratch for
// everything else) and byte-parallel row scans (4 cells per 32-bit word).
//
// Reference semantics restated (paths relative to the reference tree):
//   to_cell            src/cover/base.hpp:50-57 + costmap.cpp:71-85 (path A) / footprints.cpp:134-135 (path B)
//   walk_ray           src/cover/generators.cpp:30-52, generators.hpp:117-128
//   walk_outline       src/cover/generators.cpp:54-93
//   RunDetector        src/cover/generators.cpp:189-232 (runs, cusps, wrap-around)
//   *Store::finalize   src/cover/generators.cpp:236-255 (sort by min, pair, merge)
//   Fold<OP>           src/cover/costmap.hpp:81-133, costmap.cpp:35-46, impl/costmap.hpp:23-63
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace cvr {

constexpr int kLethal = 254;
constexpr int kNoInformation = 255;
constexpr int kInscribed = 253;
constexpr double kCoordLimit = 4194304.0;  // 2^22
constexpr int kMaxRanges = 16;             // std::sort is insertion-stable up to 16 elements (libstdc++)

enum Op : int { OP_IS_FREE = 0, OP_ACCUMULATE = 1, OP_MAX = 2 };
enum Status : int { ST_OK = 0, ST_LOGIC_ERROR = 1, ST_ROW_COMPLEXITY = 2, ST_COORD_RANGE = 3 };
enum PoseFormat : int { POSE_CS = 0, POSE_LATTICE = 1, POSE_NONE = 2 };
enum XformKind : int { XF_NONE = 0, XF_PATH_A = 1, XF_PATH_B = 2 };

/// Device copy of the costmap: rows padded to a multiple of 16 bytes so that any aligned word of a row
/// may be read; the allocation is 256-byte aligned.
struct MapView {
  const uint8_t* __restrict__ data;
  int size_x, size_y, pitch;
  double inv_res, origin_x, origin_y;
  // Optional bit planes of the current image (nullptr: not built), see k_lethal_planes in cover_b200.cu: plane k,
  // bit x of row y = "a cell of x .. x + 2^k - 1 is LETHAL" (`planes`) / "... is INSCRIBED or LETHAL" (`planes_inf`),
  // k = 0..5, stored column-major: word (c = x / 32, y) at k * plane_words + c * ypitch + y.
  const uint32_t* __restrict__ planes;
  const uint32_t* __restrict__ planes_inf;
  int ypitch;
  long long plane_words;
};

struct Pose {
  double c, s, tx, ty;  // linear part [c -s; s c] (Eigen::Rotation2D::toRotationMatrix)
  int ox, oy;           // cell offset added after discretisation (with_offset, costmap.hpp:135-155)
};

struct PoseSource {
  int format;
  long long first;         // lattice: global index of this batch's first pose
  const double* cs;        // POSE_CS: count x 4
  const double* theta_cs;  // POSE_LATTICE: ntheta x 2 (device copy)
  double x0, dx, y0, dy;
  int nx, ny;
  int off_x, off_y;        // cover_poses.cell_offset
};

struct OutView {
  uint8_t* is_free;
  double* cost;
  uint8_t* status;
  uint32_t* free_bits;          // optional bit-packed verdicts (pre-zeroed): bit (i & 31) of word (i >> 5)
  unsigned long long* not_ok;   // optional counter of poses whose status is not ST_OK (pre-zeroed)
};

/// Verdict of pose i: byte array and/or bit-packed word (bits are only ever set, the buffer starts zeroed).
__device__ __forceinline__ void emit_free(const OutView& o, long long i, bool is_free) {
  if (o.is_free) o.is_free[i] = is_free ? 1 : 0;
  if (o.free_bits && is_free) atomicOr(o.free_bits + (i >> 5), 1u << (i & 31));
}

__device__ __forceinline__ Pose load_pose(const PoseSource& src, long long i) {
  Pose T;
  T.ox = src.off_x, T.oy = src.off_y;
  if (src.format == POSE_CS) {
    const double2* q = reinterpret_cast<const double2*>(src.cs) + 2 * i;
    const double2 a = __ldg(q), b = __ldg(q + 1);
    T.c = a.x, T.s = a.y, T.tx = b.x, T.ty = b.y;
  } else if (src.format == POSE_LATTICE) {
    const long long p = src.first + i;
    const long long plane = static_cast<long long>(src.nx) * src.ny;
    const int it = static_cast<int>(p / plane);
    const int r = static_cast<int>(p - it * plane);
    const int iy = r / src.nx, ix = r - iy * src.nx;
    // t = x0 + ix*dx: multiply, then add, each rounded (no FMA) — the ABI's stated rule.
    T.tx = __dadd_rn(src.x0, __dmul_rn(static_cast<double>(ix), src.dx));
    T.ty = __dadd_rn(src.y0, __dmul_rn(static_cast<double>(iy), src.dy));
    const double2 cs = __ldg(reinterpret_cast<const double2*>(src.theta_cs) + it);
    T.c = cs.x, T.s = cs.y;
  } else {
    T.c = 1, T.s = 0, T.tx = 0, T.ty = 0;
  }
  return T;
}

/// World vertex -> map cell.  Every FP operation is an explicitly rounded intrinsic so that nvcc can
/// never contract a multiply-add: the CPU reference (x86-64 SSE2, no FMA) rounds after each step.
template <int XF>
__device__ __forceinline__ bool to_cell(const Pose& T, double px, double py, const MapView& m, int& cx, int& cy) {
  double qx, qy;
  if (XF == XF_NONE) {  // costmap.cpp:73-74
    qx = __dsub_rn(px, m.origin_x);
    qy = __dsub_rn(py, m.origin_y);
  } else {
    // Eigen Transform * Matrix: res = t; res += (l_i0*p_x) + (l_i1*p_y)
    const double lx = __dadd_rn(__dmul_rn(T.c, px), __dmul_rn(-T.s, py));
    const double ly = __dadd_rn(__dmul_rn(T.s, px), __dmul_rn(T.c, py));
    qx = __dadd_rn(T.tx, lx);
    qy = __dadd_rn(T.ty, ly);
    if (XF == XF_PATH_A) {  // costmap.cpp:81-83: origin subtracted AFTER the transform
      qx = __dsub_rn(qx, m.origin_x);
      qy = __dsub_rn(qy, m.origin_y);
    }  // XF_PATH_B: the caller folded (-origin) into T.tx/T.ty (footprints.cpp:134-135)
  }
  const double sx = __dmul_rn(qx, m.inv_res), sy = __dmul_rn(qy, m.inv_res);
  cx = __double2int_rz(sx) + T.ox;  // base.hpp:56: cast<int>() truncates toward zero; then the offset (costmap.hpp:150-153)
  cy = __double2int_rz(sy) + T.oy;
  return (fabs(sx) < kCoordLimit) && (fabs(sy) < kCoordLimit);  // false also for NaN
}

// ---------------------------------------------------------------------------------------------------
// Byte-parallel row scans.  A row pointer is 16-byte aligned; x indexes bytes of the row.
// ---------------------------------------------------------------------------------------------------

/// Non-zero iff some byte of v is zero (exact as an existence test).
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v) { return (v - 0x01010101u) & ~v & 0x80808080u; }

/// Bytes of word `w` (covering x = 4w..4w+3) that lie inside [lo, hi], as a 0xFF-per-byte mask.
__device__ __forceinline__ uint32_t keep_mask(int w, int lo, int hi) {
  uint32_t k = 0xFFFFFFFFu;
  const int a = lo - 4 * w, b = 4 * w + 3 - hi;
  if (a > 0) k &= 0xFFFFFFFFu << (8 * a);
  if (b > 0) k &= 0xFFFFFFFFu >> (8 * b);
  return k;
}

/// true iff a cell of [lo, hi] of row y is marked in the bit planes `planes` (all cells inside the map): the OR of two
/// overlapping power-of-two windows for spans of up to 32 cells - two loads, whatever the width - and of 32-cell
/// windows beyond that.
__device__ __forceinline__ bool span_any_bits(const uint32_t* __restrict__ planes, int ypitch, long long plane_words, int y, int lo, int hi) {
  const int w = hi - lo + 1;
  if (w <= 32) {
    const int k = 31 - __clz(w);
    const uint32_t* __restrict__ p = planes + k * plane_words + y;
    const int x2 = hi - (1 << k) + 1;
    const uint32_t a = __ldg(p + static_cast<long long>(lo >> 5) * ypitch) >> (lo & 31);
    const uint32_t b = __ldg(p + static_cast<long long>(x2 >> 5) * ypitch) >> (x2 & 31);
    return ((a | b) & 1u) != 0u;
  }
  const uint32_t* __restrict__ p = planes + 5 * plane_words + y;
  const int x2 = hi - 31;
  uint32_t acc = __ldg(p + static_cast<long long>(x2 >> 5) * ypitch) >> (x2 & 31);
  for (int x = lo; x < x2; x += 32) acc |= __ldg(p + static_cast<long long>(x >> 5) * ypitch) >> (x & 31);
  return (acc & 1u) != 0u;
}

/// true iff some cell of [lo,hi] equals the byte replicated in pat4.  The words of the span are loaded in
/// batches of 8 independent (predicated) loads before any of them is tested: with scattered poses every load
/// is an L2 round trip, so the memory-level parallelism has to come from inside the thread.
__device__ __forceinline__ bool row_any_eq(const uint8_t* __restrict__ row, int lo, int hi, uint32_t pat4) {
  const uint32_t* __restrict__ p = reinterpret_cast<const uint32_t*>(row) + (lo >> 2);
  const int last = (hi >> 2) - (lo >> 2);  // index of the last word of the span
  const uint32_t drop_first = ~(0xFFFFFFFFu << (8 * (lo & 3)));        // bytes before lo
  const uint32_t drop_last = ~(0xFFFFFFFFu >> (8 * (3 - (hi & 3))));   // bytes after hi
  uint32_t acc = 0;
  for (int base = 0; base <= last; base += 8) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (base + j <= last) ? __ldg(p + base + j) : ~pat4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t t = v[j] ^ pat4;  // a zero byte marks a match
      if (base + j == 0) t |= drop_first;
      if (base + j == last) t |= drop_last;
      acc |= zero_bytes(t);
    }
  }
  return acc != 0;
}

/// true iff some cell of [lo,hi] is INSCRIBED_INFLATED (253) or LETHAL (254): is_lethal_or_inflated (costmap.hpp:96-109).
__device__ __forceinline__ bool row_any_inflated(const uint8_t* __restrict__ row, int lo, int hi) {
  const uint32_t* __restrict__ p = reinterpret_cast<const uint32_t*>(row) + (lo >> 2);
  const int last = (hi >> 2) - (lo >> 2);
  const uint32_t keep_first = 0xFFFFFFFFu << (8 * (lo & 3)), keep_last = 0xFFFFFFFFu >> (8 * (3 - (hi & 3)));
  uint32_t acc = 0;
  for (int base = 0; base <= last; base += 8) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (base + j <= last) ? __ldg(p + base + j) : 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t t = v[j];  // excluded bytes become 0: neither 253 nor 254
      if (base + j == 0) t &= keep_first;
      if (base + j == last) t &= keep_last;
      acc |= zero_bytes(t ^ 0xFEFEFEFEu) | zero_bytes(t ^ 0xFDFDFDFDu);
    }
  }
  return acc != 0;
}

/// accumulate_cost over [lo,hi] in x order: adds costs to `sum`; returns 0, or -1 / -2 for the first
/// LETHAL / NO_INFORMATION cell (costmap.cpp:35-46).  Loads are issued 8 words at a time.
__device__ __forceinline__ int row_accumulate(const uint8_t* __restrict__ row, int lo, int hi, unsigned long long& sum) {
  const uint32_t* __restrict__ p = reinterpret_cast<const uint32_t*>(row) + (lo >> 2);
  const int last = (hi >> 2) - (lo >> 2);
  const uint32_t keep_first = 0xFFFFFFFFu << (8 * (lo & 3)), keep_last = 0xFFFFFFFFu >> (8 * (3 - (hi & 3)));
  unsigned int part = 0;
  for (int base = 0; base <= last; base += 8) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (base + j <= last) ? __ldg(p + base + j) : 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t u = v[j];
      if (base + j == 0) u &= keep_first;
      if (base + j == last) u &= keep_last;
      if (zero_bytes(~(u | 0x01010101u))) {  // some byte is 254 or 255: resolve the first one in x order
        for (int b = 0; b < 4; ++b) {
          const uint32_t c = (u >> (8 * b)) & 0xFFu;
          if (c == kLethal) return -1;
          if (c == kNoInformation) return -2;
          part += c;  // excluded bytes are zero
        }
      } else {
        part = __dp4a(u, 0x01010101u, part);
      }
    }
    if (part > 0x7FFFFFFFu) {
      sum += part;
      part = 0;
    }
  }
  sum += part;
  return 0;
}

__device__ __forceinline__ uint32_t row_max(const uint8_t* __restrict__ row, int lo, int hi, uint32_t best4) {
  const uint32_t* __restrict__ p = reinterpret_cast<const uint32_t*>(row) + (lo >> 2);
  const int last = (hi >> 2) - (lo >> 2);
  const uint32_t keep_first = 0xFFFFFFFFu << (8 * (lo & 3)), keep_last = 0xFFFFFFFFu >> (8 * (3 - (hi & 3)));
  for (int base = 0; base <= last; base += 8) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (base + j <= last) ? __ldg(p + base + j) : 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t u = v[j];
      if (base + j == 0) u &= keep_first;
      if (base + j == last) u &= keep_last;
      best4 = __vmaxu4(best4, u);
    }
  }
  return best4;
}

// ---------------------------------------------------------------------------------------------------
// Folds: what is_free / accumulate_cost / max do with cells (outline, ray, cell-list paths) and with
// whole row spans (area path).  `cell`/`span` return false when the walk may stop.
// ---------------------------------------------------------------------------------------------------

template <int OP>
struct Fold;

template <>
struct Fold<OP_IS_FREE> {  // std::none_of(is_lethal): out of the map counts as lethal (costmap.hpp:81-91)
  bool hit = false;
  /// one cell whose cost byte was already fetched (`inside` = the cell is inside the map)
  __device__ __forceinline__ bool consume(bool inside, int v) {
    hit = !inside || v == kLethal;
    return !hit;
  }
  __device__ __forceinline__ bool cell(const MapView& m, int x, int y) {
    if (static_cast<unsigned>(x) >= static_cast<unsigned>(m.size_x) || static_cast<unsigned>(y) >= static_cast<unsigned>(m.size_y) ||
        __ldg(m.data + static_cast<size_t>(y) * m.pitch + x) == kLethal) {
      hit = true;
      return false;
    }
    return true;
  }
  __device__ __forceinline__ bool span(const MapView& m, int y, int lo, int hi) {
    if (static_cast<unsigned>(y) >= static_cast<unsigned>(m.size_y) || lo < 0 || hi >= m.size_x ||
        (m.planes ? span_any_bits(m.planes, m.ypitch, m.plane_words, y, lo, hi)
                  : row_any_eq(m.data + static_cast<size_t>(y) * m.pitch, lo, hi, 0xFEFEFEFEu))) {
      hit = true;
      return false;
    }
    return true;
  }
  __device__ __forceinline__ void store(const OutView& o, long long i) const { emit_free(o, i, !hit); }
};

template <>
struct Fold<OP_ACCUMULATE> {  // impl/costmap.hpp:50-63: first negative value in generator order wins
  unsigned long long sum = 0;
  int early = 0;
  __device__ __forceinline__ bool consume(bool inside, int v) {
    if (!inside) {
      early = -3;
    } else if (v >= kLethal) {
      early = v == kLethal ? -1 : -2;
    } else {
      sum += v;
    }
    return early == 0;
  }
  __device__ __forceinline__ bool cell(const MapView& m, int x, int y) {
    if (static_cast<unsigned>(x) >= static_cast<unsigned>(m.size_x) || static_cast<unsigned>(y) >= static_cast<unsigned>(m.size_y)) {
      early = -3;
      return false;
    }
    const int v = __ldg(m.data + static_cast<size_t>(y) * m.pitch + x);
    if (v >= kLethal) {
      early = v == kLethal ? -1 : -2;
      return false;
    }
    sum += v;
    return true;
  }
  __device__ __forceinline__ bool span(const MapView& m, int y, int lo, int hi) {
    if (static_cast<unsigned>(y) >= static_cast<unsigned>(m.size_y) || lo < 0 || lo >= m.size_x) {
      early = -3;  // the first cell of the span is already outside
      return false;
    }
    const int hc = min(hi, m.size_x - 1);
    early = row_accumulate(m.data + static_cast<size_t>(y) * m.pitch, lo, hc, sum);
    if (early == 0 && hi >= m.size_x) early = -3;
    return early == 0;
  }
  __device__ __forceinline__ void store(const OutView& o, long long i) const {
    const double c = early ? static_cast<double>(early) : static_cast<double>(sum);
    if (o.cost) o.cost[i] = c;
    emit_free(o, i, early == 0);
  }
};

template <>
struct Fold<OP_MAX> {  // extension: max of get_cell_cost; any cell outside the map -> -3
  uint32_t best4 = 0;
  bool outside = false;
  __device__ __forceinline__ bool consume(bool inside, int v) {
    if (!inside) outside = true;
    else best4 = max(best4, static_cast<uint32_t>(v));
    return inside;
  }
  __device__ __forceinline__ bool cell(const MapView& m, int x, int y) {
    if (static_cast<unsigned>(x) >= static_cast<unsigned>(m.size_x) || static_cast<unsigned>(y) >= static_cast<unsigned>(m.size_y)) {
      outside = true;
      return false;
    }
    best4 = max(best4, static_cast<uint32_t>(__ldg(m.data + static_cast<size_t>(y) * m.pitch + x)));
    return true;
  }
  __device__ __forceinline__ bool span(const MapView& m, int y, int lo, int hi) {
    if (static_cast<unsigned>(y) >= static_cast<unsigned>(m.size_y) || lo < 0 || hi >= m.size_x) {
      outside = true;
      return false;
    }
    best4 = row_max(m.data + static_cast<size_t>(y) * m.pitch, lo, hi, best4);
    return true;
  }
  __device__ __forceinline__ int value() const {
    const uint32_t a = max(best4 & 0xFFu, (best4 >> 8) & 0xFFu), b = max((best4 >> 16) & 0xFFu, best4 >> 24);
    return static_cast<int>(max(a, b));
  }
  __device__ __forceinline__ void store(const OutView& o, long long i) const {
    const int v = value();
    if (o.cost) o.cost[i] = outside ? -3.0 : static_cast<double>(v);
    emit_free(o, i, !outside && v != kLethal);
  }
};

/// Deferred-pose list: `counter[0]` counts the poses pushed, `counter[1]` is the list's capacity.  A push beyond the
/// capacity only counts; the kernel that drains the list then sees counter[0] > counter[1] and takes EVERY pose of the
/// batch instead (results are exact either way - the list is an optimisation, its size a guess).
__device__ __forceinline__ void push_todo(unsigned int* __restrict__ list, unsigned int* __restrict__ counter, long long i) {
  const unsigned int at = atomicAdd(counter, 1u);
  if (at < counter[1]) list[at] = static_cast<unsigned int>(i);
}
/// Number of items a drain kernel has to take and whether they are list entries (else: poses 0 .. count - 1).
__device__ __forceinline__ long long todo_items(const unsigned int* __restrict__ counter, long long count, bool& listed) {
  const unsigned int n = counter[0], cap = counter[1];
  listed = n <= cap;
  return listed ? static_cast<long long>(n) : count;
}

__device__ __forceinline__ void store_failure(const OutView& o, long long i, int status) {
  if (o.is_free) o.is_free[i] = 0;
  if (o.cost) o.cost[i] = __longlong_as_double(0x7FF8000000000000LL);  // quiet NaN
  if (o.status) o.status[i] = static_cast<uint8_t>(status);
  if (o.not_ok) atomicAdd(o.not_ok, 1ull);
}

// ---------------------------------------------------------------------------------------------------
// Bresenham walkers.
// ---------------------------------------------------------------------------------------------------

/// ray_generator: `size = max(|dx|,|dy|)` cells from (bx,by), end excluded; ties make x the major axis.
template <class Sink>
__device__ __forceinline__ bool walk_ray(int bx, int by, int ex, int ey, Sink& sink) {
  const int dx = ex - bx, dy = ey - by;
  const int adx = abs(dx), ady = abs(dy);
  const int sx = (dx > 0) - (dx < 0), sy = (dy > 0) - (dy < 0);
  const bool x_major = adx >= ady;
  const int size = x_major ? adx : ady, add = x_major ? ady : adx;
  const int major_x = x_major ? sx : 0, major_y = x_major ? 0 : sy;
  const int minor_x = x_major ? 0 : sx, minor_y = x_major ? sy : 0;
  int x = bx, y = by, num = size >> 1;
  for (int k = 0; k < size; ++k) {
    if (!sink.cell(x, y)) return false;
    num += add;
    if (num >= size) {
      num -= size;
      x += minor_x;
      y += minor_y;
    }
    x += major_x;
    y += major_y;
  }
  return true;
}

/// outline_generator over the transformed polygon; vertices are discretised on the fly (two live
/// vertices + the first).  `poly` = column-major xy in shared memory.
template <int XF, class Sink>
__device__ __forceinline__ bool walk_outline(const double* __restrict__ poly, int nv, const Pose& T, const MapView& m,
                                             Sink& sink) {
  if (nv == 0) return true;
  int fx, fy;
  to_cell<XF>(T, poly[0], poly[1], m, fx, fy);
  int px = fx, py = fy, rays = 0;
  for (int i = 1; i < nv; ++i) {
    int cx, cy;
    to_cell<XF>(T, poly[2 * i], poly[2 * i + 1], m, cx, cy);
    if (cx != px || cy != py) {  // degenerate rays are skipped (generators.cpp:86-93)
      ++rays;
      if (!walk_ray(px, py, cx, cy, sink)) return false;
    }
    px = cx, py = cy;
  }
  if (rays > 1) {  // close the outline (generators.cpp:69-71)
    if (px != fx || py != fy) return walk_ray(px, py, fx, fy, sink);
    return true;
  }
  return sink.cell(px, py);  // one-cell dummy ray holding the last point (generators.cpp:72-78)
}

/// outline_generator over an already discretised vertex list (`verts[0]` is implicitly (0,0) when
/// `first_is_origin`): used by the lattice kernel, which rasterises each distinct RELATIVE shape once.
template <class Sink>
__device__ __forceinline__ bool walk_outline_cells(const int2* __restrict__ verts, int nv, Sink& sink) {
  if (nv == 0) return true;
  const int2 f = verts[0];
  int2 p = f;
  int rays = 0;
  for (int i = 1; i < nv; ++i) {
    const int2 c = verts[i];
    if (c.x != p.x || c.y != p.y) {
      ++rays;
      if (!walk_ray(p.x, p.y, c.x, c.y, sink)) return false;
    }
    p = c;
  }
  if (rays > 1) {
    if (p.x != f.x || p.y != f.y) return walk_ray(p.x, p.y, f.x, f.y, sink);
    return true;
  }
  return sink.cell(p.x, p.y);
}

/// Integer bounding box of the discretised vertices + the coordinate-range check.
template <int XF>
__device__ __forceinline__ bool polygon_bbox(const double* __restrict__ poly, int nv, const Pose& T, const MapView& m, int& x0,
                                             int& y0, int& x1, int& y1) {
  bool ok = true;
  x0 = y0 = 0x7FFFFFFF;
  x1 = y1 = -0x7FFFFFFF - 1;
  for (int i = 0; i < nv; ++i) {
    int cx, cy;
    ok &= to_cell<XF>(T, poly[2 * i], poly[2 * i + 1], m, cx, cy);
    x0 = min(x0, cx), x1 = max(x1, cx);
    y0 = min(y0, cy), y1 = max(y1, cy);
  }
  return ok;
}

// ---------------------------------------------------------------------------------------------------
// Run detection: outline cell stream -> scan-line ranges (generators.cpp:189-232).
//
// A run is a maximal cyclic stretch of consecutive outline cells on one row.  Each run contributes the
// range (x of its first cell, x of its last cell); a run where y turns around (cusp) counts twice.  The
// reference inserts ranges in the order of the runs' END indices, starting with the run that contains
// outline cell 0 — whose beginning may lie at the tail of the outline.  We stream: every run but that
// first one is emitted when it closes; the first run is emitted last through `add_first`, which gives
// it the rank the reference gives it (it wins ties of equal `min`).
// ---------------------------------------------------------------------------------------------------
template <class Store>
struct RunDetector {
  Store& store;
  int seq_base;                                  // outline cell index offset (multi-outline order)
  int idx = 0;                                   // cells seen so far
  int y_run = 0, x_first = 0, x_prev = 0;        // current run
  int y_before = 0;                              // y of the cell before the current run
  int x0 = 0, y0 = 0;                            // outline cell 0
  int f_xlast = 0, f_ynext = 0, f_end = 0;       // the first run, closed
  bool in_first = true;
  int min_x, max_x;  // x extent of the vertices = x extent of the outline when every cell is on one row

  __device__ __forceinline__ RunDetector(Store& s, int base, int bbox_x0, int bbox_x1) :
      store(s), seq_base(base), min_x(bbox_x0), max_x(bbox_x1) {}

  __device__ __forceinline__ void emit(int y, int a, int b, int yp, int yn, int end_idx, bool first) {
    const int lo = min(a, b), hi = max(a, b);
    const bool cusp = (yp - y) * (y - yn) < 0;  // generators.cpp:100-103
    store.add(y, lo, hi, seq_base + end_idx, first);
    if (cusp) store.add(y, lo, hi, seq_base + end_idx, first);
  }

  __device__ __forceinline__ bool cell(int x, int y) {
    if (idx == 0) {
      y_run = y0 = y;
      x_first = x0 = x;
    } else if (y != y_run) {
      if (in_first) {
        f_xlast = x_prev, f_ynext = y, f_end = idx - 1;
        in_first = false;
      } else {
        emit(y_run, x_first, x_prev, y_before, y, idx - 1, false);
      }
      y_before = y_run;
      y_run = y;
      x_first = x;
    }
    x_prev = x;
    ++idx;
    return true;
  }

  __device__ __forceinline__ void finish() {
    if (idx == 0) return;
    if (in_first) {  // every cell on one row (generators.cpp:197-204)
      store.add(y0, min_x, min_x, seq_base, false);
      store.add(y0, max_x, max_x, seq_base, false);
      return;
    }
    if (y_run == y0) {  // the tail run continues into cell 0: one run, ending where the first run ended
      emit(y0, x_first, f_xlast, y_before, f_ynext, f_end, true);
    } else {
      emit(y_run, x_first, x_prev, y_before, y0, idx - 1, false);
      emit(y0, x0, f_xlast, y_run, f_ynext, f_end, true);
    }
  }
};

// ---------------------------------------------------------------------------------------------------
// Row stores.
// ---------------------------------------------------------------------------------------------------

/// Two ranges per row, packed in ONE shared-memory word per (row, thread): bits 0-14 lo, 15-29 hi
/// (relative to the vertex bounding box), 30-31 count.  The pair is merged the moment the second range
/// arrives: sorted by `min` with insertion order breaking ties, result [first.min, second.max]
/// (generators.cpp:236-255).  A third range marks the pose for the list store.
struct PackedStore {
  uint32_t* tab;  // &table[0][threadIdx.x]
  int stride;     // threads per block
  int xb, yb;     // bounding-box origin
  bool overflow = false;

  __device__ __forceinline__ void clear(int rows) {
    for (int r = 0; r < rows; ++r) tab[r * stride] = 0;
  }
  __device__ __forceinline__ void add(int y, int lo, int hi, int /*seq*/, bool first) {
    uint32_t& w = tab[(y - yb) * stride];
    const uint32_t cnt = w >> 30;
    const uint32_t nlo = static_cast<uint32_t>(lo - xb), nhi = static_cast<uint32_t>(hi - xb);
    if (cnt == 0) {
      w = nlo | (nhi << 15) | (1u << 30);
    } else if (cnt == 1) {
      const uint32_t olo = w & 0x7FFFu, ohi = (w >> 15) & 0x7FFFu;
      uint32_t rlo, rhi;
      if (first) {  // the new range was inserted BEFORE the stored one
        if (olo < nlo) rlo = olo, rhi = nhi;
        else rlo = nlo, rhi = ohi;
      } else {
        if (nlo < olo) rlo = nlo, rhi = ohi;
        else rlo = olo, rhi = nhi;
      }
      w = rlo | (rhi << 15) | (2u << 30);
    } else {
      overflow = true;
    }
  }
};

/// Sorted range lists per row in global scratch, interleaved across the resident threads
/// ([row][slot][thread]) so that a warp's accesses to the same (row, slot) coalesce.  Insertion keeps
/// the list ordered by (min, outline order) = what a stable sort by `min` produces.
struct ListStore {
  int* cnt;   // [rows_cap][nthreads]
  int2* rng;  // [rows_cap][kMaxRanges][nthreads]
  int* seq;   // [rows_cap][kMaxRanges][nthreads]
  long long nthreads;
  long long me;
  int yb, rows_cap;
  bool overflow = false;   // a row needed more than kMaxRanges entries
  bool out_of_rows = false;

  __device__ __forceinline__ long long at(int row, int k) const { return (static_cast<long long>(row) * kMaxRanges + k) * nthreads + me; }
  __device__ __forceinline__ void clear(int rows) {
    for (int r = 0; r < rows; ++r) cnt[static_cast<long long>(r) * nthreads + me] = 0;
  }
  __device__ __forceinline__ void add(int y, int lo, int hi, int s, bool /*first*/) {
    const int row = y - yb;
    if (row < 0 || row >= rows_cap) {
      out_of_rows = true;
      return;
    }
    int& n = cnt[static_cast<long long>(row) * nthreads + me];
    if (n >= kMaxRanges) {  // keep counting: the parity still decides std::logic_error
      overflow = true;
      ++n;
      return;
    }
    int k = n;
    while (k > 0) {
      const int2 e = rng[at(row, k - 1)];
      const int es = seq[at(row, k - 1)];
      if (e.x < lo || (e.x == lo && es <= s)) break;
      rng[at(row, k)] = e;
      seq[at(row, k)] = es;
      --k;
    }
    rng[at(row, k)] = make_int2(lo, hi);
    seq[at(row, k)] = s;
    ++n;
  }
};

}  // namespace cvr
