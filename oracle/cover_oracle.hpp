// cover_oracle.hpp — CPU restatement of magazino/cover's hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing in the product (cover_b200/, include/) may include,
// link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / the CPU arm.
//
// Parity pin: this restatement is checked (tests/test_oracle_golden.py) against every known-answer
// vector the reference's own tests hold for the path (test/generators.cpp, test/expand.cpp incl. the
// 33 cv::fillPoly equalities, test/costmap.cpp, test/footprints.cpp:244-305) and, when /root/reference
// is present, differentially against the reference's own generators.cpp/costmap.cpp compiled over
// shim headers (oracle/_ref, see oracle/ref_shim/README.md).  What stays UNPINNED is third-party
// arithmetic that is not in /root/reference: the floating-point evaluation order of Eigen's
// `Isometry2d * Matrix2Xd` (restated from Eigen 3.3/3.4's transform_right_product_impl: res = t;
// res += L*p with L*p = (l_i0*p_0) + (l_i1*p_1), no FMA) and the costmap_2d constants 253/254/255.
//
// Every function cites the reference file:line (relative to /root/reference) it follows.  The code
// is organised differently on purpose (eager vectors and callbacks instead of iterator classes), but
// the visiting ORDER of cells, the std::sort call and all early-exit rules are the reference's.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <vector>

namespace oracle {

// costmap_2d/cost_values.h (third-party; values restated, see header note).
constexpr uint8_t kNoInformation = 255;
constexpr uint8_t kLethal = 254;
constexpr uint8_t kInscribed = 253;

// src/cover/costmap.cpp:28-33
constexpr double kCostLethal = -1.0;
constexpr double kCostNoInformation = -2.0;
constexpr double kCostOutOfMap = -3.0;

struct Cell {
  int x, y;
  bool operator==(const Cell& o) const { return x == o.x && y == o.y; }
  bool operator!=(const Cell& o) const { return !(*this == o); }
};
using Cells = std::vector<Cell>;

/// Stand-in for costmap_2d::Costmap2D: row-major bytes, index = y * size_x + x
/// (call sites: src/cover/costmap.hpp:52,89,105,122).
struct MapView {
  const uint8_t* data;
  int size_x, size_y;
  double resolution, origin_x, origin_y;
  uint8_t cost(int x, int y) const { return data[static_cast<size_t>(y) * size_x + x]; }
};

/// 2x3 affine pose (Eigen::Isometry2d: linear part + translation).
struct Pose {
  double m00, m01, m10, m11, tx, ty;
  static Pose from_cs(double c, double s, double tx, double ty) {
    // Eigen::Rotation2D::toRotationMatrix(): [c -s; s c]
    return Pose{c, -s, s, c, tx, ty};
  }
};

// ---------------------------------------------------------------------------------------------
// L0: discretisation — src/cover/base.hpp:50-57
// ---------------------------------------------------------------------------------------------

/// `(x * (1./res))` then `cast<int>()` = C++ truncation toward zero.
inline int discretise(double v, double inv_res) { return static_cast<int>(v * inv_res); }

/// Polygon layout = Eigen column-major 2xN: x0,y0,x1,y1,...  (src/cover/base.hpp:26)
inline Cells to_discrete(double res, const double* xy, size_t n) {
  if (res <= 0) throw std::invalid_argument("The resolution must be positive");
  const double inv = 1. / res;
  Cells out(n);
  for (size_t i = 0; i < n; ++i) out[i] = Cell{discretise(xy[2 * i], inv), discretise(xy[2 * i + 1], inv)};
  return out;
}

// ---------------------------------------------------------------------------------------------
// L1: ray — ctor src/cover/generators.cpp:30-52, stepping src/cover/generators.hpp:94-128
// ---------------------------------------------------------------------------------------------

struct Ray {
  Cell begin;
  Cell step_minor, step_major;
  unsigned size, add;
};

inline int sgn(int v) { return (v > 0) - (v < 0); }

inline Ray make_ray(Cell b, Cell e) {
  const int dx = e.x - b.x, dy = e.y - b.y;
  const int ax = std::abs(dx), ay = std::abs(dy);
  // maxCoeff/minCoeff keep the FIRST extremum: ties make x the major axis (generators.cpp:38-51).
  const bool x_major = ax >= ay;
  Ray r;
  r.begin = b;
  r.size = static_cast<unsigned>(x_major ? ax : ay);
  r.add = static_cast<unsigned>(x_major ? ay : ax);
  r.step_major = x_major ? Cell{sgn(dx), 0} : Cell{0, sgn(dy)};
  r.step_minor = x_major ? Cell{0, sgn(dy)} : Cell{sgn(dx), 0};
  return r;
}

/// Visits the `size` cells of the ray (start inclusive, end exclusive). `f(cell)` returns false to stop.
template <class F>
inline bool walk_ray(const Ray& r, F&& f) {
  Cell c = r.begin;
  int num = static_cast<int>(r.size / 2);  // generators.hpp:100
  for (unsigned k = 0; k < r.size; ++k) {
    if (!f(c)) return false;
    num += static_cast<int>(r.add);  // generators.hpp:118-123
    if (num >= static_cast<int>(r.size)) {
      num -= static_cast<int>(r.size);
      c.x += r.step_minor.x;
      c.y += r.step_minor.y;
    }
    c.x += r.step_major.x;
    c.y += r.step_major.y;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// L1: outline — src/cover/generators.cpp:54-93, iteration src/cover/generators.hpp:247-255
// ---------------------------------------------------------------------------------------------

struct Outline {
  std::vector<Ray> rays;
  size_t size = 0;
};

inline Outline make_outline(const Cells& sparse) {
  Outline o;
  const size_t n = sparse.size();
  if (n == 0) return o;
  auto push = [&o](Cell b, Cell e) {  // generators.cpp:86-93: skip degenerate rays
    if (b != e) {
      o.rays.push_back(make_ray(b, e));
      o.size += o.rays.back().size;
    }
  };
  for (size_t i = 0; i + 1 < n; ++i) push(sparse[i], sparse[i + 1]);
  if (o.rays.size() > 1) {
    push(sparse[n - 1], sparse[0]);  // close (generators.cpp:69-71)
  } else {
    // generators.cpp:72-78: one-cell dummy ray so that the last point is part of the outline.
    push(sparse[n - 1], Cell{sparse[n - 1].x + 1, sparse[n - 1].y + 1});
  }
  return o;
}

template <class F>
inline bool walk_outline(const Outline& o, F&& f) {
  for (const Ray& r : o.rays)
    if (!walk_ray(r, f)) return false;
  return true;
}

/// impl/expand.hpp:23-39 for an outline_generator ("to_outline", expand.hpp:68-72).
inline Cells dense_outline(const Cells& sparse) {
  const Outline o = make_outline(sparse);
  Cells out;
  out.reserve(o.size);
  walk_outline(o, [&out](Cell c) {
    out.push_back(c);
    return true;
  });
  return out;
}

// ---------------------------------------------------------------------------------------------
// L1: area — src/cover/generators.cpp:95-268, iteration src/cover/generators.hpp:402-433
// ---------------------------------------------------------------------------------------------

struct Span {
  int y, lo, hi;
};

struct Area {
  std::vector<Span> spans;       ///< merged ranges, y ascending, per-row order = sorted pair order
  size_t size = 0;               ///< number of visited cells (duplicates counted)
  int max_ranges_per_row = 0;    ///< before pairing; > 16 means std::sort was not insertion-stable
};

/// `dense` = closed dense outlines (each the expansion of an outline_generator).
/// Throws std::logic_error exactly where generators.cpp:246-249 does.
inline Area make_area(const std::vector<Cells>& dense) {
  Area a;
  size_t total = 0;
  for (const Cells& o : dense) total += o.size();
  if (total == 0) return a;  // generators.cpp:153-154

  int y_lo = std::numeric_limits<int>::max(), y_hi = std::numeric_limits<int>::min();
  for (const Cells& o : dense)
    for (const Cell& c : o) {  // generators.cpp:158-170
      y_lo = std::min(y_lo, c.y);
      y_hi = std::max(y_hi, c.y);
    }

  struct Rng {
    int y, lo, hi;
  };
  auto mk = [](int y, int p, int q) { return Rng{y, std::min(p, q), std::max(p, q)}; };  // generators.cpp:105-110
  std::vector<std::vector<Rng>> rows(static_cast<size_t>(y_hi - y_lo) + 1);

  for (const Cells& o : dense) {
    const size_t n = o.size();
    if (n == 0) continue;  // the reference indexes out of bounds here (UB); we define "skip"

    // generators.cpp:189-193: last index whose y differs from the first cell's y.
    size_t end = n - 1;
    while (end != 0 && o[end].y == o[0].y) --end;

    if (end == 0) {  // generators.cpp:197-204: everything on one row
      int mn = o[0].x, mx = o[0].x;
      for (const Cell& c : o) {
        mn = std::min(mn, c.x);
        mx = std::max(mx, c.x);
      }
      auto& row = rows[static_cast<size_t>(o[0].y - y_lo)];
      row.push_back(mk(o[0].y, mn, mn));
      row.push_back(mk(o[0].y, mx, mx));
    }

    auto wrap = [n](size_t i) { return i == n ? 0 : i; };
    size_t last = end;
    for (size_t cur = 0; cur <= end; ++cur) {  // generators.cpp:211-232
      const size_t nxt = wrap(cur + 1);
      if (o[cur].y == o[nxt].y) continue;  // run continues
      const size_t start = wrap(last + 1);
      auto& row = rows[static_cast<size_t>(o[cur].y - y_lo)];
      row.push_back(mk(o[cur].y, o[start].x, o[cur].x));
      // cusp (generators.cpp:95-103,226-229): y turns around at this run -> count it twice
      if ((o[last].y - o[cur].y) * (o[cur].y - o[nxt].y) < 0) row.push_back(mk(o[cur].y, o[start].x, o[cur].x));
      last = cur;
    }
  }

  for (auto& row : rows) {
    a.max_ranges_per_row = std::max(a.max_ranges_per_row, static_cast<int>(row.size()));
    // generators.cpp:236-241 — literally std::sort with a key-only comparator (not stable for n > 16).
    std::sort(row.begin(), row.end(), [](const Rng& l, const Rng& r) { return l.lo < r.lo; });
  }
  for (auto& row : rows) {
    if (row.size() % 2 != 0) throw std::logic_error("Even number of line-pairs expected");
    for (size_t i = 0; i < row.size(); i += 2) {  // generators.cpp:253-254: [r[2k].min, r[2k+1].max]
      a.spans.push_back(Span{row[i].y, row[i].lo, row[i + 1].hi});
    }
  }
  for (const Span& s : a.spans) a.size += static_cast<size_t>(s.hi - s.lo + 1);
  return a;
}

/// generators.hpp:417-423: rows ascending, x ascending inside each merged range.
template <class F>
inline bool walk_area(const Area& a, F&& f) {
  for (const Span& s : a.spans) {
    int x = s.lo;
    do {  // the reference always emits `min` first, then steps while x <= max
      if (!f(Cell{x, s.y})) return false;
    } while (++x <= s.hi);
  }
  return true;
}

inline Cells area_cells(const Area& a) {
  Cells out;
  out.reserve(a.size);
  walk_area(a, [&out](Cell c) {
    out.push_back(c);
    return true;
  });
  return out;
}

// ---------------------------------------------------------------------------------------------
// L3: cell predicates — src/cover/costmap.hpp:41-155, src/cover/costmap.cpp:35-46
// ---------------------------------------------------------------------------------------------

inline bool is_inside(const MapView& m, Cell c) { return c.x >= 0 && c.y >= 0 && c.x < m.size_x && c.y < m.size_y; }
inline bool is_lethal(const MapView& m, Cell c) { return !is_inside(m, c) || m.cost(c.x, c.y) == kLethal; }
inline bool is_lethal_or_inflated(const MapView& m, Cell c) {
  if (!is_inside(m, c)) return true;
  const uint8_t v = m.cost(c.x, c.y);
  return v == kInscribed || v == kLethal;
}
inline double model_cost(const MapView& m, Cell c) {
  if (!is_inside(m, c)) return kCostOutOfMap;
  const uint8_t v = m.cost(c.x, c.y);
  if (v == kLethal) return kCostLethal;
  if (v == kNoInformation) return kCostNoInformation;
  return v;
}

/// Counters so that the bench can state the ALGORITHMIC bytes: how many costmap cells
/// the reference's sequential loop reads before it returns.
struct Tally {
  int64_t cells_read = 0;
};

// ---------------------------------------------------------------------------------------------
// L3: pose checks — src/cover/impl/costmap.hpp:23-81, src/cover/costmap.cpp:59-85
// ---------------------------------------------------------------------------------------------

enum class Shape { Ray, Outline, Area };

/// Walks `sparse` (discrete vertices) as ray (needs exactly 2 cells) / outline / area in the
/// reference's generator order.
template <class F>
inline bool walk_shape(Shape shape, const Cells& sparse, F&& f) {
  switch (shape) {
    case Shape::Ray:
      return walk_ray(make_ray(sparse.at(0), sparse.at(1)), f);
    case Shape::Outline:
      return walk_outline(make_outline(sparse), f);
    case Shape::Area:
      return walk_area(make_area({dense_outline(sparse)}), f);
  }
  return true;
}

/// std::none_of(gen, with_offset(is_lethal)) — impl/costmap.hpp:23-35
template <class Walk>
inline bool is_free_walk(Walk&& walk, Cell off, const MapView& m, Tally* t = nullptr) {
  return walk([&](Cell c) {
    const Cell q{c.x + off.x, c.y + off.y};
    if (t && is_inside(m, q)) ++t->cells_read;
    return !is_lethal(m, q);
  });
}

/// impl/costmap.hpp:50-63 — sum as double, first negative value (generator order) wins.
template <class Walk>
inline double accumulate_walk(Walk&& walk, Cell off, const MapView& m, Tally* t = nullptr) {
  double sum = 0, early = 0;
  const bool done = walk([&](Cell c) {
    const Cell q{c.x + off.x, c.y + off.y};
    if (t && is_inside(m, q)) ++t->cells_read;
    const double v = model_cost(m, q);
    if (v < 0) {
      early = v;
      return false;
    }
    sum += v;
    return true;
  });
  return done ? sum : early;
}

/// EXTENSION (not in the reference): std::max fold of get_cell_cost (costmap.hpp:113-124)
/// over the generator; a cell outside the map makes the result -3 (mirrors OUT_OF_MAP).
template <class Walk>
inline double max_cost_walk(Walk&& walk, Cell off, const MapView& m, Tally* t = nullptr) {
  int best = 0;
  const bool inside = walk([&](Cell c) {
    const Cell q{c.x + off.x, c.y + off.y};
    if (!is_inside(m, q)) return false;
    if (t) ++t->cells_read;
    best = std::max(best, static_cast<int>(m.cost(q.x, q.y)));
    return true;
  });
  return inside ? static_cast<double>(best) : kCostOutOfMap;
}

/// Path A transform — src/cover/costmap.cpp:73-74,81-83: ((t + L*p) - origin), then to_discrete.
/// Eigen order (third-party, restated): res = t; res += (l_i0*p_x) + (l_i1*p_y).
inline Cells transform_path_a(const Pose& T, const double* xy, size_t n, const MapView& m) {
  if (m.resolution <= 0) throw std::invalid_argument("The resolution must be positive");
  const double inv = 1. / m.resolution;
  Cells out(n);
  for (size_t i = 0; i < n; ++i) {
    const double px = xy[2 * i], py = xy[2 * i + 1];
    const double wx = T.tx + ((T.m00 * px) + (T.m01 * py));
    const double wy = T.ty + ((T.m10 * px) + (T.m11 * py));
    out[i] = Cell{discretise(wx - m.origin_x, inv), discretise(wy - m.origin_y, inv)};
  }
  return out;
}

/// No-pose variant — src/cover/costmap.cpp:71-76: (p - origin) only.
inline Cells shift_path_a(const double* xy, size_t n, const MapView& m) {
  if (m.resolution <= 0) throw std::invalid_argument("The resolution must be positive");
  const double inv = 1. / m.resolution;
  Cells out(n);
  for (size_t i = 0; i < n; ++i)
    out[i] = Cell{discretise(xy[2 * i] - m.origin_x, inv), discretise(xy[2 * i + 1] - m.origin_y, inv)};
  return out;
}

/// Path B transform — src/cover/footprints.cpp:134-135,160,179: t' = t + (-origin); q = t' + L*p.
inline Cells transform_path_b(const Pose& T, const double* xy, size_t n, const MapView& m) {
  if (m.resolution <= 0) throw std::invalid_argument("The resolution must be positive");
  const double inv = 1. / m.resolution;
  const double tx = T.tx + (-m.origin_x), ty = T.ty + (-m.origin_y);
  Cells out(n);
  for (size_t i = 0; i < n; ++i) {
    const double px = xy[2 * i], py = xy[2 * i + 1];
    out[i] = Cell{discretise(tx + ((T.m00 * px) + (T.m01 * py)), inv), discretise(ty + ((T.m10 * px) + (T.m11 * py)), inv)};
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
// L4: footprints — src/cover/footprints.cpp:130-247
// ---------------------------------------------------------------------------------------------

struct Polygon {
  std::vector<double> xy;  ///< column-major 2xN
  size_t n() const { return xy.size() / 2; }
};

struct ContinuousFootprint {
  std::vector<Polygon> outlines, areas;
};

/// continuous_footprint::is_free — footprints.cpp:130-196.  May throw std::logic_error (area).
inline bool continuous_is_free(const ContinuousFootprint& fp, const Pose& T, const MapView& m, Tally* t = nullptr) {
  auto bbox_inside = [&m](const Cells& d) {  // footprints.cpp:139-155: SPARSE vertices only
    if (d.empty()) return true;
    int x0 = d[0].x, x1 = d[0].x, y0 = d[0].y, y1 = d[0].y;
    for (const Cell& c : d) {
      x0 = std::min(x0, c.x);
      x1 = std::max(x1, c.x);
      y0 = std::min(y0, c.y);
      y1 = std::max(y1, c.y);
    }
    if (x0 < 0 || y0 < 0) return false;
    if (x1 >= m.size_x || y1 >= m.size_y) return false;
    return true;
  };
  for (const Polygon& o : fp.outlines) {  // footprints.cpp:158-174
    const Cells d = transform_path_b(T, o.xy.data(), o.n(), m);
    if (!bbox_inside(d)) return false;
    const bool clear = walk_outline(make_outline(d), [&](Cell c) {
      if (t) ++t->cells_read;
      const uint8_t v = m.cost(c.x, c.y);  // raw getCost, no per-cell bounds test
      return !(v == kInscribed || v == kLethal);
    });
    if (!clear) return false;
  }
  for (const Polygon& a : fp.areas) {  // footprints.cpp:177-193
    const Cells d = transform_path_b(T, a.xy.data(), a.n(), m);
    if (!bbox_inside(d)) return false;
    const Area ar = make_area({dense_outline(d)});
    const bool clear = walk_area(ar, [&](Cell c) {
      if (t) ++t->cells_read;
      return m.cost(c.x, c.y) != kLethal;
    });
    if (!clear) return false;
  }
  return true;
}

/// generator_footprint / discrete_footprint — footprints.cpp:198-247: per-sub-polygon generators
/// (or their concatenated expansion, which visits the same cells in the same order).
struct DiscreteFootprint {
  Cells outlines, areas;  ///< expand_multiple of the outline / area generators (footprints.cpp:233-238)
};

/// generator_footprint::init + discrete_footprint::init — footprints.cpp:198-212,233-238.
inline DiscreteFootprint make_discrete_footprint(double res, const ContinuousFootprint& fp) {
  DiscreteFootprint d;
  for (const Polygon& a : fp.areas) {
    const Cells cs = area_cells(make_area({dense_outline(to_discrete(res, a.xy.data(), a.n()))}));
    d.areas.insert(d.areas.end(), cs.begin(), cs.end());
  }
  for (const Polygon& o : fp.outlines) {
    const Cells cs = dense_outline(to_discrete(res, o.xy.data(), o.n()));
    d.outlines.insert(d.outlines.end(), cs.begin(), cs.end());
  }
  return d;
}

/// discrete_footprint::is_free — footprints.cpp:240-247 (+ impl/costmap.hpp:39-48).
inline bool discrete_is_free(const DiscreteFootprint& fp, Cell off, const MapView& m, Tally* t = nullptr) {
  for (const Cell& c : fp.outlines) {
    const Cell q{c.x + off.x, c.y + off.y};
    if (t && is_inside(m, q)) ++t->cells_read;
    if (is_lethal_or_inflated(m, q)) return false;
  }
  for (const Cell& c : fp.areas) {
    const Cell q{c.x + off.x, c.y + off.y};
    if (t && is_inside(m, q)) ++t->cells_read;
    if (is_lethal(m, q)) return false;
  }
  return true;
}

}  // namespace oracle
