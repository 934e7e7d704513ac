// polygon_split.hpp — get_radius and the continuous-footprint split WITHOUT boost::geometry (host code, C++17).
//
// What the reference does (src/cover/footprints.cpp:55-128, impl/boost_io.hpp):
//   get_radius(polygon)               = min distance from the centroid to the polygon's edges
//   continuous_footprint(r, polygon)  : P = correct(polygon) (closed, clockwise)
//       eroded  = buffer(P, -r (1 - 1e-6))   round joins, 36 points per circle          -> OUTLINES
//       dilated = buffer(linestring(eroded.front().outer()), +r)                        (a band around the first ring)
//       diffs   = P \ dilated                                                           -> AREAS
//     polygons that are empty or have inner rings are dropped (boost_io.hpp:108-124).
// The reference's tests pin only properties of the result (test/footprints.cpp:20-166) and a safety regression
// (:205-242) - the vertex coordinates depend on boost's buffer strategies and are unpinned - so this file restates the
// CONSTRUCTION, not boost's output:
//   * P is eroded by d - 2 h instead of d, h = 1/384 of the larger bounding-box side: with the usual radius (the
//     inradius) the exact erosion is the polygon's skeleton - a point or a segment without area, which boost returns
//     as a sliver - and 2 h of width keep it a proper polygon.  A convex P is eroded exactly (its edges' half-planes
//     moved inwards: what a negative buffer of a convex polygon is, whatever the join strategy); any other P through
//     its implicit form - the signed distance to P's boundary - with marching squares on a grid of pitch h and a
//     Douglas-Peucker pass (tolerance h / 4);
//   * the areas are {p in P : dist(p, first eroded ring) > radius - 1.5 h}, again from the implicit form.  The band the
//     areas leave out is 1.5 h narrower than the one the outline ring answers for, so the two overlap by more than the
//     grid's error, and because the ring sits 2 h closer to P's boundary than boost's, the narrower band still reaches
//     over every edge it is parallel to: no rim of area is left along such an edge.
// Why that is safe: the areas are P minus (less than) the band of width `radius` around the ring that is actually
// RETURNED, so every point of P lies in an area or within `radius` of the outline ring, wherever that ring sits.  A ring
// 2 h closer to P's boundary than boost's only makes the outline test (253 or 254) trip for obstacles up to 2 h outside
// the footprint - a few millimetres on the conservative side.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <stdexcept>
#include <utility>
#include <vector>

namespace cvr_split {

struct Pt {
  double x, y;
};
using Ring = std::vector<Pt>;  // closed: front() == back()

inline double cross(const Pt& a, const Pt& b, const Pt& c) { return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x); }

/// Twice the signed area (positive: counter-clockwise); works on open and closed rings.
inline double signed_area2(const Ring& r) {
  double a = 0;
  for (size_t i = 0; i + 1 < r.size(); ++i) a += r[i].x * r[i + 1].y - r[i + 1].x * r[i].y;
  if (!r.empty() && (r.front().x != r.back().x || r.front().y != r.back().y)) a += r.back().x * r.front().y - r.front().x * r.back().y;
  return a;
}

/// boost::geometry::correct for the default polygon type: close the ring, make it clockwise (boost_io.hpp:44-60).
inline Ring correct(const double* xy, int n) {
  Ring r;
  for (int i = 0; i < n; ++i) r.push_back(Pt{xy[2 * i], xy[2 * i + 1]});
  if (r.empty()) return r;
  if (r.front().x != r.back().x || r.front().y != r.back().y) r.push_back(r.front());
  if (signed_area2(r) > 0) std::reverse(r.begin(), r.end());
  return r;
}

inline double dist_point_segment(const Pt& p, const Pt& a, const Pt& b) {
  const double vx = b.x - a.x, vy = b.y - a.y, wx = p.x - a.x, wy = p.y - a.y;
  const double vv = vx * vx + vy * vy;
  double t = vv > 0 ? (wx * vx + wy * vy) / vv : 0.0;
  t = std::min(1.0, std::max(0.0, t));
  return std::hypot(wx - t * vx, wy - t * vy);
}

/// Distance from p to the polyline through `r` (consecutive points; a closed ring is its own closing segment).
inline double dist_polyline(const Ring& r, const Pt& p) {
  if (r.size() == 1) return std::hypot(p.x - r[0].x, p.y - r[0].y);
  double d = std::numeric_limits<double>::max();
  for (size_t i = 0; i + 1 < r.size(); ++i) d = std::min(d, dist_point_segment(p, r[i], r[i + 1]));
  return d;
}

inline bool inside(const Ring& r, const Pt& p) {  // even-odd
  bool in = false;
  for (size_t i = 0; i + 1 < r.size(); ++i) {
    const Pt &a = r[i], &b = r[i + 1];
    if ((a.y > p.y) != (b.y > p.y) && p.x < a.x + (p.y - a.y) * (b.x - a.x) / (b.y - a.y)) in = !in;
  }
  return in;
}

/// get_radius (footprints.cpp:55-77): min distance from the (area-weighted) centroid to the edges; 0 for no points.
inline double get_radius(const double* xy, int n) {
  const Ring r = correct(xy, n);
  if (r.empty()) return 0;
  double a2 = 0, cx = 0, cy = 0;
  for (size_t i = 0; i + 1 < r.size(); ++i) {
    const double w = r[i].x * r[i + 1].y - r[i + 1].x * r[i].y;
    a2 += w, cx += (r[i].x + r[i + 1].x) * w, cy += (r[i].y + r[i + 1].y) * w;
  }
  Pt c;
  if (a2 != 0) {
    c = Pt{cx / (3 * a2), cy / (3 * a2)};
  } else {  // no area: the mean of the points
    c = Pt{0, 0};
    for (size_t i = 0; i + 1 < r.size() || i == 0; ++i) c.x += r[i].x, c.y += r[i].y;
    const double m = static_cast<double>(std::max<size_t>(1, r.size() - 1));
    c.x /= m, c.y /= m;
  }
  if (r.size() < 2) return 0;
  double d = std::numeric_limits<double>::max();
  for (size_t i = 0; i + 1 < r.size(); ++i) d = std::min(d, dist_point_segment(c, r[i], r[i + 1]));
  return d;
}

inline bool is_convex(const Ring& r) {
  int sign = 0;
  const size_t n = r.size() - 1;
  for (size_t i = 0; i < n; ++i) {
    const double c = cross(r[i], r[(i + 1) % n], r[(i + 2) % n]);
    if (c > 0) {
      if (sign < 0) return false;
      sign = 1;
    } else if (c < 0) {
      if (sign > 0) return false;
      sign = -1;
    }
  }
  return sign != 0;
}

/// Erosion of a convex clockwise ring by d: every edge's half-plane moved inwards (Sutherland-Hodgman).  Closed ring or empty.
inline Ring erode_convex(const Ring& r, double d) {
  std::vector<Pt> poly(r.begin(), r.end() - 1);
  const size_t n = r.size() - 1;
  for (size_t e = 0; e < n && !poly.empty(); ++e) {
    const Pt a = r[e], b = r[e + 1];
    const double len = std::hypot(b.x - a.x, b.y - a.y);
    if (len == 0) continue;
    const double nx = (b.y - a.y) / len, ny = -(b.x - a.x) / len;  // inward normal of a clockwise ring
    auto side = [&](const Pt& p) { return (p.x - a.x) * nx + (p.y - a.y) * ny - d; };
    std::vector<Pt> out;
    for (size_t i = 0; i < poly.size(); ++i) {
      const Pt p = poly[i], q = poly[(i + 1) % poly.size()];
      const double sp = side(p), sq = side(q);
      if (sp >= 0) out.push_back(p);
      if ((sp >= 0) != (sq >= 0)) {
        const double t = sp / (sp - sq);
        out.push_back(Pt{p.x + t * (q.x - p.x), p.y + t * (q.y - p.y)});
      }
    }
    poly.swap(out);
  }
  if (poly.size() < 3) return {};
  Ring ring(poly.begin(), poly.end());
  ring.push_back(ring.front());
  if (signed_area2(ring) > 0) std::reverse(ring.begin(), ring.end());
  return ring;
}

/// Douglas-Peucker on a closed ring (anchored at its first point and the point farthest from it).
inline void simplify_span(const Ring& r, size_t a, size_t b, double tol, std::vector<char>& keep) {
  if (b <= a + 1) return;
  double worst = -1;
  size_t at = a;
  for (size_t i = a + 1; i < b; ++i) {
    const double d = dist_point_segment(r[i], r[a], r[b]);
    if (d > worst) worst = d, at = i;
  }
  if (worst > tol) {
    keep[at] = 1;
    simplify_span(r, a, at, tol, keep);
    simplify_span(r, at, b, tol, keep);
  }
}
inline Ring simplify(const Ring& r, double tol) {
  if (r.size() < 5) return r;
  size_t far = 0;
  double best = -1;
  for (size_t i = 0; i + 1 < r.size(); ++i) {
    const double d = std::hypot(r[i].x - r[0].x, r[i].y - r[0].y);
    if (d > best) best = d, far = i;
  }
  std::vector<char> keep(r.size(), 0);
  keep[0] = keep[far] = keep[r.size() - 1] = 1;
  simplify_span(r, 0, far, tol, keep);
  simplify_span(r, far, r.size() - 1, tol, keep);
  Ring out;
  for (size_t i = 0; i < r.size(); ++i)
    if (keep[i]) out.push_back(r[i]);
  return out;
}

/// Closed iso-contours {f = 0} of a function sampled on a regular grid, region {f > 0} on the LEFT of the direction of
/// travel (outer boundaries counter-clockwise, holes clockwise).  Marching squares with linear interpolation; saddles are
/// resolved by the cell's mean.
inline std::vector<Ring> contours(const std::vector<double>& f, int nx, int ny, double x0, double y0, double h) {
  // grid edges: horizontal edge (i, j) joins nodes (i, j)-(i + 1, j): id = 2 * (j * nx + i); vertical (i, j)-(i, j + 1): id + 1
  auto val = [&](int i, int j) { return f[static_cast<size_t>(j) * nx + i]; };
  auto point_on = [&](long long id) {
    const int kind = static_cast<int>(id & 1);
    const long long node = id >> 1;
    const int i = static_cast<int>(node % nx), j = static_cast<int>(node / nx);
    const double a = val(i, j), b = kind ? val(i, j + 1) : val(i + 1, j);
    const double t = a / (a - b);
    return kind ? Pt{x0 + i * h, y0 + (j + t) * h} : Pt{x0 + (i + t) * h, y0 + j * h};
  };
  std::map<long long, long long> next;  // directed segment: from grid edge -> to grid edge
  for (int j = 0; j + 1 < ny; ++j) {
    for (int i = 0; i + 1 < nx; ++i) {
      const double v00 = val(i, j), v10 = val(i + 1, j), v11 = val(i + 1, j + 1), v01 = val(i, j + 1);
      const int code = (v00 > 0) | ((v10 > 0) << 1) | ((v11 > 0) << 2) | ((v01 > 0) << 3);
      if (code == 0 || code == 15) continue;
      const long long B = 2 * (static_cast<long long>(j) * nx + i), R = 2 * (static_cast<long long>(j) * nx + i + 1) + 1;
      const long long T = 2 * (static_cast<long long>(j + 1) * nx + i), L = 2 * (static_cast<long long>(j) * nx + i) + 1;
      auto seg = [&](long long from, long long to) { next[from] = to; };
      const bool centre = (v00 + v10 + v11 + v01) > 0;
      switch (code) {  // inside on the left
        case 1: seg(B, L); break;
        case 2: seg(R, B); break;
        case 3: seg(R, L); break;
        case 4: seg(T, R); break;
        case 5:
          if (centre) seg(T, L), seg(B, R);
          else seg(B, L), seg(T, R);
          break;
        case 6: seg(T, B); break;
        case 7: seg(T, L); break;
        case 8: seg(L, T); break;
        case 9: seg(B, T); break;
        case 10:
          if (centre) seg(L, B), seg(R, T);
          else seg(L, T), seg(R, B);
          break;
        case 11: seg(R, T); break;
        case 12: seg(L, R); break;
        case 13: seg(B, R); break;
        case 14: seg(L, B); break;
        default: break;
      }
    }
  }
  std::vector<Ring> rings;
  while (!next.empty()) {
    Ring ring;
    const long long start = next.begin()->first;
    long long at = start;
    for (;;) {
      const auto it = next.find(at);
      if (it == next.end()) break;  // (open chain: the region touched the grid border - the caller pads the grid)
      ring.push_back(point_on(at));
      const long long to = it->second;
      next.erase(it);
      at = to;
      if (at == start) break;
    }
    if (ring.size() >= 3 && at == start) {
      ring.push_back(ring.front());
      rings.push_back(std::move(ring));
    }
  }
  return rings;
}

/// Nearest point of the closed ring `r` to p (and its distance).
inline Pt nearest_on_ring(const Ring& r, const Pt& p, double& dist) {
  Pt best = r[0];
  dist = std::numeric_limits<double>::max();
  for (size_t i = 0; i + 1 < r.size(); ++i) {
    const Pt &a = r[i], &b = r[i + 1];
    const double vx = b.x - a.x, vy = b.y - a.y, vv = vx * vx + vy * vy;
    double t = vv > 0 ? ((p.x - a.x) * vx + (p.y - a.y) * vy) / vv : 0.0;
    t = std::min(1.0, std::max(0.0, t));
    const Pt q{a.x + t * vx, a.y + t * vy};
    const double d = std::hypot(p.x - q.x, p.y - q.y);
    if (d < dist) dist = d, best = q;
  }
  return best;
}

/// The grid leaves an area ring's vertices a hair inside P where the ring runs along P's boundary, and chamfers P's
/// corners.  boost's difference keeps P's own edges and vertices there, and the cell a vertex truncates to depends on
/// it (a coordinate of exactly 0.65 m is row 113, 0.65 - 1e-9 is row 112): vertices within `reach` of P's boundary are
/// moved ONTO it, and P's vertices that the ring passes within `reach` of are put back into the ring.
inline void snap_to_polygon(Ring& ring, const Ring& P, double reach) {
  for (Pt& v : ring) {
    double d;
    const Pt q = nearest_on_ring(P, v, d);
    if (d <= reach) v = q;
  }
  ring.back() = ring.front();
  for (size_t k = 0; k + 1 < P.size(); ++k) {
    const Pt& corner = P[k];
    double best = std::numeric_limits<double>::max();
    size_t at = 0;
    bool present = false;
    for (size_t i = 0; i + 1 < ring.size(); ++i) {
      if (ring[i].x == corner.x && ring[i].y == corner.y) present = true;
      const double d = dist_point_segment(corner, ring[i], ring[i + 1]);
      if (d < best) best = d, at = i;
    }
    if (!present && best <= reach) ring.insert(ring.begin() + static_cast<long>(at) + 1, corner);
  }
}

struct Split {
  std::vector<Ring> outlines, areas;  // closed, clockwise (boost's default polygon)
};

/// Keeps the counter-clockwise loops that contain no clockwise loop (polygons with inner rings are dropped,
/// boost_io.hpp:108-111), simplified and turned clockwise.
inline std::vector<Ring> simple_polygons(std::vector<Ring> loops, double tol, double min_width = 0) {
  std::vector<Ring> outers, holes, out;
  for (Ring& r : loops) (signed_area2(r) > 0 ? outers : holes).push_back(std::move(r));
  for (Ring& o : outers) {
    bool has_hole = false;
    for (const Ring& hl : holes) has_hole = has_hole || inside(o, hl[0]);
    if (has_hole) continue;
    Ring s = simplify(o, tol);
    if (s.size() < 4) continue;
    if (min_width > 0) {  // numerical slivers where the band is tangent to an edge of P: thinner than the grid resolves
      double perimeter = 0;
      for (size_t i = 0; i + 1 < s.size(); ++i) perimeter += std::hypot(s[i + 1].x - s[i].x, s[i + 1].y - s[i].y);
      if (std::fabs(signed_area2(s)) < min_width * perimeter) continue;  // area / perimeter < min_width / 2
    }
    std::reverse(s.begin(), s.end());
    out.push_back(std::move(s));
  }
  std::sort(out.begin(), out.end(), [](const Ring& a, const Ring& b) { return std::fabs(signed_area2(a)) > std::fabs(signed_area2(b)); });
  return out;
}

/// continuous_footprint(radius, polygon) (footprints.cpp:79-125).  Throws std::invalid_argument for radius <= 0.
inline Split split(double radius, const double* xy, int n, double slack = 0) {
  Split s;
  if (n == 0) return s;  // an empty polygon is a noop
  if (!(radius > 0)) throw std::invalid_argument("Radius must be positive");
  const Ring P = correct(xy, n);
  double bx0 = P[0].x, bx1 = P[0].x, by0 = P[0].y, by1 = P[0].y;
  for (const Pt& p : P) bx0 = std::min(bx0, p.x), bx1 = std::max(bx1, p.x), by0 = std::min(by0, p.y), by1 = std::max(by1, p.y);
  const double extent = std::max(bx1 - bx0, by1 - by0);
  if (std::fabs(signed_area2(P)) <= 1e-12 * extent * extent) return s;  // fewer than three distinct points / no area
  const double h = extent / 384.0, margin = 2.0 * h, tol = 0.25 * h;
  // (slack: the outlines that much closer to P's boundary, see cover_b200.h)
  const double d = radius * (1.0 - 1e-6) - margin - std::max(0.0, slack);
  const int pad = 4;
  const int nx = static_cast<int>(std::ceil((bx1 - bx0) / h)) + 2 * pad + 1, ny = static_cast<int>(std::ceil((by1 - by0) / h)) + 2 * pad + 1;
  const double gx0 = bx0 - pad * h, gy0 = by0 - pad * h;
  // signed distance to the boundary of P on the grid: positive inside
  std::vector<double> sdf(static_cast<size_t>(nx) * ny);
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const Pt p{gx0 + i * h, gy0 + j * h};
      const double dist = dist_polyline(P, p);
      sdf[static_cast<size_t>(j) * nx + i] = inside(P, p) ? dist : -dist;
    }
  // ---- outlines: P eroded by d
  if (d <= 0) {
    s.areas.push_back(P);
    return s;
  }
  if (is_convex(P)) {
    const Ring e = erode_convex(P, d);
    if (!e.empty()) s.outlines.push_back(e);
  } else {
    std::vector<double> f(sdf.size());
    for (size_t k = 0; k < f.size(); ++k) f[k] = sdf[k] - d;
    s.outlines = simple_polygons(contours(f, nx, ny, gx0, gy0, h), tol);
  }
  if (s.outlines.empty()) {  // the radius is too large: the polygon itself is the one area (test/footprints.cpp:90-105)
    s.areas.push_back(P);
    return s;
  }
  // ---- areas: P without the band of width `radius` around the first eroded ring
  const Ring& ring0 = s.outlines.front();
  std::vector<double> g(sdf.size());
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const size_t k = static_cast<size_t>(j) * nx + i;
      const double band = dist_polyline(ring0, Pt{gx0 + i * h, gy0 + j * h}) - (radius - 1.5 * h - std::max(0.0, slack));
      g[k] = std::min(sdf[k], band);
    }
  s.areas = simple_polygons(contours(g, nx, ny, gx0, gy0, h), tol, 0.5 * h);
  for (Ring& a : s.areas) snap_to_polygon(a, P, 1.5 * h);
  return s;
}

}  // namespace cvr_split
