// ref_capi.cpp — C ABI over the REFERENCE's own sources (compiled unmodified from /root/reference by
// oracle/Makefile against the shim headers in this directory).  Output: oracle/_ref/libcover_ref.so.
//
// TEST INFRASTRUCTURE ONLY.  Used to (a) diff-fuzz the restatement in cover_oracle.hpp against the real
// generators.cpp / costmap.cpp / footprints.cpp and (b) optionally as bench.py's `cpu_baseline.kind =
// "reference"`.  Signatures mirror oracle_capi.cpp so that pyoracle can drive either library.
#include <cover/costmap.hpp>
#include <cover/expand.hpp>
#include <cover/footprints.hpp>
#include <cover/generators.hpp>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <thread>
#include <vector>

using costmap_2d::Costmap2D;

extern "C" {
struct oracle_map {
  const uint8_t* data;
  int32_t size_x, size_y;
  double resolution, origin_x, origin_y;
};
struct oracle_lattice {
  double x0, dx;
  double y0, dy;
  const double* cs;
  int32_t nx, ny, ntheta;
};
}

namespace {

enum { SHAPE_OUTLINE = 1, SHAPE_AREA = 2 };
enum { OP_IS_FREE = 0, OP_ACCUMULATE = 1 };
enum { POSE_CS = 0, POSE_LATTICE = 1, POSE_NONE = 2 };
enum { ST_OK = 0, ST_LOGIC_ERROR = 1 };

Costmap2D wrap(const oracle_map* m) {
  return Costmap2D(m->data, static_cast<unsigned>(m->size_x), static_cast<unsigned>(m->size_y), m->resolution, m->origin_x,
                   m->origin_y);
}

cover::polygon make_poly(const double* xy, int64_t n) {
  cover::polygon p(2, n);
  std::copy(xy, xy + 2 * n, p.data());
  return p;
}

cover::discrete_polygon make_dpoly(const int32_t* xy, int64_t n) {
  cover::discrete_polygon p(2, n);
  std::copy(xy, xy + 2 * n, p.data());
  return p;
}

Eigen::Isometry2d pose_at(int fmt, const void* poses, const oracle_lattice* lat, int64_t p) {
  double c, s, tx, ty;
  if (fmt == POSE_CS) {
    const double* q = static_cast<const double*>(poses) + 4 * p;
    c = q[0], s = q[1], tx = q[2], ty = q[3];
  } else {
    const int64_t ix = p % lat->nx, iy = (p / lat->nx) % lat->ny, it = p / (static_cast<int64_t>(lat->nx) * lat->ny);
    tx = lat->x0 + (static_cast<double>(ix) * lat->dx);
    ty = lat->y0 + (static_cast<double>(iy) * lat->dy);
    c = lat->cs[2 * it], s = lat->cs[2 * it + 1];
  }
  return Eigen::Isometry2d::FromParts(c, -s, s, c, tx, ty);
}

template <class Body>
void parallel_for(int64_t n, int threads, Body&& body) {
  if (threads <= 1 || n < 2) {
    body(0, n);
    return;
  }
  std::vector<std::thread> pool;
  const int64_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const int64_t b = t * chunk, e = std::min<int64_t>(n, b + chunk);
    if (b >= e) break;
    pool.emplace_back([=, &body] { body(b, e); });
  }
  for (auto& th : pool) th.join();
}

template <class Gen>
int64_t emit(const Gen& gen, int32_t* out, int64_t cap) {
  int64_t n = 0;
  for (const auto& c : gen) {
    if (n < cap) {
      out[2 * n] = c.x();
      out[2 * n + 1] = c.y();
    }
    ++n;
  }
  return n;
}

}  // namespace

extern "C" {

int64_t oracle_ray_cells(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t* out, int64_t cap) {
  return emit(cover::ray_generator(cover::cell(x0, y0), cover::cell(x1, y1)), out, cap);
}

int64_t oracle_outline_cells(const int32_t* sparse, int64_t n, int32_t* out, int64_t cap) {
  return emit(cover::outline_generator(make_dpoly(sparse, n)), out, cap);
}

int64_t oracle_area_cells(const int32_t* dense, const int64_t* sizes, int32_t n_outlines, int32_t* out, int64_t cap,
                          int32_t* max_ranges) {
  cover::discrete_polygon_vec outlines;
  int64_t at = 0;
  for (int32_t o = 0; o < n_outlines; ++o) {
    outlines.push_back(make_dpoly(dense + 2 * at, sizes[o]));
    at += sizes[o];
  }
  if (max_ranges) *max_ranges = -1;  // not observable through the reference's interface
  try {
    return emit(cover::area_generator(outlines), out, cap);
  } catch (const std::logic_error&) {
    return -1;
  }
}

int oracle_to_discrete(double res, const double* xy, int64_t n, int32_t* out) {
  try {
    const cover::discrete_polygon d = cover::to_discrete(res, make_poly(xy, n));
    std::copy(d.data(), d.data() + 2 * n, out);
    return 0;
  } catch (const std::invalid_argument&) {
    return -1;
  }
}

/// is_free / accumulate_cost(gen, offset, map) (src/cover/impl/costmap.hpp:29-35,74-81) over the outline / area generator
/// of the transformed polygon: the reference's own offset overloads.
int oracle_polygon_batch_offset(const oracle_map* map, const double* poly_xy, int32_t n_vertices, int32_t shape, int32_t op,
                                int32_t pose_format, const void* poses, const oracle_lattice* lattice, int64_t first, int64_t count,
                                const int32_t* offset, uint8_t* is_free, double* cost, uint8_t* status, int64_t* cells_read, int32_t threads) {
  if (!map || map->resolution <= 0) return -1;
  const Costmap2D cm = wrap(map);
  const cover::polygon poly = make_poly(poly_xy, n_vertices);
  const cover::point origin(cm.getOriginX(), cm.getOriginY());
  const cover::cell off(offset ? offset[0] : 0, offset ? offset[1] : 0);
  if (cells_read) *cells_read = -1;
  parallel_for(count, threads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      try {
        const cover::polygon q = pose_format == POSE_NONE
                                     ? cover::polygon(poly.colwise() - origin)
                                     : cover::polygon((pose_at(pose_format, poses, lattice, first + i) * poly).colwise() - origin);
        bool fr = false;
        double cs = 0;
        if (shape == SHAPE_AREA) {
          const cover::area_generator gen(cm.getResolution(), q);
          if (op == OP_IS_FREE) {
            fr = cover::is_free(gen, off, cm);
          } else {
            cs = cover::accumulate_cost(gen, off, cm);
            fr = cs >= 0;
          }
        } else {
          const cover::outline_generator gen(cm.getResolution(), q);
          if (op == OP_IS_FREE) {
            fr = cover::is_free(gen, off, cm);
          } else {
            cs = cover::accumulate_cost(gen, off, cm);
            fr = cs >= 0;
          }
        }
        if (is_free) is_free[i] = fr;
        if (cost) cost[i] = cs;
        if (status) status[i] = ST_OK;
      } catch (const std::logic_error&) {
        if (is_free) is_free[i] = 0;
        if (cost) cost[i] = std::numeric_limits<double>::quiet_NaN();
        if (status) status[i] = ST_LOGIC_ERROR;
      }
    }
  });
  return 0;
}

int oracle_polygon_batch(const oracle_map* map, const double* poly_xy, int32_t n_vertices, int32_t shape, int32_t op,
                         int32_t pose_format, const void* poses, const oracle_lattice* lattice, int64_t first,
                         int64_t count, uint8_t* is_free, double* cost, uint8_t* status, int64_t* cells_read,
                         int32_t threads) {
  if (!map || map->resolution <= 0) return -1;
  const Costmap2D cm = wrap(map);
  const cover::polygon poly = make_poly(poly_xy, n_vertices);
  const cover::point origin(cm.getOriginX(), cm.getOriginY());
  if (cells_read) *cells_read = -1;
  parallel_for(count, threads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      try {
        // costmap.cpp:71-85 spelled out for the outline and accumulate variants; the verdict over the area goes through
        // the public entry itself (which transforms the polygon on its own: no second transform here)
        auto transformed = [&] {
          return pose_format == POSE_NONE ? cover::polygon(poly.colwise() - origin)
                                          : cover::polygon((pose_at(pose_format, poses, lattice, first + i) * poly).colwise() - origin);
        };
        bool fr = false;
        double cs = 0;
        if (shape == SHAPE_AREA) {
          if (op == OP_IS_FREE && pose_format != POSE_NONE) {
            fr = cover::is_free(poly, pose_at(pose_format, poses, lattice, first + i), cm);  // costmap.cpp:78-85
          } else if (op == OP_IS_FREE) {
            fr = cover::is_free(poly, cm);
          } else {
            const cover::area_generator gen(cm.getResolution(), transformed());
            cs = cover::accumulate_cost(gen, cm);
            fr = cs >= 0;
          }
        } else {
          const cover::outline_generator gen(cm.getResolution(), transformed());
          if (op == OP_IS_FREE) {
            fr = cover::is_free(gen, cm);
          } else {
            cs = cover::accumulate_cost(gen, cm);
            fr = cs >= 0;
          }
        }
        if (is_free) is_free[i] = fr;
        if (cost) cost[i] = cs;
        if (status) status[i] = ST_OK;
      } catch (const std::logic_error&) {
        if (is_free) is_free[i] = 0;
        if (cost) cost[i] = std::numeric_limits<double>::quiet_NaN();
        if (status) status[i] = ST_LOGIC_ERROR;
      }
    }
  });
  return 0;
}

/// Several outlines (holes): is_free / accumulate_cost over area_generator(resolution, polygon_vec)
/// (src/cover/generators.cpp:133-135), every outline transformed like the single polygon of costmap.cpp:78-85.
int oracle_multi_polygon_batch(const oracle_map* map, const double* polys_xy, const int32_t* starts, int32_t n_outlines, int32_t op,
                               int32_t pose_format, const void* poses, const oracle_lattice* lattice, int64_t first, int64_t count,
                               uint8_t* is_free, double* cost, uint8_t* status, int64_t* cells_read, int32_t threads) {
  if (!map || map->resolution <= 0) return -1;
  const Costmap2D cm = wrap(map);
  std::vector<cover::polygon> polys;
  for (int32_t o = 0; o < n_outlines; ++o) polys.push_back(make_poly(polys_xy + 2 * starts[o], starts[o + 1] - starts[o]));
  const cover::point origin(cm.getOriginX(), cm.getOriginY());
  if (cells_read) *cells_read = -1;
  parallel_for(count, threads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      try {
        cover::polygon_vec moved;
        for (const cover::polygon& p : polys)
          moved.push_back(pose_format == POSE_NONE ? cover::polygon(p.colwise() - origin)
                                                   : cover::polygon((pose_at(pose_format, poses, lattice, first + i) * p).colwise() - origin));
        const cover::area_generator gen(cm.getResolution(), moved);
        bool fr;
        double cs = 0;
        if (op == OP_IS_FREE) {
          fr = cover::is_free(gen, cm);
        } else {
          cs = cover::accumulate_cost(gen, cm);
          fr = cs >= 0;
        }
        if (is_free) is_free[i] = fr;
        if (cost) cost[i] = cs;
        if (status) status[i] = ST_OK;
      } catch (const std::logic_error&) {
        if (is_free) is_free[i] = 0;
        if (cost) cost[i] = std::numeric_limits<double>::quiet_NaN();
        if (status) status[i] = ST_LOGIC_ERROR;
      }
    }
  });
  return 0;
}

int oracle_rays_batch(const oracle_map* map, const int32_t* segments, int64_t count, const int32_t* offset, int32_t op,
                      uint8_t* is_free, double* cost, uint8_t* status, int64_t* cells_read, int32_t threads) {
  const Costmap2D cm = wrap(map);
  if (cells_read) *cells_read = -1;
  parallel_for(count, threads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      const int32_t* s = segments + 4 * i;
      const cover::ray_generator ray(cover::cell(s[0], s[1]), cover::cell(s[2], s[3]));
      double cs = 0;
      bool fr;
      if (op == OP_IS_FREE) {
        fr = offset ? cover::is_free(ray, cover::cell(offset[0], offset[1]), cm) : cover::is_free(ray, cm);
      } else {
        cs = offset ? cover::accumulate_cost(ray, cover::cell(offset[0], offset[1]), cm) : cover::accumulate_cost(ray, cm);
        fr = cs >= 0;
      }
      if (is_free) is_free[i] = fr;
      if (cost) cost[i] = cs;
      if (status) status[i] = ST_OK;
    }
  });
  return 0;
}

int oracle_cells_batch(const oracle_map* map, const int32_t* cells, int64_t n_cells, const int32_t* offsets, int64_t count,
                       int32_t op, uint8_t* is_free, double* cost, uint8_t* status, int64_t* cells_read, int32_t threads) {
  const Costmap2D cm = wrap(map);
  const cover::discrete_polygon dp = make_dpoly(cells, n_cells);
  if (cells_read) *cells_read = -1;
  if (op != OP_IS_FREE) return -3;  // the reference has no accumulate_cost over a discrete_polygon
  parallel_for(count, threads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      is_free[i] = cover::is_free(dp, cover::cell(offsets[2 * i], offsets[2 * i + 1]), cm);
      if (cost) cost[i] = 0;
      if (status) status[i] = ST_OK;
    }
  });
  return 0;
}

static cover::polygon_vec unpack(const double* xy, const int32_t* starts, int32_t n) {
  cover::polygon_vec out;
  for (int32_t k = 0; k < n; ++k) out.push_back(make_poly(xy + 2 * starts[k], starts[k + 1] - starts[k]));
  return out;
}

int oracle_make_discrete_footprint(double res, const double* outline_xy, const int32_t* outline_starts, int32_t n_outlines,
                                   const double* area_xy, const int32_t* area_starts, int32_t n_areas,
                                   int32_t* outline_cells, int64_t* n_outline_cells, int32_t* area_cells_out,
                                   int64_t* n_area_cells) {
  try {
    const cover::continuous_footprint cf(unpack(outline_xy, outline_starts, n_outlines), unpack(area_xy, area_starts, n_areas));
    const cover::discrete_footprint df(res, cf);  // generator_footprint::init + discrete_footprint::init
    const auto& o = df.get_outlines();
    const auto& a = df.get_areas();
    *n_outline_cells = o.cols();
    *n_area_cells = a.cols();
    if (outline_cells) std::copy(o.data(), o.data() + o.size(), outline_cells);
    if (area_cells_out) std::copy(a.data(), a.data() + a.size(), area_cells_out);
    return 0;
  } catch (const std::invalid_argument&) {
    return -1;
  } catch (const std::logic_error&) {
    return -2;
  }
}

int oracle_discrete_footprint_batch(const oracle_map* map, const int32_t* outline_cells, const int64_t* o_starts,
                                    const int32_t* area_cells_in, const int64_t* a_starts, int32_t n_footprints,
                                    const int32_t* offsets, const int32_t* which, int64_t count, uint8_t* is_free,
                                    uint8_t* status, int64_t* cells_read, int32_t threads) {
  const Costmap2D cm = wrap(map);
  std::vector<cover::discrete_footprint> bank;
  for (int32_t k = 0; k < n_footprints; ++k)
    bank.emplace_back(make_dpoly(outline_cells + 2 * o_starts[k], o_starts[k + 1] - o_starts[k]),
                      make_dpoly(area_cells_in + 2 * a_starts[k], a_starts[k + 1] - a_starts[k]));
  if (cells_read) *cells_read = -1;
  parallel_for(count, threads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      const auto& fp = bank[static_cast<size_t>(which ? which[i] : 0)];
      is_free[i] = fp.is_free(cover::cell(offsets[2 * i], offsets[2 * i + 1]), cm);
      if (status) status[i] = ST_OK;
    }
  });
  return 0;
}

int oracle_continuous_footprint_batch(const oracle_map* map, const double* outline_xy, const int32_t* outline_starts,
                                      int32_t n_outlines, const double* area_xy, const int32_t* area_starts, int32_t n_areas,
                                      int32_t pose_format, const void* poses, const oracle_lattice* lattice, int64_t first,
                                      int64_t count, uint8_t* is_free, uint8_t* status, int64_t* cells_read, int32_t threads) {
  if (!map || map->resolution <= 0) return -1;
  const Costmap2D cm = wrap(map);
  const cover::continuous_footprint fp(unpack(outline_xy, outline_starts, n_outlines), unpack(area_xy, area_starts, n_areas));
  if (cells_read) *cells_read = -1;
  parallel_for(count, threads, [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      try {
        is_free[i] = fp.is_free(pose_at(pose_format, poses, lattice, first + i), cm);
        if (status) status[i] = ST_OK;
      } catch (const std::logic_error&) {
        is_free[i] = 0;
        if (status) status[i] = ST_LOGIC_ERROR;
      }
    }
  });
  return 0;
}

int oracle_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }

}  // extern "C"
