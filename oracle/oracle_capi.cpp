// oracle_capi.cpp — C ABI over cover_oracle.hpp so that pytest (ctypes) and bench.py's CPU arm can
// drive the CPU restatement on whole pose batches.
//
// *** TEST INFRASTRUCTURE ONLY *** (see the header of cover_oracle.hpp).  This library is never
// linked into, loaded by, or called from cover_b200's product path.
//
// Batch conventions are the same as include/cover_b200.h so that every GPU parity test is a diff of
// two arrays: poses N x {c,s,tx,ty}; lattice index p = (it*ny + iy)*nx + ix with
// t = (x0 + ix*dx, y0 + iy*dy) (multiply, then add; no FMA); per-pose status bytes.
#include "cover_oracle.hpp"

#include <atomic>
#include <cstring>
#include <thread>

using namespace oracle;

extern "C" {

struct oracle_map {
  const uint8_t* data;
  int32_t size_x, size_y;
  double resolution, origin_x, origin_y;
};

struct oracle_lattice {
  double x0, dx;
  double y0, dy;
  const double* cs;  // ntheta x {cos, sin}
  int32_t nx, ny, ntheta;
};

enum { ORACLE_SHAPE_OUTLINE = 1, ORACLE_SHAPE_AREA = 2 };
enum { ORACLE_OP_IS_FREE = 0, ORACLE_OP_ACCUMULATE = 1, ORACLE_OP_MAX = 2 };
enum { ORACLE_POSE_CS = 0, ORACLE_POSE_LATTICE = 1, ORACLE_POSE_NONE = 2 };
enum { ORACLE_OK = 0, ORACLE_LOGIC_ERROR = 1, ORACLE_ROW_COMPLEXITY = 2, ORACLE_COORD_RANGE = 3 };

}  // extern "C"

namespace {

constexpr double kCoordLimit = 4194304.0;  // 2^22 cells; same rule as the CUDA library

MapView view(const oracle_map* m) { return MapView{m->data, m->size_x, m->size_y, m->resolution, m->origin_x, m->origin_y}; }

template <class Body>
void parallel_for(int64_t n, int threads, Body&& body) {
  if (threads <= 1 || n < 2) {
    body(0, n, 0);
    return;
  }
  std::vector<std::thread> pool;
  const int64_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const int64_t b = t * chunk, e = std::min<int64_t>(n, b + chunk);
    if (b >= e) break;
    pool.emplace_back([=, &body] { body(b, e, t); });
  }
  for (auto& th : pool) th.join();
}

Pose pose_at(int fmt, const void* poses, const oracle_lattice* lat, int64_t p) {
  if (fmt == ORACLE_POSE_CS) {
    const double* q = static_cast<const double*>(poses) + 4 * p;
    return Pose::from_cs(q[0], q[1], q[2], q[3]);
  }
  const int64_t ix = p % lat->nx, iy = (p / lat->nx) % lat->ny, it = p / (static_cast<int64_t>(lat->nx) * lat->ny);
  const double tx = lat->x0 + (static_cast<double>(ix) * lat->dx);
  const double ty = lat->y0 + (static_cast<double>(iy) * lat->dy);
  return Pose::from_cs(lat->cs[2 * it], lat->cs[2 * it + 1], tx, ty);
}

/// Pre-cast range check on the continuous map-frame coordinates (the int cast is UB out of range).
bool coords_in_range(const Pose* T, const double* xy, size_t n, const MapView& m, bool path_b) {
  const double inv = 1. / m.resolution;
  for (size_t i = 0; i < n; ++i) {
    double qx = xy[2 * i], qy = xy[2 * i + 1];
    if (T) {
      if (path_b) {
        qx = (T->tx + (-m.origin_x)) + ((T->m00 * xy[2 * i]) + (T->m01 * xy[2 * i + 1]));
        qy = (T->ty + (-m.origin_y)) + ((T->m10 * xy[2 * i]) + (T->m11 * xy[2 * i + 1]));
      } else {
        qx = (T->tx + ((T->m00 * xy[2 * i]) + (T->m01 * xy[2 * i + 1]))) - m.origin_x;
        qy = (T->ty + ((T->m10 * xy[2 * i]) + (T->m11 * xy[2 * i + 1]))) - m.origin_y;
      }
    } else {
      qx -= m.origin_x;
      qy -= m.origin_y;
    }
    if (!(std::fabs(qx * inv) < kCoordLimit) || !(std::fabs(qy * inv) < kCoordLimit)) return false;
  }
  return true;
}

struct OutSink {
  uint8_t* is_free;
  double* cost;
  uint8_t* status;
  void fail(int64_t p, int st) const {
    if (is_free) is_free[p] = 0;
    if (cost) cost[p] = std::numeric_limits<double>::quiet_NaN();
    if (status) status[p] = static_cast<uint8_t>(st);
  }
};

template <class Walk>
void run_op(int op, Walk&& walk, Cell off, const MapView& m, const OutSink& out, int64_t p, Tally* t) {
  if (op == ORACLE_OP_IS_FREE) {
    const bool f = is_free_walk(walk, off, m, t);
    if (out.is_free) out.is_free[p] = f;
  } else if (op == ORACLE_OP_ACCUMULATE) {
    const double c = accumulate_walk(walk, off, m, t);
    if (out.cost) out.cost[p] = c;
    if (out.is_free) out.is_free[p] = c >= 0;
  } else {
    const double c = max_cost_walk(walk, off, m, t);
    if (out.cost) out.cost[p] = c;
    if (out.is_free) out.is_free[p] = c >= 0 && c != kLethal;
  }
}

}  // namespace

extern "C" {

int oracle_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }

// ---- generator utilities (tests of L0/L1/L2) -----------------------------------------------

int oracle_to_discrete(double res, const double* xy, int64_t n, int32_t* out) {
  try {
    const Cells c = to_discrete(res, xy, static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
      out[2 * i] = c[i].x;
      out[2 * i + 1] = c[i].y;
    }
    return 0;
  } catch (const std::invalid_argument&) {
    return -1;
  }
}

static int64_t emit(const Cells& c, int32_t* out, int64_t cap) {
  const int64_t n = static_cast<int64_t>(c.size());
  for (int64_t i = 0; i < std::min(n, cap); ++i) {
    out[2 * i] = c[i].x;
    out[2 * i + 1] = c[i].y;
  }
  return n;
}

int64_t oracle_ray_cells(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t* out, int64_t cap) {
  Cells c;
  walk_ray(make_ray(Cell{x0, y0}, Cell{x1, y1}), [&c](Cell q) {
    c.push_back(q);
    return true;
  });
  return emit(c, out, cap);
}

int64_t oracle_outline_cells(const int32_t* sparse, int64_t n, int32_t* out, int64_t cap) {
  Cells s(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) s[i] = Cell{sparse[2 * i], sparse[2 * i + 1]};
  return emit(dense_outline(s), out, cap);
}

/// `dense` = concatenated dense outlines (2 x sum(sizes), column-major); returns the number of area
/// cells, or -1 for std::logic_error.  *max_ranges receives the largest pre-merge row list.
int64_t oracle_area_cells(const int32_t* dense, const int64_t* sizes, int32_t n_outlines, int32_t* out, int64_t cap,
                          int32_t* max_ranges) {
  std::vector<Cells> outlines(static_cast<size_t>(n_outlines));
  int64_t at = 0;
  for (int32_t o = 0; o < n_outlines; ++o) {
    outlines[o].resize(static_cast<size_t>(sizes[o]));
    for (int64_t i = 0; i < sizes[o]; ++i, ++at) outlines[o][i] = Cell{dense[2 * at], dense[2 * at + 1]};
  }
  try {
    const Area a = make_area(outlines);
    if (max_ranges) *max_ranges = a.max_ranges_per_row;
    return emit(area_cells(a), out, cap);
  } catch (const std::logic_error&) {
    return -1;
  }
}

// ---- polygon x poses (a12, a13 over outline/area generators; ext max) -----------------------

/// Path A: cells = to_discrete(res, (pose * polygon) - origin) (costmap.cpp:71-85), then
/// outline_generator / area_generator, then is_free / accumulate_cost / max.
int oracle_polygon_batch_offset(const oracle_map* map, const double* poly_xy, int32_t n_vertices, int32_t shape, int32_t op,
                                int32_t pose_format, const void* poses, const oracle_lattice* lattice, int64_t first, int64_t count,
                                const int32_t* offset, uint8_t* is_free, double* cost, uint8_t* status, int64_t* cells_read, int32_t threads);

int oracle_polygon_batch(const oracle_map* map, const double* poly_xy, int32_t n_vertices, int32_t shape, int32_t op,
                         int32_t pose_format, const void* poses, const oracle_lattice* lattice, int64_t first,
                         int64_t count, uint8_t* is_free, double* cost, uint8_t* status, int64_t* cells_read,
                         int32_t threads) {
  return oracle_polygon_batch_offset(map, poly_xy, n_vertices, shape, op, pose_format, poses, lattice, first, count, nullptr, is_free, cost, status,
                                     cells_read, threads);
}

/// The same with a cell offset added to every generated cell: is_free / accumulate_cost(gen, offset, map)
/// (src/cover/impl/costmap.hpp:29-35,74-81).
int oracle_polygon_batch_offset(const oracle_map* map, const double* poly_xy, int32_t n_vertices, int32_t shape, int32_t op,
                                int32_t pose_format, const void* poses, const oracle_lattice* lattice, int64_t first, int64_t count,
                                const int32_t* offset, uint8_t* is_free, double* cost, uint8_t* status, int64_t* cells_read, int32_t threads) {
  if (!map || map->resolution <= 0) return -1;
  const Cell cell_off = offset ? Cell{offset[0], offset[1]} : Cell{0, 0};
  const MapView m = view(map);
  const OutSink out{is_free, cost, status};
  std::vector<int64_t> tallies(static_cast<size_t>(std::max(threads, 1)), 0);
  parallel_for(count, threads, [&](int64_t b, int64_t e, int tid) {
    Tally t;
    for (int64_t i = b; i < e; ++i) {
      Pose T{};
      const bool has_pose = pose_format != ORACLE_POSE_NONE;
      if (has_pose) T = pose_at(pose_format, poses, lattice, first + i);
      if (!coords_in_range(has_pose ? &T : nullptr, poly_xy, n_vertices, m, false)) {
        out.fail(i, ORACLE_COORD_RANGE);
        continue;
      }
      const Cells sparse = has_pose ? transform_path_a(T, poly_xy, n_vertices, m) : shift_path_a(poly_xy, n_vertices, m);
      try {
        int st = ORACLE_OK;
        if (shape == ORACLE_SHAPE_AREA) {
          const Area a = make_area({dense_outline(sparse)});
          if (a.max_ranges_per_row > 16) st = ORACLE_ROW_COMPLEXITY;
          run_op(op, [&a](auto&& f) { return walk_area(a, f); }, cell_off, m, out, i, &t);
        } else {
          const Outline o = make_outline(sparse);
          run_op(op, [&o](auto&& f) { return walk_outline(o, f); }, cell_off, m, out, i, &t);
        }
        if (status) status[i] = static_cast<uint8_t>(st);
      } catch (const std::logic_error&) {
        out.fail(i, ORACLE_LOGIC_ERROR);
      }
    }
    tallies[tid] = t.cells_read;
  });
  if (cells_read) {
    *cells_read = 0;
    for (int64_t v : tallies) *cells_read += v;
  }
  return 0;
}

/// Several outlines (holes): area_generator(res, polygon_vec) (generators.cpp:126-135) under path A.
int oracle_multi_polygon_batch(const oracle_map* map, const double* polys_xy, const int32_t* starts, int32_t n_outlines, int32_t op,
                               int32_t pose_format, const void* poses, const oracle_lattice* lattice, int64_t first, int64_t count,
                               uint8_t* is_free, double* cost, uint8_t* status, int64_t* cells_read, int32_t threads) {
  if (!map || map->resolution <= 0) return -1;
  const MapView m = view(map);
  const OutSink out{is_free, cost, status};
  std::vector<int64_t> tallies(static_cast<size_t>(std::max(threads, 1)), 0);
  parallel_for(count, threads, [&](int64_t b, int64_t e, int tid) {
    Tally t;
    for (int64_t i = b; i < e; ++i) {
      Pose T{};
      const bool has_pose = pose_format != ORACLE_POSE_NONE;
      if (has_pose) T = pose_at(pose_format, poses, lattice, first + i);
      bool in_range = true;
      for (int32_t o = 0; o < n_outlines; ++o)
        in_range = in_range && coords_in_range(has_pose ? &T : nullptr, polys_xy + 2 * starts[o], starts[o + 1] - starts[o], m, false);
      if (!in_range) {
        out.fail(i, ORACLE_COORD_RANGE);
        continue;
      }
      try {
        std::vector<Cells> dense;
        for (int32_t o = 0; o < n_outlines; ++o) {
          const double* xy = polys_xy + 2 * starts[o];
          const size_t n = static_cast<size_t>(starts[o + 1] - starts[o]);
          dense.push_back(dense_outline(has_pose ? transform_path_a(T, xy, n, m) : shift_path_a(xy, n, m)));
        }
        const Area a = make_area(dense);
        run_op(op, [&a](auto&& f) { return walk_area(a, f); }, Cell{0, 0}, m, out, i, &t);
        if (status) status[i] = static_cast<uint8_t>(a.max_ranges_per_row > 16 ? ORACLE_ROW_COMPLEXITY : ORACLE_OK);
      } catch (const std::logic_error&) {
        out.fail(i, ORACLE_LOGIC_ERROR);
      }
    }
    tallies[tid] = t.cells_read;
  });
  if (cells_read) {
    *cells_read = 0;
    for (int64_t v : tallies) *cells_read += v;
  }
  return 0;
}

// ---- rays (a4 + a10/a13) --------------------------------------------------------------------

/// segments: N x {x0,y0,x1,y1} cells; `offset` (2 ints, may be null) is added to every cell
/// (impl/costmap.hpp:29-35,72-81).
int oracle_rays_batch(const oracle_map* map, const int32_t* segments, int64_t count, const int32_t* offset, int32_t op,
                      uint8_t* is_free, double* cost, uint8_t* status, int64_t* cells_read, int32_t threads) {
  if (!map) return -1;
  const MapView m = view(map);
  const OutSink out{is_free, cost, status};
  const Cell off = offset ? Cell{offset[0], offset[1]} : Cell{0, 0};
  std::vector<int64_t> tallies(static_cast<size_t>(std::max(threads, 1)), 0);
  parallel_for(count, threads, [&](int64_t b, int64_t e, int tid) {
    Tally t;
    for (int64_t i = b; i < e; ++i) {
      const int32_t* s = segments + 4 * i;
      const Ray r = make_ray(Cell{s[0], s[1]}, Cell{s[2], s[3]});
      run_op(op, [&r](auto&& f) { return walk_ray(r, f); }, off, m, out, i, &t);
      if (status) status[i] = ORACLE_OK;
    }
    tallies[tid] = t.cells_read;
  });
  if (cells_read) {
    *cells_read = 0;
    for (int64_t v : tallies) *cells_read += v;
  }
  return 0;
}

// ---- cached cell lists x offsets (a11) ------------------------------------------------------

/// cells: 2 x M column-major (discrete_polygon); offsets: N x {x,y}.  costmap.cpp:59-69.
int oracle_cells_batch(const oracle_map* map, const int32_t* cells, int64_t n_cells, const int32_t* offsets, int64_t count,
                       int32_t op, uint8_t* is_free, double* cost, uint8_t* status, int64_t* cells_read, int32_t threads) {
  if (!map) return -1;
  const MapView m = view(map);
  const OutSink out{is_free, cost, status};
  Cells list(static_cast<size_t>(n_cells));
  for (int64_t i = 0; i < n_cells; ++i) list[i] = Cell{cells[2 * i], cells[2 * i + 1]};
  std::vector<int64_t> tallies(static_cast<size_t>(std::max(threads, 1)), 0);
  parallel_for(count, threads, [&](int64_t b, int64_t e, int tid) {
    Tally t;
    for (int64_t i = b; i < e; ++i) {
      const Cell off{offsets[2 * i], offsets[2 * i + 1]};
      auto walk = [&list](auto&& f) {
        for (const Cell& c : list)
          if (!f(c)) return false;
        return true;
      };
      run_op(op, walk, off, m, out, i, &t);
      if (status) status[i] = ORACLE_OK;
    }
    tallies[tid] = t.cells_read;
  });
  if (cells_read) {
    *cells_read = 0;
    for (int64_t v : tallies) *cells_read += v;
  }
  return 0;
}

// ---- split footprints (a14, a15, a16) -------------------------------------------------------

/// Packed polygon list: `xy` = all vertices column-major, `starts[k]..starts[k+1]` = polygon k.
static std::vector<Polygon> unpack(const double* xy, const int32_t* starts, int32_t n) {
  std::vector<Polygon> out(static_cast<size_t>(n));
  for (int32_t k = 0; k < n; ++k) out[k].xy.assign(xy + 2 * starts[k], xy + 2 * starts[k + 1]);
  return out;
}

/// generator_footprint::init + discrete_footprint::init (footprints.cpp:198-212,233-238).
/// Two-call protocol: pass null outputs to get the sizes.
int oracle_make_discrete_footprint(double res, const double* outline_xy, const int32_t* outline_starts, int32_t n_outlines,
                                   const double* area_xy, const int32_t* area_starts, int32_t n_areas,
                                   int32_t* outline_cells, int64_t* n_outline_cells, int32_t* area_cells_out,
                                   int64_t* n_area_cells) {
  try {
    ContinuousFootprint fp{unpack(outline_xy, outline_starts, n_outlines), unpack(area_xy, area_starts, n_areas)};
    const DiscreteFootprint d = make_discrete_footprint(res, fp);
    *n_outline_cells = static_cast<int64_t>(d.outlines.size());
    *n_area_cells = static_cast<int64_t>(d.areas.size());
    if (outline_cells) emit(d.outlines, outline_cells, *n_outline_cells);
    if (area_cells_out) emit(d.areas, area_cells_out, *n_area_cells);
    return 0;
  } catch (const std::invalid_argument&) {  // derives from logic_error: must come first
    return -1;
  } catch (const std::logic_error&) {
    return -2;
  }
}

/// discrete_footprint::is_free(offset, map) for a BANK of footprints (one per heading bin):
/// footprint k owns outline cells [o_starts[k], o_starts[k+1]) and area cells [a_starts[k], ...).
/// `which` (nullable -> 0) selects the footprint of each pose.  footprints.cpp:240-247.
int oracle_discrete_footprint_batch(const oracle_map* map, const int32_t* outline_cells, const int64_t* o_starts,
                                    const int32_t* area_cells_in, const int64_t* a_starts, int32_t n_footprints,
                                    const int32_t* offsets, const int32_t* which, int64_t count, uint8_t* is_free,
                                    uint8_t* status, int64_t* cells_read, int32_t threads) {
  if (!map) return -1;
  const MapView m = view(map);
  std::vector<DiscreteFootprint> bank(static_cast<size_t>(n_footprints));
  for (int32_t k = 0; k < n_footprints; ++k) {
    for (int64_t i = o_starts[k]; i < o_starts[k + 1]; ++i) bank[k].outlines.push_back(Cell{outline_cells[2 * i], outline_cells[2 * i + 1]});
    for (int64_t i = a_starts[k]; i < a_starts[k + 1]; ++i) bank[k].areas.push_back(Cell{area_cells_in[2 * i], area_cells_in[2 * i + 1]});
  }
  std::vector<int64_t> tallies(static_cast<size_t>(std::max(threads, 1)), 0);
  parallel_for(count, threads, [&](int64_t b, int64_t e, int tid) {
    Tally t;
    for (int64_t i = b; i < e; ++i) {
      const int32_t k = which ? which[i] : 0;
      is_free[i] = discrete_is_free(bank[static_cast<size_t>(k)], Cell{offsets[2 * i], offsets[2 * i + 1]}, m, &t);
      if (status) status[i] = ORACLE_OK;
    }
    tallies[tid] = t.cells_read;
  });
  if (cells_read) {
    *cells_read = 0;
    for (int64_t v : tallies) *cells_read += v;
  }
  return 0;
}

/// continuous_footprint::is_free(pose, map) — footprints.cpp:130-196 (path B transform).
int oracle_continuous_footprint_batch(const oracle_map* map, const double* outline_xy, const int32_t* outline_starts,
                                      int32_t n_outlines, const double* area_xy, const int32_t* area_starts, int32_t n_areas,
                                      int32_t pose_format, const void* poses, const oracle_lattice* lattice, int64_t first,
                                      int64_t count, uint8_t* is_free, uint8_t* status, int64_t* cells_read, int32_t threads) {
  if (!map || map->resolution <= 0) return -1;
  const MapView m = view(map);
  const ContinuousFootprint fp{unpack(outline_xy, outline_starts, n_outlines), unpack(area_xy, area_starts, n_areas)};
  const OutSink out{is_free, nullptr, status};
  std::vector<int64_t> tallies(static_cast<size_t>(std::max(threads, 1)), 0);
  parallel_for(count, threads, [&](int64_t b, int64_t e, int tid) {
    Tally t;
    for (int64_t i = b; i < e; ++i) {
      const Pose T = pose_at(pose_format, poses, lattice, first + i);
      bool ok = true;
      for (const Polygon& p : fp.outlines) ok = ok && coords_in_range(&T, p.xy.data(), p.n(), m, true);
      for (const Polygon& p : fp.areas) ok = ok && coords_in_range(&T, p.xy.data(), p.n(), m, true);
      if (!ok) {
        out.fail(i, ORACLE_COORD_RANGE);
        continue;
      }
      try {
        is_free[i] = continuous_is_free(fp, T, m, &t);
        if (status) status[i] = ORACLE_OK;
      } catch (const std::logic_error&) {
        out.fail(i, ORACLE_LOGIC_ERROR);
      }
    }
    tallies[tid] = t.cells_read;
  });
  if (cells_read) {
    *cells_read = 0;
    for (int64_t v : tallies) *cells_read += v;
  }
  return 0;
}

}  // extern "C"
