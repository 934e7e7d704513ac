// cover_b200.hpp — header-only C++ host layer over the C ABI (cover_b200.h).
//
// It keeps the reference's surface for the hot path — the names, argument meaning and error behaviour
// of src/cover/costmap.hpp (is_free / accumulate_cost over ray / outline / area generators and cached
// discrete polygons) and src/cover/footprints.hpp (continuous_footprint / discrete_footprint::is_free)
// — and adds the batch dimension: MANY poses (or offsets) x ONE costmap per call.
//
//   reference (scalar)                                        this header (batched)
//   is_free(polygon, Isometry2d, Costmap2D)  costmap.cpp:78   is_free(polygon, poses, costmap)
//   is_free(polygon, Costmap2D)              costmap.cpp:71   is_free(polygon, costmap)
//   accumulate_cost(area_generator, map)     impl/costmap.hpp:65   accumulate_cost(polygon, poses, costmap)
//   is_free(outline_generator, map)          impl/costmap.hpp:23   outline_is_free(polygon, poses, costmap)
//   is_free(ray_generator[, offset], map)    impl/costmap.hpp:23-35   ray_is_free(segments[, offset], costmap)
//   is_free(discrete_polygon, offset, map)   costmap.cpp:64   discrete_polygon::is_free(offsets, costmap)
//   discrete_footprint::is_free(offset, map) footprints.cpp:240   discrete_footprints::is_free(offsets, which, costmap)
//   continuous_footprint::is_free(pose, map) footprints.cpp:130   continuous_footprint::is_free(poses, costmap)
//
// Types are deliberately Eigen-free but Eigen-compatible: any matrix type with `.data()` (column-major
// 2 x N doubles or ints) and `.cols()` — `cover::polygon`, `cover::discrete_polygon` — can be passed where
// a `polygon_view` / `cells_view` is expected.
//
// Errors: exactly the reference's exception types.  A non-positive resolution throws
// std::invalid_argument (base.hpp:54-55); a pose whose area outline is open throws std::logic_error
// ("Even number of line-pairs expected", generators.cpp:246-249) from the throwing overloads — the
// `*_nothrow` variants return the per-pose status array instead, which is what a batch caller wants.
// CUDA failures throw std::runtime_error.  There is no CPU fallback.
#pragma once

#include "cover_b200.h"

#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace cover_b200 {

/// Borrowed view of a 2 x N column-major double matrix (cover::polygon, base.hpp:26).
struct polygon_view {
  const double* xy = nullptr;
  int n = 0;
  polygon_view() = default;
  polygon_view(const double* p, int cols) : xy(p), n(cols) {}
  template <class Matrix, class = decltype(std::declval<const Matrix&>().cols())>
  polygon_view(const Matrix& m) : xy(m.data()), n(static_cast<int>(m.cols())) {}  // NOLINT: implicit on purpose
};

/// Borrowed view of a 2 x N column-major int matrix (cover::discrete_polygon, base.hpp:30).
struct cells_view {
  const int32_t* xy = nullptr;
  int64_t n = 0;
  cells_view() = default;
  cells_view(const int32_t* p, int64_t cols) : xy(p), n(cols) {}
  template <class Matrix, class = decltype(std::declval<const Matrix&>().cols())>
  cells_view(const Matrix& m) : xy(m.data()), n(static_cast<int64_t>(m.cols())) {}  // NOLINT
};

/// Minimal owning 2 x N matrix for callers without Eigen.
template <class T>
struct matrix2 {
  std::vector<T> v;
  matrix2() = default;
  explicit matrix2(int64_t cols) : v(static_cast<size_t>(2 * cols)) {}
  int64_t cols() const { return static_cast<int64_t>(v.size() / 2); }
  const T* data() const { return v.data(); }
  T* data() { return v.data(); }
  T& operator()(int r, int64_t c) { return v[static_cast<size_t>(2 * c + r)]; }
  const T& operator()(int r, int64_t c) const { return v[static_cast<size_t>(2 * c + r)]; }
};
using polygon = matrix2<double>;
using discrete_polygon = matrix2<int32_t>;

/// One pose = Eigen::Isometry2d{Translation2d(x, y) * Rotation2Dd(theta)} with cos/sin evaluated on the host.
struct pose {
  double c, s, x, y;
  static pose from_angle(double x, double y, double theta) { return pose{std::cos(theta), std::sin(theta), x, y}; }
};
static_assert(sizeof(pose) == 4 * sizeof(double), "pose must be 4 packed doubles (COVER_POSE_CS)");

/// (x, y, heading) lattice expanded on the device; see cover_b200.h for the index rule.
struct pose_lattice {
  double x0, dx;
  int nx;
  double y0, dy;
  int ny;
  std::vector<pose> headings;  ///< only c, s are used
  int64_t count() const { return static_cast<int64_t>(nx) * ny * static_cast<int64_t>(headings.size()); }
};

struct batch_result {
  std::vector<uint8_t> is_free;
  std::vector<double> cost;
  std::vector<uint8_t> status;
};

namespace detail {
inline void check(cover_ctx* ctx, int rc) {
  if (rc == COVER_OK) return;
  const std::string msg = ctx ? cover_last_error(ctx) : "cover_b200 error";
  if (rc == COVER_E_INVALID_ARG) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}
inline void rethrow_statuses(const std::vector<uint8_t>& status) {
  for (uint8_t s : status) {
    if (s == COVER_STATUS_LOGIC_ERROR) throw std::logic_error("Even number of line-pairs expected");
    if (s == COVER_STATUS_ROW_COMPLEXITY) throw std::runtime_error("more than 16 ranges on one scan row");
    if (s == COVER_STATUS_COORD_RANGE) throw std::out_of_range("cell coordinate outside +-2^22");
  }
}
}  // namespace detail

/// One GPU + stream (cover_ctx).  Single-threaded; one per device.
class context {
public:
  explicit context(int device = 0) {
    cover_ctx* c = nullptr;
    if (cover_ctx_create(device, &c) != COVER_OK) throw std::runtime_error("cover_b200: no usable CUDA device (there is no CPU fallback)");
    ctx_.reset(c, cover_ctx_destroy);
  }
  cover_ctx* get() const { return ctx_.get(); }
  void synchronize() const { detail::check(get(), cover_ctx_synchronize(get())); }
  void* stream() const { return cover_ctx_stream(get()); }

private:
  std::shared_ptr<cover_ctx> ctx_;
};

/// Device copy of a costmap_2d::Costmap2D (size, resolution, origin, getCharMap()).
class costmap {
public:
  costmap(const context& ctx, int size_x, int size_y, double resolution, double origin_x, double origin_y, const uint8_t* char_map) :
      ctx_(ctx) {
    cover_map* m = nullptr;
    detail::check(ctx.get(), cover_map_create(ctx.get(), size_x, size_y, resolution, origin_x, origin_y, &m));
    map_.reset(m, cover_map_destroy);
    if (char_map) upload(char_map);
  }
  void upload(const uint8_t* char_map) { detail::check(ctx_.get(), cover_map_upload(map_.get(), char_map, COVER_MEM_HOST)); }
  /// Queues the copy and returns: `char_map` must stay unchanged until the next check has returned its results.
  void upload_async(const uint8_t* char_map) { detail::check(ctx_.get(), cover_map_upload_async(map_.get(), char_map)); }
  void upload_window(const uint8_t* char_map, int x0, int y0, int w, int h) {
    detail::check(ctx_.get(), cover_map_upload_window(map_.get(), char_map, x0, y0, w, h));
  }
  const cover_map* get() const { return map_.get(); }
  const context& ctx() const { return ctx_; }

private:
  context ctx_;
  std::shared_ptr<cover_map> map_;
};

namespace detail {
inline cover_poses make_poses(const std::vector<pose>& poses) {
  cover_poses p{};
  p.format = COVER_POSE_CS, p.mem = COVER_MEM_HOST, p.first = 0, p.count = static_cast<int64_t>(poses.size()), p.data = poses.data();
  return p;
}
inline cover_poses make_poses(const pose_lattice& lat, std::vector<double>& cs, int64_t first, int64_t count) {
  cs.resize(2 * lat.headings.size());
  for (size_t k = 0; k < lat.headings.size(); ++k) cs[2 * k] = lat.headings[k].c, cs[2 * k + 1] = lat.headings[k].s;
  cover_poses p{};
  p.format = COVER_POSE_LATTICE, p.mem = COVER_MEM_HOST, p.first = first, p.count = count < 0 ? lat.count() - first : count;
  p.lattice = cover_lattice{lat.x0, lat.dx, lat.y0, lat.dy, cs.data(), lat.nx, lat.ny, static_cast<int32_t>(lat.headings.size())};
  return p;
}
inline cover_poses no_pose() {
  cover_poses p{};
  p.format = COVER_POSE_NONE, p.count = 1;
  return p;
}
using polygon_entry = int (*)(cover_ctx*, const cover_map*, const double*, int32_t, const cover_poses*, const cover_results*);
inline batch_result run(polygon_entry fn, polygon_view poly, const cover_poses& p, const costmap& map, bool want_cost) {
  batch_result r;
  r.is_free.resize(static_cast<size_t>(p.count));
  r.status.resize(static_cast<size_t>(p.count));
  if (want_cost) r.cost.resize(static_cast<size_t>(p.count));
  const cover_results out{COVER_MEM_HOST, r.is_free.data(), want_cost ? r.cost.data() : nullptr, r.status.data(), nullptr, nullptr};
  check(map.ctx().get(), fn(map.ctx().get(), map.get(), poly.xy, poly.n, &p, &out));
  return r;
}
}  // namespace detail

// ---- is_free(polygon, pose, map) / accumulate_cost over the area generator --------------------------

/// Batched is_free(const polygon&, const Eigen::Isometry2d&, const Costmap2D&) (costmap.cpp:78-85).
inline batch_result is_free_nothrow(polygon_view poly, const std::vector<pose>& poses, const costmap& map) {
  return detail::run(cover_area_is_free, poly, detail::make_poses(poses), map, false);
}
inline std::vector<uint8_t> is_free(polygon_view poly, const std::vector<pose>& poses, const costmap& map) {
  batch_result r = is_free_nothrow(poly, poses, map);
  detail::rethrow_statuses(r.status);
  return std::move(r.is_free);
}
/// Lattice form (the DWA sweep); `first`/`count` select a sub-range for pose sharding across GPUs.
inline batch_result is_free_nothrow(polygon_view poly, const pose_lattice& lattice, const costmap& map, int64_t first = 0, int64_t count = -1) {
  std::vector<double> cs;
  return detail::run(cover_area_is_free, poly, detail::make_poses(lattice, cs, first, count), map, false);
}
/// is_free(const polygon&, const Costmap2D&) (costmap.cpp:71-76).
inline bool is_free(polygon_view poly, const costmap& map) {
  const batch_result r = detail::run(cover_area_is_free, poly, detail::no_pose(), map, false);
  detail::rethrow_statuses(r.status);
  return r.is_free[0] != 0;
}
/// Batched accumulate_cost(area_generator(resolution, pose * polygon - origin), map).
inline batch_result accumulate_cost_nothrow(polygon_view poly, const std::vector<pose>& poses, const costmap& map) {
  return detail::run(cover_area_accumulate_cost, poly, detail::make_poses(poses), map, true);
}
inline std::vector<double> accumulate_cost(polygon_view poly, const std::vector<pose>& poses, const costmap& map) {
  batch_result r = accumulate_cost_nothrow(poly, poses, map);
  detail::rethrow_statuses(r.status);
  return std::move(r.cost);
}
/// Extension: max of detail::get_cell_cost over the area (-3 when a cell is outside the map).
inline batch_result max_cost_nothrow(polygon_view poly, const std::vector<pose>& poses, const costmap& map) {
  return detail::run(cover_area_max_cost, poly, detail::make_poses(poses), map, true);
}

// ---- outline generator -----------------------------------------------------------------------------

inline std::vector<uint8_t> outline_is_free(polygon_view poly, const std::vector<pose>& poses, const costmap& map) {
  batch_result r = detail::run(cover_outline_is_free, poly, detail::make_poses(poses), map, false);
  detail::rethrow_statuses(r.status);
  return std::move(r.is_free);
}
inline std::vector<double> outline_accumulate_cost(polygon_view poly, const std::vector<pose>& poses, const costmap& map) {
  batch_result r = detail::run(cover_outline_accumulate_cost, poly, detail::make_poses(poses), map, true);
  detail::rethrow_statuses(r.status);
  return std::move(r.cost);
}

// ---- ray generator ---------------------------------------------------------------------------------

struct segment {
  int32_t x0, y0, x1, y1;  ///< ray_generator(begin, end): end exclusive
};
inline std::vector<uint8_t> ray_is_free(const std::vector<segment>& rays, const costmap& map, const int32_t* offset = nullptr) {
  std::vector<uint8_t> free_(rays.size());
  const cover_results out{COVER_MEM_HOST, free_.data(), nullptr, nullptr, nullptr, nullptr};
  detail::check(map.ctx().get(), cover_ray_is_free(map.ctx().get(), map.get(), &rays.data()->x0, COVER_MEM_HOST,
                                                   static_cast<int64_t>(rays.size()), offset, &out));
  return free_;
}
inline std::vector<double> ray_accumulate_cost(const std::vector<segment>& rays, const costmap& map, const int32_t* offset = nullptr) {
  std::vector<double> cost(rays.size());
  const cover_results out{COVER_MEM_HOST, nullptr, cost.data(), nullptr, nullptr, nullptr};
  detail::check(map.ctx().get(), cover_ray_accumulate_cost(map.ctx().get(), map.get(), &rays.data()->x0, COVER_MEM_HOST,
                                                           static_cast<int64_t>(rays.size()), offset, &out));
  return cost;
}

// ---- cached discrete polygon: is_free(dp, offset, map) (costmap.cpp:64-69) ---------------------------

struct cell_offset {
  int32_t x, y;
};

class cached_cells {
public:
  cached_cells(const context& ctx, cells_view cells) : ctx_(ctx) {
    cover_cells* c = nullptr;
    detail::check(ctx.get(), cover_cells_create(ctx.get(), cells.xy, cells.n, &c));
    cells_.reset(c, cover_cells_destroy);
  }
  std::vector<uint8_t> is_free(const std::vector<cell_offset>& offsets, const costmap& map) const {
    std::vector<uint8_t> free_(offsets.size());
    const cover_results out{COVER_MEM_HOST, free_.data(), nullptr, nullptr, nullptr, nullptr};
    detail::check(ctx_.get(), cover_cells_is_free(ctx_.get(), map.get(), cells_.get(), &offsets.data()->x, COVER_MEM_HOST,
                                                  static_cast<int64_t>(offsets.size()), &out));
    return free_;
  }

private:
  context ctx_;
  std::shared_ptr<cover_cells> cells_;
};

// ---- split footprints ------------------------------------------------------------------------------

/// K discrete_footprint's (outline cells + area cells each), e.g. one per heading bin.
class discrete_footprints {
public:
  discrete_footprints(const context& ctx, const std::vector<discrete_polygon>& outlines, const std::vector<discrete_polygon>& areas) : ctx_(ctx) {
    if (outlines.size() != areas.size() || outlines.empty()) throw std::invalid_argument("one outline list and one area list per footprint");
    std::vector<int32_t> oc, ac;
    std::vector<int64_t> os{0}, as{0};
    for (size_t k = 0; k < outlines.size(); ++k) {
      oc.insert(oc.end(), outlines[k].v.begin(), outlines[k].v.end());
      ac.insert(ac.end(), areas[k].v.begin(), areas[k].v.end());
      os.push_back(os.back() + outlines[k].cols());
      as.push_back(as.back() + areas[k].cols());
    }
    cover_footprints* f = nullptr;
    detail::check(ctx.get(), cover_footprints_create(ctx.get(), oc.data(), os.data(), ac.data(), as.data(), static_cast<int32_t>(outlines.size()), &f));
    fps_.reset(f, cover_footprints_destroy);
  }
  /// discrete_footprint::is_free(offset, map) for every (offset, which[i]) pair (footprints.cpp:240-247).
  std::vector<uint8_t> is_free(const std::vector<cell_offset>& offsets, const std::vector<int32_t>& which, const costmap& map) const {
    if (!which.empty() && which.size() != offsets.size()) throw std::invalid_argument("which must be empty or one entry per offset");
    std::vector<uint8_t> free_(offsets.size());
    const cover_results out{COVER_MEM_HOST, free_.data(), nullptr, nullptr, nullptr, nullptr};
    detail::check(ctx_.get(), cover_footprints_is_free(ctx_.get(), map.get(), fps_.get(), &offsets.data()->x, which.empty() ? nullptr : which.data(),
                                                       COVER_MEM_HOST, static_cast<int64_t>(offsets.size()), &out));
    return free_;
  }

private:
  context ctx_;
  std::shared_ptr<cover_footprints> fps_;
};

/// continuous_footprint given as (outlines, areas) — the reference's public constructor (footprints.hpp:136).
class continuous_footprint {
public:
  continuous_footprint(const context& ctx, const std::vector<polygon>& outlines, const std::vector<polygon>& areas) : ctx_(ctx) {
    std::vector<double> oxy, axy;
    std::vector<int32_t> os{0}, as{0};
    for (const polygon& p : outlines) {
      oxy.insert(oxy.end(), p.v.begin(), p.v.end());
      os.push_back(os.back() + static_cast<int32_t>(p.cols()));
    }
    for (const polygon& p : areas) {
      axy.insert(axy.end(), p.v.begin(), p.v.end());
      as.push_back(as.back() + static_cast<int32_t>(p.cols()));
    }
    cover_split* s = nullptr;
    detail::check(ctx.get(), cover_split_create(ctx.get(), oxy.data(), os.data(), static_cast<int32_t>(outlines.size()), axy.data(), as.data(),
                                                static_cast<int32_t>(areas.size()), &s));
    split_.reset(s, cover_split_destroy);
  }
  /// continuous_footprint::is_free(pose, map) for every pose (footprints.cpp:130-196); throws like the scalar call.
  std::vector<uint8_t> is_free(const std::vector<pose>& poses, const costmap& map) const {
    std::vector<uint8_t> free_(poses.size()), status(poses.size());
    const cover_poses p = detail::make_poses(poses);
    const cover_results out{COVER_MEM_HOST, free_.data(), nullptr, status.data(), nullptr, nullptr};
    detail::check(ctx_.get(), cover_split_is_free(ctx_.get(), map.get(), split_.get(), &p, &out));
    detail::rethrow_statuses(status);
    return free_;
  }

private:
  context ctx_;
  std::shared_ptr<cover_split> split_;
};

}  // namespace cover_b200
