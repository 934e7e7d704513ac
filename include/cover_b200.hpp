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

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <exception>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
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
      ctx_(ctx), size_x_(size_x), size_y_(size_y), resolution_(resolution), origin_x_(origin_x), origin_y_(origin_y) {
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
  /// Extension (the step before the path): inflate the lethal cells in place; `cost_by_d2[d2]`, d2 = 0..radius_cells^2,
  /// is the layer's computeCost evaluated on the host for a cell sqrt(d2) cells from the nearest obstacle.
  void inflate(int radius_cells, const std::vector<uint8_t>& cost_by_d2) {
    if (cost_by_d2.size() != static_cast<size_t>(radius_cells) * radius_cells + 1) throw std::invalid_argument("cost_by_d2 must have radius_cells^2 + 1 entries");
    detail::check(ctx_.get(), cover_map_inflate(map_.get(), radius_cells, cost_by_d2.data()));
  }
  std::vector<uint8_t> download() const {
    std::vector<uint8_t> bytes(static_cast<size_t>(size_x_) * size_y_);
    detail::check(ctx_.get(), cover_map_download(map_.get(), bytes.data()));
    return bytes;
  }
  const cover_map* get() const { return map_.get(); }
  const context& ctx() const { return ctx_; }
  int size_x() const { return size_x_; }
  int size_y() const { return size_y_; }
  double resolution() const { return resolution_; }
  double origin_x() const { return origin_x_; }
  double origin_y() const { return origin_y_; }

private:
  context ctx_;
  std::shared_ptr<cover_map> map_;
  int size_x_, size_y_;
  double resolution_, origin_x_, origin_y_;
};

namespace detail {
inline cover_poses make_poses(const std::vector<pose>& poses, int32_t off_x = 0, int32_t off_y = 0) {
  cover_poses p{};
  p.format = COVER_POSE_CS, p.mem = COVER_MEM_HOST, p.first = 0, p.count = static_cast<int64_t>(poses.size()), p.data = poses.data();
  p.cell_offset[0] = off_x, p.cell_offset[1] = off_y;
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
/// The same sweep with bit-packed verdicts (bit i & 31 of word i >> 5 = pose i is free): 8x less to bring back.
/// `not_ok` (optional) receives the number of poses whose status is not OK.
inline std::vector<uint32_t> is_free_bits(polygon_view poly, const pose_lattice& lattice, const costmap& map, int64_t* not_ok = nullptr,
                                          int64_t first = 0, int64_t count = -1) {
  std::vector<double> cs;
  const cover_poses p = detail::make_poses(lattice, cs, first, count);
  std::vector<uint32_t> bits(static_cast<size_t>((p.count + 31) / 32));
  int64_t bad = 0;
  const cover_results out{COVER_MEM_HOST, nullptr, nullptr, nullptr, bits.data(), &bad};
  detail::check(map.ctx().get(), cover_area_is_free(map.ctx().get(), map.get(), poly.xy, poly.n, &p, &out));
  if (not_ok) *not_ok = bad;
  return bits;
}
/// accumulate_cost / outline checks over a lattice (status array instead of exceptions, like every `_nothrow`).
inline batch_result accumulate_cost_nothrow(polygon_view poly, const pose_lattice& lattice, const costmap& map, int64_t first = 0, int64_t count = -1) {
  std::vector<double> cs;
  return detail::run(cover_area_accumulate_cost, poly, detail::make_poses(lattice, cs, first, count), map, true);
}
inline batch_result outline_is_free_nothrow(polygon_view poly, const pose_lattice& lattice, const costmap& map, int64_t first = 0, int64_t count = -1) {
  std::vector<double> cs;
  return detail::run(cover_outline_is_free, poly, detail::make_poses(lattice, cs, first, count), map, false);
}
inline batch_result outline_accumulate_cost_nothrow(polygon_view poly, const pose_lattice& lattice, const costmap& map, int64_t first = 0, int64_t count = -1) {
  std::vector<double> cs;
  return detail::run(cover_outline_accumulate_cost, poly, detail::make_poses(lattice, cs, first, count), map, true);
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

// ---- the offset overloads: is_free / accumulate_cost(generator, offset, map) (impl/costmap.hpp:29-35,74-81) --------
// The offset (cells) is added to every cell the generator of the transformed polygon yields.

struct cell_shift {
  int32_t x, y;
};
inline batch_result is_free_nothrow(polygon_view poly, const std::vector<pose>& poses, cell_shift offset, const costmap& map) {
  return detail::run(cover_area_is_free, poly, detail::make_poses(poses, offset.x, offset.y), map, false);
}
inline batch_result accumulate_cost_nothrow(polygon_view poly, const std::vector<pose>& poses, cell_shift offset, const costmap& map) {
  return detail::run(cover_area_accumulate_cost, poly, detail::make_poses(poses, offset.x, offset.y), map, true);
}
inline batch_result outline_is_free_nothrow(polygon_view poly, const std::vector<pose>& poses, cell_shift offset, const costmap& map) {
  return detail::run(cover_outline_is_free, poly, detail::make_poses(poses, offset.x, offset.y), map, false);
}
inline batch_result outline_accumulate_cost_nothrow(polygon_view poly, const std::vector<pose>& poses, cell_shift offset, const costmap& map) {
  return detail::run(cover_outline_accumulate_cost, poly, detail::make_poses(poses, offset.x, offset.y), map, true);
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

// ---- areas with holes: area_generator(resolution, const polygon_vec&) (generators.cpp:126-135) -----

namespace detail {
inline batch_result run_multi(const std::vector<polygon>& outlines, int op, const std::vector<pose>& poses, const costmap& map, bool want_cost) {
  std::vector<double> xy;
  std::vector<int32_t> starts{0};
  for (const polygon& o : outlines) {
    xy.insert(xy.end(), o.v.begin(), o.v.end());
    starts.push_back(starts.back() + static_cast<int32_t>(o.cols()));
  }
  batch_result r;
  r.is_free.resize(poses.size());
  r.status.resize(poses.size());
  if (want_cost) r.cost.resize(poses.size());
  const cover_poses p = make_poses(poses);
  const cover_results out{COVER_MEM_HOST, r.is_free.data(), want_cost ? r.cost.data() : nullptr, r.status.data(), nullptr, nullptr};
  check(map.ctx().get(), cover_area_multi_check(map.ctx().get(), map.get(), xy.data(), starts.data(), static_cast<int32_t>(outlines.size()), op, &p, &out));
  return r;
}
}  // namespace detail

/// is_free over the area between several outlines (outer boundary + holes, even-odd), one verdict per pose.
inline std::vector<uint8_t> is_free(const std::vector<polygon>& outlines, const std::vector<pose>& poses, const costmap& map) {
  batch_result r = detail::run_multi(outlines, 0, poses, map, false);
  detail::rethrow_statuses(r.status);
  return std::move(r.is_free);
}
inline std::vector<double> accumulate_cost(const std::vector<polygon>& outlines, const std::vector<pose>& poses, const costmap& map) {
  batch_result r = detail::run_multi(outlines, 1, poses, map, true);
  detail::rethrow_statuses(r.status);
  return std::move(r.cost);
}

// ---- expand: to_area / to_outline (expand.hpp:68-80), one discrete_polygon per pose --------------------

namespace detail {
using expand_entry = int (*)(cover_ctx*, const cover_map*, const double*, int32_t, const cover_poses*, int32_t*, int64_t, int64_t*, uint8_t*, int32_t);
inline std::vector<discrete_polygon> expand(expand_entry fn, polygon_view poly, const std::vector<pose>& poses, const costmap& map) {
  const cover_poses p = make_poses(poses);
  std::vector<int64_t> starts(poses.size() + 1);
  std::vector<uint8_t> status(poses.size());
  check(map.ctx().get(), fn(map.ctx().get(), map.get(), poly.xy, poly.n, &p, nullptr, 0, starts.data(), status.data(), COVER_MEM_HOST));
  rethrow_statuses(status);
  std::vector<int32_t> cells(static_cast<size_t>(2 * starts.back()) + 2);
  check(map.ctx().get(), fn(map.ctx().get(), map.get(), poly.xy, poly.n, &p, cells.data(), starts.back(), starts.data(), status.data(), COVER_MEM_HOST));
  std::vector<discrete_polygon> out(poses.size());
  for (size_t i = 0; i < poses.size(); ++i) out[i].v.assign(cells.begin() + 2 * starts[i], cells.begin() + 2 * starts[i + 1]);
  return out;
}
}  // namespace detail

/// to_area(resolution, pose * polygon - origin): the dense cells in generator order (y ascending, x ascending).
inline std::vector<discrete_polygon> to_area(polygon_view poly, const std::vector<pose>& poses, const costmap& map) {
  return detail::expand(cover_area_expand, poly, poses, map);
}
/// to_outline(resolution, pose * polygon - origin): the outline cells in traversal order.
inline std::vector<discrete_polygon> to_outline(polygon_view poly, const std::vector<pose>& poses, const costmap& map) {
  return detail::expand(cover_outline_expand, poly, poses, map);
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

// ---- the split as POLYGONS: get_radius / continuous_footprint(radius, polygon) without boost::geometry --------------

/// get_radius(polygon) (footprints.cpp:55-77).
inline double get_radius(polygon_view poly) { return cover_get_radius(poly.xy, poly.n); }

struct polygon_split {
  std::vector<polygon> outlines, areas;
};
/// continuous_footprint(radius, polygon)'s constructor (footprints.cpp:79-125) on the host: (outlines, areas) for
/// `continuous_footprint(ctx, outlines, areas)`.  Throws std::invalid_argument for radius <= 0, like the reference.
inline polygon_split split_polygon(double radius, polygon_view poly, double slack = 0.0) {
  const int32_t max_vertices = 1 << 16, max_polygons = 256;
  std::vector<double> oxy(2 * static_cast<size_t>(max_vertices)), axy(2 * static_cast<size_t>(max_vertices));
  std::vector<int32_t> os(max_polygons + 1), as(max_polygons + 1);
  int32_t no = 0, na = 0;
  const int rc = cover_split_polygon_slack(radius, slack, poly.xy, poly.n, oxy.data(), os.data(), &no, axy.data(), as.data(), &na, max_vertices, max_polygons);
  if (rc == COVER_E_INVALID_ARG) throw std::invalid_argument("Radius must be positive");
  if (rc != COVER_OK) throw std::runtime_error("cover_split_polygon failed");
  polygon_split out;
  auto unpack = [](const std::vector<double>& xy, const std::vector<int32_t>& st, int32_t n, std::vector<polygon>& dst) {
    for (int32_t k = 0; k < n; ++k) {
      polygon p(st[static_cast<size_t>(k) + 1] - st[static_cast<size_t>(k)]);
      std::copy(xy.begin() + 2 * st[static_cast<size_t>(k)], xy.begin() + 2 * st[static_cast<size_t>(k) + 1], p.v.begin());
      dst.push_back(std::move(p));
    }
  };
  unpack(oxy, os, no, out.outlines);
  unpack(axy, as, na, out.areas);
  return out;
}
/// One-argument form (footprints.cpp:127-128): radius = get_radius(polygon) ("convex only" upstream).
inline polygon_split split_polygon(polygon_view poly) { return split_polygon(get_radius(poly), poly); }

// ---- grid split: continuous_footprint(radius, polygon) without boost::geometry ---------------------------
//
// The reference erodes / dilates the continuous polygon with boost::geometry (footprints.cpp:79-125) and then
// discretises per heading (discrete_footprint::init, footprints.cpp:233-238).  Here the split is made on the
// grid, from the exact to_area cells F of one heading:
//   D  = Euclidean distance (metres, cell centres) to the nearest cell outside F
//   E  = {c in F : D(c) > r}                    the always-inscribed core
//   dE = cells of E with a 4-neighbour outside E  -> OUTLINE cells (checked for 253 or 254)
//   A  = {c in F : dist(c, dE) > r}               -> AREA cells    (checked for 254)
// A lethal cell inside F is in A or within r of some q in dE, which an inflation layer with inscribed radius >= r
// marks 253/254: no miss by construction (the reference's regression, test/footprints.cpp:205-242); every q in dE
// is farther than r from any cell outside F: no false alarm.  Same construction, cell for cell, as
// cover_b200/split.py.

struct split_cells_result {
  discrete_polygon outline, area;  ///< both in row-major order (y ascending, then x)
};

namespace detail {
/// Exact squared Euclidean distance (cells^2) from every cell of a w x h grid to the nearest cell with feature != 0.
inline std::vector<int64_t> squared_distance(const std::vector<uint8_t>& feature, int w, int h) {
  const int64_t inf = std::numeric_limits<int64_t>::max() / 4;
  std::vector<int64_t> g(static_cast<size_t>(w) * h, inf), d(static_cast<size_t>(w) * h, inf);
  for (int x = 0; x < w; ++x) {  // per column: distance to the nearest feature cell of the same column
    int64_t run = inf;
    for (int y = 0; y < h; ++y) {
      run = feature[static_cast<size_t>(y) * w + x] ? 0 : (run >= inf ? inf : run + 1);
      g[static_cast<size_t>(y) * w + x] = run;
    }
    run = inf;
    for (int y = h - 1; y >= 0; --y) {
      run = feature[static_cast<size_t>(y) * w + x] ? 0 : (run >= inf ? inf : run + 1);
      int64_t& v = g[static_cast<size_t>(y) * w + x];
      v = std::min(v, run);
    }
  }
  for (int y = 0; y < h; ++y) {  // per row: min over columns of dx^2 + g^2, scanning outwards until dx^2 alone loses
    const int64_t* grow = &g[static_cast<size_t>(y) * w];
    for (int x = 0; x < w; ++x) {
      int64_t best = grow[x] >= inf ? inf : grow[x] * grow[x];
      for (int64_t dx = 1; dx * dx < best && (x - dx >= 0 || x + dx < w); ++dx) {
        if (x - dx >= 0 && grow[x - dx] < inf) best = std::min(best, dx * dx + grow[x - dx] * grow[x - dx]);
        if (x + dx < w && grow[x + dx] < inf) best = std::min(best, dx * dx + grow[x + dx] * grow[x + dx]);
      }
      d[static_cast<size_t>(y) * w + x] = best;
    }
  }
  return d;
}
}  // namespace detail

/// Splits one footprint's cell set F (any order, duplicates allowed) into skeleton-outline and remainder-area cells.
inline split_cells_result split_cells(cells_view cells, double resolution, double radius) {
  split_cells_result out;
  if (cells.n == 0) return out;
  if (!(radius > 0)) throw std::invalid_argument("Radius must be positive");  // footprints.cpp:86-87
  if (!(resolution > 0)) throw std::invalid_argument("The resolution must be positive");
  int32_t xmin = cells.xy[0], xmax = xmin, ymin = cells.xy[1], ymax = ymin;
  for (int64_t i = 0; i < cells.n; ++i) {
    xmin = std::min(xmin, cells.xy[2 * i]), xmax = std::max(xmax, cells.xy[2 * i]);
    ymin = std::min(ymin, cells.xy[2 * i + 1]), ymax = std::max(ymax, cells.xy[2 * i + 1]);
  }
  const int pad = static_cast<int>(std::ceil(radius / resolution)) + 2;
  const int x0 = xmin - pad, y0 = ymin - pad, w = xmax - x0 + pad + 1, h = ymax - y0 + pad + 1;
  const size_t n = static_cast<size_t>(w) * h;
  std::vector<uint8_t> inside(n, 0), mask(n);
  for (int64_t i = 0; i < cells.n; ++i) inside[static_cast<size_t>(cells.xy[2 * i + 1] - y0) * w + (cells.xy[2 * i] - x0)] = 1;
  auto farther_than_radius = [&](int64_t d2) { return std::sqrt(static_cast<double>(d2)) * resolution > radius; };
  auto emit = [&](const std::vector<uint8_t>& m, discrete_polygon& dst) {
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x)
        if (m[static_cast<size_t>(y) * w + x]) dst.v.push_back(x + x0), dst.v.push_back(y + y0);
  };
  for (size_t i = 0; i < n; ++i) mask[i] = !inside[i];
  const std::vector<int64_t> d_out = detail::squared_distance(mask, w, h);
  std::vector<uint8_t> core(n);
  bool any_core = false;
  for (size_t i = 0; i < n; ++i) any_core |= (core[i] = inside[i] && farther_than_radius(d_out[i])) != 0;
  if (!any_core) {  // radius too large: the whole footprint is scanned as area
    emit(inside, out.area);
    return out;
  }
  std::vector<uint8_t> ring(n, 0);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const size_t i = static_cast<size_t>(y) * w + x;
      if (!core[i]) continue;
      ring[i] = (y > 0 && !core[i - w]) || (y + 1 < h && !core[i + w]) || (x > 0 && !core[i - 1]) || (x + 1 < w && !core[i + 1]);
    }
  const std::vector<int64_t> d_ring = detail::squared_distance(ring, w, h);
  for (size_t i = 0; i < n; ++i) mask[i] = inside[i] && farther_than_radius(d_ring[i]);
  emit(ring, out.outline);
  emit(mask, out.area);
  return out;
}

/// One discrete_footprint per heading (the reference's per-heading cache of discrete_footprint::init): the polygon is
/// rotated about the map's origin cell, so the cell offsets passed to `is_free` are the cell of the robot's
/// reference point.  `headings[k]` needs only (c, s).
inline discrete_footprints heading_bank(const costmap& map, polygon_view poly, double radius, const std::vector<pose>& headings) {
  std::vector<pose> poses;
  for (const pose& h : headings) poses.push_back(pose{h.c, h.s, map.origin_x(), map.origin_y()});
  const std::vector<discrete_polygon> footprints = to_area(poly, poses, map);
  std::vector<discrete_polygon> outlines, areas;
  for (const discrete_polygon& f : footprints) {
    split_cells_result s = split_cells(f, map.resolution(), radius);
    outlines.push_back(std::move(s.outline));
    areas.push_back(std::move(s.area));
  }
  return discrete_footprints(map.ctx(), outlines, areas);
}

// ---- trajectory scoring (extension: the DWA-style caller after the pose checks) ---------------------------

struct trajectory_scores {
  std::vector<int32_t> first_blocked;  ///< first step whose verdict is 0 (`steps` if none)
  std::vector<double> score;           ///< first negative per-pose cost in step order, else the sum
};
/// Poses grouped as n_trajectories x steps (pose index = trajectory * steps + step); either input may be empty.
inline trajectory_scores trajectory_reduce(const context& ctx, const std::vector<uint8_t>& is_free, const std::vector<double>& cost, int32_t steps) {
  const size_t n = is_free.empty() ? cost.size() : is_free.size();
  if (steps <= 0 || n % static_cast<size_t>(steps) != 0 || (!is_free.empty() && !cost.empty() && is_free.size() != cost.size()))
    throw std::invalid_argument("poses must be n_trajectories x steps");
  trajectory_scores out;
  const int64_t n_traj = static_cast<int64_t>(n / steps);
  if (!is_free.empty()) out.first_blocked.resize(static_cast<size_t>(n_traj));
  if (!cost.empty()) out.score.resize(static_cast<size_t>(n_traj));
  detail::check(ctx.get(), cover_trajectory_reduce(ctx.get(), is_free.empty() ? nullptr : is_free.data(), cost.empty() ? nullptr : cost.data(), COVER_MEM_HOST,
                                                   n_traj, steps, out.first_blocked.empty() ? nullptr : out.first_blocked.data(),
                                                   out.score.empty() ? nullptr : out.score.data()));
  return out;
}

// ---- pose sharding across the GPUs of one box ---------------------------------------------------------
//
// Poses are independent units: every device gets a contiguous block of the pose range and a replica of the costmap,
// one host thread drives each device's context, and every block's results land in ITS slice of one host buffer.
// No collective, no exchange between devices.  (The same device may be listed several times: the shards then run as
// independent contexts = streams of that device.)

class multi_gpu {
public:
  /// `devices` empty = every device the process sees.
  explicit multi_gpu(std::vector<int> devices = {}) {
    if (devices.empty())
      for (int d = 0; d < cover_device_count(); ++d) devices.push_back(d);
    if (devices.empty()) throw std::runtime_error("cover_b200: no usable CUDA device (there is no CPU fallback)");
    for (int d : devices) ctxs_.emplace_back(d);
  }
  int size() const { return static_cast<int>(ctxs_.size()); }
  const context& ctx(int k) const { return ctxs_[static_cast<size_t>(k)]; }

  /// Replicates the costmap: one device copy per context, uploaded side by side.
  void set_costmap(int size_x, int size_y, double resolution, double origin_x, double origin_y, const uint8_t* char_map) {
    maps_.clear();
    for (const context& c : ctxs_) maps_.emplace_back(c, size_x, size_y, resolution, origin_x, origin_y, nullptr);
    upload(char_map);
  }
  void upload(const uint8_t* char_map) {
    each([&](int k) { maps_[static_cast<size_t>(k)].upload(char_map); });
  }
  const costmap& map(int k) const { return maps_.at(static_cast<size_t>(k)); }

  /// Contiguous, balanced block of [0, count) owned by shard k (the first count % size() shards get one item more);
  /// with `granule` > 1 the block borders are multiples of it (bit-packed verdicts: whole 32-pose words).
  std::pair<int64_t, int64_t> block(int64_t count, int k, int64_t granule = 1) const {
    const int64_t units = (count + granule - 1) / granule, n = size();
    const int64_t base = units / n, extra = units % n;
    const int64_t first = (k * base + std::min<int64_t>(k, extra)) * granule;
    const int64_t len = (base + (k < extra ? 1 : 0)) * granule;
    return {std::min(first, count), std::min(len, std::max<int64_t>(0, count - first))};
  }

  /// is_free(polygon, pose, map) for every pose: the pose vector is split, the verdict / status arrays are shared.
  batch_result is_free_nothrow(polygon_view poly, const std::vector<pose>& poses) const {
    batch_result r;
    r.is_free.resize(poses.size());
    r.status.resize(poses.size());
    each([&](int k) {
      const auto b = block(static_cast<int64_t>(poses.size()), k);
      if (b.second == 0) return;
      cover_poses p{};
      p.format = COVER_POSE_CS, p.mem = COVER_MEM_HOST, p.count = b.second, p.data = poses.data() + b.first;
      const cover_results out{COVER_MEM_HOST, r.is_free.data() + b.first, nullptr, r.status.data() + b.first, nullptr, nullptr};
      detail::check(ctxs_[static_cast<size_t>(k)].get(), cover_area_is_free(ctxs_[static_cast<size_t>(k)].get(), map(k).get(), poly.xy, poly.n, &p, &out));
    });
    return r;
  }
  /// The lattice sweep, sharded through cover_poses.first / count.
  batch_result is_free_nothrow(polygon_view poly, const pose_lattice& lattice) const {
    batch_result r;
    const int64_t n = lattice.count();
    r.is_free.resize(static_cast<size_t>(n));
    r.status.resize(static_cast<size_t>(n));
    each([&](int k) {
      const auto b = block(n, k);
      if (b.second == 0) return;
      std::vector<double> cs;
      const cover_poses p = detail::make_poses(lattice, cs, b.first, b.second);
      const cover_results out{COVER_MEM_HOST, r.is_free.data() + b.first, nullptr, r.status.data() + b.first, nullptr, nullptr};
      detail::check(ctxs_[static_cast<size_t>(k)].get(), cover_area_is_free(ctxs_[static_cast<size_t>(k)].get(), map(k).get(), poly.xy, poly.n, &p, &out));
    });
    return r;
  }
  /// The same sweep with bit-packed verdicts; `not_ok` (optional) = poses whose status is not OK, summed over the shards.
  std::vector<uint32_t> is_free_bits(polygon_view poly, const pose_lattice& lattice, int64_t* not_ok = nullptr) const {
    const int64_t n = lattice.count();
    std::vector<uint32_t> bits(static_cast<size_t>((n + 31) / 32));
    std::vector<int64_t> bad(static_cast<size_t>(size()), 0);
    each([&](int k) {
      const auto b = block(n, k, 32);
      if (b.second == 0) return;
      std::vector<double> cs;
      const cover_poses p = detail::make_poses(lattice, cs, b.first, b.second);
      const cover_results out{COVER_MEM_HOST, nullptr, nullptr, nullptr, bits.data() + b.first / 32, &bad[static_cast<size_t>(k)]};
      detail::check(ctxs_[static_cast<size_t>(k)].get(), cover_area_is_free(ctxs_[static_cast<size_t>(k)].get(), map(k).get(), poly.xy, poly.n, &p, &out));
    });
    if (not_ok) {
      *not_ok = 0;
      for (int64_t v : bad) *not_ok += v;
    }
    return bits;
  }

private:
  /// Runs body(k) for every shard on its own host thread; the first exception (if any) is rethrown here.
  template <class Body>
  void each(Body&& body) const {
    std::vector<std::exception_ptr> errors(static_cast<size_t>(size()));
    std::vector<std::thread> threads;
    for (int k = 0; k < size(); ++k)
      threads.emplace_back([&, k] {
        try {
          detail::check(ctxs_[static_cast<size_t>(k)].get(), cover_ctx_bind(ctxs_[static_cast<size_t>(k)].get()));  // this thread drives one device
          body(k);
        } catch (...) {
          errors[static_cast<size_t>(k)] = std::current_exception();
        }
      });
    for (std::thread& t : threads) t.join();
    for (const std::exception_ptr& e : errors)
      if (e) std::rethrow_exception(e);
  }
  std::vector<context> ctxs_;
  std::vector<costmap> maps_;
};

}  // namespace cover_b200
