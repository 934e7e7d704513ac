/* cover_b200.h — C ABI of the B200-native batched footprint collision checker.
 *
 * This is the drop-in boundary for ONE hot path of magazino/cover: cell generators
 * (src/cover/generators.hpp) feeding the costmap pose checks (src/cover/costmap.hpp) and the split
 * footprint checks (src/cover/footprints.hpp), batched as MANY POSES x ONE COSTMAP.
 * The reference has no FFI today (it is a header-heavy C++ library); each entry below names the
 * reference function it batches (paths relative to the reference tree) and INTEGRATION.md shows the
 * C++ overload a maintainer adds next to that function to call it.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns 0 (COVER_OK) or a negative cover_error;
 *    nothing throws across this boundary; cover_last_error(ctx) holds the message.
 *  - polygons are 2 x V doubles, COLUMN-major (x0,y0,x1,y1,...) = Eigen's layout of cover::polygon
 *    (src/cover/base.hpp:26); cell lists are 2 x N int32 column-major = cover::discrete_polygon
 *    (src/cover/base.hpp:30).
 *  - poses: COVER_POSE_CS = count x {cos, sin, tx, ty} doubles, i.e. the Eigen::Isometry2d
 *    Translation2d(tx,ty) * Rotation2Dd with the trig evaluated BY THE CALLER on the host (device
 *    sin/cos differ from glibc in the last ulp); COVER_POSE_LATTICE = a descriptor expanded on the
 *    device: pose index p = (it*ny + iy)*nx + ix, t = (x0 + ix*dx, y0 + iy*dy) (multiply then add,
 *    no FMA), heading (cos,sin) = cs[it]; COVER_POSE_NONE = the no-pose overloads (one check).
 *  - outputs are caller-owned arrays (sizes in cover_results), host or device (cover_results.mem); any
 *    of them may be NULL.  status[i] is a cover_status; for a status other than COVER_STATUS_OK the
 *    verdict is 0 and the cost NaN (the scalar reference would have thrown, see each status).
 *  - a cover_ctx is single-threaded and owns one CUDA stream; all work of a call is enqueued on it
 *    and, when any argument or result lives on the host, completed before the call returns.  With
 *    device-resident arguments and results the call is asynchronous (cover_ctx_synchronize).
 *  - maps, cell lists and footprints are immutable once uploaded and may be shared by calls on the
 *    same context.  One context per GPU; contexts are independent (pose sharding across GPUs is done
 *    by the caller: replicate the map, split the pose range, see cover_poses.first).
 *  - there is NO CPU fallback: without a CUDA device every entry fails with COVER_E_CUDA.
 */
#ifndef COVER_B200_H
#define COVER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COVER_B200_ABI_VERSION 2

typedef struct cover_ctx cover_ctx;
typedef struct cover_map cover_map;
typedef struct cover_cells cover_cells;
typedef struct cover_footprints cover_footprints;
typedef struct cover_split cover_split;

typedef enum cover_error {
  COVER_OK = 0,
  COVER_E_INVALID_ARG = -1, /* std::invalid_argument in the reference (e.g. resolution <= 0, base.hpp:54) */
  COVER_E_CUDA = -2,
  COVER_E_NOMEM = -3,
  COVER_E_UNSUPPORTED = -4
} cover_error;

typedef enum cover_status {
  COVER_STATUS_OK = 0,
  COVER_STATUS_LOGIC_ERROR = 1,    /* area_generator threw std::logic_error (generators.cpp:246-249) */
  COVER_STATUS_ROW_COMPLEXITY = 2, /* > 16 ranges on one scan row: std::sort is no longer insertion-stable
                                      (generators.cpp:236-241); the result would depend on libstdc++ */
  COVER_STATUS_COORD_RANGE = 3     /* |cell coordinate| >= 2^22: outside any costmap (int cast UB upstream) */
} cover_status;

/* COVER_MEM_HOST_ASYNC (cover_results only): host arrays - page-locked, or the copies are not asynchronous - that the
 * call fills WITHOUT waiting: it returns once kernels and downloads are queued on the context's stream, the results (and
 * the staging buffers they pass through) are valid after cover_ctx_synchronize(ctx) and until the context's next call.
 * One host thread can so keep several contexts (= several planning cycles) in flight. */
typedef enum cover_mem { COVER_MEM_HOST = 0, COVER_MEM_DEVICE = 1, COVER_MEM_HOST_ASYNC = 2 } cover_mem;
typedef enum cover_pose_format { COVER_POSE_CS = 0, COVER_POSE_LATTICE = 1, COVER_POSE_NONE = 2 } cover_pose_format;

typedef struct cover_lattice {
  double x0, dx; /* tx = x0 + ix*dx */
  double y0, dy; /* ty = y0 + iy*dy */
  const double* cs; /* HOST pointer: ntheta x {cos, sin} */
  int32_t nx, ny, ntheta;
} cover_lattice;

typedef struct cover_poses {
  int32_t format;        /* cover_pose_format */
  int32_t mem;           /* cover_mem of `data` (COVER_POSE_CS only) */
  int64_t first;         /* LATTICE: index of the first pose of this batch (pose sharding); else 0 */
  int64_t count;         /* number of poses in this batch */
  const void* data;      /* CS: count x 4 doubles */
  cover_lattice lattice; /* LATTICE */
  int32_t cell_offset[2]; /* added to every generated cell: the offset overloads is_free / accumulate_cost(gen, offset, map)
                             (src/cover/impl/costmap.hpp:29-35,74-81) for the outline / area generator of the transformed
                             polygon; {0, 0} for the plain overloads */
} cover_poses;

typedef struct cover_results {
  int32_t mem;         /* cover_mem of every pointer below */
  uint8_t* is_free;    /* count bytes, 1 = free */
  double* cost;        /* count doubles: accumulate_cost / max_cost value */
  uint8_t* status;     /* count bytes: cover_status */
  uint32_t* free_bits; /* (count + 31) / 32 words: bit (i & 31) of word (i >> 5) = is_free[i]; 8x less
                          device->host traffic than is_free for verdict-only sweeps (1e8 poses = 12.5 MB) */
  int64_t* not_ok;     /* ONE counter: number of poses whose status is not COVER_STATUS_OK, so that a caller
                          who skips `status` still learns whether the scalar reference would have thrown */
} cover_results;

/* ---- context ------------------------------------------------------------------------------------ */
int cover_abi_version(void);
/* number of CUDA devices this process sees (0 without a driver): one cover_ctx per device for pose sharding. */
int cover_device_count(void);
int cover_ctx_create(int device, cover_ctx** out);
void cover_ctx_destroy(cover_ctx* ctx);
const char* cover_last_error(const cover_ctx* ctx);
void* cover_ctx_stream(cover_ctx* ctx); /* the cudaStream_t all work of this context is enqueued on */
int cover_ctx_synchronize(cover_ctx* ctx);
/* Makes the context's device the calling thread's current CUDA device.  Every entry does that for the duration of the
 * call and restores the previous device afterwards; a host thread that drives ONE context (pose sharding: one thread
 * per GPU) should call this once first - otherwise every call switches from the thread's default device 0 and back,
 * which also creates a CUDA context on device 0 in that process. */
int cover_ctx_bind(cover_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t cover_ctx_launch_count(const cover_ctx* ctx);

/* Measurement helper: best-of-10 streaming read bandwidth (16-byte vectors, L1 bypassed) over a scratch buffer of
 * `bytes`, `reps` passes per launch.  16 MB measures the L2 -> SM path (the buffer stays in the 126 MB L2), 1 GB the
 * HBM.  bench.py reports the costmap kernels against both. */
int cover_probe_read_bandwidth(cover_ctx* ctx, int64_t bytes, int32_t reps, double* gb_per_s);

/* ---- costmap: stands in for `const costmap_2d::Costmap2D&` ----------------------------------------
 * size/resolution/origin = getSizeInCellsX/Y, getResolution, getOriginX/Y; bytes = getCharMap(),
 * row-major, index = y*size_x + x (call sites src/cover/costmap.hpp:52,89,105,122).
 * cover_map_create fails with COVER_E_INVALID_ARG for resolution <= 0 (base.hpp:54-55). */
int cover_map_create(cover_ctx* ctx, int32_t size_x, int32_t size_y, double resolution, double origin_x, double origin_y,
                     cover_map** out);
/* A new image (size_y x size_x bytes, row-major, unpadded).  mem = COVER_MEM_HOST: returns when the copy is done.
 * mem = COVER_MEM_DEVICE: the copy is queued on the context's stream (cover_ctx_stream) and the call returns at once -
 * `data` must stay valid and unchanged until that stream has passed it (cover_ctx_synchronize, a later call that
 * returns results to the host, or an event of the caller's on that stream).  Everything derived from the old image
 * (bit planes, prefix words, maximum planes) is invalidated and rebuilt by the next call that needs it. */
int cover_map_upload(cover_map* map, const uint8_t* data, int32_t mem);
/* Same image, HOST source, without waiting for the copy: it is queued as row bands on a second stream and every
 * later call that reads the map is ordered after the bands it needs on the device - a lattice check of a small
 * footprint starts on the first rows of the map while the rest is still in flight (the per-planning-cycle
 * "new costmap, then sweep" sequence).  `host_data` (page-locked memory for a truly asynchronous copy) must stay
 * unchanged until a later call that returns results to the host (COVER_MEM_HOST, not ..._ASYNC), or
 * cover_ctx_synchronize, has returned: both wait for the whole image to have landed, also when the check itself only
 * needed a few bands of it. */
int cover_map_upload_async(cover_map* map, const uint8_t* host_data);
/* re-upload the window [x0,x0+w) x [y0,y0+h) from a full-size HOST image (dirty-rectangle update). */
int cover_map_upload_window(cover_map* map, const uint8_t* full_image, int32_t x0, int32_t y0, int32_t w, int32_t h);
/* EXTENSION (the step before the path; costmap_2d::InflationLayer is not part of the reference):
 * inflate the LETHAL cells of the device map in place.  cost_by_d2[d2], d2 = 0..radius_cells^2, is the cost of a
 * cell whose nearest lethal cell is sqrt(d2) cells away (HOST array, built by the caller with the layer's
 * computeCost so that no device exp() is involved); it is combined with the old cost like
 * InflationLayer::updateCosts (NO_INFORMATION stays unless the new cost is >= 253, else max). */
int cover_map_inflate(cover_map* map, int32_t radius_cells, const uint8_t* cost_by_d2);
/* device map -> size_y x size_x host bytes (row-major, unpadded) */
int cover_map_download(const cover_map* map, uint8_t* data);
void cover_map_destroy(cover_map* map);

/* ---- polygon x poses: path A transform ((pose*polygon) - origin, src/cover/costmap.cpp:71-85) ---- */
/* is_free(const polygon&, const Eigen::Isometry2d&, const Costmap2D&)  costmap.cpp:78-85 (POSE_NONE: :71-76) */
int cover_area_is_free(cover_ctx* ctx, const cover_map* map, const double* polygon_xy, int32_t n_vertices,
                       const cover_poses* poses, const cover_results* out);
/* accumulate_cost(area_generator(res, pose*polygon - origin), map)  impl/costmap.hpp:50-70 over generators.cpp:137-268 */
int cover_area_accumulate_cost(cover_ctx* ctx, const cover_map* map, const double* polygon_xy, int32_t n_vertices,
                               const cover_poses* poses, const cover_results* out);
/* EXTENSION (no reference function): max of detail::get_cell_cost (costmap.hpp:113-124) over the area;
 * a cell outside the map gives -3. */
int cover_area_max_cost(cover_ctx* ctx, const cover_map* map, const double* polygon_xy, int32_t n_vertices,
                        const cover_poses* poses, const cover_results* out);
/* Several outlines = an area with holes: is_free / accumulate_cost / max over
 * area_generator(resolution, const polygon_vec&) (src/cover/generators.cpp:126-135; nested outlines follow the
 * generalised even-odd rule).  Outline k = vertices [starts[k], starts[k+1]) of polygons_xy; every outline is
 * transformed like the single polygon above.  op: 0 = is_free, 1 = accumulate_cost, 2 = max_cost.
 * At most 16 outlines and 256 vertices in total. */
int cover_area_multi_check(cover_ctx* ctx, const cover_map* map, const double* polygons_xy, const int32_t* starts, int32_t n_outlines,
                           int32_t op, const cover_poses* poses, const cover_results* out);
/* is_free(outline_generator(res, pose*polygon - origin), map)  impl/costmap.hpp:23-27 over generators.cpp:54-93 */
int cover_outline_is_free(cover_ctx* ctx, const cover_map* map, const double* polygon_xy, int32_t n_vertices,
                          const cover_poses* poses, const cover_results* out);
int cover_outline_accumulate_cost(cover_ctx* ctx, const cover_map* map, const double* polygon_xy, int32_t n_vertices,
                                  const cover_poses* poses, const cover_results* out);
int cover_outline_max_cost(cover_ctx* ctx, const cover_map* map, const double* polygon_xy, int32_t n_vertices,
                           const cover_poses* poses, const cover_results* out);

/* ---- expand: to_outline / to_area (src/cover/expand.hpp:68-80, impl/expand.hpp:23-39), batched over poses ----
 * The dense cells of outline_generator / area_generator(res, pose*polygon - origin) in GENERATOR ORDER
 * (the per-heading discrete-footprint cache is built from these on the device).
 * starts[count + 1] receives the prefix offsets (pose i owns cells [starts[i], starts[i+1])); with
 * cells == NULL only starts/status are produced (size query), else 2 x starts[count] int32 are written
 * (capacity = number of cells the buffer holds).  A pose whose status is not OK contributes 0 cells.
 * Host arrays only (mem must be COVER_MEM_HOST): this is a per-footprint precompute, not a per-pose check. */
int cover_area_expand(cover_ctx* ctx, const cover_map* map, const double* polygon_xy, int32_t n_vertices, const cover_poses* poses,
                      int32_t* cells, int64_t capacity, int64_t* starts, uint8_t* status, int32_t mem);
int cover_outline_expand(cover_ctx* ctx, const cover_map* map, const double* polygon_xy, int32_t n_vertices, const cover_poses* poses,
                         int32_t* cells, int64_t capacity, int64_t* starts, uint8_t* status, int32_t mem);

/* ---- rays: is_free / accumulate_cost(ray_generator(begin,end)[, offset], map) ---------------------
 * impl/costmap.hpp:23-35,65-81 over generators.cpp:30-52.  segments = count x {x0,y0,x1,y1} cells
 * (end exclusive); offset = 2 ints added to every cell, or NULL. */
int cover_ray_is_free(cover_ctx* ctx, const cover_map* map, const int32_t* segments, int32_t mem, int64_t count,
                      const int32_t* offset, const cover_results* out);
int cover_ray_accumulate_cost(cover_ctx* ctx, const cover_map* map, const int32_t* segments, int32_t mem, int64_t count,
                              const int32_t* offset, const cover_results* out);
int cover_ray_max_cost(cover_ctx* ctx, const cover_map* map, const int32_t* segments, int32_t mem, int64_t count,
                       const int32_t* offset, const cover_results* out);

/* ---- cached cell lists x offsets: is_free(const discrete_polygon&, const cell& offset, map) -------
 * src/cover/costmap.cpp:59-69.  offsets = count x {x,y}. */
int cover_cells_create(cover_ctx* ctx, const int32_t* cells, int64_t n_cells, cover_cells** out);
void cover_cells_destroy(cover_cells* cells);
int cover_cells_is_free(cover_ctx* ctx, const cover_map* map, const cover_cells* cells, const int32_t* offsets, int32_t mem,
                        int64_t count, const cover_results* out);
/* EXTENSIONS: the accumulate_cost / max folds over the same cached list, in list order. */
int cover_cells_accumulate_cost(cover_ctx* ctx, const cover_map* map, const cover_cells* cells, const int32_t* offsets,
                                int32_t mem, int64_t count, const cover_results* out);
int cover_cells_max_cost(cover_ctx* ctx, const cover_map* map, const cover_cells* cells, const int32_t* offsets, int32_t mem,
                         int64_t count, const cover_results* out);

/* ---- split footprints, offset form: generator_footprint / discrete_footprint::is_free(offset, map)
 * src/cover/footprints.cpp:214-247.  A BANK of n_footprints footprints (one per heading bin):
 * footprint k owns outline cells [o_starts[k], o_starts[k+1]) (checked for 253 or 254) and area cells
 * [a_starts[k], a_starts[k+1]) (checked for 254); a cell outside the map collides.
 * `which` (count int32, NULL = all 0) selects the footprint of each offset; an index outside [0, n_footprints)
 * is rejected with COVER_E_INVALID_ARG for host arrays and gives status COVER_STATUS_COORD_RANGE for device arrays. */
int cover_footprints_create(cover_ctx* ctx, const int32_t* outline_cells, const int64_t* o_starts, const int32_t* area_cells,
                            const int64_t* a_starts, int32_t n_footprints, cover_footprints** out);
void cover_footprints_destroy(cover_footprints* fps);
int cover_footprints_is_free(cover_ctx* ctx, const cover_map* map, const cover_footprints* fps, const int32_t* offsets,
                             const int32_t* which, int32_t mem, int64_t count, const cover_results* out);

/* ---- split footprint, pose form: continuous_footprint::is_free(pose, map) -------------------------
 * src/cover/footprints.cpp:130-196 (path B transform: t' = t + (-origin), q = t' + R*p).  The split
 * itself (outlines = always-inscribed sub-regions, areas = remainder) is given as packed polygon lists:
 * polygon k of a list = vertices [starts[k], starts[k+1]) of its xy buffer. */
int cover_split_create(cover_ctx* ctx, const double* outline_xy, const int32_t* outline_starts, int32_t n_outlines,
                       const double* area_xy, const int32_t* area_starts, int32_t n_areas, cover_split** out);
void cover_split_destroy(cover_split* split);
int cover_split_is_free(cover_ctx* ctx, const cover_map* map, const cover_split* split, const cover_poses* poses,
                        const cover_results* out);

/* ---- the split itself, on the HOST and without boost::geometry --------------------------------------
 * get_radius(polygon) (src/cover/footprints.cpp:55-77): min distance from the centroid to the edges; 0 for no points.
 * cover_split_polygon = continuous_footprint(radius, polygon) (footprints.cpp:79-125): the polygon eroded by
 * radius * (1 - 1e-6) -> outlines; the polygon minus the band of width `radius` around the first eroded ring -> areas;
 * polygons with inner rings are dropped (impl/boost_io.hpp:108-124).  Rings come back closed and clockwise, packed like
 * the inputs of cover_split_create: polygon k = vertices [starts[k], starts[k+1]).  The caller provides room for
 * max_vertices vertices per list and max_polygons (+ 1 start) per list; COVER_E_NOMEM if that is too little,
 * COVER_E_INVALID_ARG for radius <= 0 (std::invalid_argument upstream; an EMPTY polygon is a no-op first).
 * The vertex coordinates are NOT boost's (they depend on its buffer strategies and are pinned by no reference test);
 * the construction and its safety property are (cover_b200/csrc/polygon_split.hpp). */
double cover_get_radius(const double* polygon_xy, int32_t n_vertices);
int cover_split_polygon(double radius, const double* polygon_xy, int32_t n_vertices, double* outline_xy, int32_t* outline_starts,
                        int32_t* n_outlines, double* area_xy, int32_t* area_starts, int32_t* n_areas, int32_t max_vertices,
                        int32_t max_polygons);
/* EXTENSION: the same with the outlines eroded by `slack` less and the areas' band `slack` narrower (0 = the reference's
 * construction).  Everything in the polygon stays covered - the areas are the polygon minus the band around the ring
 * that is returned - but ring and areas now overlap by `slack`, which absorbs the cell rounding between "within
 * `radius` of the ring" in metres and "within the inflation layer's inscribed radius of an outline CELL" in cells
 * (the single-cell misses the reference's own regression tolerates, test/footprints.cpp:236-241): one cell of slack
 * meets that test's bar.  The price: an obstacle up to `slack` OUTSIDE the footprint may count as a collision. */
int cover_split_polygon_slack(double radius, double slack, const double* polygon_xy, int32_t n_vertices, double* outline_xy,
                              int32_t* outline_starts, int32_t* n_outlines, double* area_xy, int32_t* area_starts, int32_t* n_areas,
                              int32_t max_vertices, int32_t max_polygons);

/* ---- trajectory scoring (EXTENSION: the DWA-style caller right after the pose checks) -------
 * Poses grouped as n_trajectories x steps (pose index = trajectory * steps + step), e.g. the outputs of the
 * entries above left on the device.  first_blocked[t] = first step with verdict 0 (`steps` if none);
 * score[t] = the first negative per-pose cost in step order if any, else the sum of the per-pose costs
 * (accumulate_cost's rule, impl/costmap.hpp:50-63, lifted from cells to poses).  Either output may be NULL. */
int cover_trajectory_reduce(cover_ctx* ctx, const uint8_t* is_free, const double* cost, int32_t mem, int64_t n_trajectories,
                            int32_t steps, int32_t* first_blocked, double* score);

#ifdef __cplusplus
}
#endif
#endif /* COVER_B200_H */
