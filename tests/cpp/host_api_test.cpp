// host_api_test.cpp — the reference's own costmap / footprint unit tests (test/costmap.cpp:106-194,
// test/footprints.cpp:244-305) restated against the batched C++ host layer (include/cover_b200.hpp),
// i.e. evaluated by the CUDA kernels through the C ABI.  Exit code 0 = all checks passed.
#include "cover_b200.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace cb = cover_b200;

#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                         \
    }                                                                       \
  } while (0)

static cb::polygon make_polygon(std::initializer_list<std::pair<double, double>> pts) {
  cb::polygon p(static_cast<int64_t>(pts.size()));
  int64_t c = 0;
  for (const auto& q : pts) p(0, c) = q.first, p(1, c) = q.second, ++c;
  return p;
}

int main() {
  cb::context ctx(0);

  // costmap_fixture (test/costmap.cpp:53-65): Costmap2D(10, 20, 1, 0, 0) with three special cells
  std::vector<uint8_t> bytes(10 * 20, 0);
  bytes[5 * 10 + 5] = 254, bytes[6 * 10 + 5] = 255, bytes[7 * 10 + 5] = 253;
  cb::costmap map(ctx, 10, 20, 1.0, 0, 0, bytes.data());

  // is_free_generator / is_free_with_offset (test/costmap.cpp:106-136)
  {
    const auto f = cb::ray_is_free({{4, 0, 4, 9}, {5, 0, 5, 9}, {-4, -4, -4, 4}}, map);
    EXPECT(f[0] == 1 && f[1] == 0 && f[2] == 0);
    const int32_t off[2] = {4, 4};
    EXPECT(cb::ray_is_free({{-4, -4, -4, 4}}, map, off)[0] == 1);
  }
  // accumulate_cost_generator / out_of_bounds (test/costmap.cpp:138-194)
  {
    std::vector<uint8_t> tens(10 * 20, 10);
    cb::costmap m10(ctx, 10, 20, 1.0, 0, 0, tens.data());
    const cb::polygon box = make_polygon({{0.5, 0.5}, {9.5, 0.5}, {9.5, 19.5}, {0.5, 19.5}});
    EXPECT(cb::accumulate_cost(box, {cb::pose{1, 0, 0, 0}}, m10)[0] == 10.0 * 200);
    EXPECT(cb::accumulate_cost(box, {cb::pose{1, 0, -4, -4}}, m10)[0] == -3);
    EXPECT(cb::ray_accumulate_cost({{-1, -1, 5, 5}}, m10)[0] == -3);
    EXPECT(cb::is_free(box, m10));
  }
  // non-positive resolution -> std::invalid_argument (base.hpp:54-55)
  {
    bool threw = false;
    try {
      cb::costmap bad(ctx, 10, 10, 0.0, 0, 0, nullptr);
    } catch (const std::invalid_argument&) {
      threw = true;
    }
    EXPECT(threw);
  }
  // open outline -> std::logic_error from the throwing overload, status from the nothrow one (generators.cpp:246-249)
  {
    std::vector<uint8_t> zeros(100 * 100, 0);
    cb::costmap m(ctx, 100, 100, 0.05, 0, 0, zeros.data());
    const cb::polygon open = make_polygon({{0.0, 0.0}, {0.3, 0.45}});
    bool threw = false;
    try {
      cb::is_free(open, {cb::pose{1, 0, 2, 2}}, m);
    } catch (const std::logic_error&) {
      threw = true;
    }
    EXPECT(threw);
    EXPECT(cb::is_free_nothrow(open, {cb::pose{1, 0, 2, 2}}, m).status[0] == COVER_STATUS_LOGIC_ERROR);
  }
  // is_free_fixture / continuous_footprint (test/footprints.cpp:244-305): Costmap2D(100, 200, 0.1, -5, -10)
  {
    std::vector<uint8_t> data(100 * 200, 0);
    const cb::polygon box = make_polygon({{-0.5, -1}, {0.5, -1}, {0.5, 1}, {-0.5, 1}});
    const std::vector<cb::pose> identity{cb::pose{1, 0, 0, 0}};
    auto check = [&](const std::vector<cb::polygon>& outlines, const std::vector<cb::polygon>& areas) {
      cb::costmap m(ctx, 100, 200, 0.1, -5, -10, data.data());
      return cb::continuous_footprint(ctx, outlines, areas).is_free(identity, m)[0] != 0;
    };
    EXPECT(check({box}, {}));
    data[100 * 100 + 50] = 254;  // interior cell: invisible to an outline-only footprint
    EXPECT(check({box}, {}));
    data[90 * 100 + 45] = 253;  // worldToMap(box(0,0), box(1,0)) = (45, 90)
    EXPECT(!check({box}, {}));
    EXPECT(!check({}, {box}));
    data[100 * 100 + 50] = 253;
    EXPECT(check({}, {box}));
    // rotations by 0 and pi stay free, a translation by half the map size collides
    data.assign(data.size(), 0);
    cb::polygon inner = box;
    for (auto& v : inner.v) v *= 0.5;
    cb::costmap m(ctx, 100, 200, 0.1, -5, -10, data.data());
    cb::continuous_footprint fp(ctx, {inner}, {box});
    const auto f = fp.is_free({cb::pose::from_angle(0, 0, 0), cb::pose::from_angle(0, 0, M_PI), cb::pose{1, 0, 5, 10}}, m);
    EXPECT(f[0] == 1 && f[1] == 1 && f[2] == 0);
  }
  // lattice form == explicit poses
  {
    std::vector<uint8_t> data(400 * 400, 0);
    for (size_t i = 0; i < data.size(); i += 977) data[i] = 254;
    cb::costmap m(ctx, 400, 400, 0.05, 0, 0, data.data());
    const cb::polygon rect = make_polygon({{0.5, 0.3}, {0.5, -0.3}, {-0.5, -0.3}, {-0.5, 0.3}});
    cb::pose_lattice lat{1.0, 0.05, 200, 1.0, 0.05, 150, {cb::pose::from_angle(0, 0, 0.3), cb::pose::from_angle(0, 0, -2.0)}};
    const cb::batch_result a = cb::is_free_nothrow(rect, lat, m);
    std::vector<cb::pose> explicit_poses;
    for (size_t t = 0; t < lat.headings.size(); ++t)
      for (int iy = 0; iy < lat.ny; ++iy)
        for (int ix = 0; ix < lat.nx; ++ix) explicit_poses.push_back(cb::pose{lat.headings[t].c, lat.headings[t].s, lat.x0 + ix * lat.dx, lat.y0 + iy * lat.dy});
    const std::vector<uint8_t> b = cb::is_free(rect, explicit_poses, m);
    EXPECT(a.is_free == b);
    size_t free_count = 0;
    for (uint8_t v : b) free_count += v;
    EXPECT(free_count > 0 && free_count < b.size());
    // the other lattice forms against explicit poses: bit-packed verdicts, accumulate_cost, outline checks
    {
      int64_t not_ok = -1;
      const std::vector<uint32_t> bits = cb::is_free_bits(rect, lat, m, &not_ok);
      EXPECT(not_ok == 0);
      for (size_t i = 0; i < b.size(); ++i) EXPECT(((bits[i >> 5] >> (i & 31)) & 1u) == b[i]);
      EXPECT(cb::accumulate_cost_nothrow(rect, lat, m).cost == cb::accumulate_cost(rect, explicit_poses, m));
      EXPECT(cb::outline_is_free_nothrow(rect, lat, m).is_free == cb::outline_is_free(rect, explicit_poses, m));
      EXPECT(cb::outline_accumulate_cost_nothrow(rect, lat, m).cost == cb::outline_accumulate_cost(rect, explicit_poses, m));
    }
    // a new image queued with upload_async is what the next check sees (the library orders the reader after the copy)
    std::vector<uint8_t> cleared(data.size(), 0);
    m.upload_async(cleared.data());
    const cb::batch_result c = cb::is_free_nothrow(rect, lat, m);
    for (uint8_t v : c.is_free) EXPECT(v == 1);
    m.upload_async(data.data());
    EXPECT(cb::is_free_nothrow(rect, lat, m).is_free == b);
  }
  std::puts("host_api_test: all checks passed");
  return 0;
}
