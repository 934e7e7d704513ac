// multi_gpu_test.cpp — cover_b200::multi_gpu (include/cover_b200.hpp): pose sharding across contexts, results gathered
// into one host buffer, against a single context.  With one GPU the shards are three contexts of device 0.
#include "cover_b200.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace cover_b200;

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);        \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main() {
  const int n_dev = cover_device_count();
  CHECK(n_dev >= 1);
  std::vector<int> devices;
  if (n_dev >= 2) {
    for (int d = 0; d < n_dev; ++d) devices.push_back(d);
  } else {
    devices = {0, 0, 0};
  }
  const int sx = 900, sy = 700;
  std::vector<uint8_t> image(static_cast<size_t>(sx) * sy, 0);
  std::mt19937 rng(7);
  for (uint8_t& c : image) {
    const unsigned v = rng();
    c = (v % 500 == 0) ? 254 : static_cast<uint8_t>(v % 200);
  }
  polygon rect(4);
  const double xs[4] = {0.5, 0.5, -0.5, -0.5}, ys[4] = {0.3, -0.3, -0.3, 0.3};
  for (int v = 0; v < 4; ++v) rect(0, v) = xs[v], rect(1, v) = ys[v];

  multi_gpu gpus(devices);
  gpus.set_costmap(sx, sy, 0.05, 0.0, 0.0, image.data());
  context one(0);
  costmap map(one, sx, sy, 0.05, 0.0, 0.0, image.data());

  // block partition: contiguous, disjoint, covering; granule borders
  int64_t at = 0;
  for (int k = 0; k < gpus.size(); ++k) {
    const auto b = gpus.block(1000003, k, 32);
    CHECK(b.first == at && (b.first % 32 == 0));
    at += b.second;
  }
  CHECK(at == 1000003);

  // explicit poses
  std::vector<pose> poses;
  std::uniform_real_distribution<double> ux(1.0, 44.0), uy(1.0, 34.0), ut(-3.14159, 3.14159);
  for (int i = 0; i < 100001; ++i) poses.push_back(pose::from_angle(ux(rng), uy(rng), ut(rng)));
  const batch_result a = gpus.is_free_nothrow(rect, poses);
  const batch_result b = is_free_nothrow(rect, poses, map);
  CHECK(a.is_free == b.is_free && a.status == b.status);

  // lattice, bytes and bits
  pose_lattice lat{2.0, 0.05, 333, 3.0, 0.05, 201, {}};
  for (int t = 0; t < 5; ++t) lat.headings.push_back(pose::from_angle(0, 0, -3.0 + 1.2 * t));
  const batch_result la = gpus.is_free_nothrow(rect, lat);
  const batch_result lb = is_free_nothrow(rect, lat, map);
  CHECK(la.is_free == lb.is_free && la.status == lb.status);
  int64_t bad = -1;
  const std::vector<uint32_t> bits = gpus.is_free_bits(rect, lat, &bad);
  CHECK(bad == 0);
  for (int64_t i = 0; i < lat.count(); ++i) CHECK(((bits[static_cast<size_t>(i >> 5)] >> (i & 31)) & 1u) == lb.is_free[static_cast<size_t>(i)]);
  int64_t n_free = 0;
  for (uint8_t f : lb.is_free) n_free += f;
  CHECK(n_free > 0 && n_free < lat.count());
  std::printf("multi_gpu ok: %d shards on %d device(s), %lld of %lld lattice poses free\n", gpus.size(), n_dev, static_cast<long long>(n_free),
              static_cast<long long>(lat.count()));
  return 0;
}
