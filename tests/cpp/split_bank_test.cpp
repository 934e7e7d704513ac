// split_bank_test.cpp — the C++ host layer's grid split (include/cover_b200.hpp: to_area, split_cells,
// heading_bank), areas with holes, the inflation extension and trajectory scoring, through the C ABI.
//
//   split_bank_test <input> <output>
//
// <input>  (text): resolution radius / n_vertices / x y per vertex / n_headings / cos sin per heading
// <output> (text): per heading "H k n_outline n_area" followed by the outline cells and the area cells ("x y"),
//          which tests/test_gpu_parity.py compares cell for cell with cover_b200/split.py.
// On top of that the program runs the reference's safety regression (test/footprints.cpp:205-242) on the bank it
// built: a lethal cell anywhere on the footprint, inflated with an inscribed radius equal to the split radius, must
// be found by discrete_footprint::is_free; an empty map must be free.  Exit code 0 = all checks passed.
#include "cover_b200.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>

namespace cb = cover_b200;

#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                         \
    }                                                                       \
  } while (0)

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  std::ifstream in(argv[1]);
  double resolution, radius;
  int nv, nh;
  in >> resolution >> radius >> nv;
  cb::polygon poly(nv);
  for (int v = 0; v < nv; ++v) in >> poly(0, v) >> poly(1, v);
  in >> nh;
  std::vector<cb::pose> headings(nh);
  for (cb::pose& h : headings) in >> h.c >> h.s, h.x = h.y = 0;
  EXPECT(in.good());

  cb::context ctx(0);
  const int size = 240, centre = size / 2;
  std::vector<uint8_t> empty(static_cast<size_t>(size) * size, 0);
  cb::costmap map(ctx, size, size, resolution, -centre * resolution, -centre * resolution, empty.data());

  // ---- to_area per heading + split, written out for the cross-check against split.py
  std::vector<cb::pose> poses;
  for (const cb::pose& h : headings) poses.push_back(cb::pose{h.c, h.s, map.origin_x(), map.origin_y()});
  const std::vector<cb::discrete_polygon> footprints = cb::to_area(poly, poses, map);
  std::vector<cb::split_cells_result> splits;
  std::ofstream out(argv[2]);
  for (int k = 0; k < nh; ++k) {
    splits.push_back(cb::split_cells(footprints[k], resolution, radius));
    const cb::split_cells_result& s = splits.back();
    out << "H " << k << ' ' << s.outline.cols() << ' ' << s.area.cols() << '\n';
    for (int64_t i = 0; i < s.outline.cols(); ++i) out << s.outline(0, i) << ' ' << s.outline(1, i) << '\n';
    for (int64_t i = 0; i < s.area.cols(); ++i) out << s.area(0, i) << ' ' << s.area(1, i) << '\n';
    EXPECT(s.outline.cols() + s.area.cols() < footprints[k].cols());  // fewer cells than the full-area scan
  }
  out.close();
  // to_outline: the first cell of the outline is the cell of vertex 0
  {
    const std::vector<cb::discrete_polygon> outlines = cb::to_outline(poly, {poses[0]}, map);
    EXPECT(outlines[0].cols() > 0);
    const double qx = (poses[0].x + (poses[0].c * poly(0, 0) + -poses[0].s * poly(1, 0))) - map.origin_x();
    EXPECT(outlines[0](0, 0) == static_cast<int>(qx * (1. / resolution)));
  }
  // radius validation and the too-large radius (everything becomes area)
  {
    bool threw = false;
    try {
      cb::split_cells(footprints[0], resolution, 0.0);
    } catch (const std::invalid_argument&) {
      threw = true;
    }
    EXPECT(threw);
    const cb::split_cells_result all = cb::split_cells(footprints[0], resolution, 100.0);
    EXPECT(all.outline.cols() == 0 && all.area.cols() > 0);
  }

  // ---- the safety regression on the bank
  const cb::discrete_footprints bank = cb::heading_bank(map, poly, radius, headings);
  const int rad = static_cast<int>(std::ceil(radius / resolution));
  std::vector<uint8_t> lut(static_cast<size_t>(rad) * rad + 1, 0);  // inscribed radius = split radius, nothing beyond
  for (int d2 = 0; d2 <= rad * rad; ++d2) lut[d2] = std::sqrt(static_cast<double>(d2)) * resolution <= radius ? 253 : 0;
  lut[0] = 254;
  const std::vector<cb::cell_offset> at_centre{cb::cell_offset{centre, centre}};
  int checked = 0;
  for (int k = 0; k < nh; k += 3) {
    const std::vector<int32_t> which{k};
    EXPECT(bank.is_free(at_centre, which, map)[0] == 1);
    for (int64_t i = 0; i < footprints[k].cols(); i += 11) {  // every 11th footprint cell
      std::vector<uint8_t> data = empty;
      data[static_cast<size_t>(footprints[k](1, i) + centre) * size + (footprints[k](0, i) + centre)] = 254;
      map.upload(data.data());
      map.inflate(rad, lut);
      EXPECT(bank.is_free(at_centre, which, map)[0] == 0);
      ++checked;
    }
    map.upload(empty.data());
  }
  EXPECT(checked > 20);
  // the inflation extension itself: one obstacle becomes a disc of 253 around a 254
  {
    std::vector<uint8_t> data = empty;
    data[static_cast<size_t>(centre) * size + centre] = 254;
    map.upload(data.data());
    map.inflate(rad, lut);
    const std::vector<uint8_t> got = map.download();
    for (int y = centre - rad - 1; y <= centre + rad + 1; ++y)
      for (int x = centre - rad - 1; x <= centre + rad + 1; ++x) {
        const int d2 = (x - centre) * (x - centre) + (y - centre) * (y - centre);
        const uint8_t want = d2 == 0 ? 254 : (d2 <= rad * rad ? lut[d2] : 0);
        EXPECT(got[static_cast<size_t>(y) * size + x] == want);
      }
    map.upload(empty.data());
  }

  // ---- an area with a hole (area_generator(resolution, polygon_vec), generators.cpp:126-135)
  {
    cb::polygon outer(4), hole(4);
    const double o = 2.0, h = 1.0;
    const double ox[4] = {-o, o, o, -o}, oy[4] = {-o, -o, o, o};
    for (int v = 0; v < 4; ++v) outer(0, v) = ox[v], outer(1, v) = oy[v], hole(0, v) = ox[v] * h / o, hole(1, v) = oy[v] * h / o;
    const std::vector<cb::pose> identity{cb::pose{1, 0, 0, 0}};
    std::vector<uint8_t> data = empty;
    data[static_cast<size_t>(centre) * size + centre] = 254;  // inside the hole
    map.upload(data.data());
    EXPECT(cb::is_free({outer, hole}, identity, map)[0] == 1);
    EXPECT(cb::is_free(outer, identity, map)[0] == 0);
    data[static_cast<size_t>(centre) * size + centre + static_cast<int>(1.5 / resolution)] = 254;  // inside the ring
    map.upload(data.data());
    EXPECT(cb::is_free({outer, hole}, identity, map)[0] == 0);
    std::vector<uint8_t> sevens(empty.size(), 7);
    map.upload(sevens.data());
    const double ring = cb::accumulate_cost({outer, hole}, identity, map)[0], full = cb::accumulate_cost(outer, identity, map)[0];
    EXPECT(ring > 0 && ring < full && std::fmod(ring, 7.0) == 0);
    map.upload(empty.data());
  }

  // ---- trajectory scoring: first blocked step, accumulate-style score
  {
    const std::vector<uint8_t> free_{1, 1, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0};
    const std::vector<double> cost{1, 2, -1, 4, 5, 6, 7, 8, -3, -2, 0, 0};
    const cb::trajectory_scores t = cb::trajectory_reduce(ctx, free_, cost, 4);
    EXPECT(t.first_blocked.size() == 3 && t.first_blocked[0] == 2 && t.first_blocked[1] == 4 && t.first_blocked[2] == 0);
    EXPECT(t.score[0] == -1 && t.score[1] == 26 && t.score[2] == -3);
  }
  std::puts("split_bank_test: all checks passed");
  return 0;
}
