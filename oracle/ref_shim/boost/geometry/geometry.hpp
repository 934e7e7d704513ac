// Stub of the boost::geometry names src/cover/footprints.cpp and impl/boost_io.hpp mention, so that
// footprints.cpp compiles UNMODIFIED.  The split (buffer / difference / centroid / distance) is NOT
// available here: those stubs throw.  Only continuous_footprint::is_free, generator_footprint and
// discrete_footprint (which never touch boost) are exercised through oracle/_ref.
#pragma once
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace boost {
namespace geometry {

inline void shim_unavailable() { throw std::runtime_error("boost::geometry is not available in the oracle shim"); }

namespace model {
namespace d2 {
template <typename T>
class point_xy {
public:
  point_xy() : x_(), y_() {}
  point_xy(T x, T y) : x_(x), y_(y) {}
  template <int I>
  T get() const {
    return I == 0 ? x_ : y_;
  }
  T x() const { return x_; }
  T y() const { return y_; }

private:
  T x_, y_;
};
}  // namespace d2

template <typename Point>
class polygon {
public:
  typedef std::vector<Point> ring_type;
  ring_type& outer() { return outer_; }
  const ring_type& outer() const { return outer_; }
  std::vector<ring_type>& inners() { return inners_; }
  const std::vector<ring_type>& inners() const { return inners_; }

private:
  ring_type outer_;
  std::vector<ring_type> inners_;
};

template <typename Polygon>
class multi_polygon : public std::vector<Polygon> {};

template <typename Point>
class linestring : public std::vector<Point> {
public:
  template <typename It>
  linestring(It b, It e) : std::vector<Point>(b, e) {}
};

template <typename Point>
class referring_segment {
public:
  referring_segment(Point&, Point&) {}
};
}  // namespace model

namespace strategy {
namespace buffer {
template <typename T>
struct distance_symmetric {
  explicit distance_symmetric(T) {}
};
struct side_straight {};
struct join_round {
  explicit join_round(std::size_t) {}
};
struct end_round {
  explicit end_round(std::size_t) {}
};
struct point_circle {
  explicit point_circle(std::size_t) {}
};
}  // namespace buffer
}  // namespace strategy

template <typename G>
inline void correct(G&) {}
template <typename G>
inline bool is_empty(const G& g) {
  return g.outer().empty();
}
template <typename G, typename P>
inline void centroid(const G&, P&) {
  shim_unavailable();
}
template <typename A, typename B>
inline double distance(const A&, const B&) {
  shim_unavailable();
  return 0;
}
template <typename G, typename Out, typename D, typename S, typename J, typename E, typename C>
inline void buffer(const G&, Out&, const D&, const S&, const J&, const E&, const C&) {
  shim_unavailable();
}
template <typename A, typename B, typename Out>
inline void difference(const A&, const B&, Out&) {
  shim_unavailable();
}

}  // namespace geometry
}  // namespace boost
