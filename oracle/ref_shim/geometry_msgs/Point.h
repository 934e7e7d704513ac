#pragma once
namespace geometry_msgs {
struct Point {
  double x, y, z;
};
}  // namespace geometry_msgs
