// Stand-in for costmap_2d::Costmap2D: the accessors magazino/cover calls, over a borrowed or owned
// row-major byte array (index = my * size_x + mx, as in ros-navigation).
#pragma once
#include <vector>
namespace costmap_2d {
class Costmap2D {
public:
  Costmap2D() : sx_(0), sy_(0), res_(1), ox_(0), oy_(0), ext_(nullptr) {}
  Costmap2D(unsigned int sx, unsigned int sy, double res, double ox, double oy, unsigned char def = 0) :
      sx_(sx), sy_(sy), res_(res), ox_(ox), oy_(oy), own_(static_cast<size_t>(sx) * sy, def), ext_(nullptr) {}
  /// Borrowing constructor (shim only): wraps caller memory without a copy.
  Costmap2D(const unsigned char* data, unsigned int sx, unsigned int sy, double res, double ox, double oy) :
      sx_(sx), sy_(sy), res_(res), ox_(ox), oy_(oy), ext_(data) {}
  unsigned char getCost(unsigned int mx, unsigned int my) const { return bytes()[static_cast<size_t>(my) * sx_ + mx]; }
  void setCost(unsigned int mx, unsigned int my, unsigned char c) { own_[static_cast<size_t>(my) * sx_ + mx] = c; }
  unsigned int getSizeInCellsX() const { return sx_; }
  unsigned int getSizeInCellsY() const { return sy_; }
  double getOriginX() const { return ox_; }
  double getOriginY() const { return oy_; }
  double getResolution() const { return res_; }
  unsigned char* getCharMap() { return own_.data(); }

private:
  const unsigned char* bytes() const { return ext_ ? ext_ : own_.data(); }
  unsigned int sx_, sy_;
  double res_, ox_, oy_;
  std::vector<unsigned char> own_;
  const unsigned char* ext_;
};
}  // namespace costmap_2d
