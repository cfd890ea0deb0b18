// g2o/stuff/misc.h stand-in: the three helpers the reference calls (upstream g2o, third-party).
#ifndef REF_SHIM_G2O_MISC
#define REF_SHIM_G2O_MISC
#include <cmath>
namespace g2o {
inline double normalize_theta(double theta) {
  if (theta >= -M_PI && theta < M_PI) return theta;
  double multiplier = std::floor(theta / (2 * M_PI));
  theta = theta - multiplier * 2 * M_PI;
  if (theta >= M_PI) theta -= 2 * M_PI;
  if (theta < -M_PI) theta += 2 * M_PI;
  return theta;
}
inline double average_angle(double theta1, double theta2) {
  double x = std::cos(theta1) + std::cos(theta2), y = std::sin(theta1) + std::sin(theta2);
  if (x == 0 && y == 0) return 0;
  return std::atan2(y, x);
}
template <typename T> inline int sign(T x) { return x > 0 ? 1 : (x < 0 ? -1 : 0); }
}  // namespace g2o
#endif
