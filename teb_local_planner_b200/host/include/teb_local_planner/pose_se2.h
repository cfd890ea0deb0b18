/*
 * pose_se2.h — ROS/Eigen-free mirror of the reference's basic types for the drop-in host layer.
 *
 * Mirrors (same names, argument meaning and arithmetic):
 *   teb_local_planner::PoseSE2            include/teb_local_planner/pose_se2.h:60-400 (plus :238, average :266,
 *                                         orientationUnitVec :166)
 *   g2o::normalize_theta / average_angle  
 *   geometry_msgs::Twist / PoseStamped, tf::Pose: field-compatible PODs (ROS is not available in this image;
 *   the ROS-typed constructors of the reference are out of scope)
 */
#ifndef TEB_B200_POSE_SE2_H_
#define TEB_B200_POSE_SE2_H_

#include <cmath>
#include <vector>

namespace Eigen {
/* minimal stand-in for Eigen::Vector2d: only what the reference's public signatures on this path use */
struct Vector2d {
  double v[2];
  Vector2d() : v{0, 0} {}
  Vector2d(double x, double y) : v{x, y} {}
  double& x() { return v[0]; }
  double& y() { return v[1]; }
  const double& x() const { return v[0]; }
  const double& y() const { return v[1]; }
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
  double& coeffRef(int i) { return v[i]; }
  Vector2d operator+(const Vector2d& o) const { return Vector2d(v[0] + o.v[0], v[1] + o.v[1]); }
  Vector2d operator-(const Vector2d& o) const { return Vector2d(v[0] - o.v[0], v[1] - o.v[1]); }
  Vector2d operator*(double s) const { return Vector2d(v[0] * s, v[1] * s); }
  Vector2d operator/(double s) const { return Vector2d(v[0] / s, v[1] / s); }
  double dot(const Vector2d& o) const { return v[0] * o.v[0] + v[1] * o.v[1]; }
  double squaredNorm() const { return v[0] * v[0] + v[1] * v[1]; }
  double norm() const { return std::sqrt(squaredNorm()); }
  /* Eigen: divides by sqrt(squaredNorm()) if that is positive */
  void normalize() { const double z = squaredNorm(); if (z > 0) { const double r = std::sqrt(z); v[0] /= r; v[1] /= r; } }
  Vector2d normalized() const { Vector2d o(*this); o.normalize(); return o; }
  static Vector2d Zero() { return Vector2d(0, 0); }
};
inline Vector2d operator*(double s, const Vector2d& a) { return a * s; }
}  // namespace Eigen

namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Twist { Vector3 linear, angular; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { Pose pose; };
}  // namespace geometry_msgs

/* the one costmap query the planners make (isTrajectoryFeasible): base_local_planner::CostmapModel::footprintCost
 * returns -1 for a footprint in collision / off the map. The caller plugs its own map in by deriving from this. */
namespace base_local_planner {
class CostmapModel {
 public:
  virtual ~CostmapModel() {}
  virtual double footprintCost(double x, double y, double theta, const std::vector<geometry_msgs::Point>& footprint_spec,
                               double inscribed_radius = 0.0, double circumscribed_radius = 0.0) = 0;
};
}  // namespace base_local_planner

namespace tf {
inline double getYaw(const geometry_msgs::Quaternion& q) {
  return std::atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
}
inline geometry_msgs::Quaternion createQuaternionFromYaw(double yaw) {
  geometry_msgs::Quaternion q;
  q.z = std::sin(yaw / 2);
  q.w = std::cos(yaw / 2);
  return q;
}
typedef geometry_msgs::Pose Pose;
}  // namespace tf

namespace g2o {
inline double normalize_theta(double theta) {
  if (theta >= -M_PI && theta < M_PI) return theta;
  double multiplier = std::floor(theta / (2 * M_PI));
  theta = theta - multiplier * 2 * M_PI;
  if (theta >= M_PI) theta -= 2 * M_PI;
  if (theta < -M_PI) theta += 2 * M_PI;
  return theta;
}
inline double average_angle(double a, double b) {
  double x = std::cos(a) + std::cos(b), y = std::sin(a) + std::sin(b);
  if (x == 0 && y == 0) return 0;
  return std::atan2(y, x);
}
template <typename T> inline int sign(T x) { return x > 0 ? 1 : (x < 0 ? -1 : 0); }
}  // namespace g2o

namespace teb_local_planner {

enum class RotType { left, none, right };  /* misc.h:54 */

class PoseSE2 {
 public:
  PoseSE2() : _position(0, 0), _theta(0) {}
  PoseSE2(const Eigen::Vector2d& position, double theta) : _position(position), _theta(theta) {}
  PoseSE2(double x, double y, double theta) : _position(x, y), _theta(theta) {}
  explicit PoseSE2(const geometry_msgs::Pose& pose)
      : _position(pose.position.x, pose.position.y), _theta(tf::getYaw(pose.orientation)) {}
  Eigen::Vector2d& position() { return _position; }
  const Eigen::Vector2d& position() const { return _position; }
  double& x() { return _position.x(); }
  const double& x() const { return _position.x(); }
  double& y() { return _position.y(); }
  const double& y() const { return _position.y(); }
  double& theta() { return _theta; }
  const double& theta() const { return _theta; }
  void setZero() { _position = Eigen::Vector2d(0, 0); _theta = 0; }
  Eigen::Vector2d orientationUnitVec() const { return Eigen::Vector2d(std::cos(_theta), std::sin(_theta)); }
  void scale(double factor) { _position = _position * factor; _theta = g2o::normalize_theta(_theta * factor); }
  /* pose_se2.h:238 */
  void plus(const double* pose_as_array) {
    _position.coeffRef(0) += pose_as_array[0];
    _position.coeffRef(1) += pose_as_array[1];
    _theta = g2o::normalize_theta(_theta + pose_as_array[2]);
  }
  void averageInPlace(const PoseSE2& pose1, const PoseSE2& pose2) {
    _position = (pose1._position + pose2._position) / 2;
    _theta = g2o::average_angle(pose1._theta, pose2._theta);
  }
  /* pose_se2.h:266 */
  static PoseSE2 average(const PoseSE2& pose1, const PoseSE2& pose2) {
    return PoseSE2((pose1._position + pose2._position) / 2, g2o::average_angle(pose1._theta, pose2._theta));
  }

 private:
  Eigen::Vector2d _position;
  double _theta;
};

}  // namespace teb_local_planner
#endif
