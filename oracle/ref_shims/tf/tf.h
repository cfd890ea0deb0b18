#ifndef REF_SHIM_TF
#define REF_SHIM_TF
#include <cmath>
#include <geometry_msgs/Point.h>
namespace tf {
struct Quaternion { double x_, y_, z_, w_; Quaternion(double x = 0, double y = 0, double z = 0, double w = 1) : x_(x), y_(y), z_(z), w_(w) {} };
struct Vector3 { double x_, y_, z_; Vector3(double x = 0, double y = 0, double z = 0) : x_(x), y_(y), z_(z) {} double x() const { return x_; } double y() const { return y_; } double z() const { return z_; } double getX() const { return x_; } double getY() const { return y_; } };
struct Pose { Vector3 o; Quaternion q; const Vector3& getOrigin() const { return o; } Quaternion getRotation() const { return q; } };
typedef Pose Transform;
inline double yaw_of(double x, double y, double z, double w) { return std::atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z)); }
inline double getYaw(const Quaternion& q) { return yaw_of(q.x_, q.y_, q.z_, q.w_); }
inline double getYaw(const geometry_msgs::Quaternion& q) { return yaw_of(q.x, q.y, q.z, q.w); }
inline geometry_msgs::Quaternion createQuaternionMsgFromYaw(double yaw) { geometry_msgs::Quaternion q; q.x = 0; q.y = 0; q.z = std::sin(yaw / 2); q.w = std::cos(yaw / 2); return q; }
inline Quaternion createQuaternionFromYaw(double yaw) { return Quaternion(0, 0, std::sin(yaw / 2), std::cos(yaw / 2)); }
}
#endif
