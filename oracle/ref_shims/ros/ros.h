#ifndef REF_SHIM_ROS_ROS
#define REF_SHIM_ROS_ROS
#include <ros/console.h>
#include <boost/shared_ptr.hpp>
#include <boost/thread/mutex.hpp>
#include <boost/thread/once.hpp>
#include <boost/optional.hpp>
#include <boost/make_shared.hpp>
#include <string>
namespace ros {
class NodeHandle {};
struct Duration { double s; explicit Duration(double v = 0) : s(v) {} double toSec() const { return s; } Duration& fromSec(double v) { s = v; return *this; } };
inline bool ok() { return true; }
struct Time { double s; Time() : s(0) {} static Time now() { return Time(); } double toSec() const { return s; } Duration operator-(const Time& o) const { return Duration(s - o.s); } };
}
#endif
