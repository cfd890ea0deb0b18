#ifndef REF_SHIM_ROS_ROS
#define REF_SHIM_ROS_ROS
#include <ros/console.h>
#include <boost/shared_ptr.hpp>
#include <boost/thread/mutex.hpp>
#include <boost/thread/once.hpp>
#include <boost/optional.hpp>
#include <boost/make_shared.hpp>
#include <chrono>
#include <string>
namespace ros {
class NodeHandle {};
struct Duration { double s; explicit Duration(double v = 0) : s(v) {} double toSec() const { return s; } Duration& fromSec(double v) { s = v; return *this; } };
inline bool ok() { return true; }
/* now(): a steady clock, so that the planner's "time since the last class switch" (homotopy_class_planner.cpp:634-649)
 * is positive as it is under ROS; a constant would block every switch */
struct Time { double s; Time() : s(0) {} static Time now() { Time t; t.s = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); return t; } double toSec() const { return s; } Duration operator-(const Time& o) const { return Duration(s - o.s); } };
}
#endif
