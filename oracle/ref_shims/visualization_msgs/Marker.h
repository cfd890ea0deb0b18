#ifndef REF_SHIM_MARKER
#define REF_SHIM_MARKER
#include <geometry_msgs/Point.h>
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7, POINTS = 8 };
  int type; geometry_msgs::Pose pose; geometry_msgs::Vector3 scale; std_msgs::ColorRGBA color; std::vector<geometry_msgs::Point> points;
  Marker() : type(0) {}
};
}
#endif
