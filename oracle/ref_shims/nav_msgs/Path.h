#ifndef REF_SHIM_NAV_PATH
#define REF_SHIM_NAV_PATH
#include <geometry_msgs/Point.h>
namespace nav_msgs { struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; }; struct Odometry {}; }
#endif
