#ifndef REF_SHIM_TRAJ_MSG
#define REF_SHIM_TRAJ_MSG
#include <geometry_msgs/Point.h>
#include <ros/ros.h>
namespace teb_local_planner {
struct TrajectoryPointMsg { geometry_msgs::Pose pose; geometry_msgs::Twist velocity, acceleration; ros::Duration time_from_start; };
struct TrajectoryMsg { std_msgs::Header header; std::vector<TrajectoryPointMsg> trajectory; };
}
#endif
