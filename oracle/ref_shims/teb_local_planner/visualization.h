// Shadows include/teb_local_planner/visualization.h (ROS publishers + boost::graph, irrelevant to the hot path): the
// planner only stores the pointer and calls these when it is non-null; oracle/_ref always passes a null pointer.
#ifndef REF_SHIM_VISUALIZATION
#define REF_SHIM_VISUALIZATION
#include <teb_local_planner/teb_config.h>
#include <teb_local_planner/timed_elastic_band.h>
#include <teb_local_planner/robot_footprint_model.h>
#include <boost/shared_ptr.hpp>
namespace teb_local_planner {
class TebOptimalPlanner;
class TebVisualization {
 public:
  void publishLocalPlanAndPoses(const TimedElasticBand&) const {}
  void publishRobotFootprintModel(const PoseSE2&, const BaseRobotFootprintModel&, const std::string& = "", const std_msgs::ColorRGBA& = std_msgs::ColorRGBA()) {}
  void publishInfeasibleRobotPose(const PoseSE2&, const BaseRobotFootprintModel&, const std::vector<geometry_msgs::Point>&) {}
  void publishFeedbackMessage(const TebOptimalPlanner&, const ObstContainer&) {}
  template <class Container> void publishFeedbackMessage(const Container&, unsigned int, const ObstContainer&) {}
  template <class Graph> void publishGraph(const Graph&, const std::string& = "") {}
  template <class Container> void publishTebContainer(const Container&, const std::string& = "") {}
};
typedef boost::shared_ptr<TebVisualization> TebVisualizationPtr;
typedef boost::shared_ptr<const TebVisualization> TebVisualizationConstPtr;
}
#endif
