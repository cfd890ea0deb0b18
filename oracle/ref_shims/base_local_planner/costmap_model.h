#ifndef REF_SHIM_COSTMAP_MODEL
#define REF_SHIM_COSTMAP_MODEL
#include <geometry_msgs/Point.h>
#include <vector>
namespace base_local_planner {
class CostmapModel {
 public:
  virtual ~CostmapModel() {}
  virtual double footprintCost(double, double, double, const std::vector<geometry_msgs::Point>&, double = 0.0, double = 0.0) { return 0.0; }
};
}
#endif
