/*
 * optimal_planner.h — drop-in TebOptimalPlanner whose optimizeTEB runs on the B200 through the C-ABI.
 *
 * Mirrors teb_local_planner::PlannerInterface (include/teb_local_planner/planner_interface.h:67-200) and
 * teb_local_planner::TebOptimalPlanner (include/teb_local_planner/optimal_planner.h:100-700,
 * src/optimal_planner.cpp): same class / method names, argument meaning and bool-return error behaviour.
 * Replaced: buildGraph/optimizeGraph/clearGraph and the g2o optimizer (optimal_planner.cpp:161-179, 323-418) —
 * the whole optimizeTEB loop is one tebgpu_optimize_batch call. Not available (documented in INTEGRATION.md):
 * optimizer() accessors (no g2o object exists) and visualize(). isTrajectoryFeasible() takes the costmap through the
 * abstract base_local_planner::CostmapModel declared in pose_se2.h (costmap_2d itself is not part of this repository).
 */
#ifndef TEB_B200_OPTIMAL_PLANNER_H_
#define TEB_B200_OPTIMAL_PLANNER_H_

#include <memory>
#include <mutex>
#include <vector>

#include "teb_b200.h"
#include "teb_local_planner/obstacles.h"
#include "teb_local_planner/teb_config.h"
#include "teb_local_planner/timed_elastic_band.h"

namespace teb_local_planner {

/* placeholder for TebVisualizationPtr so that the reference's constructor signatures survive (rviz is out of scope) */
struct TebVisualization {};
typedef std::shared_ptr<TebVisualization> TebVisualizationPtr;

/* One tebgpu context (device buffers + stream), shared by every planner of a process unless told otherwise. */
class TebGpuContext {
 public:
  TebGpuContext(int max_bands, int max_poses, int max_obstacles, int max_viapoints, int device = 0,
                int max_obst_vertices = 2048);
  ~TebGpuContext();
  tebgpu_ctx* get() const { return ctx_; }
  const TebGpuLimits& limits() const { return lim_; }
  std::mutex& mutex() { return mutex_; } /* serialises set_params .. optimize sequences of planners sharing the context */
  static std::shared_ptr<TebGpuContext> shared(int min_bands = 8, int min_poses = 512, int min_obstacles = 256,
                                               int min_viapoints = 16, int min_obst_vertices = 2048);
 private:
  tebgpu_ctx* ctx_ = nullptr;
  TebGpuLimits lim_{};
  std::mutex mutex_;
};
typedef std::shared_ptr<TebGpuContext> TebGpuContextPtr;

class PlannerInterface {
 public:
  PlannerInterface() {}
  virtual ~PlannerInterface() {}
  virtual bool plan(const std::vector<geometry_msgs::PoseStamped>& initial_plan, const geometry_msgs::Twist* start_vel = NULL, bool free_goal_vel = false) = 0;
  virtual bool plan(const tf::Pose& start, const tf::Pose& goal, const geometry_msgs::Twist* start_vel = NULL, bool free_goal_vel = false) = 0;
  virtual bool plan(const PoseSE2& start, const PoseSE2& goal, const geometry_msgs::Twist* start_vel = NULL, bool free_goal_vel = false) = 0;
  virtual bool getVelocityCommand(double& vx, double& vy, double& omega, int look_ahead_poses) const = 0;
  virtual void clearPlanner() = 0;
  virtual void setPreferredTurningDir(RotType dir) { (void)dir; }
  virtual void visualize() {}
  virtual void updateRobotModel(RobotFootprintModelPtr robot_model) { (void)robot_model; }
  virtual void computeCurrentCost(std::vector<double>& cost, double obst_cost_scale = 1.0, bool alternative_time_cost = false) {
    (void)cost; (void)obst_cost_scale; (void)alternative_time_cost;
  }
  virtual bool hasDiverged() const = 0;
  /* planner_interface.h:181 */
  virtual bool isTrajectoryFeasible(base_local_planner::CostmapModel* costmap_model, const std::vector<geometry_msgs::Point>& footprint_spec,
                                    double inscribed_radius = 0.0, double circumscribed_radius = 0.0, int look_ahead_idx = -1,
                                    double feasibility_check_lookahead_distance = -1.0) = 0;
};
typedef std::shared_ptr<PlannerInterface> PlannerInterfacePtr;

/* teb_local_planner/TrajectoryPointMsg (msg/TrajectoryPointMsg.msg) without ROS */
struct TrajectoryPointMsg {
  geometry_msgs::Pose pose;
  geometry_msgs::Twist velocity;
  geometry_msgs::Twist acceleration;
  double time_from_start = 0; /* seconds */
};

class TebOptimalPlanner : public PlannerInterface {
 public:
  TebOptimalPlanner();
  TebOptimalPlanner(const TebConfig& cfg, ObstContainer* obstacles = NULL, TebVisualizationPtr visual = TebVisualizationPtr(),
                    const ViaPointContainer* via_points = NULL);
  virtual ~TebOptimalPlanner();
  void initialize(const TebConfig& cfg, ObstContainer* obstacles = NULL, TebVisualizationPtr visual = TebVisualizationPtr(),
                  const ViaPointContainer* via_points = NULL);
  void updateRobotModel(RobotFootprintModelPtr robot_model) override { robot_model_ = robot_model; }

  bool plan(const std::vector<geometry_msgs::PoseStamped>& initial_plan, const geometry_msgs::Twist* start_vel = NULL, bool free_goal_vel = false) override;
  bool plan(const tf::Pose& start, const tf::Pose& goal, const geometry_msgs::Twist* start_vel = NULL, bool free_goal_vel = false) override;
  bool plan(const PoseSE2& start, const PoseSE2& goal, const geometry_msgs::Twist* start_vel = NULL, bool free_goal_vel = false) override;
  bool getVelocityCommand(double& vx, double& vy, double& omega, int look_ahead_poses) const override;

  /* optimal_planner.h:231 */
  bool optimizeTEB(int iterations_innerloop, int iterations_outerloop, bool compute_cost_afterwards = false,
                   double obst_cost_scale = 1.0, double viapoint_cost_scale = 1.0, bool alternative_time_cost = false);

  void setVelocityStart(const geometry_msgs::Twist& vel_start);
  void setVelocityGoal(const geometry_msgs::Twist& vel_goal);
  void setVelocityGoalFree() { vel_goal_.first = false; }
  void setObstVector(ObstContainer* obst_vector) { obstacles_ = obst_vector; }
  const ObstContainer& getObstVector() const { return *obstacles_; }
  void setViaPoints(const ViaPointContainer* via_points) { via_points_ = via_points; }
  const ViaPointContainer& getViaPoints() const { return *via_points_; }
  void setVisualization(TebVisualizationPtr visualization) { visualization_ = visualization; }
  void clearPlanner() override { clearGraph(); teb_.clearTimedElasticBand(); }
  void setPreferredTurningDir(RotType dir) override { prefer_rotdir_ = dir; }
  TimedElasticBand& teb() { return teb_; }
  const TimedElasticBand& teb() const { return teb_; }
  bool isOptimized() const { return optimized_; }
  bool hasDiverged() const override;
  void computeCurrentCost(double obst_cost_scale = 1.0, double viapoint_cost_scale = 1.0, bool alternative_time_cost = false);
  void computeCurrentCost(std::vector<double>& cost, double obst_cost_scale = 1.0, bool alternative_time_cost = false) override {
    computeCurrentCost(obst_cost_scale, 1.0, alternative_time_cost);
    cost.push_back(getCurrentCost());
  }
  double getCurrentCost() const { return cost_; }
  inline void extractVelocity(const PoseSE2& pose1, const PoseSE2& pose2, double dt, double& vx, double& vy, double& omega) const;
  void getVelocityProfile(std::vector<geometry_msgs::Twist>& velocity_profile) const;
  /* optimal_planner.cpp:1197-1244; TrajectoryPointMsg = pose, velocity, time_from_start [s] */
  void getFullTrajectory(std::vector<TrajectoryPointMsg>& trajectory) const;
  /* optimal_planner.cpp:1247-1306: costmap check of the first poses (+ interpolated poses where consecutive poses are
   * farther apart than the inscribed radius / min_resolution_collision_check_angular) */
  bool isTrajectoryFeasible(base_local_planner::CostmapModel* costmap_model, const std::vector<geometry_msgs::Point>& footprint_spec,
                            double inscribed_radius = 0.0, double circumscribed_radius = 0.0, int look_ahead_idx = -1,
                            double feasibility_check_lookahead_distance = -1.0) override;
  void clearGraph() {}

  /* --- additions of this implementation (not in the reference) --- */
  void setGpuContext(TebGpuContextPtr ctx) { gpu_ = ctx; }
  int lastStatus() const { return status_; }          /* TEB_STATUS_* bits of the last optimizeTEB */
  double lastChi2() const { return chi2_; }
  /* used by HomotopyClassPlanner::optimizeAllTEBs to optimise all candidates in one batched call */
  friend class HomotopyClassPlanner;
  /* calculateEquivalenceClass for several planners in one device call: values[k] = (Re, Im) or one value per obstacle */
  static bool hSignatureBatch(const TebConfig& cfg, const std::vector<TebOptimalPlanner*>& planners,
                              std::shared_ptr<TebGpuContext> gpu, std::vector<std::vector<double>>& values);
  /* shared implementation of optimizeTebBatch / computeCurrentCost: packs the planners into one TebBatch */
  static bool runBatch(const TebConfig& cfg, const std::vector<TebOptimalPlanner*>& planners, int iterations_innerloop,
                       int iterations_outerloop, bool compute_cost_afterwards, double obst_cost_scale,
                       double viapoint_cost_scale, bool alternative_time_cost, std::shared_ptr<TebGpuContext> gpu,
                       bool cost_only, std::vector<std::vector<double>>* hsig_out = nullptr);

 protected:
  const TebConfig* cfg_ = nullptr;
  ObstContainer* obstacles_ = nullptr;
  const ViaPointContainer* via_points_ = nullptr;
  TebVisualizationPtr visualization_;
  RobotFootprintModelPtr robot_model_;
  double cost_ = HUGE_VAL;
  RotType prefer_rotdir_ = RotType::none;
  TimedElasticBand teb_;
  std::pair<bool, geometry_msgs::Twist> vel_start_;
  std::pair<bool, geometry_msgs::Twist> vel_goal_;
  bool initialized_ = false;
  bool optimized_ = false;
  int status_ = 0;
  double chi2_ = 0;
  TebGpuContextPtr gpu_;
};
typedef std::shared_ptr<TebOptimalPlanner> TebOptimalPlannerPtr;
typedef std::shared_ptr<const TebOptimalPlanner> TebOptimalPlannerConstPtr;
typedef std::vector<TebOptimalPlannerPtr> TebOptPlannerContainer;

/* batched optimizeTEB over several planners (the replacement of the per-candidate boost::thread fan-out,
 * homotopy_class_planner.cpp:466-493). Returns false if the C-ABI call failed. */
bool optimizeTebBatch(const TebConfig& cfg, const std::vector<TebOptimalPlanner*>& planners, int iterations_innerloop,
                      int iterations_outerloop, bool compute_cost_afterwards, double obst_cost_scale,
                      double viapoint_cost_scale, bool alternative_time_cost, TebGpuContextPtr gpu);

}  // namespace teb_local_planner
#endif
