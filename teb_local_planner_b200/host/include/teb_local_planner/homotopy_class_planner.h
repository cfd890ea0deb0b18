/*
 * homotopy_class_planner.h — drop-in HomotopyClassPlanner: candidate container + ONE batched GPU optimisation of all
 * candidates + best-candidate selection.
 *
 * Mirrors teb_local_planner::HomotopyClassPlanner (include/teb_local_planner/homotopy_class_planner.h:108-560,
 * src/homotopy_class_planner.cpp): plan() = updateAllTEBs -> exploreEquivalenceClassesAndInitTebs -> via-points ->
 * optimizeAllTEBs -> selectBestTeb (:107-125). optimizeAllTEBs (:466-493) and selectBestTeb (:564-667) are the
 * hot-path rows. The equivalence classes of the candidates (H-signatures, h_signature.h) are computed for ALL candidates
 * in one device call and filtered with the reference's first-come-first-serve rule (renewAndAnalyzeOldTebs :214-256,
 * addEquivalenceClassIfNew :189-211). New candidates are proposed by the graph search (graph_search.h: lrKeyPointGraph when
 * hcp.simple_exploration, else ProbRoadmapGraph, + depth-first enumeration); a proposed path becomes a band only if its
 * class is new (addAndInitNewTebs).
 */
#ifndef TEB_B200_HOMOTOPY_CLASS_PLANNER_H_
#define TEB_B200_HOMOTOPY_CLASS_PLANNER_H_

#include <chrono>
#include <random>

#include "teb_local_planner/graph_search.h"
#include "teb_local_planner/h_signature.h"
#include "teb_local_planner/optimal_planner.h"

namespace teb_local_planner {

class HomotopyClassPlanner : public PlannerInterface {
 public:
  HomotopyClassPlanner();
  HomotopyClassPlanner(const TebConfig& cfg, ObstContainer* obstacles = NULL, TebVisualizationPtr visualization = TebVisualizationPtr(),
                       const ViaPointContainer* via_points = NULL);
  virtual ~HomotopyClassPlanner() {}
  void initialize(const TebConfig& cfg, ObstContainer* obstacles = NULL, TebVisualizationPtr visualization = TebVisualizationPtr(),
                  const ViaPointContainer* via_points = NULL);
  void updateRobotModel(RobotFootprintModelPtr robot_model) override { robot_model_ = robot_model; }

  bool plan(const std::vector<geometry_msgs::PoseStamped>& initial_plan, const geometry_msgs::Twist* start_vel = NULL, bool free_goal_vel = false) override;
  bool plan(const tf::Pose& start, const tf::Pose& goal, const geometry_msgs::Twist* start_vel = NULL, bool free_goal_vel = false) override;
  bool plan(const PoseSE2& start, const PoseSE2& goal, const geometry_msgs::Twist* start_vel = NULL, bool free_goal_vel = false) override;
  bool getVelocityCommand(double& vx, double& vy, double& omega, int look_ahead_poses) const override;

  TebOptimalPlannerPtr bestTeb() const { return tebs_.empty() ? TebOptimalPlannerPtr() : tebs_.size() == 1 ? tebs_.front() : best_teb_; }
  void exploreEquivalenceClassesAndInitTebs(const PoseSE2& start, const PoseSE2& goal, double dist_to_obst,
                                            const geometry_msgs::Twist* start_vel, bool free_goal_vel = false);
  /* homotopy_class_planner.cpp:359 (start/goal straight-line init) */
  TebOptimalPlannerPtr addAndInitNewTeb(const PoseSE2& start, const PoseSE2& goal, const geometry_msgs::Twist* start_velocity,
                                        bool free_goal_vel = false);
  /* homotopy_class_planner.hpp:67-100 for a chunk of graph-search paths: bands initialised from the paths, ONE device call
   * for their equivalence classes, accepted in order while fewer than max_number_classes bands exist. false = device error */
  bool addAndInitNewTebs(const std::vector<std::vector<Eigen::Vector2d>>& paths, double start_orientation, double goal_orientation,
                         const geometry_msgs::Twist* start_velocity, bool free_goal_vel = false);
  /* homotopy_class_planner.cpp:414 (init from a reference path) */
  TebOptimalPlannerPtr addAndInitNewTeb(const std::vector<geometry_msgs::PoseStamped>& initial_plan,
                                        const geometry_msgs::Twist* start_velocity, bool free_goal_vel = false);
  void updateAllTEBs(const PoseSE2* start, const PoseSE2* goal, const geometry_msgs::Twist* start_velocity);
  /* homotopy_class_planner.cpp:304-335 */
  void updateReferenceTrajectoryViaPoints(bool all_trajectories);
  void optimizeAllTEBs(int iter_innerloop, int iter_outerloop);
  TebOptimalPlannerPtr getInitialPlanTEB();
  TebOptimalPlannerPtr selectBestTeb();
  /* equivalence classes (homotopy_class_planner.h:388-520 of the reference) */
  std::vector<EquivalenceClassPtr> calculateEquivalenceClasses(const std::vector<TebOptimalPlanner*>& planners);
  EquivalenceClassPtr calculateEquivalenceClass(TebOptimalPlanner* planner);
  bool addEquivalenceClassIfNew(const EquivalenceClassPtr& eq_class, bool lock = false);
  bool hasEquivalenceClass(const EquivalenceClassPtr& eq_class) const;
  bool isInBestTebClass(const EquivalenceClassPtr& eq_class) const;
  int numTebsInClass(const EquivalenceClassPtr& eq_class) const;
  int numTebsInBestTebClass() const;
  void renewAndAnalyzeOldTebs(bool delete_detours);
  /* :766-838: drop bands that start against the current direction of motion, could not be optimised or take much longer
   * than the best band */
  void deletePlansDetouringBackwards(const double orient_threshold, const double len_orientation_vector);
  bool computeStartOrientation(const TebOptimalPlannerPtr plan, const double len_orientation_vector, double& orientation);
  /* :539-562: with probability selection_dropping_probability a band other than the best is dropped (default 0) */
  void randomlyDropTebs();
  const EquivalenceClassContainer& getEquivalenceClassRef() const { return equivalence_classes_; }
  void clearPlanner() override;
  void setPreferredTurningDir(RotType dir) override;
  const TebOptPlannerContainer& getTrajectoryContainer() const { return tebs_; }
  bool hasDiverged() const override;
  /* homotopy_class_planner.cpp:686-707: checks the best band, drops it if infeasible and tries the next best */
  bool isTrajectoryFeasible(base_local_planner::CostmapModel* costmap_model, const std::vector<geometry_msgs::Point>& footprint_spec,
                            double inscribed_radius = 0.0, double circumscribed_radius = 0.0, int look_ahead_idx = -1,
                            double feasibility_check_lookahead_distance = -1.0) override;
  TebOptimalPlannerPtr findBestTeb();
  void computeCurrentCost(std::vector<double>& cost, double obst_cost_scale = 1.0, double viapoint_cost_scale = 1.0,
                          bool alternative_time_cost = false);
  void computeCurrentCost(std::vector<double>& cost, double obst_cost_scale, bool alternative_time_cost) override {
    computeCurrentCost(cost, obst_cost_scale, 1.0, alternative_time_cost);
  }
  int bestTebIdx() const;
  TebOptPlannerContainer::iterator removeTeb(TebOptimalPlannerPtr& teb);
  const TebConfig* config() const { return cfg_; }
  const ObstContainer* obstacles() const { return obstacles_; }
  bool isInitialized() const { return initialized_; }
  void setGpuContext(TebGpuContextPtr ctx) { gpu_ = ctx; }
  /* one context per device: optimizeAllTEBs shards the candidates over them (contiguous split, one host thread per device) */
  void setGpuContexts(const std::vector<TebGpuContextPtr>& ctxs) { gpus_ = ctxs; if (!ctxs.empty()) gpu_ = ctxs.front(); }

 protected:
  const TebConfig* cfg_ = nullptr;
  ObstContainer* obstacles_ = nullptr;
  const ViaPointContainer* via_points_ = nullptr;
  TebVisualizationPtr visualization_;
  TebOptimalPlannerPtr best_teb_;
  TebOptimalPlannerPtr last_best_teb_;
  TebOptimalPlannerPtr initial_plan_teb_;
  RobotFootprintModelPtr robot_model_;
  const std::vector<geometry_msgs::PoseStamped>* initial_plan_ = nullptr;
  TebOptPlannerContainer tebs_;
  EquivalenceClassContainer equivalence_classes_;
  EquivalenceClassPtr best_teb_eq_class_;
  EquivalenceClassPtr initial_plan_eq_class_;
  std::shared_ptr<GraphSearchInterface> graph_search_;
  std::default_random_engine random_;
  std::chrono::steady_clock::time_point last_eq_class_switching_time_;
  bool initialized_ = false;
  TebGpuContextPtr gpu_;
  std::vector<TebGpuContextPtr> gpus_;
};
typedef std::shared_ptr<HomotopyClassPlanner> HomotopyClassPlannerPtr;

}  // namespace teb_local_planner
#endif
