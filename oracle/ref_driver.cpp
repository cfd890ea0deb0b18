/*
 * ref_driver.cpp — C interface over the REFERENCE'S OWN sources, compiled where they lie under /root/reference
 * (TEST INFRASTRUCTURE; builds oracle/_ref/libteb_ref.so, see oracle/Makefile target `ref`).
 *
 * Compiled from the reference, unmodified: src/optimal_planner.cpp (optimizeTEB, buildGraph, every AddEdges*,
 * optimizeGraph, computeCurrentCost, hasDiverged), src/timed_elastic_band.cpp (+ .hpp: autoResize, initTrajectoryToGoal,
 * updateAndPruneTEB, findClosestTrajectoryPose), src/obstacles.cpp, and every header they include from
 * include/teb_local_planner/ (g2o_types/edge_*.h computeError / linearizeOplus bodies, penalties.h, misc.h,
 * distance_calculations.h, obstacles.h, robot_footprint_model.h, pose_se2.h, teb_config.h incl. the TebConfig() defaults).
 * NOT from the reference (absent third-party code, replaced by oracle/ref_shims/): Eigen (a small fixed-size matrix
 * class), boost / ROS message types (plain structs), and g2o - its vertex / edge base classes, numeric linearizeOplus,
 * Levenberg-Marquardt loop and linear solver are restated in ref_shims/g2o/ following SURVEY.md Appendix A.
 * So oracle/_ref pins the oracle restatement (oracle/teb_oracle.c) for everything that IS reference code; only the g2o
 * optimizer internals remain "restated from published behaviour".
 */
#include <teb_local_planner/optimal_planner.h>
#include <teb_local_planner/g2o_types/penalties.h>
#include <teb_local_planner/h_signature.h>
#include <teb_local_planner/homotopy_class_planner.h>

#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/teb_b200.h"

using namespace teb_local_planner;

/* teb_config.cpp is not compiled (ROS parameter server); the few out-of-line members are never called here */
namespace {

struct Scene {
  TebConfig cfg;
  ObstContainer obstacles;
  ViaPointContainer via;
};

void fill_config(const TebParams* p, TebConfig& c) {
  c.trajectory.teb_autosize = p->teb_autosize;
  c.trajectory.dt_ref = p->dt_ref;
  c.trajectory.dt_hysteresis = p->dt_hysteresis;
  c.trajectory.min_samples = p->min_samples;
  c.trajectory.max_samples = p->max_samples;
  c.trajectory.exact_arc_length = p->exact_arc_length != 0;
  c.trajectory.via_points_ordered = p->via_points_ordered != 0;
  c.trajectory.allow_init_with_backwards_motion = p->allow_init_with_backwards_motion != 0;
  c.trajectory.global_plan_overwrite_orientation = p->global_plan_overwrite_orientation != 0;
  c.trajectory.force_reinit_new_goal_dist = p->force_reinit_new_goal_dist;
  c.trajectory.force_reinit_new_goal_angular = p->force_reinit_new_goal_angular;
  c.robot.max_vel_x = p->max_vel_x;
  c.robot.max_vel_x_backwards = p->max_vel_x_backwards;
  c.robot.max_vel_y = p->max_vel_y;
  c.robot.max_vel_trans = p->max_vel_trans;
  c.robot.max_vel_theta = p->max_vel_theta;
  c.robot.acc_lim_x = p->acc_lim_x;
  c.robot.acc_lim_y = p->acc_lim_y;
  c.robot.acc_lim_theta = p->acc_lim_theta;
  c.robot.min_turning_radius = p->min_turning_radius;
  c.obstacles.min_obstacle_dist = p->min_obstacle_dist;
  c.obstacles.inflation_dist = p->inflation_dist;
  c.obstacles.dynamic_obstacle_inflation_dist = p->dynamic_obstacle_inflation_dist;
  c.obstacles.include_dynamic_obstacles = p->include_dynamic_obstacles != 0;
  c.obstacles.legacy_obstacle_association = p->legacy_obstacle_association != 0;
  c.obstacles.obstacle_poses_affected = p->obstacle_poses_affected;
  c.obstacles.obstacle_association_force_inclusion_factor = p->obstacle_association_force_inclusion_factor;
  c.obstacles.obstacle_association_cutoff_factor = p->obstacle_association_cutoff_factor;
  c.obstacles.obstacle_proximity_ratio_max_vel = p->obstacle_proximity_ratio_max_vel;
  c.obstacles.obstacle_proximity_lower_bound = p->obstacle_proximity_lower_bound;
  c.obstacles.obstacle_proximity_upper_bound = p->obstacle_proximity_upper_bound;
  c.optim.no_inner_iterations = p->no_inner_iterations;
  c.optim.no_outer_iterations = p->no_outer_iterations;
  c.optim.optimization_activate = p->optimization_activate != 0;
  c.optim.optimization_verbose = false;
  c.optim.penalty_epsilon = p->penalty_epsilon;
  c.optim.weight_max_vel_x = p->weight_max_vel_x;
  c.optim.weight_max_vel_y = p->weight_max_vel_y;
  c.optim.weight_max_vel_theta = p->weight_max_vel_theta;
  c.optim.weight_acc_lim_x = p->weight_acc_lim_x;
  c.optim.weight_acc_lim_y = p->weight_acc_lim_y;
  c.optim.weight_acc_lim_theta = p->weight_acc_lim_theta;
  c.optim.weight_kinematics_nh = p->weight_kinematics_nh;
  c.optim.weight_kinematics_forward_drive = p->weight_kinematics_forward_drive;
  c.optim.weight_kinematics_turning_radius = p->weight_kinematics_turning_radius;
  c.optim.weight_optimaltime = p->weight_optimaltime;
  c.optim.weight_shortest_path = p->weight_shortest_path;
  c.optim.weight_obstacle = p->weight_obstacle;
  c.optim.weight_inflation = p->weight_inflation;
  c.optim.weight_dynamic_obstacle = p->weight_dynamic_obstacle;
  c.optim.weight_dynamic_obstacle_inflation = p->weight_dynamic_obstacle_inflation;
  c.optim.weight_velocity_obstacle_ratio = p->weight_velocity_obstacle_ratio;
  c.optim.weight_viapoint = p->weight_viapoint;
  c.optim.weight_prefer_rotdir = p->weight_prefer_rotdir;
  c.optim.weight_adapt_factor = p->weight_adapt_factor;
  c.optim.obstacle_cost_exponent = p->obstacle_cost_exponent;
  c.hcp.selection_cost_hysteresis = p->selection_cost_hysteresis;
  c.hcp.selection_prefer_initial_plan = p->selection_prefer_initial_plan;
  c.hcp.selection_obst_cost_scale = p->selection_obst_cost_scale;
  c.hcp.selection_viapoint_cost_scale = p->selection_viapoint_cost_scale;
  c.hcp.selection_alternative_time_cost = p->selection_alternative_time_cost != 0;
  c.hcp.enable_multithreading = p->enable_multithreading != 0;
  c.hcp.h_signature_prescaler = p->h_signature_prescaler;
  c.hcp.h_signature_threshold = p->h_signature_threshold;
  /* left uninitialised by TebConfig's constructor (teb_config.h:245-390); a ROS node gets it from dynamic_reconfigure
   * (cfg/TebLocalPlannerReconfigure.cfg:350, default 1) - without this line addEquivalenceClassIfNew reads garbage */
  c.hcp.max_number_plans_in_current_class = 1;
  c.recovery.divergence_detection_enable = p->divergence_detection_enable != 0;
  c.recovery.divergence_detection_max_chi_squared = (int)p->divergence_detection_max_chi_squared;
  switch (p->footprint_type) {
    case TEB_FOOTPRINT_CIRCULAR: c.robot_model = boost::make_shared<CircularRobotFootprint>(p->footprint_radius); break;
    case TEB_FOOTPRINT_TWO_CIRCLES:
      c.robot_model = boost::make_shared<TwoCirclesRobotFootprint>(p->footprint_front_offset, p->footprint_front_radius,
                                                                    p->footprint_rear_offset, p->footprint_rear_radius);
      break;
    case TEB_FOOTPRINT_LINE:
      c.robot_model = boost::make_shared<LineRobotFootprint>(Eigen::Vector2d(p->footprint_line[0], p->footprint_line[1]),
                                                              Eigen::Vector2d(p->footprint_line[2], p->footprint_line[3]), 0.0);
      break;
    case TEB_FOOTPRINT_POLYGON: {
      Point2dContainer v;
      for (int k = 0; k < p->footprint_vertex_count; ++k) v.push_back(Eigen::Vector2d(p->footprint_vertices[2 * k], p->footprint_vertices[2 * k + 1]));
      c.robot_model = boost::make_shared<PolygonRobotFootprint>(v);
      break;
    }
    default: c.robot_model = boost::make_shared<PointRobotFootprint>(); break;
  }
}

ObstaclePtr make_obstacle(const TebObstacle& o, const double* verts) {
  ObstaclePtr ob;
  switch (o.type) {
    case TEB_OBST_CIRCULAR: ob = ObstaclePtr(new CircularObstacle(o.x, o.y, o.radius)); break;
    case TEB_OBST_LINE: {
      const double* v = verts + 2 * (size_t)o.vertex_begin;
      ob = ObstaclePtr(new LineObstacle(v[0], v[1], v[2], v[3]));
      break;
    }
    case TEB_OBST_PILL: {
      const double* v = verts + 2 * (size_t)o.vertex_begin;
      ob = ObstaclePtr(new PillObstacle(v[0], v[1], v[2], v[3], o.radius));
      break;
    }
    case TEB_OBST_POLYGON: {
      PolygonObstacle* po = new PolygonObstacle();
      for (int k = 0; k < o.vertex_count; ++k) po->pushBackVertex(verts[2 * (size_t)(o.vertex_begin + k)], verts[2 * (size_t)(o.vertex_begin + k) + 1]);
      po->finalizePolygon();
      ob = ObstaclePtr(po);
      break;
    }
    default: ob = ObstaclePtr(new PointObstacle(o.x, o.y)); break;
  }
  if (o.dynamic) ob->setCentroidVelocity(Eigen::Vector2d(o.vx, o.vy));
  return ob;
}

void fill_scene(Scene& s, const TebParams* p, const TebObstacle* obst, int M, const double* verts, const double* via, int V) {
  fill_config(p, s.cfg);
  for (int m = 0; m < M; ++m) s.obstacles.push_back(make_obstacle(obst[m], verts));
  for (int v = 0; v < V; ++v) s.via.push_back(Eigen::Vector2d(via[2 * v], via[2 * v + 1]));
}

/* opens the protected graph functions of the reference planner */
class RefPlanner : public TebOptimalPlanner {
 public:
  RefPlanner(const TebConfig& cfg, ObstContainer* obst, const ViaPointContainer* via) : TebOptimalPlanner(cfg, obst, TebVisualizationPtr(), via) {}
  using TebOptimalPlanner::buildGraph;
  using TebOptimalPlanner::clearGraph;
};

void load_band(TimedElasticBand& teb, const double* rec, int n) {
  /* same container state as initTrajectoryToGoal leaves behind: first and last pose fixed (timed_elastic_band.cpp:330,377) */
  for (int i = 0; i < n; ++i) {
    teb.addPose(rec[4 * i], rec[4 * i + 1], rec[4 * i + 2], i == 0 || i == n - 1);
    if (i < n - 1) teb.addTimeDiff(rec[4 * i + 3]);
  }
}
void store_band(const TimedElasticBand& teb, double* rec, int n_cap) {
  const int n = teb.sizePoses();
  for (int i = 0; i < n && i < n_cap; ++i) {
    rec[4 * i] = teb.Pose(i).x(); rec[4 * i + 1] = teb.Pose(i).y(); rec[4 * i + 2] = teb.Pose(i).theta();
    rec[4 * i + 3] = i < teb.sizeTimeDiffs() ? teb.TimeDiff(i) : 0.0;
  }
}
void set_velocities(RefPlanner& pl, const double* vs, const double* vg, int rotdir) {
  /* vel_start_.first / vel_goal_.first are true with zero twists after construction (optimal_planner.cpp:94-102) */
  if (vs && vs[3] != 0) {
    geometry_msgs::Twist t;
    t.linear.x = vs[0]; t.linear.y = vs[1]; t.angular.z = vs[2];
    pl.setVelocityStart(t);
  }
  if (vg) {
    if (vg[3] != 0) {
      geometry_msgs::Twist t;
      t.linear.x = vg[0]; t.linear.y = vg[1]; t.angular.z = vg[2];
      pl.setVelocityGoal(t);
    } else {
      pl.setVelocityGoalFree();
    }
  }
  pl.setPreferredTurningDir(rotdir == TEB_ROTDIR_LEFT ? RotType::left : (rotdir == TEB_ROTDIR_RIGHT ? RotType::right : RotType::none));
}

}  // namespace

extern "C" {

int32_t teb_ref_abi(void) { return 6; }

/* TebConfig::TebConfig() (teb_config.h:245-390) read back through the POD mirror */
void teb_ref_default_params(TebParams* p) {
  std::memset(p, 0, sizeof(*p));
  TebConfig c;
  p->dt_ref = c.trajectory.dt_ref; p->dt_hysteresis = c.trajectory.dt_hysteresis;
  p->force_reinit_new_goal_dist = c.trajectory.force_reinit_new_goal_dist;
  p->force_reinit_new_goal_angular = c.trajectory.force_reinit_new_goal_angular;
  p->teb_autosize = c.trajectory.teb_autosize != 0; p->min_samples = c.trajectory.min_samples; p->max_samples = c.trajectory.max_samples;
  p->exact_arc_length = c.trajectory.exact_arc_length; p->via_points_ordered = c.trajectory.via_points_ordered;
  p->allow_init_with_backwards_motion = c.trajectory.allow_init_with_backwards_motion;
  p->global_plan_overwrite_orientation = c.trajectory.global_plan_overwrite_orientation;
  p->max_vel_x = c.robot.max_vel_x; p->max_vel_x_backwards = c.robot.max_vel_x_backwards; p->max_vel_y = c.robot.max_vel_y;
  p->max_vel_trans = c.robot.max_vel_trans; p->max_vel_theta = c.robot.max_vel_theta;
  p->acc_lim_x = c.robot.acc_lim_x; p->acc_lim_y = c.robot.acc_lim_y; p->acc_lim_theta = c.robot.acc_lim_theta;
  p->min_turning_radius = c.robot.min_turning_radius;
  p->footprint_type = TEB_FOOTPRINT_POINT;
  p->min_obstacle_dist = c.obstacles.min_obstacle_dist; p->inflation_dist = c.obstacles.inflation_dist;
  p->dynamic_obstacle_inflation_dist = c.obstacles.dynamic_obstacle_inflation_dist;
  p->obstacle_association_force_inclusion_factor = c.obstacles.obstacle_association_force_inclusion_factor;
  p->obstacle_association_cutoff_factor = c.obstacles.obstacle_association_cutoff_factor;
  p->obstacle_proximity_ratio_max_vel = c.obstacles.obstacle_proximity_ratio_max_vel;
  p->obstacle_proximity_lower_bound = c.obstacles.obstacle_proximity_lower_bound;
  p->obstacle_proximity_upper_bound = c.obstacles.obstacle_proximity_upper_bound;
  p->include_dynamic_obstacles = c.obstacles.include_dynamic_obstacles;
  p->legacy_obstacle_association = c.obstacles.legacy_obstacle_association;
  p->obstacle_poses_affected = c.obstacles.obstacle_poses_affected;
  p->penalty_epsilon = c.optim.penalty_epsilon;
  p->weight_max_vel_x = c.optim.weight_max_vel_x; p->weight_max_vel_y = c.optim.weight_max_vel_y; p->weight_max_vel_theta = c.optim.weight_max_vel_theta;
  p->weight_acc_lim_x = c.optim.weight_acc_lim_x; p->weight_acc_lim_y = c.optim.weight_acc_lim_y; p->weight_acc_lim_theta = c.optim.weight_acc_lim_theta;
  p->weight_kinematics_nh = c.optim.weight_kinematics_nh; p->weight_kinematics_forward_drive = c.optim.weight_kinematics_forward_drive;
  p->weight_kinematics_turning_radius = c.optim.weight_kinematics_turning_radius;
  p->weight_optimaltime = c.optim.weight_optimaltime; p->weight_shortest_path = c.optim.weight_shortest_path;
  p->weight_obstacle = c.optim.weight_obstacle; p->weight_inflation = c.optim.weight_inflation;
  p->weight_dynamic_obstacle = c.optim.weight_dynamic_obstacle; p->weight_dynamic_obstacle_inflation = c.optim.weight_dynamic_obstacle_inflation;
  p->weight_velocity_obstacle_ratio = c.optim.weight_velocity_obstacle_ratio;
  p->weight_viapoint = c.optim.weight_viapoint; p->weight_prefer_rotdir = c.optim.weight_prefer_rotdir;
  p->weight_adapt_factor = c.optim.weight_adapt_factor; p->obstacle_cost_exponent = c.optim.obstacle_cost_exponent;
  p->no_inner_iterations = c.optim.no_inner_iterations; p->no_outer_iterations = c.optim.no_outer_iterations;
  p->optimization_activate = c.optim.optimization_activate;
  p->selection_cost_hysteresis = c.hcp.selection_cost_hysteresis; p->selection_prefer_initial_plan = c.hcp.selection_prefer_initial_plan;
  p->selection_obst_cost_scale = c.hcp.selection_obst_cost_scale; p->selection_viapoint_cost_scale = c.hcp.selection_viapoint_cost_scale;
  p->selection_alternative_time_cost = c.hcp.selection_alternative_time_cost; p->enable_multithreading = c.hcp.enable_multithreading;
  p->h_signature_prescaler = c.hcp.h_signature_prescaler; p->h_signature_threshold = c.hcp.h_signature_threshold;
  /* recovery.divergence_detection_* are left uninitialised by the reference's constructor: the
   * dynamic_reconfigure defaults are reported instead */
  p->divergence_detection_enable = 0; p->divergence_detection_max_chi_squared = 10;
}

/* penalties.h: which 0 penaltyBoundToInterval(var,a,eps) 1 (var,a,b,eps) 2 penaltyBoundFromBelow 3..5 their derivatives */
double teb_ref_penalty(int32_t which, double var, double a, double b, double eps) {
  switch (which) {
    case 0: return penaltyBoundToInterval(var, a, eps);
    case 1: return penaltyBoundToInterval(var, a, b, eps);
    case 2: return penaltyBoundFromBelow(var, a, eps);
    case 3: return penaltyBoundToIntervalDerivative(var, a, eps);
    case 4: return penaltyBoundToIntervalDerivative(var, a, b, eps);
    case 5: return penaltyBoundFromBelowDerivative(var, a, eps);
    default: return NAN;
  }
}
double teb_ref_fast_sigmoid(double x) { return fast_sigmoid(x); }

/* robot_model->calculateDistance (t < 0) / estimateSpatioTemporalDistance (t >= 0) on the reference's own classes */
double teb_ref_distance(const TebParams* p, const double* pose3, const TebObstacle* o, const double* verts, double t) {
  TebConfig cfg;
  fill_config(p, cfg);
  ObstaclePtr ob = make_obstacle(*o, verts);
  const PoseSE2 pose(pose3[0], pose3[1], pose3[2]);
  if (t < 0) return cfg.robot_model->calculateDistance(pose, ob.get());
  return cfg.robot_model->estimateSpatioTemporalDistance(pose, ob.get(), t);
}

/* TimedElasticBand::autoResize (timed_elastic_band.cpp:227-286) */
int32_t teb_ref_auto_resize(double* rec, int32_t n, int32_t n_cap, double dt_ref, double dt_hysteresis, int32_t min_samples,
                            int32_t max_samples, int32_t fast_mode) {
  TimedElasticBand teb;
  load_band(teb, rec, n);
  teb.autoResize(dt_ref, dt_hysteresis, min_samples, max_samples, fast_mode != 0);
  if (teb.sizePoses() > n_cap) return -1;
  store_band(teb, rec, n_cap);
  return teb.sizePoses();
}

/* TimedElasticBand::initTrajectoryToGoal(start, goal, diststep, max_vel_x, min_samples, guess_backwards_motion) :325-387 */
int32_t teb_ref_init_trajectory(const double* start3, const double* goal3, double diststep, double max_vel_x, int32_t min_samples,
                                int32_t backwards, double* rec, int32_t n_cap) {
  TimedElasticBand teb;
  teb.initTrajectoryToGoal(PoseSE2(start3[0], start3[1], start3[2]), PoseSE2(goal3[0], goal3[1], goal3[2]), diststep, max_vel_x,
                           min_samples, backwards != 0);
  if (teb.sizePoses() > n_cap) return -1;
  store_band(teb, rec, n_cap);
  return teb.sizePoses();
}

/* updateAndPruneTEB (timed_elastic_band.cpp:555-597) */
int32_t teb_ref_update_and_prune(double* rec, int32_t n, int32_t n_cap, const double* new_start3, const double* new_goal3, int32_t min_samples) {
  TimedElasticBand teb;
  load_band(teb, rec, n);
  const PoseSE2 s(new_start3[0], new_start3[1], new_start3[2]), g(new_goal3[0], new_goal3[1], new_goal3[2]);
  teb.updateAndPruneTEB(s, g, min_samples);
  store_band(teb, rec, n_cap);
  return teb.sizePoses();
}

/* TebOptimalPlanner::optimizeTEB (optimal_planner.cpp:182-231) on one band. stats: [0] lm_trials [1] rejected trials
 * [2] last optimize() ended with Terminate [3] a factorisation failed [4] hasDiverged() [5] isOptimized(). Returns 1/0 =
 * optimizeTEB's return value. */
int32_t teb_ref_optimize(const TebParams* p, double* rec, int32_t* n_io, int32_t n_cap, const TebObstacle* obst, int32_t M,
                         const double* verts, const double* via, int32_t V, const double* vel_start4, const double* vel_goal4,
                         int32_t rotdir, const TebOptimizeArgs* args, double* cost_out, double* stats6) {
  Scene s;
  fill_scene(s, p, obst, M, verts, via, V);
  RefPlanner pl(s.cfg, &s.obstacles, &s.via);
  load_band(pl.teb(), rec, *n_io);
  set_velocities(pl, vel_start4, vel_goal4, rotdir);
  const bool ok = pl.optimizeTEB(args->iterations_innerloop, args->iterations_outerloop, args->compute_cost_afterwards != 0,
                                 args->obst_cost_scale, args->viapoint_cost_scale, args->alternative_time_cost != 0);
  if (pl.teb().sizePoses() > n_cap) return -1;
  store_band(pl.teb(), rec, n_cap);
  *n_io = pl.teb().sizePoses();
  if (cost_out) *cost_out = pl.getCurrentCost();
  if (stats6) {
    stats6[0] = (double)pl.optimizer()->lm_trials; stats6[1] = (double)pl.optimizer()->lm_rejected;
    stats6[2] = pl.optimizer()->last_terminated; stats6[3] = pl.optimizer()->chol_failed;
    stats6[4] = pl.hasDiverged(); stats6[5] = pl.isOptimized();
  }
  return ok ? 1 : 0;
}

/* computeCurrentCost called outside optimizeTEB (optimal_planner.cpp:1041-1094: graph rebuilt with multiplier 1) */
double teb_ref_compute_cost(const TebParams* p, const double* rec, int32_t n, const TebObstacle* obst, int32_t M, const double* verts,
                            const double* via, int32_t V, const double* vel_start4, const double* vel_goal4, int32_t rotdir,
                            const TebOptimizeArgs* args) {
  Scene s;
  fill_scene(s, p, obst, M, verts, via, V);
  RefPlanner pl(s.cfg, &s.obstacles, &s.via);
  load_band(pl.teb(), rec, n);
  set_velocities(pl, vel_start4, vel_goal4, rotdir);
  pl.computeCurrentCost(args->obst_cost_scale, args->viapoint_cost_scale, args->alternative_time_cost != 0);
  return pl.getCurrentCost();
}

/* buildGraph(weight_multiplier) + computeActiveErrors + buildSystem. H_dense [N][N] (full symmetric), b [N], N = system
 * size in g2o order (dt_0, pose_1, dt_1, ...); chi2 = activeChi2. edges_out (optional, up to max_edges rows of 64
 * doubles): per active edge in insertion order: [0] error dimension, [1] vertex count, [2..4] error, [5..7] information
 * diagonal, [8 + 9 k ...] Jacobian of vertex k (row major dim x vdim, zero for fixed vertices), [53 + k] vertex dimension,
 * [58 + k] vertex id. Returns N (or -1); *n_edges receives the number of active edges. */
int32_t teb_ref_build_system(const TebParams* p, const double* rec, int32_t n, const TebObstacle* obst, int32_t M, const double* verts,
                             const double* via, int32_t V, const double* vel_start4, const double* vel_goal4, int32_t rotdir,
                             double weight_multiplier, double* H_dense, double* b, double* chi2, double* edges_out, int32_t max_edges,
                             int32_t* n_edges) {
  Scene s;
  fill_scene(s, p, obst, M, verts, via, V);
  RefPlanner pl(s.cfg, &s.obstacles, &s.via);
  load_band(pl.teb(), rec, n);
  set_velocities(pl, vel_start4, vel_goal4, rotdir);
  if (!pl.buildGraph(weight_multiplier)) return -1;
  g2o::SparseOptimizer& opt = *pl.optimizer();
  opt.initializeOptimization();
  opt.buildSystemOnly();
  const int N = opt.systemSize(), W = opt.halfBandwidth() + 1;
  if (H_dense) {
    std::memset(H_dense, 0, sizeof(double) * (size_t)N * N);
    for (int r = 0; r < N; ++r)
      for (int k = 0; k < W && k <= r; ++k) {
        const double v = opt.bandedH()[(size_t)r * W + k];
        H_dense[(size_t)r * N + (r - k)] = v;
        H_dense[(size_t)(r - k) * N + r] = v;
      }
  }
  if (b) std::memcpy(b, opt.rhs().data(), sizeof(double) * (size_t)N);
  double c2 = 0;
  const g2o::SparseOptimizer::EdgeContainer& act = opt.activeEdges();
  for (size_t k = 0; k < act.size(); ++k) {
    g2o::OptimizableGraph::Edge* e = act[k];
    const int D = e->dimension();
    for (int d = 0; d < D; ++d) c2 += e->errorData()[d] * e->informationData()[d + D * d] * e->errorData()[d];
    if (edges_out && (int)k < max_edges) {
      double* row = edges_out + 64 * k;
      std::memset(row, 0, 64 * sizeof(double));
      row[0] = D; row[1] = (double)e->numVertices();
      for (int d = 0; d < D; ++d) { row[2 + d] = e->errorData()[d]; row[5 + d] = e->informationData()[d + D * d]; }
      for (size_t v = 0; v < e->numVertices() && v < 5; ++v) {
        const int vd = e->vertexAt(v)->dimension();
        row[53 + v] = vd; row[58 + v] = e->vertexAt(v)->id();
        if (e->vertexAt(v)->fixed()) continue;
        const double* J = e->jacobianData(v); /* column major D x vd */
        for (int d = 0; d < D; ++d)
          for (int a = 0; a < vd; ++a) row[8 + 9 * v + d * vd + a] = J[d + D * a];
      }
    }
  }
  if (chi2) *chi2 = c2;
  if (n_edges) *n_edges = (int32_t)act.size();
  pl.clearGraph();
  return N;
}

/* Whole batch with the TebBatch layout: `threads` host threads, one band at a time per thread - the reference's own
 * optimizeAllTEBs model (one boost::thread per candidate, homotopy_class_planner.cpp:466-493). pin != 0: worker t runs on
 * the t-th CPU of the affinity mask. Used as the CPU arm of bench.py when this library exists. */
int32_t teb_ref_optimize_batch(const TebParams* p, const TebBatch* bt, const TebOptimizeArgs* args, int32_t threads, int32_t pin) {
  if (threads < 1) threads = 1;
  if (threads > bt->B) threads = bt->B;
  std::atomic<int> next(0), bad(0);
  auto work = [&](int index) {
    if (pin && index >= 0) {
      cpu_set_t allowed;
      CPU_ZERO(&allowed);
      if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
        const int cnt = CPU_COUNT(&allowed);
        int want = cnt > 0 ? index % cnt : 0, seen = 0;
        for (int c = 0; c < CPU_SETSIZE && cnt > 0; ++c) {
          if (!CPU_ISSET(c, &allowed)) continue;
          if (seen++ == want) { cpu_set_t one; CPU_ZERO(&one); CPU_SET(c, &one); pthread_setaffinity_np(pthread_self(), sizeof(one), &one); break; }
        }
      }
    }
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= bt->B) break;
      const int s = bt->scene_id ? bt->scene_id[b] : 0;
      int32_t nb = bt->n[b];
      double cost = HUGE_VAL, st[6];
      const double one4[4] = {0, 0, 0, 1};
      const int32_t rc = teb_ref_optimize(p, bt->poses + (size_t)b * bt->n_cap * 4, &nb, bt->n_cap, bt->obstacles + (size_t)s * bt->M_cap,
                                          bt->obst_count ? bt->obst_count[s] : 0,
                                          (bt->obst_vertices && bt->PV_cap > 0) ? bt->obst_vertices + (size_t)s * bt->PV_cap * 2 : nullptr,
                                          bt->via ? bt->via + (size_t)b * bt->V_cap * 2 : nullptr, bt->via_count ? bt->via_count[b] : 0,
                                          bt->vel_start ? bt->vel_start + 4 * b : one4, bt->vel_goal ? bt->vel_goal + 4 * b : one4,
                                          bt->prefer_rotdir ? bt->prefer_rotdir[b] : 0, args, &cost, st);
      if (rc < 0) { bad = 1; continue; }
      bt->n[b] = nb;
      if (bt->cost) bt->cost[b] = cost;
      if (bt->status) bt->status[b] = (rc == 1 ? TEB_STATUS_OPTIMIZED : 0) | (st[2] != 0 ? TEB_STATUS_TERMINATED : 0) | (st[3] != 0 ? TEB_STATUS_CHOL_FAILED : 0);
      if (bt->lm_iters) bt->lm_iters[b] = 0;
    }
  };
  if (threads == 1) {
    work(-1);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  return bad ? -1 : 0;
}

/* HomotopyClassPlanner::calculateEquivalenceClass (homotopy_class_planner.hpp:46-63) on the poses of one band, executed
 * by the reference's own h_signature.h: include_dynamic_obstacles == 0 -> HSignature (out[0..1] = Re, Im of the long
 * double value), else HSignature3d (out[0..M)); flags[0] = isValid(), flags[1] = isReasonable(). The point functor is
 * homotopy_class_planner.h:72-75 (getCplxFromVertexPosePtr), restated here because that header pulls in boost::graph. */
int32_t teb_ref_h_signature(const TebParams* p, const double* rec, int32_t n, const TebObstacle* obst, int32_t M, const double* verts,
                            int32_t use_timediffs, double* out, int32_t* flags) {
  Scene sc;
  fill_scene(sc, p, obst, M, verts, nullptr, 0);
  TimedElasticBand teb;
  load_band(teb, rec, n);
  auto cplx = [](const VertexPose* pose) { return std::complex<long double>(pose->x(), pose->y()); };
  if (sc.cfg.obstacles.include_dynamic_obstacles) {
    HSignature3d H(sc.cfg);
    if (use_timediffs) H.calculateHSignature(teb.poses().begin(), teb.poses().end(), cplx, &sc.obstacles, teb.timediffs().begin(), teb.timediffs().end());
    else H.calculateHSignature(teb.poses().begin(), teb.poses().end(), cplx, &sc.obstacles, boost::none, boost::none);
    for (int m = 0; m < M && m < (int)H.values().size(); ++m) out[m] = H.values()[m];
    if (flags) { flags[0] = H.isValid(); flags[1] = H.isReasonable(); }
  } else {
    HSignature H(sc.cfg);
    H.calculateHSignature(teb.poses().begin(), teb.poses().end(), cplx, &sc.obstacles);
    out[0] = (double)H.value().real();
    out[1] = (double)H.value().imag();
    if (flags) { flags[0] = H.isValid(); flags[1] = H.isReasonable(); }
  }
  return 0;
}


namespace {
int put_band(const TimedElasticBand& teb, double* out, int cap) {
  const int n = teb.sizePoses();
  if (4 * n > cap) return -1;
  store_band(teb, out, n);
  return 4 * n;
}
void twist_from(const double* v, geometry_msgs::Twist& t) { t.linear.x = v[0]; t.linear.y = v[1]; t.angular.z = v[2]; }
}  // namespace

/* ---- band operations of the host-side API, one entry point for the pin test (tests/test_host_pin.py compares this
 * library's answers with the drop-in layer's, operation by operation):
 *   op 1  initTrajectoryToGoal(start, goal, diststep, max_vel_x, min_samples, backwards)   args: s[3] g[3] diststep vx min_samples backwards
 *   op 2  initTrajectoryToGoal(plan, max_vel_x, max_vel_theta, estimate_orient, min_samples, backwards)   args: vx vth est min_samples backwards np plan[np][3]
 *   op 3  initTrajectoryToGoal(2-D path, ...) of timed_elastic_band.hpp   args: vx vth accx acct so go (NaN = none) min_samples backwards np pts[np][2]
 *   op 4  updateAndPruneTEB   rec in; args: s[3] g[3] min_samples
 *   op 5  findClosestTrajectoryPose(point, &dist, begin_idx)   args: px py begin -> out: idx dist
 *   op 6  getSumOfAllTimeDiffs, getAccumulatedDistance, getSumOfTimeDiffsUpToIdx(args[0])
 *   op 7  isTrajectoryInsideRegion(radius, max_dist_behind_robot, skip_poses) -> out[0]
 *   op 8  getVelocityCommand(look_ahead)   args: look_ahead max_vel_y -> out: ok vx vy omega
 *   op 9  getVelocityProfile   args: max_vel_y vs[4] vg[4] -> out: (n + 1) x (vx, vy, omega)
 *   op 10 getFullTrajectory    args: max_vel_y vs[4] vg[4] -> out: n x (x, y, yaw, vx, vy, omega, t)
 *   op 11 autoResize(dt_ref, dt_hysteresis, min_samples, max_samples, fast_mode)
 * Bands go in and out as (x, y, theta, dt) records; the return value is the number of output doubles (< 0: capacity). */


/* op 12: isTrajectoryFeasible against a costmap stand-in that RECORDS every footprint query and reports a collision
 * (-1) inside any of the given discs. args: inscribed_radius circumscribed_radius look_ahead_idx lookahead_distance
 * min_resolution_collision_check_angular n_discs (x y r)* -> out: feasible, number of queries, the queried (x, y, theta). */
struct RecordingCostmap : public base_local_planner::CostmapModel {
  std::vector<double> q;
  const double* discs = nullptr;
  int nd = 0;
  double footprintCost(double x, double y, double th, const std::vector<geometry_msgs::Point>&, double = 0.0, double = 0.0) {
    q.push_back(x); q.push_back(y); q.push_back(th);
    for (int k = 0; k < nd; ++k)
      if (std::hypot(x - discs[3 * k], y - discs[3 * k + 1]) <= discs[3 * k + 2]) return -1;
    return 0;
  }
};

int32_t teb_ref_band_op(int32_t op, const double* rec, int32_t n, const double* a, int32_t na, double* out, int32_t cap) {
  (void)na;
  TebConfig cfg;
  TimedElasticBand teb;
  if (rec && n > 0) load_band(teb, rec, n);
  switch (op) {
    case 1:
      teb.initTrajectoryToGoal(PoseSE2(a[0], a[1], a[2]), PoseSE2(a[3], a[4], a[5]), a[6], a[7], (int)a[8], a[9] != 0);
      return put_band(teb, out, cap);
    case 2: {
      std::vector<geometry_msgs::PoseStamped> plan((size_t)a[5]);
      for (size_t i = 0; i < plan.size(); ++i) {
        plan[i].pose.position.x = a[6 + 3 * i]; plan[i].pose.position.y = a[7 + 3 * i];
        plan[i].pose.orientation = tf::createQuaternionMsgFromYaw(a[8 + 3 * i]);
      }
      teb.initTrajectoryToGoal(plan, a[0], a[1], a[2] != 0, (int)a[3], a[4] != 0);
      return put_band(teb, out, cap);
    }
    case 3: {
      std::vector<Eigen::Vector2d> path((size_t)a[8]);
      for (size_t i = 0; i < path.size(); ++i) path[i] = Eigen::Vector2d(a[9 + 2 * i], a[10 + 2 * i]);
      auto opt = [](double v) { return std::isnan(v) ? boost::optional<double>() : boost::optional<double>(v); };
      teb.initTrajectoryToGoal(path.begin(), path.end(), [](const Eigen::Vector2d& p) -> const Eigen::Vector2d& { return p; }, a[0], a[1],
                               opt(a[2]), opt(a[3]), opt(a[4]), opt(a[5]), (int)a[6], a[7] != 0);
      return put_band(teb, out, cap);
    }
    case 4: {
      PoseSE2 s(a[0], a[1], a[2]), g(a[3], a[4], a[5]);
      teb.updateAndPruneTEB(s, g, (int)a[6]);
      return put_band(teb, out, cap);
    }
    case 5: {
      double dist = -1;
      out[0] = teb.findClosestTrajectoryPose(Eigen::Vector2d(a[0], a[1]), &dist, (int)a[2]);
      out[1] = dist;
      return 2;
    }
    case 6:
      out[0] = teb.getSumOfAllTimeDiffs(); out[1] = teb.getAccumulatedDistance(); out[2] = teb.getSumOfTimeDiffsUpToIdx((int)a[0]);
      return 3;
    case 7:
      out[0] = teb.isTrajectoryInsideRegion(a[0], a[1], (int)a[2]);
      return 1;
    case 11:
      teb.autoResize(a[0], a[1], (int)a[2], (int)a[3], a[4] != 0);
      return put_band(teb, out, cap);
    default: break;
  }
  /* planner-level operations */
  cfg.robot.max_vel_y = (op == 8) ? a[1] : a[0];
  ObstContainer obst;
  RefPlanner pl(cfg, &obst, nullptr);
  load_band(pl.teb(), rec, n);
  if (op == 13) { /* plan() sequences with the optimisation switched off (optimization_activate = false): cold start,
                     warm start (updateAndPruneTEB) and re-initialisation after a goal jump (optimal_planner.cpp:233-321).
                     args: kind (0 poses / 1 plans) ncalls reinit_dist reinit_ang max_vel_x max_vel_theta min_samples
                     backwards overwrite_orientation, then per call s[3] g[3] or np pts[np][3]; out: per call n, band */
    TebConfig cfg13;
    cfg13.optim.optimization_activate = false;
    cfg13.trajectory.force_reinit_new_goal_dist = a[2];
    cfg13.trajectory.force_reinit_new_goal_angular = a[3];
    cfg13.robot.max_vel_x = a[4];
    cfg13.robot.max_vel_theta = a[5];
    cfg13.trajectory.min_samples = (int)a[6];
    cfg13.trajectory.allow_init_with_backwards_motion = a[7] != 0;
    cfg13.trajectory.global_plan_overwrite_orientation = a[8] != 0;
    ObstContainer obst13;
    RefPlanner pl13(cfg13, &obst13, nullptr);
    int r = 9, w = 0;
    for (int c = 0; c < (int)a[1]; ++c) {
      if (a[0] == 0) {
        pl13.plan(PoseSE2(a[r], a[r + 1], a[r + 2]), PoseSE2(a[r + 3], a[r + 4], a[r + 5]), nullptr, false);
        r += 6;
      } else {
        const int np = (int)a[r++];
        std::vector<geometry_msgs::PoseStamped> plan((size_t)np);
        for (int i = 0; i < np; ++i, r += 3) {
          plan[i].pose.position.x = a[r]; plan[i].pose.position.y = a[r + 1];
          plan[i].pose.orientation = tf::createQuaternionMsgFromYaw(a[r + 2]);
        }
        pl13.plan(plan, nullptr, false);
      }
      const int nn = pl13.teb().sizePoses();
      if (w + 1 + 4 * nn > cap) return -1;
      out[w++] = nn;
      const int k = put_band(pl13.teb(), out + w, cap - w);
      w += k;
    }
    return w;
  }
  if (op == 12) {
    RecordingCostmap cm;
    cm.discs = a + 6; cm.nd = (int)a[5];
    TebConfig cfg12;
    cfg12.trajectory.min_resolution_collision_check_angular = a[4];
    ObstContainer obst12;
    RefPlanner pl12(cfg12, &obst12, nullptr);
    load_band(pl12.teb(), rec, n);
    std::vector<geometry_msgs::Point> footprint;
    out[0] = pl12.isTrajectoryFeasible(&cm, footprint, a[0], a[1], (int)a[2], a[3]);
    out[1] = (double)(cm.q.size() / 3);
    if (2 + (int)cm.q.size() > cap) return -1;
    for (size_t k = 0; k < cm.q.size(); ++k) out[2 + k] = cm.q[k];
    return 2 + (int)cm.q.size();
  }
  if (op == 8) {
    double vx = 0, vy = 0, om = 0;
    out[0] = pl.getVelocityCommand(vx, vy, om, (int)a[0]);
    out[1] = vx; out[2] = vy; out[3] = om;
    return 4;
  }
  geometry_msgs::Twist ts, tg;
  twist_from(a + 1, ts); twist_from(a + 5, tg);
  if (a[4] != 0) pl.setVelocityStart(ts);
  if (a[8] != 0) pl.setVelocityGoal(tg); else pl.setVelocityGoalFree();
  if (op == 9) {
    std::vector<geometry_msgs::Twist> prof;
    pl.getVelocityProfile(prof);
    if (3 * (int)prof.size() > cap) return -1;
    for (size_t i = 0; i < prof.size(); ++i) { out[3 * i] = prof[i].linear.x; out[3 * i + 1] = prof[i].linear.y; out[3 * i + 2] = prof[i].angular.z; }
    return 3 * (int)prof.size();
  }
  if (op == 10) {
    std::vector<TrajectoryPointMsg> tr;
    pl.getFullTrajectory(tr);
    if (7 * (int)tr.size() > cap) return -1;
    for (size_t i = 0; i < tr.size(); ++i) {
      double* o = out + 7 * i;
      o[0] = tr[i].pose.position.x; o[1] = tr[i].pose.position.y; o[2] = tf::getYaw(tr[i].pose.orientation);
      o[3] = tr[i].velocity.linear.x; o[4] = tr[i].velocity.linear.y; o[5] = tr[i].velocity.angular.z; o[6] = tr[i].time_from_start.toSec();
    }
    return 7 * (int)tr.size();
  }
  return -2;
}

/* HomotopyClassPlanner::exploreEquivalenceClassesAndInitTebs (homotopy_class_planner.cpp:339-365) of the reference -
 * graph_search.cpp's lrKeyPointGraph (hcp[5] == 0) or ProbRoadmapGraph (!= 0), DepthFirst, addAndInitNewTeb with the
 * path variant of initTrajectoryToGoal, H-signature filtering - on a fresh planner, `cycles` times with clearPlanner() in
 * between (the roadmap's random generator lives on, as it does between planning cycles).
 * hcp = {max_number_classes, obstacle_heading_threshold, roadmap_graph_area_width, roadmap_graph_area_length_scale,
 *        roadmap_graph_no_samples, simple_exploration}. counts[c] = candidates of cycle c; then, per candidate in order,
 * its number of poses; out = the candidates' (x, y, theta, dt) records back to back. Returns the doubles written. */
int32_t teb_ref_hcp_explore(const TebParams* p, const double* hcp, const double* start3, const double* goal3, const TebObstacle* obst,
                            int32_t M, const double* verts, int32_t cycles, double* out, int32_t cap, int32_t* counts, int32_t counts_cap) {
  Scene sc;
  fill_scene(sc, p, obst, M, verts, nullptr, 0);
  sc.cfg.hcp.max_number_classes = (int)hcp[0];
  sc.cfg.hcp.obstacle_heading_threshold = hcp[1];
  sc.cfg.hcp.roadmap_graph_area_width = hcp[2];
  sc.cfg.hcp.roadmap_graph_area_length_scale = hcp[3];
  sc.cfg.hcp.roadmap_graph_no_samples = (int)hcp[4];
  sc.cfg.hcp.simple_exploration = hcp[5] == 0;      /* true: lrKeyPointGraph, false: ProbRoadmapGraph (homotopy_class_planner.cpp:84-87) */
  sc.cfg.hcp.enable_multithreading = false;
  HomotopyClassPlanner hcpl(sc.cfg, &sc.obstacles, TebVisualizationPtr(), nullptr);
  const PoseSE2 start(start3[0], start3[1], start3[2]), goal(goal3[0], goal3[1], goal3[2]);
  int w = 0, ci = 0;
  for (int c = 0; c < cycles; ++c) {
    if (c > 0) hcpl.clearPlanner();
    hcpl.exploreEquivalenceClassesAndInitTebs(start, goal, sc.cfg.obstacles.min_obstacle_dist, nullptr, false);
    const TebOptPlannerContainer& tebs = hcpl.getTrajectoryContainer();
    if (c >= counts_cap) return -1;
    counts[c] = (int32_t)tebs.size();
    ci = ci < cycles ? cycles : ci;
    for (size_t k = 0; k < tebs.size(); ++k) {
      const TimedElasticBand& teb = tebs[k]->teb();
      const int n = teb.sizePoses();
      if (ci >= counts_cap || w + 4 * n > cap) return -1;
      counts[ci++] = n;
      store_band(teb, out + w, n);
      w += 4 * n;
    }
  }
  return w;
}

/* `cycles` calls of HomotopyClassPlanner::plan(start_c, goal) (homotopy_class_planner.cpp:107-125: updateAllTEBs,
 * exploreEquivalenceClassesAndInitTebs incl. renewAndAnalyzeOldTebs / deletePlansDetouringBackwards, optimizeAllTEBs,
 * selectBestTeb) on ONE planner, the way consecutive control cycles use it. starts = [cycles][3]. hcp as in
 * teb_ref_hcp_explore. Per cycle: counts[3 c] = candidates, counts[3 c + 1] = index of the best one (-1: none),
 * counts[3 c + 2] = plan()'s return value; then (from counts[3 cycles] on) the pose count of every candidate of every
 * cycle; out = per candidate its cost (one double) followed by its (x, y, theta, dt) records. */
int32_t teb_ref_hcp_plan(const TebParams* p, const double* hcp, const double* starts, const double* goal3, const TebObstacle* obst,
                         int32_t M, const double* verts, int32_t cycles, double* out, int32_t cap, int32_t* counts, int32_t counts_cap) {
  Scene sc;
  fill_scene(sc, p, obst, M, verts, nullptr, 0);
  sc.cfg.hcp.max_number_classes = (int)hcp[0];
  sc.cfg.hcp.obstacle_heading_threshold = hcp[1];
  sc.cfg.hcp.roadmap_graph_area_width = hcp[2];
  sc.cfg.hcp.roadmap_graph_area_length_scale = hcp[3];
  sc.cfg.hcp.roadmap_graph_no_samples = (int)hcp[4];
  sc.cfg.hcp.simple_exploration = hcp[5] == 0;
  sc.cfg.hcp.enable_multithreading = false;
  HomotopyClassPlanner hcpl(sc.cfg, &sc.obstacles, TebVisualizationPtr(), nullptr);
  const PoseSE2 goal(goal3[0], goal3[1], goal3[2]);
  int w = 0, ci = 3 * cycles;
  if (ci > counts_cap) return -1;
  for (int c = 0; c < cycles; ++c) {
    const PoseSE2 start(starts[3 * c], starts[3 * c + 1], starts[3 * c + 2]);
    const bool ok = hcpl.plan(start, goal, nullptr, false);
    const TebOptPlannerContainer& tebs = hcpl.getTrajectoryContainer();
    counts[3 * c] = (int32_t)tebs.size();
    counts[3 * c + 1] = -1;
    counts[3 * c + 2] = ok;
    TebOptimalPlannerPtr best = hcpl.bestTeb();
    for (size_t k = 0; k < tebs.size(); ++k) {
      if (tebs[k] == best) counts[3 * c + 1] = (int32_t)k;
      const TimedElasticBand& teb = tebs[k]->teb();
      const int n = teb.sizePoses();
      if (ci >= counts_cap || w + 1 + 4 * n > cap) return -1;
      counts[ci++] = n;
      out[w++] = tebs[k]->getCurrentCost();
      store_band(teb, out + w, n);
      w += 4 * n;
    }
  }
  return w;
}

}  /* extern "C" */
