/*
 * teb_config.h — ROS-free mirror of teb_local_planner::TebConfig (include/teb_local_planner/teb_config.h:62-430).
 * Same nested groups and field names (so code written against the reference's `cfg.optim.weight_obstacle` style
 * compiles unchanged); constructor defaults are the reference's constructor defaults (:245-390). The fields the
 * reference leaves uninitialised in its constructor (hcp.max_number_plans_in_current_class,
 * recovery.divergence_detection_*) take the dynamic_reconfigure defaults (cfg/TebLocalPlannerReconfigure.cfg).
 * rosparam / dynamic_reconfigure loading and the config mutex are out of scope (ROS is absent).
 * toParams() produces the POD block that crosses the C-ABI (include/teb_b200.h TebParams).
 */
#ifndef TEB_B200_TEB_CONFIG_H_
#define TEB_B200_TEB_CONFIG_H_

#include <memory>
#include <string>

#include "teb_b200.h"
#include "teb_local_planner/obstacles.h"

#define USE_ANALYTIC_JACOBI /* teb_config.h:52 (informational: every Jacobian on the device is closed form) */

namespace teb_local_planner {

class TebConfig {
 public:
  std::string odom_topic;
  std::string map_frame;
  RobotFootprintModelPtr robot_model;

  struct Trajectory {
    double teb_autosize = true;
    double dt_ref = 0.3;
    double dt_hysteresis = 0.1;
    int min_samples = 3;
    int max_samples = 500;
    bool global_plan_overwrite_orientation = true;
    bool allow_init_with_backwards_motion = false;
    double global_plan_viapoint_sep = -1;
    bool via_points_ordered = false;
    double max_global_plan_lookahead_dist = 1;
    double global_plan_prune_distance = 1;
    bool exact_arc_length = false;
    double force_reinit_new_goal_dist = 1;
    double force_reinit_new_goal_angular = 0.5 * M_PI;
    int feasibility_check_no_poses = 5;
    double feasibility_check_lookahead_distance = -1;
    bool publish_feedback = false;
    double min_resolution_collision_check_angular = M_PI;
    int control_look_ahead_poses = 1;
    int prevent_look_ahead_poses_near_goal = 0;
  } trajectory;

  struct Robot {
    double max_vel_x = 0.4;
    double max_vel_x_backwards = 0.2;
    double max_vel_y = 0.0;
    double max_vel_trans = 0.0;
    double max_vel_theta = 0.3;
    double acc_lim_x = 0.5;
    double acc_lim_y = 0.5;
    double acc_lim_theta = 0.5;
    double min_turning_radius = 0;
    double wheelbase = 1.0;
    bool cmd_angle_instead_rotvel = false;
    bool is_footprint_dynamic = false;
    bool use_proportional_saturation = false;
    double transform_tolerance = 0.5;
  } robot;

  struct GoalTolerance {
    double yaw_goal_tolerance = 0.2;
    double xy_goal_tolerance = 0.2;
    bool free_goal_vel = false;
    double trans_stopped_vel = 0.1;
    double theta_stopped_vel = 0.1;
    bool complete_global_plan = true;
  } goal_tolerance;

  struct Obstacles {
    double min_obstacle_dist = 0.5;
    double inflation_dist = 0.6;
    double dynamic_obstacle_inflation_dist = 0.6;
    bool include_dynamic_obstacles = true;
    bool include_costmap_obstacles = true;
    double costmap_obstacles_behind_robot_dist = 1.5;
    int obstacle_poses_affected = 25;
    bool legacy_obstacle_association = false;
    double obstacle_association_force_inclusion_factor = 1.5;
    double obstacle_association_cutoff_factor = 5;
    std::string costmap_converter_plugin = "";
    bool costmap_converter_spin_thread = true;
    int costmap_converter_rate = 5;
    double obstacle_proximity_ratio_max_vel = 1;
    double obstacle_proximity_lower_bound = 0;
    double obstacle_proximity_upper_bound = 0.5;
  } obstacles;

  struct Optimization {
    int no_inner_iterations = 5;
    int no_outer_iterations = 4;
    bool optimization_activate = true;
    bool optimization_verbose = false;
    double penalty_epsilon = 0.05;
    double weight_max_vel_x = 2;
    double weight_max_vel_y = 2;
    double weight_max_vel_theta = 1;
    double weight_acc_lim_x = 1;
    double weight_acc_lim_y = 1;
    double weight_acc_lim_theta = 1;
    double weight_kinematics_nh = 1000;
    double weight_kinematics_forward_drive = 1;
    double weight_kinematics_turning_radius = 1;
    double weight_optimaltime = 1;
    double weight_shortest_path = 0;
    double weight_obstacle = 50;
    double weight_inflation = 0.1;
    double weight_dynamic_obstacle = 50;
    double weight_dynamic_obstacle_inflation = 0.1;
    double weight_velocity_obstacle_ratio = 0;
    double weight_viapoint = 1;
    double weight_prefer_rotdir = 50;
    double weight_adapt_factor = 2.0;
    double obstacle_cost_exponent = 1.0;
  } optim;

  struct HomotopyClasses {
    bool enable_homotopy_class_planning = true;
    bool enable_multithreading = true;
    bool simple_exploration = false;
    int max_number_classes = 5;
    int max_number_plans_in_current_class = 1;
    double selection_cost_hysteresis = 1.0;
    double selection_prefer_initial_plan = 0.95;
    double selection_obst_cost_scale = 100.0;
    double selection_viapoint_cost_scale = 1.0;
    bool selection_alternative_time_cost = false;
    double selection_dropping_probability = 0.0;
    double switching_blocking_period = 0.0;
    int roadmap_graph_no_samples = 15;
    double roadmap_graph_area_width = 6;
    double roadmap_graph_area_length_scale = 1.0;
    double h_signature_prescaler = 1;
    double h_signature_threshold = 0.1;
    double obstacle_keypoint_offset = 0.1;
    double obstacle_heading_threshold = 0.45;
    bool viapoints_all_candidates = true;
    bool visualize_hc_graph = false;
    double visualize_with_time_as_z_axis_scale = 0.0;
    bool delete_detours_backwards = true;
    double detours_orientation_tolerance = M_PI / 2.0;
    double length_start_orientation_vector = 0.4;
    double max_ratio_detours_duration_best_duration = 3.0;
  } hcp;

  struct Recovery {
    bool shrink_horizon_backup = true;
    double shrink_horizon_min_duration = 10;
    bool oscillation_recovery = true;
    double oscillation_v_eps = 0.1;
    double oscillation_omega_eps = 0.1;
    double oscillation_recovery_min_duration = 10;
    double oscillation_filter_duration = 10;
    bool divergence_detection_enable = false;
    int divergence_detection_max_chi_squared = 10;
  } recovery;

  /* defaults = the reference's TebConfig() constructor (teb_config.h:245-390), written as member initialisers */
  TebConfig() : odom_topic("odom"), map_frame("odom"), robot_model(std::make_shared<PointRobotFootprint>()) {}

  /* POD block for the device (every field the kernels read) */
  TebParams toParams() const {
    TebParams p;
    tebgpu_default_params(&p);
    p.dt_ref = trajectory.dt_ref; p.dt_hysteresis = trajectory.dt_hysteresis;
    p.force_reinit_new_goal_dist = trajectory.force_reinit_new_goal_dist;
    p.force_reinit_new_goal_angular = trajectory.force_reinit_new_goal_angular;
    p.teb_autosize = trajectory.teb_autosize != 0; p.min_samples = trajectory.min_samples; p.max_samples = trajectory.max_samples;
    p.exact_arc_length = trajectory.exact_arc_length; p.via_points_ordered = trajectory.via_points_ordered;
    p.allow_init_with_backwards_motion = trajectory.allow_init_with_backwards_motion;
    p.global_plan_overwrite_orientation = trajectory.global_plan_overwrite_orientation;
    p.max_vel_x = robot.max_vel_x; p.max_vel_x_backwards = robot.max_vel_x_backwards; p.max_vel_y = robot.max_vel_y;
    p.max_vel_trans = robot.max_vel_trans; p.max_vel_theta = robot.max_vel_theta; p.acc_lim_x = robot.acc_lim_x;
    p.acc_lim_y = robot.acc_lim_y; p.acc_lim_theta = robot.acc_lim_theta; p.min_turning_radius = robot.min_turning_radius;
    if (robot_model) robot_model->fillParams(p);
    p.min_obstacle_dist = obstacles.min_obstacle_dist; p.inflation_dist = obstacles.inflation_dist;
    p.dynamic_obstacle_inflation_dist = obstacles.dynamic_obstacle_inflation_dist;
    p.obstacle_association_force_inclusion_factor = obstacles.obstacle_association_force_inclusion_factor;
    p.obstacle_association_cutoff_factor = obstacles.obstacle_association_cutoff_factor;
    p.obstacle_proximity_ratio_max_vel = obstacles.obstacle_proximity_ratio_max_vel;
    p.obstacle_proximity_lower_bound = obstacles.obstacle_proximity_lower_bound;
    p.obstacle_proximity_upper_bound = obstacles.obstacle_proximity_upper_bound;
    p.include_dynamic_obstacles = obstacles.include_dynamic_obstacles;
    p.legacy_obstacle_association = obstacles.legacy_obstacle_association;
    p.obstacle_poses_affected = obstacles.obstacle_poses_affected;
    p.penalty_epsilon = optim.penalty_epsilon;
    p.weight_max_vel_x = optim.weight_max_vel_x; p.weight_max_vel_y = optim.weight_max_vel_y;
    p.weight_max_vel_theta = optim.weight_max_vel_theta; p.weight_acc_lim_x = optim.weight_acc_lim_x;
    p.weight_acc_lim_y = optim.weight_acc_lim_y; p.weight_acc_lim_theta = optim.weight_acc_lim_theta;
    p.weight_kinematics_nh = optim.weight_kinematics_nh;
    p.weight_kinematics_forward_drive = optim.weight_kinematics_forward_drive;
    p.weight_kinematics_turning_radius = optim.weight_kinematics_turning_radius;
    p.weight_optimaltime = optim.weight_optimaltime; p.weight_shortest_path = optim.weight_shortest_path;
    p.weight_obstacle = optim.weight_obstacle; p.weight_inflation = optim.weight_inflation;
    p.weight_dynamic_obstacle = optim.weight_dynamic_obstacle;
    p.weight_dynamic_obstacle_inflation = optim.weight_dynamic_obstacle_inflation;
    p.weight_velocity_obstacle_ratio = optim.weight_velocity_obstacle_ratio;
    p.weight_viapoint = optim.weight_viapoint; p.weight_prefer_rotdir = optim.weight_prefer_rotdir;
    p.weight_adapt_factor = optim.weight_adapt_factor; p.obstacle_cost_exponent = optim.obstacle_cost_exponent;
    p.no_inner_iterations = optim.no_inner_iterations; p.no_outer_iterations = optim.no_outer_iterations;
    p.optimization_activate = optim.optimization_activate;
    p.selection_cost_hysteresis = hcp.selection_cost_hysteresis;
    p.selection_prefer_initial_plan = hcp.selection_prefer_initial_plan;
    p.selection_obst_cost_scale = hcp.selection_obst_cost_scale;
    p.selection_viapoint_cost_scale = hcp.selection_viapoint_cost_scale;
    p.selection_alternative_time_cost = hcp.selection_alternative_time_cost;
    p.enable_multithreading = hcp.enable_multithreading;
    p.h_signature_prescaler = hcp.h_signature_prescaler; p.h_signature_threshold = hcp.h_signature_threshold;
    p.divergence_detection_enable = recovery.divergence_detection_enable;
    p.divergence_detection_max_chi_squared = recovery.divergence_detection_max_chi_squared;
    return p;
  }
};

}  // namespace teb_local_planner
#endif
