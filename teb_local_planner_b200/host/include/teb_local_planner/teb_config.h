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
    double teb_autosize;
    double dt_ref;
    double dt_hysteresis;
    int min_samples;
    int max_samples;
    bool global_plan_overwrite_orientation;
    bool allow_init_with_backwards_motion;
    double global_plan_viapoint_sep;
    bool via_points_ordered;
    double max_global_plan_lookahead_dist;
    double global_plan_prune_distance;
    bool exact_arc_length;
    double force_reinit_new_goal_dist;
    double force_reinit_new_goal_angular;
    int feasibility_check_no_poses;
    double feasibility_check_lookahead_distance;
    bool publish_feedback;
    double min_resolution_collision_check_angular;
    int control_look_ahead_poses;
    int prevent_look_ahead_poses_near_goal;
  } trajectory;

  struct Robot {
    double max_vel_x;
    double max_vel_x_backwards;
    double max_vel_y;
    double max_vel_trans;
    double max_vel_theta;
    double acc_lim_x;
    double acc_lim_y;
    double acc_lim_theta;
    double min_turning_radius;
    double wheelbase;
    bool cmd_angle_instead_rotvel;
    bool is_footprint_dynamic;
    bool use_proportional_saturation;
    double transform_tolerance = 0.5;
  } robot;

  struct GoalTolerance {
    double yaw_goal_tolerance;
    double xy_goal_tolerance;
    bool free_goal_vel;
    double trans_stopped_vel;
    double theta_stopped_vel;
    bool complete_global_plan;
  } goal_tolerance;

  struct Obstacles {
    double min_obstacle_dist;
    double inflation_dist;
    double dynamic_obstacle_inflation_dist;
    bool include_dynamic_obstacles;
    bool include_costmap_obstacles;
    double costmap_obstacles_behind_robot_dist;
    int obstacle_poses_affected;
    bool legacy_obstacle_association;
    double obstacle_association_force_inclusion_factor;
    double obstacle_association_cutoff_factor;
    std::string costmap_converter_plugin;
    bool costmap_converter_spin_thread;
    int costmap_converter_rate;
    double obstacle_proximity_ratio_max_vel;
    double obstacle_proximity_lower_bound;
    double obstacle_proximity_upper_bound;
  } obstacles;

  struct Optimization {
    int no_inner_iterations;
    int no_outer_iterations;
    bool optimization_activate;
    bool optimization_verbose;
    double penalty_epsilon;
    double weight_max_vel_x;
    double weight_max_vel_y;
    double weight_max_vel_theta;
    double weight_acc_lim_x;
    double weight_acc_lim_y;
    double weight_acc_lim_theta;
    double weight_kinematics_nh;
    double weight_kinematics_forward_drive;
    double weight_kinematics_turning_radius;
    double weight_optimaltime;
    double weight_shortest_path;
    double weight_obstacle;
    double weight_inflation;
    double weight_dynamic_obstacle;
    double weight_dynamic_obstacle_inflation;
    double weight_velocity_obstacle_ratio;
    double weight_viapoint;
    double weight_prefer_rotdir;
    double weight_adapt_factor;
    double obstacle_cost_exponent;
  } optim;

  struct HomotopyClasses {
    bool enable_homotopy_class_planning;
    bool enable_multithreading;
    bool simple_exploration;
    int max_number_classes;
    int max_number_plans_in_current_class;
    double selection_cost_hysteresis;
    double selection_prefer_initial_plan;
    double selection_obst_cost_scale;
    double selection_viapoint_cost_scale;
    bool selection_alternative_time_cost;
    double selection_dropping_probability;
    double switching_blocking_period;
    int roadmap_graph_no_samples;
    double roadmap_graph_area_width;
    double roadmap_graph_area_length_scale;
    double h_signature_prescaler;
    double h_signature_threshold;
    double obstacle_keypoint_offset;
    double obstacle_heading_threshold;
    bool viapoints_all_candidates;
    bool visualize_hc_graph;
    double visualize_with_time_as_z_axis_scale;
    bool delete_detours_backwards;
    double detours_orientation_tolerance;
    double length_start_orientation_vector;
    double max_ratio_detours_duration_best_duration;
  } hcp;

  struct Recovery {
    bool shrink_horizon_backup;
    double shrink_horizon_min_duration;
    bool oscillation_recovery;
    double oscillation_v_eps;
    double oscillation_omega_eps;
    double oscillation_recovery_min_duration;
    double oscillation_filter_duration;
    bool divergence_detection_enable;
    int divergence_detection_max_chi_squared;
  } recovery;

  TebConfig() {
    odom_topic = "odom";
    map_frame = "odom";
    robot_model = std::make_shared<PointRobotFootprint>();
    trajectory.teb_autosize = true;
    trajectory.dt_ref = 0.3;
    trajectory.dt_hysteresis = 0.1;
    trajectory.min_samples = 3;
    trajectory.max_samples = 500;
    trajectory.global_plan_overwrite_orientation = true;
    trajectory.allow_init_with_backwards_motion = false;
    trajectory.global_plan_viapoint_sep = -1;
    trajectory.via_points_ordered = false;
    trajectory.max_global_plan_lookahead_dist = 1;
    trajectory.global_plan_prune_distance = 1;
    trajectory.exact_arc_length = false;
    trajectory.force_reinit_new_goal_dist = 1;
    trajectory.force_reinit_new_goal_angular = 0.5 * M_PI;
    trajectory.feasibility_check_no_poses = 5;
    trajectory.feasibility_check_lookahead_distance = -1;
    trajectory.publish_feedback = false;
    trajectory.min_resolution_collision_check_angular = M_PI;
    trajectory.control_look_ahead_poses = 1;
    trajectory.prevent_look_ahead_poses_near_goal = 0;
    robot.max_vel_x = 0.4;
    robot.max_vel_x_backwards = 0.2;
    robot.max_vel_y = 0.0;
    robot.max_vel_trans = 0.0;
    robot.max_vel_theta = 0.3;
    robot.acc_lim_x = 0.5;
    robot.acc_lim_y = 0.5;
    robot.acc_lim_theta = 0.5;
    robot.min_turning_radius = 0;
    robot.wheelbase = 1.0;
    robot.cmd_angle_instead_rotvel = false;
    robot.is_footprint_dynamic = false;
    robot.use_proportional_saturation = false;
    goal_tolerance.xy_goal_tolerance = 0.2;
    goal_tolerance.yaw_goal_tolerance = 0.2;
    goal_tolerance.free_goal_vel = false;
    goal_tolerance.trans_stopped_vel = 0.1;
    goal_tolerance.theta_stopped_vel = 0.1;
    goal_tolerance.complete_global_plan = true;
    obstacles.min_obstacle_dist = 0.5;
    obstacles.inflation_dist = 0.6;
    obstacles.dynamic_obstacle_inflation_dist = 0.6;
    obstacles.include_dynamic_obstacles = true;
    obstacles.include_costmap_obstacles = true;
    obstacles.costmap_obstacles_behind_robot_dist = 1.5;
    obstacles.obstacle_poses_affected = 25;
    obstacles.legacy_obstacle_association = false;
    obstacles.obstacle_association_force_inclusion_factor = 1.5;
    obstacles.obstacle_association_cutoff_factor = 5;
    obstacles.costmap_converter_plugin = "";
    obstacles.costmap_converter_spin_thread = true;
    obstacles.costmap_converter_rate = 5;
    obstacles.obstacle_proximity_ratio_max_vel = 1;
    obstacles.obstacle_proximity_lower_bound = 0;
    obstacles.obstacle_proximity_upper_bound = 0.5;
    optim.no_inner_iterations = 5;
    optim.no_outer_iterations = 4;
    optim.optimization_activate = true;
    optim.optimization_verbose = false;
    optim.penalty_epsilon = 0.05;
    optim.weight_max_vel_x = 2;
    optim.weight_max_vel_y = 2;
    optim.weight_max_vel_theta = 1;
    optim.weight_acc_lim_x = 1;
    optim.weight_acc_lim_y = 1;
    optim.weight_acc_lim_theta = 1;
    optim.weight_kinematics_nh = 1000;
    optim.weight_kinematics_forward_drive = 1;
    optim.weight_kinematics_turning_radius = 1;
    optim.weight_optimaltime = 1;
    optim.weight_shortest_path = 0;
    optim.weight_obstacle = 50;
    optim.weight_inflation = 0.1;
    optim.weight_dynamic_obstacle = 50;
    optim.weight_dynamic_obstacle_inflation = 0.1;
    optim.weight_velocity_obstacle_ratio = 0;
    optim.weight_viapoint = 1;
    optim.weight_prefer_rotdir = 50;
    optim.weight_adapt_factor = 2.0;
    optim.obstacle_cost_exponent = 1.0;
    hcp.enable_homotopy_class_planning = true;
    hcp.enable_multithreading = true;
    hcp.simple_exploration = false;
    hcp.max_number_classes = 5;
    hcp.max_number_plans_in_current_class = 1;
    hcp.selection_cost_hysteresis = 1.0;
    hcp.selection_prefer_initial_plan = 0.95;
    hcp.selection_obst_cost_scale = 100.0;
    hcp.selection_viapoint_cost_scale = 1.0;
    hcp.selection_alternative_time_cost = false;
    hcp.selection_dropping_probability = 0.0;
    hcp.obstacle_keypoint_offset = 0.1;
    hcp.obstacle_heading_threshold = 0.45;
    hcp.roadmap_graph_no_samples = 15;
    hcp.roadmap_graph_area_width = 6;
    hcp.roadmap_graph_area_length_scale = 1.0;
    hcp.h_signature_prescaler = 1;
    hcp.h_signature_threshold = 0.1;
    hcp.switching_blocking_period = 0.0;
    hcp.viapoints_all_candidates = true;
    hcp.visualize_hc_graph = false;
    hcp.visualize_with_time_as_z_axis_scale = 0.0;
    hcp.delete_detours_backwards = true;
    hcp.detours_orientation_tolerance = M_PI / 2.0;
    hcp.length_start_orientation_vector = 0.4;
    hcp.max_ratio_detours_duration_best_duration = 3.0;
    recovery.shrink_horizon_backup = true;
    recovery.shrink_horizon_min_duration = 10;
    recovery.oscillation_recovery = true;
    recovery.oscillation_v_eps = 0.1;
    recovery.oscillation_omega_eps = 0.1;
    recovery.oscillation_recovery_min_duration = 10;
    recovery.oscillation_filter_duration = 10;
    recovery.divergence_detection_enable = false;
    recovery.divergence_detection_max_chi_squared = 10;
  }

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
    p.divergence_detection_enable = recovery.divergence_detection_enable;
    p.divergence_detection_max_chi_squared = recovery.divergence_detection_max_chi_squared;
    return p;
  }
};

}  // namespace teb_local_planner
#endif
