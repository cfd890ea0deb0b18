/* timed_elastic_band.cpp — see include/teb_local_planner/timed_elastic_band.h.
 * Behaviour follows src/timed_elastic_band.cpp of the reference; line references below point there. */
#include "teb_local_planner/timed_elastic_band.h"

#include <algorithm>
#include <cmath>

#include "../../csrc/teb_resize.h"

namespace teb_local_planner {

namespace {
/* estimateDeltaT (timed_elastic_band.cpp:52-65) */
double estimateDeltaT(const PoseSE2& start, const PoseSE2& end, double max_vel_x, double max_vel_theta) {
  double dt_constant_motion = 0.1;
  if (max_vel_x > 0) {
    double trans_dist = (end.position() - start.position()).norm();
    dt_constant_motion = trans_dist / max_vel_x;
  }
  if (max_vel_theta > 0) {
    double rot_dist = std::abs(g2o::normalize_theta(end.theta() - start.theta()));
    dt_constant_motion = std::max(dt_constant_motion, rot_dist / max_vel_theta);
  }
  return dt_constant_motion;
}
}  // namespace

/* :141-150 */
void TimedElasticBand::addPoseAndTimeDiff(const PoseSE2& pose, double dt) {
  if (sizePoses() != sizeTimeDiffs()) {
    addPose(pose, false);
    addTimeDiff(dt, false);
  }
}

void TimedElasticBand::setPoseVertexFixed(int index, bool status) {
  /* only the start and the goal pose are fixed on the device path; they always are (:330, :377) */
  (void)index; (void)status;
}

/* :227-286 — same routine as the device kernel */
void TimedElasticBand::autoResize(double dt_ref, double dt_hysteresis, int min_samples, int max_samples, bool fast_mode) {
  const int n = sizePoses();
  if (n < 2) return;
  const int n_cap = std::max(2 * n + 16, max_samples + 2);
  std::vector<double> rec((size_t)4 * n_cap, 0.0);
  toRecords(rec.data());
  int nn = teb_auto_resize_records(rec.data(), n, n_cap, dt_ref, dt_hysteresis, min_samples, max_samples, fast_mode ? 1 : 0);
  if (nn < 0) return;
  fromRecords(rec.data(), nn);
}

/* :325-387 */
bool TimedElasticBand::initTrajectoryToGoal(const PoseSE2& start, const PoseSE2& goal, double diststep, double max_vel_x,
                                            int min_samples, bool guess_backwards_motion) {
  if (!isInit()) {
    addPose(start);
    setPoseVertexFixed(0, true);
    double timestep = 0.1;
    if (diststep != 0) {
      Eigen::Vector2d point_to_goal = goal.position() - start.position();
      double dir_to_goal = std::atan2(point_to_goal[1], point_to_goal[0]);
      double dx = diststep * std::cos(dir_to_goal);
      double dy = diststep * std::sin(dir_to_goal);
      double orient_init = dir_to_goal;
      if (guess_backwards_motion && point_to_goal.dot(start.orientationUnitVec()) < 0)
        orient_init = g2o::normalize_theta(orient_init + M_PI);
      double dist_to_goal = point_to_goal.norm();
      double no_steps_d = dist_to_goal / std::abs(diststep);
      unsigned int no_steps = (unsigned int)std::floor(no_steps_d);
      if (max_vel_x > 0) timestep = diststep / max_vel_x;
      for (unsigned int i = 1; i <= no_steps; i++) {
        if (i == no_steps && no_steps_d == (float)no_steps) break;
        addPoseAndTimeDiff(start.x() + i * dx, start.y() + i * dy, orient_init, timestep);
      }
    }
    if (sizePoses() < min_samples - 1) {
      while (sizePoses() < min_samples - 1) {
        PoseSE2 intermediate_pose = PoseSE2::average(BackPose(), goal);
        if (max_vel_x > 0) timestep = (intermediate_pose.position() - BackPose().position()).norm() / max_vel_x;
        addPoseAndTimeDiff(intermediate_pose, timestep);
      }
    }
    if (max_vel_x > 0) timestep = (goal.position() - BackPose().position()).norm() / max_vel_x;
    addPoseAndTimeDiff(goal, timestep);
    setPoseVertexFixed(sizePoses() - 1, true);
  } else {
    return false;
  }
  return true;
}

/* :389-452 */
bool TimedElasticBand::initTrajectoryToGoal(const std::vector<geometry_msgs::PoseStamped>& plan, double max_vel_x,
                                            double max_vel_theta, bool estimate_orient, int min_samples,
                                            bool guess_backwards_motion) {
  if (!isInit()) {
    PoseSE2 start(plan.front().pose);
    PoseSE2 goal(plan.back().pose);
    addPose(start);
    setPoseVertexFixed(0, true);
    bool backwards = false;
    if (guess_backwards_motion && (goal.position() - start.position()).dot(start.orientationUnitVec()) < 0) backwards = true;
    for (int i = 1; i < (int)plan.size() - 1; ++i) {
      double yaw;
      if (estimate_orient) {
        double dx = plan[i + 1].pose.position.x - plan[i].pose.position.x;
        double dy = plan[i + 1].pose.position.y - plan[i].pose.position.y;
        yaw = std::atan2(dy, dx);
        if (backwards) yaw = g2o::normalize_theta(yaw + M_PI);
      } else {
        yaw = tf::getYaw(plan[i].pose.orientation);
      }
      PoseSE2 intermediate_pose(plan[i].pose.position.x, plan[i].pose.position.y, yaw);
      double dt = estimateDeltaT(BackPose(), intermediate_pose, max_vel_x, max_vel_theta);
      addPoseAndTimeDiff(intermediate_pose, dt);
    }
    if (sizePoses() < min_samples - 1) {
      while (sizePoses() < min_samples - 1) {
        PoseSE2 intermediate_pose = PoseSE2::average(BackPose(), goal);
        double dt = estimateDeltaT(BackPose(), intermediate_pose, max_vel_x, max_vel_theta);
        addPoseAndTimeDiff(intermediate_pose, dt);
      }
    }
    double dt = estimateDeltaT(BackPose(), goal, max_vel_x, max_vel_theta);
    addPoseAndTimeDiff(goal, dt);
    setPoseVertexFixed(sizePoses() - 1, true);
  } else {
    return false;
  }
  return true;
}

/* :455-478 */
int TimedElasticBand::findClosestTrajectoryPose(const Eigen::Vector2d& ref_point, double* distance, int begin_idx) const {
  int n = sizePoses();
  if (begin_idx < 0 || begin_idx >= n) return -1;
  double min_dist_sq = std::numeric_limits<double>::max();
  int min_idx = -1;
  for (int i = begin_idx; i < n; i++) {
    double dist_sq = (ref_point - Pose(i).position()).squaredNorm();
    if (dist_sq < min_dist_sq) {
      min_dist_sq = dist_sq;
      min_idx = i;
    }
  }
  if (distance) *distance = std::sqrt(min_dist_sq);
  return min_idx;
}

/* :555-597 */
void TimedElasticBand::updateAndPruneTEB(const PoseSE2* new_start, const PoseSE2* new_goal, int min_samples) {
  if (new_start && sizePoses() > 0) {
    double dist_cache = (new_start->position() - Pose(0).position()).norm();
    double dist;
    int lookahead = std::min<int>(sizePoses() - min_samples, 10);
    int nearest_idx = 0;
    for (int i = 1; i <= lookahead; ++i) {
      dist = (new_start->position() - Pose(i).position()).norm();
      if (dist < dist_cache) {
        dist_cache = dist;
        nearest_idx = i;
      } else {
        break;
      }
    }
    if (nearest_idx > 0) {
      deletePoses(1, nearest_idx);
      deleteTimeDiffs(1, nearest_idx);
    }
    Pose(0) = *new_start;
  }
  if (new_goal && sizePoses() > 0) BackPose() = *new_goal;
}

double TimedElasticBand::getSumOfAllTimeDiffs() const {
  double time = 0;
  for (double dt : timediff_vec_) time += dt;
  return time;
}
double TimedElasticBand::getSumOfTimeDiffsUpToIdx(int index) const {
  double time = 0;
  for (int i = 0; i < index; ++i) time += timediff_vec_.at(i);
  return time;
}
double TimedElasticBand::getAccumulatedDistance() const {
  double dist = 0;
  for (int i = 1; i < sizePoses(); ++i) dist += (Pose(i).position() - Pose(i - 1).position()).norm();
  return dist;
}

/* :600-631 */
bool TimedElasticBand::isTrajectoryInsideRegion(double radius, double max_dist_behind_robot, int skip_poses) {
  if (sizePoses() <= 0) return true;
  double radius_sq = radius * radius;
  double max_dist_behind_robot_sq = max_dist_behind_robot * max_dist_behind_robot;
  Eigen::Vector2d robot_orient = Pose(0).orientationUnitVec();
  for (int i = 1; i < sizePoses(); i = i + skip_poses + 1) {
    Eigen::Vector2d dist_vec = Pose(i).position() - Pose(0).position();
    double dist_sq = dist_vec.squaredNorm();
    if (dist_sq > radius_sq) return false;
    if (max_dist_behind_robot >= 0 && dist_vec.dot(robot_orient) < 0 && dist_sq > max_dist_behind_robot_sq) return false;
  }
  return true;
}

void TimedElasticBand::toRecords(double* rec) const {
  const int n = sizePoses();
  for (int i = 0; i < n; ++i) {
    rec[4 * i] = pose_vec_[i].x();
    rec[4 * i + 1] = pose_vec_[i].y();
    rec[4 * i + 2] = pose_vec_[i].theta();
    rec[4 * i + 3] = i < sizeTimeDiffs() ? timediff_vec_[i] : 0.0;
  }
}
void TimedElasticBand::fromRecords(const double* rec, int n) {
  pose_vec_.resize(n);
  timediff_vec_.resize(n > 0 ? n - 1 : 0);
  for (int i = 0; i < n; ++i) {
    pose_vec_[i] = PoseSE2(rec[4 * i], rec[4 * i + 1], rec[4 * i + 2]);
    if (i < n - 1) timediff_vec_[i] = rec[4 * i + 3];
  }
}

/* timed_elastic_band.hpp:46-185: poses on the path points, heading along the incoming segment, time per segment from a
 * constant-velocity / constant-acceleration estimate; samples are added towards the goal until min_samples is met */
bool TimedElasticBand::initTrajectoryToGoal(const std::vector<Eigen::Vector2d>& path, double max_vel_x, double max_vel_theta,
                                            const double* max_acc_x, const double* max_acc_theta, const double* start_orientation,
                                            const double* goal_orientation, int min_samples, bool guess_backwards_motion) {
  (void)max_vel_theta; (void)max_acc_theta;
  if (path.empty()) return false;
  if (isInit()) return false;
  const Eigen::Vector2d start_position = path.front(), goal_position = path.back();
  bool backwards = false;
  double start_orient;
  if (start_orientation) {
    start_orient = *start_orientation;
    if (guess_backwards_motion &&
        (goal_position - start_position).dot(Eigen::Vector2d(std::cos(start_orient), std::sin(start_orient))) < 0)
      backwards = true;
  } else {
    const Eigen::Vector2d start2goal = goal_position - start_position;
    start_orient = std::atan2(start2goal[1], start2goal[0]);
  }
  const double goal_orient = goal_orientation ? *goal_orientation : start_orient;
  double timestep = 1;
  auto segment_time = [&](double length) {
    const double t_vel = length / max_vel_x;
    if (max_acc_x) {
      const double t_acc = std::sqrt(2 * length / (*max_acc_x));
      return t_vel < t_acc ? t_acc : t_vel;
    }
    return t_vel;
  };
  addPose(PoseSE2(start_position, start_orient));
  setPoseVertexFixed(0, true);
  int idx = 0;
  for (size_t k = 1; k + 1 < path.size(); ++k, ++idx) { /* middle points */
    const Eigen::Vector2d diff_last = path[k] - Pose(idx).position();
    timestep = segment_time(diff_last.norm());
    if (timestep <= 0) timestep = 0.2;
    double yaw = std::atan2(diff_last[1], diff_last[0]);
    if (backwards) yaw = g2o::normalize_theta(yaw + M_PI);
    addPoseAndTimeDiff(PoseSE2(path[k], yaw), timestep);
  }
  timestep = segment_time((goal_position - Pose(idx).position()).norm());
  const PoseSE2 goal(goal_position, goal_orient);
  while (sizePoses() < min_samples - 1) { /* each inserted pose bisects the remaining distance and time */
    timestep /= 2;
    addPoseAndTimeDiff(PoseSE2::average(BackPose(), goal), timestep);
  }
  addPoseAndTimeDiff(goal, timestep);
  setPoseVertexFixed(sizePoses() - 1, true);
  return true;
}

}  // namespace teb_local_planner
