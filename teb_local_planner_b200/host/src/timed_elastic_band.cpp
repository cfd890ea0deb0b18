/* timed_elastic_band.cpp — see include/teb_local_planner/timed_elastic_band.h.
 * Behaviour follows src/timed_elastic_band.cpp of the reference; line references below point there. */
#include "teb_local_planner/timed_elastic_band.h"

#include <algorithm>
#include <cmath>

#include "../../csrc/teb_resize.h"

namespace teb_local_planner {

namespace {

/* Time a robot limited to (v_max, w_max) needs from pose a to pose b when translation and rotation run concurrently:
 * the slower of the two decides; 0.1 s when no translational limit is given (reference estimateDeltaT, :52-65). */
double travel_time(const PoseSE2& a, const PoseSE2& b, double v_max, double w_max) {
  double t = 0.1;
  if (v_max > 0) t = (b.position() - a.position()).norm() / v_max;
  if (w_max > 0) t = std::max(t, std::abs(g2o::normalize_theta(b.theta() - a.theta())) / w_max);
  return t;
}

/* true when the goal lies behind the start pose's heading */
bool goal_is_behind(const Eigen::Vector2d& start_to_goal, const PoseSE2& start) {
  return start_to_goal.dot(start.orientationUnitVec()) < 0;
}

}  // namespace

/* a pose / time-difference pair can only be appended while the band holds one more pose than time differences (:141-150) */
void TimedElasticBand::addPoseAndTimeDiff(const PoseSE2& pose, double dt) {
  if (sizePoses() == sizeTimeDiffs()) return;
  addPose(pose, false);
  addTimeDiff(dt, false);
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

/* Straight-line band from start to goal (reference :325-387): samples every `diststep` metres along the connecting
 * line, all with the line's heading (turned by pi when driving backwards is guessed), constant time step diststep /
 * max_vel_x; a sample that would coincide with the goal is left out; midpoints towards the goal are appended until
 * min_samples is reached; the goal closes the band. Refuses to touch an initialised band. */
bool TimedElasticBand::initTrajectoryToGoal(const PoseSE2& start, const PoseSE2& goal, double diststep, double max_vel_x,
                                            int min_samples, bool guess_backwards_motion) {
  if (isInit()) return false;
  const bool timed = max_vel_x > 0;
  double step_time = 0.1;
  addPose(start);
  setPoseVertexFixed(0, true);
  if (diststep != 0) {
    const Eigen::Vector2d to_goal = goal.position() - start.position();
    const double line_heading = std::atan2(to_goal[1], to_goal[0]);
    const double step_x = diststep * std::cos(line_heading), step_y = diststep * std::sin(line_heading);
    const double heading = (guess_backwards_motion && goal_is_behind(to_goal, start)) ? g2o::normalize_theta(line_heading + M_PI)
                                                                                      : line_heading;
    const double steps_exact = to_goal.norm() / std::abs(diststep);
    const unsigned int steps = (unsigned int)std::floor(steps_exact);
    const bool last_hits_goal = steps_exact == (float)steps; /* single-precision comparison, as the reference does */
    if (timed) step_time = diststep / max_vel_x;
    for (unsigned int k = 1; k <= steps; ++k) {
      if (k == steps && last_hits_goal) break;
      addPoseAndTimeDiff(start.x() + k * step_x, start.y() + k * step_y, heading, step_time);
    }
  }
  while (sizePoses() < min_samples - 1) { /* too few samples: keep bisecting the rest of the way */
    const PoseSE2 mid = PoseSE2::average(BackPose(), goal);
    if (timed) step_time = (mid.position() - BackPose().position()).norm() / max_vel_x;
    addPoseAndTimeDiff(mid, step_time);
  }
  if (timed) step_time = (goal.position() - BackPose().position()).norm() / max_vel_x;
  addPoseAndTimeDiff(goal, step_time);
  setPoseVertexFixed(sizePoses() - 1, true);
  return true;
}

/* Band along a reference plan (reference :389-452): one pose per inner plan point, its heading either taken from the
 * plan or estimated from the direction to the next plan point (turned by pi for a guessed backwards motion), time
 * differences from travel_time(); padded with midpoints up to min_samples; closed by the plan's last pose. */
bool TimedElasticBand::initTrajectoryToGoal(const std::vector<geometry_msgs::PoseStamped>& plan, double max_vel_x,
                                            double max_vel_theta, bool estimate_orient, int min_samples,
                                            bool guess_backwards_motion) {
  if (isInit()) return false;
  const PoseSE2 first(plan.front().pose), last(plan.back().pose);
  const bool reversed = guess_backwards_motion && goal_is_behind(last.position() - first.position(), first);
  auto append = [&](const PoseSE2& pose) { addPoseAndTimeDiff(pose, travel_time(BackPose(), pose, max_vel_x, max_vel_theta)); };
  addPose(first);
  setPoseVertexFixed(0, true);
  const int inner_end = (int)plan.size() - 1;
  for (int k = 1; k < inner_end; ++k) {
    const geometry_msgs::Point& here = plan[k].pose.position;
    double heading = tf::getYaw(plan[k].pose.orientation);
    if (estimate_orient) {
      const geometry_msgs::Point& ahead = plan[k + 1].pose.position;
      heading = std::atan2(ahead.y - here.y, ahead.x - here.x);
      if (reversed) heading = g2o::normalize_theta(heading + M_PI);
    }
    append(PoseSE2(here.x, here.y, heading));
  }
  while (sizePoses() < min_samples - 1) append(PoseSE2::average(BackPose(), last));
  append(last);
  setPoseVertexFixed(sizePoses() - 1, true);
  return true;
}

/* index of the pose (from begin_idx on) nearest to ref_point, -1 for an invalid begin_idx; the first minimum wins
 * (reference :455-478) */
int TimedElasticBand::findClosestTrajectoryPose(const Eigen::Vector2d& ref_point, double* distance, int begin_idx) const {
  const int count = sizePoses();
  if (begin_idx < 0 || begin_idx >= count) return -1;
  int best = -1;
  double best_sq = std::numeric_limits<double>::max();
  for (int k = begin_idx; k < count; ++k) {
    const double sq = (ref_point - Pose(k).position()).squaredNorm();
    if (sq < best_sq) { best_sq = sq; best = k; }
  }
  if (distance) *distance = std::sqrt(best_sq);
  return best;
}

/* Warm start of the next planning cycle (reference :555-597): walking from the front, poses are dropped while they
 * keep getting closer to the new start (at most 10, and never below min_samples); the first pose is then replaced by the
 * new start and the last one by the new goal. */
void TimedElasticBand::updateAndPruneTEB(const PoseSE2* new_start, const PoseSE2* new_goal, int min_samples) {
  if (sizePoses() == 0) return;
  if (new_start) {
    const int reach = std::min<int>(sizePoses() - min_samples, 10);
    double nearest = (new_start->position() - Pose(0).position()).norm();
    int drop = 0; /* poses 1 .. drop go away */
    while (drop < reach) {
      const double d = (new_start->position() - Pose(drop + 1).position()).norm();
      if (!(d < nearest)) break;
      nearest = d;
      ++drop;
    }
    if (drop > 0) {
      deletePoses(1, drop);
      deleteTimeDiffs(1, drop);
    }
    Pose(0) = *new_start;
  }
  if (new_goal) BackPose() = *new_goal;
}

double TimedElasticBand::getSumOfAllTimeDiffs() const {
  double total = 0;
  for (double dt : timediff_vec_) total += dt;
  return total;
}
double TimedElasticBand::getSumOfTimeDiffsUpToIdx(int index) const {
  double total = 0;
  for (int k = 0; k < index; ++k) total += timediff_vec_.at(k);
  return total;
}
double TimedElasticBand::getAccumulatedDistance() const {
  double length = 0;
  for (int k = 1; k < sizePoses(); ++k) length += (Pose(k).position() - Pose(k - 1).position()).norm();
  return length;
}

/* every (skip_poses + 1)-th pose must lie within `radius` of the robot and, when max_dist_behind_robot >= 0, not further
 * than that behind it (reference :600-631) */
bool TimedElasticBand::isTrajectoryInsideRegion(double radius, double max_dist_behind_robot, int skip_poses) {
  const int count = sizePoses();
  if (count <= 0) return true;
  const Eigen::Vector2d origin = Pose(0).position(), forward = Pose(0).orientationUnitVec();
  const bool check_behind = max_dist_behind_robot >= 0;
  for (int k = 1; k < count; k += skip_poses + 1) {
    const Eigen::Vector2d offset = Pose(k).position() - origin;
    const double sq = offset.squaredNorm();
    const bool too_far = sq > radius * radius;
    const bool too_far_behind = check_behind && offset.dot(forward) < 0 && sq > max_dist_behind_robot * max_dist_behind_robot;
    if (too_far || too_far_behind) return false;
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
