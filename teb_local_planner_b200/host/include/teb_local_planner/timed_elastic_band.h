/*
 * timed_elastic_band.h — host-side band container of the drop-in planners.
 *
 * Mirrors teb_local_planner::TimedElasticBand (include/teb_local_planner/timed_elastic_band.h:86-655,
 * src/timed_elastic_band.cpp): same public methods and semantics for the part of the API the planners on the hot
 * path use. Storage differs on purpose: poses and time differences are plain value vectors (the reference owns g2o
 * vertex pointers, timed_elastic_band.cpp:80-107) because the optimiser lives on the GPU and takes packed records
 * (x, y, theta, dt). PoseVertex()/TimeDiffVertex() therefore do not exist. Start and goal pose are fixed during
 * optimisation (the only fixed vertices the reference's own init paths create, :330, :377).
 * autoResize runs the same routine as the device kernel (csrc/teb_resize.h).
 */
#ifndef TEB_B200_TIMED_ELASTIC_BAND_H_
#define TEB_B200_TIMED_ELASTIC_BAND_H_

#include <limits>
#include <vector>

#include "teb_local_planner/obstacles.h"
#include "teb_local_planner/pose_se2.h"

namespace teb_local_planner {

typedef std::vector<PoseSE2> PoseSequence;
typedef std::vector<double> TimeDiffSequence;

class TimedElasticBand {
 public:
  TimedElasticBand() {}
  virtual ~TimedElasticBand() {}

  PoseSequence& poses() { return pose_vec_; }
  const PoseSequence& poses() const { return pose_vec_; }
  TimeDiffSequence& timediffs() { return timediff_vec_; }
  const TimeDiffSequence& timediffs() const { return timediff_vec_; }
  double& TimeDiff(int index) { return timediff_vec_.at(index); }
  const double& TimeDiff(int index) const { return timediff_vec_.at(index); }
  PoseSE2& Pose(int index) { return pose_vec_.at(index); }
  const PoseSE2& Pose(int index) const { return pose_vec_.at(index); }
  PoseSE2& BackPose() { return pose_vec_.back(); }
  const PoseSE2& BackPose() const { return pose_vec_.back(); }
  double& BackTimeDiff() { return timediff_vec_.back(); }
  const double& BackTimeDiff() const { return timediff_vec_.back(); }

  void addPose(const PoseSE2& pose, bool fixed = false) { (void)fixed; pose_vec_.push_back(pose); }
  void addPose(const Eigen::Vector2d& position, double theta, bool fixed = false) { addPose(PoseSE2(position, theta), fixed); }
  void addPose(double x, double y, double theta, bool fixed = false) { addPose(PoseSE2(x, y, theta), fixed); }
  void addTimeDiff(double dt, bool fixed = false) { (void)fixed; timediff_vec_.push_back(dt); }
  void addPoseAndTimeDiff(const PoseSE2& pose, double dt);
  void addPoseAndTimeDiff(const Eigen::Vector2d& position, double theta, double dt) { addPoseAndTimeDiff(PoseSE2(position, theta), dt); }
  void addPoseAndTimeDiff(double x, double y, double theta, double dt) { addPoseAndTimeDiff(PoseSE2(x, y, theta), dt); }
  void insertPose(int index, const PoseSE2& pose) { pose_vec_.insert(pose_vec_.begin() + index, pose); }
  void insertPose(int index, double x, double y, double theta) { insertPose(index, PoseSE2(x, y, theta)); }
  void insertTimeDiff(int index, double dt) { timediff_vec_.insert(timediff_vec_.begin() + index, dt); }
  void deletePose(int index) { pose_vec_.erase(pose_vec_.begin() + index); }
  void deletePoses(int index, int number) { pose_vec_.erase(pose_vec_.begin() + index, pose_vec_.begin() + index + number); }
  void deleteTimeDiff(int index) { timediff_vec_.erase(timediff_vec_.begin() + index); }
  void deleteTimeDiffs(int index, int number) { timediff_vec_.erase(timediff_vec_.begin() + index, timediff_vec_.begin() + index + number); }

  bool initTrajectoryToGoal(const PoseSE2& start, const PoseSE2& goal, double diststep = 0, double max_vel_x = 0.5,
                            int min_samples = 3, bool guess_backwards_motion = false);
  bool initTrajectoryToGoal(const std::vector<geometry_msgs::PoseStamped>& plan, double max_vel_x, double max_vel_theta,
                            bool estimate_orient = false, int min_samples = 3, bool guess_backwards_motion = false);
  /* timed_elastic_band.hpp:46-185: initialise from a 2-D path (graph-search candidates); optional values as pointers */
  bool initTrajectoryToGoal(const std::vector<Eigen::Vector2d>& path, double max_vel_x, double max_vel_theta,
                            const double* max_acc_x, const double* max_acc_theta, const double* start_orientation,
                            const double* goal_orientation, int min_samples = 3, bool guess_backwards_motion = false);
  /* pointers instead of boost::optional<const PoseSE2&>; NULL = leave unchanged */
  void updateAndPruneTEB(const PoseSE2* new_start, const PoseSE2* new_goal, int min_samples = 3);
  void updateAndPruneTEB(const PoseSE2& new_start, const PoseSE2& new_goal, int min_samples = 3) {
    updateAndPruneTEB(&new_start, &new_goal, min_samples);
  }
  void autoResize(double dt_ref, double dt_hysteresis, int min_samples = 3, int max_samples = 1000, bool fast_mode = false);
  void setPoseVertexFixed(int index, bool status);
  void clearTimedElasticBand() { pose_vec_.clear(); timediff_vec_.clear(); }

  int findClosestTrajectoryPose(const Eigen::Vector2d& ref_point, double* distance = NULL, int begin_idx = 0) const;
  int sizePoses() const { return (int)pose_vec_.size(); }
  int sizeTimeDiffs() const { return (int)timediff_vec_.size(); }
  bool isInit() const { return !timediff_vec_.empty() && !pose_vec_.empty(); }
  double getSumOfAllTimeDiffs() const;
  double getSumOfTimeDiffsUpToIdx(int index) const;
  double getAccumulatedDistance() const;
  bool isTrajectoryInsideRegion(double radius, double max_dist_behind_robot = -1, int skip_poses = 0);

  /* packed records for the C-ABI: rec[n_cap][4] = x, y, theta, dt (dt of the last pose = 0) */
  void toRecords(double* rec) const;
  void fromRecords(const double* rec, int n);

 protected:
  PoseSequence pose_vec_;
  TimeDiffSequence timediff_vec_;
};

}  // namespace teb_local_planner
#endif
