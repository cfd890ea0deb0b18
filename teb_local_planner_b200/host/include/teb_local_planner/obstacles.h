/*
 * obstacles.h / robot footprint models — host-side mirrors for the drop-in planners.
 *
 * Mirrors include/teb_local_planner/obstacles.h (Obstacle :61, PointObstacle :305, CircularObstacle :447) and
 * include/teb_local_planner/robot_footprint_model.h (Point :134, Circular :213, TwoCircles :300). The distance
 * arithmetic itself runs on the device (csrc/teb_device.cuh footprint_distance); these classes carry the data
 * and keep the host-side queries of the reference API. Line / Pill / Polygon obstacles and Line / Polygon
 * footprints are not available yet: constructing the planner with one of them fails loudly.
 */
#ifndef TEB_B200_OBSTACLES_H_
#define TEB_B200_OBSTACLES_H_

#include <memory>
#include <vector>

#include "teb_b200.h"
#include "teb_local_planner/pose_se2.h"

namespace teb_local_planner {

class Obstacle {
 public:
  Obstacle() : dynamic_(false), centroid_velocity_(0, 0) {}
  virtual ~Obstacle() {}
  virtual const Eigen::Vector2d& getCentroid() const = 0;
  virtual double getMinimumDistance(const Eigen::Vector2d& position) const = 0;
  virtual double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& position, double t) const = 0;
  virtual bool checkCollision(const Eigen::Vector2d& position, double min_dist) const {
    return getMinimumDistance(position) < min_dist;
  }
  virtual void predictCentroidConstantVelocity(double t, Eigen::Vector2d& position) const {
    position = getCentroid() + t * getCentroidVelocity();
  }
  bool isDynamic() const { return dynamic_; }
  /* obstacles.h:206 — setting a velocity marks the obstacle as dynamic */
  void setCentroidVelocity(const Eigen::Vector2d& vel) { centroid_velocity_ = vel; dynamic_ = true; }
  const Eigen::Vector2d& getCentroidVelocity() const { return centroid_velocity_; }
  /* row of the device obstacle table (include/teb_b200.h TebObstacle) */
  virtual TebObstacle toRow() const = 0;

 protected:
  bool dynamic_;
  Eigen::Vector2d centroid_velocity_;
};
typedef std::shared_ptr<Obstacle> ObstaclePtr;
typedef std::shared_ptr<const Obstacle> ObstacleConstPtr;
typedef std::vector<ObstaclePtr> ObstContainer;

class PointObstacle : public Obstacle {
 public:
  PointObstacle() : pos_(0, 0) {}
  explicit PointObstacle(const Eigen::Vector2d& position) : pos_(position) {}
  PointObstacle(double x, double y) : pos_(x, y) {}
  const Eigen::Vector2d& getCentroid() const override { return pos_; }
  double getMinimumDistance(const Eigen::Vector2d& position) const override { return (position - pos_).norm(); }
  double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& position, double t) const override {
    return (pos_ + t * centroid_velocity_ - position).norm();
  }
  const Eigen::Vector2d& position() const { return pos_; }
  Eigen::Vector2d& position() { return pos_; }
  double& x() { return pos_.x(); }
  double& y() { return pos_.y(); }
  TebObstacle toRow() const override {
    TebObstacle o{pos_.x(), pos_.y(), centroid_velocity_.x(), centroid_velocity_.y(), 0.0, dynamic_ ? 1 : 0, TEB_OBST_POINT};
    return o;
  }

 protected:
  Eigen::Vector2d pos_;
};

class CircularObstacle : public Obstacle {
 public:
  CircularObstacle() : pos_(0, 0), radius_(0) {}
  CircularObstacle(const Eigen::Vector2d& position, double radius) : pos_(position), radius_(radius) {}
  CircularObstacle(double x, double y, double radius) : pos_(x, y), radius_(radius) {}
  const Eigen::Vector2d& getCentroid() const override { return pos_; }
  double getMinimumDistance(const Eigen::Vector2d& position) const override { return (position - pos_).norm() - radius_; }
  double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& position, double t) const override {
    return (pos_ + t * centroid_velocity_ - position).norm() - radius_;
  }
  const Eigen::Vector2d& position() const { return pos_; }
  Eigen::Vector2d& position() { return pos_; }
  double& radius() { return radius_; }
  const double& radius() const { return radius_; }
  TebObstacle toRow() const override {
    TebObstacle o{pos_.x(), pos_.y(), centroid_velocity_.x(), centroid_velocity_.y(), radius_, dynamic_ ? 1 : 0, TEB_OBST_CIRCULAR};
    return o;
  }

 protected:
  Eigen::Vector2d pos_;
  double radius_;
};

/* ------------------------------------------------------------------ robot footprint models */
class BaseRobotFootprintModel {
 public:
  virtual ~BaseRobotFootprintModel() {}
  virtual double calculateDistance(const PoseSE2& current_pose, const Obstacle* obstacle) const = 0;
  virtual double estimateSpatioTemporalDistance(const PoseSE2& current_pose, const Obstacle* obstacle, double t) const = 0;
  virtual double getInscribedRadius() = 0;
  /* fills footprint_* of the POD parameter block handed to the device */
  virtual void fillParams(TebParams& p) const = 0;
};
typedef std::shared_ptr<BaseRobotFootprintModel> RobotFootprintModelPtr;
typedef std::shared_ptr<const BaseRobotFootprintModel> RobotFootprintModelConstPtr;

class PointRobotFootprint : public BaseRobotFootprintModel {
 public:
  PointRobotFootprint() {}
  double calculateDistance(const PoseSE2& p, const Obstacle* o) const override { return o->getMinimumDistance(p.position()); }
  double estimateSpatioTemporalDistance(const PoseSE2& p, const Obstacle* o, double t) const override {
    return o->getMinimumSpatioTemporalDistance(p.position(), t);
  }
  double getInscribedRadius() override { return 0.0; }
  void fillParams(TebParams& p) const override { p.footprint_type = TEB_FOOTPRINT_POINT; }
};

class CircularRobotFootprint : public BaseRobotFootprintModel {
 public:
  explicit CircularRobotFootprint(double radius) : radius_(radius) {}
  void setRadius(double radius) { radius_ = radius; }
  double calculateDistance(const PoseSE2& p, const Obstacle* o) const override {
    return o->getMinimumDistance(p.position()) - radius_;
  }
  double estimateSpatioTemporalDistance(const PoseSE2& p, const Obstacle* o, double t) const override {
    return o->getMinimumSpatioTemporalDistance(p.position(), t) - radius_;
  }
  double getInscribedRadius() override { return radius_; }
  void fillParams(TebParams& p) const override { p.footprint_type = TEB_FOOTPRINT_CIRCULAR; p.footprint_radius = radius_; }

 private:
  double radius_;
};

class TwoCirclesRobotFootprint : public BaseRobotFootprintModel {
 public:
  TwoCirclesRobotFootprint(double front_offset, double front_radius, double rear_offset, double rear_radius)
      : front_offset_(front_offset), front_radius_(front_radius), rear_offset_(rear_offset), rear_radius_(rear_radius) {}
  void setParameters(double front_offset, double front_radius, double rear_offset, double rear_radius) {
    front_offset_ = front_offset; front_radius_ = front_radius; rear_offset_ = rear_offset; rear_radius_ = rear_radius;
  }
  double calculateDistance(const PoseSE2& p, const Obstacle* o) const override {
    Eigen::Vector2d dir = p.orientationUnitVec();
    double dist_front = o->getMinimumDistance(p.position() + front_offset_ * dir) - front_radius_;
    double dist_rear = o->getMinimumDistance(p.position() - rear_offset_ * dir) - rear_radius_;
    return std::min(dist_front, dist_rear);
  }
  double estimateSpatioTemporalDistance(const PoseSE2& p, const Obstacle* o, double t) const override {
    Eigen::Vector2d dir = p.orientationUnitVec();
    double dist_front = o->getMinimumSpatioTemporalDistance(p.position() + front_offset_ * dir, t) - front_radius_;
    double dist_rear = o->getMinimumSpatioTemporalDistance(p.position() - rear_offset_ * dir, t) - rear_radius_;
    return std::min(dist_front, dist_rear);
  }
  double getInscribedRadius() override {
    double min_longitudinal = std::min(rear_offset_ + rear_radius_, front_offset_ + front_radius_);
    double min_lateral = std::min(rear_radius_, front_radius_);
    return std::min(min_longitudinal, min_lateral);
  }
  void fillParams(TebParams& p) const override {
    p.footprint_type = TEB_FOOTPRINT_TWO_CIRCLES;
    p.footprint_front_offset = front_offset_; p.footprint_front_radius = front_radius_;
    p.footprint_rear_offset = rear_offset_; p.footprint_rear_radius = rear_radius_;
  }

 private:
  double front_offset_, front_radius_, rear_offset_, rear_radius_;
};

typedef std::vector<Eigen::Vector2d> ViaPointContainer;  /* optimal_planner.h:87 */

}  // namespace teb_local_planner
#endif
