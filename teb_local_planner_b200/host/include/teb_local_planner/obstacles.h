/*
 * obstacles.h / robot footprint models — host-side mirrors for the drop-in planners.
 *
 * Mirrors include/teb_local_planner/obstacles.h (Obstacle :61, PointObstacle :305, CircularObstacle :447,
 * LineObstacle :597, PillObstacle :746, PolygonObstacle :893), include/teb_local_planner/robot_footprint_model.h
 * (Point :134, Circular :213, TwoCircles :300, Line :439, Polygon :635) and the 2-D helpers of
 * include/teb_local_planner/distance_calculations.h:60-262. The distance arithmetic of the optimisation runs on the
 * device (csrc/teb_device.cuh footprint_distance / generic_distance); these classes carry the data, fill the obstacle
 * table / vertex pool / footprint parameters and keep the host-side queries of the reference API.
 */
#ifndef TEB_B200_OBSTACLES_H_
#define TEB_B200_OBSTACLES_H_

#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <vector>

#include "teb_b200.h"
#include "teb_local_planner/pose_se2.h"

namespace teb_local_planner {

typedef std::vector<Eigen::Vector2d> Point2dContainer; /* obstacles.h typedefs / distance_calculations.h:51 */

/* ------------------------------------------------------------------ distance_calculations.h:60-262 (2-D part) */
inline Eigen::Vector2d closest_point_on_line_segment_2d(const Eigen::Vector2d& point, const Eigen::Vector2d& line_start,
                                                        const Eigen::Vector2d& line_end) {
  const Eigen::Vector2d dir = line_end - line_start;
  const double len2 = dir.dot(dir);
  if (len2 == 0) return line_start;
  const double u = (point - line_start).dot(dir) / len2;
  if (u <= 0) return line_start;
  if (u >= 1) return line_end;
  return line_start + u * dir;
}
inline double distance_point_to_segment_2d(const Eigen::Vector2d& point, const Eigen::Vector2d& line_start,
                                           const Eigen::Vector2d& line_end) {
  return (point - closest_point_on_line_segment_2d(point, line_start, line_end)).norm();
}
inline bool check_line_segments_intersection_2d(const Eigen::Vector2d& line1_start, const Eigen::Vector2d& line1_end,
                                                const Eigen::Vector2d& line2_start, const Eigen::Vector2d& line2_end,
                                                Eigen::Vector2d* intersection = nullptr) {
  const Eigen::Vector2d d1 = line1_end - line1_start, d2 = line2_end - line2_start, w = line1_start - line2_start;
  const double denom = d1.x() * d2.y() - d2.x() * d1.y();
  if (denom == 0) return false; /* parallel or collinear */
  const bool pos = denom > 0;
  const double s_num = d1.x() * w.y() - d1.y() * w.x();
  if ((s_num < 0) == pos) return false;
  const double t_num = d2.x() * w.y() - d2.y() * w.x();
  if ((t_num < 0) == pos) return false;
  if (((s_num > denom) == pos) || ((t_num > denom) == pos)) return false;
  if (intersection) *intersection = line1_start + (t_num / denom) * d1;
  return true;
}
inline double distance_segment_to_segment_2d(const Eigen::Vector2d& line1_start, const Eigen::Vector2d& line1_end,
                                             const Eigen::Vector2d& line2_start, const Eigen::Vector2d& line2_end) {
  if (check_line_segments_intersection_2d(line1_start, line1_end, line2_start, line2_end)) return 0;
  double d = distance_point_to_segment_2d(line1_start, line2_start, line2_end);
  d = std::min(d, distance_point_to_segment_2d(line1_end, line2_start, line2_end));
  d = std::min(d, distance_point_to_segment_2d(line2_start, line1_start, line1_end));
  return std::min(d, distance_point_to_segment_2d(line2_end, line1_start, line1_end));
}
/* number of edges of a vertex list: 1 vertex = point, 2 = one segment, more = closed polygon */
inline int polygon_edge_count(const Point2dContainer& v) { return v.size() <= 2 ? (int)v.size() - 1 : (int)v.size(); }
inline double distance_point_to_polygon_2d(const Eigen::Vector2d& point, const Point2dContainer& vertices) {
  if (vertices.size() == 1) return (point - vertices.front()).norm();
  double d = HUGE_VAL;
  for (int i = 0; i < polygon_edge_count(vertices); ++i)
    d = std::min(d, distance_point_to_segment_2d(point, vertices[i], vertices[(i + 1) % vertices.size()]));
  return d;
}
inline double distance_segment_to_polygon_2d(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end,
                                             const Point2dContainer& vertices) {
  if (vertices.size() == 1) return distance_point_to_segment_2d(vertices.front(), line_start, line_end);
  double d = HUGE_VAL;
  for (int i = 0; i < polygon_edge_count(vertices); ++i)
    d = std::min(d, distance_segment_to_segment_2d(line_start, line_end, vertices[i], vertices[(i + 1) % vertices.size()]));
  return d;
}
inline double distance_polygon_to_polygon_2d(const Point2dContainer& vertices1, const Point2dContainer& vertices2) {
  if (vertices1.size() == 1) return distance_point_to_polygon_2d(vertices1.front(), vertices2);
  double d = HUGE_VAL;
  for (int i = 0; i < polygon_edge_count(vertices1); ++i)
    d = std::min(d, distance_segment_to_polygon_2d(vertices1[i], vertices1[(i + 1) % vertices1.size()], vertices2));
  return d;
}

class Obstacle {
 public:
  Obstacle() : dynamic_(false), centroid_velocity_(0, 0) {}
  virtual ~Obstacle() {}
  virtual const Eigen::Vector2d& getCentroid() const = 0;
  virtual double getMinimumDistance(const Eigen::Vector2d& position) const = 0;
  virtual double getMinimumDistance(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end) const = 0;
  virtual double getMinimumDistance(const Point2dContainer& polygon) const = 0;
  virtual double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& position, double t) const = 0;
  virtual double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end, double t) const = 0;
  virtual double getMinimumSpatioTemporalDistance(const Point2dContainer& polygon, double t) const = 0;
  virtual bool checkCollision(const Eigen::Vector2d& position, double min_dist) const {
    return getMinimumDistance(position) < min_dist;
  }
  /* does the segment come closer than min_dist (Point / Circular) resp. cross the shape (Line / Pill / Polygon,
   * which ignore min_dist like the reference: obstacles.h:339, :483, :647, :794; obstacles.cpp:176-191) */
  virtual bool checkLineIntersection(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end, double min_dist = 0) const = 0;
  virtual void predictCentroidConstantVelocity(double t, Eigen::Vector2d& position) const {
    position = getCentroid() + t * getCentroidVelocity();
  }
  bool isDynamic() const { return dynamic_; }
  /* obstacles.h:206 — setting a velocity marks the obstacle as dynamic */
  void setCentroidVelocity(const Eigen::Vector2d& vel) { centroid_velocity_ = vel; dynamic_ = true; }
  const Eigen::Vector2d& getCentroidVelocity() const { return centroid_velocity_; }
  /* row of the device obstacle table (include/teb_b200.h TebObstacle); vertex-list shapes append their vertices to
   * the scene's pool (x0, y0, x1, y1, ...) and record the range */
  virtual TebObstacle toRow() const = 0;
  virtual void appendVertices(std::vector<double>& pool, TebObstacle& row) const { (void)pool; (void)row; }

 protected:
  static void appendList(const Point2dContainer& v, std::vector<double>& pool, TebObstacle& row) {
    row.vertex_begin = (int32_t)(pool.size() / 2);
    row.vertex_count = (int32_t)v.size();
    for (const Eigen::Vector2d& p : v) { pool.push_back(p.x()); pool.push_back(p.y()); }
  }
  static Point2dContainer shifted(const Point2dContainer& v, const Eigen::Vector2d& off) {
    Point2dContainer out(v);
    for (Eigen::Vector2d& p : out) p = p + off;
    return out;
  }

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
  double getMinimumDistance(const Eigen::Vector2d& a, const Eigen::Vector2d& b) const override { return distance_point_to_segment_2d(pos_, a, b); }
  double getMinimumDistance(const Point2dContainer& polygon) const override { return distance_point_to_polygon_2d(pos_, polygon); }
  double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& position, double t) const override {
    return (pos_ + t * centroid_velocity_ - position).norm();
  }
  double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& a, const Eigen::Vector2d& b, double t) const override {
    return distance_point_to_segment_2d(pos_ + t * centroid_velocity_, a, b);
  }
  double getMinimumSpatioTemporalDistance(const Point2dContainer& polygon, double t) const override {
    return distance_point_to_polygon_2d(pos_ + t * centroid_velocity_, polygon);
  }
  bool checkLineIntersection(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end, double min_dist = 0) const override {
    /* nearest point of the segment to the centre (parameter clamped to [0, 1]), then the point test */
    const Eigen::Vector2d a = line_end - line_start, b = pos_ - line_start;
    double t = a.dot(b) / a.dot(a);
    if (t < 0) t = 0; else if (t > 1) t = 1;
    return checkCollision(line_start + a * t, min_dist);
  }
  const Eigen::Vector2d& position() const { return pos_; }
  Eigen::Vector2d& position() { return pos_; }
  double& x() { return pos_.x(); }
  double& y() { return pos_.y(); }
  TebObstacle toRow() const override {
    TebObstacle o{pos_.x(), pos_.y(), centroid_velocity_.x(), centroid_velocity_.y(), 0.0, dynamic_ ? 1 : 0, TEB_OBST_POINT, 0, 0, 0.0};
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
  double getMinimumDistance(const Eigen::Vector2d& a, const Eigen::Vector2d& b) const override { return distance_point_to_segment_2d(pos_, a, b) - radius_; }
  double getMinimumDistance(const Point2dContainer& polygon) const override { return distance_point_to_polygon_2d(pos_, polygon) - radius_; }
  double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& position, double t) const override {
    return (pos_ + t * centroid_velocity_ - position).norm() - radius_;
  }
  double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& a, const Eigen::Vector2d& b, double t) const override {
    return distance_point_to_segment_2d(pos_ + t * centroid_velocity_, a, b) - radius_;
  }
  double getMinimumSpatioTemporalDistance(const Point2dContainer& polygon, double t) const override {
    return distance_point_to_polygon_2d(pos_ + t * centroid_velocity_, polygon) - radius_;
  }
  bool checkLineIntersection(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end, double min_dist = 0) const override {
    /* nearest point of the segment to the centre (parameter clamped to [0, 1]), then the point test */
    const Eigen::Vector2d a = line_end - line_start, b = pos_ - line_start;
    double t = a.dot(b) / a.dot(a);
    if (t < 0) t = 0; else if (t > 1) t = 1;
    return checkCollision(line_start + a * t, min_dist);
  }
  const Eigen::Vector2d& position() const { return pos_; }
  Eigen::Vector2d& position() { return pos_; }
  double& radius() { return radius_; }
  const double& radius() const { return radius_; }
  TebObstacle toRow() const override {
    TebObstacle o{pos_.x(), pos_.y(), centroid_velocity_.x(), centroid_velocity_.y(), radius_, dynamic_ ? 1 : 0, TEB_OBST_CIRCULAR, 0, 0, 0.0};
    return o;
  }

 protected:
  Eigen::Vector2d pos_;
  double radius_;
};


/* LineObstacle obstacles.h:597-740 and PillObstacle :746-890 (a line with a radius) */
class LineObstacle : public Obstacle {
 public:
  LineObstacle() : radius_(0) { verts_.resize(2); calcCentroid(); }
  LineObstacle(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end) : radius_(0) {
    verts_.push_back(line_start); verts_.push_back(line_end); calcCentroid();
  }
  LineObstacle(double x1, double y1, double x2, double y2) : LineObstacle(Eigen::Vector2d(x1, y1), Eigen::Vector2d(x2, y2)) {}
  const Eigen::Vector2d& getCentroid() const override { return centroid_; }
  bool checkCollision(const Eigen::Vector2d& point, double min_dist) const override { return getMinimumDistance(point) <= min_dist; }
  double getMinimumDistance(const Eigen::Vector2d& position) const override {
    return distance_point_to_segment_2d(position, verts_[0], verts_[1]) - radius_;
  }
  double getMinimumDistance(const Eigen::Vector2d& a, const Eigen::Vector2d& b) const override {
    return distance_segment_to_segment_2d(verts_[0], verts_[1], a, b) - radius_;
  }
  double getMinimumDistance(const Point2dContainer& polygon) const override {
    return distance_segment_to_polygon_2d(verts_[0], verts_[1], polygon) - radius_;
  }
  double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& position, double t) const override {
    const Eigen::Vector2d off = t * centroid_velocity_;
    return distance_point_to_segment_2d(position, verts_[0] + off, verts_[1] + off) - radius_;
  }
  double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& a, const Eigen::Vector2d& b, double t) const override {
    const Eigen::Vector2d off = t * centroid_velocity_;
    return distance_segment_to_segment_2d(verts_[0] + off, verts_[1] + off, a, b) - radius_;
  }
  double getMinimumSpatioTemporalDistance(const Point2dContainer& polygon, double t) const override {
    const Eigen::Vector2d off = t * centroid_velocity_;
    return distance_segment_to_polygon_2d(verts_[0] + off, verts_[1] + off, polygon) - radius_;
  }
  bool checkLineIntersection(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end, double min_dist = 0) const override {
    (void)min_dist;
    return check_line_segments_intersection_2d(line_start, line_end, verts_[0], verts_[1]);
  }
  const Eigen::Vector2d& start() const { return verts_[0]; }
  const Eigen::Vector2d& end() const { return verts_[1]; }
  void setStart(const Eigen::Vector2d& p) { verts_[0] = p; calcCentroid(); }
  void setEnd(const Eigen::Vector2d& p) { verts_[1] = p; calcCentroid(); }
  TebObstacle toRow() const override {
    TebObstacle o{centroid_.x(), centroid_.y(), centroid_velocity_.x(), centroid_velocity_.y(), radius_, dynamic_ ? 1 : 0,
                  radius_ > 0 ? TEB_OBST_PILL : TEB_OBST_LINE, 0, 0, 0.0};
    return o;
  }
  void appendVertices(std::vector<double>& pool, TebObstacle& row) const override { appendList(verts_, pool, row); }

 protected:
  void calcCentroid() { centroid_ = 0.5 * (verts_[0] + verts_[1]); }
  Point2dContainer verts_;
  Eigen::Vector2d centroid_;
  double radius_;
};

class PillObstacle : public LineObstacle {
 public:
  PillObstacle() { radius_ = 0; }
  PillObstacle(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end, double radius) : LineObstacle(line_start, line_end) {
    radius_ = radius;
  }
  PillObstacle(double x1, double y1, double x2, double y2, double radius) : LineObstacle(x1, y1, x2, y2) { radius_ = radius; }
  void setRadius(double radius) { radius_ = radius; }
  double radius() const { return radius_; }
  TebObstacle toRow() const override {
    TebObstacle o = LineObstacle::toRow();
    o.type = TEB_OBST_PILL;
    return o;
  }
};

/* PolygonObstacle obstacles.h:893-1045; vertices in order, not closed; finalizePolygon() computes the centroid
 * (src/obstacles.cpp:56-119) */
class PolygonObstacle : public Obstacle {
 public:
  PolygonObstacle() : finalized_(false), centroid_(NAN, NAN) {}
  explicit PolygonObstacle(const Point2dContainer& vertices) : verts_(vertices), finalized_(false) { finalizePolygon(); }
  void pushBackVertex(const Eigen::Vector2d& vertex) { verts_.push_back(vertex); finalized_ = false; }
  void pushBackVertex(double x, double y) { pushBackVertex(Eigen::Vector2d(x, y)); }
  void finalizePolygon() {
    if (verts_.size() >= 2 && (verts_.front() - verts_.back()).norm() <= 1e-12 * std::max(1.0, verts_.front().norm()))
      verts_.pop_back(); /* fixPolygonClosure: the first vertex must not be repeated at the end */
    calcCentroid();
    finalized_ = true;
  }
  void clearVertices() { verts_.clear(); finalized_ = false; }
  int noVertices() const { return (int)verts_.size(); }
  const Point2dContainer& vertices() const { return verts_; }
  const Eigen::Vector2d& getCentroid() const override { return centroid_; }
  bool checkCollision(const Eigen::Vector2d& point, double min_dist) const override {
    if (noVertices() == 2) return getMinimumDistance(point) <= min_dist;
    bool inside = false; /* ray casting; points exactly on an edge may go either way, as in the reference */
    for (int i = 0, j = noVertices() - 1; i < noVertices(); j = i++) {
      const Eigen::Vector2d &a = verts_[i], &b = verts_[j];
      if (((a.y() > point.y()) != (b.y() > point.y())) &&
          (point.x() < (b.x() - a.x()) * (point.y() - a.y()) / (b.y() - a.y()) + a.x()))
        inside = !inside;
    }
    if (inside) return true;
    return min_dist == 0 ? false : getMinimumDistance(point) < min_dist;
  }
  bool checkLineIntersection(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end, double min_dist = 0) const override {
    (void)min_dist;
    for (int i = 0; i + 1 < noVertices(); ++i)
      if (check_line_segments_intersection_2d(line_start, line_end, verts_[i], verts_[i + 1])) return true;
    if (noVertices() <= 2) return false;
    return check_line_segments_intersection_2d(line_start, line_end, verts_.back(), verts_.front());
  }
  double getMinimumDistance(const Eigen::Vector2d& position) const override { return distance_point_to_polygon_2d(position, verts_); }
  double getMinimumDistance(const Eigen::Vector2d& a, const Eigen::Vector2d& b) const override {
    return distance_segment_to_polygon_2d(a, b, verts_);
  }
  double getMinimumDistance(const Point2dContainer& polygon) const override { return distance_polygon_to_polygon_2d(polygon, verts_); }
  double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& position, double t) const override {
    return distance_point_to_polygon_2d(position, shifted(verts_, t * centroid_velocity_));
  }
  double getMinimumSpatioTemporalDistance(const Eigen::Vector2d& a, const Eigen::Vector2d& b, double t) const override {
    return distance_segment_to_polygon_2d(a, b, shifted(verts_, t * centroid_velocity_));
  }
  double getMinimumSpatioTemporalDistance(const Point2dContainer& polygon, double t) const override {
    return distance_polygon_to_polygon_2d(polygon, shifted(verts_, t * centroid_velocity_));
  }
  TebObstacle toRow() const override {
    TebObstacle o{centroid_.x(), centroid_.y(), centroid_velocity_.x(), centroid_velocity_.y(), 0.0, dynamic_ ? 1 : 0, TEB_OBST_POLYGON, 0, 0, 0.0};
    return o;
  }
  void appendVertices(std::vector<double>& pool, TebObstacle& row) const override { appendList(verts_, pool, row); }

 protected:
  void calcCentroid() {
    const int k = noVertices();
    if (k == 0) { centroid_ = Eigen::Vector2d(NAN, NAN); return; }
    if (k == 1) { centroid_ = verts_[0]; return; }
    if (k == 2) { centroid_ = 0.5 * (verts_[0] + verts_[1]); return; }
    double area2 = 0; /* twice the signed area */
    Eigen::Vector2d acc(0, 0);
    for (int i = 0; i < k; ++i) {
      const Eigen::Vector2d &a = verts_[i], &b = verts_[(i + 1) % k];
      const double cr = a.x() * b.y() - b.x() * a.y();
      area2 += cr;
      acc = acc + (a + b) * cr;
    }
    if (area2 != 0) { centroid_ = acc / (3 * area2); return; }
    /* all vertices on one line: midpoint of the two vertices that are farthest apart */
    int bi = 0, bj = 0;
    double far = 0;
    for (int i = 0; i < k; ++i)
      for (int j = i + 1; j < k; ++j) {
        const double d = (verts_[i] - verts_[j]).norm();
        if (d > far) { far = d; bi = i; bj = j; }
      }
    centroid_ = 0.5 * (verts_[bi] + verts_[bj]);
  }
  Point2dContainer verts_;
  bool finalized_;
  Eigen::Vector2d centroid_;
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


/* LineRobotFootprint robot_footprint_model.h:439-625 */
class LineRobotFootprint : public BaseRobotFootprintModel {
 public:
  LineRobotFootprint(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end, double min_obstacle_dist = 0.0)
      : line_start_(line_start), line_end_(line_end), min_obstacle_dist_(min_obstacle_dist) {}
  void setLine(const Eigen::Vector2d& line_start, const Eigen::Vector2d& line_end) { line_start_ = line_start; line_end_ = line_end; }
  double calculateDistance(const PoseSE2& p, const Obstacle* o) const override {
    Eigen::Vector2d a, b;
    transformToWorld(p, a, b);
    return o->getMinimumDistance(a, b);
  }
  double estimateSpatioTemporalDistance(const PoseSE2& p, const Obstacle* o, double t) const override {
    Eigen::Vector2d a, b;
    transformToWorld(p, a, b);
    return o->getMinimumSpatioTemporalDistance(a, b, t);
  }
  double getInscribedRadius() override { return 0.0; }
  void fillParams(TebParams& p) const override {
    p.footprint_type = TEB_FOOTPRINT_LINE;
    p.footprint_line[0] = line_start_.x(); p.footprint_line[1] = line_start_.y();
    p.footprint_line[2] = line_end_.x(); p.footprint_line[3] = line_end_.y();
  }

 private:
  void transformToWorld(const PoseSE2& pose, Eigen::Vector2d& a, Eigen::Vector2d& b) const {
    const double c = std::cos(pose.theta()), s = std::sin(pose.theta());
    a = Eigen::Vector2d(pose.x() + c * line_start_.x() - s * line_start_.y(), pose.y() + s * line_start_.x() + c * line_start_.y());
    b = Eigen::Vector2d(pose.x() + c * line_end_.x() - s * line_end_.y(), pose.y() + s * line_end_.x() + c * line_end_.y());
  }
  Eigen::Vector2d line_start_, line_end_;
  double min_obstacle_dist_;
};

/* PolygonRobotFootprint robot_footprint_model.h:635-775; at most TEB_MAX_FOOTPRINT_VERTICES vertices on the device */
class PolygonRobotFootprint : public BaseRobotFootprintModel {
 public:
  explicit PolygonRobotFootprint(const Point2dContainer& vertices) : vertices_(vertices) {}
  void setVertices(const Point2dContainer& vertices) { vertices_ = vertices; }
  double calculateDistance(const PoseSE2& p, const Obstacle* o) const override { return o->getMinimumDistance(world(p)); }
  double estimateSpatioTemporalDistance(const PoseSE2& p, const Obstacle* o, double t) const override {
    return o->getMinimumSpatioTemporalDistance(world(p), t);
  }
  double getInscribedRadius() override {
    if (vertices_.size() <= 2) return 0.0;
    const Eigen::Vector2d centre(0, 0);
    double r = std::numeric_limits<double>::max();
    for (size_t i = 0; i < vertices_.size(); ++i) {
      const Eigen::Vector2d &a = vertices_[i], &b = vertices_[(i + 1) % vertices_.size()];
      r = std::min(r, std::min(a.norm(), distance_point_to_segment_2d(centre, a, b)));
    }
    return r;
  }
  void fillParams(TebParams& p) const override {
    p.footprint_type = TEB_FOOTPRINT_POLYGON;
    /* more vertices than the device table holds: the count is passed through and tebgpu_set_params refuses it */
    p.footprint_vertex_count = (int32_t)vertices_.size();
    for (size_t k = 0; k < vertices_.size() && k < TEB_MAX_FOOTPRINT_VERTICES; ++k) {
      p.footprint_vertices[2 * k] = vertices_[k].x();
      p.footprint_vertices[2 * k + 1] = vertices_[k].y();
    }
  }

 private:
  Point2dContainer world(const PoseSE2& pose) const {
    const double c = std::cos(pose.theta()), s = std::sin(pose.theta());
    Point2dContainer w(vertices_.size());
    for (size_t i = 0; i < vertices_.size(); ++i)
      w[i] = Eigen::Vector2d(pose.x() + c * vertices_[i].x() - s * vertices_[i].y(), pose.y() + s * vertices_[i].x() + c * vertices_[i].y());
    return w;
  }
  Point2dContainer vertices_;
};

typedef std::vector<Eigen::Vector2d> ViaPointContainer;  /* optimal_planner.h:87 */

}  // namespace teb_local_planner
#endif
