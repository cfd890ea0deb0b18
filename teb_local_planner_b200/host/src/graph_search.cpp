/* graph_search.cpp — see include/teb_local_planner/graph_search.h. Line references: src/graph_search.cpp of the reference. */
#include "teb_local_planner/graph_search.h"

#include <algorithm>
#include <cfloat>
#include <cmath>

#include "teb_local_planner/homotopy_class_planner.h"

namespace teb_local_planner {

bool GraphSearchInterface::classesFull() const {
  return (int)hcp_->getTrajectoryContainer().size() >= cfg_->hcp.max_number_classes;
}

void GraphSearchInterface::flushPaths(double start_orientation, double goal_orientation, const geometry_msgs::Twist* start_velocity,
                                      bool free_goal_vel) {
  if (pending_.empty()) return;
  /* candidates are initialised, classified in one device call and accepted in enumeration order */
  if (!hcp_->addAndInitNewTebs(pending_, start_orientation, goal_orientation, start_velocity, free_goal_vel)) stop_ = true;
  pending_.clear();
  if (classesFull()) stop_ = true;
}

/* :45-88: all simple forward paths from visited.back() to goal, goal-adjacent first, then recursion in edge order */
void GraphSearchInterface::DepthFirst(std::vector<int>& visited, int goal, double start_orientation, double goal_orientation,
                                      const geometry_msgs::Twist* start_velocity, bool free_goal_vel) {
  if (stop_) return;
  const int back = visited.back();
  for (int v : graph_.adj[back]) {
    if (std::find(visited.begin(), visited.end(), v) != visited.end()) continue;
    if (v == goal) {
      std::vector<Eigen::Vector2d> path;
      for (int u : visited) path.push_back(graph_.pos[u]);
      path.push_back(graph_.pos[goal]);
      pending_.push_back(path);
      if ((int)pending_.size() >= CHUNK) flushPaths(start_orientation, goal_orientation, start_velocity, free_goal_vel);
      break;
    }
  }
  for (int v : graph_.adj[back]) {
    if (stop_) return;
    if (std::find(visited.begin(), visited.end(), v) != visited.end() || v == goal) continue;
    visited.push_back(v);
    DepthFirst(visited, goal, start_orientation, goal_orientation, start_velocity, free_goal_vel);
    visited.pop_back();
  }
}

/* :92-216 */
void lrKeyPointGraph::createGraph(const PoseSE2& start, const PoseSE2& goal, double dist_to_obst, double obstacle_heading_threshold,
                                  const geometry_msgs::Twist* start_velocity, bool free_goal_vel) {
  clearGraph();
  pending_.clear();
  stop_ = false;
  if (classesFull()) return;
  Eigen::Vector2d diff = goal.position() - start.position();
  if (diff.norm() < cfg_->goal_tolerance.xy_goal_tolerance) {
    /* goal reached up to the tolerance: one short straight band to correct the orientation */
    if (hcp_->getTrajectoryContainer().empty()) hcp_->addAndInitNewTeb(start, goal, start_velocity, free_goal_vel);
    return;
  }
  Eigen::Vector2d normal(-diff[1], diff[0]);
  normal.normalize();
  normal = normal * dist_to_obst;
  const int start_vtx = graph_.addVertex(start.position());
  diff.normalize();
  int nearest_u = -1, nearest_v = -1;
  double min_dist = DBL_MAX;
  const ObstContainer* obstacles = hcp_->obstacles();
  if (obstacles) {
    for (const ObstaclePtr& ob : *obstacles) {
      const Eigen::Vector2d start2obst = ob->getCentroid() - start.position();
      const double dist = start2obst.norm();
      if (start2obst.dot(diff) / dist < 0.1) continue; /* not in front of the start */
      const int u = graph_.addVertex(ob->getCentroid() + normal);
      const int v = graph_.addVertex(ob->getCentroid() - normal);
      if (obstacle_heading_threshold && dist < min_dist) { min_dist = dist; nearest_u = u; nearest_v = v; }
    }
  }
  const int goal_vtx = graph_.addVertex(goal.position());
  const int nv = (int)graph_.pos.size();
  for (int i = 0; i < nv - 1; ++i) { /* the goal has no out edges */
    for (int j = 0; j < nv; ++j) {
      if (i == j) continue;
      Eigen::Vector2d distij = graph_.pos[j] - graph_.pos[i];
      distij.normalize();
      if (distij.dot(diff) <= obstacle_heading_threshold) continue; /* backwards / too far sideways */
      if (obstacle_heading_threshold && i == start_vtx && min_dist != DBL_MAX && (j == nearest_u || j == nearest_v)) {
        Eigen::Vector2d keypoint_dist = graph_.pos[j] - start.position();
        keypoint_dist.normalize();
        const Eigen::Vector2d start_orient_vec(std::cos(start.theta()), std::sin(start.theta()));
        if (start_orient_vec.dot(keypoint_dist) <= obstacle_heading_threshold) continue;
      }
      bool collision = false;
      if (obstacles)
        for (const ObstaclePtr& ob : *obstacles)
          if (ob->checkLineIntersection(graph_.pos[i], graph_.pos[j], 0.5 * dist_to_obst)) { collision = true; break; }
      if (collision) continue;
      graph_.addEdge(i, j);
    }
  }
  std::vector<int> visited(1, start_vtx);
  DepthFirst(visited, goal_vtx, start.theta(), goal.theta(), start_velocity, free_goal_vel);
  flushPaths(start.theta(), goal.theta(), start_velocity, free_goal_vel);
}

double ProbRoadmapGraph::uniform(double lo, double hi) {
  for (;;) {
    const double r = (double)rnd_generator_() / 4294967296.0 * (hi - lo) + lo;
    if (r < hi) return r;
  }
}

/* :220-342 */
void ProbRoadmapGraph::createGraph(const PoseSE2& start, const PoseSE2& goal, double dist_to_obst, double obstacle_heading_threshold,
                                   const geometry_msgs::Twist* start_velocity, bool free_goal_vel) {
  clearGraph();
  pending_.clear();
  stop_ = false;
  if (classesFull()) return;
  Eigen::Vector2d diff = goal.position() - start.position();
  const double start_goal_dist = diff.norm();
  if (start_goal_dist < cfg_->goal_tolerance.xy_goal_tolerance) {
    if (hcp_->getTrajectoryContainer().empty()) hcp_->addAndInitNewTeb(start, goal, start_velocity, free_goal_vel);
    return;
  }
  Eigen::Vector2d normal(-diff[1], diff[0]);
  normal.normalize();
  const double area_width = cfg_->hcp.roadmap_graph_area_width;
  const double len_scale = cfg_->hcp.roadmap_graph_area_length_scale;
  const double x_hi = start_goal_dist * len_scale;
  const double phi = std::atan2(diff[1], diff[0]);
  const double cphi = std::cos(phi), sphi = std::sin(phi);
  Eigen::Vector2d area_origin;
  if (len_scale != 1.0) area_origin = start.position() + 0.5 * (1.0 - len_scale) * start_goal_dist * diff.normalized() - 0.5 * area_width * normal;
  else area_origin = start.position() - 0.5 * area_width * normal;
  const int start_vtx = graph_.addVertex(start.position());
  diff.normalize();
  for (int i = 0; i < cfg_->hcp.roadmap_graph_no_samples; ++i) {
    const double sx = uniform(0, x_hi);
    const double sy = uniform(0, area_width);
    graph_.addVertex(area_origin + Eigen::Vector2d(cphi * sx - sphi * sy, sphi * sx + cphi * sy));
  }
  const int goal_vtx = graph_.addVertex(goal.position());
  const int nv = (int)graph_.pos.size();
  const ObstContainer* obstacles = hcp_->obstacles();
  for (int i = 0; i < nv - 1; ++i) {
    for (int j = 0; j < nv; ++j) {
      if (i == j) continue;
      Eigen::Vector2d distij = graph_.pos[j] - graph_.pos[i];
      distij.normalize();
      if (distij.dot(diff) <= obstacle_heading_threshold) continue;
      bool collision = false;
      if (obstacles)
        for (const ObstaclePtr& ob : *obstacles)
          if (ob->checkLineIntersection(graph_.pos[i], graph_.pos[j], dist_to_obst)) { collision = true; break; }
      if (collision) continue;
      graph_.addEdge(i, j);
    }
  }
  std::vector<int> visited(1, start_vtx);
  DepthFirst(visited, goal_vtx, start.theta(), goal.theta(), start_velocity, free_goal_vel);
  flushPaths(start.theta(), goal.theta(), start_velocity, free_goal_vel);
}

}  // namespace teb_local_planner
