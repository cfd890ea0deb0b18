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

bool GraphSearchInterface::beginSearch(const PoseSE2& start, const PoseSE2& goal, const geometry_msgs::Twist* start_velocity,
                                       bool free_goal_vel) {
  clearGraph();
  pending_.clear();
  stop_ = false;
  if (classesFull()) return true;
  if ((goal.position() - start.position()).norm() < cfg_->goal_tolerance.xy_goal_tolerance) {
    if (hcp_->getTrajectoryContainer().empty()) hcp_->addAndInitNewTeb(start, goal, start_velocity, free_goal_vel);
    return true;
  }
  return false;
}

template <typename Veto>
void GraphSearchInterface::connectForward(const Eigen::Vector2d& travel_dir, double heading_threshold, double clearance, Veto veto) {
  const ObstContainer* obstacles = hcp_->obstacles();
  const int vertices = (int)graph_.pos.size();
  for (int u = 0; u + 1 < vertices; ++u) {
    for (int v = 0; v < vertices; ++v) {
      if (u == v) continue;
      const Eigen::Vector2d dir = (graph_.pos[v] - graph_.pos[u]).normalized();
      if (dir.dot(travel_dir) <= heading_threshold) continue; /* backwards, or too far sideways */
      if (veto(u, v)) continue;
      bool blocked = false;
      if (obstacles)
        for (const ObstaclePtr& ob : *obstacles)
          if (ob->checkLineIntersection(graph_.pos[u], graph_.pos[v], clearance)) { blocked = true; break; }
      if (!blocked) graph_.addEdge(u, v);
    }
  }
}

void GraphSearchInterface::enumerate(int start_vertex, int goal_vertex, const PoseSE2& start, const PoseSE2& goal,
                                     const geometry_msgs::Twist* start_velocity, bool free_goal_vel) {
  std::vector<int> visited(1, start_vertex);
  DepthFirst(visited, goal_vertex, start.theta(), goal.theta(), start_velocity, free_goal_vel);
  flushPaths(start.theta(), goal.theta(), start_velocity, free_goal_vel);
}

/* Key-point graph (reference :92-216): two vertices per obstacle in front of the start, dist_to_obst to the left and to
 * the right of its centroid (relative to the start -> goal line), then forward edges that clear the obstacles by half
 * that distance. With a heading threshold, the two key points of the obstacle nearest to the start are only reachable
 * from the start if they also lie ahead of the robot's current heading. */
void lrKeyPointGraph::createGraph(const PoseSE2& start, const PoseSE2& goal, double dist_to_obst, double obstacle_heading_threshold,
                                  const geometry_msgs::Twist* start_velocity, bool free_goal_vel) {
  if (beginSearch(start, goal, start_velocity, free_goal_vel)) return;
  const Eigen::Vector2d line = goal.position() - start.position();
  const Eigen::Vector2d travel_dir = line.normalized();
  const Eigen::Vector2d side = Eigen::Vector2d(-line[1], line[0]).normalized() * dist_to_obst;
  const int start_vertex = graph_.addVertex(start.position());
  int near_left = -1, near_right = -1;
  double near_dist = DBL_MAX;
  if (const ObstContainer* obstacles = hcp_->obstacles()) {
    for (const ObstaclePtr& ob : *obstacles) {
      const Eigen::Vector2d to_obst = ob->getCentroid() - start.position();
      const double range = to_obst.norm();
      if (to_obst.dot(travel_dir) / range < 0.1) continue; /* beside or behind the start: no key points */
      const int left = graph_.addVertex(ob->getCentroid() + side);
      const int right = graph_.addVertex(ob->getCentroid() - side);
      if (obstacle_heading_threshold && range < near_dist) { near_dist = range; near_left = left; near_right = right; }
    }
  }
  const int goal_vertex = graph_.addVertex(goal.position());
  const Eigen::Vector2d heading(std::cos(start.theta()), std::sin(start.theta()));
  const bool guard_nearest = obstacle_heading_threshold && near_dist != DBL_MAX;
  connectForward(travel_dir, obstacle_heading_threshold, 0.5 * dist_to_obst, [&](int u, int v) {
    if (!guard_nearest || u != start_vertex || (v != near_left && v != near_right)) return false;
    const Eigen::Vector2d to_key = (graph_.pos[v] - start.position()).normalized();
    return heading.dot(to_key) <= obstacle_heading_threshold;
  });
  enumerate(start_vertex, goal_vertex, start, goal, start_velocity, free_goal_vel);
}

/* boost::random::mt19937 (default seed) + uniform_real_distribution<double>: one 32-bit draw per sample, scaled to
 * [lo, hi), redrawn in the (rounding) case that it reaches hi */
double ProbRoadmapGraph::uniform(double lo, double hi) {
  double r;
  do {
    r = (double)rnd_generator_() / 4294967296.0 * (hi - lo) + lo;
  } while (!(r < hi));
  return r;
}

/* Probabilistic roadmap (reference :220-342): roadmap_graph_no_samples vertices drawn uniformly in a rectangle of width
 * roadmap_graph_area_width around the start -> goal line (its length scaled by roadmap_graph_area_length_scale about the
 * midpoint), then forward edges that clear the obstacles by dist_to_obst. */
void ProbRoadmapGraph::createGraph(const PoseSE2& start, const PoseSE2& goal, double dist_to_obst, double obstacle_heading_threshold,
                                   const geometry_msgs::Twist* start_velocity, bool free_goal_vel) {
  if (beginSearch(start, goal, start_velocity, free_goal_vel)) return;
  const Eigen::Vector2d line = goal.position() - start.position();
  const double span = line.norm();
  const Eigen::Vector2d travel_dir = line.normalized();
  const Eigen::Vector2d across = Eigen::Vector2d(-line[1], line[0]).normalized();
  const double width = cfg_->hcp.roadmap_graph_area_width, stretch = cfg_->hcp.roadmap_graph_area_length_scale;
  const double length = span * stretch;
  const double angle = std::atan2(line[1], line[0]);
  const double ca = std::cos(angle), sa = std::sin(angle);
  Eigen::Vector2d corner = start.position() - 0.5 * width * across; /* rectangle corner: sample (0, 0) */
  if (stretch != 1.0) corner = start.position() + 0.5 * (1.0 - stretch) * span * travel_dir - 0.5 * width * across;
  const int start_vertex = graph_.addVertex(start.position());
  for (int k = 0; k < cfg_->hcp.roadmap_graph_no_samples; ++k) {
    /* the reference builds the sample as Vector2d(distribution_x(gen), distribution_y(gen)) (graph_search.cpp:274); the
     * evaluation order of the two arguments is unspecified and GCC - what the reference is built with - goes right to
     * left: the y sample comes out of the generator FIRST (pinned against the reference compiled with the same compiler,
     * tests/test_reference_pin.py) */
    const double aside = uniform(0, width);
    const double along = uniform(0, length);
    graph_.addVertex(corner + Eigen::Vector2d(ca * along - sa * aside, sa * along + ca * aside));
  }
  const int goal_vertex = graph_.addVertex(goal.position());
  connectForward(travel_dir, obstacle_heading_threshold, dist_to_obst, [](int, int) { return false; });
  enumerate(start_vertex, goal_vertex, start, goal, start_velocity, free_goal_vel);
}

}  // namespace teb_local_planner
