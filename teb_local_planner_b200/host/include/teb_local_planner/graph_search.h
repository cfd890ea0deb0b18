/*
 * graph_search.h — exploration of candidate paths in distinct homotopy classes (host side).
 *
 * Mirrors include/teb_local_planner/graph_search.h / src/graph_search.cpp of the reference: lrKeyPointGraph (:92-216, two
 * key points left / right of every obstacle), ProbRoadmapGraph (:220-342, uniformly sampled vertices in a rectangle
 * between start and goal) and the depth-first enumeration of all forward paths (:45-88). Differences in structure, not in
 * behaviour: the graph is a plain adjacency list (no boost::graph), and the enumerated paths are handed to the planner in
 * chunks so that their H-signatures come from ONE device call per chunk (HomotopyClassPlanner::addAndInitNewTebs); the
 * planner accepts them in enumeration order and the search stops as soon as max_number_classes is reached, which is what
 * the reference's per-path test does.
 */
#ifndef TEB_B200_GRAPH_SEARCH_H_
#define TEB_B200_GRAPH_SEARCH_H_

#include <cstdint>
#include <random>
#include <vector>

#include "teb_local_planner/obstacles.h"
#include "teb_local_planner/teb_config.h"

namespace teb_local_planner {

class HomotopyClassPlanner;

struct HcGraph {
  std::vector<Eigen::Vector2d> pos;        /* vertex positions, index = vertex descriptor */
  std::vector<std::vector<int>> adj;       /* out edges in insertion order (boost::adjacency_list<listS, vecS, directedS>) */
  int addVertex(const Eigen::Vector2d& p) { pos.push_back(p); adj.emplace_back(); return (int)pos.size() - 1; }
  void addEdge(int u, int v) { adj[u].push_back(v); }
  void clear() { pos.clear(); adj.clear(); }
};

class GraphSearchInterface {
 public:
  virtual ~GraphSearchInterface() {}
  virtual void createGraph(const PoseSE2& start, const PoseSE2& goal, double dist_to_obst, double obstacle_heading_threshold,
                           const geometry_msgs::Twist* start_velocity, bool free_goal_vel = false) = 0;
  void clearGraph() { graph_.clear(); }
  HcGraph graph_;

 protected:
  GraphSearchInterface(const TebConfig& cfg, HomotopyClassPlanner* hcp) : cfg_(&cfg), hcp_(hcp) {}
  /* graph_search.cpp:45-88; paths are flushed to the planner every CHUNK complete paths */
  void DepthFirst(std::vector<int>& visited, int goal, double start_orientation, double goal_orientation,
                  const geometry_msgs::Twist* start_velocity, bool free_goal_vel);
  void flushPaths(double start_orientation, double goal_orientation, const geometry_msgs::Twist* start_velocity, bool free_goal_vel);
  bool classesFull() const;
  /* what both graph builders share:
   *   beginSearch     resets the search state; true when nothing is left to explore (class budget spent, or start and goal
   *                   closer than xy_goal_tolerance: then at most one straight band is created to fix the orientation)
   *   connectForward  directed edge u -> v for every ordered pair (the goal, the last vertex, has no out edges) whose
   *                   direction agrees with `travel_dir` by more than `heading_threshold`, that `veto` does not reject and
   *                   whose segment keeps `clearance` from every obstacle
   *   enumerate       depth-first enumeration of all start -> goal paths + hand-over of the last partial chunk */
  bool beginSearch(const PoseSE2& start, const PoseSE2& goal, const geometry_msgs::Twist* start_velocity, bool free_goal_vel);
  template <typename Veto>
  void connectForward(const Eigen::Vector2d& travel_dir, double heading_threshold, double clearance, Veto veto);
  void enumerate(int start_vertex, int goal_vertex, const PoseSE2& start, const PoseSE2& goal,
                 const geometry_msgs::Twist* start_velocity, bool free_goal_vel);
  const TebConfig* cfg_;
  HomotopyClassPlanner* hcp_;
  std::vector<std::vector<Eigen::Vector2d>> pending_;
  bool stop_ = false;
  static constexpr int CHUNK = 16;
};

class lrKeyPointGraph : public GraphSearchInterface {
 public:
  lrKeyPointGraph(const TebConfig& cfg, HomotopyClassPlanner* hcp) : GraphSearchInterface(cfg, hcp) {}
  void createGraph(const PoseSE2& start, const PoseSE2& goal, double dist_to_obst, double obstacle_heading_threshold,
                   const geometry_msgs::Twist* start_velocity, bool free_goal_vel = false) override;
};

class ProbRoadmapGraph : public GraphSearchInterface {
 public:
  ProbRoadmapGraph(const TebConfig& cfg, HomotopyClassPlanner* hcp) : GraphSearchInterface(cfg, hcp) {}
  void createGraph(const PoseSE2& start, const PoseSE2& goal, double dist_to_obst, double obstacle_heading_threshold,
                   const geometry_msgs::Twist* start_velocity, bool free_goal_vel = false) override;

 private:
  /* boost::random::mt19937 (default seed 5489) + boost::random::uniform_real_distribution<double>: one 32-bit draw per
   * sample, value = draw / 2^32 * (max - min) + min, redrawn if it reaches max. A roadmap sample consumes its y draw
   * before its x draw (GCC's right-to-left argument evaluation in graph_search.cpp:274; pinned in tests/test_reference_pin.py) */
  double uniform(double lo, double hi);
  std::mt19937 rnd_generator_;
};

}  // namespace teb_local_planner
#endif
