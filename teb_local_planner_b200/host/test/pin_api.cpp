/*
 * pin_api.cpp — TEST INFRASTRUCTURE: the band operations of the drop-in layer (TimedElasticBand, TebOptimalPlanner's
 * velocity / trajectory read-outs) behind one C entry point with the same operation codes as oracle/ref_driver.cpp's
 * teb_ref_band_op, so that tests/test_host_pin.py can compare the drop-in layer with the REFERENCE's own code
 * (oracle/_ref/libteb_ref.so), operation by operation, bit for bit. Nothing here touches a GPU.
 */
#include <cmath>
#include <cstdint>
#include <vector>

#include "teb_local_planner/optimal_planner.h"
#include "teb_local_planner/timed_elastic_band.h"

using namespace teb_local_planner;

namespace {
void load_band(TimedElasticBand& teb, const double* rec, int n) {
  for (int i = 0; i < n; ++i) {
    teb.addPose(rec[4 * i], rec[4 * i + 1], rec[4 * i + 2], i == 0 || i == n - 1);
    if (i < n - 1) teb.addTimeDiff(rec[4 * i + 3]);
  }
}
int put_band(const TimedElasticBand& teb, double* out, int cap) {
  const int n = teb.sizePoses();
  if (4 * n > cap) return -1;
  for (int i = 0; i < n; ++i) {
    out[4 * i] = teb.Pose(i).x(); out[4 * i + 1] = teb.Pose(i).y(); out[4 * i + 2] = teb.Pose(i).theta();
    out[4 * i + 3] = i < teb.sizeTimeDiffs() ? teb.TimeDiff(i) : 0.0;
  }
  return 4 * n;
}
void twist_from(const double* v, geometry_msgs::Twist& t) { t.linear.x = v[0]; t.linear.y = v[1]; t.angular.z = v[2]; }
}  // namespace


/* op 12: isTrajectoryFeasible against a costmap stand-in that RECORDS every footprint query and reports a collision
 * (-1) inside any of the given discs. args: inscribed_radius circumscribed_radius look_ahead_idx lookahead_distance
 * min_resolution_collision_check_angular n_discs (x y r)* -> out: feasible, number of queries, the queried (x, y, theta). */
struct RecordingCostmap : public base_local_planner::CostmapModel {
  std::vector<double> q;
  const double* discs = nullptr;
  int nd = 0;
  double footprintCost(double x, double y, double th, const std::vector<geometry_msgs::Point>&, double = 0.0, double = 0.0) override {
    q.push_back(x); q.push_back(y); q.push_back(th);
    for (int k = 0; k < nd; ++k)
      if (std::hypot(x - discs[3 * k], y - discs[3 * k + 1]) <= discs[3 * k + 2]) return -1;
    return 0;
  }
};

extern "C" int32_t teb_host_band_op(int32_t op, const double* rec, int32_t n, const double* a, int32_t na, double* out, int32_t cap) {
  (void)na;
  TimedElasticBand teb;
  if (rec && n > 0) load_band(teb, rec, n);
  switch (op) {
    case 1:
      teb.initTrajectoryToGoal(PoseSE2(a[0], a[1], a[2]), PoseSE2(a[3], a[4], a[5]), a[6], a[7], (int)a[8], a[9] != 0);
      return put_band(teb, out, cap);
    case 2: {
      std::vector<geometry_msgs::PoseStamped> plan((size_t)a[5]);
      for (size_t i = 0; i < plan.size(); ++i) {
        plan[i].pose.position.x = a[6 + 3 * i]; plan[i].pose.position.y = a[7 + 3 * i];
        plan[i].pose.orientation = tf::createQuaternionFromYaw(a[8 + 3 * i]);
      }
      teb.initTrajectoryToGoal(plan, a[0], a[1], a[2] != 0, (int)a[3], a[4] != 0);
      return put_band(teb, out, cap);
    }
    case 3: {
      std::vector<Eigen::Vector2d> path((size_t)a[8]);
      for (size_t i = 0; i < path.size(); ++i) path[i] = Eigen::Vector2d(a[9 + 2 * i], a[10 + 2 * i]);
      auto opt = [&](int k) -> const double* { return std::isnan(a[k]) ? nullptr : &a[k]; };
      teb.initTrajectoryToGoal(path, a[0], a[1], opt(2), opt(3), opt(4), opt(5), (int)a[6], a[7] != 0);
      return put_band(teb, out, cap);
    }
    case 4: {
      PoseSE2 s(a[0], a[1], a[2]), g(a[3], a[4], a[5]);
      teb.updateAndPruneTEB(s, g, (int)a[6]);
      return put_band(teb, out, cap);
    }
    case 5: {
      double dist = -1;
      out[0] = teb.findClosestTrajectoryPose(Eigen::Vector2d(a[0], a[1]), &dist, (int)a[2]);
      out[1] = dist;
      return 2;
    }
    case 6:
      out[0] = teb.getSumOfAllTimeDiffs(); out[1] = teb.getAccumulatedDistance(); out[2] = teb.getSumOfTimeDiffsUpToIdx((int)a[0]);
      return 3;
    case 7:
      out[0] = teb.isTrajectoryInsideRegion(a[0], a[1], (int)a[2]);
      return 1;
    case 11:
      teb.autoResize(a[0], a[1], (int)a[2], (int)a[3], a[4] != 0);
      return put_band(teb, out, cap);
    default: break;
  }
  TebConfig cfg;
  cfg.robot.max_vel_y = (op == 8) ? a[1] : a[0];
  ObstContainer obst;
  TebOptimalPlanner pl(cfg, &obst);
  load_band(pl.teb(), rec, n);
  if (op == 13) { /* plan() sequences with the optimisation switched off (optimization_activate = false): cold start,
                     warm start (updateAndPruneTEB) and re-initialisation after a goal jump (optimal_planner.cpp:233-321).
                     args: kind (0 poses / 1 plans) ncalls reinit_dist reinit_ang max_vel_x max_vel_theta min_samples
                     backwards overwrite_orientation, then per call s[3] g[3] or np pts[np][3]; out: per call n, band */
    TebConfig cfg13;
    cfg13.optim.optimization_activate = false;
    cfg13.trajectory.force_reinit_new_goal_dist = a[2];
    cfg13.trajectory.force_reinit_new_goal_angular = a[3];
    cfg13.robot.max_vel_x = a[4];
    cfg13.robot.max_vel_theta = a[5];
    cfg13.trajectory.min_samples = (int)a[6];
    cfg13.trajectory.allow_init_with_backwards_motion = a[7] != 0;
    cfg13.trajectory.global_plan_overwrite_orientation = a[8] != 0;
    ObstContainer obst13;
    TebOptimalPlanner pl13(cfg13, &obst13);
    int r = 9, w = 0;
    for (int c = 0; c < (int)a[1]; ++c) {
      if (a[0] == 0) {
        pl13.plan(PoseSE2(a[r], a[r + 1], a[r + 2]), PoseSE2(a[r + 3], a[r + 4], a[r + 5]), nullptr, false);
        r += 6;
      } else {
        const int np = (int)a[r++];
        std::vector<geometry_msgs::PoseStamped> plan((size_t)np);
        for (int i = 0; i < np; ++i, r += 3) {
          plan[i].pose.position.x = a[r]; plan[i].pose.position.y = a[r + 1];
          plan[i].pose.orientation = tf::createQuaternionFromYaw(a[r + 2]);
        }
        pl13.plan(plan, nullptr, false);
      }
      const int nn = pl13.teb().sizePoses();
      if (w + 1 + 4 * nn > cap) return -1;
      out[w++] = nn;
      const int k = put_band(pl13.teb(), out + w, cap - w);
      w += k;
    }
    return w;
  }
  if (op == 12) {
    RecordingCostmap cm;
    cm.discs = a + 6; cm.nd = (int)a[5];
    TebConfig cfg12;
    cfg12.trajectory.min_resolution_collision_check_angular = a[4];
    ObstContainer obst12;
    TebOptimalPlanner pl12(cfg12, &obst12);
    load_band(pl12.teb(), rec, n);
    std::vector<geometry_msgs::Point> footprint;
    out[0] = pl12.isTrajectoryFeasible(&cm, footprint, a[0], a[1], (int)a[2], a[3]);
    out[1] = (double)(cm.q.size() / 3);
    if (2 + (int)cm.q.size() > cap) return -1;
    for (size_t k = 0; k < cm.q.size(); ++k) out[2 + k] = cm.q[k];
    return 2 + (int)cm.q.size();
  }
  if (op == 8) {
    double vx = 0, vy = 0, om = 0;
    out[0] = pl.getVelocityCommand(vx, vy, om, (int)a[0]);
    out[1] = vx; out[2] = vy; out[3] = om;
    return 4;
  }
  geometry_msgs::Twist ts, tg;
  twist_from(a + 1, ts); twist_from(a + 5, tg);
  if (a[4] != 0) pl.setVelocityStart(ts);
  if (a[8] != 0) pl.setVelocityGoal(tg); else pl.setVelocityGoalFree();
  if (op == 9) {
    std::vector<geometry_msgs::Twist> prof;
    pl.getVelocityProfile(prof);
    if (3 * (int)prof.size() > cap) return -1;
    for (size_t i = 0; i < prof.size(); ++i) { out[3 * i] = prof[i].linear.x; out[3 * i + 1] = prof[i].linear.y; out[3 * i + 2] = prof[i].angular.z; }
    return 3 * (int)prof.size();
  }
  if (op == 10) {
    std::vector<TrajectoryPointMsg> tr;
    pl.getFullTrajectory(tr);
    if (7 * (int)tr.size() > cap) return -1;
    for (size_t i = 0; i < tr.size(); ++i) {
      double* o = out + 7 * i;
      o[0] = tr[i].pose.position.x; o[1] = tr[i].pose.position.y; o[2] = tf::getYaw(tr[i].pose.orientation);
      o[3] = tr[i].velocity.linear.x; o[4] = tr[i].velocity.linear.y; o[5] = tr[i].velocity.angular.z; o[6] = tr[i].time_from_start;
    }
    return 7 * (int)tr.size();
  }
  return -2;
}
