"""Named feature scenarios (parameters + a small batch) shared by the oracle-pinning tests and the GPU parity tests.
Each one switches on a branch of the reference's hot path (SURVEY.md par. 8a rows)."""
import numpy as np

from teb_local_planner_b200 import abi, scenes


def _base(cfg, candidates, seed, autosize=False):
    return scenes.make_config_batch(cfg, candidates=candidates, seed=seed, autosize=autosize)


def scenario(name, candidates=3):
    """-> (TebParams, HostBatch). Scene 0 holds the obstacles of every band."""
    if name in ("C1", "C2", "C3", "C4"):
        return _base(name, candidates, 21)
    if name in ("C1_autosize", "C2_autosize", "C3_autosize", "C4_autosize"):
        return _base(name[:2], candidates, 7, autosize=True)
    if name == "all_edge_families":   # two circles, cost exponent, exact arc length, shortest path, rotdir, car-like, velocities
        p, hb = _base("C3", 4, 5)
        p.footprint_type = abi.TEB_FOOTPRINT_TWO_CIRCLES
        p.footprint_front_offset, p.footprint_front_radius = 0.3, 0.2
        p.footprint_rear_offset, p.footprint_rear_radius = 0.2, 0.25
        p.obstacle_cost_exponent = 2.0
        p.exact_arc_length = 1
        p.weight_shortest_path = 0.5
        hb.prefer_rotdir[:] = [abi.TEB_ROTDIR_LEFT, abi.TEB_ROTDIR_RIGHT, 0, abi.TEB_ROTDIR_LEFT]
        hb.vel_start[:, 0], hb.vel_start[:, 2] = 0.3, -0.1
        hb.vel_goal[1, 3] = 0.0        # free_goal_vel for band 1
        hb.vel_goal[2, 0] = 0.2
        return p, hb
    if name in ("legacy", "vor", "legacy_vor"):
        p, hb = _base("C3", candidates, 4)
        p.legacy_obstacle_association = int(name.startswith("legacy"))
        p.obstacle_poses_affected = 10
        p.weight_velocity_obstacle_ratio = 2.0 if name.endswith("vor") else 0.0
        p.obstacle_proximity_lower_bound, p.obstacle_proximity_upper_bound = 0.2, 1.0
        return p, hb
    if name in ("holonomic", "holonomic_no_acc_y", "holonomic_trans"):
        p, hb = _base("C3", candidates, 8)
        p.max_vel_y = 0.3
        p.acc_lim_y = 0.0 if name == "holonomic_no_acc_y" else 0.5
        p.max_vel_trans = 0.45 if name == "holonomic_trans" else 0.0
        p.weight_kinematics_nh = 1.0
        p.weight_max_vel_y, p.weight_acc_lim_y = 2.0, 1.5
        rng = np.random.default_rng(3)
        for b in range(hb.B):
            n = hb.n[b]
            hb.poses[b, 1:n - 1, 2] += rng.normal(0, 0.35, n - 2)
            hb.poses[b, :n - 1, 3] *= 0.8
        hb.vel_start[:, 0], hb.vel_start[:, 1], hb.vel_start[:, 2] = 0.25, 0.1, -0.1
        hb.vel_goal[1, 3] = 0.0
        hb.vel_goal[2 % hb.B, 1] = 0.15
        return p, hb
    if name.startswith("shapes_"):     # Line / Pill / Polygon obstacles, static and moving, with a footprint model
        footprint = name[len("shapes_"):]
        p, hb = _base("C4", candidates, 6)
        hb = scenes.add_shape_obstacles(hb, seed=1)
        hb.obstacles["dynamic"][0, ::3] = 0
        if footprint == "line":
            scenes.set_line_footprint(p)
        elif footprint == "polygon":
            scenes.set_polygon_footprint(p)
        elif footprint == "circular":
            p.footprint_type, p.footprint_radius = abi.TEB_FOOTPRINT_CIRCULAR, 0.2
        elif footprint == "two_circles":
            p.footprint_type = abi.TEB_FOOTPRINT_TWO_CIRCLES
            p.footprint_front_offset, p.footprint_front_radius, p.footprint_rear_offset, p.footprint_rear_radius = 0.3, 0.15, 0.2, 0.2
        elif footprint == "legacy_polygon":
            scenes.set_polygon_footprint(p)
            p.legacy_obstacle_association, p.obstacle_poses_affected = 1, 8
        return p, hb
    if name == "via_ordered":          # via_points_ordered (optimal_planner.cpp:690-705)
        p, hb = _base("C4", candidates, 9)
        p.via_points_ordered = 1
        hb.via[:, :, 1] += 0.4         # off the band, so that the closest-pose search matters
        return p, hb
    if name == "via_unordered_static":  # via-points without the dynamic-obstacle edges
        p, hb = _base("C4", candidates, 9)
        p.include_dynamic_obstacles = 0
        hb.obstacles["dynamic"][...] = 0
        return p, hb
    if name == "divergence":           # recovery.divergence_detection_enable: batch statistics refresh the cached errors
        p, hb = _base("C2", candidates, 11)
        p.divergence_detection_enable = 1
        p.divergence_detection_max_chi_squared = 10
        return p, hb
    raise KeyError(name)


ALL = ["C1", "C2", "C3", "C4", "C1_autosize", "C2_autosize", "C3_autosize", "C4_autosize", "all_edge_families", "legacy", "vor",
       "legacy_vor", "holonomic", "holonomic_no_acc_y", "holonomic_trans", "shapes_point", "shapes_line", "shapes_polygon",
       "shapes_two_circles", "shapes_circular", "shapes_legacy_polygon", "via_ordered", "via_unordered_static", "divergence"]


def band_kwargs(hb, b):
    """keyword arguments of the single-band oracle / reference entry points for band b (scene of the band)"""
    s = int(hb.scene_id[b])
    return dict(obstacles=hb.obstacles[s][:hb.obst_count[s]], via=hb.via[b][:hb.via_count[b]] if hb.V_cap else None,
                vel_start=hb.vel_start[b], vel_goal=hb.vel_goal[b], rotdir=int(hb.prefer_rotdir[b]),
                obst_vertices=hb.obst_vertices[s] if hb.PV_cap > 0 else None)
