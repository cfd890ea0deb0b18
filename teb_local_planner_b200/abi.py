"""ctypes mirror of include/teb_b200.h (the C-ABI of the B200-native TEB optimizer).

Field order and types must match the header exactly; `check_sizes(lib)` compares against
`tebgpu_sizeof()` exported by the library.
"""
import ctypes as C

import numpy as np

TEBGPU_OK = 0
TEBGPU_ERR_INVALID_ARG = -1
TEBGPU_ERR_CUDA = -2
TEBGPU_ERR_UNSUPPORTED = -3
TEBGPU_ERR_CAPACITY = -4
TEBGPU_ERR_NO_DEVICE = -5

TEB_STATUS_OPTIMIZED = 1
TEB_STATUS_TOO_FEW_POSES = 2
TEB_STATUS_CHOL_FAILED = 4
TEB_STATUS_NONFINITE = 8
TEB_STATUS_TERMINATED = 16
TEB_STATUS_DISABLED = 32
TEB_STATUS_CAPACITY = 64
TEB_STATUS_BAD_INPUT = 128

TEB_FOOTPRINT_POINT, TEB_FOOTPRINT_CIRCULAR, TEB_FOOTPRINT_TWO_CIRCLES, TEB_FOOTPRINT_LINE, TEB_FOOTPRINT_POLYGON = range(5)
TEB_OBST_POINT, TEB_OBST_CIRCULAR, TEB_OBST_LINE, TEB_OBST_PILL, TEB_OBST_POLYGON = range(5)
TEB_MAX_FOOTPRINT_VERTICES = 16
TEB_ROTDIR_NONE, TEB_ROTDIR_LEFT, TEB_ROTDIR_RIGHT = 0, 1, 2

_d, _i = C.c_double, C.c_int32


class TebParams(C.Structure):
    _fields_ = [
        # trajectory
        ("dt_ref", _d), ("dt_hysteresis", _d),
        ("force_reinit_new_goal_dist", _d), ("force_reinit_new_goal_angular", _d),
        ("teb_autosize", _i), ("min_samples", _i), ("max_samples", _i), ("exact_arc_length", _i),
        ("via_points_ordered", _i), ("allow_init_with_backwards_motion", _i),
        ("global_plan_overwrite_orientation", _i), ("_pad0", _i),
        # robot
        ("max_vel_x", _d), ("max_vel_x_backwards", _d), ("max_vel_y", _d), ("max_vel_trans", _d),
        ("max_vel_theta", _d), ("acc_lim_x", _d), ("acc_lim_y", _d), ("acc_lim_theta", _d),
        ("min_turning_radius", _d),
        # footprint
        ("footprint_radius", _d), ("footprint_front_offset", _d), ("footprint_front_radius", _d),
        ("footprint_rear_offset", _d), ("footprint_rear_radius", _d),
        ("footprint_type", _i), ("footprint_vertex_count", _i),
        ("footprint_line", _d * 4), ("footprint_vertices", _d * (2 * 16)),
        # obstacles
        ("min_obstacle_dist", _d), ("inflation_dist", _d), ("dynamic_obstacle_inflation_dist", _d),
        ("obstacle_association_force_inclusion_factor", _d), ("obstacle_association_cutoff_factor", _d),
        ("obstacle_proximity_ratio_max_vel", _d), ("obstacle_proximity_lower_bound", _d),
        ("obstacle_proximity_upper_bound", _d),
        ("include_dynamic_obstacles", _i), ("legacy_obstacle_association", _i),
        ("obstacle_poses_affected", _i), ("_pad2", _i),
        # optim
        ("penalty_epsilon", _d), ("weight_max_vel_x", _d), ("weight_max_vel_y", _d),
        ("weight_max_vel_theta", _d), ("weight_acc_lim_x", _d), ("weight_acc_lim_y", _d),
        ("weight_acc_lim_theta", _d), ("weight_kinematics_nh", _d),
        ("weight_kinematics_forward_drive", _d), ("weight_kinematics_turning_radius", _d),
        ("weight_optimaltime", _d), ("weight_shortest_path", _d), ("weight_obstacle", _d),
        ("weight_inflation", _d), ("weight_dynamic_obstacle", _d),
        ("weight_dynamic_obstacle_inflation", _d), ("weight_velocity_obstacle_ratio", _d),
        ("weight_viapoint", _d), ("weight_prefer_rotdir", _d), ("weight_adapt_factor", _d),
        ("obstacle_cost_exponent", _d),
        ("no_inner_iterations", _i), ("no_outer_iterations", _i), ("optimization_activate", _i), ("_pad3", _i),
        # hcp
        ("selection_cost_hysteresis", _d), ("selection_prefer_initial_plan", _d),
        ("selection_obst_cost_scale", _d), ("selection_viapoint_cost_scale", _d),
        ("selection_alternative_time_cost", _i), ("enable_multithreading", _i),
        ("h_signature_prescaler", _d), ("h_signature_threshold", _d),
        # recovery
        ("divergence_detection_enable", _i), ("_pad4", _i),
        ("divergence_detection_max_chi_squared", _d),
    ]


class TebObstacle(C.Structure):
    _fields_ = [("x", _d), ("y", _d), ("vx", _d), ("vy", _d), ("radius", _d), ("dynamic", _i), ("type", _i),
                ("vertex_begin", _i), ("vertex_count", _i), ("_pad", _d)]


OBST_DTYPE = np.dtype([("x", "f8"), ("y", "f8"), ("vx", "f8"), ("vy", "f8"), ("radius", "f8"),
                       ("dynamic", "i4"), ("type", "i4"), ("vertex_begin", "i4"), ("vertex_count", "i4"), ("_pad", "f8")])
assert OBST_DTYPE.itemsize == C.sizeof(TebObstacle) == 64


class TebBatch(C.Structure):
    _fields_ = [
        ("B", _i), ("n_cap", _i), ("S", _i), ("M_cap", _i), ("V_cap", _i), ("PV_cap", _i),
        ("poses", C.c_void_p), ("n", C.c_void_p), ("scene_id", C.c_void_p),
        ("obstacles", C.c_void_p), ("obst_count", C.c_void_p), ("obst_vertices", C.c_void_p),
        ("via", C.c_void_p), ("via_count", C.c_void_p),
        ("vel_start", C.c_void_p), ("vel_goal", C.c_void_p), ("prefer_rotdir", C.c_void_p),
        ("cost", C.c_void_p), ("chi2", C.c_void_p), ("status", C.c_void_p), ("lm_iters", C.c_void_p),
    ]


class TebOptimizeArgs(C.Structure):
    _fields_ = [("iterations_innerloop", _i), ("iterations_outerloop", _i),
                ("compute_cost_afterwards", _i), ("alternative_time_cost", _i),
                ("obst_cost_scale", _d), ("viapoint_cost_scale", _d)]


class TebGpuLimits(C.Structure):
    _fields_ = [("max_bands", _i), ("max_poses", _i), ("max_scenes", _i), ("max_obstacles", _i),
                ("max_viapoints", _i), ("max_obst_vertices", _i)]


def default_params() -> TebParams:
    """TebConfig() constructor defaults (reference include/teb_local_planner/teb_config.h:245-390).

    Kept in Python as well so host-side tests do not need the CUDA library; the C-ABI's
    tebgpu_default_params() must return the same values (tests compare them).
    """
    p = TebParams()
    p.dt_ref, p.dt_hysteresis = 0.3, 0.1
    p.force_reinit_new_goal_dist, p.force_reinit_new_goal_angular = 1.0, 0.5 * np.pi
    p.teb_autosize, p.min_samples, p.max_samples, p.exact_arc_length = 1, 3, 500, 0
    p.via_points_ordered, p.allow_init_with_backwards_motion, p.global_plan_overwrite_orientation = 0, 0, 1
    p.max_vel_x, p.max_vel_x_backwards, p.max_vel_y, p.max_vel_trans, p.max_vel_theta = 0.4, 0.2, 0.0, 0.0, 0.3
    p.acc_lim_x, p.acc_lim_y, p.acc_lim_theta, p.min_turning_radius = 0.5, 0.5, 0.5, 0.0
    p.footprint_type = TEB_FOOTPRINT_POINT
    p.min_obstacle_dist, p.inflation_dist, p.dynamic_obstacle_inflation_dist = 0.5, 0.6, 0.6
    p.obstacle_association_force_inclusion_factor, p.obstacle_association_cutoff_factor = 1.5, 5.0
    p.obstacle_proximity_ratio_max_vel, p.obstacle_proximity_lower_bound, p.obstacle_proximity_upper_bound = 1.0, 0.0, 0.5
    p.include_dynamic_obstacles, p.legacy_obstacle_association, p.obstacle_poses_affected = 1, 0, 25
    p.penalty_epsilon = 0.05
    p.weight_max_vel_x, p.weight_max_vel_y, p.weight_max_vel_theta = 2.0, 2.0, 1.0
    p.weight_acc_lim_x, p.weight_acc_lim_y, p.weight_acc_lim_theta = 1.0, 1.0, 1.0
    p.weight_kinematics_nh, p.weight_kinematics_forward_drive, p.weight_kinematics_turning_radius = 1000.0, 1.0, 1.0
    p.weight_optimaltime, p.weight_shortest_path = 1.0, 0.0
    p.weight_obstacle, p.weight_inflation = 50.0, 0.1
    p.weight_dynamic_obstacle, p.weight_dynamic_obstacle_inflation = 50.0, 0.1
    p.weight_velocity_obstacle_ratio, p.weight_viapoint, p.weight_prefer_rotdir = 0.0, 1.0, 50.0
    p.weight_adapt_factor, p.obstacle_cost_exponent = 2.0, 1.0
    p.no_inner_iterations, p.no_outer_iterations, p.optimization_activate = 5, 4, 1
    p.selection_cost_hysteresis, p.selection_prefer_initial_plan = 1.0, 0.95
    p.selection_obst_cost_scale, p.selection_viapoint_cost_scale = 100.0, 1.0
    p.selection_alternative_time_cost, p.enable_multithreading = 0, 1
    p.h_signature_prescaler, p.h_signature_threshold = 1.0, 0.1
    p.divergence_detection_enable, p.divergence_detection_max_chi_squared = 0, 10.0
    return p


def ptr(a):
    """Address of a numpy array (or None), as c_void_p value."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


class HostBatch:
    """Owns the numpy arrays behind a host TebBatch (keeps them alive while the struct is in use)."""

    def __init__(self, poses, n, obstacles, obst_count, scene_id=None, via=None, via_count=None,
                 vel_start=None, vel_goal=None, prefer_rotdir=None, obst_vertices=None):
        self.poses = np.ascontiguousarray(poses, dtype=np.float64)
        B, n_cap, four = self.poses.shape
        assert four == 4
        self.n = np.ascontiguousarray(n, dtype=np.int32)
        self.obstacles = np.ascontiguousarray(obstacles, dtype=OBST_DTYPE)
        if self.obstacles.ndim == 1:
            self.obstacles = self.obstacles[None, :]
        S, M_cap = self.obstacles.shape
        self.obst_count = np.ascontiguousarray(obst_count, dtype=np.int32).reshape(S)
        # vertex pool of the Line / Pill / Polygon obstacles: [S][PV_cap][2]
        if obst_vertices is None:
            self.obst_vertices = np.zeros((S, 0, 2), np.float64)
        else:
            self.obst_vertices = np.ascontiguousarray(obst_vertices, dtype=np.float64)
            if self.obst_vertices.ndim == 2:
                self.obst_vertices = self.obst_vertices[None]
        assert self.obst_vertices.shape[0] == S and self.obst_vertices.shape[2] == 2
        self.PV_cap = self.obst_vertices.shape[1]
        self.scene_id = (np.zeros(B, np.int32) if scene_id is None
                         else np.ascontiguousarray(scene_id, dtype=np.int32))
        if via is None:
            self.via = np.zeros((B, 0, 2), np.float64)
            self.via_count = np.zeros(B, np.int32)
        else:
            self.via = np.ascontiguousarray(via, dtype=np.float64)
            self.via_count = np.ascontiguousarray(via_count, dtype=np.int32)
        V_cap = self.via.shape[1]
        z = np.zeros((B, 4), np.float64)
        z[:, 3] = 1.0  # vel_start_.first = vel_goal_.first = true (optimal_planner.cpp:94-102)
        self.vel_start = z.copy() if vel_start is None else np.ascontiguousarray(vel_start, dtype=np.float64)
        self.vel_goal = z.copy() if vel_goal is None else np.ascontiguousarray(vel_goal, dtype=np.float64)
        self.prefer_rotdir = (np.zeros(B, np.int32) if prefer_rotdir is None
                              else np.ascontiguousarray(prefer_rotdir, dtype=np.int32))
        self.cost = np.full(B, np.inf)
        self.chi2 = np.zeros(B)
        self.status = np.zeros(B, np.int32)
        self.lm_iters = np.zeros(B, np.int32)
        self.B, self.n_cap, self.S, self.M_cap, self.V_cap = B, n_cap, S, M_cap, V_cap

    def struct(self) -> TebBatch:
        b = TebBatch()
        b.B, b.n_cap, b.S, b.M_cap, b.V_cap = self.B, self.n_cap, self.S, self.M_cap, self.V_cap
        b.poses, b.n, b.scene_id = ptr(self.poses), ptr(self.n), ptr(self.scene_id)
        b.obstacles, b.obst_count = ptr(self.obstacles), ptr(self.obst_count)
        b.PV_cap = self.PV_cap
        b.obst_vertices = ptr(self.obst_vertices) if self.PV_cap > 0 else None
        b.via = ptr(self.via) if self.V_cap > 0 else None
        b.via_count = ptr(self.via_count)
        b.vel_start, b.vel_goal, b.prefer_rotdir = ptr(self.vel_start), ptr(self.vel_goal), ptr(self.prefer_rotdir)
        b.cost, b.chi2, b.status, b.lm_iters = ptr(self.cost), ptr(self.chi2), ptr(self.status), ptr(self.lm_iters)
        return b

    def copy(self):
        h = HostBatch(self.poses.copy(), self.n.copy(), self.obstacles.copy(), self.obst_count.copy(),
                      self.scene_id.copy(), self.via.copy() if self.V_cap > 0 else None,
                      self.via_count.copy() if self.V_cap > 0 else None,
                      self.vel_start.copy(), self.vel_goal.copy(), self.prefer_rotdir.copy(),
                      self.obst_vertices.copy() if self.PV_cap > 0 else None)
        return h


def make_args(inner=5, outer=4, compute_cost=True, obst_scale=1.0, via_scale=1.0, alt_time=False) -> TebOptimizeArgs:
    a = TebOptimizeArgs()
    a.iterations_innerloop, a.iterations_outerloop = inner, outer
    a.compute_cost_afterwards, a.alternative_time_cost = int(compute_cost), int(alt_time)
    a.obst_cost_scale, a.viapoint_cost_scale = obst_scale, via_scale
    return a
