"""TEST INFRASTRUCTURE -- numpy restatement of the planner's candidate preparation (not on the product path).

Follows the reference between the rivals' predictions and the candidate solves:
  * rival ordering            OvertakeTrajPlanner.get_local_traj, car_racing/planning/overtake_traj_planner.py:69-77
  * veh_infos                 :87-92 (row `num` = insertion order, column 1 = max ey over the prediction)
  * max_delta_v               planner_helper.get_agent_info, car_racing/planning/planner_helper.py:177-205
  * Bezier control points     planner_helper.get_bezier_control_points, planner_helper.py:46-136
  * Bezier curve samples      planner_helper.get_bezier_curve, planner_helper.py:139-153; overtake_traj_planner.py:105-111
  * per-stage targets         generate_traj_per_region, overtake_traj_planner.py:329-334 (interp1d of the samples)
Pinned: tests/golden/planner_prep_golden.npz holds what the unmodified reference returns for ten cases
(tests/golden/make_planner_prep_golden.py); tests/test_planner_prepare.py checks this file against it.
Only tests/ may import this module.
"""
import numpy as np


def sort_rivals(eys):
    """Order of `sorted_vehicles` as indices into the insertion order (overtake_traj_planner.py:69-77): a rival goes
    to the front if its ey is >= the current first one's, to the back otherwise -- not a full sort for > 2 rivals."""
    order = []
    for i, ey in enumerate(eys):
        if not order:
            order.append(i)
        elif ey >= eys[order[0]]:
            order.insert(0, i)
        elif ey <= eys[order[0]]:
            order.append(i)
    return order


def interp_lin(xs, ys, x):
    """scipy interp1d(kind='linear') on float64 data (= numpy.interp inside the range); ValueError outside, as
    interp1d's default bounds_error does."""
    if x < xs[0] or x > xs[-1]:
        raise ValueError("A value in x_new is outside the interpolation range.")
    return float(np.interp(x, xs, ys))


def bezier_control_points(xcurv_ego, veh_ey_max, max_delta_v, prediction_factor, track_width, lap_length, veh_width, opt_traj):
    """planner_helper.py:46-136.  veh_ey_max[i] = veh_infos[i, 1] (insertion order -- the reference indexes it with the
    region index all the same); opt_traj (T,2) = columns s, ey of the optimal trajectory.  Returns (num_veh+1, 4, 2)."""
    num_veh = len(veh_ey_max)
    cp = np.zeros((num_veh + 1, 4, 2))
    os_, oe = opt_traj[:, 0], opt_traj[:, 1]
    for index in range(num_veh + 1):
        s0 = xcurv_ego[4]
        s3 = xcurv_ego[4] + prediction_factor * max_delta_v + 4
        if s0 > s3:                                              # :63-79
            s1 = (s3 + lap_length - s0) / 3.0 + s0
            s2 = 2.0 * (s3 + lap_length - s0) / 3.0 + s0
            s3 = s3 + lap_length
        else:                                                    # :81-90
            s1 = (s3 - s0) / 3.0 + s0
            s2 = 2.0 * (s3 - s0) / 3.0 + s0
        # ey0: the look-up is made (and may raise) although the value is overwritten (:92-100)
        if s0 < 0:
            interp_lin(os_, oe, s0 + lap_length)
        elif s0 < os_[0]:
            pass
        else:
            interp_lin(os_, oe, s0)
        ey0 = xcurv_ego[5]
        if index == 0:                                           # :103-110
            ey1 = 0.8 * track_width - (-veh_ey_max[index] - 0.5 * veh_width) * 0.2
        elif index == num_veh:                                   # :112-118
            ey1 = -0.8 * track_width + ((veh_ey_max[index - 1] - 0.5 * veh_width)) * 0.2
        else:                                                    # :119-125
            ey1 = 0.7 * (veh_ey_max[index] + 0.5 * veh_width) + 0.3 * (veh_ey_max[index - 1] - 0.5 * veh_width)
        if s3 >= lap_length:                                     # :127-135
            ey3 = oe[0] if s3 - lap_length <= os_[0] else interp_lin(os_, oe, s3 - lap_length)
        else:
            ey3 = oe[0] if s3 <= os_[0] else interp_lin(os_, oe, s3)
        cp[index, :, 0] = s0, s1, s2, s3
        cp[index, :, 1] = ey0, ey1, ey1, ey3
    return cp


def bezier_curves(cp, N):
    """planner_helper.py:139-153 sampled at t = j (1/N), j = 0..N (overtake_traj_planner.py:105-111)."""
    C = cp.shape[0]
    out = np.zeros((C, N + 1, 2))
    for index in range(C):
        for j in range(N + 1):
            t = j * (1.0 / N)
            for d in range(2):
                p0, p1, p2, p3 = cp[index, :, d]
                out[index, j, d] = p0 * ((1 - t) ** 3) + 3 * p1 * t * ((1 - t) ** 2) + 3 * p2 * (t ** 2) * (1 - t) + p3 * (t ** 3)
    return out


def prepare(ego_x, xcurv_ego, obs_sorted, insertion, rival_vx, prediction_factor, track_width, lap_length, veh_width, opt_traj, N):
    """obs_sorted (num_veh, 2, N+1): s and ey predictions in sorted_vehicles order; insertion[i] = position in
    sorted_vehicles of the rival that was i-th in vehicles_interest; rival_vx in sorted order.
    Returns control points, curve samples and the per-stage targets of every region."""
    num_veh = obs_sorted.shape[0]
    veh_ey_max = np.array([obs_sorted[insertion[i], 1].max() for i in range(num_veh)])
    max_delta_v = max(abs(ego_x[0] - v) for v in rival_vx)
    cp = bezier_control_points(xcurv_ego, veh_ey_max, max_delta_v, prediction_factor, track_width, lap_length, veh_width, opt_traj)
    bez = bezier_curves(cp, N)
    C = num_veh + 1
    s_ref, ey_ref = np.zeros((C, N + 1)), np.zeros((C, N + 1))
    for c in range(C):
        for j in range(N + 1):
            s_tmp = np.clip(ego_x[4] + 1.0 * j * ego_x[0] * 0.1, bez[c, 0, 0], bez[c, -1, 0])
            s_ref[c, j], ey_ref[c, j] = s_tmp, interp_lin(bez[c, :, 0], bez[c, :, 1], s_tmp)
    return dict(ctrl=cp, bezier=bez, s_ref=s_ref, ey_ref=ey_ref, max_delta_v=max_delta_v)
