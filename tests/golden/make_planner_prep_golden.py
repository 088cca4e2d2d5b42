"""Generates tests/golden/planner_prep_golden.npz: what the UNMODIFIED reference computes between the rivals' predictions
and the candidate solves (build container only; the .npz travels).

`OvertakeTrajPlanner.get_local_traj` (car_racing/planning/overtake_traj_planner.py:44-161) is run as it stands, imported
from /root/reference with empty stubs for casadi / matplotlib / cvxopt / pathos (none of them is touched on this path):
rival ordering (:69-77), `veh_infos` (:87-92), `get_agent_info` and `get_bezier_control_points`
(planning/planner_helper.py:177-205, 46-136), the sampled Bezier curves and their interp1d (:105-117).  Its
`solve_optimization_problem` is replaced by a probe that records the prepared planner state -- `sorted_vehicles`,
`obs_infos`, `bezier_xcurvs`, `bezier_funcs` evaluated at the candidates' target abscissae -- and returns a dummy result;
`get_traj_xglob` only feeds plots and gets a track whose `get_global_position` returns zeros.

    python tests/golden/make_planner_prep_golden.py
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    _stub("casadi")
    solvers = _stub("cvxopt.solvers", qp=None)
    _stub("cvxopt", matrix=None, spmatrix=None, solvers=solvers)
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.animation"]:
        _stub(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    _stub("pathos.multiprocessing", ProcessingPool=None)
    _stub("pathos")
    sys.path.insert(0, os.path.join(REF, "car_racing"))
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        from planning import overtake_traj_planner as ref_planner
        from utils import base as ref_base
    finally:
        os.chdir(cwd)
    return ref_planner, ref_base


class Rival:
    """Duck-typed NoDynamicsModel: constant speed along s, constant or drifting ey."""
    no_dynamics = True

    def __init__(self, s0, ey0, v, ey_rate):
        self.param = types.SimpleNamespace(length=0.4, width=0.2)
        self.s0, self.ey0, self.v, self.ey_rate = s0, ey0, v, ey_rate
        self.xcurv = np.array([v, 0.0, 0.0, 0.0, s0, ey0])

    def get_trajectory_nsteps(self, t0, delta_t, n):
        traj = np.zeros((6, n))
        k = np.arange(n)
        traj[0] = self.v
        traj[4] = self.s0 + self.v * delta_t * k
        traj[5] = self.ey0 + self.ey_rate * delta_t * k
        return traj, None


def case(rng, num_veh, lap, s_ego=None, order=None):
    vx = rng.uniform(1.0, 1.6)
    s = rng.uniform(2.0, lap - 7.0) if s_ego is None else s_ego
    ego_x = np.array([vx, rng.uniform(-.02, .02), rng.uniform(-.05, .05), rng.uniform(-.03, .03), s, rng.uniform(-.3, .3)])
    eys = rng.uniform(-0.6, 0.6, size=num_veh)
    if order == "ascending":
        eys = np.sort(eys)
    rivals = {}
    for j in range(num_veh):
        rivals["car%d" % (j + 1)] = Rival(s + rng.uniform(0.3, 1.2), eys[j], rng.uniform(0.7, 1.3), rng.uniform(-0.3, 0.3))
    return ego_x, rivals


def main():
    ref_planner, ref_base = import_reference()
    lap = 19.131304718436152                      # goggle, SURVEY 8(d)
    opt = np.genfromtxt(os.path.join(REF, "data/optimal_traj/xcurv_goggle.csv"), delimiter=",")
    A = np.genfromtxt(os.path.join(REF, "data/sys/LTI/matrix_A.csv"), delimiter=",")
    B = np.genfromtxt(os.path.join(REF, "data/sys/LTI/matrix_B.csv"), delimiter=",")
    track = types.SimpleNamespace(lap_length=lap, width=1.0, get_global_position=lambda s, ey: (0.0, 0.0))
    rng = np.random.default_rng(0)
    specs = [(1, None, None), (2, None, None), (2, None, "ascending"), (3, None, None), (3, None, None), (3, None, "ascending"),
             (2, lap - 3.0, None),      # s3 beyond the start line: ey3 looked up one lap back
             (2, lap - 4.3, None),      # s3 - lap below the first abscissa of the optimal trajectory
             (1, 0.05, None),           # s0 below the first abscissa of the optimal trajectory
             (4, None, None)]
    store = {"opt_traj": opt[:, 4:6].copy(), "lap_length": np.array(lap), "num_cases": np.array(len(specs))}
    for ci, (num_veh, s_ego, order) in enumerate(specs):
        ego_x, rivals = case(rng, num_veh, lap, s_ego, order)
        prm = ref_base.RacingGameParam(matrix_A=A, matrix_B=B, timestep=0.1)
        pl = ref_planner.OvertakeTrajPlanner(prm)
        vehicles = {"ego": types.SimpleNamespace(param=types.SimpleNamespace(length=0.4, width=0.2), xcurv=ego_x.copy())}
        vehicles.update(rivals)
        pl.vehicles, pl.agent_name, pl.track, pl.opti_traj_xcurv = vehicles, "ego", track, opt
        N = prm.num_horizon_planner
        seen = {}

        def probe(self=pl, seen=seen, N=N, ego_x=ego_x):
            C = len(self.sorted_vehicles) + 1
            seen["sorted"] = list(self.sorted_vehicles)
            seen["obs"] = np.array([self.obs_infos[n] for n in self.sorted_vehicles])
            seen["bezier_xcurvs"] = self.bezier_xcurvs.copy()
            ey_ref = np.zeros((C, N + 1))
            s_ref = np.zeros((C, N + 1))
            for c in range(C):      # the lookups generate_traj_per_region does (:329-334)
                for j in range(N + 1):
                    s_tmp = np.clip(ego_x[4] + 1.0 * j * ego_x[0] * 0.1, self.bezier_xcurvs[c, 0, 0], self.bezier_xcurvs[c, -1, 0])
                    s_ref[c, j], ey_ref[c, j] = s_tmp, float(self.bezier_funcs[c](s_tmp))
            seen["s_ref"], seen["ey_ref"] = s_ref, ey_ref
            return np.zeros((N + 1, 6)), 0, np.zeros(C), np.zeros((C, 6, N + 1))
        pl.solve_optimization_problem = probe
        # control points as the reference computes them (get_local_traj does not keep them)
        names = list(rivals)
        pl.get_local_traj(ego_x.copy(), 0.0, dict(rivals), None, None, None, None, None)
        veh_infos = np.zeros((num_veh, 3))
        for num, name in enumerate(names):
            tr, _ = rivals[name].get_trajectory_nsteps(0.0, 0.1, N + 1)
            veh_infos[num] = rivals[name].xcurv[4], tr[5].max(), tr[5].min()
        info = ref_planner.get_agent_info(vehicles, seen["sorted"], track)
        ctrl = ref_planner.get_bezier_control_points(dict(rivals), veh_infos, info, prm, track, opt, seen["sorted"], ego_x.copy())
        d = dict(ego_x=ego_x, num_veh=np.array(num_veh), insertion=np.array([seen["sorted"].index(n) for n in names]),
                 rival_vx=np.array([rivals[n].xcurv[0] for n in seen["sorted"]]), obs=seen["obs"][:, 4:6, :],
                 max_delta_v=np.array(info.max_delta_v), ctrl=ctrl, bezier_xcurvs=seen["bezier_xcurvs"], s_ref=seen["s_ref"],
                 ey_ref=seen["ey_ref"], prediction_factor=np.array(prm.planning_prediction_factor))
        for k, v in d.items():
            store["case%d/%s" % (ci, k)] = np.asarray(v)
        print("case", ci, "num_veh", num_veh, "sorted", seen["sorted"], "insertion->sorted", d["insertion"].tolist(),
              "s0 %.3f s3 %.3f" % (ctrl[0, 0, 0], ctrl[0, 3, 0]), flush=True)
    out = os.path.join(HERE, "planner_prep_golden.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, len(store), "arrays,", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
