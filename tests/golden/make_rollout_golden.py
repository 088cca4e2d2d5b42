"""Generates tests/golden/rollout_golden.npz with the UNMODIFIED reference prediction of a dynamic rival:
offboard.DynamicBicycleModel.get_trajectory_nsteps / get_estimation (car_racing/racing/offboard.py:51-94), the zero-input
kinematic rollout the MPC-CBF controller (control.py:505-507, realtime_flag) and the planner (overtake_traj_planner.py:84-86)
ask a rival with dynamics for.  Build container only.

    python tests/golden/make_rollout_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = "/root/reference"
from make_ilqr_golden import import_reference_control   # noqa: E402


def main():
    import_reference_control()
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        from racing import offboard
    finally:
        os.chdir(cwd)
    from utils import racing_env
    rng = np.random.default_rng(11)
    out = {}
    for name in ["ellipse", "l_shape", "goggle"]:
        spec = np.genfromtxt(os.path.join(REF, "data/track_layout/%s.csv" % name), delimiter=",")
        track = racing_env.ClosedTrack(spec, 1.0)
        pat = np.asarray(track.point_and_tangent, float)
        Bn = 12
        for n in (11, 21):
            xc0, xg0 = np.zeros((Bn, 6)), np.zeros((Bn, 6))
            xc, xg = np.zeros((Bn, 6, n)), np.zeros((Bn, 6, n))
            for b in range(Bn):
                s0 = rng.uniform(0.2, track.lap_length - 0.2) if b < Bn - 3 else track.lap_length - rng.uniform(0.05, 0.6)   # last 3 wrap
                xc0[b] = [rng.uniform(0.5, 1.8), rng.uniform(-.1, .1), rng.uniform(-.4, .4), rng.uniform(-.2, .2), s0, rng.uniform(-.6, .6)]
                xg0[b] = [xc0[b, 0], xc0[b, 1], xc0[b, 2], rng.uniform(-3, 3), rng.uniform(-5, 5), rng.uniform(-5, 5)]
                me = types.SimpleNamespace(xcurv=xc0[b].copy(), xglob=xg0[b].copy(), lap_length=track.lap_length,
                                           point_and_tangent=track.point_and_tangent, timestep=0.1)
                me.get_estimation = types.MethodType(offboard.DynamicBicycleModel.get_estimation, me)
                xc[b], xg[b] = offboard.DynamicBicycleModel.get_trajectory_nsteps(me, n)
            out.update({"%s/n%d/xcurv0" % (name, n): xc0, "%s/n%d/xglob0" % (name, n): xg0, "%s/n%d/xcurv" % (name, n): xc,
                        "%s/n%d/xglob" % (name, n): xg})
        out["%s/pat" % name] = pat
        out["%s/lap_length" % name] = np.array(track.lap_length)
        print(name, "lap", track.lap_length, "wrapped rows:", int((xc[:, 4, -1] < xc[:, 4, 0]).sum()), flush=True)
    path = os.path.join(HERE, "rollout_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
