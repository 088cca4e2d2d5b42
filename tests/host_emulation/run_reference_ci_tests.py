"""The reference's OWN CI test functions, imported from /root/reference/tests/auto_*.py and called UNMODIFIED, with the drop-in installed:

    test_tracking          (auto_control_test.py:   PID, then MPC-LTI, 20 s each)
    test_racing            (auto_mpccbf_test.py:    MPC-CBF past 3 rivals, 40 s)
    test_racing_overtake   (auto_racing_game_test.py: PID lap, MPC-LTI lap, LMPC lap, LMPC + overtaking planner lap with 2 rivals)

Those functions assert nothing themselves: they pass when they run to the end.  Here they additionally leave a summary of what
happened (laps, lap times, where the ego ended relative to the rivals, smallest clearance).  The only edits are outside the test functions:
`car_racing_b200.install_all()`, empty stubs for the modules that are not installed (casadi, cvxopt, pathos, matplotlib -- none is called any
more once the drop-in is in) and no-op plotting / animation methods on the simulator (matplotlib, out of scope).

Build container only (needs /root/reference).  TEST INFRASTRUCTURE: without a GPU the shims call the host-compiled copy of the library
(tests/host_emulation/build_emu_library.py); `--lib product` on a GPU box.

    python tests/host_emulation/run_reference_ci_tests.py --tests test_tracking test_racing test_racing_overtake --out profiles/r04b_reference_ci_tests.json
"""
import argparse
import builtins
import contextlib
import importlib.util
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
FILES = {"test_tracking": "auto_control_test.py", "test_racing": "auto_mpccbf_test.py", "test_racing_overtake": "auto_racing_game_test.py"}


def _clearance(ego_log, car_log, lap):
    """Over the steps both vehicles logged (0.4 x 0.2 cars, base.py:700): the smallest box distance max(|ds| - length, |dey| - width), and the
    smallest value of the reference's own degree-6 shape (ds / 0.4)^6 + (dey / 0.2)^6 - 1 (control.py:527-557 without margin and slack; >= 0 = no contact
    in the reference's sense -- the shape's corners are rounded, so it can be positive where the box distance is slightly negative)."""
    n = min(len(ego_log), len(car_log))
    box, shape = np.inf, np.inf
    for a, b in zip(ego_log[-n:], car_log[-n:]):
        ds = (a[4] - b[4] + lap / 2) % lap - lap / 2
        box = min(box, max(abs(ds) - 0.4, abs(a[5] - b[5]) - 0.2))
        shape = min(shape, (ds / 0.4) ** 6 + ((a[5] - b[5]) / 0.2) ** 6 - 1.0)
    return float(box), float(shape)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tests", nargs="+", default=list(FILES))
    ap.add_argument("--lib", choices=["emu", "product", "auto"], default="auto")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import dropin_sim
    import car_racing_b200 as crb
    from car_racing_b200 import _capi, batch
    lib = args.lib
    if lib == "auto":
        import torch
        lib = "product" if torch.cuda.is_available() else "emu"
    if lib == "emu":
        spec = importlib.util.spec_from_file_location("build_emu_library", os.path.join(HERE, "build_emu_library.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _capi.LIB_PATH, _capi._lib, batch._default_handle = mod.build(), None, None
    control, offboard, base, racing_env = dropin_sim.import_reference()
    from planning.overtake_traj_planner import OvertakeTrajPlanner
    crb.install_all(control, base, OvertakeTrajPlanner, offboard)
    for name in ("plot_simulation", "plot_state", "plot_input", "animate"):          # matplotlib: out of scope
        setattr(offboard.CarRacingSim, name, lambda self, *a, **k: None)
    sims = []
    init = offboard.CarRacingSim.__init__

    def recording_init(self, *a, **k):
        init(self, *a, **k)
        sims.append(self)
    offboard.CarRacingSim.__init__ = recording_init
    out = {"library": lib, "tests": {}}
    real_open = builtins.open
    scratch_dir = tempfile.mkdtemp(prefix="b200mpc_ref_ci_")
    for test in args.tests:
        spec = importlib.util.spec_from_file_location("ref_" + test, os.path.join(dropin_sim.REF, "tests", FILES[test]))
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        del sims[:]
        buf = io.StringIO()
        written = {}

        def guarded_open(file, mode="r", *a, **k):          # the tests pickle the simulator into data/simulator/ of the reference checkout, which
            if isinstance(file, str) and any(c in mode for c in "wax+") and (not os.path.isabs(file) or file.startswith(dropin_sim.REF)):
                dst = os.path.join(scratch_dir, os.path.basename(file))               # must stay untouched: written to a scratch directory instead
                written[file] = dst
                return real_open(dst, mode, *a, **k)
            return real_open(file, mode, *a, **k)
        builtins.open = guarded_open
        t0 = time.time()
        try:
            with contextlib.redirect_stdout(buf):
                getattr(module, test)()                                               # the reference's test function, as it is
        finally:
            builtins.open = real_open
        wall = time.time() - t0
        sim = sims[-1]
        ego = sim.vehicles["ego"]
        lap = sim.track.lap_length
        text = buf.getvalue()
        r = {"file": "tests/" + FILES[test], "passed": True, "wall_s": wall, "ego_laps": int(ego.laps), "ego_time_s": float(ego.time),
             "ego_s_final": float(ego.xcurv[4]), "ego_vx_final": float(ego.xcurv[0]), "ego_ey_final": float(ego.xcurv[5]),
             "lap_times_printed_by_the_test": [ln for ln in text.splitlines() if ln.startswith("lap time")],
             "non_convergence_messages": text.count("solver fail"),
             "files_the_test_pickled": {k: os.path.getsize(v) for k, v in written.items()}}
        ego_log = [x for lp in ego.xcurvs for x in lp] + list(ego.lap_xcurvs)            # completed laps + the lap in progress (base.py:76-93)
        r["max_abs_ey"] = float(max(abs(x[5]) for x in ego_log)) if ego_log else None
        r["track_half_width"] = float(sim.track.width)
        rivals = {}
        for name, car in sim.vehicles.items():
            if name == "ego":
                continue
            ds = float((ego.xcurv[4] - car.xcurv[4] + lap / 2) % lap - lap / 2)      # position on the track relative to the rival (> 0: ahead)
            car_log = [x for lp in car.xcurvs for x in lp] + list(car.lap_xcurvs)
            rivals[name] = {"ego_ahead_of_rival_on_track_at_end_m": ds,
                            "ego_total_s_minus_rival_total_s_at_end_m": float(ego.xcurv[4] + ego.laps * lap - car.xcurv[4]),   # rivals' s is not lap-wrapped
 "min_box_clearance_m": _clearance(ego_log, car_log, lap)[0] if car_log else None,
                            "min_degree6_shape_minus_1": _clearance(ego_log, car_log, lap)[1] if car_log else None}
        r["rivals"] = rivals
        out["tests"][test] = r
        print("[%s] ran to the end in %.0f s: laps %d, ego s %.2f vx %.2f, max|ey| %s, rivals %s, non-convergence messages %d" % (
            test, wall, ego.laps, ego.xcurv[4], ego.xcurv[0], r["max_abs_ey"], json.dumps(rivals), r["non_convergence_messages"]), flush=True)
        for ln in r["lap_times_printed_by_the_test"]:
            print("   ", ln)
        if args.out:
            with open(os.path.join(ROOT, args.out) if not os.path.isabs(args.out) else args.out, "w") as f:
                json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
