"""Why the barrier value dips below 0 while the ego passes car1 in the reference's MPC-CBF scenario (car_racing/tests/mpccbf_test.py, zero noise):
per step, the number of rivals M that pass the reference's proximity filter (control.py:499-523), solver status / iterations, realised against
predicted next state, the realised barrier value and the largest slack.  Output of one run: profiles/r04b_mpccbf_filter_flicker.txt.
Build container only (needs /root/reference); TEST INFRASTRUCTURE (the shims call the host-compiled library)."""
import sys, os, json, importlib.util
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
import dropin_sim
import car_racing_b200 as crb
from car_racing_b200 import _capi, batch
spec = importlib.util.spec_from_file_location("b", os.path.join(HERE, "build_emu_library.py")); mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
_capi.LIB_PATH, _capi._lib, batch._default_handle = mod.build(), None, None
control, offboard, base, racing_env = dropin_sim.import_reference()
crb.install(control)
shim = control.mpccbf
log = []
def logged(*a, **kw):
    u, r = shim(*a, return_details=True, **kw)
    log.append(r)
    return u
control.mpccbf = logged
rivals = [(4.0, 0.2, 0.1), (10.0, 0.2, -0.1)]
sim, ego, track = dropin_sim.build_sim(offboard, base, racing_env, "mpccbf", rivals)
ego.set_zero_noise()
prev = None
for k in range(140):
    x_before = np.array(ego.xcurv, float).copy()
    for name in sim.vehicles:
        sim.vehicles[name].forward_one_step(sim.vehicles[name].realtime_flag)
    x = np.array(ego.xcurv, float)
    r = log[-1]
    t = (k + 1) * 0.1
    s1, e1 = 4.0 + 0.2 * t, 0.1
    ds, de = x[4] - s1, x[5] - e1
    h = (ds / 0.4) ** 6 + (de / 0.2) ** 6 - 1.2
    xp = r["x"][0, 1]
    if abs(ds) < 0.8:
        sg = r.get("sigma")
        print("k=%3d M=%d st=%d it=%3d  real s,ey=(%.4f,%.4f) pred=(%.4f,%.4f) d=(%.4f,%.4f)  h_real=%.3f  sigma_max=%s" % (
            k, 0 if sg is None else np.shape(sg)[1] if np.ndim(sg) > 1 else 0, r["status"][0], r["iters"][0], x[4], x[5], xp[4], xp[5], x[4]-xp[4], x[5]-xp[5], h, None if sg is None or np.size(sg) == 0 else float(np.max(sg))))
