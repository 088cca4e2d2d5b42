"""Plant step: the numpy restatement against golden vectors produced by the reference's own
DynamicBicycleModel.forward_dynamics (tests/golden/make_plant_golden.py)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plant_golden.npz")


def test_plant_restatement_matches_reference_golden():
    import plant_numpy
    g = np.load(GOLD)
    assert plant_numpy.substeps(0.1) == 100
    for name in ("ellipse", "l_shape"):
        xc, xg, us, dr = g["xcurv_" + name], g["xglob_" + name], g["u_" + name], g["draws_" + name]
        for k in range(us.shape[1]):
            nc, ng = plant_numpy.plant_step(xc[:, k], xg[:, k], us[:, k], dr[:, k], g["dyn"], g["pat_" + name],
                                            float(g["lap_length_" + name]))
            assert np.abs(nc - xc[:, k + 1]).max() < 1e-12 and np.abs(ng - xg[:, k + 1]).max() < 1e-12
