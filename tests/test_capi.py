"""CPU tests: the C-ABI library loads and exports every symbol include/b200mpc.h declares; packers."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib_or_skip():
    from car_racing_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _capi, _capi.lib()


def test_library_exports_every_declared_symbol():
    _capi, L = _lib_or_skip()
    hdr = open(os.path.join(ROOT, "include", "b200mpc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b200mpc_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/b200mpc.h but not exported"
    assert L.b200mpc_version() == 100


def test_struct_layouts_and_record_sizes():
    _capi, L = _lib_or_skip()
    from car_racing_b200 import batch
    assert C.sizeof(_capi.CbfParams) == 16 + 8 * (36 + 12 + 36 + 4 + 2 + 8)
    assert C.sizeof(_capi.IpmOptions) == 8 + 4 + 4 + 8 * 6 + 4 + 4
    for N in (1, 10, 20, 50, 64):
        for M in range(_capi.MMAX + 1):
            for ps in (0, 1):
                assert L.b200mpc_cbf_record_doubles(N, M, ps) == batch.cbf_record_doubles(N, M, ps)
                for fl in (3, 4):
                    assert L.b200mpc_cbf_record_doubles_ex(N, M, ps, fl) == batch.cbf_record_doubles(N, M, ps, fl)
        assert L.b200mpc_ilqr_record_doubles(N) == batch.ilqr_record_doubles(N)
    assert C.sizeof(_capi.PlantParams) == 16 + 12 * 8
    assert _capi.make_plant_params(9, 25.4).n_sub == 100
    assert C.sizeof(_capi.SysidParams) == 4 * 4 + 4 * 4 + 8 + 3 * 8
    assert C.sizeof(_capi.LmpcParams) == 8 + 8 * (36 + 4 + 4 + 6 + 2 + 2)
    for N, K in ((2, 1), (12, 44), (16, 64), (10, 33)):
        assert L.b200mpc_lmpc_record_doubles(N, K) == batch.lmpc_record_doubles(N, K)
    assert L.b200mpc_lmpc_record_doubles(17, 44) < 0 and L.b200mpc_lmpc_record_doubles(12, 65) < 0
    assert batch.cbf_record_doubles(20, 3, 0) * 8 == 1136      # algorithmic input bytes
    assert L.b200mpc_cbf_record_doubles(65, 0, 0) < 0 and L.b200mpc_cbf_record_doubles(20, _capi.MMAX + 1, 0) < 0
    o = _capi.default_options()
    assert o.tol == 1e-8 and o.max_iter == 200 and o.rho == 1e3 and o.start == _capi.START_ROLLOUT and o.max_reset == 5


def test_no_device_is_a_loud_error():
    _capi, L = _lib_or_skip()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_capi.B200MPCError, match="no CPU fallback"):
        _capi.Handle().ptr


def test_pack_cbf_layout():
    from car_racing_b200 import _capi, batch
    N, M, B = 4, 2, 3
    rng = np.random.default_rng(0)
    x0 = rng.normal(size=(B, 6)); obs = rng.normal(size=(B, M, 2, N + 1)); lo = rng.normal(size=(B, M))
    xt = rng.normal(size=6)
    rec, m, ps = batch.pack_cbf(x0, xt, obs, lo, N)
    assert (m, ps) == (M, False) and rec.shape == (B, batch.cbf_record_doubles(N, M, 0))
    assert (rec[:, :6] == x0).all() and (rec[:, 6:8] == lo).all() and (rec[:, 8:14] == xt).all()
    assert (rec[1, 14:14 + N + 1] == obs[1, 0, 0]).all() and (rec[1, 14 + 3 * (N + 1):14 + 4 * (N + 1)] == obs[1, 1, 1]).all()
    xts = rng.normal(size=(B, N + 1, 6))
    rec, m, ps = batch.pack_cbf(x0, xts, obs, None, N)
    assert ps and (rec[2, 8:8 + 6 * (N + 1)] == xts[2].ravel()).all() and (rec[:, 6:8] == 0).all()
    rec, m, ps = batch.pack_cbf(x0, xt, np.zeros((B, 0, 2, N + 1)), None, N)
    assert m == 0 and rec.shape[1] == 12
    with pytest.raises(ValueError):
        batch.pack_cbf(x0, xt, np.zeros((B, _capi.MMAX + 1, 2, N + 1)), None, N)
    # per-rival sizes: appended block, flag RIVAL_SIZE; only wd given: the planner blocks travel together
    sz = np.array([[0.5, 0.2], [0.4, 0.3]])
    rec, m, ps = batch.pack_cbf(x0, xt, obs, lo, N, sizes=sz)
    assert rec.shape[1] == batch.cbf_record_doubles(N, M, 0, _capi.FLAG_RIVAL_SIZE) and (rec[:, -4:] == sz.ravel()).all()
    rec, m, ps = batch.pack_cbf(x0, xt, np.zeros((B, 0, 2, N + 1)), None, N, wd=np.ones((B, N)))
    assert batch.cbf_flags(0, wd=1) == 3 and rec.shape[1] == batch.cbf_record_doubles(N, 0, 0, 3)
    with pytest.raises(ValueError):
        batch.pack_cbf(x0, xt, obs, lo, N, xlb=np.zeros((B, N + 1, 2)))


def test_handle_pickles_without_native_state():
    import pickle
    from car_racing_b200 import _capi
    h = _capi.Handle(device=0, max_batch=7)
    h2 = pickle.loads(pickle.dumps(h))
    assert h2.device == 0 and h2.max_batch == 7 and h2._h is None
