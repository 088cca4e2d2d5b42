"""car_racing_b200 -- B200-native batched MPC solver behind the controller API of
HybridRobotics/car-racing.  Importing the package does not touch the GPU;
the first solve loads car_racing_b200/libb200mpc.so (hand-written sm_100a kernels behind a
C-ABI, include/b200mpc.h) and fails loudly if it is missing -- there is no CPU fallback.
"""
from . import rivals, scenarios  # noqa: F401
from ._capi import B200MPCError, Handle, default_options  # noqa: F401
from .batch import (CbfPipeline, IlqrPipeline, LmpcPipeline, pack_cbf, pack_ilqr, pack_lmpc, solve_cbf_batch, solve_cbf_packed, solve_ilqr_batch, solve_lmpc_batch,
                    estimate_abc_batch, plant_step_batch, rival_rollout_batch, curv_to_glob_batch)  # noqa: F401
from .control import estimate_ABC, ilqr, install, install_all, lmpc, mpc_lti, mpc_multi_agents, mpccbf, pid  # noqa: F401

__version__ = "0.1.0"
