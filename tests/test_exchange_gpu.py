"""-m gpu: the exchange window (b200mpc_comm_*, csrc/exchange.cuh) between real processes -- CUDA IPC mapping of the peers'
windows, records stored from the solver kernels' epilogue, arrival counters, argmin kernel, acknowledgements.  Two ranks;
with one GPU they share it (IPC works across processes on one device), with two or more each rank has its own and the
free-running variant (no barrier between the ranks for 40 uses of 2 slots) runs as well."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(world, extra=()):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "helpers", "exchange_worker.py"), *extra], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "EXCHANGE_OK" in outs[0], outs[0]


def test_two_ranks_exchange_records_through_peer_windows():
    _run(2)


def test_free_running_ranks_flow_control():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs one GPU per rank (a spinning kernel does not yield a shared GPU)")
    _run(2, ("--free-running",))
