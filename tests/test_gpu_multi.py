"""
Multi-GPU parity (-m gpu; every test skips below 2 visible GPUs): the sharded MixPE sum through
pgx_mix_reduce (peer memory over NVLink, csrc/pgx_comm.cu) against the oracle sum over all streams.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import pygmu2_b200 as pg
import pygmu2_oracle as orc
from pygmu2_b200 import _lib, dist as pd, workloads as wl
from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _need_gpus(n):
    if _lib.device_count() < n:
        pytest.skip(f"needs {n} GPUs, {_lib.device_count()} visible")


def test_mix_reduce_two_devices_in_one_process_matches_oracle_sum():
    """Two ranks inside this process (device 0 and 1, plain peer access): ragged shards, ragged pulls, reduce of
    every pull onto rank 1 (a non-zero root)."""
    _need_gpus(2)
    N, L, B = 11, 1500, 128
    pulls = (128, 300, 77, 512, 263)
    n = sum(pulls)
    comms = [pd.MixComm(d, d, 2, root=1, max_floats=512, connect=False) for d in range(2)]
    pd.MixComm.connect_local(comms)
    banks, xs = [], []
    for r in range(2):
        lo, hi = pd.shard_bounds(N, 2, r)
        banks.append(pg.ConvolveBank(np.stack([wl.c4_ir(s, L) for s in range(lo, hi)]), hi - lo, 1, block=B,
                                     max_pull=512, device=r))
        banks[-1].attach_comm(comms[r])
        xs.append(np.stack([wl.c4_input(n, s) for s in range(lo, hi)])[:, None, :])
    out, pos = [], 0
    for d in pulls:
        ys = [np.full((1, d), np.nan, np.float32) for _ in range(2)]
        tk = [banks[r].submit(np.ascontiguousarray(xs[r][:, :, pos:pos + d]), ys[r], mix=True, reduce=True) for r in range(2)]
        for r in range(2):
            banks[r].wait(tk[r])
        assert np.all(np.isnan(ys[0]))          # only the root delivers
        out.append(ys[1])
        pos += d
    y = np.concatenate(out, axis=1)[0]
    ref = np.zeros(n)
    for s in range(N):
        ref += orc.OracleConvolve(wl.c4_ir(s, L), 1).render(wl.c4_input(n, s)).astype(np.float64)[:, 0]
    assert rel_err(y, ref) <= TOL
    for c in comms:
        c.check()
    for b in banks:
        b.close()


def test_mix_reduce_contract():
    bank = pg.ConvolveBank(np.ones((1, 8), np.float32), 2, 1, block=16, max_pull=64)
    x = np.zeros((2, 1, 16), np.float32)
    with pytest.raises(ValueError, match="no communicator"):
        bank.submit(x, np.zeros((1, 16), np.float32), mix=True, reduce=True)
    comm = pd.MixComm(0, 0, 1, max_floats=8)               # a one-rank world is already connected
    bank.attach_comm(comm)
    with pytest.raises(ValueError, match="PGX_PULL_MIX"):
        bank.submit(x, np.zeros((2, 1, 16), np.float32), mix=False, reduce=True)
    with pytest.raises(ValueError, match="exceed"):
        bank.submit(x, np.zeros((1, 16), np.float32), mix=True, reduce=True)
    with pytest.raises(ValueError):
        pd.MixComm(0, 3, 2)
    bank.close()


@pytest.mark.parametrize("world", [2])
def test_sharded_mix_torchrun(world):
    """One process per GPU under torchrun: pipelined submits with the peer-memory reduce, parity on rank 0."""
    _need_gpus(world)
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "mp_sharded_mix.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "peer-memory reduce err" in res.stdout
