"""CPU test of the N>1 plumbing: world_size-2 gloo run of the cell sharding + gather + reassembly used by
dointegration.fun (the compute is replaced by a deterministic stand-in; no GPU needed)."""
import os
import socket
import subprocess
import sys


from geosmie_b200 import dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
from geosmie_b200 import dist
comm = dist.Comm.from_env(backend="gloo")
cells = [(li, rhi) for li in range(7) for rhi in range(5)]
mine = dist.shard(cells, comm.rank, comm.world)
# stand-in for gm_table_run: values that encode the cell identity
scal = np.array([[[li * 100 + rhi + 0.25 * k for k in range(11)]] for li, rhi in mine], dtype=float).reshape(len(mine), 1, 11)
phase = np.array([[[li + 0.001 * rhi + q] * 13 for q in range(4)] for li, rhi in mine], dtype=float).reshape(len(mine), 4, 13)
out = comm.gather_cells(scal, phase)
if comm.rank == 0:
    pos = dist.unshard_order(len(cells), comm.world)
    s, p = out[0][pos], out[1][pos]
    ok = all(s[i, 0, 0] == li * 100 + rhi and p[i, 2, 5] == li + 0.001 * rhi + 2 for i, (li, rhi) in enumerate(cells))
    print("GATHER_OK" if ok and s.shape == (35, 1, 11) and p.shape == (35, 4, 13) else "GATHER_BAD")
# ragged gather as used by dointegration.fun (finished rows; a rank may own no wavelength at all)
rows = np.full((3, 5), float(comm.rank)) if comm.rank == 0 else np.zeros((0, 5))
g = comm.gather_rows(rows)
if comm.rank == 0:
    print("RAGGED_OK" if g.shape == (3, 5) and np.all(g == 0.0) else "RAGGED_BAD")
comm.close()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_and_unshard_are_inverse():
    for n, w in ((35, 2), (36, 8), (5, 8), (1, 1)):
        order = [i for r in range(w) for i in dist.shard(range(n), r, w)]
        pos = dist.unshard_order(n, w)
        assert [order[p] for p in pos] == list(range(n))


def test_gloo_world2_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0], outs
    assert "RAGGED_OK" in outs[0], outs
