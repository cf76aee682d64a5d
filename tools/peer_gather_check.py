#!/usr/bin/env python3
"""Multi-GPU check of the peer-memory gather (run under torchrun, >= 2 GPUs):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_gather_check.py

Every rank evaluates its own shard of optics_BC cells (different refractive indices per rank), then the results reach
rank 0 three ways: (a) copy-engine puts into rank 0's IPC-mapped buffer (gm_peer_put), (b) P2P stores of k_finalize /
k_gsf through the mapped pointers (gm_table_set_mirror), (c) an NCCL gather of the same device tensors.  (a) and (b) must
equal (c) bit for bit, for both slots of the double buffer.  Prints PEER_GATHER_OK / PEER_GATHER_FAIL on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as td

from geosmie_b200 import _lib, dist, workloads

comm = dist.Comm.from_env()
rank, world, dev = comm.rank, comm.world, comm.device
h = _lib.Handle.get(dev.index)
NANG = 371
ang = np.concatenate([np.linspace(0., 1., 100, endpoint=False), np.linspace(1., 10., 100, endpoint=False),
                      np.linspace(10., 180., 171, endpoint=True)])
cells = [(li, ri) for li in range(rank, 61, world) for ri in range(0, 36, 6)][:40]
plan = workloads.bin_plan("bc", 0, cells=cells)
ncell = len(plan.cells)
ncell_t = torch.tensor([ncell], device=dev)
td.all_reduce(ncell_t, op=td.ReduceOp.MIN)
ncell = int(ncell_t.item())
table = _lib.Table(plan.xx, plan.nmax, np.cos(np.radians(ang)), h)
mz_np, wp_np, _, tpc = plan.tasks()
mz = torch.from_numpy(np.ascontiguousarray(mz_np[:ncell]).view(np.float64).reshape(ncell, 2).copy()).to(dev)
w = torch.from_numpy(np.ascontiguousarray(wp_np[:ncell])).to(dev)
nscal, nph, nco = ncell * _lib.GM_NSCAL, ncell * 4 * NANG, ncell * 6 * 129
scal = torch.zeros(nscal, dtype=torch.float64, device=dev)
phase = torch.zeros(nph, dtype=torch.float64, device=dev)
coef = torch.zeros(nco, dtype=torch.float64, device=dev)
cn = torch.zeros(ncell, dtype=torch.float64, device=dev)

pg = comm.peer_gather((nscal + nph + nco) * 8, nslot=2, handle=h)
ok = True
if pg is None:
    if rank == 0:
        print("peer gather unavailable on this box")
    ok = False
else:
    # reference: NCCL gather of the locally written results
    table.run_dev(ncell, mz.data_ptr(), mz.data_ptr(), 1, w.data_ptr(), 0, scal.data_ptr(), phase.data_ptr())
    h.gsf_expand_phase4_dev(ang, ncell, phase.data_ptr(), coef.data_ptr(), cn.data_ptr())
    h.sync()
    local = torch.cat([scal, phase, coef])
    ref = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
    td.gather(local, ref, dst=0)
    for slot in (0, 1):
        for mode in ("put", "store"):
            # wipe this rank's segment (through the mapping), then fill it again
            zero = torch.zeros_like(local)
            pg.put(slot, zero.data_ptr(), local.numel() * 8)
            pg.complete()
            if mode == "put":
                table.set_mirror(None, None)
                table.run_dev(ncell, mz.data_ptr(), mz.data_ptr(), 1, w.data_ptr(), 0, scal.data_ptr(), phase.data_ptr())
                h.gsf_expand_phase4_dev(ang, ncell, phase.data_ptr(), coef.data_ptr(), cn.data_ptr())
                pg.put(slot, scal.data_ptr(), nscal * 8, 0)
                pg.put(slot, phase.data_ptr(), nph * 8, nscal * 8)
                pg.put(slot, coef.data_ptr(), nco * 8, (nscal + nph) * 8)
            else:
                table.set_mirror(pg.seg_ptr(slot, 0), pg.seg_ptr(slot, nscal * 8))
                table.run_dev(ncell, mz.data_ptr(), mz.data_ptr(), 1, w.data_ptr(), 0, scal.data_ptr(), phase.data_ptr())
                h.gsf_expand_phase4_dev(ang, ncell, phase.data_ptr(), pg.seg_ptr(slot, (nscal + nph) * 8), cn.data_ptr())
                table.set_mirror(None, None)
            pg.complete()
            if rank == 0:
                got = pg.read(slot, [local.numel() * 8] * world)
                for r in range(world):
                    a, b = got[r].view(np.float64), ref[r].cpu().numpy()
                    same = np.array_equal(a, b) and np.abs(b).max() > 0
                    distinct = r == 0 or not np.array_equal(b, ref[0].cpu().numpy())
                    ok &= bool(same and distinct)
                    print("slot %d %-5s rank %d: %s (%d doubles, max|v| %.3e)" % (slot, mode, r, "identical" if same else "DIFFERENT",
                                                                                 a.size, np.abs(b).max()), flush=True)
            td.barrier()
table.close()
comm.barrier()
comm.close()
if rank == 0:
    print("PEER_GATHER_OK" if ok else "PEER_GATHER_FAIL")
    sys.exit(0 if ok else 1)
