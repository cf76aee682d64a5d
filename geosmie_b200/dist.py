"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

The table grid shards by independent (wavelength, RH) cells: rank r evaluates cells r, r+W, r+2W, ... of
each bin with no data-path collective; the reduced sums (a few KB per cell) are gathered to rank 0 once per bin.
"""
import os

import numpy as np


def shard(items, rank, world):
    """Round-robin shard of a list of independent work items."""
    return list(items)[rank::world]


def unshard_order(n, world):
    """Position of global item i in the rank-major concatenation of the shards."""
    order = [i for r in range(world) for i in range(n)[r::world]]
    pos = np.empty(n, dtype=np.int64)
    pos[np.array(order, dtype=np.int64)] = np.arange(n)
    return pos


class Comm(object):
    def __init__(self, rank, world, device=None, backend=None):
        import torch
        import torch.distributed as td
        self.torch, self.td = torch, td
        self.rank, self.world = rank, world
        self.backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        self.device = torch.device("cuda", device) if self.backend == "nccl" else torch.device("cpu")
        if self.backend == "nccl":
            torch.cuda.set_device(self.device)
        self._own = False
        if not td.is_initialized():
            td.init_process_group(self.backend, rank=rank, world_size=world)
            self._own = True

    @classmethod
    def from_env(cls, backend=None):
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", str(rank)))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        return cls(rank, world, device=local, backend=backend)

    def barrier(self):
        self.td.barrier()

    def gather_rows(self, rows):
        """rows: 2-D float64 array (possibly a different row count per rank) -> on rank 0 the rank-major concatenation,
        elsewhere None.  NCCL gather of device tensors over NVLink (gloo on CPU)."""
        torch, td = self.torch, self.td
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=self.device)
        counts = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.world)]
        td.all_gather(counts, n)
        counts = [int(c.item()) for c in counts]
        width = rows.shape[1]
        nmax = max(counts)
        buf = torch.zeros((nmax, width), dtype=torch.float64, device=self.device)
        if rows.shape[0]:
            buf[:rows.shape[0]] = torch.from_numpy(rows).to(self.device)
        out = [torch.empty_like(buf) for _ in range(self.world)] if self.rank == 0 else None
        td.gather(buf, out, dst=0)
        if self.rank != 0:
            return None
        return np.concatenate([out[r][:counts[r]].cpu().numpy() for r in range(self.world)], axis=0)

    def gather_cells(self, scal, phase):
        """(scal [nt][nmode][11], phase [nt][4][nang]) per rank -> rank-major concatenation on rank 0."""
        nt = scal.shape[0]
        ws, wp = int(np.prod(scal.shape[1:])), int(np.prod(phase.shape[1:]))
        rows = np.concatenate([scal.reshape(nt, ws), phase.reshape(nt, wp)], axis=1)
        g = self.gather_rows(rows)
        if g is None:
            return None
        return g[:, :ws].reshape((-1,) + scal.shape[1:]), g[:, ws:].reshape((-1,) + phase.shape[1:])

    def close(self):
        if self._own and self.td.is_initialized():
            self.td.destroy_process_group()
