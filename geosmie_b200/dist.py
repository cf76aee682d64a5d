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
        self.device = torch.device("cuda", device) if self.backend == "nccl" else torch.device("cpu")   # device of the control tensors
        self.gpu = device if (device is not None and torch.cuda.is_available()) else None                # GPU index of this rank
        if self.gpu is not None:
            torch.cuda.set_device(self.gpu)
        self._own = False
        self._peer, self._peer_failed = None, False
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

    def peer_gather(self, seg_bytes, nslot=1, handle=None):
        """A PeerGather (rank 0's buffer mapped by every rank through CUDA IPC) with segments of at least `seg_bytes`, or
        None when there is no GPU (gloo tests on CPU), it is switched off (GEOSMIE_GATHER=nccl) or could not be set up on every rank
        (then the NCCL collectives below are used: still a GPU path)."""
        if self.gpu is None or os.environ.get("GEOSMIE_GATHER", "peer") == "nccl" or self._peer_failed:
            return None
        pg = self._peer
        if pg is not None and pg.seg >= seg_bytes and pg.nslot >= nslot:
            return pg
        if pg is not None:
            pg.close()
            self._peer = None
        try:
            self._peer = PeerGather(self, seg_bytes, nslot=nslot, handle=handle)
        except PeerUnavailable as e:
            if self.rank == 0:
                print("geosmie_b200.dist: peer-memory gather unavailable (%s); using the NCCL gather" % e, flush=True)
            self._peer_failed = True
            self._peer = None
        return self._peer

    def gather_rows(self, rows):
        """rows: 2-D float64 array (possibly a different row count per rank) -> on rank 0 the rank-major concatenation,
        elsewhere None.  On GPUs the rows travel over NVLink: written by every rank's copy engine straight into rank 0's
        buffer (PeerGather), else by an NCCL gather of device tensors; gloo on CPU."""
        torch, td = self.torch, self.td
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=self.device)
        counts = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.world)]
        td.all_gather(counts, n)
        counts = [int(c.item()) for c in counts]
        width = rows.shape[1]
        nmax = max(counts)
        pg = self.peer_gather(nmax * width * 8) if nmax * width else None
        if pg is not None:
            stage = torch.from_numpy(rows).to(torch.device("cuda", self.gpu)) if rows.size else None
            if stage is not None:
                pg.put(0, stage.data_ptr(), rows.nbytes)
            pg.complete()                                       # every rank's rows have landed on rank 0
            out = pg.read(0, [c * width * 8 for c in counts]) if self.rank == 0 else None
            td.barrier()                                        # the slot may be overwritten by the next gather only now
            if self.rank != 0:
                return None
            return np.concatenate([o.view(np.float64).reshape(c, width) for o, c in zip(out, counts)], axis=0)
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
        if self._peer is not None:
            self._peer.close()
            self._peer = None
        if self._own and self.td.is_initialized():
            self.td.destroy_process_group()


class PeerUnavailable(RuntimeError):
    pass


class PeerGather(object):
    """Gather over NVLink peer memory (C ABI: gm_peer_*, include/geosmie_b200.h "multi-GPU exchange").

    Rank 0 owns one device buffer laid out [slot][rank][seg bytes]; every other rank maps it into its own address space
    through a CUDA IPC handle (broadcast once through torch.distributed) and writes its segment itself -- with its copy
    engine (`put`, no SM time on either GPU, overlaps the next step's kernels) or directly from the producing kernels
    (`seg_ptr` handed to gm_table_set_mirror / gm_gsf_expand_phase4_dev).  `complete()` = all puts of this rank finished
    + barrier: afterwards the slot is complete on rank 0.  `nslot` > 1 lets step k+1 fill another slot while slot k is read.
    """

    def __init__(self, comm, seg_bytes, nslot=1, handle=None):
        from . import _lib
        self.comm = comm
        self.h = handle or _lib.Handle.get(comm.gpu)
        self.seg = (int(seg_bytes) + 255) // 256 * 256
        self.nslot = int(nslot)
        self.base = None
        torch, td = comm.torch, comm.td
        payload, err = [None], ""
        if comm.rank == 0:
            try:
                self.base, payload[0] = self.h.peer_alloc(self.seg * comm.world * self.nslot)
            except Exception as e:   # noqa: BLE001 -- reported to every rank below
                err = str(e)
        td.broadcast_object_list(payload, src=0)
        if comm.rank != 0 and payload[0] is not None:
            try:
                self.base = self.h.peer_open(payload[0])
            except Exception as e:   # noqa: BLE001
                err = str(e)
        ok = torch.tensor([1 if self.base else 0], dtype=torch.int32, device=comm.device)
        td.all_reduce(ok, op=td.ReduceOp.MIN)
        if int(ok.item()) == 0:
            self.close()
            raise PeerUnavailable(err or "another rank could not map the exchange buffer")

    def seg_ptr(self, slot, offset=0, rank=None):
        """Device pointer (valid in THIS process) of byte `offset` of the segment of `rank` (default: this rank)."""
        r = self.comm.rank if rank is None else rank
        return self.base + (slot * self.comm.world + r) * self.seg + offset

    def put(self, slot, src_ptr, nbytes, offset=0):
        assert offset + nbytes <= self.seg
        self.h.peer_put(self.seg_ptr(slot, offset), src_ptr, nbytes)

    def join(self):
        self.h.peer_join()

    def sync(self):
        self.h.peer_sync()

    def complete(self):
        self.h.sync()          # kernels that store into the buffer directly (mirror mode)
        self.h.peer_sync()     # copy-engine puts
        self.comm.td.barrier()

    def read(self, slot, nbytes_per_rank):
        """Rank 0: host copies (uint8 arrays) of the first nbytes_per_rank[r] bytes of every rank's segment."""
        torch = self.comm.torch
        outs = [torch.empty(max(int(n), 1), dtype=torch.uint8).pin_memory() for n in nbytes_per_rank]
        for r, (o, n) in enumerate(zip(outs, nbytes_per_rank)):
            if n:
                self.h.peer_put(o.data_ptr(), self.seg_ptr(slot, 0, rank=r), int(n))
        self.h.peer_sync()
        return [o.numpy()[:int(n)] for o, n in zip(outs, nbytes_per_rank)]

    def close(self):
        """Collective: the importing ranks unmap the buffer before rank 0 frees it."""
        try:
            self.h.peer_sync()
            if self.base and self.comm.rank != 0:
                self.h.peer_close(self.base)
                self.base = None
            if self.comm.td.is_initialized():
                self.comm.td.barrier()
            if self.base:
                self.h.peer_free(self.base)
        finally:
            self.base = None
