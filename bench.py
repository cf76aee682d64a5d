#!/usr/bin/env python3
"""bench.py -- headline benchmark of the GEOSmie Mie lookup-table hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): the sulfate table optics_SU -- su.json's lognormal bin on its 4459-point size grid
x 61 wavelengths x 36 RH = 2196 cells = 9,791,964 Mie particle-evaluations per step (371 angles each), evaluated DENSE
(zero-weight particles included, i.e. reference-equivalent work), followed by the GSF moment expansion of all cells.
Refractive indices are the OPAC sulfate / HITRAN water values recorded in tests/golden/hostlogic.npz.

One "step" = one pass of the hot path over that batch:
  value  particle-evals/s with inputs (m, weights) resident in HBM (device-pointer C ABI, CUDA-event timed);
  e2e    the same through the host-buffer C ABI the table driver uses (gm_table_run_psd: per-cell refractive indices and
         size-distribution parameters in pinned host memory -> H2D -> kernels -> D2H of the reduced sums and GSF moments
         into pinned host memory, all inside the timed region);
  roofline   FP64 tensor (DMMA) roofline of the dominant kernel (k_gram on optics_SU), duration from CUDA events recorded
             around each of its launches on the launching stream inside the timed steps; the other kernels of the step
             are listed under roofline.kernels with their own bounds;
  cpu_baseline   the CPU oracle port (oracle/mie_oracle.c, OpenMP) on a bounded sample of the same cells.
With --gpus N (torchrun) every rank evaluates its own 2196-cell shard of an N-times larger grid (weak scaling) and the
finished rows are gathered to rank 0 over NVLink inside the timed region: by default every rank's copy engine writes them
into rank 0's CUDA-IPC-mapped buffer (geosmie_b200.dist.PeerGather); GEOSMIE_GATHER=store|nccl select the fused P2P-store
variant or the NCCL gather.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mie_particle_evals_per_sec"
UNIT = "particle-evals/s"
NANG = 371
FP64_PEAK_TFLOPS = 37.1   # measured on this pool's B200 with tools/fp64_peak.cu (DMMA m8n8k4), profiles/r01_fp64_peak.json
FP64_DFMA_PEAK_TFLOPS = 33.5   # same tool, plain DFMA (the pipe k_coeff runs on)
# dram__bytes_read.sum + dram__bytes_write.sum of k_gram per SU cell, from the `ncu --set full` capture
# profiles/r01k_gram_su_732cells.ncu-rep (968.82 MB + 196.93 MB over 732 dense cells)
GRAM_DRAM_BYTES_PER_CELL = (968.823552e6 + 196.926976e6) / 732


def build_su_plan():
    """Host inputs of the SU table exactly as dointegration.fun derives them (grid, m(lambda, RH), number weights)."""
    from geosmie_b200 import workloads
    return workloads.bin_plan("su", 0)


def flop_model(nmax, nmx_sum, ncell_factor=1):
    """Algorithmic FP64 flop of one pass.
    `contract`: the per-angle S+/S- formulation (k_contract): 4 FMA per (particle, n, angle) + 8 FMA per (particle, angle);
    `survey`:   the SURVEY 8d model F(p) = nmax (16 N_ang + 94) + 14 nmx + 22 N_ang (the reference's four complex dots);
    `gram`:     the Gram formulation (k_gram): four N x N blocks, K = 2 per particle -> 16 nmax^2 flop per particle, no
                tile padding counted;
    `gram_exec`: the DMMA flop k_gram actually issues (8-row tiles, full symmetric blocks, stacked tile for nmax <= 4);
    `coeff`:    14 nmx + 94 nmax per particle (recurrences, a_n, b_n, efficiencies)."""
    nm = np.asarray(nmax, dtype=np.float64)
    snm = float(nm.sum()) * ncell_factor
    npart = float(len(nm)) * ncell_factor
    ng = (len(nm) + 31) // 32
    pad = np.zeros(ng * 32)
    pad[:len(nm)] = nm
    gm = pad.reshape(ng, 32).max(axis=1)
    tg = np.ceil(gm / 8.0)
    in_gram = gm <= 64
    # DMMAs k_gram issues per group of 32 particles (16 k-steps): stacked c+/c- tile when nmax <= 4 (2 per k-step), 4 tg^2 up
    # to 4 tiles, 12 warps x tg x ceil((tg+2)/3) column tiles above (the ragged last third recomputes one tile)
    dm = np.where(gm <= 4, 32.0, np.where(tg <= 4, 64.0 * tg * tg, 16.0 * 12.0 * tg * np.floor((tg + 2) / 3)))
    return {"contract": snm * 8.0 * NANG + npart * 16.0 * NANG,
            "survey": snm * (16.0 * NANG + 94.0) + 14.0 * nmx_sum + npart * 22.0 * NANG,
            "gram": 16.0 * float((nm[np.repeat(in_gram, 32)[:len(nm)]] ** 2).sum()) * ncell_factor,
            "gram_exec": 512.0 * float(dm[in_gram].sum()) * ncell_factor,
            "coeff": 14.0 * nmx_sum + 94.0 * snm}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        # under-load samples: the upper half of the SM clock readings
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


class FastPowerSampler(threading.Thread):
    """Best-effort second sampler (50 ms period): lowest SM clock, highest power draw and the HW power-brake flag during the
    timed region.  Added to find out why k_coeff runs 1.6-1.9 ms instead of 1.03 ms on some ranks of a multi-GPU job although
    the 100 ms sampler sees 1965 MHz everywhere.  A driver that does not know one of the fields makes
    nvidia-smi print nothing; the result is then simply empty."""

    Q = "clocks.sm,power.draw,clocks_event_reasons.hw_power_brake_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, pw, brake, cap = [], [], False, False
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                pw.append(float(r[1]))
                brake |= r[2].lower().startswith("active")
                cap |= r[3].lower().startswith("active")
            except Exception:
                continue
        if not sm:
            return {}
        return {"sm_mhz_min": min(sm), "power_w_max": max(pw), "hw_power_brake": brake, "sw_power_cap": cap, "fast_samples": len(sm)}


def cpu_baseline(plan, seconds=12.0, threads=None):
    """Oracle port timed on the host cores on a bounded sample of SU cells (every cell costs the same: the grid is
    shared, only m changes).  Returns the cpu_baseline dict."""
    from oracle import mie_oracle as mo
    threads = threads or len(os.sched_getaffinity(0))
    cost = np.cos(np.radians(_angles()))
    t0 = time.time()
    sr = mo.SizeRange(plan.xx, cost)       # scipy Bessel pre-computation, once per bin like the reference
    t_pre = time.time() - t0
    ncell = len(plan.cells)
    order = np.random.default_rng(0).permutation(ncell)
    done, t_run = 0, 0.0
    for ci in order:
        m = plan.m[ci, 0]
        t1 = time.time()
        q, _, mu = sr.run(float(m.real), float(m.imag), want_mueller=True, nthreads=threads)
        mo.raw_sums(plan.xx, q, mu, plan.w[ci, 0])
        t_run += time.time() - t1
        done += 1
        if t_run > seconds:
            break
    evals = done * plan.xx.size
    return {"value": evals / t_run, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d of %d SU cells x %d particles x %d angles (oracle/mie_oracle.c + OpenMP, %.1f s; one-off scipy Bessel "
                      "pre-computation %.1f s not included)" % (done, ncell, plan.xx.size, NANG, t_run, t_pre)}, t_run


def _angles():
    return np.concatenate([np.linspace(0., 1., 100, endpoint=False), np.linspace(1., 10., 100, endpoint=False),
                           np.linspace(10., 180., 171, endpoint=True)])


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Python + numba and cannot
    travel to the GPU box; its algorithm is timed through the oracle port (plain C + OpenMP on all host threads --
    faster than the reference's own numba loops) on a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    plan = build_su_plan()
    vals = []
    for s in range(args.warmup + args.steps):
        cb, t = cpu_baseline(plan, seconds=float(os.environ.get("GEOSMIE_REF_SECONDS", "4.0")))
        if s >= args.warmup:
            vals.append((cb, t))
    v = float(np.mean([c["value"] for c, _ in vals]))
    cb = vals[-1][0]
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([t for _, t in vals])), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "su.json parameters, OPAC sulfate + HITRAN water indices",
            "config": {"workload": "optics_SU dense: 4459 x (61 lambda x 36 RH) cells x 371 angles (bounded sample per step)"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--elide", action="store_true", help="also report the zero-weight-elided rate (extra, not the headline)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from geosmie_b200 import _lib, workloads
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as td
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # control plane (barrier, timing exchange, IPC handle broadcast): NCCL; GEOSMIE_BENCH_CONTROL=gloo is an experiment that
        # keeps NCCL out of the process entirely (the rows then travel over the peer-memory gather only)
        control = os.environ.get("GEOSMIE_BENCH_CONTROL", "nccl")
        if control == "gloo":
            td.init_process_group("gloo", rank=rank, world_size=world)
        else:
            td.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ctrl_dev = torch.device("cpu") if (world > 1 and os.environ.get("GEOSMIE_BENCH_CONTROL", "nccl") == "gloo") else dev

    plan = build_su_plan()
    ang = _angles()
    cost = np.cos(np.radians(ang))
    ncell, nx = len(plan.cells), plan.xx.size
    h = _lib.Handle.get(local)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    table = _lib.Table(plan.xx, plan.nmax, cost, h)
    table.set_timing(True)

    mz_np, wp_np, ws_np, tpc = plan.tasks()
    assert tpc == 1
    # pinned host inputs (e2e) and device-resident copies (value)
    mz_h = torch.from_numpy(np.ascontiguousarray(mz_np).view(np.float64).reshape(ncell, 2).copy()).pin_memory()
    w_h = torch.from_numpy(np.ascontiguousarray(wp_np)).pin_memory()
    mz_d, w_d = mz_h.to(dev), w_h.to(dev)
    # two sets of device outputs: with more than one rank the copy engine may still be reading the rows of step k-1 while
    # step k computes (the fences gm_peer_wait / gm_peer_mark order the reuse two steps later)
    outs_d = [(torch.empty((ncell, 1, _lib.GM_NSCAL), dtype=torch.float64, device=dev),
               torch.empty((ncell, 4, NANG), dtype=torch.float64, device=dev),
               torch.empty((ncell, 6, 129), dtype=torch.float64, device=dev)) for _ in range(2 if world > 1 else 1)]
    scal_d, phase_d, coef_d = outs_d[0]
    cn_d = torch.empty((ncell,), dtype=torch.float64, device=dev)
    scal_h = torch.empty((ncell, 1, _lib.GM_NSCAL), dtype=torch.float64).pin_memory()
    phase_h = torch.empty((ncell, 4, NANG), dtype=torch.float64).pin_memory()
    coef_h = torch.empty((ncell, 6, 129), dtype=torch.float64).pin_memory()
    # multi-GPU: the finished rows of a step (sums, phase sums, GSF moments: 18 KB per cell) go to rank 0 over NVLink.
    # GEOSMIE_GATHER = peer (default): every rank's copy engine writes its rows into rank 0's buffer, mapped through CUDA
    #                  IPC (gm_peer_put: no SM time on either GPU); the transfer of step k overlaps the kernels of step k+1;
    #                = store: k_finalize and k_gsf store their results through the mapped pointers themselves (P2P stores);
    #                = nccl: NCCL gather on NCCL's stream (two packed buffers).
    # Every transfer is complete before the timed region ends (drain).
    pending = [None, None]
    gather_mode = "none"
    pg = None
    nscal, nph, nco = ncell * _lib.GM_NSCAL, ncell * 4 * NANG, ncell * 6 * 129
    if world > 1:
        from geosmie_b200 import dist
        comm = dist.Comm(rank, world, device=local, backend="gloo" if ctrl_dev.type == "cpu" else None)
        gather_mode = os.environ.get("GEOSMIE_GATHER", "peer")
        if ctrl_dev.type == "cpu" and gather_mode == "nccl":
            raise SystemExit("GEOSMIE_BENCH_CONTROL=gloo has no NCCL gather: use GEOSMIE_GATHER=peer|store|off")
        if gather_mode == "off":             # diagnostic: independent replicas, nothing leaves the GPU
            pass
        elif gather_mode in ("peer", "store"):
            pg = comm.peer_gather((nscal + nph + nco) * 8, nslot=2, handle=h)
            if pg is None:
                gather_mode = "nccl"
        if gather_mode == "nccl":
            width = _lib.GM_NSCAL + 4 * NANG + 6 * 129
            packed = [torch.empty((ncell, width), dtype=torch.float64, device=dev) for _ in range(2)]
            gather_buf = [[torch.empty_like(packed[0]) for _ in range(world)] if rank == 0 else None for _ in range(2)]
    step_no = [0]

    def gather_rows(src=None):
        """src = (scal, phase, coef) device pointers of this step's results (default: the bench's device tensors of this
        step's parity)."""
        i = step_no[0] & 1
        step_no[0] += 1
        sd, pd, cd = outs_d[i % len(outs_d)]
        if gather_mode == "off":
            return
        if gather_mode in ("peer", "store"):
            if gather_mode == "store" and src is None:
                return                          # the kernels of this step already stored into slot i (see step_device)
            ps, pp_, pc = src or (sd.data_ptr(), pd.data_ptr(), cd.data_ptr())
            pg.put(i, ps, nscal * 8, 0)
            pg.put(i, pp_, nph * 8, nscal * 8)
            pg.put(i, pc, nco * 8, (nscal + nph) * 8)
            h.peer_mark(i)
            return
        if pending[i] is not None:
            pending[i].wait()                   # the buffer pair of two steps ago is free again
        if src is None:
            parts = [sd.reshape(ncell, -1), pd.reshape(ncell, -1), cd.reshape(ncell, -1)]
        else:
            parts = [_dev_view(p, n).reshape(ncell, -1) for p, n in zip(src, (nscal, nph, nco))]
        torch.cat(parts, dim=1, out=packed[i])
        pending[i] = td.gather(packed[i], gather_buf[i], dst=0, async_op=True)
        if os.environ.get("GEOSMIE_BENCH_GATHER_SYNC"):      # diagnostic: no overlap with the next step
            pending[i].wait()
            pending[i] = None

    def _dev_view(p, n):
        class _A(object):
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(p), False), "version": 3}
        return torch.as_tensor(_A(), device=dev)

    def drain():
        if pg is not None:
            pg.join()                           # the compute stream (and the closing event) waits for the exchange stream
        for i in range(2):
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    def step_device():
        i = step_no[0] & 1
        sd, pd, cd = outs_d[i % len(outs_d)]
        coef_ptr = cd.data_ptr()
        if gather_mode == "peer":
            h.peer_wait(i)                                   # the put of two steps ago has finished reading this output set
        elif gather_mode == "store":
            table.set_mirror(pg.seg_ptr(i, 0), pg.seg_ptr(i, nscal * 8))
            coef_ptr = pg.seg_ptr(i, (nscal + nph) * 8)     # k_gsf writes the moments straight into rank 0's buffer
        table.run_dev(ncell, mz_d.data_ptr(), mz_d.data_ptr(), 1, w_d.data_ptr(), 0, sd.data_ptr(), pd.data_ptr(), elide=False)
        h.gsf_expand_phase4_dev(ang, ncell, pd.data_ptr(), coef_ptr, cn_d.data_ptr())
        if world > 1:
            gather_rows()

    # the public call of the table build (dointegration.fun -> BinPlan.evaluate -> gm_table_run_psd): per-cell refractive
    # index and PSD parameters in, reduced sums out; the number weights are generated on the device
    plan_psd = workloads.bin_plan("su", 0, device_psd=True)
    mz_psd, psd_par, psd_frac, _ = plan_psd.tasks_psd()
    psd_kind = plan_psd.psd_kind
    table.set_dr(plan_psd.dr)
    mz_hn, w_hn = mz_h.numpy(), w_h.numpy()          # views of the pinned host buffers
    scal_hn, phase_hn = scal_h.numpy(), phase_h.numpy()

    coef_hn = coef_h.numpy()

    def step_e2e():
        # the user-facing call with HOST buffers (gm_table_run_psd with the fused GSF stage): H2D of this step's inputs,
        # kernels, and D2H of the reduced sums and GSF moments pipelined batch by batch inside the library
        if pg is not None:
            h.peer_wait(2)                      # the library's device copies of the last step's rows have been sent
        table.run_psd(mz_psd, mz_psd, psd_kind, psd_par, psd_frac, elide=False, out=(scal_hn, phase_hn))
        if world > 1:
            gather_rows(e2e_src[0])
            if pg is not None:
                h.peer_mark(2)

    per_rank_ms = []            # [timed region][rank] ms per step (multi-GPU diagnostics)
    stagger_ms = float(os.environ.get("GEOSMIE_BENCH_STAGGER_MS", "0") or 0.0)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if stagger_ms > 0.0 and world > 1:
            # experiment (off by default, GEOSMIE_BENCH_STAGGER_MS=<step ms>): rank r starts r/world of a step late, INSIDE the
            # timed region, so that the ranks do not run the same kernel at the same instant
            time.sleep(1e-3 * stagger_ms * rank / world)
        for _ in range(steps):
            fn()
        if world > 1:
            drain()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=ctrl_dev)
            every = [torch.zeros_like(t) for _ in range(world)]
            td.all_gather(every, t)
            per_rank_ms.append([float(x.item()) / steps for x in every])
            ms = max(float(x.item()) for x in every)              # the slowest rank defines the step
        return ms / steps

    for _ in range(max(args.warmup, 3)):
        step_device()
    launches0 = h.launch_count()
    sampler = ClockSampler(local)          # every rank samples its own GPU; rank 0 reports its own and the slowest
    sampler.start()
    fast = FastPowerSampler(local)
    fast.start()
    time.sleep(0.3)
    ms_dev = timed(step_device, args.steps)
    launches = (h.launch_count() - launches0) // args.steps
    kms = table.last_kernel_ms()            # CUDA events around the launches of the LAST timed step
    per_rank_kernel_ms = None
    if world > 1:
        per_rank_kernel_ms = [None] * world
        td.all_gather_object(per_rank_kernel_ms, {k: round(kms[k], 4) for k in ("k_coeff", "k_gram", "k_gram_sum_eval", "k_finalize")})
    stats = table.last_stats()
    table.set_mirror(None, None)
    table.set_gsf(ang, 129, False, coef_hn, None)
    e2e_src = [None]
    if world > 1:
        table.run_psd(mz_psd, mz_psd, psd_kind, psd_par, psd_frac, elide=False, out=(scal_hn, phase_hn))
        e2e_src[0] = table.device_outputs() + (table.gsf_device()[0],)   # the library's device copies of the results
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    table.set_gsf(None)
    clocks = sampler.stop()
    clocks.update(fast.stop())
    if world > 1:
        allc = [None] * world
        td.all_gather_object(allc, clocks)
        clocks = dict(allc[0])
        clocks["per_rank_sm_mhz"] = [c.get("sm_mhz") for c in allc]
        clocks["per_rank_sm_mhz_min"] = [c.get("sm_mhz_min") for c in allc]
        clocks["per_rank_power_w_max"] = [c.get("power_w_max") for c in allc]
        clocks["hw_power_brake"] = any(c.get("hw_power_brake") for c in allc)
        clocks["sw_power_cap"] = any(c.get("sw_power_cap") for c in allc)
        clocks["reasons"] = sorted(set(r for c in allc for r in c.get("reasons", [])))
        known = [c["sm_mhz"] for c in allc if c.get("sm_mhz")]
        if known:
            clocks["sm_mhz"] = min(known)              # the slowest GPU of the job

    total_evals = float(ncell) * nx * world
    value = total_evals / (ms_dev * 1e-3)
    e2e = total_evals / (ms_e2e * 1e-3)
    fl = flop_model(plan.nmax, stats["sum_nmx"], ncell)

    def tf(flop, ms):
        return flop / (ms * 1e-3) / 1e12 if ms and ms > 0 else None

    kernels = {
        "k_gram": {"bound": "tensor", "ms": kms["k_gram"], "achieved": tf(fl["gram"], kms["k_gram"]), "peak": FP64_PEAK_TFLOPS,
                   "unit": "TFLOP/s", "executed": tf(fl["gram_exec"], kms["k_gram"]),
                   "note": "achieved = 16 nmax^2 flop per particle (no tile padding); executed = DMMA flop actually issued"},
        "k_coeff": {"bound": "fp64 vector pipe (latency-bound recurrences)", "ms": kms["k_coeff"],
                    "achieved": tf(fl["coeff"], kms["k_coeff"]), "peak": FP64_DFMA_PEAK_TFLOPS, "unit": "TFLOP/s"},
        "k_gram_sum+k_gram_eval": {"bound": "tensor", "ms": kms["k_gram_sum_eval"],
                                   "achieved": tf(8.0 * NANG * float(plan.nmax.max()) ** 2 * ncell, kms["k_gram_sum_eval"]),
                                   "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                                   "note": "4 quadratic forms of size max(nmax) at 371 angles per cell"},
        "k_contract": {"bound": "tensor", "ms": kms["k_contract"], "achieved": None, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                       "note": "per-angle contraction of groups with nmax > 64: none in optics_SU"},
        "k_finalize": {"bound": "hbm", "ms": kms["k_finalize"]},
    }
    for k in kernels.values():
        if k.get("achieved") and k.get("peak"):
            k["frac"] = k["achieved"] / k["peak"]
    dom = max(("k_gram", "k_coeff", "k_gram_sum+k_gram_eval", "k_contract"), key=lambda k: kernels[k]["ms"] or 0.0)
    D = kernels[dom]
    ach = D.get("achieved")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "su.json parameters; OPAC sulfate + HITRAN water refractive indices (tests/golden/hostlogic.npz)",
        "config": {"workload": "optics_SU dense table build: 1 bin x 4459 sizes x 61 lambda x 36 RH = 2196 cells, 371 angles, "
                               "then 129 GSF moments x 6 per cell", "cells_per_gpu": ncell, "nx": nx, "nang": NANG,
                   "l2_policy": "inputs larger than L2 per step: 78 MB weights + 2.9 GB coefficient stream + 0.62 GB partial Gram blocks re-written every step",
                   "parallelism": "cells sharded, %d rank(s), gather to rank 0: %s (overlapped with the next step, complete inside the timed region)"
                                  % (world, {"peer": "copy-engine puts into rank 0's IPC-mapped buffer over NVLink (gm_peer_put)",
                                             "store": "P2P stores of k_finalize / k_gsf into rank 0's IPC-mapped buffer",
                                             "nccl": "NCCL gather", "none": "none (1 rank)", "off": "SWITCHED OFF (diagnostic run)"}[gather_mode])},
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(mz_psd.nbytes * 2 + psd_par.nbytes + psd_frac.nbytes),
                "api": "gm_table_run_psd with the fused GSF stage (host buffers: per-cell m and PSD parameters in, reduced sums and GSF moments out)",
                "d2h_bytes_per_step": int(scal_h.numel() * 8 + phase_h.numel() * 8 + coef_h.numel() * 8)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor" if D["bound"] == "tensor" else "fp64",
                     "kernel": (dom + " (FP64 DMMA m8n8k4)") if D["bound"] == "tensor" else dom,
                     "achieved": ach, "peak": D.get("peak"), "unit": "TFLOP/s", "frac": D.get("frac"),
                     "executed_frac": (D["executed"] / D["peak"]) if D.get("executed") else None,
                     "traffic": GRAM_DRAM_BYTES_PER_CELL * ncell if dom == "k_gram" else None,
                     "traffic_note": "DRAM bytes of the k_gram launch of one step, scaled per cell from the ncu capture in "
                                     "profiles/ (algorithmic: 32 B x sum(nmax) coefficient stream = %.2e B read + partial Gram "
                                     "blocks written)" % (32.0 * float(np.sum(plan.nmax)) * ncell),
                     "peak_source": "measured on this pool: tools/fp64_peak.cu DMMA burst 37.1 TFLOP/s, DFMA 33.5 TFLOP/s "
                                    "(MEASURED_PEAKS.json has no FP64 entry; nominal 37 TFLOP/s)",
                     "flop_per_launch_set": fl["gram"] if dom == "k_gram" else None,
                     "kernel_ms_per_step": kms, "kernels": kernels,
                     "step_tflops_survey_flop_model_per_gpu": fl["survey"] / (ms_dev * 1e-3) / 1e12,
                     "step_tflops_per_angle_model_per_gpu": fl["contract"] / (ms_dev * 1e-3) / 1e12,
                     "model_note": "TFLOP/s the reference formulation (SURVEY 8d) / the per-angle S+/S- formulation would need to "
                                   "do this step in the same time; both exceed the 37.1 TFLOP/s peak because the Gram form needs "
                                   "~9x fewer flop when nmax << N_ang",
                     "kernel_share_of_step": (D["ms"] / ms_dev) if ms_dev else None},
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_baseline(plan)
    if args.elide and rank == 0:
        def step_elide():
            table.run_dev(ncell, mz_d.data_ptr(), mz_d.data_ptr(), 1, w_d.data_ptr(), 0, scal_d.data_ptr(), phase_d.data_ptr(), elide=True)
        for _ in range(2):
            step_elide()
        ms_el = timed(step_elide, args.steps) if world == 1 else None
        line["elided"] = {"ms_per_step": ms_el, "grid_evals_per_sec": total_evals / world / (ms_el * 1e-3) if ms_el else None,
                          "evaluated": table.last_stats()["evals"]}
    if world > 1:
        line["per_rank_ms_per_step"] = {"device": per_rank_ms[0], "e2e": per_rank_ms[1] if len(per_rank_ms) > 1 else None}
        line["per_rank_kernel_ms"] = per_rank_kernel_ms
    if rank == 0:
        print(json.dumps(line))
    table.close()
    if world > 1:
        comm.close()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
