#!/usr/bin/env python3
"""bench.py -- headline benchmark of the GEOSmie Mie lookup-table hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workloads ss,su] [--no-lut] [--no-cpu-baseline]

Workloads (BASELINE.json configs):
  optics_SS (configs[2], the HEADLINE: the largest single-GPU configuration): the sea-salt table -- ss.json's five Gong bins on
            their 5963 / 6244 / 5889 / 5963 / 5608-point size grids (x up to 2513, nmax up to 2567) x 61 wavelengths x 36 RH =
            10,980 cells = 65,148,732 Mie particle-evaluations per step, 371 angles each;
  optics_SU (configs[1], reported under "workloads"): the sulfate table -- 4459 sizes x 2196 cells = 9,791,964 evaluations.
Both are evaluated DENSE (zero-weight particles included = the work the reference does) and followed by the GSF moment expansion
of every cell.  Refractive indices are the OPAC / HITRAN values recorded in tests/golden/hostlogic.npz.

One "step" = one pass of the hot path over the whole table:
  value  particle-evals/s with the inputs (m, number weights) resident in HBM (device-pointer C ABI, CUDA-event timed);
  e2e    the same through the host-buffer C ABI the table driver uses (gm_table_run_psd with the fused GSF stage: per-cell
         refractive indices and size-distribution parameters in host memory -> H2D -> kernels -> D2H of the reduced sums and GSF
         moments into pinned host memory, all inside the timed region; the per-bin tables are created once, like the reference's
         MultipleMie.preCalculate outside its cell loops);
  roofline   the kernel with the largest CUDA-event time inside the timed steps (k_contract, FP64 DMMA, on optics_SS); duration
             from CUDA events recorded around each of its launches on the launching stream; `traffic` = DRAM bytes of those
             launches from the ncu pass recorded in profiles/r02_dram_bytes.json; every kernel of the step is listed under
             roofline.kernels;
  lut_build_s   wall time of the real table build: runoptics.main (inputs, kernels, post-processing, file) followed by rungsf.main
                (moments into the file) for optics_SU / optics_BC / optics_SS, at N GPUs under torchrun (cells sharded);
  cpu_baseline  the UNMODIFIED reference (baseline/_ref, Python + numba) on all host cores on a bounded proportional sample of the
                same table (baseline/ref_runner.py), plus the C + OpenMP oracle port as a second figure.
With --gpus N (torchrun) every rank evaluates its own copy of the table (weak scaling) and the finished rows (sums, phase sums,
GSF moments: 18 KB per cell) are gathered to rank 0 over NVLink inside the timed region: every rank's copy engine writes them into
rank 0's CUDA-IPC-mapped buffer (geosmie_b200.dist.PeerGather); GEOSMIE_GATHER=nccl selects the NCCL gather, =off none.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mie_particle_evals_per_sec"
UNIT = "particle-evals/s"
NANG = 371
NG = 129
FP64_PEAK_TFLOPS = 37.1        # measured on this pool's B200 with tools/fp64_peak.cu (DMMA m8n8k4), profiles/r01_fp64_peak.json
FP64_DFMA_PEAK_TFLOPS = 33.5   # same tool, plain DFMA
PEAK_SOURCE = ("measured on this pool: tools/fp64_peak.cu DMMA burst 37.1 TFLOP/s, DFMA 33.5 TFLOP/s (MEASURED_PEAKS.json has no "
               "FP64 entry; nominal 37 TFLOP/s)")
WORKLOADS = {
    "ss": {"name": "optics_SS", "bins": 5,
           "text": "optics_SS dense table build: 5 bins x (5963, 6244, 5889, 5963, 5608) sizes x 61 lambda x 36 RH = 10,980 cells, "
                   "371 angles, then 129 GSF moments x 6 per cell"},
    "su": {"name": "optics_SU", "bins": 1,
           "text": "optics_SU dense table build: 1 bin x 4459 sizes x 61 lambda x 36 RH = 2196 cells, 371 angles, then 129 GSF "
                   "moments x 6 per cell"},
}
DRAM_BYTES_FILE = os.path.join(ROOT, "profiles", "r02_dram_bytes.json")


def _angles():
    return np.concatenate([np.linspace(0., 1., 100, endpoint=False), np.linspace(1., 10., 100, endpoint=False),
                           np.linspace(10., 180., 171, endpoint=True)])


# ----------------------------------------------------------------------------------------------- flop models
def flop_model(nmax, nmx_sum, ncell_factor=1):
    """Algorithmic FP64 flop of one pass over one bin.
    `contract`: the per-angle S+/S- formulation for EVERY particle: 4 FMA per (particle, n, angle) + 8 FMA per (particle, angle);
    `contract_direct`: the same for the particles k_contract actually handles (groups with max nmax > 64);
    `survey`:   the SURVEY 8d model F(p) = nmax (16 N_ang + 94) + 14 nmx + 22 N_ang (the reference's four complex dots);
    `gram`:     the Gram formulation: four N x N blocks, K = 2 per particle -> 16 nmax^2 flop per particle of the Gram groups, no
                tile padding counted (`gram_big`: the classes k_gram handles, max nmax 9..64; `gram_small`: k_small's, <= 8);
    `gram_exec`: the DMMA flop k_gram + k_small actually issue (8-row tiles, full symmetric blocks, stacked tile for nmax <= 4);
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
    small = gm <= 8
    pg = np.repeat(in_gram, 32)[:len(nm)]
    ps = np.repeat(small, 32)[:len(nm)]
    ngram = 8.0 * float(tg[in_gram].max()) if in_gram.any() else 0.0
    dm = np.where(gm <= 4, 32.0, np.where(tg <= 4, 64.0 * tg * tg, 16.0 * 12.0 * tg * np.floor((tg + 2) / 3)))
    return {"contract": snm * 8.0 * NANG + npart * 16.0 * NANG,
            "contract_direct": (float(nm[~pg].sum()) * 8.0 * NANG + float((~pg).sum()) * 16.0 * NANG) * ncell_factor,
            "survey": snm * (16.0 * NANG + 94.0) + 14.0 * nmx_sum + npart * 22.0 * NANG,
            "gram": 16.0 * float((nm[pg] ** 2).sum()) * ncell_factor,
            "gram_big": 16.0 * float((nm[pg & ~ps] ** 2).sum()) * ncell_factor,
            "gram_small": 16.0 * float((nm[ps] ** 2).sum()) * ncell_factor,
            "gram_exec": 512.0 * float(dm[in_gram].sum()) * ncell_factor,
            "gram_exec_big": 512.0 * float(dm[in_gram & ~small].sum()) * ncell_factor,
            "coeff": 14.0 * nmx_sum + 94.0 * snm,
            "coeff_small": 94.0 * float(nm[ps].sum()) * ncell_factor,      # + 14 nmx of those particles (not split out by the kernel stats)
            # k_gram_eval + k_gram_interp: the four quadratic forms at the 2 N + 1 Chebyshev nodes (N = 8 x largest Gram class), then the
            # barycentric interpolation of the four degree-2N polynomials to the table's angles
            "eval": ((8.0 * (2 * ngram + 1) * ngram ** 2 + 8.0 * NANG * (2 * ngram + 1)) if pg.any() else 0.0) * ncell_factor}


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons / power during the timed regions (B200_PROFILING.md recipe), 100 ms period."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

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
        sm, pw, mx, reasons = [], [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        # under-load samples: the readings taken while the GPU drew more than half of the highest power seen
        load = [s for s, p in zip(sm, pw) if pw and p >= 0.5 * max(pw)]
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": len(load), "sm_mhz_min_under_load": min(load) if load else None,
                "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------------------------- CPU arms
def _limit_blas():
    """The oracle port mixes OpenMP regions with numpy; idle OpenBLAS workers spinning between the regions cost it 3.5x in round 1
    (0.185 against 0.65 M evals/s on one box, depending on whether torchrun had set OMP_NUM_THREADS=1)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1, user_api="blas")
    except Exception:
        return None


def port_baseline(sp, seconds=6.0, threads=None):
    """The C + OpenMP oracle port (oracle/mie_oracle.c) on a bounded sample of cells of every bin of the table."""
    from geosmie_b200 import workloads
    from oracle import mie_oracle as mo
    threads = threads or len(os.sched_getaffinity(0))
    cost = np.cos(np.radians(_angles()))
    lim = _limit_blas()
    nb = WORKLOADS[sp]["bins"]
    stride = 1 if sp == "su" else 8
    plans, srs = [], []
    for b in range(nb):
        plan = workloads.bin_plan(sp, b, cells=[(7 * k % 61, 5 * k % 36) for k in range(64)])
        sel = np.arange(0, plan.xx.size, stride)
        plans.append((plan, sel))
        srs.append(mo.SizeRange(plan.xx[sel], cost))     # scipy Bessel pre-computation, once per bin like the reference
    evals, t_run, k = 0, 0.0, 0
    while t_run < seconds:
        for b in range(nb):
            plan, sel = plans[b]
            ci = k % len(plan.cells)
            m = plan.m[ci, 0]
            t1 = time.time()
            q, _, mu = srs[b].run(float(m.real), float(m.imag), want_mueller=True, nthreads=threads)
            mo.raw_sums(plan.xx[sel], q, mu, plan.w[ci, 0][sel])
            t_run += time.time() - t1
            evals += sel.size
        k += 1
    del lim
    return {"value": evals / t_run, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d cell sets (one cell of each of the %d bins of %s, every %d-th grid point, 371 angles) with oracle/mie_oracle.c + "
                      "OpenMP, %.1f s; scipy Bessel pre-computation per bin untimed" % (k, nb, WORKLOADS[sp]["name"], stride, t_run)}


def reference_baseline(sp, seconds=12.0, stride=None, cores=None, farm=None):
    """The unmodified reference on all host cores (baseline/ref_runner.py).  Returns (cpu_baseline dict, farm)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_runner
    stride = stride or (20 if sp == "ss" else 8)
    own = farm is None
    if own:
        farm = ref_runner.Farm(sp, stride, cores=cores)
    evals, wall, core, nsets = 0, 0.0, 0.0, 0
    while wall < seconds:
        n, w, c = farm.step()
        evals, wall, core, nsets = evals + n, wall + w, core + c, nsets + farm.cores
    out = {"value": evals / wall, "unit": UNIT, "cores": farm.cores, "kind": "reference",
           "single_core_value": evals / core,
           "sample": ref_runner.describe(sp, stride, farm.cores, nsets, farm.nx) + "; %.1f s wall, %.1f s untimed preparation" % (wall, farm.prep_s)}
    if own:
        farm.close()
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (baseline/_ref, unmodified Python + numba) on all host
    cores, on the headline workload's config; each step is a bounded proportional sample of the table (one cell set per core)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sp = args.workloads.split(",")[0]
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_runner
    if ref_runner.reference_root() is None:
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is missing (baseline/install_reference.sh needs /root/reference)"}))
        return
    stride = int(os.environ.get("GEOSMIE_REF_STRIDE", "20" if sp == "ss" else "8"))
    farm = ref_runner.Farm(sp, stride)
    vals, walls, evs, nsets = [], [], 0, 0
    for s in range(args.warmup + args.steps):
        n, wall, core = farm.step()
        if s >= args.warmup:
            vals.append(n / wall)
            walls.append(wall)
            evs += n
            nsets += farm.cores
    farm.close()
    v = evs / sum(walls)
    cb = {"value": v, "unit": UNIT, "cores": farm.cores, "kind": "reference",
          "sample": ref_runner.describe(sp, stride, farm.cores, nsets, farm.nx) + "; one cell set per core and step"}
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "%s.json parameters, OPAC + HITRAN water refractive indices" % sp,
            "config": {"workload": WORKLOADS[sp]["text"] + " (bounded proportional sample per step)"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
class TableBench(object):
    """One species table on this rank's GPU: per-bin tables, device-resident inputs and outputs, the two kinds of step."""

    def __init__(self, sp, torch, dev, h, stream, world, rank):
        from geosmie_b200 import _lib, workloads
        self.sp, self.torch, self.dev, self.h, self.world, self.rank = sp, torch, dev, h, world, rank
        self.ang = _angles()
        cost = np.cos(np.radians(self.ang))
        self.bins = []
        nout = 2 if world > 1 else 1
        self.cells, self.evals = 0, 0
        for b in range(WORKLOADS[sp]["bins"]):
            plan = workloads.bin_plan(sp, b, device_psd=True)
            mz, par, frac, tpc = plan.tasks_psd()
            assert tpc == 1
            ncell, nx = len(plan.cells), plan.xx.size
            table = _lib.Table(plan.xx, plan.nmax, cost, h)
            table.set_timing(True)
            if plan.psd_kind == _lib.PSD_LOGNORM:
                table.set_dr(plan.dr)
            # the number weights of every cell, generated once by the library's own PSD kernel and kept resident for the `value` steps
            table.run_psd(mz, mz, plan.psd_kind, par, frac, elide=False)
            w_d = torch.from_numpy(np.ascontiguousarray(table.get_weights()[:, 0, :])).to(dev)
            mz_h = torch.from_numpy(np.ascontiguousarray(mz).view(np.float64).reshape(ncell, 2).copy()).pin_memory()
            B = {"plan": plan, "table": table, "ncell": ncell, "nx": nx, "mz": mz, "par": par, "frac": frac, "w_d": w_d,
                 "mz_d": mz_h.to(dev), "cn_d": torch.empty((ncell,), dtype=torch.float64, device=dev),
                 "outs": [(torch.empty((ncell, 1, _lib.GM_NSCAL), dtype=torch.float64, device=dev),
                           torch.empty((ncell, 4, NANG), dtype=torch.float64, device=dev),
                           torch.empty((ncell, 6, NG), dtype=torch.float64, device=dev)) for _ in range(nout)],
                 "scal_h": torch.empty((ncell, 1, _lib.GM_NSCAL), dtype=torch.float64).pin_memory().numpy(),
                 "phase_h": torch.empty((ncell, 4, NANG), dtype=torch.float64).pin_memory().numpy(),
                 "coef_h": torch.empty((ncell, 6, NG), dtype=torch.float64).pin_memory().numpy(),
                 "row_bytes": ncell * (_lib.GM_NSCAL + 4 * NANG + 6 * NG) * 8, "e2e_src": None}
            self.bins.append(B)
            self.cells += ncell
            self.evals += ncell * nx
        self.nscal = _lib.GM_NSCAL
        self.step_no = 0
        self.gather_mode, self.pg, self.comm = "none", None, None
        self.pending = [None, None]

    # ---- multi-GPU gather of the finished rows to rank 0
    def setup_gather(self, comm, td):
        self.comm, self.td = comm, td
        self.gather_mode = os.environ.get("GEOSMIE_GATHER", "peer")
        if self.gather_mode == "peer":
            self.pg = comm.peer_gather(sum(B["row_bytes"] for B in self.bins), nslot=2, handle=self.h)
            if self.pg is None:
                self.gather_mode = "nccl"
        if self.gather_mode == "nccl":
            width = self.nscal + 4 * NANG + 6 * NG
            for B in self.bins:
                B["packed"] = [self.torch.empty((B["ncell"], width), dtype=self.torch.float64, device=self.dev) for _ in range(2)]
                B["gbuf"] = [[self.torch.empty_like(B["packed"][0]) for _ in range(self.world)] if self.rank == 0 else None for _ in range(2)]

    def _view(self, p, n):
        class _A(object):
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(p), False), "version": 3}
        return self.torch.as_tensor(_A(), device=self.dev)

    def _gather_bin(self, B, slot, off, src):
        n = B["ncell"]
        ns, nph, nco = n * self.nscal, n * 4 * NANG, n * 6 * NG
        if self.gather_mode == "peer":
            for p, nb in zip(src, (ns, nph, nco)):
                self.pg.put(slot, p, nb * 8, off)
                off += nb * 8
        elif self.gather_mode == "nccl":
            key = id(B)
            if self.pending[slot] is None:
                self.pending[slot] = {}
            if self.pending[slot].get(key) is not None:
                self.pending[slot][key].wait()
            parts = [self._view(p, nb).reshape(n, -1) for p, nb in zip(src, (ns, nph, nco))]
            self.torch.cat(parts, dim=1, out=B["packed"][slot])
            self.pending[slot][key] = self.td.gather(B["packed"][slot], B["gbuf"][slot], dst=0, async_op=True)
        return off

    def drain(self):
        if self.pg is not None:
            self.pg.join()
        for slot in range(2):
            if self.pending[slot]:
                for w in self.pending[slot].values():
                    if w is not None:
                        w.wait()
                self.pending[slot] = None

    # ---- the two kinds of step
    def step_device(self):
        slot = self.step_no & 1
        self.step_no += 1
        if self.gather_mode == "peer":
            self.h.peer_wait(slot)               # the puts of two steps ago have finished reading this output set
        off = 0
        for B in self.bins:
            sd, pd, cd = B["outs"][slot % len(B["outs"])]
            B["table"].run_dev(B["ncell"], B["mz_d"].data_ptr(), B["mz_d"].data_ptr(), 1, B["w_d"].data_ptr(), 0, sd.data_ptr(),
                               pd.data_ptr(), elide=False)
            self.h.gsf_expand_phase4_dev(self.ang, B["ncell"], pd.data_ptr(), cd.data_ptr(), B["cn_d"].data_ptr())
            if self.world > 1 and self.gather_mode != "off":
                off = self._gather_bin(B, slot, off, (sd.data_ptr(), pd.data_ptr(), cd.data_ptr()))
        if self.gather_mode == "peer":
            self.h.peer_mark(slot)

    def prepare_e2e(self):
        for B in self.bins:
            B["table"].set_mirror(None, None)
            B["table"].set_gsf(self.ang, NG, False, B["coef_h"], None)

    def step_e2e(self):
        """The user-facing call with HOST buffers (gm_table_run_psd with the fused GSF stage), bin after bin."""
        slot = self.step_no & 1
        self.step_no += 1
        if self.pg is not None:
            self.h.peer_wait(2)                  # the library's device copies of the last step's rows have been sent
        off = 0
        for B in self.bins:
            t = B["table"]
            t.run_psd(B["mz"], B["mz"], B["plan"].psd_kind, B["par"], B["frac"], elide=False, out=(B["scal_h"], B["phase_h"]))
            if self.world > 1 and self.gather_mode != "off":
                # the library's device copies of this bin's rows (owned by the bin's table: they stay valid until its next call)
                src = t.device_outputs() + (t.gsf_device()[0],)
                off = self._gather_bin(B, slot, off, src)
        if self.pg is not None:
            self.h.peer_mark(2)

    def finish_e2e(self):
        for B in self.bins:
            B["table"].set_gsf(None)

    def kernel_ms(self):
        tot = {}
        for B in self.bins:
            k = B["table"].last_kernel_ms()
            for name, v in k.items():
                if name == "launches":
                    for kk, n in v.items():
                        tot.setdefault("launches", {}).setdefault(kk, 0)
                        tot["launches"][kk] += n
                else:
                    tot[name] = tot.get(name, 0.0) + v
        return tot

    def stats(self):
        tot = {}
        for B in self.bins:
            for k, v in B["table"].last_stats().items():
                tot[k] = tot.get(k, 0.0) + v
        return tot

    def flops(self, nmx_sum):
        tot = {}
        for B in self.bins:
            f = flop_model(B["plan"].nmax, 0.0, B["ncell"])
            for k, v in f.items():
                tot[k] = tot.get(k, 0.0) + v
        tot["coeff"] += 14.0 * nmx_sum
        tot["survey"] += 14.0 * nmx_sum
        return tot

    def close(self):
        for B in self.bins:
            B["table"].close()


def _dram_bytes():
    """DRAM bytes per kernel and per cell of each workload from the ncu pass of the final kernels (tools/dram_bytes.py ->
    profiles/r02_dram_bytes.json: dram__bytes_read.sum + dram__bytes_write.sum of every launch of one dense step)."""
    try:
        with open(DRAM_BYTES_FILE) as fp:
            return json.load(fp)
    except Exception:
        return {}


def lut_build(sp, world, rank, td, torch, dense=False):
    """Wall seconds of `runoptics --name <sp>.json` followed by `rungsf --filename optics_<sp>.nomom.nc4` in a scratch directory
    (median of three runs with the CUDA context, the library and the allocations warm; the first run's time is reported as `cold_s`)."""
    from geosmie_b200 import runoptics, workloads
    from geosmie_b200.gsf import rungsf
    import contextlib
    res = {}
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        with contextlib.redirect_stdout(sys.stderr):      # the drivers print progress like the reference's; stdout carries the JSON line only
            cfg = workloads.write_run_dir(d, sp)
            os.chdir(d)
            try:
                runs = []
                for attempt in ("cold_s", "s1", "s2", "s3"):
                    out = os.path.join(d, "out_" + attempt)
                    os.makedirs(out)
                    if world > 1:
                        td.barrier()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    runoptics.main(["--name", cfg, "--dest", out] + (["--dense"] if dense else []))
                    t1 = time.perf_counter()
                    if rank == 0:
                        rungsf.main(["--filename", os.path.join(out, "optics_%s.nomom.nc4" % sp), "--dest", out])
                    torch.cuda.synchronize()
                    if world > 1:
                        td.barrier()
                    t2 = time.perf_counter()
                    if attempt == "cold_s":
                        res["cold_s"], res["runoptics_cold_s"], res["rungsf_cold_s"] = t2 - t0, t1 - t0, t2 - t1
                    else:
                        runs.append((t2 - t0, t1 - t0, t2 - t1))
                    shutil.rmtree(out, ignore_errors=True)
                runs.sort()
                res["s"], res["runoptics_s"], res["rungsf_s"] = runs[len(runs) // 2]      # the median of the three warm runs
                res["warm_runs_s"] = [r[0] for r in runs]
            finally:
                os.chdir(old)
    return res


def fine_grid_build(sp, nlam, world, rank, td, torch, comm):
    """BASELINE config 5: the species table on `nlam` wavelengths (cells sharded over the ranks, rows gathered to rank 0; the phase
    matrices never leave the GPUs) followed by the RRTMG band averages of all columns.  Wall seconds of the second of two builds."""
    import contextlib
    from geosmie_b200 import bandaverage, dointegration as DI, workloads
    files, cfg = workloads.fine_grid_files(sp, nlam)
    res = {"n_lambda": nlam}
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        for name, text in files.items():
            os.makedirs(os.path.dirname(os.path.join(d, name)), exist_ok=True)
            with open(os.path.join(d, name), "w") as fp:
                fp.write(text)
        os.chdir(d)
        try:
            with contextlib.redirect_stdout(sys.stderr):
                warm = []
                for attempt in ("cold_s", "s1", "s2"):
                    if world > 1:
                        td.barrier()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    out = DI.fun(cfg, "json", d, False, write=False, comm=comm, keep_phase=False)
                    t1 = time.perf_counter()
                    if rank == 0:
                        vals = out["vals"]
                        for var in bandaverage.varsToAverage:
                            a = vals[var].transpose(0, 2, 1)                      # (bin, rh, lambda)
                            bandaverage.average_columns(out["wavelength"], a.reshape(a.shape[0] * a.shape[1], -1), "RRTMG")
                        res["cells"] = int(vals["qext"].size)
                    torch.cuda.synchronize()
                    if world > 1:
                        td.barrier()
                    t2 = time.perf_counter()
                    if attempt == "cold_s":
                        res["cold_s"], res["table_cold_s"], res["bands_cold_s"] = t2 - t0, t1 - t0, t2 - t1
                    else:
                        warm.append((t2 - t0, t1 - t0, t2 - t1))
                warm.sort()
                res["s"], res["table_s"], res["bands_s"] = warm[0]           # the faster of the two warm builds
                res["warm_runs_s"] = [w[0] for w in warm]
        finally:
            os.chdir(old)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workloads", default="ss,su", help="comma list; the first one is the headline (default ss,su)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lut", action="store_true", help="skip the runoptics / rungsf table-build timing")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        print("bench.py: --warmup %d raised to 3 (timing rule: at least 3 warm-up steps)" % args.warmup, file=sys.stderr)
        args.warmup = 3

    import torch
    from geosmie_b200 import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td = None
    comm = None
    if world > 1:
        import torch.distributed as td
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from geosmie_b200 import dist
        comm = dist.Comm(rank, world, device=local)
    h = _lib.Handle.get(local)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    dram = _dram_bytes()

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, drain):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        if world > 1:
            drain()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        per_rank = None
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            every = [torch.zeros_like(t) for _ in range(world)]
            td.all_gather(every, t)
            per_rank = [float(x.item()) / steps for x in every]
            ms = max(float(x.item()) for x in every)              # the slowest rank defines the step
        return ms / steps, per_rank

    sampler = ClockSampler(local)          # every rank samples its own GPU
    sampler.start()
    results = {}
    order = [w for w in args.workloads.split(",") if w in WORKLOADS]
    for sp in order:
        tb = TableBench(sp, torch, dev, h, stream, world, rank)
        if world > 1:
            tb.setup_gather(comm, td)
        for _ in range(args.warmup):
            tb.step_device()
        launches0 = h.launch_count()
        ms_dev, pr_dev = timed(tb.step_device, args.steps, tb.drain)
        launches = (h.launch_count() - launches0) // args.steps
        kms = tb.kernel_ms()            # CUDA events around every launch of the LAST timed step
        stats = tb.stats()
        tb.prepare_e2e()
        if world > 1:
            for B in tb.bins:           # the library's own device copies of the results exist after one call
                B["table"].run_psd(B["mz"], B["mz"], B["plan"].psd_kind, B["par"], B["frac"], elide=False, out=(B["scal_h"], B["phase_h"]))
        for _ in range(2):
            tb.step_e2e()
        ms_e2e, pr_e2e = timed(tb.step_e2e, args.steps, tb.drain)
        tb.finish_e2e()
        fl = tb.flops(stats["sum_nmx"])
        total_evals = float(tb.evals) * world

        def tf(flop, ms):
            return flop / (ms * 1e-3) / 1e12 if ms and ms > 0 else None

        db = dram.get(WORKLOADS[sp]["name"], {})
        kernels = {
            "k_contract": {"bound": "tensor", "ms": kms.get("k_contract", 0.0), "achieved": tf(fl["contract_direct"], kms.get("k_contract", 0.0)),
                           "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "flop": fl["contract_direct"],
                           "note": "per-angle S+/S- contraction (FP64 DMMA) of the groups with max nmax > 64: (8 nmax + 16) N_ang flop per particle"},
            "k_gram": {"bound": "tensor", "ms": kms.get("k_gram", 0.0), "achieved": tf(fl["gram_big"], kms.get("k_gram", 0.0)), "peak": FP64_PEAK_TFLOPS,
                       "unit": "TFLOP/s", "flop": fl["gram_big"], "executed": tf(fl["gram_exec_big"], kms.get("k_gram", 0.0)),
                       "note": "Gram blocks of the groups with max nmax 9..64: achieved = 16 nmax^2 flop per particle (no tile padding); executed = DMMA flop issued"},
            "k_small": {"bound": "fp64 pipe (DFMA recurrences + DMMA Gram blocks in one kernel)", "ms": kms.get("k_small", 0.0),
                        "achieved": tf(fl["coeff_small"] + fl["gram_small"], kms.get("k_small", 0.0)), "peak": FP64_DFMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                        "flop": fl["coeff_small"] + fl["gram_small"],
                        "note": "fused coefficients + Gram blocks of the groups with max nmax <= 8: 94 nmax (+ 14 nmx, not counted) + 16 nmax^2 flop per particle"},
            "k_coeff": {"bound": "fp64 vector pipe (latency-bound recurrences)", "ms": kms.get("k_coeff", 0.0),
                        "achieved": tf(fl["coeff"] - fl["coeff_small"], kms.get("k_coeff", 0.0)), "peak": FP64_DFMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                        "flop": fl["coeff"] - fl["coeff_small"]},
            "k_gram_sum+k_gram_eval": {"bound": "tensor", "ms": kms.get("k_gram_sum_eval", 0.0), "achieved": tf(fl["eval"], kms.get("k_gram_sum_eval", 0.0)),
                                       "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "flop": fl["eval"],
                                       "note": "4 quadratic forms of size N = 8 x (largest Gram class) at the 2 N + 1 Chebyshev nodes per cell + interpolation of the four polynomials to the 371 angles"},
            "k_finalize": {"bound": "hbm", "ms": kms.get("k_finalize", 0.0)},
        }
        for name, k in kernels.items():
            if k.get("achieved") and k.get("peak"):
                k["frac"] = k["achieved"] / k["peak"]
            key = name.split("+")[0] if name != "k_gram_sum+k_gram_eval" else "k_gram_sum_eval"
            if db.get(key) is not None:
                k["dram_bytes_per_step"] = db[key] * tb.cells          # recorded per cell
                if k["ms"]:
                    k["dram_gbs"] = k["dram_bytes_per_step"] / (k["ms"] * 1e-3) / 1e9
        dom = max((n for n in kernels if n != "k_finalize"), key=lambda n: kernels[n]["ms"] or 0.0)
        D = kernels[dom]
        nl = (kms.get("launches") or {}).get({"k_gram_sum+k_gram_eval": "k_gram_sum_eval"}.get(dom, dom), 0) or 1
        roofline = {"bound": "tensor" if D["bound"] == "tensor" else "fp64",
                    "kernel": (dom + " (FP64 DMMA m8n8k4)") if D["bound"] == "tensor" else dom,
                    "achieved": D.get("achieved"), "peak": D.get("peak"), "unit": "TFLOP/s", "frac": D.get("frac"),
                    "executed_frac": (D["executed"] / D["peak"]) if D.get("executed") else D.get("frac"),
                    "launches_per_step": nl, "avg_launch_ms": D["ms"] / nl, "flop_per_launch": (D.get("flop") or 0.0) / nl,
                    "traffic": (D["dram_bytes_per_step"] / nl) if D.get("dram_bytes_per_step") else None,
                    "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (average over the launches of one step), from the ncu "
                                    "pass of this revision's kernels recorded in profiles/r02_dram_bytes.json; algorithmic bytes per cell "
                                    "(SURVEY 8d): nx*20 + N_ang*48 + 240 = %.0f B" % (np.mean([B["nx"] for B in tb.bins]) * 20 + NANG * 48 + 240),
                    "peak_source": PEAK_SOURCE, "kernel_ms_per_step": kms, "kernels": kernels,
                    "step_tflops_survey_flop_model_per_gpu": fl["survey"] / (ms_dev * 1e-3) / 1e12,
                    "step_tflops_per_angle_model_per_gpu": fl["contract"] / (ms_dev * 1e-3) / 1e12,
                    "model_note": "TFLOP/s the reference formulation (SURVEY 8d) / the per-angle S+/S- formulation would need to do this step "
                                  "in the same time (throughput equivalents: the Gram form of the small-nmax groups needs fewer flop)",
                    "kernel_share_of_step": (D["ms"] / ms_dev) if ms_dev else None}
        res = {"value": total_evals / (ms_dev * 1e-3), "unit": UNIT, "ms_per_step": ms_dev, "cells_per_gpu": tb.cells,
               "particle_evals_per_gpu_per_step": tb.evals,
               "e2e": {"value": total_evals / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                       "h2d_bytes_per_step": int(sum(B["mz"].nbytes * 2 + B["par"].nbytes + B["frac"].nbytes for B in tb.bins)),
                       "d2h_bytes_per_step": int(sum(B["scal_h"].nbytes + B["phase_h"].nbytes + B["coef_h"].nbytes for B in tb.bins)),
                       "api": "gm_table_run_psd with the fused GSF stage, one call per bin (host buffers: per-cell m and PSD parameters in, reduced "
                              "sums and GSF moments out)"},
               "gpu_launches": int(launches), "roofline": roofline, "gather": tb.gather_mode}
        if world > 1:
            res["per_rank_ms_per_step"] = {"device": pr_dev, "e2e": pr_e2e}
            per_rank_k = [None] * world
            td.all_gather_object(per_rank_k, {k: round(v, 4) for k, v in kms.items() if k != "launches"})
            res["per_rank_kernel_ms"] = per_rank_k
        results[sp] = res
        tb.close()
    clocks = sampler.stop()
    if world > 1:
        allc = [None] * world
        td.all_gather_object(allc, clocks)
        clocks = dict(allc[0])
        clocks["per_rank_sm_mhz"] = [c.get("sm_mhz") for c in allc]
        clocks["reasons"] = sorted(set(r for c in allc for r in c.get("reasons", [])))
        known = [c["sm_mhz"] for c in allc if c.get("sm_mhz")]
        if known:
            clocks["sm_mhz"] = min(known)              # the slowest GPU of the job

    lut = None
    if not args.no_lut:
        lut = {"what": "wall seconds of runoptics.main + rungsf.main (inputs, kernels, post-processing, file, GSF moments) per table at "
                       "%d GPU(s); zero-weight particles elided as runoptics does by default; median of three runs in the warm process (cold_s = the first run; warm_runs_s = all three)" % world}
        for sp in ("su", "bc", "ss"):
            try:
                lut["optics_" + sp.upper()] = lut_build(sp, world, rank, td, torch)
            except Exception as e:   # noqa: BLE001 -- reported in the line, never hidden
                lut["optics_" + sp.upper()] = {"error": "%s: %s" % (type(e).__name__, e)}

        try:
            lut["optics_SS_2048_wavelengths_to_RRTMG_bands"] = fine_grid_build("ss", 2048, world, rank, td, torch, comm)
            lut["optics_SS_2048_wavelengths_to_RRTMG_bands"]["what"] = (
                "BASELINE config 5: dointegration.fun on a 2048-wavelength grid (368,640 cells, sharded over the ranks) + RRTMG band "
                "averages of the eight variables; the build whose cost is on the GPUs (the 61-wavelength tables above are bound by "
                "host work and file I/O: their kernels take 3-35 ms)")
        except Exception as e:   # noqa: BLE001
            lut["optics_SS_2048_wavelengths_to_RRTMG_bands"] = {"error": "%s: %s" % (type(e).__name__, e)}

    head = order[0]
    H = results[head]
    line = {
        "metric": METRIC, "value": H["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": H["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "%s.json parameters; OPAC + HITRAN water refractive indices (tests/golden/hostlogic.npz); synthetic = the shipped configuration itself" % head,
        "config": {"workload": WORKLOADS[head]["text"], "cells_per_gpu": H["cells_per_gpu"], "nang": NANG,
                   "l2_policy": "inputs larger than L2 per step: the coefficient stream (up to 4 GB per batch), number weights and partial sums are "
                                "re-written every step",
                   "parallelism": "every rank evaluates its own copy of the table (weak scaling), %d rank(s); rows gathered to rank 0: %s"
                                  % (world, {"peer": "copy-engine puts into rank 0's IPC-mapped buffer over NVLink (gm_peer_put), overlapped with the "
                                                     "next kernels, complete inside the timed region", "nccl": "NCCL gather", "none": "none (1 rank)",
                                             "off": "SWITCHED OFF (diagnostic run)"}[H["gather"]])},
        "e2e": H["e2e"], "gpu_launches": H["gpu_launches"], "roofline": H["roofline"], "clocks": clocks,
        "workloads": {WORKLOADS[sp]["name"]: dict(results[sp], config=WORKLOADS[sp]["text"]) for sp in order},
    }
    if lut is not None:
        line["lut_build_s"] = lut
    for k in ("per_rank_ms_per_step", "per_rank_kernel_ms"):
        if k in H:
            line[k] = H[k]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = reference_baseline(head, seconds=float(os.environ.get("GEOSMIE_REF_SECONDS", "12")))
        except Exception as e:   # noqa: BLE001
            line["cpu_baseline"] = port_baseline(head)
            line["cpu_baseline"]["note"] = "reference unavailable (%s): oracle port timed instead" % e
        line["cpu_baseline"]["port"] = port_baseline(head)
        for sp in order[1:]:
            try:
                line["workloads"][WORKLOADS[sp]["name"]]["cpu_baseline"] = reference_baseline(sp, seconds=6.0)
            except Exception as e:   # noqa: BLE001
                line["workloads"][WORKLOADS[sp]["name"]]["cpu_baseline"] = {"error": str(e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        comm.close()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
