#!/usr/bin/env python
"""bench.py — galaxy-band fluxes/s of the photometry path (BASELINE.json metric) on N B200s.

One step = one pass of the hot path (both get_flux passes + totals, src/egg-gencat.cpp:2097-2239)
over one synthetic catalog.  The contract's line is BASELINE.json configs[1]: area=0.08 mmin=7.5
(1.95e5 galaxies), the 30 UV-to-far-IR bands of SURVEY.md §8d, ce01 IR library, reference defaults
(igm=inoue14, emission lines on).  Each rank owns one GPU and its own catalog of that size (weak
scaling, no data-path collective: galaxies are independent).

  value        : kernel-only throughput, inputs and outputs resident in HBM, CUDA events on the launch stream
  e2e          : the same catalog through eggphot_compute (host pinned buffers, H2D + kernels + D2H timed)
  roofline     : HBM roof (contract) with the measured DRAM traffic of the committed ncu capture
  roofline_fp  : the binding roof of SURVEY.md §8d: FLOP per unit recomputed from this run's node counts over the
                 FMA peak measured on this GPU by a micro-benchmark (eggphot_measure_fma_peak)
  other_modes  : short runs (5 steps) of fp64, configs[0], configs[2] and of the node-per-lane kernel
  configs3     : BASELINE.json configs[3] (1.1e8 galaxies x 30 bands over 8 GPUs): this rank's contiguous shard of
                 1.375e7 galaxies, kernel-only and end to end (--workload cfg4 runs it as the main line)
  --impl reference : the CPU restatement of the reference path (oracle, single thread and all host threads) on a
                 bounded sample of the same workload (the reference itself cannot be built here: vif is absent)
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
np.seterr(all="ignore")

NGAL_CFG2 = 195_000
NGAL_CFG4_TOTAL = 110_000_000  # area=100 deg2, mmin=8.0
WORKLOADS = {
    "cfg1": "egg-gencat area=0.08 mmin=9.0 4 HST bands ce01 (configs[0])",
    "cfg2": "egg-gencat area=0.08 mmin=7.5 30 bands ce01 (configs[1])",
    "cfg3": "egg-gencat COSMOS-like, opt_lib hd (re-binned) + cs17, 40 bands + UVJ rfbands (configs[2])",
    "cfg4": "egg-gencat area=100 Euclid-depth, 1.1e8 galaxies x 30 bands ce01, contiguous shards of 1.375e7 galaxies per GPU (configs[3])",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ngal", type=int, default=0, help="galaxies per GPU (default: the workload's own size)")
    ap.add_argument("--precision", default="fp32c", choices=["fp32c", "fp64"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other_modes / configs3 sub-records")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="BASELINE.json configs[0..3]; the contract's bench line is cfg2 (= configs[1], the default)")
    return ap.parse_args()


def default_ngal(workload):
    return NGAL_CFG4_TOTAL // 8 if workload == "cfg4" else NGAL_CFG2


def make_inputs(ngal, seed, workload="cfg2"):
    """-> optical lib, IR lib, galaxy columns, bands, line list, rfbands (all synthetic)"""
    from egg_b200 import synth
    if workload == "cfg3":
        opt, ir = synth.make_opt_lib(seed=3, hd=True), synth.make_cs17_lib()
        gal = synth.draw_galaxies(ngal, opt_lib=opt, ir_lib=ir, mmin=8.0, seed=seed)
        return opt, ir, gal, synth.get_filters(synth.B40), synth.LINE_LAMBDA, synth.get_filters(synth.RF3)
    opt, ir = synth.make_opt_lib(), synth.ce01_lib()
    mmin = {"cfg1": 9.0, "cfg2": 7.5, "cfg4": 8.0}[workload]
    if ngal > 3_000_000:  # draw a large shard in pieces (bounded temporaries)
        parts = [synth.draw_galaxies(min(2_750_000, ngal - f), opt_lib=opt, ir_lib=ir, mmin=mmin, seed=seed * 1000 + k)
                 for k, f in enumerate(range(0, ngal, 2_750_000))]
        gal = {k: (np.concatenate([p[k] for p in parts]) if isinstance(parts[0][k], np.ndarray) else parts[0][k]) for k in parts[0]}
    else:
        gal = synth.draw_galaxies(ngal, opt_lib=opt, ir_lib=ir, mmin=mmin, seed=seed)
    return opt, ir, gal, synth.get_filters(synth.B4 if workload == "cfg1" else synth.B30), synth.LINE_LAMBDA, []


def host_threads():
    """Host threads this process may use.  torch.distributed.run exports OMP_NUM_THREADS=1: the CPU arm must not
    inherit that, so the count comes from the affinity mask and is passed to the oracle explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def time_oracle(opt, ir, bands, lines, gal, nsample, nthreads, reps=1, rfb=()):
    from oracle.oracle import Oracle
    orc = Oracle(opt, ir, bands, rfb, line_lambda=lines)
    sub = {k: (v[:nsample] if hasattr(v, "shape") and v.shape[:1] == gal["z"].shape else v) for k, v in gal.items()}
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        orc.compute(sub, nthreads=nthreads)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return nsample * len(bands) / best, best


def profile_numbers(ngal, workload, precision):
    """Counters of the committed `ncu --set full` capture of this command (profiles/r2_lane_kernel_counters.json):
    DRAM bytes per launch for roofline.traffic, and the pipe utilisations.  None for any other size / workload."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_lane_kernel_counters.json")))
        same = int(t["ngal"]) == int(ngal) and workload == t.get("workload", "cfg2") and precision == t.get("precision", "fp32c")
        return t if same else None
    except Exception:
        return None


def flop_per_unit(ph, bands, z, has_rf=0):
    """SURVEY.md §8d convention: 10 FLOP per node of the merged (filter + SED) grid of either component + 4 FLOP per SED
    node touched, averaged over this run's galaxies; a unit = one (galaxy, band) triple of fluxes."""
    xd, xb = ph.grid("disk"), ph.grid("bulge")
    zs = np.asarray(z, np.float64)
    zs = zs[:: max(1, len(zs) // 4000)]
    opz = 1.0 + zs
    merged = touched = 0.0
    for lam, _ in bands:
        lam = np.asarray(lam, np.float64)
        for x in (xd, xb):
            if len(x) < 2:
                continue
            lo = np.searchsorted(x, lam[0] / opz, side="right")
            hi = np.searchsorted(x, lam[-1] / opz, side="left")
            inside = np.maximum(hi - lo, 0).astype(np.float64)
            covered = (x[0] * opz <= lam[0]) & (x[-1] * opz >= lam[-1])
            merged += float(np.mean(np.where(covered, inside + len(lam), 0.0)))
            touched += float(np.mean(np.where(covered, inside, 0.0)))
    nb = max(len(bands), 1)
    return (10.0 * merged + 4.0 * touched) / nb, merged / nb, touched / nb


class ClockSampler:
    def __init__(self, dev):
        self.dev, self.samples, self.reasons, self.stop = dev, [], set(), False
        self.max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(dev)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4,
                 "hw_power_brake": 0x80}
        while not self.stop and nv:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_reference(a, rank, world):
    """--impl reference: the CPU path on this box's host cores, bounded sample per step (rank 0 only)."""
    if rank != 0:
        return
    nthreads = host_threads()
    n = a.ngal or default_ngal(a.workload)
    opt, ir, gal, bands, lines, rfb = make_inputs(min(n, 195_000), 42, a.workload)
    nb = len(bands)
    # size the per-step sample so that warm-up + steps stay within ~2 minutes; the whole catalog of the step when it fits
    rate, _ = time_oracle(opt, ir, bands, lines, gal, 4000, nthreads, rfb=rfb)
    budget = 120.0 / max(a.steps + a.warmup, 1)
    nsample = int(max(1000, min(len(gal["z"]), rate / nb * min(budget, a.cpu_seconds))))
    for _ in range(a.warmup):
        time_oracle(opt, ir, bands, lines, gal, nsample, nthreads, rfb=rfb)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        time_oracle(opt, ir, bands, lines, gal, nsample, nthreads, rfb=rfb)
    dt = (time.perf_counter() - t0) / a.steps
    val = nsample * nb / dt
    n1 = int(max(500, min(nsample, rate / nthreads / nb * 6.0)))
    rate1, _ = time_oracle(opt, ir, bands, lines, gal, n1, 1, rfb=rfb)
    line = {"impl": "reference", "metric": "galaxy-band fluxes/sec", "value": val, "unit": "galaxy-band fluxes/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[a.workload], "nband": nb, "sample_galaxies_per_step": nsample,
                       "same_catalog_size_as_gpu_arm": bool(nsample == n),
                       "note": "CPU restatement of the reference path (oracle port; vif-based reference cannot be built)"},
            "cpu_baseline": {"value": val, "unit": "galaxy-band fluxes/s", "cores": nthreads, "kind": "port",
                             "sample": f"{nsample} galaxies x {nb} bands per step, OpenMP over galaxies, {nthreads} threads",
                             "single_thread": {"value": rate1, "sample": f"{n1} galaxies x {nb} bands"}},
            "e2e": {"value": val, "unit": "galaxy-band fluxes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


class Runner:
    """One workload on this rank's GPU: device-resident and host-resident timing of the same catalog."""

    def __init__(self, a, workload, precision, ngal, rank, world, local, dev):
        import torch
        import egg_b200
        self.torch, self.a, self.world, self.dev, self.local = torch, a, world, dev, local
        self.workload, self.precision, self.n = workload, precision, ngal
        self.opt, self.ir, self.gal, self.bands, self.lines, self.rfb = make_inputs(ngal, 42 + rank, workload)
        self.nb = len(self.bands)
        self.ph = egg_b200.Photometry(self.opt, self.ir, self.bands, self.rfb, line_lambda=self.lines, precision=precision, device=local)
        cols = ["z", "d", "m_disk", "m_bulge", "m2lcor", "lir", "igm_abs", "vdisp", "line_lum_disk", "line_lum_bulge",
                "opt_sed_disk", "opt_sed_bulge", "ir_sed"] + (["fpah", "mdust"] if workload == "cfg3" else [])
        self.cols = cols
        gal = self.gal
        self.host = {k: torch.from_numpy(np.ascontiguousarray(gal[k]).view(np.int64) if gal[k].dtype == np.uint64
                                         else np.ascontiguousarray(gal[k])).pin_memory() for k in cols}
        self.devt = {k: v.to(dev) for k, v in self.host.items()}
        n, nb, rfb = self.n, self.nb, self.rfb
        self.outs = {k: torch.empty((n, nb), dtype=torch.float32, device=dev) for k in ("flux", "flux_disk", "flux_bulge")}
        self.outs.update({k: torch.empty((n, len(rfb)), dtype=torch.float32, device=dev) for k in ("rfmag", "rfmag_disk", "rfmag_bulge") if rfb})
        self.in_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.out_bytes = sum(v.numel() * v.element_size() for v in self.outs.values())

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def kernel(self, steps, warmup, flush, clocks=False):
        """-> ms per step (max over ranks), launches per step, wall ms per step, clock summary"""
        torch = self.torch
        gp = {k: v.data_ptr() for k, v in self.devt.items()}
        op = {k: v.data_ptr() for k, v in self.outs.items()}
        stream = torch.cuda.current_stream()

        def step():
            flush.zero_()  # evict L2 between timed iterations (untimed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            self.ph.compute_device(gp, self.n, op, stream=stream.cuda_stream)
            e1.record(stream)
            return e0, e1

        for _ in range(max(warmup, 3)):
            step()
        self.barrier()
        launches = self.ph.stats()["kernel_launches"]  # (also clears the library's walk-kernel timers before the timed steps)
        clk = ClockSampler(self.local) if clocks else None
        if clk:
            clk.__enter__()
        self.barrier()
        t0 = time.perf_counter()
        evs = [step() for _ in range(steps)]
        self.barrier()
        wall = time.perf_counter() - t0
        if clk:
            clk.__exit__()
        self.ph.synchronize()
        # the walk kernel alone (the dominant kernel): CUDA events around its launches on the launching stream, inside the library
        self.walk_ms = float(self.ph.stats().get("last_kernel_ms", 0.0))
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, launches, wall * 1e3 / steps, (clk.summary() if clk else None)

    def e2e(self, steps):
        torch = self.torch
        hgal = {k: (self.host[k].numpy().view(np.uint64) if self.host[k].dtype == torch.int64 else self.host[k].numpy()) for k in self.cols}
        hout = {k: torch.empty(tuple(v.shape), dtype=torch.float32).pin_memory().numpy() for k, v in self.outs.items()}
        for _ in range(2):
            self.ph.compute(hgal, out=hout)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.ph.compute(hgal, out=hout)
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return {"value": self.world * self.n * self.nb / float(dt.item()), "unit": "galaxy-band fluxes/s",
                "h2d_bytes_per_step": int(self.in_bytes), "d2h_bytes_per_step": int(self.out_bytes),
                "ms_per_step": float(dt.item()) * 1e3}

    def close(self):
        self.ph.close()
        self.devt = self.outs = self.host = None
        self.torch.cuda.empty_cache()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the photometry path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints one JSON line
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    n = a.ngal or default_ngal(a.workload)
    R = Runner(a, a.workload, a.precision, n, rank, world, local, dev)
    nb = R.nb
    ms_per_step, launches, wall_ms, clocks = R.kernel(a.steps, a.warmup, flush, clocks=True)
    value = world * n * nb / (ms_per_step * 1e-3)
    walk_ms = R.walk_ms
    e2e = None if a.no_e2e else R.e2e(a.steps)
    fpu, merged, touched = flop_per_unit(R.ph, R.bands, R.gal["z"])
    peak = R.ph.fma_peak()
    stats = R.ph.stats()
    in_bytes, out_bytes = R.in_bytes, R.out_bytes
    cpu = None
    if rank == 0 and not a.no_cpu:
        nthreads = host_threads()
        rate1, _ = time_oracle(R.opt, R.ir, R.bands, R.lines, R.gal, 2000, nthreads, rfb=R.rfb)
        ns = int(max(2000, min(n, rate1 / nb * a.cpu_seconds)))
        rate, secs = time_oracle(R.opt, R.ir, R.bands, R.lines, R.gal, ns, nthreads, rfb=R.rfb)
        cpu = {"value": rate, "unit": "galaxy-band fluxes/s", "cores": nthreads, "kind": "port",
               "sample": f"first {ns} galaxies of the step's catalog x {nb} bands, {secs:.1f} s, OpenMP over galaxies, {nthreads} threads"}
    R.close()

    # ---- sub-records: the other arithmetic mode, the other configs, the other kernel, and configs[3]
    extra, cfg4 = {}, None
    if not a.no_extra:
        def short(workload, precision, ngal, steps=5, env=None, e2e_steps=0):
            old = {}
            for k, v in (env or {}).items():
                old[k] = os.environ.get(k)
                os.environ[k] = v
            try:
                r = Runner(a, workload, precision, ngal, rank, world, local, dev)
                ms, _, _, _ = r.kernel(steps, 3, flush)
                rec = {"workload": WORKLOADS[workload], "precision": precision, "ngal_per_gpu": ngal, "nband": r.nb,
                       "nrfband": len(r.rfb), "steps": steps, "ms_per_step": ms, "value": world * ngal * r.nb / (ms * 1e-3),
                       "unit": "galaxy-band fluxes/s", "ngroups": r.ph.stats()["ngroups"]}
                if e2e_steps:
                    rec["e2e"] = r.e2e(e2e_steps)
                r.close()
                return rec
            except Exception as ex:  # a sub-record must not take the contract's line down with it
                return {"workload": WORKLOADS[workload], "precision": precision, "error": repr(ex)[:300]}
            finally:
                for k, v in old.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v

        if a.workload == "cfg2":
            other = "fp64" if a.precision == "fp32c" else "fp32c"
            extra[other] = short("cfg2", other, n)
            extra["configs0_b4"] = short("cfg1", a.precision, n)
            extra["configs2_b40_cs17_rf3"] = short("cfg3", a.precision, n)
            extra["node_per_lane_kernel"] = short("cfg2", a.precision, n, env={"EGGPHOT_KERNEL": "node"})
            cfg4 = short("cfg4", a.precision, default_ngal("cfg4"), steps=2, e2e_steps=2)
            if "value" in cfg4:
                cfg4["ngal_total"] = world * default_ngal("cfg4")
                cfg4["note"] = ("every rank computes its own contiguous shard of 1.375e7 galaxies (drawn with its own seed); at 8 GPUs "
                                "this is configs[3] in full, 1.1e8 galaxies x 30 bands")

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # algorithmic HBM bytes per unit: 3 f32 outputs + per-galaxy inputs / nband
        bytes_per_unit = out_bytes / n / nb + in_bytes / n / nb
        # the contract's roofline is that of the dominant kernel: algorithmic bytes of a launch / its own duration, measured live
        dom_ms = walk_ms if walk_ms > 0 else ms_per_step
        achieved = n * nb * bytes_per_unit / (dom_ms * 1e-3) / 1e9
        achieved_step = n * nb * bytes_per_unit / (ms_per_step * 1e-3) / 1e9
        prof = profile_numbers(n, a.workload, a.precision)
        fp_peak = peak["fp32_tflops"] if a.precision == "fp32c" else peak["fp64_tflops"]
        fp_ach = value / world * fpu / 1e12
        line = {"metric": "galaxy-band fluxes/sec", "value": value, "unit": "galaxy-band fluxes/s", "n_gpus": world,
                "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if a.precision == "fp32c" else "f64",
                "data": "synthetic",
                "config": {"workload": WORKLOADS[a.workload], "ngal_per_gpu": n, "nband": nb, "nrfband": len(R.rfb),
                           "precision": a.precision + (" (SEDs and products f32, quadrature weights f64)" if a.precision == "fp32c" else ""),
                           "kernel": "node-per-lane" if os.environ.get("EGGPHOT_KERNEL", "") in ("node", "1") else "lane-per-galaxy",
                           "ngroups": stats["ngroups"],
                           "l2": "flushed between timed iterations (512 MiB memset)", "lines": True, "igm": "inoue14"},
                "e2e": e2e, "gpu_launches": int(launches * a.steps),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                             "frac": achieved / hbm_peak,
                             "kernel": "lane_walk_kernel" if walk_ms > 0 else "whole step",
                             "kernel_ms": dom_ms, "kernel_share_of_step": dom_ms / ms_per_step,
                             "achieved_whole_step": achieved_step, "frac_whole_step": achieved_step / hbm_peak,
                             "traffic": float(prof["dominant_kernel_dram_bytes"]) if prof and "dominant_kernel_dram_bytes" in prof else None,
                             "traffic_whole_step": float(prof["dram_bytes_per_launch"]) if prof else None,
                             "traffic_frac_of_peak": (float(prof["dram_bytes_per_launch"]) / (ms_per_step * 1e-3) / 1e9 / hbm_peak) if prof else None,
                             "bytes_per_unit": bytes_per_unit,
                             "algorithmic_bytes_per_launch": n * nb * bytes_per_unit,
                             "peak_source": "MEASURED_PEAKS.json (hbm_gbs, measured copy)" if peaks else "fallback",
                             "note": "kernel = the walk kernel (the quadrature), timed live with CUDA events around its launches; algorithmic "
                                     "bytes are the catalog columns in and out (SURVEY.md §8d). The measured traffic (ncu, "
                                     "profiles/r2_lane_kernel_counters.json) is far larger: the rest-frame SED rows [node][32 galaxies] of "
                                     "every batch are written once by the build kernel and read by every band table whose window covers "
                                     "them, and the batches in flight do not fit the L2 (DESIGN.md §5, §10). Neither figure binds: "
                                     "the walk kernel is bound by instruction issue and the shared-memory pipe (`pipes`), the FP roofline is "
                                     "roofline_fp"},
                "roofline_fp": {"bound": "cuda-core FMA pipes (SURVEY.md §8d)", "flop_per_unit": fpu,
                                "merged_nodes_per_unit": merged, "sed_nodes_per_unit": touched,
                                "achieved": fp_ach, "peak": fp_peak, "unit": "TFLOP/s", "frac": fp_ach / fp_peak if fp_peak else None,
                                "peak_fp32_tflops": peak["fp32_tflops"], "peak_fp64_tflops": peak["fp64_tflops"],
                                "peak_source": "eggphot_measure_fma_peak: independent FMA chains on every SM, best of five, this run",
                                "convention": "10 FLOP per node of the merged filter+SED grid of either component + 4 per SED node, "
                                              "node counts of this run's catalog; the kernel itself spends far fewer (one weight per "
                                              "SED node whatever the filter's node count), in double"},
                "pipes": ({k: prof[k] for k in ("issue_active_pct", "lsu_pipe_pct", "smem_bank_conflict_frac", "fp64_pipe_pct",
                                                 "fma_pipe_pct", "alu_pipe_pct", "source") if k in prof} if prof else None),
                "cpu_baseline": cpu, "clocks": clocks, "wall_ms_per_step": wall_ms,
                "other_modes": extra or None, "configs3": cfg4}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
