#!/usr/bin/env python
"""Benchmark of the vertical forward operator on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A *step* is one pass of the hot path over one batch: BASELINE.json configs[1] -- one synthetic
Chapman day profile, X-mode, 174 sounding frequencies 0.1..17.4 MHz, n_points = 20000 -- per GPU.
``value`` is virtual heights (profile x frequency outputs, NaN rows included) per second with the
inputs resident in HBM, timed with CUDA events around every step (L2 flushed between steps);
``e2e`` is the same metric through the numpy drop-in ``vertical_forward_operator`` (host buffers,
H2D + kernel + D2H inside the timed region).  ``--impl reference`` times the reference's CPU
algorithm (numpy restatement, one process per host core) on the same workload.
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "virtual heights/sec (profile x freq) X-mode n=20000"
UNIT = "vh/s"
MODE = "X"
N_POINTS = 20000
FLOPS_PER_POINT = 77          # algorithmic FP64 operations per grid point of a live row


def workload(rank=0):
    from pyrayhf_b200 import synth
    den, bmag, bpsi, alt = synth.bench_day_profile(rank=rank)
    return synth.default_freq(), den, bmag, bpsi, alt


def workload_name():
    return ("configs[1]: single synthetic Chapman day profile (lat 4.5, lon 0 [+1 deg per rank], dipole B), "
            "X-mode, 174 freqs 0.1-17.4 MHz, n_points=20000, 620 altitudes")


def algorithmic_flops(vh, den):
    """W = sum over rows of W_vh: 77 N + 4 At + 8 live, 4 At dead."""
    at = int(np.argmax(den))
    live = int(np.isfinite(vh).sum())
    dead = vh.size - live
    return live * (FLOPS_PER_POINT * N_POINTS + 4 * at + 8) + dead * 4 * at, live, at


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ----------------------------------------------------------------------------------------------
def cpu_numpy_port(steps, warmup, budget_s=150.0):
    """Times the numpy port on every host core.  Returns dict(value, cores, sample, ms_per_step, vh)."""
    from oracle.cpu_baseline import NumpyPortPool
    freq, den, bmag, bpsi, alt = workload(0)
    pool = NumpyPortPool()
    try:
        # size the per-step sample from one probing pass over 29 of the 174 rows
        wall, _, _ = pool.one_pass(freq[::6], den, bmag, bpsi, alt, MODE, N_POINTS)
        wall, rows, _ = pool.one_pass(freq[::6], den, bmag, bpsi, alt, MODE, N_POINTS)
        per_row = wall / (rows / pool.cores)
        target = max(0.25, min(per_row * freq.size, budget_s / max(steps + warmup, 1)))
        n_rows = int(max(6, min(freq.size, target / per_row)))
        sel = np.round(np.linspace(0, freq.size - 1, n_rows)).astype(int)
        fsel = freq[sel]
        for _ in range(warmup):
            pool.one_pass(fsel, den, bmag, bpsi, alt, MODE, N_POINTS)
        total_wall, total_rows, vh = 0.0, 0, None
        for _ in range(steps):
            wall, rows, vh = pool.one_pass(fsel, den, bmag, bpsi, alt, MODE, N_POINTS)
            total_wall += wall
            total_rows += rows
    finally:
        pool.close()
    return dict(value=total_rows / total_wall, cores=pool.cores, ms_per_step=1e3 * total_wall / steps,
                sample="%d of the 174 frequency rows of the workload per process per step, %d processes "
                       "(one per host core), %d steps" % (n_rows, pool.cores, steps),
                rows=sel, vh=vh)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    warnings.simplefilter("ignore")
    r = cpu_numpy_port(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name()},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": r["sample"],
                         "what": "numpy restatement of PyRayHF.library.vertical_forward_operator "
                                 "(oracle/vfo_oracle.py, bit-identical to the reference in the dev container)"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """In-process NVML sampling of SM clock and throttle reasons (called between timed steps)."""

    def __init__(self, index):
        self.ok = False
        self.sm, self.reasons = [], set()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {
                "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
            }
            for k, bit in names.items():
                if mask & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def summary(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(self.max),
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def next_rows_section(torch, dev, stream, ctx, freq, den, bmag, bpsi, alt, with_cpu):
    """Throughput of the operators beside the headline path: the standalone regrid stage
    (HBM-bound), the batched Snell's-law tracers and the inversion residual.  Device-resident inputs, CUDA events
    on the launching stream, L2 flushed by the caller's buffer between repetitions is not needed here: every
    repetition streams more than L2 (regrid) or is compute-bound (tracers)."""
    import ctypes
    vp = ctypes.c_void_p
    L = ctx.lib
    sp = vp(stream.cuda_stream)

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ev = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            ev.append((a, b))
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in ev]))

    t = lambda v: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev)   # noqa: E731
    out = {}
    # ---- regrid_to_nonuniform_grid: five [F x N] arrays written ----
    n_freq, n_pts = freq.size, N_POINTS
    tf, td, tb, tp, ta = t(freq * 1e6), t(den), t(bmag), t(bpsi), t(alt)
    hc = torch.empty(n_freq, dtype=torch.float64, device=dev)
    big = [torch.empty((n_freq, n_pts), dtype=torch.float64, device=dev) for _ in range(5)]
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    ms = timed(lambda: ctx.check(L.prhf_regrid_f64(ctx.handle, vp(tf.data_ptr()), n_freq, vp(td.data_ptr()),
                                                   vp(tb.data_ptr()), vp(tp.data_ptr()), vp(ta.data_ptr()), alt.size, 1,
                                                   n_pts, vp(hc.data_ptr()), *[vp(b.data_ptr()) for b in big],
                                                   vp(st.data_ptr()), sp)))
    hbm_peak = measured_peaks().get("hbm_gbs", 6544.7)
    gbs = 5 * n_freq * n_pts * 8 / (ms * 1e-3) / 1e9
    out["regrid_stage"] = {"workload": "regrid_to_nonuniform_grid, %d freqs x %d points, 5 arrays written" % (n_freq, n_pts),
                           "ms": ms, "write_GBps": gbs, "frac_of_measured_hbm": gbs / hbm_peak,
                           "note": "row-setup kernel + write kernel; bound = HBM writes (MEASURED_PEAKS.json hbm_gbs)"}
    # ---- elementwise stages on 16 Mi elements: find_X (1 read + 1 write), find_mu_mup (3 reads + 2 writes) ----
    n_el = 1 << 24
    ex = torch.rand(n_el, dtype=torch.float64, device=dev) * 0.9
    ey = torch.rand(n_el, dtype=torch.float64, device=dev) * 0.5 + 0.05
    ep = torch.rand(n_el, dtype=torch.float64, device=dev) * 90.0
    eo1, eo2 = torch.empty_like(ex), torch.empty_like(ex)
    fscal = torch.full((1,), 5e6, dtype=torch.float64, device=dev)
    eden = ex * 1e11
    ms = timed(lambda: ctx.check(L.prhf_find_x_f64(ctx.handle, vp(eden.data_ptr()), 1, vp(fscal.data_ptr()), 0, n_el,
                                                   vp(eo1.data_ptr()), None, sp)))
    gbs = 2 * n_el * 8 / (ms * 1e-3) / 1e9
    out["find_X_stage"] = {"elements": n_el, "ms": ms, "GBps": gbs, "frac_of_measured_hbm": gbs / hbm_peak,
                           "note": "bit-identical to numpy: one IEEE sqrt and one IEEE division per element keep it "
                                   "FP64-bound (the fused operator uses reciprocal seeds instead)"}
    ms = timed(lambda: ctx.check(L.prhf_mu_mup_f64(ctx.handle, vp(ex.data_ptr()), vp(ey.data_ptr()), vp(ep.data_ptr()),
                                                   n_el, 1, 0, 0, vp(eo1.data_ptr()), vp(eo2.data_ptr()), sp)))
    gbs = 5 * n_el * 8 / (ms * 1e-3) / 1e9
    out["find_mu_mup_stage"] = {"elements": n_el, "ms": ms, "GBps": gbs, "frac_of_measured_hbm": gbs / hbm_peak,
                                "note": "sincos + Appleton-Hartree per element: FP64-bound, not HBM-bound"}
    del ex, ey, ep, eo1, eo2, eden
    # ---- Snell tracers: 174 frequencies x 64 elevations over the same profile ----
    elev = np.linspace(5.0, 88.0, 64)
    f_r = np.repeat(freq * 1e6, elev.size)
    e_r = np.tile(elev, freq.size)
    t_f, t_e = t(f_r), t(e_r)
    scal = torch.empty((f_r.size, 5), dtype=torch.float64, device=dev)
    npth = torch.zeros(f_r.size, dtype=torch.int32, device=dev)
    for geo, name in ((0, "cartesian"), (1, "spherical")):
        ms = timed(lambda: ctx.check(L.prhf_snell_f64(ctx.handle, vp(t_f.data_ptr()), vp(t_e.data_ptr()), f_r.size,
                                                      vp(ta.data_ptr()), vp(td.data_ptr()), vp(tb.data_ptr()),
                                                      vp(tp.data_ptr()), alt.size, 1, geo, 0, 1.0, 200.0, 400, 6371.0,
                                                      vp(scal.data_ptr()), None, None, 0, vp(npth.data_ptr()), sp)))
        entry = {"rays": int(f_r.size), "rays_with_a_path": int((npth > 0).sum().item()), "ms": ms,
                 "rays_per_s": f_r.size / (ms * 1e-3)}
        if with_cpu:
            from oracle import snell_oracle
            idx = np.linspace(0, f_r.size - 1, 12).astype(int)
            t0 = time.perf_counter()
            for i in idx:
                snell_oracle.trace(f_r[i], e_r[i], alt, den, bmag, bpsi, 'X', name)
            cpu = (time.perf_counter() - t0) / idx.size
            entry["cpu_port_rays_per_s_1_core"] = 1.0 / cpu
        out["snell_" + name] = entry
    # ---- inversion residual: chi2 of 4096 candidate curves ----
    vm = torch.rand((4096, n_freq), dtype=torch.float64, device=dev) * 300 + 100
    vo = torch.rand(n_freq, dtype=torch.float64, device=dev) * 300 + 100
    chi = torch.empty(4096, dtype=torch.float64, device=dev)
    ms = timed(lambda: ctx.check(L.prhf_residual_f64(ctx.handle, vp(vm.data_ptr()), vp(vo.data_ptr()), 4096, n_freq,
                                                     None, vp(chi.data_ptr()), sp)))
    out["residual_chi2"] = {"candidates": 4096, "ms": ms, "read_GBps": 4096 * n_freq * 8 / (ms * 1e-3) / 1e9}
    return out


def run_b200_arm(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import pyrayhf_b200
    from pyrayhf_b200 import _cabi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run "
                             "(--nproc-per-node %d)" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION/INFO; stdout carries exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    freq, den, bmag, bpsi, alt = workload(rank)
    vp = ctypes.c_void_p
    ctx = _cabi.context(local)
    t_freq, t_den, t_b, t_psi, t_alt = (torch.from_numpy(np.ascontiguousarray(v)).to(dev)
                                         for v in (freq, den[None], bmag[None], bpsi[None], alt))
    out = torch.empty((1, freq.size), dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step_device():
        rc = ctx.lib.prhf_vfo_f64(ctx.handle, vp(t_freq.data_ptr()), freq.size, 0, vp(t_den.data_ptr()),
                                  vp(t_b.data_ptr()), vp(t_psi.data_ptr()), vp(t_alt.data_ptr()), 0, 1, alt.size,
                                  1, N_POINTS, 0, vp(out.data_ptr()), vp(status.data_ptr()), vp(stream.cuda_stream))
        ctx.check(rc)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    clocks = ClockSampler(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: events around every step, L2 flushed between steps ----
    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    barrier()
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        if k % 4 == 0:
            clocks.sample()
    barrier()
    launches = ctx.launch_count() - launches0
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(step_ms.sum())
    vh_dev = out.cpu().numpy()[0]

    # ---- end to end through the public numpy API (host buffers in, host buffer out) ----
    for _ in range(args.warmup):
        pyrayhf_b200.vertical_forward_operator(freq, den, bmag, bpsi, alt, MODE, N_POINTS, device=local)
    barrier()
    e2e_s = 0.0
    vh_e2e = None
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vh_e2e = pyrayhf_b200.vertical_forward_operator(freq, den, bmag, bpsi, alt, MODE, N_POINTS, device=local)
        e2e_s += time.perf_counter() - t0
        if k % 4 == 0:
            clocks.sample()
    barrier()

    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0].item()), float(t[1].item())
    units = freq.size * world                      # virtual heights per step, all ranks
    value = units * args.steps / (total_ms * 1e-3)
    e2e_value = units * args.steps / e2e_s

    # ---- the dominant kernel alone (roofline): measurement mode of the C ABI puts CUDA events around the
    #      row-setup kernel and the tile kernel of every call (L2 flushed between calls as above) ----
    ctx.kernel_timing(True)
    n_kt = max(5, min(args.steps, 20))
    for _ in range(n_kt):
        flush.zero_()
        step_device()
    rows_ms, tile_ms, pairs = ctx.kernel_timing(False)
    torch.cuda.synchronize()

    # ---- batched form of the same operator (extra, not the headline): 512 profiles, X-mode, n=20000 ----
    batched = None
    if rank == 0 and not args.no_batched:
        from pyrayhf_b200 import synth
        lat, lon = synth.grid_subset(512)
        bden, bb, bpsi2 = synth.profiles_at(lat, lon, alt)
        tb = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, bden, bb, bpsi2, alt)]
        bout = torch.empty((512, freq.size), dtype=torch.float64, device=dev)
        for _ in range(3):
            pyrayhf_b200.vertical_forward_operator_batched(*tb, MODE, N_POINTS, out=bout, errors='nan')
        torch.cuda.synchronize()
        bev = []
        for _ in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            pyrayhf_b200.vertical_forward_operator_batched(*tb, MODE, N_POINTS, out=bout, errors='nan')
            b.record(stream)
            bev.append((a, b))
        torch.cuda.synchronize()
        b_ms = float(np.mean([a.elapsed_time(b) for a, b in bev]))
        ctx.kernel_timing(True)
        for _ in range(3):
            pyrayhf_b200.vertical_forward_operator_batched(*tb, MODE, N_POINTS, out=bout, errors='nan')
        _, b_tile_ms, b_pairs = ctx.kernel_timing(False)
        bvh = bout.cpu().numpy()
        b_live = int(np.isfinite(bvh).sum())
        b_flops = sum(algorithmic_flops(bvh[q], bden[q])[0] for q in range(512))
        for _ in range(2):      # first calls size the pinned / device arena and capture the graph
            pyrayhf_b200.vertical_forward_operator_batched(freq, bden, bb, bpsi2, alt, MODE, N_POINTS, errors='nan')
        t0 = time.perf_counter()
        for _ in range(3):
            pyrayhf_b200.vertical_forward_operator_batched(freq, bden, bb, bpsi2, alt, MODE, N_POINTS, errors='nan')
        b_e2e = (time.perf_counter() - t0) / 3
        batched = {"workload": "512 synthetic profiles (seeded subset of the 1-degree grid), X-mode, 174 freqs, n_points=20000",
                   "value": 512 * freq.size / (b_ms * 1e-3), "unit": UNIT, "ms_per_step": b_ms, "live_rows": b_live,
                   "grid_points_per_s": b_live * N_POINTS / (b_ms * 1e-3),
                   "tile_kernel_ms": b_tile_ms / max(b_pairs, 1),
                   "tile_kernel_tflops": b_flops / (b_tile_ms / max(b_pairs, 1) * 1e-3) / 1e12,
                   "e2e_value": 512 * freq.size / b_e2e, "e2e_ms_per_step": 1e3 * b_e2e,
                   "e2e_h2d_bytes_per_step": int((freq.size + alt.size + 3 * 512 * alt.size) * 8),
                   "e2e_d2h_bytes_per_step": int(512 * freq.size * 8 + 512 * 4)}

    if rank == 0:
        flops, live, at = algorithmic_flops(vh_dev, den)
        kernel_ms = tile_ms / max(pairs, 1)               # the tile kernel alone, CUDA events on its stream
        peak_tf = ctx.measure_fp64_peak()
        achieved_tf = flops / (kernel_ms * 1e-3) / 1e12
        hbm_bytes = 8 * freq.size + (3 * alt.size + alt.size + freq.size) * 8
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(), "n_profiles_per_gpu": 1, "n_freq": int(freq.size),
                       "n_points": N_POINTS, "mode": MODE, "live_rows": live, "truncated_levels": at,
                       "grid_points_per_s": live * N_POINTS * world / (total_ms / args.steps * 1e-3),
                       "l2": "flushed between timed steps (256 MiB write); every step timed with its own CUDA events"},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic,
                         "kernel": ("prhf::vfo_solo_kernel<1,false> (row setup + grid points of a tile in one launch)"
                                    if launches == args.steps else "prhf::vfo_tile_kernel<1,false>"),
                         "kernel_ms": kernel_ms,
                         "rows_kernel_ms": rows_ms / max(pairs, 1),
                         "step_frac": (flops / (total_ms / args.steps * 1e-3) / 1e12) / peak_tf,
                         "algorithmic_flops_per_launch": flops,
                         "note": "achieved = W (77 flop per grid point of a reflecting row + 4 per profile level per row, "
                                 "SURVEY 8d) / duration of the dominant kernel from CUDA events on its stream "
                                 "(prhf_kernel_timing); step_frac uses the whole timed step",
                         "peak_source": "prhf_measure_fp64_peak: dependent-free DFMA kernel timed live on this GPU "
                                        "(MEASURED_PEAKS.json has no FP64 entry)",
                         "hbm_algorithmic_bytes_per_launch": hbm_bytes,
                         "hbm_frac_of_measured": (hbm_bytes / (kernel_ms * 1e-3) / 1e9) /
                                                 measured_peaks().get("hbm_gbs", 6650.0)},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int((freq.size + 4 * alt.size) * 8),
                    "d2h_bytes_per_step": int(freq.size * 8 + 4),
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "pyrayhf_b200.vertical_forward_operator(numpy...) -> prhf_vfo_host_f64"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if batched:
            batched["tile_kernel_frac_of_fp64_peak"] = batched["tile_kernel_tflops"] / peak_tf
            line["batched"] = batched
        if world == 1 and not args.no_batched:
            warnings.simplefilter("ignore")
            line["next_rows"] = next_rows_section(torch, dev, stream, ctx, freq, den, bmag, bpsi, alt,
                                                  with_cpu=not args.no_cpu_baseline)
        if world == 1 and not args.no_cpu_baseline:
            warnings.simplefilter("ignore")
            cb = cpu_numpy_port(steps=2, warmup=0, budget_s=30.0)
            ref_rows = cb["vh"]
            got = vh_e2e[cb["rows"]]
            mask_mismatch = int(np.sum(np.isnan(got) != np.isnan(ref_rows)))
            m = np.isfinite(ref_rows) & np.isfinite(got)
            relerr = float(np.max(np.abs(got[m] - ref_rows[m]) / np.abs(ref_rows[m]))) if m.any() else 0.0
            line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port",
                                    "sample": cb["sample"]}
            try:
                from oracle.cpu_baseline import c_port_rate
                rate, cores = c_port_rate(freq, den, bmag, bpsi, alt, MODE, N_POINTS)
                line["cpu_baseline"]["c_port_value"] = rate
                line["cpu_baseline"]["c_port_note"] = "scalar C restatement, %d pthreads (extra, not the baseline)" % cores
            except Exception as exc:       # the C oracle is optional here
                line["cpu_baseline"]["c_port_note"] = "unavailable: %s" % exc
            line["parity"] = {"rows_checked": int(ref_rows.size), "nan_mask_mismatches": mask_mismatch,
                              "max_rel_err_vs_numpy_port": relerr, "tolerance": 1e-9}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """Native libraries (NCCL's version banner, for one) print to file descriptor 1.  The contract is ONE JSON line
    on stdout, so everything written to fd 1 while the benchmark runs goes to stderr; `emit` writes the line to the
    real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    text = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, text)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
