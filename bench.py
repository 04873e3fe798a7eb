#!/usr/bin/env python
"""Benchmark of the vertical forward operator on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path over ONE FIXED batch, the unit BASELINE.json configs[3] is made of: 8 members
of the data-assimilation ensemble x 8 192 perturbed Chapman + dipole profiles = 65 536 profiles x 174 sounding
frequencies, X-mode, n_points = 20 000.  The batch is sharded by profile over the N ranks (one process per GPU,
interleaved) through ``pyrayhf_b200.sharding.ShardedForwardOperator``; every rank holds only its shard and the
``[65 536 x 174]`` result is gathered into ONE page-locked host buffer on rank 0 INSIDE the timed region
(``"scaling": "strong"``: the total work is fixed as N grows).

``value``  virtual heights (profile x frequency outputs, NaN rows included) per second with the shards resident in
           HBM when the timed region starts (built on the device from 40 bytes of layer parameters per profile);
           every step timed with CUDA events on the launching stream, L2 flushed between steps, max over ranks.
``e2e``    the same batch through the same public call with HOST arrays: every rank's [P/N x 620] x 3 inputs start
           in page-locked host memory, are copied to its GPU inside the step (pipelined with the kernels) and the
           result lands in rank 0's host buffer; wall clock, max over ranks.
``latency`` BASELINE.json configs[1] (ONE profile, 174 frequencies, X-mode, n = 20 000) through the numpy drop-in
           ``vertical_forward_operator`` -- the single-call story, with its own roofline.
``--impl reference`` times the reference's own ``vertical_forward_operator`` (unmodified copy in ``oracle/_ref``,
           else the numpy port) on the box's host cores on a bounded sample of the same batch.
"""
import argparse
import importlib.util
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
MEMBERS = 8                   # ensemble members per step
BASE_PROFILES = 8192          # perturbed profiles per member (BASELINE.json configs[3])
N_PROFILES = MEMBERS * BASE_PROFILES
LAYOUT = "interleaved"
FLOPS_PER_POINT = 77          # algorithmic FP64 operations per grid point of a row that reflects
FLOPS_PER_CLAMPED_ROW = 77    # rows that reflect at/below the first level: ONE evaluation of mu' (the weights telescope)


def load_synth():
    """The synthetic-input generator, loaded BY PATH: the reference arm must not import the product package
    (``pyrayhf_b200/__init__`` loads the CUDA extension); ``synth.py`` needs numpy only."""
    name = "_prhf_bench_synth"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "pyrayhf_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def workload_parameters(synth):
    """[65 536 x 5] layer parameters {foF2, hmF2, H, foE, latitude}: members 0..7 of the config-4 ensemble over the
    first 8 192 profiles of the seeded shuffle of the 1-degree grid (SURVEY.md 8d "Config 4")."""
    lat, lon = synth.grid_subset(BASE_PROFILES)
    rows = [np.stack(synth.ensemble_member_parameters(lat, lon, m), axis=1) for m in range(MEMBERS)]
    return np.ascontiguousarray(np.concatenate(rows, axis=0))


def config_dict(world):
    """Identical in both arms (the driver compares them)."""
    return {
        "workload": "BASELINE configs[3] unit: %d ensemble members x %d perturbed Chapman+dipole profiles = %d profiles "
                    "x 174 freqs 0.1-17.4 MHz, X-mode, n_points=%d, 620 altitudes; ONE fixed batch per step"
                    % (MEMBERS, BASE_PROFILES, N_PROFILES, N_POINTS),
        "n_profiles": N_PROFILES, "n_freq": 174, "n_alt": 620, "n_points": N_POINTS, "mode": MODE,
        "sharding": "by profile, %s, over %d rank(s), one process per GPU; no data-path collective; result gathered "
                    "into rank 0's page-locked host buffer inside the timed region" % (LAYOUT, world),
        "l2": "flushed between timed steps (256 MiB write); inputs per rank (%.0f MB) also exceed the 126 MB L2"
              % (N_PROFILES / world * 620 * 3 * 8 / 1e6),
    }


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ----------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, budget_s=240.0, cores=None):
    """The reference's CPU path on a bounded sample of the workload: every step hands ONE profile (all 174
    frequencies, X-mode, n = 20 000) to each of ``cores`` processes; consecutive steps take consecutive profiles of
    member 0, so K steps cover K x cores distinct profiles.  Cost per profile does not depend on the profile
    (BASELINE.md section 2), so the rate extrapolates linearly to the 65 536-profile batch.
    Also: the ONE-process rate (how the reference actually runs)."""
    from oracle import cpu_baseline as cb
    synth = load_synth()
    params = workload_parameters(synth)
    alt, freq = synth.default_alt(), synth.default_freq()
    pool = cb.ReferencePool(cores)
    try:
        k = pool.cores

        def profiles(step):
            sel = (np.arange(k) + step * k) % N_PROFILES
            return sel, synth.profiles_from_parameters(*params[sel].T, alt=alt)

        sel, (den, bmag, bpsi) = profiles(0)
        t_one, vh_one = cb.one_process(freq, den[0], bmag[0], bpsi[0], alt, MODE, N_POINTS, repeats=2)
        wall, _ = pool.one_pass(freq, den, bmag, bpsi, alt, MODE, N_POINTS)          # sizing pass (untimed)
        steps_fit = max(1, int(budget_s / max(wall, 1e-3)) - warmup)
        steps_run = max(1, min(steps, steps_fit))
        for w in range(warmup):
            _, (d, b, p) = profiles(w)
            pool.one_pass(freq, d, b, p, alt, MODE, N_POINTS)
        total_wall, total_vh, checked = 0.0, 0, []
        for s in range(steps_run):
            sel, (d, b, p) = profiles(s)
            wall, vh = pool.one_pass(freq, d, b, p, alt, MODE, N_POINTS)
            total_wall += wall
            total_vh += vh.size
            if s == 0:
                checked = [sel, vh]
    finally:
        pool.close()
    return dict(value=total_vh / total_wall, cores=pool.cores, kind=pool.kind, ms_per_step=1e3 * total_wall / steps_run,
                steps_run=steps_run, one_process_value=freq.size / t_one, one_process_s_per_profile=t_one,
                sample="%d profiles of the batch per step (one per process, all 174 frequencies, X-mode, n=20000), "
                       "%d processes = host cores, %d steps = %d distinct profiles; rate extrapolates linearly to "
                       "the %d-profile batch" % (pool.cores, pool.cores, steps_run, steps_run * pool.cores, N_PROFILES),
                rows=checked[0], vh=checked[1])


def cpu_baseline_block(r):
    what = ("PyRayHF.library.vertical_forward_operator, unmodified copy in oracle/_ref (oracle/make_ref.py)"
            if r["kind"] == "reference" else
            "numpy restatement oracle/vfo_oracle.py (oracle/_ref absent): bit-identical to the reference in the dev "
            "container, ~25 % faster")
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
            "what": what, "one_process_value": r["one_process_value"],
            "one_process_s_per_profile": r["one_process_s_per_profile"]}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    warnings.simplefilter("ignore")
    r = cpu_reference(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": cpu_baseline_block(r),
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "steps_run": r["steps_run"],
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


class BackgroundClocks:
    """Samples the clocks from a thread WHILE a step runs (a 250 ms step gives the thread time to see load)."""

    def __init__(self, sampler, period=0.05):
        import threading
        self.s, self.period = sampler, period
        self.stop = threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.wait(self.period):
            self.s.sample()

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def committed_traffic(kernel):
    """DRAM bytes per launch of `kernel` (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture
    of this bench command, committed as profiles/traffic.json: {kernel: {"dram_bytes_per_launch": ...}}), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def row_classes(torch, freq_mhz, den, bmag, mode):
    """Counts, on the GPU with torch (bench accounting, not the product path), how the rows of a shard split into
    rows that ENTER the grid loop (reflect above the first level), rows clamped to the first level (reflect at or
    below it: constant mu', only the weights are summed) and dead rows; plus sum over rows of the truncated length.
    library.py:371-399 decides this: valid = running max of X (O) / X+Y (X) over the levels below the peak >= 1."""
    cp2, gp = 8.97866275 ** 2, 2.799249247e10
    f = freq_mhz * 1e6
    kx = (cp2 / (f * f))[None, :, None]
    ky = (gp / f)[None, :, None] if mode == "X" else None
    n_loop = n_clamp = n_dead = at_sum = 0
    n_alt = den.shape[1]
    lev = torch.arange(n_alt, device=den.device)[None, None, :]
    for p0 in range(0, den.shape[0], 1024):
        d = den[p0:p0 + 1024]
        nt = torch.argmax(d, dim=1)
        crit = d[:, None, :] * kx
        if ky is not None:
            crit = crit + bmag[p0:p0 + 1024][:, None, :] * ky
        crit = torch.where(lev < nt[:, None, None], crit, torch.full_like(crit, -1.0))
        live = crit.max(dim=2).values >= 1.0
        clamp = live & (crit[:, :, 0] >= 1.0)
        n_loop += int((live & ~clamp).sum().item())
        n_clamp += int(clamp.sum().item())
        n_dead += int((~live).sum().item())
        at_sum += int(nt.sum().item()) * freq_mhz.numel()
    return n_loop, n_clamp, n_dead, at_sum


def algorithmic_flops(n_loop, n_clamp, at_sum, n_points):
    """W: 77 N + 8 per row that enters the grid loop, 77 + 8 per row clamped to the first level
    (constant mu': one evaluation, the weights telescope), 4 per truncated level of EVERY row (critical curve)."""
    return n_loop * (FLOPS_PER_POINT * n_points + 8) + n_clamp * (FLOPS_PER_CLAMPED_ROW + 8) + 4 * at_sum


def timed(torch, stream, fn, reps=5, warm=2, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        ev.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def next_rows_section(torch, dev, stream, ctx, freq, den, bmag, bpsi, alt, with_cpu):
    """Throughput of the operators beside the headline path: the standalone regrid stage
    (HBM-bound), the elementwise stages, the batched Snell's-law tracers and the inversion objective."""
    import ctypes
    vp = ctypes.c_void_p
    L = ctx.lib
    sp = vp(stream.cuda_stream)
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev)   # noqa: E731
    out = {}
    hbm_peak = measured_peaks().get("hbm_gbs", 6650.0)
    # ---- regrid_to_nonuniform_grid: five [F x N] arrays written ----
    n_freq, n_pts = freq.size, N_POINTS
    tf, td, tb, tp, ta = t(freq * 1e6), t(den), t(bmag), t(bpsi), t(alt)
    hc = torch.empty(n_freq, dtype=torch.float64, device=dev)
    big = [torch.empty((n_freq, n_pts), dtype=torch.float64, device=dev) for _ in range(5)]
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    ms = timed(torch, stream, lambda: ctx.check(L.prhf_regrid_f64(
        ctx.handle, vp(tf.data_ptr()), n_freq, vp(td.data_ptr()), vp(tb.data_ptr()), vp(tp.data_ptr()), vp(ta.data_ptr()),
        alt.size, 1, n_pts, vp(hc.data_ptr()), *[vp(b.data_ptr()) for b in big], vp(st.data_ptr()), sp)))
    gbs = 5 * n_freq * n_pts * 8 / (ms * 1e-3) / 1e9
    out["regrid_stage"] = {"workload": "regrid_to_nonuniform_grid, %d freqs x %d points, 5 arrays written" % (n_freq, n_pts),
                           "ms": ms, "write_GBps": gbs, "frac_of_measured_hbm": gbs / hbm_peak,
                           "note": "row-setup kernel + write kernel; bound = HBM writes (MEASURED_PEAKS.json hbm_gbs)"}
    del big
    # ---- elementwise stages on 16 Mi elements: find_X (1 read + 1 write), find_mu_mup (3 reads + 2 writes) ----
    n_el = 1 << 24
    ex = torch.rand(n_el, dtype=torch.float64, device=dev) * 0.9
    ey = torch.rand(n_el, dtype=torch.float64, device=dev) * 0.5 + 0.05
    ep = torch.rand(n_el, dtype=torch.float64, device=dev) * 90.0
    eo1, eo2 = torch.empty_like(ex), torch.empty_like(ex)
    fscal = torch.full((1,), 5e6, dtype=torch.float64, device=dev)
    eden = ex * 1e11
    ms = timed(torch, stream, lambda: ctx.check(L.prhf_find_x_f64(ctx.handle, vp(eden.data_ptr()), 1, vp(fscal.data_ptr()),
                                                                  0, n_el, vp(eo1.data_ptr()), None, sp)))
    gbs = 2 * n_el * 8 / (ms * 1e-3) / 1e9
    out["find_X_stage"] = {"elements": n_el, "ms": ms, "GBps": gbs, "frac_of_measured_hbm": gbs / hbm_peak,
                           "note": "bit-identical to numpy: one IEEE sqrt and one IEEE division per element keep it "
                                   "FP64-bound (the fused operator uses reciprocal seeds instead)"}
    ms = timed(torch, stream, lambda: ctx.check(L.prhf_mu_mup_f64(ctx.handle, vp(ex.data_ptr()), vp(ey.data_ptr()),
                                                                  vp(ep.data_ptr()), n_el, 1, 0, 0, vp(eo1.data_ptr()),
                                                                  vp(eo2.data_ptr()), sp)))
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
        ms = timed(torch, stream, lambda: ctx.check(L.prhf_snell_f64(
            ctx.handle, vp(t_f.data_ptr()), vp(t_e.data_ptr()), f_r.size, vp(ta.data_ptr()), vp(td.data_ptr()),
            vp(tb.data_ptr()), vp(tp.data_ptr()), alt.size, 1, geo, 0, 1.0, 200.0, 400, 6371.0, vp(scal.data_ptr()), None,
            None, 0, vp(npth.data_ptr()), sp)))
        entry = {"rays": int(f_r.size), "rays_with_a_path": int((npth > 0).sum().item()), "ms": ms,
                 "rays_per_s": f_r.size / (ms * 1e-3)}
        # the same fan through the fan entry: refractive-index field once per frequency instead of once per ray
        scal_pairs = scal.clone()
        t_ff, t_fe = t(freq * 1e6), t(elev)
        ms_fan = timed(torch, stream, lambda: ctx.check(L.prhf_snell_fan_f64(
            ctx.handle, vp(t_ff.data_ptr()), freq.size, vp(t_fe.data_ptr()), elev.size, vp(ta.data_ptr()),
            vp(td.data_ptr()), vp(tb.data_ptr()), vp(tp.data_ptr()), alt.size, 1, geo, 0, 1.0, 200.0, 400, 6371.0,
            vp(scal.data_ptr()), None, None, 0, vp(npth.data_ptr()), sp)))
        entry["fan_entry_ms"] = ms_fan
        entry["fan_entry_rays_per_s"] = f_r.size / (ms_fan * 1e-3)
        entry["fan_entry_equals_per_ray_entry_bitwise"] = bool(torch.equal(
            torch.nan_to_num(scal, nan=-1.0), torch.nan_to_num(scal_pairs, nan=-1.0)))
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
    ms = timed(torch, stream, lambda: ctx.check(L.prhf_residual_f64(ctx.handle, vp(vm.data_ptr()), vp(vo.data_ptr()), 4096,
                                                                    n_freq, None, vp(chi.data_ptr()), sp)))
    out["residual_chi2"] = {"candidates": 4096, "ms": ms, "read_GBps": 4096 * n_freq * 8 / (ms * 1e-3) / 1e9}
    return out


def latency_block(torch, dev, stream, ctx, synth, flush, steps, warmup, clocks, peak_tf):
    """BASELINE.json configs[1]: ONE profile x 174 frequencies, X-mode, n = 20 000 -- device-timed single launch
    and the numpy drop-in call end to end, with the roofline of the single-launch kernel."""
    import ctypes
    import pyrayhf_b200
    vp = ctypes.c_void_p
    den, bmag, bpsi, alt = synth.bench_day_profile()
    freq = synth.default_freq()
    t_freq, t_den, t_b, t_psi, t_alt = (torch.from_numpy(np.ascontiguousarray(v)).to(dev)
                                         for v in (freq, den[None], bmag[None], bpsi[None], alt))
    out = torch.empty((1, freq.size), dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)

    def step_device():
        ctx.check(ctx.lib.prhf_vfo_f64(ctx.handle, vp(t_freq.data_ptr()), freq.size, 0, vp(t_den.data_ptr()),
                                       vp(t_b.data_ptr()), vp(t_psi.data_ptr()), vp(t_alt.data_ptr()), 0, 1, alt.size,
                                       1, N_POINTS, 0, vp(out.data_ptr()), vp(status.data_ptr()),
                                       vp(stream.cuda_stream)))

    for _ in range(warmup):
        flush.zero_()
        step_device()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.zero_()
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    launches = ctx.launch_count() - launches0
    dev_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    vh_dev = out.cpu().numpy()[0]
    for _ in range(warmup):
        pyrayhf_b200.vertical_forward_operator(freq, den, bmag, bpsi, alt, MODE, N_POINTS, device=dev.index)
    e2e_s, vh_e2e, e2e_each = 0.0, None, []
    for k in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vh_e2e = pyrayhf_b200.vertical_forward_operator(freq, den, bmag, bpsi, alt, MODE, N_POINTS, device=dev.index)
        e2e_each.append(time.perf_counter() - t0)
        e2e_s += e2e_each[-1]
        if k % 16 == 0:
            clocks.sample()
    ctx.kernel_timing(True)
    for _ in range(max(5, min(steps, 20))):
        flush.zero_()
        step_device()
    rows_ms, tile_ms, pairs = ctx.kernel_timing(False)
    torch.cuda.synchronize()
    n_loop, n_clamp, n_dead, at_sum = row_classes(torch, t_freq, t_den, t_b, MODE)
    flops = algorithmic_flops(n_loop, n_clamp, at_sum, N_POINTS)
    kernel_ms = tile_ms / max(pairs, 1)
    achieved = flops / (kernel_ms * 1e-3) / 1e12
    solo = launches == steps
    kname = "vfo_solo_kernel<1,0>" if solo else "vfo_tile_kernel<1,0>"
    # inputs once (3 profile arrays + altitudes + frequencies), outputs, and the stretched-grid tables (m, dm, E)
    hbm_bytes = (4 * alt.size + freq.size) * 8 + 8 * freq.size + 3 * (N_POINTS + 4) * 8
    return {
        "workload": "BASELINE configs[1]: single synthetic Chapman day profile (lat 4.5, lon 0, dipole B), X-mode, "
                    "174 freqs 0.1-17.4 MHz, n_points=20000, 620 altitudes",
        "device_us_per_call": 1e3 * dev_ms, "device_value": freq.size / (dev_ms * 1e-3),
        "e2e_us_per_call": 1e6 * e2e_s / steps, "e2e_us_per_call_median": 1e6 * float(np.median(e2e_each)),
        "e2e_value": freq.size * steps / e2e_s, "unit": UNIT, "calls": steps,
        "e2e_api": "pyrayhf_b200.vertical_forward_operator(numpy...) -> prhf_vfo_host_f64",
        "e2e_h2d_bytes_per_call": int((freq.size + 4 * alt.size) * 8), "e2e_d2h_bytes_per_call": int(freq.size * 8 + 4),
        "launches_per_call": launches / steps,
        "rows": {"enter_grid_loop": n_loop, "clamped_to_first_level": n_clamp, "no_reflection": n_dead,
                 "finite_results": int(np.isfinite(vh_dev).sum())},
        "roofline": {"bound": "fp64", "kernel": "prhf::" + kname, "achieved": achieved, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved / peak_tf, "kernel_ms": kernel_ms,
                     "algorithmic_flops_per_launch": flops, "traffic": committed_traffic(kname),
                     "hbm_algorithmic_bytes_per_launch": hbm_bytes},
    }, (freq, den, bmag, bpsi, alt, vh_e2e)


def config3_block(torch, dev, stream, synth, flush):
    """BASELINE configs[2]: the 65 341 profiles of the 1-degree global grid, O and X mode, device-resident."""
    import pyrayhf_b200
    lat, lon = synth.global_grid_points()
    alt, freq = synth.default_alt(), synth.default_freq()
    fof2, hmf2, scale_h, foe = synth.layer_parameters(lat, lon)
    den, bmag, bpsi = synth.profiles_from_parameters_device(fof2, hmf2, scale_h, foe, lat, alt=alt, device=dev)
    t_freq = torch.from_numpy(freq).to(dev)
    t_alt = torch.from_numpy(alt).to(dev)
    out = torch.empty((lat.size, freq.size), dtype=torch.float64, device=dev)
    block = {"workload": "BASELINE configs[2]: %d profiles (1-degree global lat/lon grid of Chapman layers + dipole B) "
                         "x 174 freqs, device-resident" % lat.size}
    for mode, n in (("X", 20000), ("O", 20000), ("X", 200), ("O", 200)):
        ms = timed(torch, stream, lambda: pyrayhf_b200.vertical_forward_operator_batched(
            t_freq, den, bmag, bpsi, t_alt, mode, n, out=out, errors='nan'), reps=3, warm=1, flush=flush)
        n_loop, n_clamp, n_dead, _ = row_classes(torch, t_freq, den, bmag, mode)
        block["%s_n%d" % (mode, n)] = {"ms": ms, "value": lat.size * freq.size / (ms * 1e-3), "unit": UNIT,
                                       "grid_points_per_s": n_loop * n / (ms * 1e-3),
                                       "rows_entering_grid_loop": n_loop}
    return block


def inversion_block(torch, dev, synth):
    """one brute-force fit as minimize_parameters runs it (library.py:672-825) -- a 31 x 31 grid of
    (hmF2, B_bot) candidates x the observed frequencies, O-mode, n_points = 200 (the reference's default) -- as one
    device pipeline: profiles built on the GPU, forward operator, residual, argmin; 16 bytes come back per fit."""
    import pyrayhf_b200 as prhf
    alt, freq = synth.default_alt(), synth.default_freq()
    _, bmag, bpsi = synth.profiles_at([20.0], [0.0], alt)
    bmag, bpsi = bmag[0], bpsi[0]
    builder = prhf.chapman_profile_builder(3.0)
    nm_true = (9.6e6 / 8.97866275) ** 2
    truth = builder(nm_true, np.array([301.3]), np.array([47.4]), alt).cpu().numpy()[0]
    out = {"workload": "31 x 31 (hmF2, B_bot) brute grid = 961 candidate profiles per fit, O-mode, Chapman profile "
                       "builder on the device, observations = the frequencies of 0.1-17.4 MHz that reflect"}
    for n_points in (200, 2000):
        vh_true = prhf.vertical_forward_operator(freq, truth, bmag, bpsi, alt, 'O', n_points)
        f_in, obs = prhf.inversion.sort_observations(freq, vh_true)
        nm = prhf.nmf2_from_max_frequency(f_in[-1], alt, bmag, 310.0, 'O')
        hm_grid, bb_grid = np.linspace(279.0, 341.0, 31), np.linspace(40.0, 55.0, 31)
        fit = lambda: prhf.brute_force_search(f_in, obs, alt, bmag, bpsi, nm, hm_grid, bb_grid, builder, 'O', n_points)  # noqa: E731
        for _ in range(3):
            res = fit()
        torch.cuda.synchronize()
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            res = fit()
        dt = (time.perf_counter() - t0) / reps
        out["n%d" % n_points] = {"ms_per_fit": 1e3 * dt, "fits_per_s": 1.0 / dt, "candidates_per_s": 961 / dt,
                                 "vh_per_s": 961 * f_in.size / dt, "frequencies": int(f_in.size),
                                 "d2h_bytes_per_fit": 16,
                                 "h2d_bytes_per_fit": int(961 * 5 * 8 + (4 * alt.size + 2 * f_in.size) * 8),
                                 "best": [res[0], res[1]]}
    return out


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    import pyrayhf_b200
    from pyrayhf_b200 import _cabi, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run "
                         "(--nproc-per-node %d)" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    synth = load_synth()
    ctx = _cabi.context(local)
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    clocks = ClockSampler(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- the batch: every rank builds ITS shard on the device from 40 bytes of parameters per profile ----
    alt, freq = synth.default_alt(), synth.default_freq()
    idx = sharding.shard_indices(N_PROFILES, world, rank, LAYOUT)
    params = workload_parameters(synth)[idx]
    with torch.cuda.stream(stream):
        den, bmag, bpsi = synth.profiles_from_parameters_device(*params.T, alt=alt, device=dev)
        t_freq = torch.from_numpy(freq).to(dev)
        t_alt = torch.from_numpy(alt).to(dev)
    op = sharding.ShardedForwardOperator(N_PROFILES, freq.size, layout=LAYOUT, gather_to=0)
    n_local = idx.size

    def step_resident():
        return op(t_freq, den, bmag, bpsi, t_alt, MODE, N_POINTS, errors='nan', stream=stream.cuda_stream)

    # ---- device-resident throughput: CUDA events around every step, L2 flushed between steps ----
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            flush.zero_()
            step_resident()
        barrier()
        launches0 = ctx.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        wall0 = time.perf_counter()
        with BackgroundClocks(clocks):
            for k in range(args.steps):
                flush.zero_()
                if world > 1:
                    dist.barrier()                 # every rank starts the step together
                ev[k][0].record(stream)
                result = step_resident()           # returns after this rank's rows are in the host buffer + barrier
                ev[k][1].record(stream)
        barrier()
        wall_s = time.perf_counter() - wall0
    launches = ctx.launch_count() - launches0
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    vh_resident = np.array(result[:256], copy=True) if rank == 0 else None

    # ---- end to end: the same call with HOST arrays (page-locked), H2D + kernels + D2H inside the step ----
    h_den, h_b, h_psi = (pyrayhf_b200.pinned_empty((n_local, alt.size)) for _ in range(3))
    for h, d in ((h_den, den), (h_b, bmag), (h_psi, bpsi)):
        torch.from_numpy(h).copy_(d)
    torch.cuda.synchronize()

    def step_host():
        return op(freq, h_den, h_b, h_psi, alt, MODE, N_POINTS, errors='nan')

    for _ in range(args.warmup):
        step_host()
    barrier()
    e2e_s = 0.0
    for k in range(args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        result = step_host()
        e2e_s += time.perf_counter() - t0
    barrier()
    vh_e2e = np.array(result[:256], copy=True) if rank == 0 else None

    # ---- the dominant kernel alone (roofline): measurement mode of the C ABI puts CUDA events around the row-setup
    #      kernel and the tile kernel of every chunk (this serialises them; L2 flushed between passes) ----
    ctx.kernel_timing(True)
    n_kt = 3
    with torch.cuda.stream(stream):
        for _ in range(n_kt):
            flush.zero_()
            step_resident()
    rows_ms, tile_ms, pairs = ctx.kernel_timing(False)
    barrier()
    with torch.cuda.stream(stream):
        n_loop, n_clamp, n_dead, at_sum = row_classes(torch, t_freq, den, bmag, MODE)
    flops_shard = algorithmic_flops(n_loop, n_clamp, at_sum, N_POINTS)

    t = torch.tensor([total_ms, e2e_s, wall_s], dtype=torch.float64, device=dev)
    cnt = torch.tensor([launches, n_loop, n_clamp, n_dead], dtype=torch.int64, device=dev)
    per_rank_ms = [total_ms / args.steps]
    if world > 1:
        gathered = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, torch.tensor([total_ms / args.steps], dtype=torch.float64, device=dev))
        per_rank_ms = [float(g.item()) for g in gathered]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    total_ms, e2e_s, wall_s = (float(v) for v in t.tolist())
    launches_all, loop_all, clamp_all, dead_all = (int(v) for v in cnt.tolist())
    units = N_PROFILES * freq.size                 # virtual heights per step, whole job
    value = units * args.steps / (total_ms * 1e-3)
    e2e_value = units * args.steps / e2e_s

    if rank != 0:
        op.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peak_tf = ctx.measure_fp64_peak()
    launches_per_step = pairs / n_kt                                  # chunks of the streaming entry on this rank
    kernel_ms = tile_ms / max(pairs, 1)                               # one tile-kernel launch, CUDA events on its stream
    flops_launch = flops_shard / max(launches_per_step, 1)
    achieved_tf = flops_launch / (kernel_ms * 1e-3) / 1e12
    prof_per_launch = n_local / max(launches_per_step, 1)
    # per launch: the chunk's three profile arrays + its results + status, the shared altitude / frequency vectors and
    # the stretched-grid tables (m, dm, E: the grid loop itself reads only its seeds from E)
    hbm_bytes = (prof_per_launch * (3 * alt.size * 8 + freq.size * 8 + 4) + (alt.size + freq.size) * 8 +
                 3 * (N_POINTS + 4) * 8)
    kname = "vfo_queue_kernel<1,0>"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(world),
        "workload_stats": {
            "rows_per_step": units, "rows_entering_grid_loop": loop_all, "rows_clamped_to_first_level": clamp_all,
            "rows_without_reflection": dead_all, "grid_points_per_s": loop_all * N_POINTS / (total_ms / args.steps * 1e-3),
            "ms_per_step_by_rank": per_rank_ms, "wall_s_timed_loop_incl_flush_and_barriers": wall_s,
            "chunks_per_rank_per_step": launches_per_step,
            "timing": "CUDA events on the launching stream around every step (the stream waits for the last copy-out; "
                      "the call returns after this rank's rows are in the host buffer and the ranks' barrier), summed "
                      "over steps, max over ranks",
        },
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf, "traffic": committed_traffic(kname),
                     "kernel": "prhf::" + kname + " (rows that reflect queued by the row-setup kernel; 128-thread CTAs, "
                               "one per resident slot, eight per SM, drawing whole-row tiles by ticket; rank 0's shard)",
                     "kernel_ms": kernel_ms, "rows_kernel_ms": rows_ms / max(pairs, 1),
                     "launches_per_step_per_rank": launches_per_step,
                     "step_frac": (flops_shard / (total_ms / args.steps * 1e-3) / 1e12) / peak_tf,
                     "algorithmic_flops_per_launch": flops_launch,
                     "note": "achieved = W (77 flop per grid point of a row that enters the grid loop, 77 per row "
                             "clamped to the first level, 4 per truncated level of every row; SURVEY 8d) of one "
                             "launch / its duration from CUDA events on its stream (prhf_kernel_timing); step_frac "
                             "uses rank 0's whole timed step (row setup, copies and barrier included)",
                     "peak_source": "prhf_measure_fp64_peak: dependent-free DFMA kernel timed live on this GPU "
                                    "(MEASURED_PEAKS.json has no FP64 entry)",
                     "hbm_algorithmic_bytes_per_launch": hbm_bytes,
                     "hbm_frac_of_measured": (hbm_bytes / (kernel_ms * 1e-3) / 1e9) /
                                             measured_peaks().get("hbm_gbs", 6650.0)},
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(N_PROFILES * 3 * alt.size * 8 + world * (alt.size + freq.size) * 8),
                "d2h_bytes_per_step": int(units * 8 + N_PROFILES * 4),
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "pyrayhf_b200.sharding.ShardedForwardOperator(...)(freq, den, bmag, bpsi, alt, 'X', 20000) with "
                       "page-locked numpy shards -> prhf_vfo_stream_f64 -> rank 0's page-locked [P x F] buffer"},
        "gpu_launches": int(launches_all),
    }

    # ---- parity of what was just timed (first rows of the gathered result: with interleaved shards they come from
    #      every rank) against the CPU reference, X-mode rule: NaN masks equal, rel. error <= 1e-9 ----
    warnings.simplefilter("ignore")
    par = {"tolerance": 1e-9, "resident_equals_e2e_bitwise": bool(np.array_equal(vh_resident, vh_e2e, equal_nan=True))}
    ref_rows = None
    if not args.no_cpu_baseline:
        if world == 1:
            cb = cpu_reference(steps=2, warmup=0, budget_s=30.0, cores=min(os.cpu_count() or 1, 256))
            line["cpu_baseline"] = cpu_baseline_block(cb)
            try:
                from oracle.cpu_baseline import c_port_rate
                d1, b1, p1, _ = synth.bench_day_profile()
                rate, cores = c_port_rate(freq, d1, b1, p1, alt, MODE, N_POINTS)
                line["cpu_baseline"]["c_port_value"] = rate
                line["cpu_baseline"]["c_port_note"] = "scalar C restatement, %d pthreads (extra, not the baseline)" % cores
            except Exception as exc:       # the C oracle is optional here
                line["cpu_baseline"]["c_port_note"] = "unavailable: %s" % exc
            sel, ref_rows = cb["rows"], cb["vh"]
            par["checked_against"] = cb["kind"]
        else:
            from oracle import cpu_baseline as cbm
            sel = np.arange(2)
            d, b, p = synth.profiles_from_parameters(*workload_parameters(synth)[sel].T, alt=alt)
            ref_rows = np.stack([cbm._work((freq, d[q], b[q], p[q], alt, MODE, N_POINTS))[1] for q in range(2)])
            par["checked_against"] = cbm.baseline_kind()
        got = vh_e2e[sel]
        m = np.isfinite(ref_rows) & np.isfinite(got)
        par["x_mode"] = {"profiles_checked": int(len(sel)), "rows_checked": int(ref_rows.size),
                         "nan_mask_mismatches": int(np.sum(np.isnan(got) != np.isnan(ref_rows))),
                         "max_rel_err_vs_reference": float(np.max(np.abs(got[m] - ref_rows[m]) / np.abs(ref_rows[m])))}
        # O-mode (not the headline mode): the float64 reference is itself 1e-6..1e-4 from an exact evaluation of its own
        # formulas (cancellation in library.py:229), so the rule is 1e-9 against the long-double truth and inside the
        # reference's own rounding ball
        try:
            from oracle import scalar, vfo_oracle
            ko = 2
            d, b, p = synth.profiles_from_parameters(*workload_parameters(synth)[:ko].T, alt=alt)
            got_o = pyrayhf_b200.vertical_forward_operator_batched(freq, d, b, p, alt, "O", N_POINTS, errors='nan')
            mult = vfo_oracle.stretch_multiplier(N_POINTS)
            truth = scalar.vertical_forward_operator_batched(freq, d, b, p, alt, "O", N_POINTS, variant=1, multiplier=mult)[0]
            lit = scalar.vertical_forward_operator_batched(freq, d, b, p, alt, "O", N_POINTS, variant=0, multiplier=mult)[0]
            m = np.isfinite(truth)
            par["o_mode"] = {"profiles_checked": ko, "rows_checked": int(truth.size),
                             "nan_mask_mismatches": int(np.sum(np.isnan(got_o) != np.isnan(lit))),
                             "max_rel_err_vs_long_double_truth": float(np.max(np.abs(got_o[m] - truth[m]) / np.abs(truth[m]))),
                             "float64_restatement_vs_truth": float(np.max(np.abs(lit[m] - truth[m]) / np.abs(truth[m])))}
        except Exception as exc:
            par["o_mode"] = {"unavailable": str(exc)}
    line["parity"] = par

    # ---- single-profile latency (configs[1]) and, on one GPU, the rows beside the path ----
    with torch.cuda.stream(stream):
        # (200 calls of ~50 us: a mean over 20 calls moves by several microseconds with one scheduling hiccup)
        lat_block, single = latency_block(torch, dev, stream, ctx, synth, flush, max(args.steps, 200),
                                          max(args.warmup, 10), clocks, peak_tf)
        line["latency"] = lat_block
        if world == 1 and not args.no_extras:
            line["config3"] = config3_block(torch, dev, stream, synth, flush)
            sf, sd, sb, sp_, sa, _ = single
            line["next_rows"] = next_rows_section(torch, dev, stream, ctx, sf, sd, sb, sp_, sa,
                                                  with_cpu=not args.no_cpu_baseline)
            line["inversion"] = inversion_block(torch, dev, synth)
    line["clocks"] = clocks.summary()
    emit(line)
    op.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """Native libraries (NCCL's version banner at NCCL_DEBUG=INFO, for one) print to file descriptor 1.  The contract
    is ONE JSON line on stdout, so everything written to fd 1 while the benchmark runs goes to stderr (where the
    driver's NCCL log ends up as well); `emit` writes the line to the real stdout."""
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config3 / next_rows blocks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
