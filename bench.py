#!/usr/bin/env python
"""bench.py -- headline benchmark: FP64 CG iterations/s on the 512^3 7-point Poisson system (BASELINE config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid 512]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is ONE CG ITERATION of Poisson3DSolver's solver over the whole grid (matrix-free operator apply + p.Ap,
x/r update + r.r, p update).  The reference's ConjugateGrad has no tolerance test, so solve(K) performs exactly K
iterations (unless |p.Ap| < 1e-12 fires, which RHS-R does not do before ~2k iterations at this size) and
iterations/s is a like-for-like unit.  The 512^3 total is fixed for every N ("strong" scaling, z-slab partition).

  value   device-resident: b, x, r, p already in HBM; CUDA events on the launching stream around the K-iteration
          device loop (inside the library), barrier + synchronize on both sides, max over ranks.
  e2e     the same solve through the reference-facing C-ABI call with HOST buffers (atk_cg_solve_host): pinned b is
          uploaded, K iterations run, x is downloaded; wall clock around the call, max over ranks.
  roofline  per-kernel CUDA-event times from a second run of the same loop launched kernel by kernel.
  extra   side measurements of the same hot path: at N = 1 SpMV GB/s (matrix-free stencil, SELL-32, CSR-stream) and
          GMRES(300) on bcsstk08; at N > 1 the distributed SpMV alone (matrix-free and assembled SELL-32), aggregate.

--impl reference times the reference's own CPU implementation (oracle/_ref, the unmodified headers; or the C port
when that library is absent) on a bounded sample of the same workload and prints the same JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "poisson512_cg_iterations_per_sec"
UNIT = "iterations/s"
SEED = 12345


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# =====================================================================================================================
def workload_name(g):
    """config.workload, identical in both arms: the operator's storage (matrix-free here, assembled CSC in the reference)
    is an implementation property and is reported under config.operator instead."""
    return f"poisson3d_cg_{g}^3"


def reference_sample(grid, sample_grid, steps, warmup):
    """The reference's own CPU implementation of the path (oracle/_ref = the unmodified headers; the C port when that
    library is absent) on a bounded sample of the workload: the same RHS-R Poisson system on a sample_grid^3 sub-grid,
    `warmup` untimed + `steps` timed CG iterations through the reference's public API (SparseMatrixCSC assembled by
    Poisson3DSolver::buildLaplacianMatrix, ConjugateGrad(A, b, maxIter, tol).solve(x): the ctor's copy of A is part of the
    timed call, as it is for a user).  The reference is serial: 1 thread whatever the host has.
    Returns (seconds for `steps` iterations, seconds of assembly, kind)."""
    from oracle import pyoracle

    s = sample_grid
    if pyoracle.ref_available():
        R = pyoracle.Ref()
        t0 = time.perf_counter()
        M = R.mat_poisson(s, s, s)
        build = time.perf_counter() - t0
        b = R.poisson_rhs(s, s, s, rhs_kind=1, seed=SEED)
        if warmup > 0:
            R.cg_solve(M, b, warmup)
        t0 = time.perf_counter()
        R.cg_solve(M, b, steps)
        return time.perf_counter() - t0, build, "reference"
    P = pyoracle.Port()
    b = P.poisson_rhs(s, s, s, rhs_kind=1, seed=SEED)
    t0 = time.perf_counter()
    A = P.poisson_csc(s, s, s)
    build = time.perf_counter() - t0
    if warmup > 0:
        P.cg(A, b, warmup)
    t0 = time.perf_counter()
    P.cg(A, b, steps)
    return time.perf_counter() - t0, build, "port"


def pick_sample_grid(requested, grid):
    """256^3 (1/8 of the rows of 512^3; ~8 GB in the reference, ~1.1 s per iteration) when the host has the memory,
    else 128^3.  Larger samples extrapolate more honestly: the reference's per-row cost grows with size (cache misses)."""
    if requested:
        return min(requested, grid)
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    return min(grid, 256 if avail >= 16 * 2 ** 30 else 128)


def run_reference(args):
    """--impl reference: rank 0 alone times the reference's CPU path; a step = one CG iteration on the sample sub-grid.
    `ms_per_step` and `steps` are what was actually timed; `value` is that rate scaled by rows to the full workload and
    says so (`extrapolated`, `sample_grid`, `row_scale`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g = args.grid
    s = pick_sample_grid(args.ref_sample_grid, g)
    scale = (s ** 3) / float(g ** 3)
    secs, build, kind = reference_sample(g, s, args.steps, args.warmup)
    it_s_sample = args.steps / secs
    value = it_s_sample * scale
    sample = (f"{s}^3 sub-grid ({s ** 3} of {g ** 3} rows), {args.warmup} untimed + {args.steps} timed CG iterations of the "
              f"assembled-CSC reference solver in {secs:.2f} s (+ {build:.2f} s assembly, untimed); iterations/s scaled by rows "
              f"x{scale:.6f}; 1 thread (the reference is serial; {os.cpu_count()} host cores present)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "extrapolated": s != g, "sample_grid": s, "row_scale": scale, "sample_iterations_per_sec": it_s_sample,
        "config": {"workload": workload_name(g), "grid": [g, g, g], "rows": g ** 3, "rhs": f"RHS-R seed {SEED}",
                   "operator": "assembled SparseMatrixCSC (the reference's storage)", "step": "one CG iteration",
                   "timed": f"steps x ms_per_step is the measured time on the {s}^3 sample; value = sample rate x row_scale"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count(), "sample_iterations_per_sec": it_s_sample,
                         "extrapolated": s != g, "sample_grid": s},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =====================================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist

    from algotoolkit_b200 import capi, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: bench.py has no CPU path for --impl ours")
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        blob = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            blob = torch.tensor(list(capi.Context.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(blob, 0)
        uid = bytes(blob.cpu().tolist())
    ctx = capi.Context(local_rank, rank, world, uid)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g = args.grid
    cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
    op = ctx.operator_stencil7(g, g, g, cx, cy, cz)
    n_loc, N = op.n_rows_local, g ** 3
    plane = g * g
    k0 = op.row_begin // plane
    k1 = k0 + n_loc // plane

    # ---- synthetic RHS-R of this rank's slab, in pinned host memory
    import ctypes as C

    L = ctx.L
    hb, hx = C.c_void_p(), C.c_void_p()
    capi._check(L.atk_host_alloc(n_loc * 8, C.byref(hb)))
    capi._check(L.atk_host_alloc(n_loc * 8, C.byref(hx)))
    b_host = np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_double)), shape=(n_loc,))
    x_host = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_double)), shape=(n_loc,))
    synthetic.poisson_rhs_r(g, g, g, k0, k1, seed=SEED, out=b_host)
    bn, dn = C.c_int(-1), C.c_int(-1)
    L.atk_host_numa_node(hb, C.byref(bn), C.byref(dn))
    numa_info = {"pinned_buffer_node": bn.value, "gpu_node": dn.value}

    db = ctx.vector(b_host)
    dx, dr, dp = ctx.empty(n_loc), ctx.empty(n_loc), ctx.empty(n_loc)

    def fresh_state():
        dx.zero()
        dr.copy_from(db)
        dp.copy_from(db)

    # ---- warm-up (W iterations, untimed)
    fresh_state()
    ctx.cg_solve(op, db, dx, dr, dp, max(args.warmup, 3))

    sampler = ClockSampler(local_rank)
    # ---- device-resident timed region: exactly K iterations
    fresh_state()
    barrier()
    launches0 = ctx.launch_count()
    with sampler:
        res = ctx.cg_solve(op, db, dx, dr, dp, args.steps)
        barrier()
    launches = ctx.launch_count() - launches0
    dev_ms = max_over_ranks(res["device_ms"])
    if res["iters"] != args.steps:
        raise SystemExit(f"solver stopped after {res['iters']} of {args.steps} iterations (|pAp|<1e-12 exit)")
    value = args.steps / (dev_ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call: H2D b, K iterations, D2H x
    e2e_times = []
    for rep in range(2):  # the first call also pays one-time workspace allocation; the second is the steady state reported
        info = capi.CgInfo()
        barrier()
        t0 = time.perf_counter()
        capi._check(L.atk_cg_solve_host(ctx.h, op.h, hb, hx, None, None, 1, args.steps, C.byref(info)))
        ctx.synchronize()
        e2e_times.append(max_over_ranks(time.perf_counter() - t0))
    t_e2e = e2e_times[-1]
    e2e_value = args.steps / t_e2e
    x_norm2 = float(np.dot(x_host, x_host))  # the host reads the result it received

    # ---- per-kernel times (kernel-by-kernel launch of the same loop) for the roofline
    fresh_state()
    barrier()
    kt_steps = min(args.steps, 50)
    kres = ctx.cg_solve(op, db, dx, dr, dp, kt_steps, time_kernels=True)
    kms = [max_over_ranks(v) / kt_steps for v in kres["kernel_ms"]]  # ms per launch: K1, K2, K3
    peak, peak_src = measured_peaks()
    if kres["fused"] and kres["persistent"]:
        # ONE persistent launch per solve: the same two passes as below, separated by in-kernel grid /
        # cross-GPU reductions; per-phase times (reduction included) from %globaltimer stamps inside the kernel
        bytes_per_row = [48.0, 24.0]
        names = ["cg_persistent_kernel phase A (= KA)", "cg_persistent_kernel phase B (= KB)"]
    elif kres["fused"]:
        # fused two-kernel formulation: KA = p update + deferred x update + stencil + p.Ap reads r,p,x and
        # writes p,Ap,x (48 B/row); KB = r update + r.r reads r,Ap, writes r (24 B/row).  72 B/row per iteration.
        bytes_per_row = [48.0, 24.0]
        names = ["stencil7_cg_fused_smem_kernel (KA)", "cg_fused_update_r_kernel (KB)"]
    else:
        bytes_per_row = [16.0, 48.0, 24.0]  # stencil apply+dot, x/r update + r.r, p update
        names = ["stencil7_kernel (K1)", "cg_update_xr_kernel (K2)", "cg_update_p_kernel (K3)"]
    kms = kms[:len(names)]
    iter_bytes_per_row = sum(bytes_per_row)
    per_kernel = []
    for nm, bpr, ms in zip(names, bytes_per_row, kms):
        gbs = bpr * n_loc / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        per_kernel.append({"kernel": nm, "ms_per_launch": ms, "algorithmic_bytes": bpr * n_loc, "achieved_gbs": gbs,
                           "frac": gbs / peak, "share_of_step": ms / sum(kms) if sum(kms) > 0 else 0.0})
    if res["persistent"]:
        # the timed region IS one launch of one kernel: its algorithmic bytes are K iterations x 72 B/row plus the start-up
        # sums (r.r, b.b: 16 B/row) and the final pending x / p updates (read x, p, r; write x, p: 40 B/row), its duration
        # the CUDA-event time of the timed region itself
        launch_bytes = (72.0 * args.steps + 56.0) * n_loc
        gbs = launch_bytes / (dev_ms * 1e-3) / 1e9
        dom = {"kernel": "cg_persistent_kernel<32,4> (whole solve: start-up sums + K iterations + final updates, one launch)",
               "ms_per_launch": dev_ms, "algorithmic_bytes": launch_bytes, "achieved_gbs": gbs, "frac": gbs / peak}
        traffic_key = "cg_persistent_kernel"
    else:
        dom = max(per_kernel, key=lambda d: d["ms_per_launch"])
        traffic_key = dom["kernel"]
    traffic = None
    try:  # DRAM bytes of the dominant kernel from the committed ncu --set full capture of the shipped kernel (N = 1, same
        # grid); per launch, scaled to this run's step count where the capture is of a whole-solve launch
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ent = json.load(f).get(traffic_key)
        if ent and ent["grid"] == g:
            if "dram_bytes_per_row_per_step" in ent:
                traffic = (ent["dram_bytes_per_row_per_step"] * args.steps + ent.get("dram_bytes_per_row_fixed", 0.0)) * n_loc
            elif ent["n_gpus"] == world:
                traffic = ent["dram_bytes_per_launch"]
    except Exception:
        pass
    t_it = dev_ms * 1e-3 / args.steps
    iter_gbs = iter_bytes_per_row * N / t_it / 1e9 / world      # bytes this formulation must move, per GPU

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(g), "grid": [g, g, g], "rows": N, "rows_per_gpu": n_loc,
                   "rhs": f"RHS-R seed {SEED}", "operator": "matrix-free 7-point stencil, Dirichlet shell = identity rows",
                   "parallelism": f"z-slab x{world}", "l2": "inputs larger than L2 (5 vectors x %.2f GB per GPU)" % (n_loc * 8 / 1e9),
                   "step": "one CG iteration"},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_loc * 8 / args.steps,
                "d2h_bytes_per_step": n_loc * 8 / args.steps, "seconds": t_e2e, "first_call_seconds": e2e_times[0],
                "device_loop_seconds": info.device_ms * 1e-3, "x_norm": float(np.sqrt(x_norm2)),
                "phases_ms": {"h2d": max_over_ranks(info.h2d_ms), "solve": max_over_ranks(info.solve_ms),
                              "d2h": max_over_ranks(info.d2h_ms),
                              "other_host": t_e2e * 1e3 - max_over_ranks(info.h2d_ms + info.solve_ms + info.d2h_ms)},
                "numa": numa_info,
                "note": "one atk_cg_solve_host call = K steps; pinned b up, x down, per rank"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                     "frac": dom["frac"], "traffic": traffic, "algorithmic_bytes": dom["algorithmic_bytes"], "peak_source": peak_src,
                     "formulation": ("persistent one-launch loop, phases A+B, 72 B/row" if res["persistent"] else
                                     "fused KA+KB, 72 B/row" if kres["fused"] else "K1+K2+K3, 88 B/row"),
                     "iteration_bytes_per_row": iter_bytes_per_row,
                     "iteration_achieved_gbs_per_gpu": iter_gbs, "iteration_frac": iter_gbs / peak,
                     "iteration_frac_of_8TBs": iter_gbs / 8000.0,
                     "per_kernel": per_kernel},
        "rs_final": res["rs"],
    }

    def allreduce_sum(vals):
        if world == 1:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return [float(v) for v in t.cpu()]

    # ---- full-size property (the reference cannot run 512^3: ~55 GB, hours): after the K timed iterations the
    # recurrence residual sqrt(r.r) must equal the true residual ||b - A x|| recomputed from x
    fresh_state()
    res_p = ctx.cg_solve(op, db, dx, dr, dp, args.steps)
    op.apply(dx, dp)  # dp <- A x  (p is scratch from here on)
    ctx.synchronize()
    d_true = b_host - dp.numpy()
    true_res = float(np.sqrt(allreduce_sum([float(d_true @ d_true)])[0]))
    b_norm = float(np.sqrt(allreduce_sum([float(b_host @ b_host)])[0]))
    line["parity"] = {"full_size": {"grid": g, "checked_by": "properties only (the reference needs ~55 GB and hours at 512^3)",
                                    "recurrence_residual": float(np.sqrt(res_p["rs"])), "true_residual": true_res,
                                    "rel_diff_vs_b_norm": abs(float(np.sqrt(res_p["rs"])) - true_res) / b_norm,
                                    "ok": bool(abs(float(np.sqrt(res_p["rs"])) - true_res) <= 1e-10 * b_norm)}}
    if not args.no_parity:
        try:
            line["parity"].update(parity_checks(ctx, capi, synthetic, world, allreduce_sum))
        except Exception as e:
            line["parity"].update({"ok": False, "error": f"{type(e).__name__}: {e}"})
        line["parity"]["ok"] = bool(line["parity"].get("ok", False) and line["parity"]["full_size"]["ok"])
    else:
        line["parity"]["ok"] = line["parity"]["full_size"]["ok"]

    # ---- secondary measurements of the same hot path (N=1 only; a few seconds): SpMV GB/s and GMRES restarts/s
    if world == 1 and not args.no_extra:
        try:
            line["extra"] = secondary_measurements(ctx, capi, synthetic, op, g, peak)
        except Exception as e:  # the headline above must survive a failure of the side measurements
            line["extra_error"] = f"{type(e).__name__}: {e}"
    elif not args.no_extra:
        # N > 1: the distributed matrix-free SpMV alone (z-slabs, halo planes exchanged over NCCL on the communication
        # stream while the interior planes run); every rank takes part, time = max over ranks, bytes = whole grid
        xs = ctx.vector(np.random.default_rng(rank).standard_normal(op.n_rows_local))
        ys = ctx.empty(op.n_rows_local)
        for _ in range(3):
            op.apply(xs, ys)
        ctx.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            op.apply(xs, ys)
        ctx.synchronize()
        tt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        xs.free()
        ys.free()
        tsec = float(tt.item())
        line["extra"] = {"spmv_stencil7_matrix_free": {
            "grid": g, "n_gpus": world, "ms": tsec * 1e3, "algorithmic_bytes": 16.0 * N, "GBs": 16.0 * N / tsec / 1e9,
            "GBs_per_gpu": 16.0 * N / tsec / 1e9 / world, "frac_of_measured_peak_per_gpu": 16.0 * N / tsec / 1e9 / world / peak,
            "note": "aggregate over ranks; includes the halo exchange of every apply"}}
        # the assembled 256^3 Poisson matrix in SELL-32, partitioned by row blocks (z-slabs), neighbour ranges over NCCL
        ga = 256
        ca = synthetic.poisson_coeffs(ga, ga, ga)
        opa = ctx.operator_poisson_assembled(ga, ga, ga, ca[0], ca[1], ca[2], capi.FORMAT_SELL)
        nnz_loc = torch.tensor([float(opa.nnz)], dtype=torch.float64, device="cuda")
        dist.all_reduce(nnz_loc)
        xs = ctx.vector(np.random.default_rng(rank).standard_normal(opa.n_rows_local))
        ys = ctx.empty(opa.n_rows_local)
        for _ in range(3):
            opa.apply(xs, ys)
        ctx.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            opa.apply(xs, ys)
        ctx.synchronize()
        tt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ta = float(tt.item())
        alg = 12.0 * float(nnz_loc.item()) + 4.0 * (ga ** 3 + world) + 16.0 * ga ** 3
        line["extra"]["spmv_sell32_assembled"] = {
            "grid": ga, "n_gpus": world, "nnz": int(nnz_loc.item()), "ms": ta * 1e3, "algorithmic_bytes": alg, "GBs": alg / ta / 1e9,
            "GBs_per_gpu": alg / ta / 1e9 / world, "frac_of_measured_peak_per_gpu": alg / ta / 1e9 / world / peak,
            "note": "row blocks = z-slabs; aggregate over ranks; includes the neighbour exchange of every apply"}
        xs.free()
        ys.free()
        opa.destroy()

    # ---- BASELINE config 5 (GMRES(300), 27-point, 16.7 M rows) at this N, both orthogonalisation modes
    if not args.no_extra and not args.no_config5:
        try:
            line.setdefault("extra", {})["gmres300_op27_256"] = config5_gmres(ctx, capi, synthetic, world, allreduce_sum,
                                                                                max_over_ranks, peak)
        except Exception as e:
            line.setdefault("extra", {})["gmres300_op27_256_error"] = f"{type(e).__name__}: {e}"

    # ---- CPU baseline beside it (rank 0, N=1 only): the unmodified reference on a bounded sample
    if world == 1 and not args.no_cpu_baseline:
        s = pick_sample_grid(args.ref_sample_grid, g)
        it = args.ref_sample_iters if args.ref_sample_iters > 0 else (10 if s >= 256 else 150)
        secs, build, kind = reference_sample(g, s, it, 0)
        scale = s ** 3 / float(N)
        line["cpu_baseline"] = {
            "value": it / secs * scale, "unit": UNIT, "cores": 1, "kind": kind, "extrapolated": s != g, "sample_grid": s,
            "row_scale": scale, "sample_iterations_per_sec": it / secs,
            "sample": f"{s}^3 sub-grid, {it} CG iterations in {secs:.2f} s ({it / secs:.2f} it/s; + {build:.2f} s assembly, untimed), "
                      f"scaled by rows x{scale:.6f} to the {g}^3 workload; the reference is single-threaded "
                      f"({os.cpu_count()} host cores present)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not line["parity"].get("ok", False):
        raise SystemExit("parity check failed: " + json.dumps(line["parity"]))


def parity_checks(ctx, capi, synthetic, world, allreduce_sum):
    """Results parity of the SHARDED solvers against fixtures produced by the unmodified reference
    (tests/golden/golden_large.json, tests/golden/make_golden_large.py), outside every timed region:
      * Poisson CG 128^3 RHS-R to the |pAp| < 1e-12 exit (ConjugateGradient.hpp:49-83): iteration count within +-2 of the
        reference's 576, ||x|| and 512 sampled entries within 1e-10 relative;
      * GMRES(60) x 2 restarts on the 27-point operator at 32^3 (GMRES.hpp:62-112), per-projection and batched MGS:
        residual ladder, ||x|| and 512 sampled entries within 1e-10 relative.
    Every rank solves its z-slab; norms and sample errors are summed over ranks."""
    REL = 1e-10
    with open(os.path.join(ROOT, "tests", "golden", "golden_large.json")) as f:
        gl = json.load(f)
    out = {"tolerance": REL, "fixtures": "tests/golden/golden_large.json (unmodified reference, same inputs)"}
    ok = True

    def slab_errors(x, lo, fx):
        idx, want = np.array(fx["sample_index"]), np.array(fx["sample_x"])
        m = (idx >= lo) & (idx < lo + x.size)
        d = x[idx[m] - lo] - want[m]
        xx, dd = allreduce_sum([float(x @ x), float(d @ d)])
        return abs(np.sqrt(xx) - fx["x_norm"]) / fx["x_norm"], np.sqrt(dd) / np.linalg.norm(want)

    # ---- CG, BASELINE config 3/4 operator at 128^3
    fx = gl["128"]
    n = fx["n"]
    cx, cy, cz = synthetic.poisson_coeffs(n, n, n)
    op = ctx.operator_stencil7(n, n, n, cx, cy, cz)
    plane = n * n
    k0 = op.row_begin // plane
    b_loc = synthetic.poisson_rhs_r(n, n, n, k0, k0 + op.n_rows_local // plane, seed=SEED)
    res = ctx.cg_solve_host(op, b_loc, 100000, want_rp=False)
    e_norm, e_samp = slab_errors(res["x"], op.row_begin, fx)
    c_ok = bool(res["stopped_on_pAp"] and abs(res["iters"] - fx["iterations"]) <= 2 and e_norm <= REL and e_samp <= REL)
    out["cg_poisson128_to_exit"] = {"iterations": res["iters"], "reference_iterations": fx["iterations"],
                                    "stopped_on_pAp": bool(res["stopped_on_pAp"]), "x_norm_rel_err": e_norm,
                                    "sample512_rel_err": e_samp, "peer_memory_mode": bool(res["peer"]),
                                    "fused": bool(res["fused"]), "ok": c_ok}
    ok = ok and c_ok
    op.destroy()

    # ---- GMRES, BASELINE config 5 operator at 32^3
    fx = gl["gmres_op27_32_K60"]
    n = fx["n"]
    op = ctx.operator_stencil27(n, n, n, synthetic.stencil27_coefficients())
    x_ext = synthetic.u01_array(777, op.row_begin, op.n_rows_local)
    b_loc = op.apply_numpy(x_ext)
    for name, ortho in (("gmres60_op27_32_mgs", capi.GMRES_MGS), ("gmres60_op27_32_mgs_batched", capi.GMRES_MGS_BATCHED)):
        r = ctx.gmres_solve_host(op, b_loc, np.zeros(op.n_rows_local), fx["max_iter"], fx["K"], 1e-6, ortho=ortho)
        e_norm, e_samp = slab_errors(r["x"], op.row_begin, fx)
        lad = (len(r["resid"]) == len(fx["resid"]) and
               bool(np.allclose(r["resid"], fx["resid"], rtol=REL, atol=0.0)))
        g_ok = bool(lad and r["status"] == fx["status"] and e_norm <= REL and e_samp <= REL)
        out[name] = {"residuals": [float(v) for v in r["resid"]], "reference_residuals": fx["resid"],
                     "x_norm_rel_err": e_norm, "sample512_rel_err": e_samp, "ok": g_ok}
        ok = ok and g_ok
    op.destroy()
    out["ok"] = bool(ok)
    return out


def gmres_projection_traffic(name, launches, K):
    """HBM bytes per row and projection the path that ran actually moves, recognised by its launch count for one restart of
    K steps (K(K-1)/2 projections): the persistent step kernels (2 launches per step) read each Krylov column once from HBM
    (8 B; w stays on chip), the fused per-projection launch (1 per projection) moves 32 B, the dot + axmy pair (2 per
    projection) the 40 B SURVEY 8(d) counts; batched MGS reads every column twice (16 B)."""
    if name == "batched_mgs":
        return 16.0, "two passes over the columns"
    pairs = K * (K - 1) / 2.0
    if launches < 0.5 * pairs:
        return 8.0, "persistent step kernel (w on chip, one column read per projection)"
    if launches < 1.5 * pairs:
        return 32.0, "fused update+products launch per projection"
    return 40.0, "dot and axmy launch per projection"


def config5_gmres(ctx, capi, synthetic, world, allreduce_sum, max_over_ranks, peak, grid=256, K=300):
    """BASELINE config 5: GMRES(300), one restart cycle, on the synthetic 27-point operator at 256^3 (16.7 M rows), z-slab
    sharded over all ranks.  Both orthogonalisation modes; the residual the solver reports is checked against the true
    residual ||b - A x|| recomputed from the returned x.  Algorithmic bytes of the projections:
    exact MGS 40*N*K(K-1)/2, batched 16*N*K(K-1)/2 (+ 32*N per step)."""
    N = grid ** 3
    op = ctx.operator_stencil27(grid, grid, grid, synthetic.stencil27_coefficients())
    n_loc = op.n_rows_local
    x_ext = synthetic.u01_array(777, op.row_begin, n_loc)
    dx = ctx.vector(x_ext)
    db = ctx.empty(n_loc)
    op.apply(dx, db)
    ctx.synchronize()
    b_loc = db.numpy()
    out = {"grid": grid, "rows": N, "K": K, "restarts": 1, "n_gpus": world}
    for name, ortho, bpr in (("batched_mgs", capi.GMRES_MGS_BATCHED, 16.0), ("exact_mgs", capi.GMRES_MGS, 40.0)):
        r = ctx.gmres_solve_host(op, b_loc, np.zeros(n_loc), 1, K, 1e-6, ortho=ortho)
        secs = max_over_ranks(r["device_ms"]) * 1e-3
        dx.upload(r["x"])
        op.apply(dx, db)
        ctx.synchronize()
        d = b_loc - db.numpy()
        true_res = float(np.sqrt(allreduce_sum([float(d @ d)])[0]))
        proj_bytes = bpr * N * K * (K - 1) / 2.0
        moved_bpr, path = gmres_projection_traffic(name, r["kernel_launches"], K)
        moved = moved_bpr * N * K * (K - 1) / 2.0
        out[name] = {"seconds_per_restart": secs, "arnoldi_steps_per_s": r["arnoldi_steps"] / secs,
                     "path": path, "moved_bytes_per_row_per_projection": moved_bpr,
                     "moved_GBs_per_gpu": moved / secs / 1e9 / world,
                     "moved_frac_of_measured_peak_per_gpu": moved / secs / 1e9 / world / peak,
                     "residual_after_restart": float(r["resid"][-1]), "true_residual": true_res,
                     "residual_rel_diff": abs(float(r["resid"][-1]) - true_res) / true_res,
                     "projection_GBs_per_gpu": proj_bytes / secs / 1e9 / world,
                     "projection_frac_of_measured_peak_per_gpu": proj_bytes / secs / 1e9 / world / peak,
                     "kernel_launches": r["kernel_launches"]}
    out["note"] = ("projection_* use SURVEY 8(d)'s algorithmic bytes of the unfused formulation (exact MGS 40 B, batched 16 B per row and "
                   "projection) and exceed the HBM peak where a fused path moves fewer; moved_* use the bytes the path that ran moves")
    a, b = out["batched_mgs"]["residual_after_restart"], out["exact_mgs"]["residual_after_restart"]
    out["modes_rel_diff"] = abs(a - b) / abs(b)
    dx.free()
    db.free()
    op.destroy()
    return out



def secondary_measurements(ctx, capi, synthetic, stencil_op, g, peak):
    """SpMV HBM GB/s (matrix-free stencil at the bench grid, SELL-32 / CSR-vector on the assembled 256^3 Poisson matrix)
    and GMRES(300) on bcsstk08 (BASELINE config 1) in both reduction modes.  CUDA-event timed inside the library where
    the call reports it, host-timed around synchronised launches otherwise."""
    out = {}

    def time_apply(op, reps=20):
        x = ctx.vector(np.random.default_rng(0).standard_normal(op.n_cols))
        y = ctx.empty(op.n_rows_local)
        for _ in range(3):
            op.apply(x, y)
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            op.apply(x, y)
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / reps
        x.free()
        y.free()
        return dt

    n = g ** 3
    t = time_apply(stencil_op)
    out["spmv_stencil7_matrix_free"] = {"grid": g, "ms": t * 1e3, "algorithmic_bytes": 16.0 * n, "GBs": 16.0 * n / t / 1e9,
                                        "frac_of_measured_peak": 16.0 * n / t / 1e9 / peak}
    ga = 256
    cx, cy, cz = synthetic.poisson_coeffs(ga, ga, ga)
    for name, fmt in (("spmv_sell32_assembled", capi.FORMAT_SELL), ("spmv_csr_stream_assembled", capi.FORMAT_CSR_VECTOR)):
        t0 = time.perf_counter()
        op = ctx.operator_poisson_assembled(ga, ga, ga, cx, cy, cz, fmt)
        ctx.synchronize()
        t_build = time.perf_counter() - t0
        t = time_apply(op)
        alg = 12.0 * op.nnz + 4.0 * (ga ** 3 + 1) + 16.0 * ga ** 3
        out[name] = {"grid": ga, "nnz": op.nnz, "ms": t * 1e3, "algorithmic_bytes": alg, "GBs": alg / t / 1e9,
                     "frac_of_measured_peak": alg / t / 1e9 / peak, "device_assembly_and_conversion_s": t_build}
        op.destroy()
    # GMRES(300) on bcsstk08 as the reference loads it (lower triangle; the file header is ignored)
    try:
        from algotoolkit_b200 import mmio

        A = mmio.read_matrix_market_csc(os.path.join(ROOT, "tests", "golden", "data", "bcsstk08.mtx"))
        op = ctx.operator_from_csc(A["n_rows"], A["n_cols"], A["col_ptr"], A["row_idx"], A["vals"])
        x_ext = synthetic.u01_array(777, 0, A["n_cols"])
        b = op.apply_numpy(x_ext)
        for mode, name in ((capi.REDUCE_TREE, "tree"), (capi.REDUCE_SEQUENTIAL, "sequential")):
            ctx.set_reduction(mode)
            res = ctx.gmres_solve_host(op, b, np.zeros(A["n_rows"]), 2, 300, 1e-6)
            out[f"gmres300_bcsstk08_{name}"] = {"restarts": res["restarts"], "arnoldi_steps": res["arnoldi_steps"],
                                                "restarts_per_s": res["restarts"] / (res["device_ms"] * 1e-3),
                                                "arnoldi_steps_per_s": res["arnoldi_steps"] / (res["device_ms"] * 1e-3),
                                                "residual_after_restart_1": float(res["resid"][1])}
        ctx.set_reduction(capi.REDUCE_TREE)
        op.destroy()
    except Exception as e:  # the headline number must not depend on the secondary block
        out["gmres300_bcsstk08_error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--ref-sample-grid", type=int, default=0, help="0: 256 when the host has >= 16 GB free, else 128")
    ap.add_argument("--ref-sample-iters", type=int, default=0, help="cpu_baseline sample length; 0: 10 at 256^3, 150 at 128^3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
