#!/usr/bin/env python
"""bench.py — measurement contract of the uspmv-b200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

One "step" = one SpMV (one pass of the hot path over the resident matrix).  At N = 1 the workload is
BASELINE.json configs[1]: 3-D 7-point Laplacian 256^3 (16.8 M rows, 117 M nnz), SELL-C-sigma C = 32, dp.
At N > 1 (torchrun, one rank per GPU) every rank owns one 256 x 256 x 256 z-slab of a 256 x 256 x (256 N) grid
(weak scaling), with the halo exchange of remote x elements inside every step.

Prints ONE JSON line on rank 0 (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SpMV GFLOP/s (2*nnz/t) on the SELL-C-sigma path; achieved HBM GB/s and roofline fraction in `roofline`"
UNIT = "GFLOP/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="laplace7_256", help="laplace7_<n> | stencil27_<n> (n^3 grid per GPU) | powerlaw_<log2 rows> (config 4, needs --ap)")
    ap.add_argument("--ap", default=None, choices=["ap[dp_sp]", "ap[dp_hp]", "ap[sp_hp]", "ap[dp_sp_hp]"],
                    help="adaptive precision on the power-law matrix: thresholds 1.0 / 1e-2, rows split over the ranks by seg-nnz")
    ap.add_argument("--C", type=int, default=32)
    ap.add_argument("--sigma", type=int, default=1)
    ap.add_argument("--vt", default="dp", choices=["dp", "sp", "hp"])
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="CPU baseline time box")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="e2e: do not restrict the CPU affinity to the GPU's NUMA node while allocating pinned buffers")
    ap.add_argument("--strong", action="store_true", help="N>1: ONE n^3 grid cut into N z-slabs (config 5) instead of n^3 per GPU")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: exchange first, then one full SpMV (the reference's order)")
    ap.add_argument("--p2p-mode", type=int, default=2, choices=[0, 1, 2],
                    help="N>1, --halo p2p: 2 = ONE fused kernel per step, 1 = push / wait kernels next to the interior kernel, 0 = exchange first")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="N>1: NVLink peer stores + epoch flags, or NCCL send/recv")
    ap.add_argument("--bvs", type=int, default=1, help="block_vec_size > 1: SpMMV (config 3), block-vector halo exchange when N>1")
    ap.add_argument("--layout", default="rowwise", choices=["rowwise", "colwise"], help="block vector layout for --bvs > 1")
    ap.add_argument("--solve", action="store_true", help="solve mode: a step is { halo exchange ; SpMV ; swap } on two device buffers")
    ap.add_argument("--no-other-configs", dest="other_configs", action="store_false",
                    help="skip the other BASELINE.json configs (N = 1: SpMMV bvs 4/8 dp+sp, the AP power-law matrix; N > 1: 27-point strong-scaling slab)")
    ap.add_argument("--no-config4", action="store_true", help="other_configs: skip the power-law AP matrix")
    ap.add_argument("--no-banded", action="store_true", help="other_configs: config 4 with the fused kernel only (no column-banded plan)")
    ap.add_argument("--config4-log2-rows", type=int, default=25)
    ap.add_argument("--config5-n", type=int, default=512, help="other_configs at N > 1: grid edge of the 27-point strong-scaling matrix")
    ap.add_argument("--set", action="append", default=[], metavar="NAME=VALUE", help="library option for A/B runs (uspmv_set_option), repeatable")
    ap.add_argument("--steady-steps", type=int, default=1000, help="also report ms per step over this many steps (0 = skip)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference's own CUDA kernel (oracle/_ref, recompiled for sm_100)")
    return ap.parse_args()


def workload_dims(name):
    kind, n = name.rsplit("_", 1)
    n = int(n)
    pts = {"laplace7": 7, "stencil27": 27, "powerlaw": 0}[kind]
    return pts, n


def build_powerlaw_ap(pkg, ctx, log2_rows, mode, C, sigma, rank, world, plan="fused", alg_bytes=None, n_bands=0):
    """BASELINE config 4: power-law matrix with 2^log2_rows rows, generated on the device, rows split by
    seg-nnz, per-rank partition_precisions with t1 = 1.0, t2 = 1e-2.  Every rank generates only its own rows (twice at N > 1: once on
    an equal-rows split to count the elements per row for the seg-nnz walk, once on the final split).
    plan = "banded": the column-banded execution plan (one GPU only)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    eng, d = pkg.engine, pkg.dist
    n = 1 << log2_rows
    if world > 1:
        eq = np.arange(world + 1, dtype=np.int64) * (n // world)
        eq[world] = n
        mtx = eng.MtxData.powerlaw(n, int(eq[rank]), int(eq[rank + 1]), ctx=ctx)
        Iv = pkg.validate.device_int_tensor(mtx.device_arrays()["I"].value, mtx.nnz, torch.device("cuda", ctx.device))
        cnt = torch.bincount(Iv, minlength=int(eq[rank + 1] - eq[rank])).to(torch.int32)
        sizes = [int(eq[r + 1] - eq[r]) for r in range(world)]
        bufs = [torch.empty(sz, dtype=torch.int32, device="cuda") for sz in sizes]
        dist.all_gather(bufs, cnt)
        wsa = d.seg_nnz_from_row_counts(torch.cat(bufs).cpu().numpy(), world)
        if (int(wsa[rank]), int(wsa[rank + 1])) != (int(eq[rank]), int(eq[rank + 1])):
            del mtx, Iv
            mtx = eng.MtxData.powerlaw(n, int(wsa[rank]), int(wsa[rank + 1]), ctx=ctx)
    else:
        wsa = np.array([0, n], np.int32)
        mtx = eng.MtxData.powerlaw(n, 0, n, ctx=ctx)
    if plan == "banded":
        if world != 1:
            raise ValueError("the column-banded plan is a one-GPU plan")
        return d.BandedApSpmv(ctx, mtx, mode, 1.0, 1e-2, C, sigma, n_bands=n_bands, algorithmic_bytes=alg_bytes), wsa
    return d.DistributedApSpmv(ctx, wsa, mtx, mode, 1.0, 1e-2, C, sigma, rank, world, keep_coo=True), wsa


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def algorithmic_bytes(n_elements, n_chunks, n_cols_plus_halo, n_rows, vsize, bvs=1):
    """SURVEY.md §8d / reference memory model main.cpp:655-663."""
    return n_elements * (vsize + 4) + n_chunks * 8 + bvs * vsize * n_cols_plus_halo + bvs * vsize * n_rows


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# the reference's own CPU path (oracle/_ref), used by --impl reference and by the cpu_baseline leg
# ------------------------------------------------------------------------------------------------------
def cpu_reference_run(pts, n, C, sigma, vt, steps, warmup, time_box=None):
    """Builds the matrix with the reference's convert_to_scs and times its OpenMP kernel (spmv_omp_scs_adv,
    kernels.hpp:265-301) on all host cores.  Returns (seconds_per_spmv, steps_done, nnz, threads, build_seconds)."""
    import numpy as np
    ncpu = len(os.sched_getaffinity(0)) or os.cpu_count() or 1
    if os.environ.get("TORCHELASTIC_RUN_ID") or os.environ.get("OMP_NUM_THREADS") in (None, "1"):
        os.environ["OMP_NUM_THREADS"] = str(ncpu)  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    from oracle import bindings
    if not bindings.ref_available():
        raise RuntimeError("oracle/_ref is not built (run __graft_entry__.build() where /root/reference exists)")
    orc = bindings.Oracle()
    ref = bindings.Ref("col")
    try:  # libgomp may have been initialised (with torchrun's OMP_NUM_THREADS=1) before the lines above ran
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
    except (OSError, KeyError, ValueError):
        pass
    t0 = time.time()
    n_rows, n_cols, I, J, V = orc.stencil_coo(pts, n, n, n)
    s = ref.convert_to_scs(n_rows, n_cols, I, J, V, C, sigma, vt, permute_cols=True)
    build_s = time.time() - t0
    nnz = len(I)
    del I, J, V
    npt = bindings.NPT[bindings.VT[vt]]
    x = np.full(max(s.n_rows_padded, n_cols), 5.0, npt)  # DefaultValues x = 5.0, classes_structs.hpp:1792-1810
    y = np.zeros(s.n_rows_padded, npt)
    vtc = bindings.VT[vt]
    adv = C in (2, 4, 8, 16, 32, 64, 128)

    def one():
        ref.spmv_scs_raw(vtc, adv, C, s.n_chunks, s.chunk_ptrs, s.chunk_lengths, s.col_idxs, s.values, x, y)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        one()
        done += 1
        if time_box is not None and time.perf_counter() - t0 > time_box:
            break
    dt = (time.perf_counter() - t0) / done
    _REF_MATRIX.clear()
    _REF_MATRIX.update({"key": (pts, n, C, sigma, vt), "scs": s, "x": x, "y": y.copy()})
    return dt, done, nnz, ref.omp_threads(), build_s


_REF_MATRIX = {}  # the reference-built matrix of the CPU leg, reused by the gpu_baseline leg


def gpu_reference_kernel_run(pts, n, C, sigma, vt, steps):
    """gpu_baseline: the reference's OWN CUDA kernel spmv_gpu_scs_adv (code/kernels.hpp:685-775), unmodified, recompiled for sm_100
    (oracle/_ref/libuspmv_ref_gpu.so, test infrastructure), on the matrix the reference's convert_to_scs built for the CPU leg, launched
    like the harness does (THREADS_PER_BLOCK = 128, one thread per padded row).  Its y must equal the reference's CPU y."""
    import numpy as np
    try:
        from oracle import bindings
        if not bindings.ref_gpu_available():
            return {"value": None, "unit": UNIT, "kind": "reference CUDA kernel", "sample": "unavailable: oracle/_ref/libuspmv_ref_gpu.so not built"}
        if vt not in ("dp", "sp"):
            return {"value": None, "unit": UNIT, "kind": "reference CUDA kernel", "sample": "unavailable: the reference's GPU half path is not built"}
        if _REF_MATRIX.get("key") != (pts, n, C, sigma, vt):
            return {"value": None, "unit": UNIT, "kind": "reference CUDA kernel", "sample": "unavailable: no reference-built matrix (CPU leg skipped)"}
        s, x, y_cpu = _REF_MATRIX["scs"], _REF_MATRIX["x"], _REF_MATRIX["y"]
        g = bindings.RefGpu()
        crs = C == 1 and sigma == 1
        kern = "csr" if crs else ("scs_adv" if C in (2, 4, 8, 16, 32, 64, 128) else "scs")
        y, ms = g.spmv(kern, vt, C, s.n_chunks, s.chunk_ptrs, s.chunk_lengths, s.col_idxs, s.values, x, warmup=5, steps=min(steps, 200))
        nnz = int(s.nnz)
        same = bool(np.array_equal(y, y_cpu))
        close = bool(np.allclose(y, y_cpu, rtol={"dp": 1e-12, "sp": 1e-5}[vt], atol=0))
        cusp = None
        if bindings.cusparse_available():  # the reference's own second GPU baseline: its USE_CUSPARSE mode (utilities.hpp:3380-3550)
            try:
                cs = bindings.CuSparse()
                kind = "csr" if crs else "sell"
                yc, msc = cs.spmv(kind, vt, s.n_rows, len(x), nnz, C, s.chunk_ptrs, s.col_idxs, s.values, x, warmup=5, steps=min(steps, 100))
                cusp = {"value": 2.0 * nnz / (msc / 1e3) / 1e9, "unit": UNIT, "ms_per_step": msc,
                        "kernel": ("cusparseSpMV, CSR" if crs else f"cusparseSpMV on cusparseCreateSlicedEll(slice size {C}) made from the same SELL-C-sigma arrays") +
                                  ", CUSPARSE_SPMV_ALG_DEFAULT (utilities.hpp:3393-3457)",
                        "y_close_to_reference_cpu_y": bool(np.allclose(yc, y_cpu[: s.n_rows], rtol={"dp": 1e-12, "sp": 1e-5}[vt], atol=1e-300))}
            except Exception as e:
                cusp = {"value": None, "sample": f"unavailable: {e}"}
        return {"value": 2.0 * nnz / (ms / 1e3) / 1e9, "unit": UNIT, "ms_per_step": ms, "kind": "reference CUDA kernel", "cusparse": cusp,
                "kernel": {"csr": "spmv_gpu_csr (kernels.hpp:631-659)", "scs_adv": "spmv_gpu_scs_adv / scs_impl_gpu<C> (kernels.hpp:685-775)",
                           "scs": "spmv_gpu_scs (kernels.hpp:579-608)"}[kern],
                "build": f"unmodified reference source, nvcc -O3 -gencode arch=compute_100,code=sm_100, THREADS_PER_BLOCK={g.threads_per_block}",
                "y_equals_reference_cpu_y": same, "y_close_to_reference_cpu_y": close,
                "sample": f"full matrix built by the reference's convert_to_scs ({nnz} nnz), device-resident, {min(steps, 200)} launches, CUDA events"}
    except Exception as e:  # a baseline leg must never take the line down
        return {"value": None, "unit": UNIT, "kind": "reference CUDA kernel", "sample": f"unavailable: {e}"}


def workload_label(args, pts, n, world):
    """config.workload — the SAME string on both arms (the driver compares them).  The reference arm is a single-process CPU run
    (no MPI on the box), so at N > 1 it times one rank's share of this workload."""
    if args.workload.startswith("powerlaw"):
        return f"{args.workload}: power-law matrix with 2^{n} rows (config 4), {args.ap} t1=1.0 t2=1e-2, scs C={args.C} sigma={args.sigma}"
    grid = (f"{pts}-point stencil on ONE {n}^3 grid cut into {world} z-slabs" if (args.strong and world > 1) else
            f"{pts}-point stencil on a {n}^3 grid per GPU")
    return (f"{args.workload}: {grid}, scs C={args.C} sigma={args.sigma} {args.vt} " +
            (f"SpMMV block_vec_size={args.bvs} {args.layout}" if args.bvs > 1 else "SpMV") +
            (", solve mode: each step = halo exchange + SpMV + swap on two device buffers" if args.solve else ""))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pts, n = workload_dims(args.workload)
    steps, warmup = min(args.steps, 200), min(args.warmup, 20)
    dt, done, nnz, threads, build_s = cpu_reference_run(pts, n, args.C, args.sigma, args.vt, steps, warmup, time_box=120.0)
    gf = 2.0 * nnz / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gf, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": {"dp": "f64", "sp": "f32", "hp": "f16"}[args.vt],
        "data": "synthetic", "config": {"workload": workload_label(args, pts, n, args.gpus),
                                        "arm": "one rank's slab on the host CPU (single process, no MPI on the box)",
                                        "kernel": "reference spmv_omp_scs_adv (kernels.hpp:265-301) via oracle/_ref, built by the reference's convert_to_scs",
                                        "x": "constant 5.0 (reference default)"},
        "cpu_baseline": {"value": gf, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"full {args.workload} matrix ({nnz} nnz), {done} SpMVs after {warmup} warm-ups; build {build_s:.1f} s untimed"},
        "e2e": {"value": gf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def timed_steps(runner, steps, warmup, world, rank, local_rank, capi, sample_clocks=True):
    """W warm-up steps, then EXACTLY `steps` steps between barrier + synchronize, CUDA events on the launching stream, max over
    ranks; clocks sampled with nvidia-smi during the region.  Returns (ms per step, launches, clocks)."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(warmup, 3)):
        runner.step()
    barrier()
    sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if world > 1:
        # The ranks leave the host barrier 100+ us apart; a rank whose neighbour starts late waits for that neighbour's first push, so
        # the skew would be charged to the K timed steps (20 x 0.25 ms).  ONE more untimed step after the barrier aligns the device
        # timelines (every rank's kernel ends only after its neighbours' pushes have arrived); the timed region is still exactly K steps.
        runner.step()
    l0 = capi.kernel_launches()
    e0.record()
    for _ in range(steps):
        runner.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = capi.kernel_launches() - l0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps, launches, clocks


def reduce_max(v, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v) if v == v and v != float("inf") else 1e300], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(vals, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]


def checked(runner, vt, world):
    """One validated run of the runner's step (x = f(global row), NaN-poisoned halo, y against an independent reference product
    on every rank — the solve-mode validation of the reference, main.cpp:528-631,968-990).  Returns (ok, max_rel_err, tol)."""
    tol = {"dp": 1e-12, "sp": 1e-5, "hp": 1e-2}[vt]
    err = reduce_max(runner.validate(), world)
    return bool(err <= tol), err, tol


def traffic_of(key):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if key and os.path.exists(tp):
        with open(tp) as f:
            return json.load(f).get(key)
    return None


def other_config(pkg, ctx, args, rank, world, local_rank, peak, name, make_runner, vt, bvs=1, steps=None, traffic_key=None):
    """One more BASELINE.json config measured with the headline's protocol: build, validate, time, roofline."""
    import torch
    capi = pkg.capi
    steps = steps or max(20, min(args.steps, 200))
    t0 = time.time()
    runner = make_runner()
    build_s = time.time() - t0
    ok, err, tol = checked(runner, vt, world)
    if hasattr(runner, "_coo"):
        runner._coo = None  # the COO triplets were only kept for the checked step
    ms, launches, clocks = timed_steps(runner, steps, max(args.warmup, 3), world, rank, local_rank, capi)
    if runner.p2p is not None:
        runner.p2p.sync()  # raises if any step's flag wait timed out
    kern_ms = runner.time_kernel(steps)
    vsize = {"dp": 8, "sp": 4, "hp": 2}[vt]
    if hasattr(runner, "algorithmic_bytes"):
        bytes_local = runner.algorithmic_bytes()
    else:
        bytes_local = algorithmic_bytes(runner.n_elements, runner.n_chunks, runner.n_cols_local + runner.n_halo, runner.n_rows_padded, vsize, bvs)
    nnz_total, bytes_total = reduce_sum([runner.nnz * bvs, bytes_local], world)
    achieved = bytes_local / (kern_ms / 1e3) / 1e9
    out = {"config": name, "value": 2.0 * nnz_total / (ms / 1e3) / 1e9, "unit": UNIT, "ms_per_step": ms, "steps": steps, "n_gpus": world,
           "dtype": {"dp": "f64", "sp": "f32", "hp": "f16"}[vt], "validated": ok, "max_rel_err": err, "tolerance": tol,
           "nnz": int(nnz_total), "gbs": bytes_total / (ms / 1e3) / 1e9,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_of(traffic_key),
                        "traffic_source": "a committed ncu capture of this kernel on this workload (profiles/traffic.json), NOT measured in this run"
                                          if traffic_of(traffic_key) else None,
                        "kernel": runner.kernel_name, "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": int(bytes_local)},
           "gpu_launches": int(launches), "clocks": clocks, "build_s": round(build_s, 2)}
    extra = getattr(runner, "describe", None)
    if extra:
        out.update(extra())
    runner.close()
    del runner
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    is_ap = args.workload.startswith("powerlaw")
    if is_ap and not args.ap:
        raise SystemExit("--workload powerlaw_<log2 rows> needs --ap")
    block_or_solve = args.bvs > 1 or args.solve or is_ap
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:  # the distributed runners also serve N = 1 (an arena with no peers): SpMMV, solve, AP and the other_configs block
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{29400 + os.getpid() % 500}", rank=0, world_size=1,
                                device_id=torch.device("cuda", local_rank))

    pkg = importlib.import_module("ultimate-spmv_b200")
    eng, capi = pkg.engine, pkg.capi
    pts, n = workload_dims(args.workload)
    vt = args.vt
    vsize = {"dp": 8, "sp": 4, "hp": 2}[vt]
    ctx = eng.default_context(local_rank)
    for kv in args.set:
        k, _, v = kv.partition("=")
        capi.set_option(k, int(v))

    wsa = None
    if is_ap:
        runner, wsa = build_powerlaw_ap(pkg, ctx, n, args.ap, args.C, args.sigma, rank, world)
        vt = "sp" if args.ap == "ap[sp_hp]" else "dp"
    elif world == 1 and not block_or_solve:
        runner = pkg.engine.SingleGpuSpmv(ctx, pts, n, args.C, args.sigma, vt)
    else:
        runner = pkg.dist.DistributedSpmv(ctx, pts, n, args.C, args.sigma, vt, rank, world, halo=args.halo,
                                          overlap=(False if args.no_overlap else (True if args.p2p_mode == 2 else args.p2p_mode)),
                                          strong=args.strong, bvs=args.bvs, layout=args.layout,
                                          n_buf=2 if (args.solve or (args.bvs == 1 and args.halo == "p2p")) else 1)
    p2p = getattr(runner, "p2p", None)

    # ---- checked step BEFORE timing, on every rank -----------------------------------------------------------------
    ok, max_err, tol = checked(runner, vt, world)
    if hasattr(runner, "_coo"):
        runner._coo = None
    if not ok and rank == 0:
        print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "validated": False, "max_rel_err": max_err,
                          "tolerance": tol, "error": "the checked step disagrees with the reference product; nothing was timed"}), flush=True)
    if not ok:
        if dist.is_initialized():
            dist.destroy_process_group()
        sys.exit(3)

    if args.solve:
        state = {"buf": 0}

        def solve_step():
            runner.p2p.spmv_buf(runner.scs, state["buf"], state["buf"] ^ 1, torch.cuda.current_stream(), runner.comm_stream)
            state["buf"] ^= 1
        runner.step = solve_step
        for b in runner.p2p.bufs:
            b.fill_(0.0)  # x stays finite over thousands of revisions of the stencil operator
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    nnz_local = runner.nnz * args.bvs
    if is_ap:
        bytes_local = runner.algorithmic_bytes()
    else:
        bytes_local = algorithmic_bytes(runner.n_elements, runner.n_chunks, runner.n_cols_local + runner.n_halo, runner.n_rows_padded, vsize, args.bvs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ms_step, launches, clocks = timed_steps(runner, args.steps, args.warmup, world, rank, local_rank, capi)
    exchange_errors = 0
    if p2p is not None:
        try:
            p2p.sync()  # a bounded flag wait that timed out in ANY step raises here
        except capi.UspmvError as e:
            exchange_errors = 1
            print(f"rank {rank}: {e}", file=sys.stderr, flush=True)
    exchange_errors = int(reduce_sum([exchange_errors], world)[0])
    nnz_total, bytes_total = reduce_sum([nnz_local, bytes_local], world)
    sec_per_step = ms_step / 1e3
    gflops = 2.0 * nnz_total / sec_per_step / 1e9
    # the same protocol over a longer run (informative: how much of `ms_per_step` is start-up of a K-step region)
    steady_ms = None
    if args.steady_steps > 0:
        steady_ms, _, _ = timed_steps(runner, args.steady_steps, 3, world, rank, local_rank, capi, sample_clocks=False)

    # kernel-only duration of the dominant kernel (SpMV) measured with CUDA events on its stream
    kern_ms = runner.time_kernel(max(args.steps, 20))
    peak, peak_src = measured_peak()
    achieved = bytes_local / (kern_ms / 1e3) / 1e9
    traffic, traffic_src = None, "not captured for this workload / rank count"
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1:
        with open(tp) as f:
            tj = json.load(f)
        key = f"{args.workload}|C{args.C}|s{args.sigma}|{vt}" + (f"|bvs{args.bvs}{args.layout}" if args.bvs > 1 else "")
        if key in tj:
            traffic, traffic_src = tj[key], tj.get("_source", "ncu capture") + " — a committed capture of this kernel on this workload, NOT measured in this run"

    # end-to-end through the host-buffer C-ABI call (pinned host x / y, copies inside the timed region)
    e2e = None
    if not args.no_e2e and not args.solve and not is_ap:
        e2e_steps = max(3, min(args.steps, 20))
        # pinned staging buffers next to the GPU's PCIe root: first-touch under an affinity restricted to the GPU's NUMA node
        bound = None if args.no_numa_bind else pkg.dist.bind_to_gpu_numa_node(local_rank)
        sec = runner.time_e2e(e2e_steps, barrier)
        sec = reduce_max(sec, world)
        e2e_err = reduce_max(getattr(runner, "e2e_max_rel_err", float("nan")), world)
        e2e = {"value": 2.0 * nnz_total / sec / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(runner.e2e_h2d_bytes),
               "d2h_bytes_per_step": int(runner.e2e_d2h_bytes), "steps": e2e_steps,
               "validated": bool(e2e_err <= tol), "max_rel_err": e2e_err,
               "api": ("uspmv_spmv_host_submit/_wait (C ABI, pinned host x/y; every step copies its own x in and its own y out, "
                       "3 steps in flight so H2D / kernel / D2H of neighbouring steps overlap)") if (world == 1 and not block_or_solve) else
                      ("uspmv_p2p_spmv_host_submit/_wait (C ABI, pinned host x/y per rank; every step copies its rank's x slab in and its y out, "
                       "2 steps in flight over the two arena buffers, halo exchange inside every step)") if getattr(runner, "e2e_pipelined", False) else
                      "host x slab -> device, halo exchange + SpMV, y -> host, sync, per step",
               "bound_by": ("host<->device copies: %.1f GB/s in + %.1f GB/s out per GPU, %.1f GB/s aggregate over %d GPU(s) (this VM exposes one NUMA "
                            "node and no PCIe topology; aggregate host memory / PCIe root bandwidth saturates beyond 2 GPUs)" %
                            (runner.e2e_h2d_bytes / sec / 1e9, runner.e2e_d2h_bytes / sec / 1e9,
                             world * (runner.e2e_h2d_bytes + runner.e2e_d2h_bytes) / sec / 1e9, world)),
               "host_numa": ({"node": bound["node"], "cpus": bound["cpus"]} if bound else "not exposed by this VM (single NUMA node, no PCIe topology in sysfs)")}
        if world == 1 and not block_or_solve:
            sec1 = runner.time_e2e(max(3, e2e_steps // 2), barrier, pipelined=False)
            e2e["single_call_value"] = 2.0 * nnz_total / sec1 / 1e9
            e2e["single_call_api"] = "uspmv_spmv_host: H2D(x) + SpMV + D2H(y) + sync, one step at a time"
        if bound:
            os.sched_setaffinity(0, bound["previous"])  # the CPU baseline below uses all host cores again

    label = workload_label(args, pts, n, world)
    head_cfg = ({"workload": label, "rows_per_gpu": runner.n_rows, "nnz_per_gpu": int(nnz_local), "n_elements_per_gpu": [int(v) for v in runner.n_elements],
                 "partition": "none" if world == 1 else f"seg_nnz over {world} ranks (work_sharing_arr {[int(v) for v in wsa]}), halo exchange every step via p2p",
                 "halo_elements_per_gpu": int(runner.n_halo), "x": "constant 1.0",
                 "l2": "matrix parts larger than the 126 MB L2; no explicit flush"} if is_ap else
                {"workload": label, "rows_per_gpu": runner.n_rows, "nnz_per_gpu": int(nnz_local), "n_elements_per_gpu": int(runner.n_elements),
                 "partition": "none" if world == 1 else f"seg_rows z-slabs x{world}, halo exchange every step (comm_halos=1) via {args.halo}",
                 "l2": "inputs (>= 1.4 GB per GPU) are larger than the 126 MB L2; no explicit flush",
                 "x": "constant 5.0 (reference default) in the timed steps; x = sin(0.37 g) + 1.5 in the checked step"})
    kernel_name = runner.kernel_name
    runner.close()
    del runner
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configs, same protocol (build -> checked step -> timed steps -> roofline) -------------------
    others = []
    if args.other_configs and not (is_ap or args.solve or args.bvs > 1):
        D = pkg.dist
        if world == 1:
            for ovt in ("dp", "sp"):
                for b in (4, 8):
                    others.append(other_config(
                        pkg, ctx, args, rank, world, local_rank, peak, f"config 3: {args.workload} SpMMV block_vec_size={b} rowwise {ovt}, scs C=32 sigma=1",
                        lambda ovt=ovt, b=b: D.DistributedSpmv(ctx, pts, n, 32, 1, ovt, rank, world, bvs=b, layout="rowwise"), ovt, bvs=b))
            if not args.no_config4:
                for plan in (("fused",) if args.no_banded else ("fused", "banded")):
                    others.append(other_config(
                        pkg, ctx, args, rank, world, local_rank, peak,
                        f"config 4: power-law matrix 2^{args.config4_log2_rows} rows, ap[dp_sp_hp] t1=1.0 t2=1e-2, scs C=32 sigma=512, plan={plan}",
                        lambda plan=plan: build_powerlaw_ap(pkg, ctx, args.config4_log2_rows, "ap[dp_sp_hp]", 32, 512, rank, world, plan=plan,
                                                            alg_bytes=(others[-1]["roofline"]["algorithmic_bytes_per_launch"] if plan == "banded" else None))[0],
                        "dp", steps=max(10, min(args.steps, 50)),
                        traffic_key=(f"config4|powerlaw_{args.config4_log2_rows}|ap[dp_sp_hp]|C32|s512" if plan == "fused" else None)))
        elif not args.strong and pts == 7:
            n5 = args.config5_n
            if n5 % world == 0:
                a5 = argparse.Namespace(**vars(args))
                a5.workload, a5.strong, a5.C, a5.sigma, a5.vt, a5.bvs, a5.solve = f"stencil27_{n5}", True, 32, 1, "dp", 1, False
                others.append(other_config(
                    pkg, ctx, args, rank, world, local_rank, peak, "config 5: " + workload_label(a5, 27, n5, world),
                    lambda: D.DistributedSpmv(ctx, 27, n5, 32, 1, "dp", rank, world, strong=True), "dp", steps=max(10, min(args.steps, 100))))
                others[-1]["scaling"] = "strong"

    cpu = None
    gpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not is_ap:
        try:
            dt, done, nnz_c, threads, build_s = cpu_reference_run(pts, n, args.C, args.sigma, vt, 10 ** 9, 10, time_box=args.cpu_seconds)
            cpu = {"value": 2.0 * nnz_c / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"full {args.workload} matrix, {done} SpMVs in a {args.cpu_seconds:.0f} s box after 10 warm-ups (reference "
                             f"spmv_omp_scs_adv via oracle/_ref; build {build_s:.1f} s untimed)"}
        except Exception as e:  # the CPU leg must never take the GPU line down
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        if not args.no_gpu_baseline and args.bvs == 1 and not args.solve:
            gpu_base = gpu_reference_kernel_run(pts, n, args.C, args.sigma, vt, max(args.steps, 50))
            if gpu_base and gpu_base.get("value"):
                gpu_base["speedup_vs_it"] = gflops / gpu_base["value"]

    if rank == 0:
        line = {
            "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak", "vs_baseline": None,
            "dtype": {"dp": "f64", "sp": "f32", "hp": "f16"}[vt], "data": "synthetic",
            "config": head_cfg,
            "validated": ok, "max_rel_err": max_err, "tolerance": tol, "exchange_errors": exchange_errors,
            "validation": "before timing, on every rank: x = sin(0.37 g) + 1.5 (g = global row), halo tail of x poisoned with NaN, 2 steps, "
                          "y against the product formed from the matrix definition (no SELL-C-sigma structure, no exchange involved); max over ranks of "
                          "|y - ref| / sum|a||x|; after the timed steps the arena's error word is read (exchange_errors)",
            "gbs": bytes_total / sec_per_step / 1e9,
            "timing": ("CUDA events on the launching stream around exactly K steps, max over ranks" +
                       ("; after the host barrier ONE untimed step aligns the ranks' device timelines (host barrier exit skew is not part of a step)"
                        if world > 1 else "")),
            "steady_state": ({"steps": args.steady_steps, "ms_per_step": steady_ms, "value": 2.0 * nnz_total / (steady_ms / 1e3) / 1e9, "unit": UNIT}
                             if steady_ms else None),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "kernel": kernel_name, "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_launch": int(bytes_local), "peak_source": peak_src},
            "library_options": args.set or None,
            "cpu_baseline": cpu, "gpu_baseline": gpu_base, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "other_configs": others,
        }
        print(json.dumps(line), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()
    if exchange_errors or any(not o["validated"] for o in others):
        sys.exit(4)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
