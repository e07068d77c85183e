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
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="N>1: NVLink peer stores + epoch flags, or NCCL send/recv")
    ap.add_argument("--bvs", type=int, default=1, help="block_vec_size > 1: SpMMV (config 3), block-vector halo exchange when N>1")
    ap.add_argument("--layout", default="rowwise", choices=["rowwise", "colwise"], help="block vector layout for --bvs > 1")
    ap.add_argument("--solve", action="store_true", help="solve mode: a step is { halo exchange ; SpMV ; swap } on two device buffers")
    return ap.parse_args()


def workload_dims(name):
    kind, n = name.rsplit("_", 1)
    n = int(n)
    pts = {"laplace7": 7, "stencil27": 27, "powerlaw": 0}[kind]
    return pts, n


def build_powerlaw_ap(pkg, ctx, log2_rows, mode, C, sigma, rank, world):
    """BASELINE config 4: power-law matrix with 2^log2_rows rows, rows split by seg-nnz, per-rank
    partition_precisions with t1 = 1.0, t2 = 1e-2.  Every rank generates only its own rows (twice: once on an equal-rows split to
    count the elements per row for the seg-nnz walk, once on the final split)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    mats, d = pkg.matrices, pkg.dist
    n = 1 << log2_rows
    target = n * 15
    slab = 1 << 20

    def gen(r0, r1):
        parts = [mats.powerlaw_coo(n, target, row0=a, row1=min(r1, a + slab)) for a in range(r0, r1, slab)]
        I = np.concatenate([p[2].astype(np.int64) + (a - r0) for p, a in zip(parts, range(r0, r1, slab))]).astype(np.int32)
        return I, np.concatenate([p[3] for p in parts]), np.concatenate([p[4] for p in parts])
    if world > 1:
        eq = np.arange(world + 1, dtype=np.int64) * (n // world)
        eq[world] = n
        I, J, V = gen(int(eq[rank]), int(eq[rank + 1]))
        cnt = torch.from_numpy(np.bincount(I, minlength=int(eq[rank + 1] - eq[rank])).astype(np.int32)).cuda()
        sizes = [int(eq[r + 1] - eq[r]) for r in range(world)]
        bufs = [torch.empty(sz, dtype=torch.int32, device="cuda") for sz in sizes]
        dist.all_gather(bufs, cnt)
        wsa = d.seg_nnz_from_row_counts(torch.cat(bufs).cpu().numpy(), world)
        if (int(wsa[rank]), int(wsa[rank + 1])) != (int(eq[rank]), int(eq[rank + 1])):
            I, J, V = gen(int(wsa[rank]), int(wsa[rank + 1]))
    else:
        wsa = np.array([0, n], np.int32)
        I, J, V = gen(0, n)
    n_loc = int(wsa[rank + 1] - wsa[rank])
    return d.DistributedApSpmv(ctx, wsa, (n_loc, n, I, J, V), mode, 1.0, 1e-2, C, sigma, rank, world), wsa


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
    return dt, done, nnz, ref.omp_threads(), build_s


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
        "data": "synthetic", "config": {"workload": f"{args.workload} scs C={args.C} sigma={args.sigma} {args.vt} SpMV (one rank's slab)",
                                        "kernel": "reference spmv_omp_scs_adv (kernels.hpp:265-301) via oracle/_ref, built by the reference's convert_to_scs",
                                        "x": "constant 5.0"},
        "cpu_baseline": {"value": gf, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"full {args.workload} matrix ({nnz} nnz), {done} SpMVs after {warmup} warm-ups; build {build_s:.1f} s untimed"},
        "e2e": {"value": gf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
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
    elif block_or_solve:  # the distributed runner also serves N = 1 for these modes (an arena with no peers)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{29400 + os.getpid() % 500}", rank=0, world_size=1,
                                device_id=torch.device("cuda", local_rank))

    pkg = importlib.import_module("ultimate-spmv_b200")
    eng, capi = pkg.engine, pkg.capi
    pts, n = workload_dims(args.workload)
    vt = args.vt
    vsize = {"dp": 8, "sp": 4, "hp": 2}[vt]
    tdt = {"dp": torch.float64, "sp": torch.float32, "hp": torch.float16}[vt]
    ctx = eng.default_context(local_rank)

    wsa = None
    if is_ap:
        runner, wsa = build_powerlaw_ap(pkg, ctx, n, args.ap, args.C, args.sigma, rank, world)
        vt = "sp" if args.ap == "ap[sp_hp]" else "dp"
    elif world == 1 and not block_or_solve:
        runner = pkg.engine.SingleGpuSpmv(ctx, pts, n, args.C, args.sigma, vt)
    else:
        runner = pkg.dist.DistributedSpmv(ctx, pts, n, args.C, args.sigma, vt, rank, world, halo=args.halo, overlap=not args.no_overlap,
                                          strong=args.strong, bvs=args.bvs, layout=args.layout,
                                          n_buf=2 if (args.solve or (args.bvs == 1 and args.halo == "p2p")) else 1)
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

    for _ in range(max(args.warmup, 3)):
        runner.step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = capi.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        runner.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = capi.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(nnz_local), float(bytes_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    nnz_total, bytes_total = float(tot[0].item()), float(tot[1].item())
    sec_per_step = ms / 1e3 / args.steps
    gflops = 2.0 * nnz_total / sec_per_step / 1e9

    # kernel-only duration of the dominant kernel (SpMV) measured with CUDA events on its stream
    kern_ms = runner.time_kernel(args.steps)
    peak, peak_src = measured_peak()
    achieved = bytes_local / (kern_ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(f"{args.workload}|C{args.C}|s{args.sigma}|{vt}")

    # end-to-end through the host-buffer C-ABI call (pinned host x / y, copies inside the timed region)
    e2e = None
    if not args.no_e2e and not args.solve and not is_ap:
        e2e_steps = max(3, min(args.steps, 20))
        # pinned staging buffers next to the GPU's PCIe root: first-touch under an affinity restricted to the GPU's NUMA node
        bound = None if args.no_numa_bind else pkg.dist.bind_to_gpu_numa_node(local_rank)
        sec = runner.time_e2e(e2e_steps, barrier)
        te = torch.tensor([sec], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": 2.0 * nnz_total / float(te.item()) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(runner.e2e_h2d_bytes),
               "d2h_bytes_per_step": int(runner.e2e_d2h_bytes), "steps": e2e_steps,
               "api": ("uspmv_spmv_host_submit/_wait (C ABI, pinned host x/y; every step copies its own x in and its own y out, "
                       "3 steps in flight so H2D / kernel / D2H of neighbouring steps overlap)") if (world == 1 and not block_or_solve) else
                      ("uspmv_p2p_spmv_host_submit/_wait (C ABI, pinned host x/y per rank; every step copies its rank's x slab in and its y out, "
                       "2 steps in flight over the two arena buffers, halo exchange inside every step)") if getattr(runner, "e2e_pipelined", False) else
                      "host x slab -> device, halo exchange + SpMV, y -> host, sync, per step",
               "host_numa": ({"node": bound["node"], "cpus": bound["cpus"]} if bound else None)}
        if world == 1 and not block_or_solve:
            sec1 = runner.time_e2e(max(3, e2e_steps // 2), barrier, pipelined=False)
            e2e["single_call_value"] = 2.0 * nnz_total / sec1 / 1e9
            e2e["single_call_api"] = "uspmv_spmv_host: H2D(x) + SpMV + D2H(y) + sync, one step at a time"
        if bound:
            os.sched_setaffinity(0, bound["previous"])  # the CPU baseline below uses all host cores again

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not is_ap:
        try:
            dt, done, nnz_c, threads, build_s = cpu_reference_run(pts, n, args.C, args.sigma, vt, 10 ** 9, 10, time_box=args.cpu_seconds)
            cpu = {"value": 2.0 * nnz_c / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"full {args.workload} matrix, {done} SpMVs in a {args.cpu_seconds:.0f} s box after 10 warm-ups (reference "
                             f"spmv_omp_scs_adv via oracle/_ref; build {build_s:.1f} s untimed)"}
        except Exception as e:  # the CPU leg must never take the GPU line down
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak", "vs_baseline": None,
            "dtype": {"dp": "f64", "sp": "f32", "hp": "f16"}[vt], "data": "synthetic",
            "config": {"workload": f"{args.workload}: power-law matrix with 2^{n} rows (config 4), {args.ap} t1=1.0 t2=1e-2, scs C={args.C} sigma={args.sigma}"} |
                      {"rows_per_gpu": runner.n_rows, "nnz_per_gpu": int(nnz_local), "n_elements_per_gpu": [int(v) for v in runner.n_elements],
                       "partition": "none" if world == 1 else f"seg_nnz over {world} ranks (work_sharing_arr {[int(v) for v in wsa]}), halo exchange every step via p2p",
                       "halo_elements_per_gpu": int(runner.n_halo), "x": "constant 1.0",
                       "l2": "matrix parts larger than the 126 MB L2; no explicit flush"} if is_ap else
                      {"workload": (f"{args.workload}: {pts}-point stencil on ONE {n}^3 grid cut into {world} z-slabs" if (args.strong and world > 1) else
                                    f"{args.workload}: {pts}-point stencil on a {n}^3 grid per GPU") + f", scs C={args.C} sigma={args.sigma} {vt} " +
                                   (f"SpMMV block_vec_size={args.bvs} {args.layout}" if args.bvs > 1 else "SpMV") +
                                   (", solve mode: each step = halo exchange + SpMV + swap on two device buffers" if args.solve else ""),
                       "rows_per_gpu": runner.n_rows, "nnz_per_gpu": int(nnz_local), "n_elements_per_gpu": int(runner.n_elements),
                       "partition": "none" if world == 1 else f"seg_rows z-slabs x{world}, halo exchange every step (comm_halos=1) via {args.halo}",
                       "l2": "inputs (>= 1.4 GB per GPU) are larger than the 126 MB L2; no explicit flush",
                       "x": "constant 5.0 (reference default)"},
            "gbs": bytes_total / sec_per_step / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": runner.kernel_name, "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": int(bytes_local), "peak_source": peak_src},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
