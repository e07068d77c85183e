"""BASELINE config 4 on one GPU at the specified parameters (2^25 rows, ~5.0e8 nnz, C = 32, sigma = 512, ap[dp_sp_hp] t1 = 1, t2 = 1e-2):
the fused AP kernel against the column-banded plan at K = 4 / 8 / 16 bands, plus plain dp and sigma = 16384 for comparison.
Writes one JSON file (argv[1]).  python scripts/config4_probe.py out.json [log2_rows] [sigma,...] [K,...]"""
import importlib, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("ultimate-spmv_b200"); eng, capi = pkg.engine, pkg.capi
out_path = sys.argv[1]
LOG2 = int(sys.argv[2]) if len(sys.argv) > 2 else 25
SIGMAS = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [512]
KS = [int(v) for v in sys.argv[4].split(",") if v] if len(sys.argv) > 4 else [4, 8, 16]
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6458.1
n = 1 << LOG2
mode, t1, t2 = "ap[dp_sp_hp]", 1.0, 1e-2
res = {"log2_rows": LOG2, "peak_gbs": PEAK, "cases": []}

def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

t0 = time.time(); mtx = eng.MtxData.powerlaw(n); torch.cuda.synchronize(); gen_s = time.time() - t0
res["nnz"], res["generate_s"] = mtx.nnz, round(gen_s, 2)
print(f"generated 2^{LOG2} rows, {mtx.nnz} nnz ({mtx.nnz / n:.3f} per row) in {gen_s:.2f} s", flush=True)
x = torch.rand(n, dtype=torch.float64, device="cuda") + 0.5
vs = (8, 4, 2)
for sigma in SIGMAS:
    # ---- fused AP kernel on the reference-format structures
    t0 = time.time()
    coos = eng.partition_precisions(mtx, mode, t1, t2)
    parts = [None] * 3
    parts[0] = eng.convert_to_scs(coos[0], 32, sigma, "dp")
    torch.cuda.synchronize(); sort_s = time.time() - t0
    o2n = parts[0].export().old_to_new if False else None
    import ctypes as C
    perm_d = pkg.validate.device_int_tensor(parts[0].device_arrays()["old_to_new"].value, n, x.device)
    perm_h = perm_d.cpu().numpy()
    parts[1] = eng.convert_to_scs(coos[1], 32, sigma, "sp", fixed_permutation=perm_h)
    parts[2] = eng.convert_to_scs(coos[2], 32, sigma, "hp", fixed_permutation=perm_h)
    del coos
    torch.cuda.synchronize(); build_s = time.time() - t0
    ne = [p.n_elements for p in parts]
    alg = sum(ne[k] * (vs[k] + 4) + 8 * parts[0].n_chunks for k in range(3)) + 8 * n + 8 * parts[0].n_rows_padded
    nnz_bytes = sum(int(c) * (vs[k] + 4) for k, c in enumerate([p.nnz for p in parts])) + 16 * n
    y = torch.zeros(parts[0].n_rows_padded, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: eng.ap_spmv(mode, parts[0], parts[1], parts[2], x, y), 10)
    y_fused = y.clone()
    for av in [int(v) for v in os.environ.get("AP_VARIANTS", "").split(",") if v]:
        capi.set_option("ap_variant", av)
        msv = timeit(lambda: eng.ap_spmv(mode, parts[0], parts[1], parts[2], x, y), 10)
        print(json.dumps({"plan": "fused", "sigma": sigma, "ap_variant": av, "ms": msv, "same_y": bool(torch.equal(y, y_fused))}), flush=True)
        res["cases"].append({"plan": "fused", "sigma": sigma, "ap_variant": av, "ms": msv})
    capi.set_option("ap_variant", 0)
    case = {"plan": "fused", "sigma": sigma, "ms": ms, "n_elements": ne, "part_nnz": [p.nnz for p in parts], "algorithmic_bytes": alg, "bytes_without_padding": nnz_bytes,
            "frac_of_peak_algorithmic": alg / (ms / 1e3) / 1e9 / PEAK, "gflops": 2 * mtx.nnz / (ms / 1e3) / 1e9, "build_s": round(build_s, 2),
            "first_part_build_s": round(sort_s, 2)}
    print(json.dumps(case), flush=True); res["cases"].append(case)
    del parts
    torch.cuda.empty_cache()
    # ---- plain dp
    try:
        t0 = time.time(); s = eng.convert_to_scs(mtx, 32, sigma, "dp"); torch.cuda.synchronize(); b = time.time() - t0
        yd = torch.zeros(s.n_rows_padded, dtype=torch.float64, device="cuda")
        ms = timeit(lambda: eng.spmv_unpermuted(s, x, yd), 10)
        algd = s.n_elements * 12 + 8 * s.n_chunks + 16 * n
        case = {"plan": "plain dp (unpermuted x / y)", "sigma": sigma, "ms": ms, "n_elements": s.n_elements, "algorithmic_bytes": algd,
                "frac_of_peak_algorithmic": algd / (ms / 1e3) / 1e9 / PEAK, "gflops": 2 * mtx.nnz / (ms / 1e3) / 1e9, "build_s": round(b, 2)}
        print(json.dumps(case), flush=True); res["cases"].append(case)
        del s, yd
    except Exception as e:  # sigma = 512: 2.67e9 stored elements exceed the reference's int index type in ONE structure
        print(f"plain dp sigma={sigma}: {e}", flush=True)
        res["cases"].append({"plan": "plain dp", "sigma": sigma, "error": str(e)})
    torch.cuda.empty_cache()
    # ---- column-banded plan
    for K in KS:
        t0 = time.time()
        try:
            plan = eng.BandedPlan(mtx, 32, sigma, "dp", ap=mode, t1=t1, t2=t2, n_bands=K)
        except Exception as e:
            print(f"banded K={K} sigma={sigma}: {e}", flush=True)
            res["cases"].append({"plan": f"banded K={K}", "sigma": sigma, "error": str(e)})
            continue
        torch.cuda.synchronize(); b = time.time() - t0
        yb = torch.zeros(plan.n_rows_padded, dtype=torch.float64, device="cuda")
        ms = timeit(lambda: plan.spmv(x, yb), 10)
        # same row order as the fused structures (sigma-sort of the dp part) -> compare directly
        scale = float(y_fused.abs().max())
        err = float((yb - y_fused).abs().max()) / scale
        case = {"plan": f"banded K={K}", "sigma": sigma, "ms": ms, "n_elements_all_bands": plan.n_elements, "algorithmic_bytes": alg,
                "frac_of_peak_algorithmic": alg / (ms / 1e3) / 1e9 / PEAK, "gflops": 2 * mtx.nnz / (ms / 1e3) / 1e9, "build_s": round(b, 2),
                "max_abs_diff_vs_fused_over_max_abs_y": err}
        print(json.dumps(case), flush=True); res["cases"].append(case)
        del plan, yb
        torch.cuda.empty_cache()
json.dump(res, open(out_path, "w"), indent=1)
