#!/usr/bin/env python
"""bench.py -- headline benchmark of the banded hot path (BASELINE.json).

A "step" = one gbmv  y <- A*x  over the C2 workload: Float64, n = 2^27, (l,u) = (4,3)  [80*n bytes].
  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference ...                        the reference's CPU path (OpenBLAS dgbmv_) on host cores
Prints ONE JSON line.  See DESIGN.md "Measurement" for what every field means.
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

METRIC = "gbmv HBM GB/s"
UNIT = "GB/s"
N_C2, KL, KU = 1 << 27, 4, 3
LDA = KL + KU + 1


def algo_bytes(n, lda=LDA, beta_nonzero=False):
    """8*n*lda (band data) + 8n (x) + 8n (y written) [+ 8n when y is also read]."""
    return 8 * n * (lda + 2 + (1 if beta_nonzero else 0))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference's own path = OpenBLAS dgbmv_ (src/generic/matmul.jl:21-23), driven from C
# ---------------------------------------------------------------------------------------------------
def cpu_driver():
    import ctypes as C

    import oracle

    oracle.build()
    L = C.CDLL(os.path.join(ROOT, "oracle", "libblasdriver.so"))
    assert L.drv_open(oracle.openblas_path().encode()) == 0
    L.drv_gbmv.restype = C.c_double
    L.drv_gbmv.argtypes = [C.c_int64] * 4 + [C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_void_p]
    return L


def cpu_gbmv_time(L, n, data, x, y, threads=1, reps=3):
    L.drv_set_threads(int(threads))
    best = 1e30
    for _ in range(reps):
        t = L.drv_gbmv(n, n, KL, KU, 1.0, data.ctypes.data, LDA, x.ctypes.data, 0.0, y.ctypes.data)
        best = min(best, t)
    return best


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (OpenBLAS 0.3.30 dgbmv_64_, the Fortran
    entry point BandedMatrices.jl ccalls) on the box's host cores, same metric / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L = cpu_driver()
    n = args.n if args.n else (1 << 25)  # bounded sample of the C2 workload: 2^25 rows (2.7 GB), ~0.4 s per step
    rng = np.random.default_rng(1)
    data = np.asfortranarray(rng.random((LDA, n)))
    x = rng.random(n)
    y = np.zeros(n)
    cores = os.cpu_count() or 1
    # OpenBLAS runs dgbmv single-threaded for kl+ku < 15 whatever the thread setting; give it every core anyway
    L.drv_set_threads(cores)
    for _ in range(max(1, args.warmup)):
        L.drv_gbmv(n, n, KL, KU, 1.0, data.ctypes.data, LDA, x.ctypes.data, 0.0, y.ctypes.data)
    t = 0.0
    for _ in range(args.steps):
        t += L.drv_gbmv(n, n, KL, KU, 1.0, data.ctypes.data, LDA, x.ctypes.data, 0.0, y.ctypes.data)
    ms = 1e3 * t / args.steps
    val = algo_bytes(n) / (ms * 1e-3) / 1e9
    sample = f"n=2^{int(np.log2(n))} rows of the C2 workload per step (OpenBLAS 0.3.30 dgbmv_64_, threads={cores}; narrow-band gbmv is single-threaded inside OpenBLAS)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 gbmv Float64 (l,u)=(4,3), y=A*x", "n": n, "bytes_per_step": algo_bytes(n)},
        "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import bandedmatrices_b200 as bm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n if args.n else N_C2
    # strong scaling: the n rows (= data columns) are cut into `world` contiguous slabs
    c0, c1 = (n * rank) // world, (n * (rank + 1)) // world
    nl = c1 - c0
    hd = bm.handle(local)
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    A = bm.BandedMatrix(torch.rand((nl, LDA), dtype=torch.float64, device="cuda", generator=g), nl, KL, KU)
    x = torch.rand(nl, dtype=torch.float64, device="cuda", generator=g)
    y = torch.empty(nl, dtype=torch.float64, device="cuda")

    if world > 1:
        from bandedmatrices_b200.sharded import ShardedGbmv

        op = ShardedGbmv(n, c0, c1, KL, KU, A, rank, world)
        step = lambda: op(1.0, x, 0.0, y)  # noqa: E731
    else:
        step = lambda: bm.mul_(y, A, x, 1.0, 0.0)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = hd.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev0.record()
    for i in range(args.steps):
        step()
        marks[i].record()  # per-launch boundaries on the launching stream, inside the timed region (no host sync)
    ev1.record()
    barrier()
    launches = hd.launches - l0
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms = float(t.item()) / args.steps
    value = algo_bytes(n) / (ms * 1e-3) / 1e9

    # ---- per-launch duration of the dominant kernel for the roofline: consecutive event marks of the timed region itself
    # (rank-local, launching stream, queue never empty => no launch gaps, and at N>1 the ranks are in their steady state;
    # timing launches one by one with a host sync in between would add launch latency and cross-rank start skew) ----
    bounds = [ev0] + marks
    kt = [bounds[i].elapsed_time(bounds[i + 1]) for i in range(args.steps)]
    k_ms = float(np.mean(kt))
    peak, peak_src = peaks()
    achieved = algo_bytes(nl) / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gbmv_c2_traffic.json")
    if os.path.exists(tp) and world == 1 and n == N_C2:
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2 gbmv Float64 n=2^27 (l,u)=(4,3), y=A*x (alpha=1,beta=0), row-sharded over n_gpus",
                       "n": n, "bytes_per_step": algo_bytes(n), "l2": "inputs (10.7 GB) larger than L2; no flush needed",
                       "parallelism": f"rows/{world}" + (" + x halo over NVLink peer stores" if world > 1 else "")},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                         "kernel": "gbmv_n_systolic<8,8>", "kernel_ms": round(k_ms, 4), "kernel_ms_min": round(float(np.min(kt)), 4),
                         "frac_of_nominal_8TBs": round(achieved / 8000.0, 4)},
            "clocks": clocks, "gpu_launches": launches,
        }

    # ---- e2e through the host-buffer C ABI (pinned host arrays, H2D + kernel + D2H inside the timed region) ----
    # At N > 1 every rank moves its own row slab through its own PCIe link at the same time; the value is the whole-job
    # aggregate (all rows / max-over-ranks time).  The host-buffer entry point is per slab (no x halo), so this is a
    # throughput measurement of the same work, not a sharded product.
    if not args.no_e2e:
        ne = n if world == 1 else nl
        hA = torch.empty((ne, LDA), dtype=torch.float64).pin_memory()
        hx = torch.empty(ne, dtype=torch.float64).pin_memory()
        hy = torch.empty(ne, dtype=torch.float64).pin_memory()
        hA.copy_(A.data)  # device -> pinned host, direct
        hx.copy_(x)
        torch.cuda.synchronize()
        dA_np, x_np, y_np = hA.numpy().T, hx.numpy(), hy.numpy()  # (LDA x n) Fortran view of the same memory
        bm.gbmv_host("N", ne, KL, KU, 1.0, dA_np, x_np, 0.0, y_np, device=local)  # warm-up (scratch allocation)
        reps = 3
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            bm.gbmv_host("N", ne, KL, KU, 1.0, dA_np, x_np, 0.0, y_np, device=local)
        dt = (time.perf_counter() - t0) / reps
        if world > 1:
            td = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dt = float(td.item())
        same = bool(torch.equal(hy.cuda(), y)) if world == 1 else None
    if rank == 0 and not args.no_e2e:
        out["e2e"] = {"value": round(algo_bytes(n) / dt / 1e9, 2), "unit": UNIT, "h2d_bytes_per_step": 8 * n * (LDA + 1),
                      "d2h_bytes_per_step": 8 * n, "ms_per_step": round(dt * 1e3, 2),
                      "api": "bmb200_dgbmv_host (pinned host arrays" + ("" if world == 1 else "; one row slab per rank, concurrently") + ")",
                      "matches_device_path": same, "rows": n}

        # ---- CPU baseline beside it: OpenBLAS dgbmv_ on the same host arrays, bounded sample ----
        if not args.no_cpu:
            try:
                L = cpu_driver()
                ns = min(ne, 1 << 26)  # 2^26 rows = half of C2: ~0.75 s per call
                yc = np.zeros(ns)
                tb = cpu_gbmv_time(L, ns, dA_np[:, :ns], x_np[:ns], yc, threads=1, reps=3)
                bit_same = bool(np.array_equal(yc[: ns - KU], y_np[: ns - KU]))
                out["cpu_baseline"] = {"value": round(algo_bytes(ns) / tb / 1e9, 3), "unit": UNIT, "cores": 1, "kind": "reference",
                                       "sample": f"first 2^{int(np.log2(ns))} rows of the same C2 inputs, OpenBLAS 0.3.30 dgbmv_64_ "
                                                 f"(the Fortran entry point the reference ccalls; single-threaded inside OpenBLAS for kl+ku<15; "
                                                 f"host has {os.cpu_count()} cores), best of 3",
                                       "gpu_result_bit_identical_on_sample": bit_same}
            except Exception as e:  # the baseline is reporting only; never let it kill the bench line
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"failed: {e}"}
    if rank == 0 and args.extras and world == 1:
        try:
            from bench_extras import run_extras

            out["extras"] = run_extras(bm)
            if not args.no_cpu:
                from bench_extras import cpu_extras

                out["extras"]["cpu_reference"] = cpu_extras(cpu_driver(), os.cpu_count() or 1)
        except Exception as e:  # noqa: BLE001
            out["extras"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=0, help="override the row count (debugging only; the headline is n=2^27)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--extras", action="store_true", help="also time C1/C3/C4 (banded matmul, LU+solve) into 'extras'")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
