#!/usr/bin/env python
"""bench.py -- headline benchmark of the banded hot path (BASELINE.json).

A "step" = one gbmv  y <- A*x  over the C2 workload: Float64, n = 2^27, (l,u) = (4,3)  [80*n bytes].
  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference ...                        the reference's CPU path (OpenBLAS dgbmv_) on host cores
Prints ONE JSON line.  Besides the headline (C2) it carries "configs": the other BASELINE.json configurations (C1, C3, C4,
C5 -- banded matmul and LU+solve), each timed, set against its roofline and the CPU reference, and verified against OpenBLAS
at the full benchmark size (bench_configs.py).  See DESIGN.md "Measurement" for what every field means.
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
    """SM clock / throttle reasons sampled DURING the timed region: an in-process NVML poller (5 ms period; the timed region of
    the default run is ~35 ms, too short for an `nvidia-smi -lms` child to get a sample in), nvidia-smi as the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.nv, self.h = index, [], None, None, None
        self.sm, self.reason_bits, self.mx, self.run = [], 0, None, False
        try:
            import pynvml as nv
            nv.nvmlInit()
            try:
                pr = torch.cuda.get_device_properties(index)
                self.h = nv.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
            except Exception:
                self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.nv = nv
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self.run = True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
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
        if self.nv is not None:
            self.run = False
            self.t.join(timeout=1.0)
            nv, bits = self.nv, self.reason_bits
            table = [("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", 0x8), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", 0x20), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", 0x4)]
            reasons = sorted(name for name, attr, dflt in table if bits & int(getattr(nv, attr, dflt)))
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                    "samples": len(self.sm), "source": "nvml, 5 ms period, inside the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


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
    n = args.n if args.n else N_C2  # the full C2 workload (10.7 GB of host arrays, ~1.5 s per step): same config as our arm
    rng = np.random.default_rng(1)
    blk = 1 << 20  # a random 2^20-column block tiled over the matrix: the values do not matter to dgbmv_'s speed
    data = np.empty((LDA, n), order="F")
    tile = np.asfortranarray(rng.random((LDA, min(blk, n))))
    for j in range(0, n, blk):
        data[:, j:j + blk] = tile[:, : min(blk, n - j)]
    x = np.tile(rng.random(min(blk, n)), -(-n // blk))[:n].copy()
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
    sample = f"the full C2 workload per step, n=2^{int(np.log2(n))} rows (OpenBLAS 0.3.30 dgbmv_64_, threads={cores}; narrow-band gbmv is single-threaded inside OpenBLAS)"
    configs = None
    if args.configs:
        try:  # the reference's CPU timings of the other BASELINE configs (bounded samples; see bench_configs.py)
            from bench_extras import cpu_extras

            configs = cpu_extras(L, cores)
        except Exception as e:  # noqa: BLE001
            configs = {"error": repr(e)}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 gbmv Float64 n=2^27 (l,u)=(4,3), y=A*x (alpha=1,beta=0), row-sharded over n_gpus", "n": n,
                   "bytes_per_step": algo_bytes(n)},
        "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "configs": configs,
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

    sampler = ClockSampler(local)
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
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
                         "frac": round(achieved / peak, 4), "traffic": traffic,
                         "traffic_source": "ncu --set full capture of this command committed as profiles/gbmv_c2_traffic.json (not re-measured in this run)" if traffic else None,
                         "peak_source": peak_src,
                         "kernel": "gbmv_n_systolic<8,8>", "kernel_ms": round(k_ms, 4), "kernel_ms_min": round(float(np.min(kt)), 4),
                         "frac_of_nominal_8TBs": round(achieved / 8000.0, 4)},
            "clocks": clocks, "gpu_launches": launches,
        }

    # ---- N > 1: the sharded result checked against OpenBLAS on the rows that straddle EVERY slab boundary ----
    sharded_check = None
    if world > 1:
        Wd = 128  # columns gathered either side of a boundary; rows within 64 of it are compared
        pack = torch.zeros((2, Wd, LDA + 2), dtype=torch.float64, device="cuda")
        pack[0, :, :LDA], pack[0, :, LDA], pack[0, :, LDA + 1] = A.data[:Wd], x[:Wd], y[:Wd]
        pack[1, :, :LDA], pack[1, :, LDA], pack[1, :, LDA + 1] = A.data[nl - Wd:], x[nl - Wd:], y[nl - Wd:]
        allp = [torch.empty_like(pack) for _ in range(world)]
        dist.all_gather(allp, pack)
        if rank == 0 and not args.no_cpu:
            try:
                L = cpu_driver()
                L.drv_set_threads(1)
                bad, rows = 0, 0
                for r in range(1, world):
                    win = torch.cat([allp[r - 1][1], allp[r][0]], dim=0).cpu().numpy()  # 2*Wd columns around the boundary
                    d_h = np.asfortranarray(win[:, :LDA].T)
                    x_h, y_gpu = np.ascontiguousarray(win[:, LDA]), win[:, LDA + 1]
                    y_h = np.zeros(2 * Wd)
                    L.drv_gbmv(2 * Wd, 2 * Wd, KL, KU, 1.0, d_h.ctypes.data, LDA, x_h.ctypes.data, 0.0, y_h.ctypes.data)
                    sl = slice(Wd - 64, Wd + 64)  # interior rows of the window: same terms, same order as in the full product
                    bad += int((y_h[sl] != y_gpu[sl]).sum())
                    rows += 128
                sharded_check = {"sharded_bit_identical": bad == 0, "boundary_rows_checked": rows, "mismatches": bad,
                                 "against": "OpenBLAS dgbmv_64_ on the 256-column windows around every slab boundary"}
            except Exception as e:  # noqa: BLE001
                sharded_check = {"sharded_bit_identical": None, "error": repr(e)}
        op.check()  # raises if a halo wait timed out anywhere in the run
    if rank == 0 and sharded_check is not None:
        out["sharded_check"] = sharded_check

    # ---- e2e through the public host-buffer entry points (pinned host arrays, H2D + kernel + D2H inside the timed region).
    # N = 1: bmb200_dgbmv_host (chunked H2D overlapped with the kernel).  N > 1: the SHARDED product from host buffers --
    # every rank uploads its extended slab (static data halo included) and its slice of x, the x halo moves between the GPUs
    # inside the kernel exactly as in the device-resident run, and y comes back to the host; aggregate over ranks. ----
    if not args.no_e2e:
        reps = 3
        if world == 1:
            hA = torch.empty((n, LDA), dtype=torch.float64).pin_memory()
            hx = torch.empty(n, dtype=torch.float64).pin_memory()
            hy = torch.empty(n, dtype=torch.float64).pin_memory()
            hA.copy_(A.data)  # device -> pinned host, direct
            hx.copy_(x)
            torch.cuda.synchronize()
            dA_np, x_np, y_np = hA.numpy().T, hx.numpy(), hy.numpy()  # (LDA x n) Fortran view of the same memory
            e2e_step = lambda: bm.gbmv_host("N", n, KL, KU, 1.0, dA_np, x_np, 0.0, y_np, device=local)  # noqa: E731
            h2d, d2h = 8 * n * (LDA + 1), 8 * n
            api = "bmb200_dgbmv_host (pinned host arrays)"
        else:
            ext = op.data_ext
            hA = torch.empty(ext.shape, dtype=torch.float64).pin_memory()
            hx = torch.empty(nl, dtype=torch.float64).pin_memory()
            hy = torch.empty(nl, dtype=torch.float64).pin_memory()
            hA.copy_(ext)
            hx.copy_(x)
            torch.cuda.synchronize()
            y2 = torch.empty_like(y)

            def e2e_step():
                ext.copy_(hA, non_blocking=True)
                x.copy_(hx, non_blocking=True)
                op(1.0, x, 0.0, y2)
                hy.copy_(y2, non_blocking=True)
                torch.cuda.synchronize()

            h2d, d2h = 8 * (ext.numel() + nl), 8 * nl
            api = "ShardedGbmv from pinned host buffers (extended slab + x slice up, in-kernel x halo over NVLink, y slice down), all ranks concurrently"
        e2e_step()  # warm-up (scratch allocation)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            e2e_step()
        if world > 1:
            barrier()
        dt = (time.perf_counter() - t0) / reps
        if world > 1:
            td = torch.tensor([dt, float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
            tmax = td[:1].clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(td)
            dt, h2d, d2h = float(tmax.item()), int(td[1].item()), int(td[2].item())
            same_t = torch.tensor([1 if torch.equal(hy.cuda(), y) else 0], device="cuda")
            dist.all_reduce(same_t, op=dist.ReduceOp.MIN)
            same = bool(same_t.item())
        else:
            same = bool(torch.equal(hy.cuda(), y))
    if rank == 0 and not args.no_e2e:
        out["e2e"] = {"value": round(algo_bytes(n) / dt / 1e9, 2), "unit": UNIT, "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "ms_per_step": round(dt * 1e3, 2), "api": api,
                      "matches_device_path": same, "rows": n}

        # ---- CPU baseline beside it: OpenBLAS dgbmv_ on the same host arrays, bounded sample ----
        if not args.no_cpu:
            try:
                L = cpu_driver()
                ne = n if world == 1 else nl
                ns = min(ne, 1 << 26)  # 2^26 rows = half of C2: ~0.75 s per call
                hs = max(0, (hA.shape[0] - ne) if world > 1 else 0)  # rank 0's slab has no left halo
                dA_s, x_s, y_s = hA.numpy().T[:, :ns], hx.numpy()[:ns], hy.numpy()[:ns]
                yc = np.zeros(ns)
                tb = cpu_gbmv_time(L, ns, dA_s, x_s, yc, threads=1, reps=3)
                bit_same = bool(np.array_equal(yc[: ns - KU], y_s[: ns - KU]))
                out["cpu_baseline"] = {"value": round(algo_bytes(ns) / tb / 1e9, 3), "unit": UNIT, "cores": 1, "kind": "reference",
                                       "sample": f"first 2^{int(np.log2(ns))} rows of the same C2 inputs, OpenBLAS 0.3.30 dgbmv_64_ "
                                                 f"(the Fortran entry point the reference ccalls; single-threaded inside OpenBLAS for kl+ku<15; "
                                                 f"host has {os.cpu_count()} cores), best of 3",
                                       "gpu_result_bit_identical_on_sample": bit_same}
            except Exception as e:  # the baseline is reporting only; never let it kill the bench line
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"failed: {e}"}
        del hA, hx, hy
    # ---- free the headline's buffers, then the other BASELINE configs (C1, C3, C4, C5), timed and verified ----
    del A, x, y
    if world > 1:
        op.close()
        del op
    torch.cuda.empty_cache()
    if args.configs:
        from bench_configs import cpu_protos, run_configs

        Lc = None
        if not args.no_cpu:
            try:
                Lc = cpu_protos(cpu_driver())
            except Exception:  # noqa: BLE001
                Lc = None
        cfg = run_configs(bm, Lc, peak, rank, world)
        if rank == 0:
            out["configs"] = cfg
            out["gpu_launches_configs"] = hd.launches - l0 - launches if world == 1 else None
    if rank == 0 and args.extras and world == 1:
        try:
            from bench_extras import run_secondary

            out["extras"] = run_secondary(bm)
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
    ap.add_argument("--no-configs", dest="configs", action="store_false",
                    help="headline only: skip the C1/C3/C4/C5 entries (banded matmul, LU+solve) of 'configs'")
    ap.add_argument("--extras", action="store_true", help="also time the triangular / symmetric band and elementwise kernels into 'extras'")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
