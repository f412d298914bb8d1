import sys, torch
sys.path.insert(0, ".")
import bandedmatrices_b200 as bm
for n, k in ((1 << 20, 4), (1 << 20, 40), (1 << 18, 1024)):
    for uplo in "UL":
        d = torch.rand((n, k + 1), dtype=torch.float64, device="cuda") / (2 * k)
        d[:, k if uplo == "U" else 0] = 2.0
        ts = []
        for r in range(3):
            x = torch.ones(n, dtype=torch.float64, device="cuda")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); bm.tbsv_(uplo, "T", "N", n, k, d, x); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"tbsv {uplo} T n={n} k={k}: {min(ts[1:]):.2f} ms ({1e6*min(ts[1:])/n:.0f} ns/col)")
