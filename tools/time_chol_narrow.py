import sys
sys.path.insert(0, ".")
import bandedmatrices_b200 as bm
for flag in (0, 1):
    bm.handle(0).tune("pb_nodiag", flag)
    print("pb_nodiag =", flag)
    for kd in (2, 4, 7, 8):
        sys.argv = ["x", "524288", str(kd), "U", "1"]
        try:
            exec(open("tools/time_chol.py").read().split("B = torch.ones")[0])
        except SystemExit:
            pass
