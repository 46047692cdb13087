"""bring-up: dictionary calls at a doubleFast level, usingDict vs CDict vs oracle"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import zref, zstd_b200
ctx = zstd_b200.ZSTD_CCtx()
for dname in ("zdict-16k-synthetic-seed77", "raw-32k"):
    d = zref.synthetic(32 << 10, 123, 0.5) if dname == "raw-32k" else zref.golden_input(dname)
    for level in (3, 1):
        cd = zstd_b200.ZSTD_CDict(d, level)
        for n in (1024, 5000, 20000, 70000, 131072, 140000, 300000):
            src = zref.synthetic(n, 8)
            want = zref.oracle_compress_using_dict(src, d, level)
            a = ctx.compress_using_dict(src, d, level)
            b = ctx.compress_using_cdict(src, cd)
            def first_diff(x, y):
                m = min(len(x), len(y))
                return next((i for i in range(m) if x[i] != y[i]), m)
            print(dname, level, n, "usingDict", "ok" if a == want else f"DIFF@{first_diff(a, want)} len {len(a)} vs {len(want)}",
                  "| cdict", "ok" if b == want else f"DIFF@{first_diff(b, want)} len {len(b)} vs {len(want)}", flush=True)
        cd.close()
