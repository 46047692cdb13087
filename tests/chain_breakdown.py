"""Per-kernel CUDA-event times of short device-resident calls (one wave, one stream): where the dependency chain of
a single call goes.  python tests/chain_breakdown.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, zref, zstd_b200
src = zref.datagen(64 << 20, 50)
d_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
cap = zstd_b200.ZSTD_compressBound(len(src))
d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
ctx = zstd_b200.ZSTD_CCtx()
for level in (1, 3):
    for kib in (128, 1024, 16 << 10, 64 << 10):
        n = kib << 10
        best = None
        for i in range(6):
            ctx.compress_device(d_dst.data_ptr(), cap, d_src.data_ptr(), n, level=level)
            st = ctx.stats()
            row = (st.kernel_ms, st.cand_ms, st.parse_ms, st.literals_ms, st.sequences_ms, st.stitch_ms)
            best = row if best is None or row[0] < best[0] else best
        print(f"level {level} {kib:6d} KiB: total {best[0]:.3f} ms = cand {best[1]:.3f} + parse/merge {best[2]:.3f} + literals {best[3]:.3f} + sequences {best[4]:.3f} + stitch {best[5]:.3f}", flush=True)
