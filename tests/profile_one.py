"""Run one device-resident compression (for ncu).  python tests/profile_one.py <size_MiB> <P> <level> [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, zref, zstd_b200
mib, p, level = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 1
src = zref.datagen(mib << 20, p) if zref.have_datagen() else zref.synthetic(mib << 20, 1, p / 100)
t = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
cap = zstd_b200.ZSTD_compressBound(len(src))
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
ctx = zstd_b200.ZSTD_CCtx()
for i in range(iters):
    n = ctx.compress_device(out.data_ptr(), cap, t.data_ptr(), len(src), level)
    st = ctx.stats()
    print(f"iter {i}: {len(src)} -> {n}  kernel {st.kernel_ms:.3f} ms (match {st.match_ms:.3f}) = {len(src)/st.kernel_ms/1e6:.2f} GB/s, launches {st.launches}")
