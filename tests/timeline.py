"""Print the wave executor's timeline for one host-buffer call and one device-buffer call (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, zref, zstd_b200
G = 1 << 30
src = zref.datagen(G, 50)
L = zstd_b200.lib()
cap = zstd_b200.ZSTD_compressBound(G)
h_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).pin_memory()
h_dst = torch.empty(cap, dtype=torch.uint8).pin_memory()
ctx = zstd_b200.ZSTD_CCtx()
for i in range(4):
    if i == 3:
        os.environ["ZSTDB200_TIMELINE"] = "1"
    t0 = time.perf_counter()
    r = L.ZSTD_compressCCtx(ctx._h, h_dst.data_ptr(), cap, h_src.data_ptr(), G, 1)
    print(f"host call {i}: {1e3*(time.perf_counter()-t0):.2f} ms", flush=True)
os.environ.pop("ZSTDB200_TIMELINE")
d_src = h_src.cuda(); d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
for i in range(4):
    if i == 3:
        os.environ["ZSTDB200_TIMELINE"] = "1"
    r = ctx.compress_device(d_dst.data_ptr(), cap, d_src.data_ptr(), G, level=1)
    print(f"device call {i}: {ctx.stats().total_ms:.2f} ms", flush=True)
