"""Host-buffer (end-to-end) wave-executor sweep on one GPU (development tool):
   python tests/e2e_sweep.py [iters]
ZSTD_compressCCtx on pinned host buffers, 1 GiB datagen -P50 level 1, for several (blocks per wave, waves in
flight) settings; plus a plain pinned H2D / D2H copy of the same bytes as the PCIe floor."""
import os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch, zref, zstd_b200


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    G = 1 << 30
    src = zref.datagen(G, 50)
    L = zstd_b200.lib()
    cap = zstd_b200.ZSTD_compressBound(G)
    h_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).pin_memory()
    h_dst = torch.empty(cap, dtype=torch.uint8).pin_memory()
    d = torch.empty(G, dtype=torch.uint8, device="cuda")
    # PCIe floor: upload 1 GiB while downloading the compressed size
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for k in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d.copy_(h_src, non_blocking=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        with torch.cuda.stream(s1):
            d.copy_(h_src, non_blocking=True)
        with torch.cuda.stream(s2):
            h_dst[:341 << 20].copy_(d[:341 << 20], non_blocking=True)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"PCIe floor: H2D 1 GiB {1e3*(t1-t0):.2f} ms ({G/(t1-t0)/1e9:.1f} GB/s); H2D 1 GiB + D2H 341 MiB concurrently {1e3*(t2-t1):.2f} ms", flush=True)
    base = None
    for wbk, slots in ((512, 8), (512, 6), (384, 8), (256, 8), (256, 12), (128, 12), (768, 6)):
        os.environ["ZSTDB200_HOST_WAVE_BLOCKS"] = str(wbk); os.environ["ZSTDB200_WAVE_SLOTS"] = str(slots)
        ctx = zstd_b200.ZSTD_CCtx()
        ts = []
        for i in range(iters + 2):
            t0 = time.perf_counter()
            r = L.ZSTD_compressCCtx(ctx._h, h_dst.data_ptr(), cap, h_src.data_ptr(), G, 1)
            t1 = time.perf_counter()
            assert not L.ZSTD_isError(r)
            if i >= 2:
                ts.append(t1 - t0)
        crc = zlib.crc32(h_dst[:r].numpy().tobytes())
        base = base or (r, crc)
        print(f"wave {wbk:4d} slots {slots:2d}: best {1e3*min(ts):.2f} ms  median {1e3*sorted(ts)[len(ts)//2]:.2f} ms  {G/min(ts)/1e9:.1f} GB/s  same bytes {(r, crc) == base}", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
