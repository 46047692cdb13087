"""Device-resident decompression timing on one GPU: frames written by this library and by the reference encoder.
   python tests/bench_decode.py [size_MiB]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, zref, zstd_b200

if len(sys.argv) > 1 and sys.argv[1] == "--json":              # bench.py's round-trip leg: one workload, one JSON line
    import json
    size, p, level, dev = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    torch.cuda.set_device(dev)
    src = zref.datagen(size, p, 0) if zref.have_datagen() else zref.synthetic(size, 0, p / 100.0)
    c, d = zstd_b200.ZSTD_CCtx(device=dev), zstd_b200.ZSTD_DCtx(device=dev)
    d_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
    cap = zstd_b200.ZSTD_compressBound(size)
    d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = c.compress_device(d_c.data_ptr(), cap, d_src.data_ptr(), size, level)
    d_out = torch.empty(size, dtype=torch.uint8, device="cuda")
    best, m = None, 0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m = d.decompress_device(d_out.data_ptr(), size, d_c.data_ptr(), n)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        best = t if best is None else min(best, t)
    st = d.stats()
    print(json.dumps({"gpu_roundtrip_ok": bool(m == size and torch.equal(d_out, d_src)), "value": round(size / best / 1e6, 3),
                      "unit": "GB/s (output bytes, device buffers, ZSTDB200_decompressDevice)", "ms": round(best, 3), "compressed_bytes": int(n),
                      "kernel_ms": {"literals": round(st.literals_ms, 3), "sequences": round(st.sequences_ms, 3), "place": round(st.place_ms, 3), "matches": round(st.execute_ms, 3)}}))
    sys.exit(0)

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
c, d = zstd_b200.ZSTD_CCtx(), zstd_b200.ZSTD_DCtx()
for p, level in ((50, 1), (90, 3), (30, -3)):
    src = zref.datagen(mib << 20, p)
    d_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
    cap = zstd_b200.ZSTD_compressBound(len(src))
    d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = c.compress_device(d_c.data_ptr(), cap, d_src.data_ptr(), len(src), level)
    d_out = torch.empty(len(src), dtype=torch.uint8, device="cuda")
    for what in ("ours", "reference"):
        if what == "reference":
            if not zref.have_ref() or mib > 256:
                continue
            frame = zref.ref_compress(src, level)
            d_c = torch.frombuffer(bytearray(frame), dtype=torch.uint8).cuda(); n = len(frame)
        best = None
        for _ in range(3):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            m = d.decompress_device(d_out.data_ptr(), len(src), d_c.data_ptr(), n)
            ev1.record(); torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            best = ms if best is None else min(best, ms)
        st = d.stats()
        ok = m == len(src) and torch.equal(d_out, d_src)
        print(f"P{p} level {level} {what:9s}: {len(src) >> 20} MiB <- {n} B  {best:.2f} ms = {len(src) / best / 1e6:.1f} GB/s  "
              f"[kernels {st.kernel_ms:.2f}: literals {st.literals_ms:.2f} sequences {st.sequences_ms:.2f} place {st.place_ms:.2f} matches {st.execute_ms:.2f}; {st.nbBlocks} blocks]  ok {ok}", flush=True)

# many frames in one call: the match stage runs one CTA per frame, so frames decode side by side
if mib >= 256:
    for p, level, fs, what in ((30, -3, 64 << 20, "64 MiB frames (config 3's shape)"), (50, 1, 64 << 20, "64 MiB frames"), (50, 1, 1 << 20, "1 MiB frames"), (50, 1, 1024, "1 KiB records")):
        n = (mib << 20) if fs >= (1 << 20) else (128 << 20)
        src = zref.datagen(n, p)
        d_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
        offs = list(range(0, n, fs)); sizes = [fs] * len(offs)
        cap = sum(zstd_b200.ZSTD_compressBound(x) + 32 for x in sizes)
        d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
        total, csz = c.compress_frames(d_c.data_ptr(), cap, d_src.data_ptr(), offs, sizes, level=level)
        d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
        best = None
        for _ in range(3):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            m = d.decompress_device(d_out.data_ptr(), n, d_c.data_ptr(), total)
            ev1.record(); torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            best = ms if best is None else min(best, ms)
        st = d.stats()
        ok = m == n and torch.equal(d_out, d_src)
        print(f"P{p} level {level}, {len(offs)} x {what}: {n >> 20} MiB <- {total} B  {best:.2f} ms = {n / best / 1e6:.1f} GB/s  "
              f"[kernels {st.kernel_ms:.2f}: literals {st.literals_ms:.2f} sequences {st.sequences_ms:.2f} place {st.place_ms:.2f} matches {st.execute_ms:.2f}]  ok {ok}", flush=True)
