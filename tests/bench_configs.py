"""Device-resident timing of the BASELINE.json parity configs on one GPU (not the bench.py headline):
   python tests/bench_configs.py            -> one line per config: GB/s, size vs reference, round trip."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, zref, zstd_b200


def run(name, src, level, frame_size=None, dict_bytes=None, iters=3, ref_sample=None):
    ctx = zstd_b200.ZSTD_CCtx()
    n = len(src)
    d_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
    if frame_size is None:
        offs, sizes = [0], [n]
    else:
        offs = list(range(0, n, frame_size)); sizes = [min(frame_size, n - o) for o in offs]
    cap = sum(zstd_b200.ZSTD_compressBound(s) + 32 for s in sizes)
    d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    best = None
    for _ in range(iters):
        total, csz = ctx.compress_frames(d_dst.data_ptr(), cap, d_src.data_ptr(), offs, sizes, level=level, dict_bytes=dict_bytes)
        ms = ctx.stats().kernel_ms
        best = ms if best is None else min(best, ms)
    out = bytes(d_dst[:total].cpu().numpy())
    ok = None
    delta = None
    if zref.have_ref():
        if dict_bytes is None:
            ok = zref.ref_decompress(out, n) == src
        else:
            pos = 0; ok = True
            for i in range(0, len(sizes), max(1, len(sizes) // 64)):
                start = sum(csz[:i])
                ok &= zref.ref_decompress_using_dict(out[start:start + csz[i]], dict_bytes, sizes[i]) == src[offs[i]:offs[i] + sizes[i]]
        # reference size on a sample of frames
        k = ref_sample or len(sizes)
        ref_tot = ours_tot = 0
        for i in range(0, len(sizes), max(1, len(sizes) // k)):
            piece = src[offs[i]:offs[i] + sizes[i]]
            ref_tot += len(zref.ref_compress(piece, level) if dict_bytes is None else zref.ref_compress_using_dict(piece, dict_bytes, level))
            ours_tot += csz[i]
        delta = (ours_tot - ref_tot) / ref_tot
    st = ctx.stats()
    print(f"{name}: {n/best/1e6:.1f} GB/s ({best:.2f} ms, {len(sizes)} frames, {st.nbBlocks} blocks, {st.launches} launches), ratio x{n/total:.3f}, "
          f"size vs reference {delta:+.3%}, decodes: {ok}  [cand {st.cand_ms:.2f} parse {st.parse_ms:.2f} lit {st.literals_ms:.2f} seq {st.sequences_ms:.2f} stitch {st.stitch_ms:.2f}]", flush=True)
    ctx.close()


def main():
    G = 1 << 30
    which = sys.argv[1:] or ["c2", "c3", "c4", "c5"]
    if "c2" in which:
        run("C2 datagen -g1GB -P50, level 1, one frame", zref.datagen(G, 50), 1)
    if "c3" in which:
        run("C3 datagen -g1GB -P30 (1/8 of 8GB), --fast=3, 64 MiB frames", zref.datagen(G, 30), -3, frame_size=64 << 20)
    if "c4" in which:
        run("C4 datagen -g1GB -P90 (half of 2GB), level 3 dfast, one frame", zref.datagen(G, 90), 3)
    if "c5" in which:
        R = zref.ref()
        R.ZDICT_trainFromBuffer.restype = ctypes.c_size_t
        R.ZDICT_trainFromBuffer.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint]
        rec, nrec = 1024, 131072
        src = zref.datagen(rec * nrec, 50)
        sizes = (ctypes.c_size_t * 20000)(*([rec] * 20000))
        dbuf = ctypes.create_string_buffer(16 << 10)
        dn = R.ZDICT_trainFromBuffer(dbuf, 16 << 10, src, sizes, 20000)
        run("C5 131072 x 1 KiB records (P50) + 16 KiB ZDICT dictionary, level 1", src, 1, frame_size=rec, dict_bytes=dbuf.raw[:dn], ref_sample=512)
    if "c5full" in which:                                   # BASELINE config 5 at its full size: 1 Mi records
        R = zref.ref()
        rec, nrec = 1024, 1 << 20
        src = zref.datagen(rec * nrec, 50)
        d = zref.train_dict(src, rec, 20000, 16 << 10)
        run("C5 full: 1048576 x 1 KiB records (P50) + 16 KiB ZDICT dictionary, level 1", src, 1, frame_size=rec, dict_bytes=d, ref_sample=512)


if __name__ == "__main__":
    main()
