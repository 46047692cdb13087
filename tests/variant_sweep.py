"""Compare development builds of the library (tools/build_variant.sh) on one GPU:
   python tests/variant_sweep.py c2 base e12 e16
One subprocess per variant (ZSTDB200_LIB selects the .so): serial-mode per-kernel times, wave-mode total,
and a CRC of the output (all variants must produce the same bytes)."""
import os, sys, subprocess, statistics, zlib
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def child(which, iters):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import torch, zref, zstd_b200
    G = 1 << 30
    src, level = (zref.datagen(G, 50), 1) if which == "c2" else (zref.datagen(G, 90), 3)
    d_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
    cap = zstd_b200.ZSTD_compressBound(G) + 32
    d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    out = []
    for serial in (1, 0):
        os.environ["ZSTDB200_SERIAL"] = str(serial)
        ctx = zstd_b200.ZSTD_CCtx()
        rows = []
        for i in range(iters + 2):
            total = ctx.compress_device(d_dst.data_ptr(), cap, d_src.data_ptr(), G, level=level)
            st = ctx.stats()
            if i >= 2:
                rows.append((st.total_ms, st.cand_ms, st.parse_ms, st.literals_ms, st.sequences_ms, st.stitch_ms))
        best = min(rows)
        crc = zlib.crc32(d_dst[:total].cpu().numpy().tobytes())
        if serial:
            out.append(f"serial {best[0]:.2f} ms [cand {best[1]:.2f} parse {best[2]:.2f} lit {best[3]:.2f} seq {best[4]:.2f} stitch {best[5]:.2f}]")
        else:
            out.append(f"waves {best[0]:.2f} ms = {G/best[0]/1e6:.1f} GB/s")
        ctx.close()
    print(f"{os.path.basename(os.environ.get('ZSTDB200_LIB', 'default')):28s} {'  '.join(out)}  size {total} crc {crc:08x}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
    else:
        which = sys.argv[1]
        for name in sys.argv[2:]:
            env = dict(os.environ)
            if name != "default":
                env["ZSTDB200_LIB"] = os.path.join(ROOT, "zstd_b200", "variants", f"libzstd_b200_{name}.so")
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", which, "4"], env=env, check=False, timeout=300)
