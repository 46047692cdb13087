"""Latency-side view of development builds (tools/build_variant.sh): end-to-end time of 1 GiB from pinned host
buffers, and device-resident times of short inputs (where a call is one dependency chain, not a throughput problem).
   python tests/latency_sweep.py default nohist"""
import os, sys, subprocess, time
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def child():
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import torch, zref, zstd_b200
    G = 1 << 30
    src = zref.datagen(G, 50)
    L = zstd_b200.lib()
    cap = zstd_b200.ZSTD_compressBound(G)
    h_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).pin_memory()
    h_dst = torch.empty(cap, dtype=torch.uint8).pin_memory()
    ctx = zstd_b200.ZSTD_CCtx()
    ts = []
    for i in range(8):
        t0 = time.perf_counter()
        r = L.ZSTD_compressCCtx(ctx._h, h_dst.data_ptr(), cap, h_src.data_ptr(), G, 1)
        ts.append(time.perf_counter() - t0)
    out = [f"e2e 1 GiB {1e3*min(ts[3:]):.2f} ms ({G/min(ts[3:])/1e9:.1f} GB/s)"]
    ctx.close()
    d_src = h_src.cuda(); d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    ctx = zstd_b200.ZSTD_CCtx()
    for mib in (1, 4, 16, 64, 256):
        n = mib << 20
        best = 1e9
        for i in range(6):
            ctx.compress_device(d_dst.data_ptr(), cap, d_src.data_ptr(), n, level=1)
            best = min(best, ctx.stats().kernel_ms)
        out.append(f"{mib} MiB {best:.2f} ms ({n/best/1e6:.1f} GB/s)")
    print(f"{os.path.basename(os.environ.get('ZSTDB200_LIB', 'default')):26s} " + " | ".join(out), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child()
    else:
        for name in sys.argv[1:]:
            env = dict(os.environ)
            if name != "default":
                env["ZSTDB200_LIB"] = os.path.join(ROOT, "zstd_b200", "variants", f"libzstd_b200_{name}.so")
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, check=False, timeout=300)
