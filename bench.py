#!/usr/bin/env python
"""bench.py — compression throughput of the zstd hot path on B200 (BASELINE.json metric).

N=1 workload = BASELINE.json configs[1]: `datagen -g1GB -P50`, level 1, one frame, 128 KiB blocks.
N>1: every rank compresses its own 1 GiB shard as independent frames (weak scaling, no data-path
collective); the compressed buffers are gathered to rank 0 over NCCL inside the timed region.

  python bench.py --gpus N --steps K --warmup W            # our CUDA path
  python bench.py --impl reference --steps K --warmup W    # reference libzstd on the host cores
"""
import argparse
import ctypes
import json
import os

# the host path keeps ~10 streams busy (8 wave streams + upload + download): give every one its own hardware
# queue, else a download can sit behind another wave's kernels (profiles/r1_e2e_timeline.md).  Must be set
# before the CUDA context exists; INTEGRATION.md tells embedders to do the same.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "compress GB/s (input) at level 1"
GiB = 1 << 30


def load_input(size, p=50, seed=0):
    """(bytes, description).  Reference datagen when its binary travelled with the repo."""
    import zref
    if zref.have_datagen():
        return zref.datagen(size, p, seed), f"synthetic: reference tests/datagen -g{size} -P{p} -s{seed}"
    return zref.synthetic(size, seed, p / 100.0), f"synthetic: zbo_synthetic(n={size}, match_prob={p/100}) (datagen binary absent)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.index, self.proc, self.windows = [], index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark(self):
        """Start (or restart) a timed window: only samples taken inside windows are reported."""
        self.windows.append([time.perf_counter(), None])

    def unmark(self):
        self.windows[-1][1] = time.perf_counter()

    def stop(self):
        if self.proc:
            self.proc.terminate()
        inside = [s for t, s in self.samples if any(a <= t <= (b or 1e30) for a, b in self.windows)]
        sm = sorted(int(s[0]) for s in inside if s and s[0].isdigit())
        mx = max([int(s[1]) for s in inside if len(s) > 1 and s[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in inside for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm),
                "sampled": "nvidia-smi -lms 20 during the timed device-resident and end-to-end loops"}


def ref_lib():
    import zref
    R = zref.ref()
    R.ZSTD_createCCtx.restype = ctypes.c_void_p
    R.ZSTD_CCtx_setParameter.restype = ctypes.c_size_t
    R.ZSTD_CCtx_setParameter.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    R.ZSTD_compress2.restype = ctypes.c_size_t
    R.ZSTD_compress2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    return R


class RefMT:
    """One ZSTD_CCtx with nbWorkers = threads (ZSTDMT, zstdmt_compress.c) and one destination buffer, both kept
    across calls: the unmodified reference from oracle/_ref/libzstd_ref.so (-O3 -DZSTD_MULTITHREAD)."""

    def __init__(self, size, level, threads):
        self.R = ref_lib()
        self.cctx = self.R.ZSTD_createCCtx()
        self.R.ZSTD_CCtx_setParameter(self.cctx, 100, level)            # ZSTD_c_compressionLevel
        if threads > 1:
            self.R.ZSTD_CCtx_setParameter(self.cctx, 400, threads)      # ZSTD_c_nbWorkers
        self.cap = self.R.ZSTD_compressBound(size)
        self.dst = ctypes.create_string_buffer(self.cap)

    def run(self, src):
        """(seconds, compressed size) of one ZSTD_compress2 over `src`."""
        t0 = time.perf_counter()
        csize = self.R.ZSTD_compress2(self.cctx, self.dst, self.cap, src, len(src))
        return time.perf_counter() - t0, csize

    def close(self):
        self.R.ZSTD_freeCCtx(self.cctx)


def cpu_reference_time(src, level, threads, repeats=1):
    """Seconds for one ZSTD_compress2(nbWorkers=threads) of `src`; the context is warmed by one untimed call."""
    m = RefMT(len(src), level, threads)
    m.run(src)
    best, csize = None, 0
    for _ in range(repeats):
        dt, csize = m.run(src)
        best = dt if best is None else min(best, dt)
    m.close()
    return best, csize


class RefSliced:
    """`threads` host threads, each with a private ZSTD_CCtx and destination buffer (kept across calls),
    compressing equal contiguous slices of the input into independent frames (contrib/pzstd's decomposition)."""

    def __init__(self, buf, level, threads):
        R = self.R = ref_lib()
        R.ZSTD_compressCCtx.restype = ctypes.c_size_t
        R.ZSTD_compressCCtx.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        self.buf, self.level, self.threads = buf, level, threads
        self.n = len(buf)
        self.per = (self.n + threads - 1) // threads
        self.addr = ctypes.addressof(buf)
        self.ctxs = [R.ZSTD_createCCtx() for _ in range(threads)]
        self.caps = [R.ZSTD_compressBound(min(self.per, self.n - i * self.per)) if i * self.per < self.n else 0 for i in range(threads)]
        self.dsts = [ctypes.create_string_buffer(max(c, 1)) for c in self.caps]

    def run(self):
        out = [0] * self.threads

        def work(i):
            lo = i * self.per
            if lo < self.n:
                out[i] = self.R.ZSTD_compressCCtx(self.ctxs[i], self.dsts[i], self.caps[i], self.addr + lo, min(self.per, self.n - lo), self.level)

        ths = [threading.Thread(target=work, args=(i,)) for i in range(self.threads)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0, sum(out)

    def close(self):
        for c in self.ctxs:
            self.R.ZSTD_freeCCtx(c)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation with all host threads (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import zref
    if not zref.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libzstd_ref.so missing (reference not built on this box)"}))
        return
    cores = os.cpu_count() or 1
    size = args.size
    src, desc = load_input(size)
    hbuf = (ctypes.c_char * size).from_buffer_copy(src)
    mt, sl = RefMT(size, args.level, cores), RefSliced(hbuf, args.level, cores)
    # contexts, thread pools and destination buffers live across steps (as our own arm's do); at least two untimed
    # passes whatever --warmup says: the first pass of a 128-thread run pays first-touch page faults of ~1.3 GiB
    for _ in range(max(2, args.warmup)):
        mt.run(src)
        sl.run()
    # two stock ways to use every host thread: one frame through ZSTDMT (nbWorkers), or N independent frames
    dt_mt = dt_sl = 0.0
    csize = csl = 0
    for _ in range(args.steps):
        dt, csize = mt.run(src)
        dt_mt += dt
    for _ in range(args.steps):
        dt, csl = sl.run()
        dt_sl += dt
    mt.close(); sl.close()
    mode = "ZSTD_compress2 nbWorkers" if dt_mt <= dt_sl else "independent frames, one ZSTD_compressCCtx thread per slice"
    dt = min(dt_mt, dt_sl)
    if dt_sl < dt_mt:
        csize = csl
    v = size * args.steps / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": desc,
            "config": {"workload": f"datagen -g{size} -P50, level {args.level}, {cores} host threads, best of: ZSTD_compress2 nbWorkers ({size*args.steps/dt_mt/1e9:.2f} GB/s) / independent slices ({size*args.steps/dt_sl/1e9:.2f} GB/s); used: {mode}", "compressed_bytes": csize},
            "cpu_baseline": {"value": round(v, 4), "unit": "GB/s", "cores": cores, "kind": "reference",
                             "sample": f"whole {size}-byte buffer per step; {mode}"},
            "e2e": {"value": round(v, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import zstd_b200
    import zref

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (zstd_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    size = args.size
    src, desc = load_input(size, seed=rank)                 # weak scaling: every rank owns one shard
    ctx = zstd_b200.ZSTD_CCtx(device=local)
    cap = zstd_b200.ZSTD_compressBound(size)
    d_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
    d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    h_src = torch.frombuffer(bytearray(src), dtype=torch.uint8).pin_memory()
    h_dst = torch.empty(cap, dtype=torch.uint8).pin_memory()
    L = zstd_b200.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        n = ctx.compress_device(d_dst.data_ptr(), cap, d_src.data_ptr(), size, args.level)
        if world > 1:
            from zstd_b200.sharding import gather_compressed
            gather_compressed(d_dst[:n], dst=0)
        return n

    def step_e2e():
        r = L.ZSTD_compressCCtx(ctx._h, h_dst.data_ptr(), cap, h_src.data_ptr(), size, args.level)
        assert not L.ZSTD_isError(r), L.ZSTD_getErrorName(r)
        return r

    # ---- device-resident throughput (`value`) ----
    csize = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                     # nvidia-smi needs a moment to start: launch it before the warm-up
    for _ in range(args.warmup):
        csize = step_device()
    barrier()
    sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = []
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        csize = step_device()
        stats.append(ctx.stats())
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    sampler.unmark()
    ms = ev0.elapsed_time(ev1)
    launches = sum(s.launches for s in stats)
    # per-kernel CUDA-event times come from a serial-mode context (one wave, one stream): in the default
    # mode waves on several streams overlap and a kernel's start->end no longer measures that kernel alone
    os.environ["ZSTDB200_SERIAL"] = "1"
    sctx = zstd_b200.ZSTD_CCtx(device=local)
    del os.environ["ZSTDB200_SERIAL"]
    sstats = []
    for i in range(2 + 3):
        sctx.compress_device(d_dst.data_ptr(), cap, d_src.data_ptr(), size, args.level)
        if i >= 2:
            sstats.append(sctx.stats())
    kern = {k: sum(getattr(s, k) for s in sstats) / len(sstats) for k in ("kernel_ms", "cand_ms", "parse_ms", "literals_ms", "sequences_ms", "stitch_ms")}
    sctx.close()
    torch.cuda.synchronize()

    # ---- end to end through the reference-facing C ABI with pinned HOST buffers (a context of its own, as an
    # application that only ever passes host pointers would have) ----
    got_dev = d_dst[:csize].clone()
    ctx.close()
    ctx = zstd_b200.ZSTD_CCtx(device=local)
    for _ in range(max(3, args.warmup)):
        step_e2e()
    barrier()
    sampler.mark()
    te0 = time.perf_counter()
    ce = 0
    for _ in range(args.steps):
        ce = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - te0
    sampler.unmark()
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # parity spot-check of what was timed (outside the timed region)
    got = bytes(got_dev.cpu().numpy())
    assert bytes(h_dst[:ce].numpy()) == got, "host-path and device-path frames differ"
    ok_rt = zref.ref_decompress(got, size) == src if zref.have_ref() else None
    hbm, peak_src = peaks()
    value = size * world * args.steps / (ms / 1e3) / 1e9
    e2e = size * world * args.steps / e2e_s / 1e9
    dom = max(("cand_ms", "parse_ms", "literals_ms", "sequences_ms", "stitch_ms"), key=lambda k: kern[k])
    achieved = (size + csize) / (kern[dom] / 1e3) / 1e9
    # DRAM bytes of that kernel per launch: from the committed ncu capture of this workload (not measured live)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tj = json.load(f)
        k = tj["kernels"][dom.replace("_ms", "")]
        if size == GiB and args.level == 1:
            traffic, traffic_src = k["dram_read_bytes"] + k["dram_write_bytes"], "profiles/r1_traffic.json (ncu, same workload)"
    except Exception:
        pass
    # CPU baseline on this box: reference libzstd, bounded sample
    cpu = None
    if zref.have_ref() and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample = src[: min(size, 256 << 20)]
        t1, c1 = cpu_reference_time(sample, args.level, 1)
        tn, cn = cpu_reference_time(src, args.level, cores)        # warmed inside
        hbuf = (ctypes.c_char * size).from_buffer_copy(src)
        sl = RefSliced(hbuf, args.level, cores)
        sl.run(); sl.run()
        ts, _ = sl.run()
        sl.close()
        _, cref = cpu_reference_time(src, args.level, 1) if size <= (256 << 20) else (0, None)
        cpu = {"value": round(size / min(tn, ts) / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "reference",
               "sample": f"whole {size}-byte input, {cores} threads: ZSTD_compress2 nbWorkers {size/tn/1e9:.2f} GB/s, independent slices {size/ts/1e9:.2f} GB/s; 1 thread on the first {len(sample)} bytes: {len(sample)/t1/1e9:.3f} GB/s",
               "ref_compressed_bytes": cn}
    line = {"metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": desc,
            "config": {"workload": f"datagen -g{size} -P50 per GPU, level {args.level}, one frame per GPU, 128 KiB blocks", "l2": "input 1 GiB per step > 126 MB L2 (no reuse between steps)",
                       "compressed_bytes": csize, "roundtrip_ok": ok_rt,
                       "size_delta_vs_ref": (round((csize - cpu["ref_compressed_bytes"]) / cpu["ref_compressed_bytes"], 5) if cpu else None)},
            "kernel_ms": dict({k: round(v, 3) for k, v in kern.items()}, mode="serial (ZSTDB200_SERIAL=1): one wave on one stream, CUDA events around each kernel"),
            "roofline": {"bound": "hbm", "kernel": dom.replace("_ms", ""), "achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s",
                         "frac": round(achieved / hbm, 4), "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes": size + csize, "read_only_frac": round(size / (kern[dom] / 1e3) / 1e9 / hbm, 4)},
            "cpu_baseline": cpu,
            "e2e": {"value": round(e2e, 3), "unit": "GB/s", "h2d_bytes_per_step": size, "d2h_bytes_per_step": int(ce), "api": "ZSTD_compressCCtx(host pinned src/dst)"},
            "gpu_launches": launches, "clocks": clocks, "wall_s": round(wall, 3)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--size", type=int, default=GiB)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
