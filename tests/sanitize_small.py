"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel, every mode, tiny inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, zref, zstd_b200
ctx = zstd_b200.ZSTD_CCtx()
cases = [b"", b"abcdefg", zref.synthetic(1000, 1), zref.synthetic(70_000, 2), zref.synthetic(300_000, 3, 0.8), bytes(200_000),
         zref.random_bytes(150_000, 4), b"abc" * 30_000]
for src in cases:
    for level in (1, -3, 3):
        got = ctx.compress(src, level)
        assert got == zref.oracle_compress(src, level)
d = zref.golden_input("zdict-16k-synthetic-seed77")
recs = zref.synthetic(1024 * 16, 5)
for i in range(16):
    r = recs[i * 1024:(i + 1) * 1024]
    assert ctx.compress_using_dict(r, d, 1) == zref.oracle_compress_using_dict(r, d, 1)
src = zref.synthetic(1024 * 64, 9)
t = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
cap = 64 * (zstd_b200.ZSTD_compressBound(1024) + 32)
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
total, csz = ctx.compress_frames(out.data_ptr(), cap, t.data_ptr(), [i * 1024 for i in range(64)], [1024] * 64, level=1, dict_bytes=d)
# digested dictionary: single calls and a batch (cached table image, warp-per-block merge / copy kernels)
cd = zstd_b200.ZSTD_CDict(d, 1)
for i in range(4):
    r = recs[i * 1024:(i + 1) * 1024]
    assert ctx.compress_using_cdict(r, cd) == zref.oracle_compress_using_dict(r, d, 1)
total2, csz2 = ctx.compress_frames_using_cdict(out.data_ptr(), cap, t.data_ptr(), [i * 1024 for i in range(64)], [1024] * 64, cd)
assert (total2, csz2) == (total, csz)
cd.close()
# sizes around the parse-segment boundaries, several blocks, host path (waves) and device path
for n in (16383, 16385, 16391, 131071, 131073, 147461, 400_003):
    src = zref.synthetic(n, n % 11, 0.6)
    for level in (1, 3):
        assert ctx.compress(src, level) == zref.oracle_compress(src, level)
print("sanitize workload ok", total)
# device source of EXACTLY srcSize bytes from cudaMalloc (no caching-allocator slack behind it): reads past the end would show
import ctypes
rt = ctypes.CDLL("libcudart.so")
for n in (131072, 131072 * 4 + 8, 1000):
    src = zref.synthetic(n, 3, 0.7)
    p = ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(n)) == 0
    assert rt.cudaMemcpy(p, src, ctypes.c_size_t(n), 1) == 0
    capn = zstd_b200.ZSTD_compressBound(n)
    o = torch.empty(capn, dtype=torch.uint8, device="cuda")
    for level in (1, 3):
        k = ctx.compress_device(o.data_ptr(), capn, p.value, n, level)
        assert bytes(o[:k].cpu().numpy()) == zref.oracle_compress(src, level)
    rt.cudaFree(p)
# checksums on the device, streaming, a frame in parts
c2 = zstd_b200.ZSTD_CCtx(); c2.set_parameter("checksum_flag", 1)
n = c2.compress_device(out.data_ptr(), cap, t.data_ptr(), len(src), level=1)
c2.close()
# decompression: own frames, dictionary frames, device path
dctx = zstd_b200.ZSTD_DCtx()
for src in cases:
    for level in (1, 3):
        assert dctx.decompress(ctx.compress(src, level), len(src)) == src
big = zref.synthetic(600_000, 12, 0.5)
f = ctx.compress(big, 1)
d_in = torch.frombuffer(bytearray(f), dtype=torch.uint8).cuda()
d_o = torch.empty(len(big), dtype=torch.uint8, device="cuda")
assert dctx.decompress_device(d_o.data_ptr(), len(big), d_in.data_ptr(), len(f)) == len(big)
assert bytes(d_o.cpu().numpy()) == big
dctx.close()
print("sanitize workload (round 2 additions) ok")
