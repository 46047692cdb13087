"""Ad-hoc GPU bring-up script (not a pytest): GPU frame vs oracle frame on a list of cases,
with a readable diff.  Usage on the GPU box: python tests/gpu_debug.py [quick|full]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import zref
import zstd_b200


def parse_blocks(frame):
    """[(type, lastBlock, payloadOffset, payloadSize)] of a single frame"""
    fhd = frame[4]
    ss = (fhd >> 5) & 1
    fcs = fhd >> 6
    did = fhd & 3
    pos = 5 + (0 if ss else 1) + [0, 1, 2, 4][did] + ([1, 2, 4, 8][fcs] if (fcs or ss) else 0)
    out = []
    while pos + 3 <= len(frame):
        h = frame[pos] | (frame[pos + 1] << 8) | (frame[pos + 2] << 16)
        last, typ, sz = h & 1, (h >> 1) & 3, h >> 3
        psz = 1 if typ == 1 else sz
        out.append((typ, last, pos + 3, psz, sz))
        pos += 3 + psz
        if last:
            break
    return out


def check(name, src, level, ctx):
    t0 = time.time()
    try:
        got = ctx.compress(src, level)
    except Exception as e:
        print(f"[FAIL] {name} L{level}: GPU raised {e}")
        return False
    st = ctx.stats()
    want = zref.oracle_compress(src, level)
    ok = got == want
    rt = None
    if zref.have_ref():
        try:
            rt = zref.ref_decompress(got, len(src)) == src
        except Exception as e:
            rt = f"decode error: {e}"
    msg = f"[{'ok' if ok and rt in (True, None) else 'FAIL'}] {name} L{level}: n={len(src)} gpu={len(got)} oracle={len(want)} roundtrip={rt} kernel={st.kernel_ms:.2f}ms match={st.match_ms:.2f}ms"
    print(msg)
    if not ok:
        n = min(len(got), len(want))
        d = next((i for i in range(n) if got[i] != want[i]), n)
        print(f"      first diff at byte {d}")
        try:
            bg, bw = parse_blocks(got), parse_blocks(want)
            print(f"      blocks gpu={len(bg)} oracle={len(bw)}")
            for i, (a, b) in enumerate(zip(bg, bw)):
                if a != b or got[a[2]:a[2] + a[3]] != want[b[2]:b[2] + b[3]]:
                    pa, pb = got[a[2]:a[2] + a[3]], want[b[2]:b[2] + b[3]]
                    dd = next((k for k in range(min(len(pa), len(pb))) if pa[k] != pb[k]), min(len(pa), len(pb)))
                    print(f"      block {i}: gpu(type={a[0]},size={a[3]}) oracle(type={b[0]},size={b[3]}) first payload diff at {dd}")
                    print(f"        gpu   : {pa[max(0,dd-8):dd+24].hex()}")
                    print(f"        oracle: {pb[max(0,dd-8):dd+24].hex()}")
                    break
        except Exception as e:
            print("      (block parse failed:", e, ")")
    return ok


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    ctx = zstd_b200.ZSTD_CCtx()
    cases = []
    cases.append(("empty", b""))
    cases.append(("one", b"a"))
    cases.append(("six", b"abcdef"))
    cases.append(("seven", b"abcdefg"))
    cases.append(("tiny-rep", b"abcabcabcabcabcabcabcabcabcabc" * 3))
    cases.append(("zeros-300", bytes(300)))
    cases.append(("zeros-1M", bytes(1 << 20)))
    cases.append(("period3", b"abc" * 50000))
    cases.append(("period40", bytes(range(40)) * 9000))
    cases.append(("rand-100k", zref.random_bytes(100000, 1)))
    cases.append(("rand-300k", zref.random_bytes(300000, 2)))
    for n in (100, 1000, 5000, 20000, 70000, 131072, 131073, 200000, 262144, 400000):
        cases.append((f"syn-{n}", zref.synthetic(n, seed=n)))
    cases.append(("syn-4M-p30", zref.synthetic(4 << 20, seed=5, match_prob=0.3)))
    cases.append(("syn-4M-p90", zref.synthetic(4 << 20, seed=6, match_prob=0.9)))
    if zref.have_datagen():
        cases.append(("datagen-16M-P50", zref.datagen(16 << 20, 50)))
    nbad = 0
    for name, src in cases:
        for level in (1, -3, 3):
            nbad += not check(name, src, level, ctx)
    print("failures:", nbad)
    if mode == "full" and zref.have_datagen():
        import torch
        src = zref.datagen(1 << 30, 50)
        t = torch.frombuffer(bytearray(src), dtype=torch.uint8).cuda()
        cap = zstd_b200.ZSTD_compressBound(len(src))
        out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        for it in range(3):
            n = ctx.compress_device(out.data_ptr(), cap, t.data_ptr(), len(src), 1)
            st = ctx.stats()
            print(f"1GiB device: csize={n} kernel={st.kernel_ms:.2f}ms match={st.match_ms:.2f}ms -> {len(src)/st.kernel_ms/1e6:.2f} GB/s")
        got = bytes(out[:n].cpu().numpy())
        print("1GiB roundtrip:", zref.ref_decompress(got, len(src)) == src)
    return nbad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
