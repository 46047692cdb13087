"""Experiment driver (CPU only): oracle size vs reference size over data types x levels, with the oracle's knobs."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import zref

class Tun(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint) for n in ("tableN", "tableNLong", "tableFmt", "insStep", "primeBytes", "chunkBlocks", "batch", "spare")]

def tun():
    return Tun.in_dll(zref.oracle(), "zbo_tun")

_cache = {}
def inputs(mib=8):
    if mib in _cache: return _cache[mib]
    n = mib << 20
    d = {f"P{p}": zref.datagen(n, p) for p in (30, 50, 90)}
    d["syn60"] = zref.synthetic(n, 1, 0.6)
    for g in ("large-literal-and-match-lengths", "http", "PR-3517-block-splitter-corruption-test"):
        d[g[:8]] = zref.golden_input(g)
    _cache[mib] = d
    return d

_ref = {}
def run(levels=(1, 3, -3), mib=8, check=False, names=None, **kw):
    t = tun()
    for f, _ in Tun._fields_: setattr(t, f, 0)
    for k, v in kw.items(): setattr(t, k, v)
    row = []
    for name, data in inputs(mib).items():
        if names and name not in names: continue
        for lv in levels:
            key = (name, lv, mib)
            if key not in _ref: _ref[key] = len(zref.ref_compress(data, lv))
            c = zref.oracle_compress(data, lv)
            if check: assert zref.ref_decompress(c, len(data)) == data, (name, lv)
            row.append(f"{name}@{lv}:{100.0 * (len(c) - _ref[key]) / _ref[key]:+.2f}")
    print(kw, " ".join(row), flush=True)

if __name__ == "__main__":
    run(check=True)

def small(levels=(1, 3, -3), sizes=(1024, 4096, 16384, 65536, 262144), nframes=64, **kw):
    """many small frames cut from datagen streams: total oracle size vs total reference size"""
    t = tun()
    for f, _ in Tun._fields_: setattr(t, f, 0)
    for k, v in kw.items(): setattr(t, k, v)
    row = []
    for p in (30, 50, 90):
        big = zref.datagen(16 << 20, p)
        for sz in sizes:
            nf = min(nframes, len(big) // sz)
            for lv in levels:
                key = ("small", p, sz, lv, nf)
                if key not in _ref: _ref[key] = sum(len(zref.ref_compress(big[i * sz:(i + 1) * sz], lv)) for i in range(nf))
                o = sum(len(zref.oracle_compress(big[i * sz:(i + 1) * sz], lv)) for i in range(nf))
                row.append(f"P{p}/{sz >> 10}K@{lv}:{100.0 * (o - _ref[key]) / _ref[key]:+.2f}")
    print(kw, " ".join(row), flush=True)
