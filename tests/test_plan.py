"""Host logic of the product without a GPU: the planner of libzstd_b200.so (parameter derivation per level / size /
dictionary, block partition, block flags, history and insertion phases) against the oracle's plan, which is itself
pinned to the reference's ZSTD_getCParams (tests/test_oracle_frames.py::test_cparams_match_reference)."""
import ctypes

import pytest

import zref
import zstd_b200

SIZES = [0, 1, 6, 7, 100, 1024, 4096, 16 << 10, (16 << 10) + 1, 65535, 128 << 10, (128 << 10) + 1, 200_000, 256 << 10, (256 << 10) + 1,
         1 << 20, (8 << 20) + 12345, 64 << 20, (1 << 30) + 77]


class OPlan(ctypes.Structure):
    _fields_ = [("mls", ctypes.c_uint), ("hashLog", ctypes.c_uint), ("longHashLog", ctypes.c_uint), ("stepSize", ctypes.c_uint),
                ("insPeriod", ctypes.c_uint), ("insPeriodLong", ctypes.c_uint), ("startRep", ctypes.c_uint * 2),
                ("frameStart", ctypes.c_size_t), ("primeBytes", ctypes.c_uint), ("strategy", ctypes.c_uint),
                ("windowLog", ctypes.c_uint), ("litCompressionDisabled", ctypes.c_uint)]


class OCParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint) for n in ("windowLog", "chainLog", "hashLog", "searchLog", "minMatch", "targetLength", "strategy")]


def oracle_plan(level, size, dict_size):
    O = zref.oracle()
    O.zbo_getCParams.restype = OCParams
    O.zbo_getCParams.argtypes = [ctypes.c_int, ctypes.c_ulonglong, ctypes.c_size_t]
    O.zbo_makePlan.argtypes = [ctypes.POINTER(OPlan), ctypes.POINTER(OCParams)]
    cp = O.zbo_getCParams(level, size, dict_size)
    if dict_size and cp.strategy != 1:
        cp.strategy = 1                     # dictionary calls run the two-segment fast match-finder (oracle/zb_frame.c)
    pl = OPlan()
    O.zbo_makePlan(ctypes.byref(pl), ctypes.byref(cp))
    return cp, pl


@pytest.mark.parametrize("dict_size", [0, 16 << 10, 112 << 10])
@pytest.mark.parametrize("level", [-7, -3, -1, 0, 1, 2, 3, 4, 9, 19, 22, 50])
def test_planner_matches_oracle(level, dict_size):
    L = zstd_b200.lib()
    L.ZSTDB200_describePlan.restype = ctypes.c_size_t
    L.ZSTDB200_describePlan.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
    n = len(SIZES)
    sizes = (ctypes.c_size_t * n)(*SIZES)
    out = (ctypes.c_uint * (14 * n))()
    dict_tail = min(dict_size, 64 << 10)
    total = L.ZSTDB200_describePlan(sizes, n, level, dict_size, dict_tail, out)
    blocks = 0
    for f, size in enumerate(SIZES):
        r = out[14 * f:14 * f + 14]
        cp, pl = oracle_plan(level, size, dict_size)
        assert (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8]) == \
               (pl.strategy, pl.mls, pl.hashLog, pl.longHashLog, pl.stepSize, pl.litCompressionDisabled, max(cp.windowLog, 10), pl.insPeriod, pl.insPeriodLong), (size, list(r))
        block_max = min(1 << max(cp.windowLog, 10), 128 << 10)
        nb = max(1, -(-size // block_max))
        assert r[9] == nb and r[10] == min(size, block_max)
        first = 1 | (2 if nb == 1 else 0) | (4 if dict_tail else 0)                       # ZB_FLAG_FIRST | LAST | DICT
        assert r[11] == first
        last_pos = (nb - 1) * block_max
        hist = dict_tail if (nb == 1 and dict_tail) else min(last_pos, 64 << 10)
        assert r[12] == hist
        if not (nb == 1 and dict_tail):
            assert r[13] == (last_pos - hist) % pl.insPeriod
        else:
            assert r[13] == (pl.insPeriod - dict_tail % pl.insPeriod) % pl.insPeriod
        blocks += nb
    assert total == blocks


def test_planner_template_path_equals_slow_path():
    """Runs of equal single-block frames take the planner's template path: same plan as frames planned one by one."""
    L = zstd_b200.lib()
    L.ZSTDB200_describePlan.restype = ctypes.c_size_t
    L.ZSTDB200_describePlan.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
    sizes = [1024] * 5 + [5000] + [1024] * 3 + [0, 0] + [200_000] + [1024] * 2
    arr = (ctypes.c_size_t * len(sizes))(*sizes)
    out = (ctypes.c_uint * (14 * len(sizes)))()
    L.ZSTDB200_describePlan(arr, len(sizes), 1, 16 << 10, 16 << 10, out)
    for f, size in enumerate(sizes):
        one = (ctypes.c_uint * 14)()
        L.ZSTDB200_describePlan((ctypes.c_size_t * 1)(size), 1, 1, 16 << 10, 16 << 10, one)
        assert list(out[14 * f:14 * f + 14]) == list(one), (f, size)


@pytest.mark.skipif(not zref.have_ref(), reason="oracle/_ref/libzstd_ref.so not built")
def test_xxh64_matches_reference():
    """The checksum of ZSTD_c_checksumFlag frames is XXH64(content, 0) & 0xFFFFFFFF (zstd_compress.c:5297-5303)."""
    L, R = zstd_b200.lib(), zref.ref()
    L.ZSTDB200_xxh64.restype = ctypes.c_ulonglong
    L.ZSTDB200_xxh64.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    R.ZSTD_XXH64.restype = ctypes.c_ulonglong
    R.ZSTD_XXH64.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_ulonglong]
    data = zref.random_bytes(100_000, 3)
    for n in list(range(0, 70)) + [255, 256, 257, 4095, 4096, 65537, 100_000]:
        assert L.ZSTDB200_xxh64(data[:n], n) == R.ZSTD_XXH64(data[:n], n, 0), n
