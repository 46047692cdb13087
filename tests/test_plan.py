"""Host logic of the product without a GPU: the planner of libzstd_b200.so (parameter derivation per level / size /
dictionary, chunk and block partition, block flags, history reach) against the oracle's plan, which is itself
pinned to the reference's ZSTD_getCParams (tests/test_oracle_frames.py::test_cparams_match_reference)."""
import ctypes

import pytest

import zref
import zstd_b200

SIZES = [0, 1, 6, 7, 100, 1024, 4096, 16 << 10, (16 << 10) + 1, 65535, 128 << 10, (128 << 10) + 1, 200_000, 256 << 10, (256 << 10) + 1,
         512 << 10, (512 << 10) + 1, 1 << 20, (8 << 20) + 12345, 64 << 20, (1 << 30) + 77]
PRIME = 128 << 10
CHUNK_BLOCKS = 4


class OPlan(ctypes.Structure):
    _fields_ = [("mls", ctypes.c_uint), ("tableN", ctypes.c_uint), ("tableNLong", ctypes.c_uint), ("stepSize", ctypes.c_uint),
                ("insStep", ctypes.c_uint), ("chunkBlocks", ctypes.c_uint), ("startRep", ctypes.c_uint * 2), ("codeRep", ctypes.c_uint * 3),
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
    pl = OPlan()
    O.zbo_makePlan(ctypes.byref(pl), ctypes.byref(cp))
    return cp, pl


@pytest.mark.parametrize("dict_size", [0, 16 << 10, 112 << 10, 300 << 10])
@pytest.mark.parametrize("level", [-7, -3, -1, 0, 1, 2, 3, 4, 9, 19, 22, 50])
def test_planner_matches_oracle(level, dict_size):
    L = zstd_b200.lib()
    L.ZSTDB200_describePlan.restype = ctypes.c_size_t
    L.ZSTDB200_describePlan.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
    n = len(SIZES)
    sizes = (ctypes.c_size_t * n)(*SIZES)
    out = (ctypes.c_uint * (16 * n))()
    dict_tail = min(dict_size, PRIME)
    total = L.ZSTDB200_describePlan(sizes, n, level, dict_size, dict_tail, out)
    blocks = 0
    for f, size in enumerate(SIZES):
        r = out[16 * f:16 * f + 16]
        cp, pl = oracle_plan(level, size, dict_size)
        assert pl.primeBytes == PRIME and pl.chunkBlocks == CHUNK_BLOCKS
        wlog = max(cp.windowLog, 10)
        assert (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]) == \
               (pl.strategy, pl.mls, pl.tableN, pl.tableNLong, pl.stepSize, pl.litCompressionDisabled, wlog, pl.insStep), (size, list(r))
        block_max = min(1 << wlog, 128 << 10)
        nb = max(1, -(-size // block_max))
        assert r[8] == nb and r[9] == min(size, block_max)
        first = 1 | (2 if nb == 1 else 0) | (4 if dict_tail else 0)                       # ZB_FLAG_FIRST | LAST | DICT
        assert r[10] == first and r[12] == dict_tail
        # the last block: its chunk's history start and the window bound what it can reach (oracle/zb_match.c block_low)
        chunk_bytes = CHUNK_BLOCKS * block_max
        last_bs = (nb - 1) * block_max
        cs = last_bs - last_bs % chunk_bytes
        chunk_hist = dict_tail if cs == 0 else min(PRIME, cs)
        low = dict_tail + cs - chunk_hist
        be = dict_tail + size
        if be > (1 << wlog) and be - (1 << wlog) > low:
            low = be - (1 << wlog)
        assert r[11] == dict_tail + last_bs - low, (size, list(r))
        nc = max(1, -(-size // chunk_bytes))
        assert r[13] == nc and r[14] == chunk_hist and r[15] == size - cs
        blocks += nb
    assert total == blocks


def test_strict_levels_switch():
    """levels above the doubleFast rows are served by the strongest doubleFast row unless the caller asks for an error"""
    L = zstd_b200.lib()
    assert hasattr(L, "ZSTDB200_setStrictLevels")
