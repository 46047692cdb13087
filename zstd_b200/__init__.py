"""zstd_b200 — Python host-side mirror of the libzstd C API served by libzstd_b200.so.

The product is the C-ABI shared library (include/zstd_b200.h); this module is the thin ctypes
binding a Python caller (tests, bench.py, torch.distributed sharding) uses.  Function names,
argument meaning and error behaviour follow the reference's simple API
(/root/reference/lib/zstd.h:155 ZSTD_compress, :274 ZSTD_compressCCtx, :944 ZSTD_compress_usingDict).

There is no CPU fallback: importing works anywhere, but every compress call raises ZstdError
when the CUDA library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZSTDB200_LIB") or os.path.join(_HERE, "libzstd_b200.so")   # override: development variants only

_lib = None
_sz = ctypes.c_size_t
_vp = ctypes.c_void_p


class ZstdError(RuntimeError):
    """Raised with the reference's error name (lib/common/error_private.c:14-62)."""

    def __init__(self, code: int, name: str):
        super().__init__(f"zstd_b200 error {code}: {name}")
        self.code = code
        self.name = name


class Stats(ctypes.Structure):
    _fields_ = [("kernel_ms", ctypes.c_float), ("match_ms", ctypes.c_float), ("cand_ms", ctypes.c_float), ("parse_ms", ctypes.c_float), ("literals_ms", ctypes.c_float),
                ("sequences_ms", ctypes.c_float), ("stitch_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("launches", ctypes.c_uint), ("nbBlocks", ctypes.c_uint),
                ("h2d_bytes", _sz), ("d2h_bytes", _sz)]


class DStats(ctypes.Structure):
    _fields_ = [("kernel_ms", ctypes.c_float), ("literals_ms", ctypes.c_float), ("sequences_ms", ctypes.c_float), ("place_ms", ctypes.c_float),
                ("execute_ms", ctypes.c_float), ("launches", ctypes.c_uint), ("nbBlocks", ctypes.c_uint), ("nbFrames", ctypes.c_uint), ("h2d_bytes", _sz), ("d2h_bytes", _sz)]


def lib() -> ctypes.CDLL:
    """Load libzstd_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). zstd_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL)
    L.ZSTD_compress.restype = _sz
    L.ZSTD_compress.argtypes = [_vp, _sz, _vp, _sz, ctypes.c_int]
    L.ZSTD_createCCtx.restype = _vp
    L.ZSTD_createCCtx.argtypes = []
    L.ZSTD_freeCCtx.restype = _sz
    L.ZSTD_freeCCtx.argtypes = [_vp]
    L.ZSTD_compressCCtx.restype = _sz
    L.ZSTD_compressCCtx.argtypes = [_vp, _vp, _sz, _vp, _sz, ctypes.c_int]
    L.ZSTD_compress_usingDict.restype = _sz
    L.ZSTD_compress_usingDict.argtypes = [_vp, _vp, _sz, _vp, _sz, _vp, _sz, ctypes.c_int]
    L.ZSTD_compressBound.restype = _sz
    L.ZSTD_compressBound.argtypes = [_sz]
    L.ZSTD_isError.restype = ctypes.c_uint
    L.ZSTD_isError.argtypes = [_sz]
    L.ZSTD_getErrorName.restype = ctypes.c_char_p
    L.ZSTD_getErrorName.argtypes = [_sz]
    L.ZSTD_getErrorCode.restype = ctypes.c_int
    L.ZSTD_getErrorCode.argtypes = [_sz]
    for f in ("ZSTD_minCLevel", "ZSTD_maxCLevel", "ZSTD_defaultCLevel"):
        getattr(L, f).restype = ctypes.c_int
        getattr(L, f).argtypes = []
    L.ZSTD_versionNumber.restype = ctypes.c_uint
    L.ZSTD_versionString.restype = ctypes.c_char_p
    L.ZSTDB200_compressDevice.restype = _sz
    L.ZSTDB200_compressDevice.argtypes = [_vp, _vp, _sz, _vp, _sz, ctypes.c_int, _vp]
    L.ZSTDB200_compressFrames.restype = _sz
    L.ZSTDB200_compressFrames.argtypes = [_vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _sz, _vp, ctypes.c_int, ctypes.c_int, _vp]
    L.ZSTDB200_getLastStats.restype = None
    L.ZSTDB200_getLastStats.argtypes = [_vp, ctypes.POINTER(Stats)]
    L.ZSTDB200_setDevice.restype = ctypes.c_int
    L.ZSTDB200_setDevice.argtypes = [ctypes.c_int]
    L.ZSTDB200_deviceAvailable.restype = ctypes.c_int
    if os.environ.get("ZSTDB200_LIB") and not hasattr(L, "ZSTD_createCDict"):      # an older development build
        _lib = L
        return L
    L.ZSTD_createCDict.restype = _vp
    L.ZSTD_createCDict.argtypes = [_vp, _sz, ctypes.c_int]
    L.ZSTD_freeCDict.restype = _sz
    L.ZSTD_freeCDict.argtypes = [_vp]
    L.ZSTD_compress_usingCDict.restype = _sz
    L.ZSTD_compress_usingCDict.argtypes = [_vp, _vp, _sz, _vp, _sz, _vp]
    L.ZSTD_getDictID_fromCDict.restype = ctypes.c_uint
    L.ZSTD_getDictID_fromCDict.argtypes = [_vp]
    L.ZSTD_getDictID_fromDict.restype = ctypes.c_uint
    L.ZSTD_getDictID_fromDict.argtypes = [_vp, _sz]
    L.ZSTD_CCtx_setParameter.restype = _sz
    L.ZSTD_CCtx_setParameter.argtypes = [_vp, ctypes.c_int, ctypes.c_int]
    L.ZSTD_CCtx_reset.restype = _sz
    L.ZSTD_CCtx_reset.argtypes = [_vp, ctypes.c_int]
    L.ZSTD_CCtx_loadDictionary.restype = _sz
    L.ZSTD_CCtx_loadDictionary.argtypes = [_vp, _vp, _sz]
    L.ZSTD_CCtx_refCDict.restype = _sz
    L.ZSTD_CCtx_refCDict.argtypes = [_vp, _vp]
    L.ZSTD_compress2.restype = _sz
    L.ZSTD_compress2.argtypes = [_vp, _vp, _sz, _vp, _sz]
    L.ZSTD_compressStream2.restype = _sz
    L.ZSTD_compressStream2.argtypes = [_vp, _vp, _vp, ctypes.c_int]
    L.ZSTDB200_compressFrames_usingCDict.restype = _sz
    L.ZSTDB200_compressFrames_usingCDict.argtypes = [_vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _vp, ctypes.c_int, _vp]
    if hasattr(L, "ZSTD_createDCtx"):
        L.ZSTD_createDCtx.restype = _vp
        L.ZSTD_createDCtx.argtypes = []
        L.ZSTD_freeDCtx.restype = _sz
        L.ZSTD_freeDCtx.argtypes = [_vp]
        L.ZSTD_decompressDCtx.restype = _sz
        L.ZSTD_decompressDCtx.argtypes = [_vp, _vp, _sz, _vp, _sz]
        L.ZSTD_decompress.restype = _sz
        L.ZSTD_decompress.argtypes = [_vp, _sz, _vp, _sz]
        L.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
        L.ZSTD_getFrameContentSize.argtypes = [_vp, _sz]
        L.ZSTD_findFrameCompressedSize.restype = _sz
        L.ZSTD_findFrameCompressedSize.argtypes = [_vp, _sz]
        L.ZSTDB200_decompressDevice.restype = _sz
        L.ZSTDB200_decompressDevice.argtypes = [_vp, _vp, _sz, _vp, _sz, _vp]
        L.ZSTDB200_getLastDStats.restype = None
        L.ZSTDB200_getLastDStats.argtypes = [_vp, ctypes.POINTER(DStats)]
    _lib = L
    return L


def _check(code: int) -> int:
    L = lib()
    if L.ZSTD_isError(code):
        raise ZstdError(L.ZSTD_getErrorCode(code), L.ZSTD_getErrorName(code).decode())
    return code


def ZSTD_compressBound(src_size: int) -> int:
    return lib().ZSTD_compressBound(src_size)


def _buf(data) -> Tuple[ctypes.c_void_p, int, object]:
    """(pointer, nbytes, keepalive) for bytes / bytearray / memoryview / numpy arrays."""
    if isinstance(data, (bytes, bytearray)):
        keep = (ctypes.c_char * len(data)).from_buffer_copy(data) if isinstance(data, bytes) else (ctypes.c_char * len(data)).from_buffer(data)
        return ctypes.cast(keep, _vp), len(data), keep
    mv = memoryview(data).cast("B")
    keep = (ctypes.c_char * len(mv)).from_buffer(mv) if not mv.readonly else (ctypes.c_char * len(mv)).from_buffer_copy(mv)
    return ctypes.cast(keep, _vp), len(mv), keep


class ZSTD_CDict:
    """Digested dictionary (lib/zstd.h:967-995): parsed once, resident on the GPU from its first use."""

    def __init__(self, dict_bytes, level: int = 3):
        p, n, keep = _buf(dict_bytes)
        self._h = lib().ZSTD_createCDict(p, n, level)
        if not self._h:
            raise ZstdError(30, "ZSTD_createCDict failed (dictionary corrupted or out of memory)")

    @property
    def dict_id(self) -> int:
        return int(lib().ZSTD_getDictID_fromCDict(self._h))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.ZSTD_freeCDict(h)

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter teardown
            pass


class ZSTD_CCtx:
    """Reusable compression context (lib/zstd.h:259-264): owns the device workspace and a stream."""

    def __init__(self, device: Optional[int] = None):
        L = lib()
        if device is not None:
            L.ZSTDB200_setDevice(int(device))
        self._h = L.ZSTD_createCCtx()
        if not self._h:
            raise MemoryError("ZSTD_createCCtx failed")

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.ZSTD_freeCCtx(h)

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter teardown
            pass

    # -- reference-identical calls (host buffers) --
    def compress(self, src, level: int = 3, dst_capacity: Optional[int] = None) -> bytes:
        """ZSTD_compressCCtx: one complete frame (content size in header, no checksum)."""
        p, n, keep = _buf(src)
        cap = ZSTD_compressBound(n) if dst_capacity is None else dst_capacity
        dst = ctypes.create_string_buffer(max(cap, 1))
        r = _check(lib().ZSTD_compressCCtx(self._h, dst, cap, p, n, level))
        return dst.raw[:r]

    def compress_using_dict(self, src, dict_bytes, level: int = 3) -> bytes:
        p, n, keep = _buf(src)
        dp, dn, dkeep = _buf(dict_bytes)
        cap = ZSTD_compressBound(n)
        dst = ctypes.create_string_buffer(max(cap, 1))
        r = _check(lib().ZSTD_compress_usingDict(self._h, dst, cap, p, n, dp, dn, level))
        return dst.raw[:r]

    def compress_using_cdict(self, src, cdict: "ZSTD_CDict") -> bytes:
        """ZSTD_compress_usingCDict (lib/zstd.h:987): level and dictionary come from the CDict."""
        p, n, keep = _buf(src)
        cap = ZSTD_compressBound(n)
        dst = ctypes.create_string_buffer(max(cap, 1))
        r = _check(lib().ZSTD_compress_usingCDict(self._h, dst, cap, p, n, cdict._h))
        return dst.raw[:r]

    # -- advanced one-shot API (lib/zstd.h:337-603) --
    PARAMS = {"compression_level": 100, "content_size_flag": 200, "checksum_flag": 201, "dict_id_flag": 202, "nb_workers": 400}

    def set_parameter(self, name_or_id, value: int) -> None:
        """ZSTD_CCtx_setParameter: sticky until ZSTD_CCtx_reset(parameters)."""
        pid = self.PARAMS.get(name_or_id, name_or_id)
        _check(lib().ZSTD_CCtx_setParameter(self._h, int(pid), int(value)))

    def reset(self, directive: int = 3) -> None:
        """ZSTD_CCtx_reset: 1 session only, 2 parameters, 3 both."""
        _check(lib().ZSTD_CCtx_reset(self._h, directive))

    def load_dictionary(self, dict_bytes) -> None:
        p, n, keep = (None, 0, None) if not dict_bytes else _buf(dict_bytes)
        _check(lib().ZSTD_CCtx_loadDictionary(self._h, p, n))

    def ref_cdict(self, cdict: Optional["ZSTD_CDict"]) -> None:
        _check(lib().ZSTD_CCtx_refCDict(self._h, cdict._h if cdict is not None else None))

    def compress2(self, src) -> bytes:
        """ZSTD_compress2 with the context's sticky parameters / dictionary."""
        p, n, keep = _buf(src)
        cap = ZSTD_compressBound(n) + 8
        dst = ctypes.create_string_buffer(max(cap, 1))
        r = _check(lib().ZSTD_compress2(self._h, dst, cap, p, n))
        return dst.raw[:r]

    # -- B200 extensions --
    def compress_device(self, d_dst: int, dst_capacity: int, d_src: int, src_size: int, level: int = 3, stream: int = 0) -> int:
        """One frame, device pointers (ints, e.g. torch.Tensor.data_ptr()).  Returns compressed size."""
        return _check(lib().ZSTDB200_compressDevice(self._h, d_dst, dst_capacity, d_src, src_size, level, stream))

    def compress_frame_part(self, d_dst: int, dst_capacity: int, d_part: int, frame_size: int, part_begin: int, part_size: int,
                            level: int = 3, stream: int = 0) -> int:
        """This rank's share of a frame several GPUs compress together (ZSTDB200_compressFramePart).  d_part: device address
        of the frame's byte part_begin - min(part_begin, halo).  Returns the bytes this share contributes."""
        L = lib()
        L.ZSTDB200_compressFramePart.restype = _sz
        L.ZSTDB200_compressFramePart.argtypes = [_vp, _vp, _sz, _vp, _sz, _sz, _sz, ctypes.c_int, _vp]
        return _check(L.ZSTDB200_compressFramePart(self._h, d_dst, dst_capacity, d_part, frame_size, part_begin, part_size, level, stream))

    def compress_frames(self, dst: int, dst_capacity: int, src: int, offsets: Sequence[int], sizes: Sequence[int],
                        level: int = 3, device_memory: bool = True, dict_bytes=None, stream: int = 0):
        """Many independent frames in one call.  Returns (total_bytes, [compressed size per frame])."""
        n = len(sizes)
        offs = (_sz * n)(*offsets)
        szs = (_sz * n)(*sizes)
        csz = (_sz * n)()
        dp, dn, dkeep = (None, 0, None) if dict_bytes is None else _buf(dict_bytes)
        r = _check(lib().ZSTDB200_compressFrames(self._h, dst, dst_capacity, src, offs, szs, n, dp, dn, csz, level,
                                                1 if device_memory else 0, stream))
        return r, list(csz)

    def compress_frames_using_cdict(self, dst: int, dst_capacity: int, src: int, offsets: Sequence[int], sizes: Sequence[int],
                                    cdict: "ZSTD_CDict", device_memory: bool = True, stream: int = 0):
        """Many independent frames against one digested dictionary.  Returns (total_bytes, [size per frame])."""
        n = len(sizes)
        offs = (_sz * n)(*offsets)
        szs = (_sz * n)(*sizes)
        csz = (_sz * n)()
        r = _check(lib().ZSTDB200_compressFrames_usingCDict(self._h, dst, dst_capacity, src, offs, szs, n, cdict._h, csz,
                                                           1 if device_memory else 0, stream))
        return r, list(csz)

    def stats(self) -> Stats:
        s = Stats()
        lib().ZSTDB200_getLastStats(self._h, ctypes.byref(s))
        return s


class ZSTD_DCtx:
    """Reusable decompression context (lib/zstd.h:289-299): owns the device workspace and a stream."""

    def __init__(self, device: Optional[int] = None):
        L = lib()
        if device is not None:
            L.ZSTDB200_setDevice(int(device))
        self._h = L.ZSTD_createDCtx()
        if not self._h:
            raise MemoryError("ZSTD_createDCtx failed")

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.ZSTD_freeDCtx(h)

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter teardown
            pass

    def decompress(self, frames, max_size: Optional[int] = None) -> bytes:
        """ZSTD_decompressDCtx: one or more concatenated frames (host buffers).  max_size defaults to the content size
        the first frame's header states."""
        p, n, keep = _buf(frames)
        if max_size is None:
            cs = lib().ZSTD_getFrameContentSize(p, n)
            if cs >= (1 << 64) - 2:
                raise ZstdError(72, "content size unknown: pass max_size")
            max_size = cs
        dst = ctypes.create_string_buffer(max(max_size, 1))
        r = _check(lib().ZSTD_decompressDCtx(self._h, dst, max_size, p, n))
        return dst.raw[:r]

    def decompress_device(self, d_dst: int, dst_capacity: int, d_src: int, src_size: int, stream: int = 0) -> int:
        """Frames in device memory -> device memory (ints, e.g. torch.Tensor.data_ptr()).  Returns the decompressed size."""
        return _check(lib().ZSTDB200_decompressDevice(self._h, d_dst, dst_capacity, d_src, src_size, stream))

    def stats(self) -> DStats:
        s = DStats()
        lib().ZSTDB200_getLastDStats(self._h, ctypes.byref(s))
        return s


def ZSTD_decompress(frames, max_size: Optional[int] = None) -> bytes:
    """lib/zstd.h:170 — temporary context, host buffers."""
    d = ZSTD_DCtx()
    try:
        return d.decompress(frames, max_size)
    finally:
        d.close()


def ZSTD_compress(src, level: int = 3) -> bytes:
    """lib/zstd.h:155 — temporary context, host buffers."""
    p, n, keep = _buf(src)
    cap = ZSTD_compressBound(n)
    dst = ctypes.create_string_buffer(max(cap, 1))
    r = _check(lib().ZSTD_compress(dst, cap, p, n, level))
    return dst.raw[:r]


def seek_table(c_sizes: Sequence[int], d_sizes: Sequence[int]) -> bytes:
    """ZSTDB200_writeSeekTable: the seekable-format footer for a run of frames (append it behind them)."""
    n = len(c_sizes)
    cs, ds = (_sz * n)(*c_sizes), (_sz * n)(*d_sizes)
    cap = 17 + 8 * n
    dst = ctypes.create_string_buffer(cap)
    L = lib()
    L.ZSTDB200_writeSeekTable.restype = _sz
    L.ZSTDB200_writeSeekTable.argtypes = [_vp, _sz, _vp, _vp, _sz]
    r = _check(L.ZSTDB200_writeSeekTable(dst, cap, cs, ds, n))
    return dst.raw[:r]


def device_available() -> bool:
    try:
        return bool(lib().ZSTDB200_deviceAvailable())
    except (ImportError, OSError):
        return False
