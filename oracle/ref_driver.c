/* ref_driver.c — TEST / BENCH INFRASTRUCTURE ONLY (linked into oracle/_ref/libzstd_ref.so with the unmodified reference).
 * Loops over many frames inside ONE C call, so that a host thread of bench.py's reference arm spends its time in the
 * reference's ZSTD_compressCCtx / ZSTD_compress_usingCDict and not in the Python interpreter (contrib/pzstd and
 * contrib/largeNbDicts/largeNbDicts.c:553 are the reference's own drivers of these shapes). */
#include <stddef.h>
#include "zstd.h"

/* frames [first, first+count) of src (offsets / sizes arrays) -> dst back to back; returns the bytes written or an error code */
size_t refdrv_frames(ZSTD_CCtx* cctx, const void* src, const size_t* offs, const size_t* sizes, size_t first, size_t count,
                     void* dst, size_t dstCapacity, int level)
{
    size_t pos = 0, i;
    for (i = first; i < first + count; i++) {
        size_t const r = ZSTD_compressCCtx(cctx, (char*)dst + pos, dstCapacity - pos, (const char*)src + offs[i], sizes[i], level);
        if (ZSTD_isError(r)) return r;
        pos += r;
    }
    return pos;
}

/* records of recSize bytes [first, first+count) compressed one by one against a digested dictionary */
size_t refdrv_records_cdict(ZSTD_CCtx* cctx, const ZSTD_CDict* cdict, const void* src, size_t recSize, size_t first, size_t count,
                            void* dst, size_t dstCapacity)
{
    size_t pos = 0, i;
    for (i = first; i < first + count; i++) {
        size_t const r = ZSTD_compress_usingCDict(cctx, (char*)dst + pos, dstCapacity - pos, (const char*)src + i * recSize, recSize, cdict);
        if (ZSTD_isError(r)) return r;
        pos += r;
    }
    return pos;
}
