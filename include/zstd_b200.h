/* zstd_b200.h — C ABI of libzstd_b200.so: the B200-native drop-in for zstd's per-block
 * compression hot path (fast / doubleFast match-finder + Huffman literals + FSE sequences).
 *
 * Section 1 re-declares, with identical names, signatures, argument meaning and error
 * behaviour, the reference entry points this library replaces (a language binding that
 * dlopen()s libzstd for these symbols can be pointed at libzstd_b200.so unchanged).
 * Section 2 adds device-pointer and many-frame entry points that have no reference
 * counterpart (the reference has no device memory and no batch call); they are what the
 * benchmark's device-resident `value` and the multi-GPU sharding use.
 *
 * All compression work is done by sm_100a CUDA kernels.  There is no CPU fallback: when no
 * CUDA device is usable every compress call returns ZSTD_error_GENERIC (code 1).
 */
#ifndef ZSTD_B200_H
#define ZSTD_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#  define ZSTDB200_API __attribute__((visibility("default")))
#else
#  define ZSTDB200_API
#endif

/* =====================  1. reference-identical entry points  ===================== */

typedef struct ZSTD_CCtx_s ZSTD_CCtx;                 /* opaque, /root/reference/lib/zstd.h:262 */

/* lib/zstd.h:155 — one call = one complete frame, content size in the header, no checksum.
 * Returns compressed size, or an error code testable with ZSTD_isError(). */
ZSTDB200_API size_t ZSTD_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int compressionLevel);

/* lib/zstd.h:263-264 — context lifecycle; ZSTD_freeCCtx accepts NULL. */
ZSTDB200_API ZSTD_CCtx* ZSTD_createCCtx(void);
ZSTDB200_API size_t     ZSTD_freeCCtx(ZSTD_CCtx* cctx);

/* lib/zstd.h:274 — same as ZSTD_compress with an explicit, reusable context. */
ZSTDB200_API size_t ZSTD_compressCCtx(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity,
                                      const void* src, size_t srcSize, int compressionLevel);

/* lib/zstd.h:944 — compression with a (raw-content or zstd-format) dictionary. */
ZSTDB200_API size_t ZSTD_compress_usingDict(ZSTD_CCtx* ctx, void* dst, size_t dstCapacity,
                                            const void* src, size_t srcSize,
                                            const void* dict, size_t dictSize, int compressionLevel);

/* lib/zstd.h:967-995 — digested dictionary: the dictionary is parsed once, its content tail, entropy tables
 * and primed match-finder tables stay resident on the GPU across calls (the reference's CDict keeps the
 * same things in host memory, zstd_compress.c:5477-5642).  A CDict may be shared by any number of contexts
 * on one device.  ZSTD_createCDict returns NULL for a zstd-format dictionary whose entropy tables are
 * corrupted; ZSTD_freeCDict accepts NULL.  ZSTD_compress_usingCDict compresses at the CDict's level
 * (zstd_compress.c:5836) and produces the bytes ZSTD_compress_usingDict produces for the same inputs. */
typedef struct ZSTD_CDict_s ZSTD_CDict;
ZSTDB200_API ZSTD_CDict* ZSTD_createCDict(const void* dictBuffer, size_t dictSize, int compressionLevel);
ZSTDB200_API size_t      ZSTD_freeCDict(ZSTD_CDict* cdict);
ZSTDB200_API size_t      ZSTD_compress_usingCDict(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity,
                                                  const void* src, size_t srcSize, const ZSTD_CDict* cdict);
/* lib/zstd.h:1099-1111 */
ZSTDB200_API unsigned    ZSTD_getDictID_fromCDict(const ZSTD_CDict* cdict);
ZSTDB200_API unsigned    ZSTD_getDictID_fromDict(const void* dict, size_t dictSize);

/* lib/zstd.h:337-603 — the advanced one-shot API most bindings use today: sticky parameters on the context, then
 * ZSTD_compress2.  Honoured: ZSTD_c_compressionLevel, ZSTD_c_checksumFlag (XXH64 of the content: a serial recurrence per
 * frame — hashed by host threads while the GPU compresses when the input is in host memory, ~10 GB/s per frame; by one
 * warp per frame on the device for device buffers, ~1 GB/s per frame, all frames of a call side by side),
 * ZSTD_c_dictIDFlag, ZSTD_c_contentSizeFlag (the size is always written).  Accepted and ignored: ZSTD_c_nbWorkers,
 * ZSTD_c_jobSize, ZSTD_c_overlapLog.  windowLog .. strategy and the long-distance-matching parameters only at 0;
 * anything else returns ZSTD_error_parameter_unsupported (40).  ZSTD_CCtx_loadDictionary copies and digests the
 * dictionary (lib/zstd.h:1088); ZSTD_CCtx_refCDict borrows a CDict, whose level then applies (:1102). */
typedef enum {
    ZSTD_c_compressionLevel = 100, ZSTD_c_windowLog = 101, ZSTD_c_hashLog = 102, ZSTD_c_chainLog = 103, ZSTD_c_searchLog = 104,
    ZSTD_c_minMatch = 105, ZSTD_c_targetLength = 106, ZSTD_c_strategy = 107,
    ZSTD_c_enableLongDistanceMatching = 160,
    ZSTD_c_contentSizeFlag = 200, ZSTD_c_checksumFlag = 201, ZSTD_c_dictIDFlag = 202,
    ZSTD_c_nbWorkers = 400, ZSTD_c_jobSize = 401, ZSTD_c_overlapLog = 402
} ZSTD_cParameter;
typedef enum { ZSTD_reset_session_only = 1, ZSTD_reset_parameters = 2, ZSTD_reset_session_and_parameters = 3 } ZSTD_ResetDirective;
ZSTDB200_API size_t ZSTD_CCtx_setParameter(ZSTD_CCtx* cctx, ZSTD_cParameter param, int value);
ZSTDB200_API size_t ZSTD_CCtx_setPledgedSrcSize(ZSTD_CCtx* cctx, unsigned long long pledgedSrcSize);
ZSTDB200_API size_t ZSTD_CCtx_reset(ZSTD_CCtx* cctx, ZSTD_ResetDirective reset);
ZSTDB200_API size_t ZSTD_CCtx_loadDictionary(ZSTD_CCtx* cctx, const void* dict, size_t dictSize);
ZSTDB200_API size_t ZSTD_CCtx_refCDict(ZSTD_CCtx* cctx, const ZSTD_CDict* cdict);
ZSTDB200_API size_t ZSTD_compress2(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize);

/* lib/zstd.h:681-862 — streaming.  The GPU works on whole frames, so the stream front end collects input in the context
 * and emits FRAMES: ZSTD_e_continue buffers (256 MiB of input become a frame of their own), ZSTD_e_flush turns what is
 * buffered into a frame now, ZSTD_e_end does the same and ends the session.  The output is therefore a sequence of frames —
 * every zstd decoder reads it as the concatenation of their contents (lib/zstd.h:160-162) — where the reference writes one
 * frame; ZSTD_getFrameContentSize() of such output describes its first frame only.  A session's first call that brings
 * everything with ZSTD_e_end and room for ZSTD_compressBound(size) bytes is compressed straight from the caller's buffers.
 * Return value as in the reference: > 0 while output is still waiting (call again with more room), 0 when flushed / ended;
 * with ZSTD_e_continue a hint for the next input size. */
typedef struct ZSTD_inBuffer_s  { const void* src; size_t size; size_t pos; } ZSTD_inBuffer;
typedef struct ZSTD_outBuffer_s { void* dst; size_t size; size_t pos; } ZSTD_outBuffer;
typedef enum { ZSTD_e_continue = 0, ZSTD_e_flush = 1, ZSTD_e_end = 2 } ZSTD_EndDirective;
ZSTDB200_API size_t ZSTD_compressStream2(ZSTD_CCtx* cctx, ZSTD_outBuffer* output, ZSTD_inBuffer* input, ZSTD_EndDirective endOp);
typedef ZSTD_CCtx ZSTD_CStream;                       /* lib/zstd.h:822: the same object */
ZSTDB200_API ZSTD_CStream* ZSTD_createCStream(void);
ZSTDB200_API size_t ZSTD_freeCStream(ZSTD_CStream* zcs);
ZSTDB200_API size_t ZSTD_initCStream(ZSTD_CStream* zcs, int compressionLevel);
ZSTDB200_API size_t ZSTD_compressStream(ZSTD_CStream* zcs, ZSTD_outBuffer* output, ZSTD_inBuffer* input);
ZSTDB200_API size_t ZSTD_flushStream(ZSTD_CStream* zcs, ZSTD_outBuffer* output);
ZSTDB200_API size_t ZSTD_endStream(ZSTD_CStream* zcs, ZSTD_outBuffer* output);
ZSTDB200_API size_t ZSTD_CStreamInSize(void);
ZSTDB200_API size_t ZSTD_CStreamOutSize(void);

/* lib/zstd.h:236,242-246,114-120 ; lib/zstd_errors.h:106 */
ZSTDB200_API size_t      ZSTD_compressBound(size_t srcSize);
ZSTDB200_API unsigned    ZSTD_isError(size_t code);
ZSTDB200_API const char* ZSTD_getErrorName(size_t code);
ZSTDB200_API int         ZSTD_getErrorCode(size_t functionResult);     /* ZSTD_ErrorCode as int */
ZSTDB200_API int         ZSTD_minCLevel(void);
ZSTDB200_API int         ZSTD_maxCLevel(void);
ZSTDB200_API int         ZSTD_defaultCLevel(void);
ZSTDB200_API unsigned    ZSTD_versionNumber(void);
ZSTDB200_API const char* ZSTD_versionString(void);

/* lib/zstd.h:170-299 — decompression.  Every frame the format allows is accepted: this library's
 * own and the reference encoder's at any level, concatenated frames, skippable frames, frames without a content size,
 * window sizes up to 128 MiB (the reference decoder's default limit, ZSTD_WINDOWLOG_LIMIT_DEFAULT = 27), raw-content
 * and zstd-format dictionaries (ZSTD_decompress_usingDict, lib/zstd.h:955: the dictionary's content is the history in front
 * of every frame, its entropy tables and repeat offsets what a frame's first blocks may reuse; dictionary_wrong = 32 when a
 * frame names another dictionary ID).
 * The content checksum of a frame that carries one is verified (host buffers; checksum_wrong = 22).
 * All decoding work is done by CUDA kernels (zb_decode.cu); no CPU fallback. */
typedef struct ZSTD_DCtx_s ZSTD_DCtx;
ZSTDB200_API size_t     ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
ZSTDB200_API ZSTD_DCtx* ZSTD_createDCtx(void);
ZSTDB200_API size_t     ZSTD_freeDCtx(ZSTD_DCtx* dctx);                                    /* accepts NULL */
ZSTDB200_API size_t     ZSTD_decompressDCtx(ZSTD_DCtx* dctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize);
ZSTDB200_API size_t     ZSTD_decompress_usingDict(ZSTD_DCtx* dctx, void* dst, size_t dstCapacity, const void* src, size_t srcSize,
                                                  const void* dict, size_t dictSize);
/* lib/zstd.h:195-227 — header readers (host code).  ZSTD_CONTENTSIZE_UNKNOWN = (0ULL - 1), ZSTD_CONTENTSIZE_ERROR = (0ULL - 2). */
ZSTDB200_API unsigned long long ZSTD_getFrameContentSize(const void* src, size_t srcSize);
ZSTDB200_API size_t     ZSTD_findFrameCompressedSize(const void* src, size_t srcSize);

/* lib/zstd.h:880-924 — streaming decompression.  Whole frames are decoded on the GPU: compressed bytes are collected in
 * the context until a frame is complete, then decoded and handed out as the caller makes room.  Returns 0 when a frame has
 * been decoded and handed out completely, else a hint (> 0) for the next call, or an error code. */
typedef ZSTD_DCtx ZSTD_DStream;
ZSTDB200_API ZSTD_DStream* ZSTD_createDStream(void);
ZSTDB200_API size_t ZSTD_freeDStream(ZSTD_DStream* zds);
ZSTDB200_API size_t ZSTD_initDStream(ZSTD_DStream* zds);
ZSTDB200_API size_t ZSTD_decompressStream(ZSTD_DStream* zds, ZSTD_outBuffer* output, ZSTD_inBuffer* input);
ZSTDB200_API size_t ZSTD_DStreamInSize(void);
ZSTDB200_API size_t ZSTD_DStreamOutSize(void);

/* =====================  2. B200 extensions (no reference counterpart)  ===================== */

/* Decompress frames whose bytes are in device memory into device memory.  The frame / block headers are a chain that has to
 * be followed in order: for inputs of up to 512 MiB the compressed bytes are copied to a page-locked host buffer and walked
 * there, beyond that one device thread follows them (about 1 us per block); everything else is block-parallel.  Content
 * checksums are not verified on this path.  `stream`: as for ZSTDB200_compressDevice.  Returns the decompressed size. */
ZSTDB200_API size_t ZSTDB200_decompressDevice(ZSTD_DCtx* dctx, void* d_dst, size_t dstCapacity, const void* d_src, size_t srcSize, void* stream);
/* same with a dictionary (host memory; uploaded by the call) */
ZSTDB200_API size_t ZSTDB200_decompressDevice_usingDict(ZSTD_DCtx* dctx, void* d_dst, size_t dstCapacity, const void* d_src, size_t srcSize,
                                                        const void* dict, size_t dictSize, void* stream);
typedef struct {
    float kernel_ms;         /* literals kernel start -> match kernel end */
    float literals_ms, sequences_ms, place_ms, execute_ms;       /* D1, D2, D4 (literal placement), D5 (match copies) */
    unsigned launches, nbBlocks, nbFrames;
    size_t h2d_bytes, d2h_bytes;
} ZSTDB200_dstats;
ZSTDB200_API void ZSTDB200_getLastDStats(const ZSTD_DCtx* dctx, ZSTDB200_dstats* out);


/* Compress one frame whose input and output already live in device memory (HBM).  The call returns when the frame is
 * complete (it synchronises to read the size).
 * `stream` is a cudaStream_t.  Non-NULL: all work of the call is enqueued on that stream, behind whatever the caller
 * queued there before (the way to compress the output of a kernel that is still running).  NULL: the context's own
 * NON-BLOCKING streams are used — they are NOT ordered after the legacy default stream or any other stream, so the
 * producer of d_src must have completed (e.g. cudaStreamSynchronize) before the call.  Large NULL-stream calls run as
 * several waves on several streams.  d_src needs no padding: no byte outside [d_src, d_src + srcSize) is read. */
ZSTDB200_API size_t ZSTDB200_compressDevice(ZSTD_CCtx* cctx, void* d_dst, size_t dstCapacity,
                                            const void* d_src, size_t srcSize, int compressionLevel, void* stream);

/* One frame compressed by several GPUs (the reference's counterpart: the jobs of ZSTDMT, zstdmt_compress.c:1168-1227 —
 * every job reads an overlap of the input in front of it, only the first writes the frame header, only the last the end
 * mark).  Each rank calls this for its share [partBegin, partBegin + partSize) of a frame of frameSize bytes; partBegin must
 * be a multiple of ZSTDB200_framePartAlignment() and d_part must point at the frame's byte partBegin - min(partBegin,
 * ZSTDB200_framePartHalo()): the rank needs that much of the preceding input.  The ranks' outputs, concatenated in
 * order, are byte for byte the frame a single ZSTDB200_compressDevice call produces (zstd_b200/sharding.py gathers them).
 * No content checksum (ZSTD_c_checksumFlag must be off). */
ZSTDB200_API size_t ZSTDB200_framePartAlignment(void);
ZSTDB200_API size_t ZSTDB200_framePartHalo(void);
ZSTDB200_API size_t ZSTDB200_compressFramePart(ZSTD_CCtx* cctx, void* d_dst, size_t dstCapacity, const void* d_part,
                                               size_t frameSize, size_t partBegin, size_t partSize, int compressionLevel, void* stream);

/* Compress nbFrames independent inputs src[frameOffsets[i] .. +frameSizes[i]) into nbFrames
 * complete frames written back to back into dst (the decoder accepts the concatenation,
 * lib/zstd.h:160-162).  cSizes[i] (host array, may be NULL) receives each frame's size.
 * With dict != NULL every frame is compressed as ZSTD_compress_usingDict would (config 5).
 * `deviceMemory` != 0: src/dst are device pointers (dict and the offset arrays stay on the host).
 * The context's sticky ZSTD_c_checksumFlag / ZSTD_c_dictIDFlag apply to every frame of the call (level and dictionary are arguments). */
ZSTDB200_API size_t ZSTDB200_compressFrames(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity,
                                            const void* src, const size_t* frameOffsets, const size_t* frameSizes,
                                            size_t nbFrames, const void* dict, size_t dictSize,
                                            size_t* cSizes, int compressionLevel, int deviceMemory, void* stream);

/* Same with a digested dictionary (level = the CDict's): what a record store calling
 * ZSTD_compress_usingCDict in a loop would batch (contrib/largeNbDicts/largeNbDicts.c:553). */
ZSTDB200_API size_t ZSTDB200_compressFrames_usingCDict(ZSTD_CCtx* cctx, void* dst, size_t dstCapacity,
                                            const void* src, const size_t* frameOffsets, const size_t* frameSizes,
                                            size_t nbFrames, const ZSTD_CDict* cdict,
                                            size_t* cSizes, int deviceMemory, void* stream);

/* Timing / evidence of the last call on this context (CUDA events on the launching stream). */
typedef struct {
    float  kernel_ms;        /* first kernel start -> last kernel end */
    float  match_ms;         /* K1a candidate walk + K1b parse */
    float  cand_ms, parse_ms; /* K1a, K1b separately (single parameter group only, else 0) */
    float  literals_ms;      /* K2 */
    float  sequences_ms;     /* K3 */
    float  stitch_ms;        /* K4 scan + copy */
    float  total_ms;         /* including host<->device copies, when the call made any */
    unsigned launches;       /* kernels launched by the call */
    unsigned nbBlocks;       /* 128 KiB blocks processed */
    size_t h2d_bytes, d2h_bytes;
} ZSTDB200_stats;
ZSTDB200_API void ZSTDB200_getLastStats(const ZSTD_CCtx* cctx, ZSTDB200_stats* out);

/* Seek table for the frames of a ZSTDB200_compressFrames call (or of several ranks' outputs laid end to end): the
 * reference's seekable format (contrib/seekable_format/zstd_seekable_compression_format.md; its writer is
 * ZSTD_seekable_writeSeekTable, zstdseek_compress.c:297).  Append the bytes behind the frames and the reference's
 * ZSTD_seekable_* readers can decompress any range.  cSizes / dSizes: compressed and decompressed size of every frame
 * (compressed < 4 GiB, decompressed <= 1 GiB each: the format's limit, zstd_seekable.h:19).  Host code.  Returns 17 + 8 * nbFrames, or an error code. */
ZSTDB200_API size_t ZSTDB200_writeSeekTable(void* dst, size_t dstCapacity, const size_t* cSizes, const size_t* dSizes, size_t nbFrames);

/* XXH64 (seed 0) as used for the frame checksum (lib/common/xxhash.h); host code, no GPU (test hook). */
ZSTDB200_API unsigned long long ZSTDB200_xxh64(const void* data, size_t size);

/* The host planner's view of a call, computed without a GPU (test hook: tests/test_plan.py compares it with the
 * oracle's plan).  out: nbFrames x 16 unsigned = strategy, mls, tableN, tableNLong, stepSize, litDisabled, windowLog,
 * insStep, blocks of the frame, first block's size and flags, last block's history reach, dictionary part of the first
 * block's history, chunks of the frame, history walked by the last chunk, size of the last chunk.
 * Returns the total number of blocks. */
ZSTDB200_API size_t ZSTDB200_describePlan(const size_t* frameSizes, size_t nbFrames, int compressionLevel,
                                          size_t dictSize, size_t dictTail, unsigned* out);

/* COMPRESSION LEVELS.  This library implements the reference's `fast` and `doubleFast` strategies, i.e. negative
 * levels and levels 1-4 (clevels.h:25-50; level 4 only for inputs > 256 KiB).  A level whose reference strategy is
 * greedy or stronger (5 ... 22) is NOT implemented: by default such a call is served by the strongest doubleFast row
 * of its size class and returns a valid frame that is LARGER (typically 20-40 %) than what the reference produces at
 * that level.  ZSTDB200_setStrictLevels(1) turns that into ZSTD_error_parameter_unsupported for the whole process. */
ZSTDB200_API void ZSTDB200_setStrictLevels(int on);

/* Which CUDA device contexts created FROM NOW ON belong to (default: the creating thread's current device).  A context
 * keeps the device it was created for; every call restores the calling thread's current device before it returns. */
ZSTDB200_API int  ZSTDB200_setDevice(int device);
ZSTDB200_API int  ZSTDB200_deviceAvailable(void);

#ifdef __cplusplus
}
#endif
#endif /* ZSTD_B200_H */
