/* zb_common.h — shared host/device definitions of the block-parallel plan.
 *
 * Unit of work = one zstd block (<= 128 KiB, ZSTD_BLOCKSIZE_MAX, /root/reference/lib/zstd.h:142).
 * Every block is compressed independently of its neighbours: private hash table primed from the
 * history bytes that precede it, encoder repcodes start invalid, fresh entropy tables.  The
 * per-block state the reference carries across blocks (ZSTD_blockState_t,
 * lib/compress/zstd_compress_internal.h:263-267) therefore does not exist here.
 */
#ifndef ZB_COMMON_H
#define ZB_COMMON_H
#include <stdint.h>
#include <stddef.h>

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

#define ZB_BLOCK_MAX     (128u << 10)
#define ZB_PRIME_BYTES   (128u << 10)        /* history primed into a chunk's table (ZSTDMT overlap idea, zstdmt_compress.c:1182-1227) */
#define ZB_CHUNK_BLOCKS  4u                  /* blocks per chunk: one hash table lives through a chunk */
#define ZB_BATCH         1024u               /* positions of one walk batch = threads of the walk CTA */
#define ZB_TAG_BITS      11u                 /* table entry = (position + 1) << 11 | tag, positions relative to the chunk's history start */
#define ZB_FAST_HASHLOG_MAX   14u
#ifndef ZB_DFAST_SHORT_MAX
#define ZB_DFAST_SHORT_MAX    28672u         /* buckets: 112 KiB of shared memory, two walk CTAs per SM */
#endif
#ifndef ZB_DFAST_LONGLOG_MAX
#define ZB_DFAST_LONGLOG_MAX  14u            /* 64 KiB */
#endif
#define ZB_FAR           0xFFFFu             /* dist16 value: the distance is in the far array */
#define ZB_MAX_SEQ       (ZB_BLOCK_MAX / 4)  /* every sequence carries a match of >= 4 bytes */
#define ZB_SEQ_STRIDE    (ZB_MAX_SEQ + 8)    /* u64 per block */
#define ZB_LIT_STRIDE    (ZB_BLOCK_MAX + 256)/* bytes per block */
#define ZB_BODY_STRIDE   (ZB_BLOCK_MAX + 1024)/* staging for one compressed block body */
#define ZB_STATE_STRIDE  (ZB_MAX_SEQ)        /* u16 per FSE stream per block (aliases the dist area: 3*32768 <= 131072) */

/* block types, /root/reference/lib/common/zstd_internal.h:90 */
#define ZB_BT_RAW 0
#define ZB_BT_RLE 1
#define ZB_BT_COMPRESSED 2

#define ZB_FLAG_FIRST 1u       /* first block of its frame */
#define ZB_FLAG_LAST  2u       /* last block of its frame  */
#define ZB_FLAG_DICT  4u       /* its history (histLen bytes) is the tail of the call's dictionary content */

/* error codes = lib/zstd_errors.h:64-101 */
#define ZB_ERR(code) ((size_t)-(long)(code))
#define ZB_error_GENERIC 1
#define ZB_error_prefix_unknown 10
#define ZB_error_dictionary_corrupted 30
#define ZB_error_dictionary_wrong 32
#define ZB_error_parameter_unsupported 40
#define ZB_error_tableLog_tooLarge 44
#define ZB_error_maxSymbolValue_tooLarge 46
#define ZB_error_stage_wrong 60
#define ZB_error_memory_allocation 64
#define ZB_error_workSpace_tooSmall 66
#define ZB_error_dstSize_tooSmall 70
#define ZB_error_srcSize_wrong 72
#define ZB_error_dstBuffer_null 74
#define ZB_error_maxCode 120

typedef struct {
    u64 srcOff;        /* block start, byte offset into the input buffer */
    u32 size;          /* block size */
    u32 histLen;       /* bytes of history a match of this block may reach back into (chunk history and window) */
    u32 frame;         /* index into ZbFrame[] */
    u32 flags;         /* ZB_FLAG_* */
    u32 dictLen;       /* the oldest dictLen bytes of that history are the tail of the call's dictionary content (ZB_FLAG_DICT) */
    u32 pad;
} ZbBlock;

/* one chunk = up to ZB_CHUNK_BLOCKS consecutive blocks of a frame: the unit of the candidate walk */
typedef struct {
    u64 srcOff;        /* chunk start, byte offset into the input buffer */
    u32 size;          /* bytes in the chunk */
    u32 histLen;       /* bytes walked in front of the chunk to prime the table (<= ZB_PRIME_BYTES) */
    u32 dictLen;       /* the oldest dictLen of them are the dictionary content's tail (first chunk of a frame only) */
    u32 firstBlock;    /* index of the chunk's first block in the call's block array */
    u32 blockLog;      /* log2 of the frame's block size: block k of the chunk starts at k << blockLog */
    u32 pad;
} ZbChunk;

typedef struct {
    u64 srcOff;        /* frame input start */
    u64 srcSize;
    u32 firstBlock;    /* index of its first ZbBlock */
    u32 nbBlocks;
    u32 windowLog;
    u32 dictID;
    u32 checksum;      /* 1: Content_Checksum_flag set, 4 bytes are reserved behind the last block (filled in by the host) */
    u32 pad;
} ZbFrame;

typedef struct {       /* produced on the device, one per block */
    u32 nbSeq;
    u32 litSize;       /* all literals of the block incl. the trailing run */
    u32 litSecSize;    /* bytes of the literals section written at body[0..) */
    u32 bodySize;      /* final payload size (compressed body, 1 for RLE, block size for raw) */
    u32 type;          /* ZB_BT_* */
    u32 forceRaw;      /* block too small to try (zstd_compress.c:3216) or entropy stage gave up */
    u32 rleByte;
    u32 pad;
} ZbBlockMeta;

/* FSE compression table in our own layout (the reference's is common/fse.h:249) */
typedef struct {
    u32 tableLog;
    u32 maxSymbolValue;
    u16 nextState[512];
    int deltaFindState[64];
    u32 deltaNbBits[64];
} ZbdFseCTable;

/* entropy state a zstd-format dictionary installs as the "previous block" of a frame's first block
 * (ZSTD_loadCEntropy, /root/reference/lib/compress/zstd_compress.c:4987-5076); built on the host (zb_dict.cu) */
typedef struct {
    u32 present;
    u32 hufRepeat;             /* HUF_repeat: 0 none, 1 check, 2 valid */
    u32 hufMaxSymbol;
    u32 fseRepeat[3];          /* FSE_repeat per stream: 0 = LL, 1 = OF, 2 = ML */
    u32 rep[3];
    u32 dictID;
    u32 hufEnc[256];           /* code | nbBits << 16 */
    ZbdFseCTable fse[3];
} ZbDictEntropy;

typedef struct {
    u32 strategy;      /* 1 = fast, 2 = dfast */
    u32 mls;           /* bytes hashed by the (short) table: 4..8 */
    u32 tableN;        /* buckets of the (short) table; bucket = (hash32 * tableN) >> 32 */
    u32 tableNLong;    /* dfast: buckets of the 8-byte-hash table */
    u32 stepSize;      /* zstd_fast.c:200 */
    u32 litDisabled;   /* zstd_compress_internal.h:621-633 */
    u32 windowLog;
    u32 insStep;       /* positions without a candidate enter the table when ((pos - low) % step) < 2, step = insStep + walked / 128 */
    u32 startRep[2];   /* repcodes the search of a frame's first segment starts with (zstd-format dictionary), 0 = none */
    u32 codeRep[3];    /* repcode history the decoder holds at a frame's first block: {1,4,8} or the dictionary's */
} ZbParams;

/* Per-block strides of the workspace arrays of one call, derived from its largest block (M = that size
 * rounded up to 64): 128 KiB blocks use the ZB_*_STRIDE values above, a call of 1 KiB records 1/128 of them. */
/* fast strategy: a block is parsed in segments of ZB_PARSE_SEG bytes, one warp each (a segment behaves like a block
 * for the parse; candidates, literals and sequences stay the block's).  Per segment, for the merge kernel: */
#define ZB_PARSE_SEG   (16u << 10)
#define ZB_PARSE_SEGS  (ZB_BLOCK_MAX / ZB_PARSE_SEG)
typedef struct {
    u32 nbSeq;         /* raw sequences of the segment, stored from seq slot k * ZB_PARSE_SEG / 4 */
    u32 pad[3];
} ZbSegMeta;

typedef struct {
    u32 dist;          /* u16 per block : candidate distances (+ as many u32 of "far" distances); K3 reuses the u16 area for 3 x state records */
    u32 seq;           /* u64 per block : packed sequences */
    u32 lit;           /* bytes per block : literal bytes (multiple of 16) */
    u32 body;          /* bytes per block : compressed block body staging (multiple of 16) */
    u32 state;         /* u16 per FSE stream per block inside the dist area (3 * state <= dist) */
} ZbStrides;

#endif
