/* zb_api.cu — host driver + C ABI of libzstd_b200.so.
 *
 * Mirrors what the reference does around the per-block hot path:
 *   ZSTD_compress / ZSTD_compressCCtx / ZSTD_compress_usingDict  (/root/reference/lib/compress/zstd_compress.c:5398-5440)
 *   parameter derivation   ZSTD_getCParams_internal :7123-7146 + ZSTD_adjustCParams_internal :1465-1602 (compress/clevels.h:25-130)
 *   block planning         ZSTD_compress_frameChunk :4527-4623 (here: all blocks of a call at once)
 * and hands every block to the CUDA kernels (zb_match.cu, zb_literals.cu, zb_sequences.cu,
 * zb_stitch.cu).  No compression work is done on the host; without a CUDA device every compress
 * entry point fails with ZSTD_error_GENERIC.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <mutex>
#include <thread>
#include <new>
#include <time.h>
#include "../../include/zstd_b200.h"
#include "zb_common.h"
#include "zb_kernels.h"

static inline u32 hb32(u32 v) { return 31u - (u32)__builtin_clz(v); }
static inline bool zb_isErr(size_t c) { return c > ZB_ERR(ZB_error_maxCode); }

/* ------------------------------------------------------------------ parameters */
typedef struct { u32 windowLog, chainLog, hashLog, searchLog, minMatch, targetLength, strategy; } ZbCParams;
struct Row { u8 W, C, H, S, L, TL, strat; };
/* compress/clevels.h:25-130, rows 0..4; strategies above dfast are out of scope: a level whose
 * reference strategy is greedy or stronger is served by the strongest dfast row of its size class */
static const Row kRows[4][5] = {
    { {19,12,13,1,6,1,1}, {19,13,14,1,7,0,1}, {20,15,16,1,6,0,1}, {21,16,17,1,5,0,2}, {21,18,18,1,5,0,2} },
    { {18,12,13,1,5,1,1}, {18,13,14,1,6,0,1}, {18,14,14,1,5,0,2}, {18,16,16,1,4,0,2}, {18,16,16,1,4,0,2} },
    { {17,12,12,1,5,1,1}, {17,12,13,1,6,0,1}, {17,13,15,1,5,0,1}, {17,15,16,2,5,0,2}, {17,17,17,2,4,0,2} },
    { {14,12,13,1,5,1,1}, {14,14,15,1,5,0,1}, {14,14,15,1,4,0,1}, {14,14,15,2,4,0,2}, {14,14,15,2,4,0,2} },
};

static ZbCParams zb_getCParams(int level, u64 srcSize, size_t dictSize)
{
    u64 const rSize = srcSize + dictSize;
    u32 const tableID = (rSize <= 256u * 1024) + (rSize <= 128u * 1024) + (rSize <= 16u * 1024);
    int const r = level == 0 ? 3 : (level < 0 ? 0 : (level > 4 ? 4 : level));
    Row const x = kRows[tableID][r];
    ZbCParams cp = { x.W, x.C, x.H, x.S, x.L, x.TL, x.strat };
    if (level < 0) {
        int const minLevel = -(int)ZB_BLOCK_MAX;                 /* ZSTD_minCLevel, zstd_compress.c:7038 */
        cp.targetLength = (u32)(-(level < minLevel ? minLevel : level));
    }
    {   u64 const maxWindowResize = 1ull << 30;                  /* zstd_compress.c:1537-1547 */
        if (srcSize <= maxWindowResize && dictSize <= maxWindowResize) {
            u32 const tSize = (u32)(srcSize + dictSize);
            u32 const srcLog = (tSize < (1u << 6)) ? 6 : hb32(tSize - 1) + 1;
            if (cp.windowLog > srcLog) cp.windowLog = srcLog;
        }
        u32 dawl = cp.windowLog;                                  /* ZSTD_dictAndWindowLog :1432-1459 */
        if (dictSize) {
            u64 const windowSize = 1ull << cp.windowLog;
            u64 const dictAndWindowSize = dictSize + windowSize;
            if (windowSize >= dictSize + srcSize) dawl = cp.windowLog;
            else if (dictAndWindowSize >= (1ull << 31)) dawl = 31;
            else dawl = hb32((u32)dictAndWindowSize - 1) + 1;
        }
        if (cp.hashLog > dawl + 1) cp.hashLog = dawl + 1;
        if (cp.chainLog > dawl) cp.chainLog -= (cp.chainLog - dawl);
        if (cp.windowLog < 10) cp.windowLog = 10;                 /* ZSTD_WINDOWLOG_ABSOLUTEMIN */
    }
    return cp;
}

static ZbParams zb_makeParams(const ZbCParams& cp)
{
    ZbParams p; memset(&p, 0, sizeof(p));
    p.strategy = cp.strategy;
    p.windowLog = cp.windowLog;
    p.mls = cp.minMatch < 4 ? 4 : (cp.minMatch > 8 ? 8 : cp.minMatch);
    /* Table sizes and the insertion pattern are set against the reference's compressed size on datagen P30 / P50 / P90
     * (tools/exp_size.py): an occurrence stays in the table until a different string takes its
     * bucket (positions that find a candidate are not inserted), so a table somewhat smaller than the reference's holds
     * as many useful candidates.  The oracle's zbo_makePlan is the same rule (tests/test_plan.py). */
    if (cp.strategy == 1) {
        u32 const hl = cp.hashLog > ZB_FAST_HASHLOG_MAX ? ZB_FAST_HASHLOG_MAX : cp.hashLog;
        p.stepSize = cp.targetLength + !cp.targetLength + 1;     /* zstd_fast.c:200 */
        if (cp.targetLength == 0) { p.tableN = 3u << (hl - 2); p.insStep = 3; }
        else                      { p.tableN = 7u << (hl - 3); p.insStep = p.stepSize >= 5u ? p.stepSize - 1u : p.stepSize; }   /* never the parse's own probe spacing from 5 on: equal periods lock the probed positions out of phase with the inserted ones (level -7: +8.6 % instead of -6.6 % on datagen -P90) */
    } else {
        p.stepSize = 1;
        p.tableN = 1u << cp.chainLog;                            /* short table (zstd_double_fast.c:116) */
        if (p.tableN > ZB_DFAST_SHORT_MAX) p.tableN = ZB_DFAST_SHORT_MAX;
        p.tableNLong = 1u << (cp.hashLog > ZB_DFAST_LONGLOG_MAX ? ZB_DFAST_LONGLOG_MAX : cp.hashLog);
        p.insStep = 2;
    }
    p.codeRep[0] = 1; p.codeRep[1] = 4; p.codeRep[2] = 8;        /* zstd_internal.h:69 */
    p.litDisabled = (cp.strategy == 1) && (cp.targetLength > 0); /* zstd_compress_internal.h:621-633 */
    return p;
}

#define ZB_IMAGE_WORDS (ZB_DFAST_SHORT_MAX + (1u << ZB_DFAST_LONGLOG_MAX))   /* largest table pair: short table, then (doubleFast) the long table */
#define ZB_IMAGE_BYTES (ZB_IMAGE_WORDS * 4u)
#define ZB_MAX_IMAGES 4

/* ------------------------------------------------------------------ context */
#include <atomic>
static std::atomic<int> g_device(-1);

/* restores the calling thread's current device when a call returns (a context works on the device it was created for) */
struct ZbDeviceGuard {
    int prev;
    ZbDeviceGuard() : prev(-1) { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~ZbDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define ZB_WAVE_SLOTS_MAX 14u
#define ZB_WAVE_SLOTS_DEFAULT 4u
#define ZB_HOST_WAVE_SLOTS_DEFAULT 8u   /* measured best with 384-block waves (tests/e2e_sweep.py, profiles/r1_e2e_timeline.md) */
#define ZB_HOST_WAVE_BLOCKS 384u     /* 48 MiB of input per wave */
/* Digested dictionary (lib/zstd.h:979, zstd_compress.c:5477-5642): the content tail, its entropy tables and
 * the primed hash-table images live on the device across calls; any number of contexts may use it. */
struct ZSTD_CDict_s {
    int level;
    u8* content;                   /* host copy of the whole dictionary (ZSTD_dlm_byCopy) */
    size_t size;
    size_t contentOff, tail;       /* entropy header size, bytes of content that blocks can see */
    ZbDictEntropy entropy;         /* parsed on the host at creation */
    std::mutex* lock;              /* guards the lazily created device state below */
    int device;                    /* -1 until first use */
    u8* d_dict; ZbDictEntropy* d_de; u8* d_image; ZbChunk* d_dictChunk;
    u32 nbImages; ZbParams imagePrm[ZB_MAX_IMAGES];
};

struct ZbPlan;
static void zb_freePlan(struct ZbPlan* p);       /* defined behind ZbPlan */
struct ZSTD_CCtx_s {
    int device;                    /* -1 until the first call created the stream and events on bindDevice */
    int bindDevice;                /* device captured by ZSTD_createCCtx */
    cudaStream_t stream;
    /* per-block workspace */
    size_t capBlocks, capFrames, capWaves;
    size_t capHeavyBytes[9];       /* meta, seqs, lits, body, dist, dist2, segmeta, far, far2 */
    size_t capChunks; ZbChunk* d_chunks;
    ZbSegMeta* d_segmeta;          /* K1b -> K1c: per parse segment counts */
    u32 devWaveBlocks;             /* device-memory calls: blocks per wave (0 = always one wave) */
    cudaStream_t waveStream[ZB_WAVE_SLOTS_MAX + 2];
    u32 waveSlots;                 /* waves in flight, device-memory calls */
    u32 hostWaveSlots;             /* waves in flight, host-memory calls */
    u32 hostWaveBlocks;            /* host-memory calls: blocks per wave */
    ZbBlock* d_blocks; ZbFrame* d_frames; ZbBlockMeta* d_meta;
    u64* d_seqs; u8* d_lits; u8* d_body; u16* d_dist;   /* d_dist: K1a->K1b candidate distances, then K3's FSE state records */
    u16* d_dist2;                                        /* dfast only: short-hash candidate distances */
    u32* d_far; u32* d_far2;                             /* candidate distances >= 0xFFFF (dist16 = ZB_FAR) */
    u64* d_outOffsets; u64* d_frameSizes; u64* d_totals;    /* d_totals[w]: bytes produced up to and including wave w */
    /* host-pointer path staging */
    u8* d_in; size_t d_inCap; u8* d_out; size_t d_outCap;
    u8* d_dict;                    /* dictionary content tail (<= ZB_PRIME_BYTES), 32 bytes of padding on both sides */
    u8* d_image;                   /* tables primed from the dictionary tail, one per parameter group (ZB_MAX_IMAGES) */
    ZbChunk* d_dictChunk;          /* pseudo chunk descriptors for zb_launch_dict_image */
    ZbDictEntropy dictEntropy;     /* host copy of the current call's dictionary entropy state */
    ZbDictEntropy* d_de;           /* device copy */
    const ZbDictEntropy* d_deActive; /* d_de when the current call's dictionary is zstd-format, else NULL */
    u64* h_totals;                 /* pinned mirror of d_totals */
    cudaEvent_t evStart, evK0, evMid, evK1, evK2, evK3, evKEnd, evEnd;
    ZSTDB200_stats stats;
    /* advanced one-shot API (ZSTD_CCtx_setParameter + ZSTD_compress2, lib/zstd.h:337-603): sticky parameters */
    int advLevel, advChecksum, advNoDictID;
    ZSTD_CDict* advLocalDict;      /* ZSTD_CCtx_loadDictionary: owned copy, digested at its first use */
    const ZSTD_CDict* advRefCDict; /* ZSTD_CCtx_refCDict: borrowed */
    /* per-call frame options, consumed by the planner */
    u32 callChecksum, callNoDictID;
    u64 callPartBegin, callPartEnd;   /* ZSTDB200_compressFramePart: this call's share of the frame (0, 0 = all of it) */
    struct ZbPlan* plan;           /* the call's plan; its vectors are reused (a million records are 100 MB of descriptors: fresh pages cost more than filling them) */
    /* streaming front end (ZSTD_compressStream2 with ZSTD_e_continue / ZSTD_e_flush): input collected on the host, compressed
     * output waiting to be handed out */
    u8* stIn; size_t stInSize, stInCap;
    u8* stOut; size_t stOutSize, stOutPos, stOutCap;
    int stFrames;                  /* frames produced in the current session */
};

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    if (getenv("ZSTDB200_DEBUG")) fprintf(stderr, "zstd_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
    return ZB_ERR(e_ == cudaErrorMemoryAllocation ? ZB_error_memory_allocation : ZB_error_GENERIC); } } while (0)

static double zb_now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

extern "C" int ZSTDB200_setDevice(int device) { g_device.store(device); return 0; }
extern "C" int zb_boundDevice(void) { return g_device.load(); }          /* for the decompression contexts (zb_decode.cu) */
extern "C" int ZSTDB200_deviceAvailable(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n > 0;
}

extern "C" ZSTD_CCtx* ZSTD_createCCtx(void)
{
    ZSTD_CCtx* c = (ZSTD_CCtx*)calloc(1, sizeof(ZSTD_CCtx));
    if (!c) return NULL;
    c->device = -1;
    /* the context belongs to the device that is selected NOW (ZSTDB200_setDevice, else the thread's current device) */
    c->bindDevice = g_device.load();
    if (c->bindDevice < 0) { int d = -1; if (cudaGetDevice(&d) == cudaSuccess) c->bindDevice = d; else cudaGetLastError(); }
    c->advLevel = 3;                                                         /* ZSTD_CLEVEL_DEFAULT */
    {   const char* s = getenv("ZSTDB200_SERIAL"); const char* w = getenv("ZSTDB200_WAVE_BLOCKS");
        c->devWaveBlocks = (s && atoi(s)) ? 0u : (w ? (u32)atoi(w) : 1024u);     /* 128 MiB waves: tests/wave_sweep.py */
        const char* n = getenv("ZSTDB200_WAVE_SLOTS"); const char* h = getenv("ZSTDB200_HOST_WAVE_BLOCKS");
        c->waveSlots = n ? (u32)atoi(n) : ZB_WAVE_SLOTS_DEFAULT;
        c->hostWaveSlots = n ? (u32)atoi(n) : ZB_HOST_WAVE_SLOTS_DEFAULT;
        if (c->waveSlots < 1u) c->waveSlots = 1u;
        if (c->waveSlots > ZB_WAVE_SLOTS_MAX) c->waveSlots = ZB_WAVE_SLOTS_MAX;
        if (c->hostWaveSlots < 1u) c->hostWaveSlots = 1u;
        if (c->hostWaveSlots > ZB_WAVE_SLOTS_MAX) c->hostWaveSlots = ZB_WAVE_SLOTS_MAX;
        c->hostWaveBlocks = (h && atoi(h) > 0) ? (u32)atoi(h) : ZB_HOST_WAVE_BLOCKS; }
    return c;
}

static size_t zb_ctxInit(ZSTD_CCtx* c)
{
    if (c->device >= 0) { CK(cudaSetDevice(c->device)); return 0; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return ZB_ERR(ZB_error_GENERIC); }
    int dev = c->bindDevice;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
    CK(cudaSetDevice(dev));
    /* everything or nothing: a partial failure leaves the context uninitialised (device stays -1) */
    cudaStream_t st = nullptr; cudaEvent_t ev[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; i++) e = cudaEventCreate(&ev[i]);
    if (e == cudaSuccess) {
        /* the predefined FSE tables live in device memory (one copy per device; re-uploading the same bytes is harmless) */
        static ZbdFseCTable defaults[3]; static std::once_flag once;
        std::call_once(once, [] { zb_buildDefaultTables(defaults); });
        e = zb_upload_default_tables(defaults, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (e != cudaSuccess) {
        for (int i = 0; i < 8; i++) if (ev[i]) cudaEventDestroy(ev[i]);
        if (st) cudaStreamDestroy(st);
        cudaGetLastError();
        return ZB_ERR(e == cudaErrorMemoryAllocation ? ZB_error_memory_allocation : ZB_error_GENERIC);
    }
    c->stream = st;
    c->evStart = ev[0]; c->evK0 = ev[1]; c->evK1 = ev[2]; c->evK2 = ev[3]; c->evK3 = ev[4]; c->evMid = ev[5]; c->evKEnd = ev[6]; c->evEnd = ev[7];
    c->device = dev;
    return 0;
}

static void zb_freeWorkspace(ZSTD_CCtx* c)
{
    cudaFree(c->d_blocks); cudaFree(c->d_frames); cudaFree(c->d_meta); cudaFree(c->d_seqs); cudaFree(c->d_lits);
    cudaFree(c->d_body); cudaFree(c->d_dist); cudaFree(c->d_dist2); c->d_dist2 = NULL; cudaFree(c->d_segmeta); c->d_segmeta = NULL;
    cudaFree(c->d_far); cudaFree(c->d_far2); c->d_far = NULL; c->d_far2 = NULL; cudaFree(c->d_chunks); c->d_chunks = NULL; c->capChunks = 0; cudaFree(c->d_outOffsets); cudaFree(c->d_frameSizes); cudaFree(c->d_totals);
    cudaFreeHost(c->h_totals);
    c->d_blocks = NULL; c->d_frames = NULL; c->d_meta = NULL; c->d_seqs = NULL; c->d_lits = NULL;
    c->d_body = NULL; c->d_dist = NULL; c->d_outOffsets = NULL; c->d_frameSizes = NULL; c->d_totals = NULL; c->h_totals = NULL;
    c->capBlocks = 0; c->capFrames = 0; c->capWaves = 0; memset(c->capHeavyBytes, 0, sizeof(c->capHeavyBytes));
}

extern "C" size_t ZSTD_freeCDict(ZSTD_CDict* cd);
extern "C" size_t ZSTD_freeCCtx(ZSTD_CCtx* c)
{
    if (!c) return 0;
    ZbDeviceGuard guard;
    ZSTD_freeCDict(c->advLocalDict);
    free(c->stIn); free(c->stOut);
    zb_freePlan(c->plan);
    if (c->device >= 0) {
        cudaSetDevice(c->device);
        zb_freeWorkspace(c);
        cudaFree(c->d_in); cudaFree(c->d_out); cudaFree(c->d_dict); cudaFree(c->d_de); cudaFree(c->d_image); cudaFree(c->d_dictChunk);
        for (u32 s = 0; s < ZB_WAVE_SLOTS_MAX + 2u; s++) if (c->waveStream[s]) cudaStreamDestroy(c->waveStream[s]);
        cudaEventDestroy(c->evStart); cudaEventDestroy(c->evK0); cudaEventDestroy(c->evK1);
        cudaEventDestroy(c->evK2); cudaEventDestroy(c->evK3); cudaEventDestroy(c->evMid);
        cudaEventDestroy(c->evKEnd); cudaEventDestroy(c->evEnd);
        cudaStreamDestroy(c->stream);
    }
    free(c);
    return 0;
}

/* descriptors (per block / per frame, small) and the heavy per-block workspace are sized separately:
 * the host-pointer path runs the blocks in waves that share a few workspace slots */
static size_t zb_ensureDesc(ZSTD_CCtx* c, size_t nbBlocks, size_t nbFrames, size_t nbWaves, size_t nbChunks)
{
    if (nbChunks > c->capChunks) {
        cudaFree(c->d_chunks); c->d_chunks = NULL; c->capChunks = 0;
        CK(cudaMalloc(&c->d_chunks, nbChunks * sizeof(ZbChunk)));
        c->capChunks = nbChunks;
    }
    if (nbBlocks > c->capBlocks) {
        cudaFree(c->d_blocks); cudaFree(c->d_outOffsets); c->d_blocks = NULL; c->d_outOffsets = NULL; c->capBlocks = 0;
        CK(cudaMalloc(&c->d_blocks, nbBlocks * sizeof(ZbBlock)));
        CK(cudaMalloc(&c->d_outOffsets, (nbBlocks + 1) * sizeof(u64)));
        c->capBlocks = nbBlocks;
    }
    if (nbFrames > c->capFrames) {
        cudaFree(c->d_frames); cudaFree(c->d_frameSizes); c->d_frames = NULL; c->d_frameSizes = NULL; c->capFrames = 0;
        CK(cudaMalloc(&c->d_frames, nbFrames * sizeof(ZbFrame)));
        CK(cudaMalloc(&c->d_frameSizes, nbFrames * sizeof(u64)));
        c->capFrames = nbFrames;
    }
    if (nbWaves > c->capWaves) {
        cudaFree(c->d_totals); cudaFreeHost(c->h_totals); c->d_totals = NULL; c->h_totals = NULL; c->capWaves = 0;
        CK(cudaMalloc(&c->d_totals, nbWaves * sizeof(u64)));
        CK(cudaMallocHost(&c->h_totals, nbWaves * sizeof(u64)));
        c->capWaves = nbWaves;
    }
    return 0;
}
static ZbStrides zb_strides(u32 maxBlock)
{
    u32 const M = ((maxBlock < 64u ? 64u : maxBlock) + 63u) & ~63u;
    ZbStrides sd; sd.dist = M; sd.seq = M / 4u + 8u; sd.lit = M + 256u; sd.body = M + 1024u; sd.state = M / 4u;
    return sd;
}
/* workspace for nbSlotBlocks blocks laid out with the strides `sd` (capacities are kept in bytes) */
static size_t zb_ensureHeavy(ZSTD_CCtx* c, size_t nbSlotBlocks, const ZbStrides& sd, bool needDist2)
{
    size_t const nb = nbSlotBlocks;
    size_t const need[9] = { nb * sizeof(ZbBlockMeta), nb * sd.seq * sizeof(u64), nb * (size_t)sd.lit, nb * (size_t)sd.body,
                             nb * (size_t)sd.dist * sizeof(u16), needDist2 ? nb * (size_t)sd.dist * sizeof(u16) : 0,
                             nb * ((sd.dist + ZB_PARSE_SEG - 1u) / ZB_PARSE_SEG) * sizeof(ZbSegMeta),
                             nb * (size_t)sd.dist * sizeof(u32), needDist2 ? nb * (size_t)sd.dist * sizeof(u32) : 0 };
    void** const ptr[9] = { (void**)&c->d_meta, (void**)&c->d_seqs, (void**)&c->d_lits, (void**)&c->d_body, (void**)&c->d_dist, (void**)&c->d_dist2,
                            (void**)&c->d_segmeta, (void**)&c->d_far, (void**)&c->d_far2 };
    for (int i = 0; i < 9; i++) {
        if (need[i] <= c->capHeavyBytes[i]) continue;
        cudaFree(*ptr[i]); *ptr[i] = NULL; c->capHeavyBytes[i] = 0;
        CK(cudaMalloc(ptr[i], need[i] + 256));
        c->capHeavyBytes[i] = need[i];
    }
    return 0;
}

/* ------------------------------------------------------------------ dictionaries (zstd_compress.c:5119-5156)
 * < 8 bytes: ignored (:5132).  No magic 0xEC30A437: raw content (:5143-5148).  With it: zstd-format
 * dictionary = magic, dictID, Huffman table, OF/ML/LL FSE tables, 3 repcodes, content (:4987-5076).
 * The CONTENT of either kind is the history of each frame's first block; a zstd-format dictionary's Huffman /
 * FSE tables are that block's "previous" entropy state (treeless literals, set_repeat sequence tables) and
 * its repcodes start the block.  Parsing: zb_dict.cu. */
/* ------------------------------------------------------------------ planning (ZSTD_compress_frameChunk, zstd_compress.c:4527) */
struct ZbGroup { ZbParams prm; u32 b0, b1, c0, c1; const u32* image; };   /* image: tables walked over the dictionary tail, or NULL */
/* Descriptor arrays of a plan: page-locked (so that the upload of a million frames' descriptors is a DMA at PCIe speed, not a
 * staged copy of pageable memory) and kept across calls.  Without a CUDA device (the CPU tests' ZSTDB200_describePlan)
 * ordinary memory is used. */
template <typename T> struct ZbVec {
    T* p; size_t n, cap; bool pinned;
    ZbVec() : p(NULL), n(0), cap(0), pinned(false) {}
    ~ZbVec() { release(); }
    ZbVec(const ZbVec&) = delete; ZbVec& operator=(const ZbVec&) = delete;
    void release() { if (p) { if (pinned) cudaFreeHost(p); else free(p); } p = NULL; n = cap = 0; }
    void reserve(size_t want) {
        if (want <= cap) return;
        size_t const c = want < 2 * cap ? 2 * cap : want;
        T* q = NULL; bool pin = true;
        if (cudaMallocHost((void**)&q, c * sizeof(T)) != cudaSuccess) { cudaGetLastError(); pin = false; q = (T*)malloc(c * sizeof(T)); }
        if (n) memcpy(q, p, n * sizeof(T));
        size_t const keep = n;
        release(); p = q; n = keep; cap = c; pinned = pin;
    }
    void push_back(const T& v) { if (n == cap) reserve(cap ? 2 * cap : 64); p[n++] = v; }
    void resize(size_t m) { reserve(m); if (m > n) memset((void*)(p + n), 0, (m - n) * sizeof(T)); n = m; }
    void clear() { n = 0; }
    size_t size() const { return n; }
    T* data() { return p; }
    const T* data() const { return p; }
    T& back() { return p[n - 1]; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
};
struct ZbPlan { ZbVec<ZbBlock> blocks; ZbVec<ZbChunk> chunks; ZbVec<ZbFrame> frames; std::vector<ZbGroup> groups; ZbStrides sd; bool unsupported;
                void reset() { blocks.clear(); chunks.clear(); frames.clear(); groups.clear(); unsupported = false; } };   /* keeps its memory: a context plans call after call */

static void zb_freePlan(ZbPlan* p) { delete p; }
static int g_strictLevels = 0;
/* Levels whose reference strategy is greedy or stronger (>= 5; 4 for frames <= 256 KiB / <= 16 KiB) have no counterpart here:
 * by default they are served by the strongest doubleFast row of their size class (larger output than the reference's at
 * that level); after ZSTDB200_setStrictLevels(1) such calls fail with parameter_unsupported instead. */
extern "C" void ZSTDB200_setStrictLevels(int on) { g_strictLevels = on; }

/* partBegin / partEnd: only the blocks that start inside [partBegin, partEnd) of the (single) frame are planned — one rank's
 * share of a frame that several GPUs compress together (ZSTDB200_compressFramePart); the geometry is the whole frame's. */
static void zb_plan(ZbPlan& P, u32 frameChecksum, const size_t* frameOffsets, const size_t* frameSizes, size_t nbFrames, int level,
                    size_t dictSize, size_t dictTail, u32 dictID, const u32* dictRep, u64 partBegin = 0, u64 partEnd = ~0ull)
{
    P.frames.resize(nbFrames);
    P.blocks.reserve(nbFrames);
    P.chunks.reserve(nbFrames);
    P.unsupported = g_strictLevels && level > 4;
    u32 maxBlock = 0;
    /* calls made of many equal, single-block frames (config 5: a million 1 KiB records): the frame before is the template */
    u64 tplSize = ~0ull; ZbFrame tplFrame; ZbBlock tplBlock; ZbChunk tplChunk;
    memset(&tplFrame, 0, sizeof(tplFrame)); memset(&tplBlock, 0, sizeof(tplBlock)); memset(&tplChunk, 0, sizeof(tplChunk));
    for (size_t f = 0; f < nbFrames; f++) {
        u64 const fsz = frameSizes[f];
        if (fsz == tplSize) {
            ZbFrame fr = tplFrame; fr.srcOff = frameOffsets[f]; fr.firstBlock = (u32)P.blocks.size();
            ZbBlock b = tplBlock; b.srcOff = fr.srcOff; b.frame = (u32)f;
            ZbChunk ch = tplChunk; ch.srcOff = fr.srcOff; ch.firstBlock = fr.firstBlock;
            P.blocks.push_back(b);
            P.chunks.push_back(ch);
            P.frames[f] = fr;
            P.groups.back().b1 = (u32)P.blocks.size();
            P.groups.back().c1 = (u32)P.chunks.size();
            continue;
        }
        ZbCParams cp = zb_getCParams(level, fsz, dictSize);
        ZbParams prm = zb_makeParams(cp);
        if (dictRep) {                                           /* a zstd-format dictionary's repcodes (zstd_compress.c:5054-5056) */
            prm.codeRep[0] = dictRep[0]; prm.codeRep[1] = dictRep[1]; prm.codeRep[2] = dictRep[2];
            prm.startRep[0] = dictRep[0] <= dictTail ? dictRep[0] : 0u; prm.startRep[1] = dictRep[1] <= dictTail ? dictRep[1] : 0u;
        }
        size_t const blockMax = ((size_t)1 << cp.windowLog) < ZB_BLOCK_MAX ? ((size_t)1 << cp.windowLog) : ZB_BLOCK_MAX;   /* zstd_compress.c:2124 */
        u64 const chunkBytes = (u64)ZB_CHUNK_BLOCKS * blockMax;
        u64 const W = 1ull << cp.windowLog;
        ZbFrame fr; fr.srcOff = frameOffsets[f]; fr.srcSize = fsz; fr.firstBlock = (u32)P.blocks.size();
        fr.windowLog = cp.windowLog; fr.dictID = dictID; fr.checksum = frameChecksum; fr.pad = 0;
        u32 const firstChunk = (u32)P.chunks.size();
        u64 pos = 0;
        do {
            u64 const bsz = (fsz - pos) < blockMax ? (fsz - pos) : blockMax;
            if (pos < partBegin || pos >= partEnd) { pos += bsz; continue; }      /* another rank's block */
            if (pos % chunkBytes == 0) {                         /* a new chunk starts with this block */
                ZbChunk ch; memset(&ch, 0, sizeof(ch));
                ch.srcOff = fr.srcOff + pos; ch.size = (u32)((fsz - pos) < chunkBytes ? (fsz - pos) : chunkBytes);
                ch.histLen = pos == 0 ? (u32)dictTail : (u32)(pos < ZB_PRIME_BYTES ? pos : ZB_PRIME_BYTES);
                ch.dictLen = pos == 0 ? (u32)dictTail : 0u;
                ch.firstBlock = (u32)P.blocks.size(); ch.blockLog = hb32((u32)blockMax);
                P.chunks.push_back(ch);
            }
            ZbChunk const& ch = P.chunks.back();
            /* positions in [dictionary tail | frame] coordinates: the frame starts at dictTail */
            u64 const chunkPos = ch.srcOff - fr.srcOff;
            u64 const chunkLow = dictTail + chunkPos - ch.histLen;
            u64 const bsBuf = dictTail + pos, beBuf = bsBuf + bsz;
            u64 low = chunkLow;
            if (beBuf > W && beBuf - W > low) low = beBuf - W;   /* ZSTD_window_enforceMaxDist at the block's end, zstd_compress_internal.h:1173 */
            ZbBlock b; memset(&b, 0, sizeof(b));
            b.srcOff = fr.srcOff + pos; b.size = (u32)bsz;
            b.histLen = (u32)(bsBuf - low);
            b.dictLen = low < dictTail ? (u32)(dictTail - low) : 0u;
            b.frame = (u32)f; b.flags = (pos == 0 ? ZB_FLAG_FIRST : 0u) | (pos + bsz == fsz ? ZB_FLAG_LAST : 0u) | (b.dictLen ? ZB_FLAG_DICT : 0u);
            P.blocks.push_back(b);
            if (b.size > maxBlock) maxBlock = b.size;
            pos += bsz;
        } while (pos < fsz);
        fr.nbBlocks = (u32)P.blocks.size() - fr.firstBlock;
        P.frames[f] = fr;
        if (fr.nbBlocks == 1u) { tplSize = fsz; tplFrame = fr; tplBlock = P.blocks.back(); tplChunk = P.chunks.back(); } else tplSize = ~0ull;
        if (P.groups.empty() || memcmp(&P.groups.back().prm, &prm, sizeof(prm)) != 0) {
            ZbGroup g; g.prm = prm; g.b0 = fr.firstBlock; g.b1 = (u32)P.blocks.size(); g.c0 = firstChunk; g.c1 = (u32)P.chunks.size(); g.image = NULL; P.groups.push_back(g);
        } else { P.groups.back().b1 = (u32)P.blocks.size(); P.groups.back().c1 = (u32)P.chunks.size(); }
    }
    P.sd = zb_strides(maxBlock);
}

/* parse + upload the dictionary content tail; returns 0 or an error.  *d_dictEnd = NULL when no dictionary applies */
/* device buffers of one dictionary: content tail with 32 bytes of padding on both sides, entropy state,
 * table images, pseudo block descriptors */
static size_t zb_allocDictBuffers(u8** d_dict, ZbDictEntropy** d_de, u8** d_image, ZbChunk** d_dictChunk)
{
    if (*d_dict) return 0;
    CK(cudaMalloc(d_dict, ZB_PRIME_BYTES + 64)); CK(cudaMalloc(d_de, sizeof(ZbDictEntropy)));
    CK(cudaMalloc(d_image, (size_t)ZB_IMAGE_BYTES * ZB_MAX_IMAGES)); CK(cudaMalloc(d_dictChunk, ZB_MAX_IMAGES * sizeof(ZbChunk)));
    return 0;
}
static size_t zb_uploadDict(u8* d_dict, ZbDictEntropy* d_de, const ZbDictEntropy* de, const u8* content, size_t contentSize, size_t tail, cudaStream_t stream)
{
    CK(cudaMemsetAsync(d_dict, 0, ZB_PRIME_BYTES + 64, stream));
    if (tail) CK(cudaMemcpyAsync(d_dict + 32, content + (contentSize - tail), tail, cudaMemcpyHostToDevice, stream));
    if (de->present) CK(cudaMemcpyAsync(d_de, de, sizeof(ZbDictEntropy), cudaMemcpyHostToDevice, stream));
    return 0;
}

/* Makes the call's dictionary resident: either the caller's raw bytes (parsed and uploaded now, into the
 * context's buffers) or a digested ZSTD_CDict (uploaded on its first use, then only referenced). */
static size_t zb_prepareDict(ZSTD_CCtx* c, const void* dict, size_t dictSize, const ZSTD_CDict* cdictC, cudaStream_t stream,
                             size_t* effDictSize, size_t* dictTail, u32* dictID, const u8** d_dictEnd)
{
    *effDictSize = 0; *dictTail = 0; *dictID = 0; *d_dictEnd = NULL;
    c->dictEntropy.present = 0; c->d_deActive = NULL;
    if (cdictC) {
        ZSTD_CDict* const cd = const_cast<ZSTD_CDict*>(cdictC);      /* the lazily created device state is guarded by cd->lock */
        if (cd->size < 8) return 0;                                  /* zstd_compress.c:5130 : tiny dictionaries are ignored */
        std::lock_guard<std::mutex> g(*cd->lock);
        if (cd->device >= 0 && cd->device != c->device) return ZB_ERR(ZB_error_parameter_unsupported);   /* one device per CDict */
        if (cd->device < 0) {
            {   size_t const e = zb_allocDictBuffers(&cd->d_dict, &cd->d_de, &cd->d_image, &cd->d_dictChunk); if (zb_isErr(e)) return e; }
            {   size_t const e = zb_uploadDict(cd->d_dict, cd->d_de, &cd->entropy, cd->content + cd->contentOff, cd->size - cd->contentOff, cd->tail, stream); if (zb_isErr(e)) return e; }
            CK(cudaStreamSynchronize(stream));                       /* other contexts (other streams) may use it right away */
            cd->device = c->device;
        }
        c->dictEntropy = cd->entropy;
        *dictID = cd->entropy.present ? cd->entropy.dictID : 0u;
        *effDictSize = cd->size; *dictTail = cd->tail;
        *d_dictEnd = cd->d_dict + 32 + cd->tail;
        if (cd->entropy.present) c->d_deActive = cd->d_de;
        return 0;
    }
    if (!dict || dictSize < 8) return 0;
    size_t const contentOff = zb_loadDictionary(&c->dictEntropy, (const u8*)dict, dictSize);
    if (zb_isErr(contentOff)) return contentOff;
    *dictID = c->dictEntropy.present ? c->dictEntropy.dictID : 0u;
    size_t const contentSize = dictSize - contentOff;
    size_t const tail = contentSize < ZB_PRIME_BYTES ? contentSize : ZB_PRIME_BYTES;
    *effDictSize = dictSize; *dictTail = tail;
    {   size_t const e = zb_allocDictBuffers(&c->d_dict, &c->d_de, &c->d_image, &c->d_dictChunk); if (zb_isErr(e)) return e; }
    {   size_t const e = zb_uploadDict(c->d_dict, c->d_de, &c->dictEntropy, (const u8*)dict + contentOff, contentSize, tail, stream); if (zb_isErr(e)) return e; }
    *d_dictEnd = c->d_dict + 32 + tail;
    if (c->dictEntropy.present) c->d_deActive = c->d_de;
    return 0;
}

/* One table image per parameter group (many small frames share one dictionary: priming its tail once
 * instead of once per frame is what the reference's CDict does on the CPU, zstd_compress.c:5477).
 * A ZSTD_CDict keeps its images across calls. */
static size_t zb_buildDictImages(ZSTD_CCtx* c, ZbPlan& P, const ZSTD_CDict* cdictC, const u8* d_dictEnd, size_t dictTail, cudaStream_t stream)
{
    if (!d_dictEnd || dictTail < 8) return 0;
    ZSTD_CDict* const cd = const_cast<ZSTD_CDict*>(cdictC);
    if (!cd && P.groups.size() > ZB_MAX_IMAGES) return 0;
    u8* const images = cd ? cd->d_image : c->d_image;
    ZbChunk* const dchk = cd ? cd->d_dictChunk : c->d_dictChunk;
    std::unique_lock<std::mutex> g;
    if (cd) g = std::unique_lock<std::mutex>(*cd->lock);
    u32 next = cd ? cd->nbImages : 0u;
    bool built = false;
    for (size_t gi = 0; gi < P.groups.size(); gi++) {
        ZbParams const& prm = P.groups[gi].prm;
        u32 slot = ~0u;
        if (cd) for (u32 i = 0; i < cd->nbImages; i++) if (memcmp(&cd->imagePrm[i], &prm, sizeof(prm)) == 0) slot = i;
        if (slot == ~0u) {
            if (next >= ZB_MAX_IMAGES) continue;                         /* no room: this group walks the dictionary per frame */
            slot = next++;
            ZbChunk ch; memset(&ch, 0, sizeof(ch));
            ch.histLen = (u32)dictTail; ch.dictLen = (u32)dictTail; ch.size = 0; ch.blockLog = 17;
            CK(cudaMemcpyAsync(dchk + slot, &ch, sizeof(ZbChunk), cudaMemcpyHostToDevice, stream));      /* pageable source: staged before the call returns */
            u32* const img = (u32*)(images + (size_t)slot * ZB_IMAGE_BYTES);
            CK(zb_launch_dict_image(d_dictEnd, dchk + slot, &prm, img, stream));
            if (cd) { cd->imagePrm[slot] = prm; cd->nbImages = next; built = true; }
        }
        P.groups[gi].image = (const u32*)(images + (size_t)slot * ZB_IMAGE_BYTES);
    }
    if (built) CK(cudaStreamSynchronize(stream));                        /* a cached image must be complete before another context reads it */
    return 0;
}

/* K1..K3 for blocks [b0, b1) = chunks [c0, c1) using workspace slot positions [slot0, slot0 + (b1-b0)) */
static size_t zb_runBlocks(ZSTD_CCtx* c, const ZbPlan& P, const u8* d_src, const u8* d_dictEnd, u32 b0, u32 b1, u32 c0, u32 c1, size_t slot0,
                           cudaStream_t stream, bool timed, unsigned* launches)
{
    for (int phase = 0; phase < 3; phase++) {
        for (size_t g = 0; g < P.groups.size(); g++) {
            ZbGroup const& G = P.groups[g];
            u32 const lo = G.b0 > b0 ? G.b0 : b0, hi = G.b1 < b1 ? G.b1 : b1;
            u32 const clo = G.c0 > c0 ? G.c0 : c0, chi = G.c1 < c1 ? G.c1 : c1;
            if (lo >= hi) continue;
            size_t const s = slot0 + (lo - b0);
            if (phase == 0) {
                bool const df = G.prm.strategy == 2;
                CK(zb_launch_match(d_src, d_dictEnd, d_dictEnd ? G.image : (const u32*)0, c->d_blocks + lo, hi - lo, c->d_chunks + clo, chi - clo, lo, &G.prm, &P.sd,
                                   c->d_dist + s * P.sd.dist, c->d_far + s * P.sd.dist, df ? c->d_dist2 + s * P.sd.dist : (u16*)0, df ? c->d_far2 + s * P.sd.dist : (u32*)0,
                                   c->d_seqs + s * P.sd.seq, c->d_lits + s * P.sd.lit, c->d_meta + s, c->d_segmeta + s * ((P.sd.dist + ZB_PARSE_SEG - 1u) / ZB_PARSE_SEG),
                                   (timed && P.groups.size() == 1) ? c->evMid : (cudaEvent_t)0, stream));
                *launches += df ? 4 : 3;       /* walk(s), parse, merge */
            } else if (phase == 1) {
                CK(zb_launch_literals(c->d_blocks + lo, hi - lo, &G.prm, &P.sd, c->d_deActive, c->d_lits + s * P.sd.lit, c->d_body + s * P.sd.body, c->d_meta + s, stream));
                *launches += 1;
            } else {
                CK(zb_launch_sequences(d_src, c->d_blocks + lo, hi - lo, &G.prm, &P.sd, c->d_deActive, c->d_seqs + s * P.sd.seq, c->d_dist + s * P.sd.dist,
                                       c->d_body + s * P.sd.body, c->d_meta + s, stream));
                *launches += 1;
            }
        }
        if (timed) CK(cudaEventRecord(phase == 0 ? c->evK1 : (phase == 1 ? c->evK2 : c->evK3), stream));
    }
    return 0;
}

/* ------------------------------------------------------------------ core: frames already in device memory (one wave) */
static size_t zb_compressFramesDevice(ZSTD_CCtx* c, u8* d_dst, size_t dstCapacity, const u8* d_src,
                                      const size_t* frameOffsets, const size_t* frameSizes, size_t nbFrames,
                                      const void* dict, size_t dictSize, const ZSTD_CDict* cdict, size_t* cSizes, int level, cudaStream_t stream)
{
    size_t effDict = 0, dictTail = 0; u32 dictID = 0; const u8* d_dictEnd = NULL;
    {   size_t const e = zb_prepareDict(c, dict, dictSize, cdict, stream, &effDict, &dictTail, &dictID, &d_dictEnd); if (zb_isErr(e)) return e; }
    if (!c->plan) { c->plan = new (std::nothrow) ZbPlan(); if (!c->plan) return ZB_ERR(ZB_error_memory_allocation); }
    ZbPlan& P = *c->plan; P.reset();
    zb_plan(P, c->callChecksum, frameOffsets, frameSizes, nbFrames, level, effDict, dictTail, c->callNoDictID ? 0u : dictID, c->dictEntropy.present ? c->dictEntropy.rep : NULL,
            c->callPartBegin, c->callPartEnd ? c->callPartEnd : ~0ull);
    if (P.unsupported) return ZB_ERR(ZB_error_parameter_unsupported);
    u32 const nbBlocks = (u32)P.blocks.size();
    {   size_t e = zb_ensureDesc(c, nbBlocks, nbFrames, 1, P.chunks.size()); if (zb_isErr(e)) return e;
        bool d2 = false; for (size_t g = 0; g < P.groups.size(); g++) d2 |= (P.groups[g].prm.strategy == 2);
        e = zb_ensureHeavy(c, nbBlocks, P.sd, d2); if (zb_isErr(e)) return e; }
    CK(cudaMemcpyAsync(c->d_blocks, P.blocks.data(), nbBlocks * sizeof(ZbBlock), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(c->d_frames, P.frames.data(), nbFrames * sizeof(ZbFrame), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(c->d_chunks, P.chunks.data(), P.chunks.size() * sizeof(ZbChunk), cudaMemcpyHostToDevice, stream));
    CK(cudaEventRecord(c->evK0, stream));
    unsigned launches = 0;
    if (nbFrames >= 8 || cdict) { size_t const e = zb_buildDictImages(c, P, cdict, d_dictEnd, dictTail, stream); if (zb_isErr(e)) return e; }
    {   size_t const e = zb_runBlocks(c, P, d_src, d_dictEnd, 0, nbBlocks, 0, (u32)P.chunks.size(), 0, stream, true, &launches); if (zb_isErr(e)) return e; }
    CK(zb_launch_stitch(d_src, c->d_blocks, nbBlocks, c->d_frames, c->d_body, P.sd.body, c->d_meta, c->d_outOffsets, NULL, c->d_totals, d_dst, dstCapacity, stream));
    launches += 2;
    if (c->callChecksum) { CK(zb_launch_checksums(d_src, c->d_frames, (u32)nbFrames, c->d_outOffsets, d_dst, dstCapacity, stream)); launches++; }   /* device buffers: hashed on the device, a warp per frame */
    if (cSizes) { CK(zb_launch_frame_sizes(c->d_frames, (u32)nbFrames, c->d_outOffsets, c->d_frameSizes, stream)); launches++; }
    CK(cudaEventRecord(c->evKEnd, stream));
    CK(cudaMemcpyAsync(c->h_totals, c->d_totals, sizeof(u64), cudaMemcpyDeviceToHost, stream));
    if (cSizes) {
        std::vector<u64> tmp(nbFrames);
        CK(cudaMemcpyAsync(tmp.data(), c->d_frameSizes, nbFrames * sizeof(u64), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (size_t f = 0; f < nbFrames; f++) cSizes[f] = (size_t)tmp[f];
    } else CK(cudaStreamSynchronize(stream));
    c->stats.launches = launches;
    c->stats.nbBlocks = nbBlocks;
    {   float ms = 0; cudaEventElapsedTime(&ms, c->evK0, c->evKEnd); c->stats.kernel_ms = ms;
        cudaEventElapsedTime(&ms, c->evK0, c->evK1); c->stats.match_ms = ms;
        if (P.groups.size() == 1) { cudaEventElapsedTime(&ms, c->evK0, c->evMid); c->stats.cand_ms = ms; cudaEventElapsedTime(&ms, c->evMid, c->evK1); c->stats.parse_ms = ms; }
        cudaEventElapsedTime(&ms, c->evK1, c->evK2); c->stats.literals_ms = ms;
        cudaEventElapsedTime(&ms, c->evK2, c->evK3); c->stats.sequences_ms = ms;
        cudaEventElapsedTime(&ms, c->evK3, c->evKEnd); c->stats.stitch_ms = ms; }
    u64 const total = c->h_totals[0];
    if (total > dstCapacity) return ZB_ERR(ZB_error_dstSize_tooSmall);
    return (size_t)total;
}

/* events of one call: destroyed on every exit path (the CK() macro returns early on a CUDA error) */
struct ZbEventSet {
    std::vector<cudaEvent_t> v;
    explicit ZbEventSet(size_t n) : v(n, (cudaEvent_t)0) {}
    ~ZbEventSet() { for (size_t i = 0; i < v.size(); i++) if (v[i]) cudaEventDestroy(v[i]); }
    cudaEvent_t& operator[](size_t i) { return v[i]; }
};

/* XXH64 (lib/common/xxhash.h: XXH64_update / XXH64_digest, seed 0) of the frame's content: the frame checksum is
 * its low 32 bits (zstd_compress.c:5297-5303).  A serial recurrence over 32-byte stripes: it runs on the calling
 * host thread while the GPU works (about 10 GB/s — a checksummed frame is bound by this pass, not by the GPU). */
static u64 zb_xxh64(const u8* p, size_t len)
{
    u64 const P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull, P4 = 0x85EBCA77C2B2AE63ull, P5 = 0x27D4EB2F165667C5ull;
    auto rotl = [](u64 x, int r) { return (x << r) | (x >> (64 - r)); };
    auto rd64 = [](const u8* q) { u64 v; memcpy(&v, q, 8); return v; };
    auto rd32 = [](const u8* q) { u32 v; memcpy(&v, q, 4); return v; };
    auto round = [&](u64 acc, u64 in) { return rotl(acc + in * P2, 31) * P1; };
    auto merge = [&](u64 acc, u64 v) { return (acc ^ round(0, v)) * P1 + P4; };
    const u8* const end = p + len;
    u64 h;
    if (len >= 32) {
        u64 v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;
        const u8* const limit = end - 32;
        do { v1 = round(v1, rd64(p)); v2 = round(v2, rd64(p + 8)); v3 = round(v3, rd64(p + 16)); v4 = round(v4, rd64(p + 24)); p += 32; } while (p <= limit);
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
        h = merge(h, v1); h = merge(h, v2); h = merge(h, v3); h = merge(h, v4);
    } else h = P5;
    h += (u64)len;
    while (p + 8 <= end) { h ^= round(0, rd64(p)); h = rotl(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= (u64)rd32(p) * P1; h = rotl(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (u64)(*p) * P5; h = rotl(h, 11) * P1; p++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* ------------------------------------------------------------------ host pointers: pipelined waves
 * H2D copy of wave w+1 | kernels of waves w, w-1, ... (one stream + workspace slot each) | D2H of finished waves.
 * A block needs ~ms of latency end to end (one warp walks it), so several waves are kept in flight. */

static size_t zb_compressFramesWaves(ZSTD_CCtx* c, u8* dst, size_t dstCapacity, const u8* src,
                                     const size_t* frameOffsets, const size_t* frameSizes, size_t nbFrames,
                                     const void* dict, size_t dictSize, const ZSTD_CDict* cdict, size_t* cSizes, int level, bool deviceMemory)
{
    u32 const waveBlocks128 = deviceMemory ? c->devWaveBlocks : c->hostWaveBlocks;       /* wave size in 128 KiB blocks */
    u32 const ZB_WAVE_SLOTS = deviceMemory ? c->waveSlots : c->hostWaveSlots;
    /* streams are created on first use: every stream beyond the hardware queue count (8 by default) shares a
     * queue with another one, and a download queued behind another wave's kernels stalls the whole pipeline
     * (measured: 16 streams -> every download waited for the last upload; profiles/r1_e2e_timeline.md) */
    auto getStream = [&](u32 i, cudaStream_t* out) -> size_t {
        if (!c->waveStream[i]) CK(cudaStreamCreateWithFlags(&c->waveStream[i], cudaStreamNonBlocking));
        *out = c->waveStream[i];
        return 0;
    };
    cudaStream_t sCopy, sD2H = (cudaStream_t)0;
    if (deviceMemory) {                                           /* creation order as measured: wave streams first */
        cudaStream_t t;
        for (u32 i = 0; i < ZB_WAVE_SLOTS; i++) { size_t const e = getStream(i, &t); if (zb_isErr(e)) return e; }
    }
    {   size_t const e = getStream(ZB_WAVE_SLOTS_MAX, &sCopy); if (zb_isErr(e)) return e; }
    size_t effDict = 0, dictTail = 0; u32 dictID = 0; const u8* d_dictEnd = NULL;
    {   size_t const e = zb_prepareDict(c, dict, dictSize, cdict, sCopy, &effDict, &dictTail, &dictID, &d_dictEnd); if (zb_isErr(e)) return e; }
    if (!c->plan) { c->plan = new (std::nothrow) ZbPlan(); if (!c->plan) return ZB_ERR(ZB_error_memory_allocation); }
    ZbPlan& P = *c->plan; P.reset();
    zb_plan(P, c->callChecksum, frameOffsets, frameSizes, nbFrames, level, effDict, dictTail, c->callNoDictID ? 0u : dictID, c->dictEntropy.present ? c->dictEntropy.rep : NULL,
            c->callPartBegin, c->callPartEnd ? c->callPartEnd : ~0ull);
    if (P.unsupported) return ZB_ERR(ZB_error_parameter_unsupported);
    u32 const nbBlocks = (u32)P.blocks.size();
    /* a wave is sized in bytes of input (and of workspace): calls made of small blocks get proportionally more blocks per wave */
    u32 const ZB_WAVE_BLOCKS = (u32)((u64)waveBlocks128 * (ZB_BLOCK_MAX / P.sd.dist) > (1u << 22) ? (1u << 22) : waveBlocks128 * (ZB_BLOCK_MAX / P.sd.dist));
    /* wave boundaries.  Host path: the call ends when the LAST wave has gone through every kernel, so the
     * final waves shrink (1/2, 1/4, 1/8 of a wave): less work behind the last upload. */
    std::vector<u32> wb, wc;                                      /* wave w = blocks [wb[w], wb[w+1]) = chunks [wc[w], wc[w+1]) */
    {   std::vector<u32> tail, target;
        u32 left = nbBlocks;
        if (!deviceMemory) for (u32 sz = ZB_WAVE_BLOCKS / 8u; sz >= 32u && sz < ZB_WAVE_BLOCKS && left > 2u * sz; sz *= 2u) { tail.push_back(sz); left -= sz; }
        for (u32 b = 0; b < left; ) { u32 const e = (left - b > ZB_WAVE_BLOCKS) ? b + ZB_WAVE_BLOCKS : left; target.push_back(e - b); b = e; }
        for (size_t i = tail.size(); i-- > 0; ) target.push_back(tail[i]);
        /* a chunk is never split between waves: waves are filled chunk by chunk up to their target size */
        wb.push_back(0); wc.push_back(0);
        u32 acc = 0; size_t ti = 0;
        for (u32 ci = 0; ci < (u32)P.chunks.size(); ci++) {
            u32 const nextFirst = ci + 1u < (u32)P.chunks.size() ? P.chunks[ci + 1u].firstBlock : nbBlocks;
            acc += nextFirst - P.chunks[ci].firstBlock;
            if (ti < target.size() && acc >= target[ti] && ci + 1u < (u32)P.chunks.size()) { wb.push_back(nextFirst); wc.push_back(ci + 1u); acc = 0; ti++; }
        }
        wb.push_back(nbBlocks); wc.push_back((u32)P.chunks.size());
    }
    u32 maxWaveBlocks = 0;
    for (size_t w = 0; w + 1 < wb.size(); w++) if (wb[w + 1] - wb[w] > maxWaveBlocks) maxWaveBlocks = wb[w + 1] - wb[w];
    u32 const nbWaves = (u32)wb.size() - 1u;
    u32 const slots = nbWaves < ZB_WAVE_SLOTS ? nbWaves : ZB_WAVE_SLOTS;
    size_t inEnd = 0, bound = 0;
    for (size_t f = 0; f < nbFrames; f++) {
        if (frameOffsets[f] + frameSizes[f] > inEnd) inEnd = frameOffsets[f] + frameSizes[f];
        bound += ZSTD_compressBound(frameSizes[f]) + 32;
    }
    size_t const outCap = deviceMemory ? dstCapacity : (dstCapacity < bound ? dstCapacity : bound);
    {   size_t e = zb_ensureDesc(c, nbBlocks, nbFrames, nbWaves, P.chunks.size()); if (zb_isErr(e)) return e;
        bool d2 = false; for (size_t g = 0; g < P.groups.size(); g++) d2 |= (P.groups[g].prm.strategy == 2);
        e = zb_ensureHeavy(c, (size_t)slots * maxWaveBlocks, P.sd, d2); if (zb_isErr(e)) return e; }
    u8* d_in; u8* d_out;
    if (deviceMemory) { d_in = (u8*)src; d_out = dst; }
    else {
        if (inEnd + 16 > c->d_inCap) { cudaFree(c->d_in); c->d_in = NULL; c->d_inCap = 0; CK(cudaMalloc(&c->d_in, inEnd + 16)); c->d_inCap = inEnd + 16; }
        if (outCap + 16 > c->d_outCap) { cudaFree(c->d_out); c->d_out = NULL; c->d_outCap = 0; CK(cudaMalloc(&c->d_out, outCap + 16)); c->d_outCap = outCap + 16; }
        d_in = c->d_in; d_out = c->d_out;
        size_t const e = getStream(ZB_WAVE_SLOTS_MAX + 1u, &sD2H); if (zb_isErr(e)) return e;
    }
    bool const download = !deviceMemory;
    bool const timeline = getenv("ZSTDB200_TIMELINE") != NULL;      /* development: print each wave's milestones */
    unsigned const evFlags = timeline ? cudaEventDefault : cudaEventDisableTiming;
    ZbEventSet evH2D(nbWaves), evStitch(nbWaves), evSize(nbWaves), evD2H(timeline ? nbWaves : 0);
    std::vector<double> hostDone(nbWaves, 0.0);
    double const hostT0 = zb_now();
    for (u32 w = 0; w < nbWaves; w++) {
        CK(cudaEventCreateWithFlags(&evH2D[w], evFlags));
        CK(cudaEventCreateWithFlags(&evStitch[w], evFlags));
        CK(cudaEventCreateWithFlags(&evSize[w], cudaEventDisableTiming));
        if (timeline) CK(cudaEventCreate(&evD2H[w]));
    }
    unsigned launches = 0;
    size_t err = 0;
    CK(cudaEventRecord(c->evStart, sCopy));
    CK(cudaMemcpyAsync(c->d_blocks, P.blocks.data(), nbBlocks * sizeof(ZbBlock), cudaMemcpyHostToDevice, sCopy));
    CK(cudaMemcpyAsync(c->d_frames, P.frames.data(), nbFrames * sizeof(ZbFrame), cudaMemcpyHostToDevice, sCopy));
    CK(cudaMemcpyAsync(c->d_chunks, P.chunks.data(), P.chunks.size() * sizeof(ZbChunk), cudaMemcpyHostToDevice, sCopy));
    if (nbFrames >= 8 || cdict) { size_t const e = zb_buildDictImages(c, P, cdict, d_dictEnd, dictTail, sCopy); if (zb_isErr(e)) return e; }
    cudaStream_t lastStream = sCopy;
    for (u32 w = 0; w < nbWaves && !err; w++) {
        u32 const b0 = wb[w], b1 = wb[w + 1];
        /* input bytes of the wave (frames are laid out in offset order; history was uploaded by earlier waves) */
        u64 lo = ~0ull, hi = 0;
        for (u32 b = b0; b < b1; b++) { u64 const a = P.blocks[b].srcOff, e = a + P.blocks[b].size; if (a < lo) lo = a; if (e > hi) hi = e; }
        if (hi > lo && !deviceMemory) CK(cudaMemcpyAsync(d_in + lo, src + lo, hi - lo, cudaMemcpyHostToDevice, sCopy));
        CK(cudaEventRecord(evH2D[w], sCopy));
        cudaStream_t st;
        {   size_t const e = getStream(w % slots, &st); if (zb_isErr(e)) return e; }
        lastStream = st;
        CK(cudaStreamWaitEvent(st, evH2D[w], 0));
        err = zb_runBlocks(c, P, d_in, d_dictEnd, b0, b1, wc[w], wc[w + 1], (size_t)(w % slots) * maxWaveBlocks, st, false, &launches);
        if (err) break;
        if (w > 0) CK(cudaStreamWaitEvent(st, evStitch[w - 1], 0));
        size_t const s0 = (size_t)(w % slots) * maxWaveBlocks;
        CK(zb_launch_stitch(d_in, c->d_blocks + b0, b1 - b0, c->d_frames, c->d_body + s0 * P.sd.body, P.sd.body, c->d_meta + s0,
                            c->d_outOffsets + b0, w > 0 ? c->d_totals + (w - 1) : NULL, c->d_totals + w, d_out, outCap, st));
        launches += 2;
        CK(cudaEventRecord(evStitch[w], st));
        /* the wave's size goes to the host behind the event the next wave's stitch waits for: a store into mapped
         * host memory from inside the scan kernel would add a PCIe round trip to every link of that chain */
        if (download || timeline || w + 1 == nbWaves) {
            CK(cudaMemcpyAsync(c->h_totals + w, c->d_totals + w, sizeof(u64), cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(evSize[w], st));
        }
    }
    double const hostEnq = zb_now() - hostT0;
    /* content checksums.  Host buffers: XXH64 on host threads while the GPU works (a serial recurrence per frame: ~10 GB/s
     * per thread).  Device buffers: a warp per frame once the last wave is stitched. */
    std::vector<u64> xxh;
    if (c->callChecksum && !err && !deviceMemory) {
        xxh.resize(nbFrames);
        size_t const nt = nbFrames < 8 ? nbFrames : 8;
        if (nt <= 1) xxh[0] = zb_xxh64(src + frameOffsets[0], frameSizes[0]);
        else {
            std::vector<std::thread> th;
            for (size_t t = 0; t < nt; t++) th.emplace_back([&, t] { for (size_t f = t; f < nbFrames; f += nt) xxh[f] = zb_xxh64(src + frameOffsets[f], frameSizes[f]); });
            for (size_t t = 0; t < nt; t++) th[t].join();
        }
    }
    if (c->callChecksum && !err && deviceMemory) { CK(zb_launch_checksums(d_in, c->d_frames, (u32)nbFrames, c->d_outOffsets, d_out, outCap, lastStream)); launches++; }
    bool const wantSizes = cSizes != NULL || (c->callChecksum && !deviceMemory);
    std::vector<u64> fsz;
    u64 prev = 0, total = 0;
    if (download || timeline) {
        /* drain: as each wave's size becomes known, ship its bytes */
        for (u32 w = 0; w < nbWaves && !err; w++) {
            CK(cudaEventSynchronize(evSize[w]));
            hostDone[w] = zb_now() - hostT0;
            total = c->h_totals[w];
            if (download && total <= outCap && total > prev) CK(cudaMemcpyAsync(dst + prev, d_out + prev, total - prev, cudaMemcpyDeviceToHost, sD2H));
            if (total <= outCap) prev = total;
            if (timeline) CK(cudaEventRecord(evD2H[w], download ? sD2H : lastStream));
        }
    }
    if (!err && wantSizes) {
        CK(zb_launch_frame_sizes(c->d_frames, (u32)nbFrames, c->d_outOffsets, c->d_frameSizes, lastStream));
        fsz.resize(nbFrames);
        CK(cudaMemcpyAsync(fsz.data(), c->d_frameSizes, nbFrames * sizeof(u64), cudaMemcpyDeviceToHost, lastStream));
        CK(cudaStreamSynchronize(lastStream));
        if (cSizes) for (size_t f = 0; f < nbFrames; f++) cSizes[f] = (size_t)fsz[f];
    }
    /* the last wave's stitch is ordered behind every earlier one (evStitch chain) */
    if (download) { CK(cudaEventRecord(c->evEnd, sD2H)); CK(cudaStreamSynchronize(sD2H)); }
    else CK(cudaEventRecord(c->evEnd, lastStream));
    for (u32 s = 0; s < slots; s++) if (c->waveStream[s]) CK(cudaStreamSynchronize(c->waveStream[s]));
    CK(cudaStreamSynchronize(sCopy));
    if (!err && nbWaves) total = c->h_totals[nbWaves - 1];
    if (!err && c->callChecksum && !deviceMemory && total <= dstCapacity) {                               /* the 4 bytes the size scan left free behind every frame */
        u64 end = 0;
        for (size_t f = 0; f < nbFrames; f++) {
            end += fsz[f];
            if (end < 4 || end > total) break;
            u32 const ck = (u32)xxh[f];
            dst[end - 4] = (u8)ck; dst[end - 3] = (u8)(ck >> 8); dst[end - 2] = (u8)(ck >> 16); dst[end - 1] = (u8)(ck >> 24);
        }
    }
    if (timeline) {
        fprintf(stderr, "zstd_b200 timeline (ms after the call's first enqueue; host enqueue loop took %.3f ms; %s)\n", 1e3 * hostEnq,
                deviceMemory ? "device buffers" : "per-wave downloads");
        for (u32 w = 0; w < nbWaves; w++) {
            float a = 0, b = 0, e = 0;
            cudaEventElapsedTime(&a, c->evStart, evH2D[w]); cudaEventElapsedTime(&b, c->evStart, evStitch[w]);
            cudaEventElapsedTime(&e, c->evStart, evD2H[w]);
            fprintf(stderr, "  wave %2u blocks %5u..%5u : uploaded %7.3f  stitched %7.3f (host saw it %7.3f)  downloaded %7.3f\n",
                    w, wb[w], wb[w + 1], a, b, 1e3 * hostDone[w], e);
        }
    }
    if (err) return err;
    {   float ms = 0; cudaEventElapsedTime(&ms, c->evStart, c->evEnd); c->stats.total_ms = ms; c->stats.kernel_ms = ms; }
    c->stats.launches = launches; c->stats.nbBlocks = nbBlocks;
    if (!deviceMemory) { c->stats.h2d_bytes = inEnd; c->stats.d2h_bytes = (size_t)total; }
    if (total > dstCapacity) return ZB_ERR(ZB_error_dstSize_tooSmall);
    return (size_t)total;
}

static size_t zb_compressFramesAny(ZSTD_CCtx* c, void* dst, size_t dstCapacity,
                                   const void* src, const size_t* frameOffsets, const size_t* frameSizes,
                                   size_t nbFrames, const void* dict, size_t dictSize, const ZSTD_CDict* cdict,
                                   size_t* cSizes, int level, int deviceMemory, void* streamv)
{
    if (!c) return ZB_ERR(ZB_error_GENERIC);
    if (nbFrames == 0) return 0;
    ZbDeviceGuard guard;
    {   size_t const e = zb_ctxInit(c); if (zb_isErr(e)) return e; }
    memset(&c->stats, 0, sizeof(c->stats));
    if (deviceMemory) {
        /* device-resident input: large calls are cut into waves on several streams, so that the shared-memory
         * bound candidate walk of one wave overlaps the register-only parse / entropy kernels of another;
         * ZSTDB200_SERIAL=1 (or a caller-supplied stream) keeps one wave on one stream — the mode whose
         * per-kernel event times are meaningful */
        u64 bytes = 0, nb = 0; u32 maxBlock = 0;
        for (size_t f = 0; f < nbFrames; f++) {
            bytes += frameSizes[f]; nb += (frameSizes[f] + ZB_BLOCK_MAX - 1) / ZB_BLOCK_MAX + (frameSizes[f] == 0);
            u32 const m = frameSizes[f] < ZB_BLOCK_MAX ? (u32)frameSizes[f] : ZB_BLOCK_MAX; if (m > maxBlock) maxBlock = m;
        }
        ZbStrides const sd = zb_strides(maxBlock);
        u64 const wsBytes = nb * ((u64)sd.dist * 6u + (u64)sd.seq * 8u + sd.lit + sd.body);        /* one-wave workspace */
        if (!streamv && c->devWaveBlocks && (bytes >= 2ull * c->devWaveBlocks * ZB_BLOCK_MAX || wsBytes > (12ull << 30)))
            return zb_compressFramesWaves(c, (u8*)dst, dstCapacity, (const u8*)src, frameOffsets, frameSizes, nbFrames, dict, dictSize, cdict, cSizes, level, true);
        cudaStream_t stream = streamv ? (cudaStream_t)streamv : c->stream;
        size_t const r = zb_compressFramesDevice(c, (u8*)dst, dstCapacity, (const u8*)src, frameOffsets, frameSizes, nbFrames, dict, dictSize, cdict, cSizes, level, stream);
        c->stats.total_ms = c->stats.kernel_ms;
        return r;
    }
    return zb_compressFramesWaves(c, (u8*)dst, dstCapacity, (const u8*)src, frameOffsets, frameSizes, nbFrames, dict, dictSize, cdict, cSizes, level, false);
}

extern "C" size_t ZSTDB200_compressFrames(ZSTD_CCtx* c, void* dst, size_t dstCapacity,
                                          const void* src, const size_t* frameOffsets, const size_t* frameSizes,
                                          size_t nbFrames, const void* dict, size_t dictSize,
                                          size_t* cSizes, int level, int deviceMemory, void* streamv)
{
    if (!c) return ZB_ERR(ZB_error_GENERIC);
    c->callChecksum = (u32)c->advChecksum; c->callNoDictID = (u32)c->advNoDictID;      /* the sticky frame parameters of ZSTD_CCtx_setParameter apply */
    size_t const r = zb_compressFramesAny(c, dst, dstCapacity, src, frameOffsets, frameSizes, nbFrames, dict, dictSize, NULL, cSizes, level, deviceMemory, streamv);
    c->callChecksum = 0; c->callNoDictID = 0;
    return r;
}

extern "C" size_t ZSTDB200_compressFrames_usingCDict(ZSTD_CCtx* c, void* dst, size_t dstCapacity,
                                                     const void* src, const size_t* frameOffsets, const size_t* frameSizes,
                                                     size_t nbFrames, const ZSTD_CDict* cdict,
                                                     size_t* cSizes, int deviceMemory, void* streamv)
{
    if (!cdict) return ZB_ERR(ZB_error_dictionary_wrong);                        /* zstd_compress.c:5753 */
    if (!c) return ZB_ERR(ZB_error_GENERIC);
    c->callChecksum = (u32)c->advChecksum; c->callNoDictID = (u32)c->advNoDictID;
    size_t const r = zb_compressFramesAny(c, dst, dstCapacity, src, frameOffsets, frameSizes, nbFrames, NULL, 0, cdict, cSizes, cdict->level, deviceMemory, streamv);
    c->callChecksum = 0; c->callNoDictID = 0;
    return r;
}

/* ------------------------------------------------------------------ digested dictionaries (lib/zstd.h:967-995) */
extern "C" ZSTD_CDict* ZSTD_createCDict(const void* dict, size_t dictSize, int level)     /* zstd_compress.c:5633 */
{
    ZSTD_CDict* cd = (ZSTD_CDict*)calloc(1, sizeof(ZSTD_CDict));
    if (!cd) return NULL;
    cd->level = level == 0 ? 3 : level;                                          /* ZSTD_CLEVEL_DEFAULT, :5640 */
    cd->device = -1;
    cd->size = dict ? dictSize : 0;
    cd->content = (u8*)malloc(cd->size ? cd->size : 1);
    cd->lock = new (std::nothrow) std::mutex();
    if (!cd->content || !cd->lock) { free(cd->content); delete cd->lock; free(cd); return NULL; }
    if (cd->size) memcpy(cd->content, dict, cd->size);                           /* ZSTD_dlm_byCopy */
    if (cd->size >= 8) {
        size_t const off = zb_loadDictionary(&cd->entropy, cd->content, cd->size);
        if (zb_isErr(off)) { free(cd->content); delete cd->lock; free(cd); return NULL; }   /* corrupted entropy tables: creation fails (:5600-5612) */
        cd->contentOff = off;
        size_t const contentSize = cd->size - off;
        cd->tail = contentSize < ZB_PRIME_BYTES ? contentSize : ZB_PRIME_BYTES;
    }
    return cd;
}

extern "C" size_t ZSTD_freeCDict(ZSTD_CDict* cd)                                            /* accepts NULL, zstd_compress.c:5655 */
{
    if (!cd) return 0;
    if (cd->device >= 0) {
        int prev = -1; cudaGetDevice(&prev);
        cudaSetDevice(cd->device);
        cudaFree(cd->d_dict); cudaFree(cd->d_de); cudaFree(cd->d_image); cudaFree(cd->d_dictChunk);
        if (prev >= 0) cudaSetDevice(prev);
    }
    free(cd->content); delete cd->lock; free(cd);
    return 0;
}

extern "C" unsigned ZSTD_getDictID_fromCDict(const ZSTD_CDict* cd)                           /* zstd_compress.c:5738 */
{
    return (cd && cd->size >= 8 && cd->entropy.present) ? cd->entropy.dictID : 0u;
}

extern "C" unsigned ZSTD_getDictID_fromDict(const void* dict, size_t dictSize)              /* lib/decompress/zstd_ddict.c:227, zstd.h:1105 */
{
    const u8* d = (const u8*)dict;
    if (!d || dictSize < 8) return 0;
    if ((d[0] | (d[1] << 8) | (d[2] << 16) | ((u32)d[3] << 24)) != 0xEC30A437u) return 0;  /* ZSTD_MAGIC_DICTIONARY */
    return d[4] | (d[5] << 8) | (d[6] << 16) | ((u32)d[7] << 24);
}

extern "C" size_t ZSTD_compress_usingCDict(ZSTD_CCtx* c, void* dst, size_t dstCapacity, const void* src, size_t srcSize,
                                           const ZSTD_CDict* cdict)                          /* zstd_compress.c:5836 */
{
    size_t const off = 0;
    if (!c) return ZB_ERR(ZB_error_GENERIC);
    if (!cdict) return ZB_ERR(ZB_error_dictionary_wrong);
    if (dstCapacity && !dst) return ZB_ERR(ZB_error_dstBuffer_null);
    if (dstCapacity < 18) return ZB_ERR(ZB_error_dstSize_tooSmall);
    return zb_compressFramesAny(c, dst, dstCapacity, src, &off, &srcSize, 1, NULL, 0, cdict, NULL, cdict->level, 0, NULL);
}

/* One rank's share of a frame that several GPUs compress together (the reference's counterpart are the jobs
 * of ZSTDMT, zstdmt_compress.c:1168-1227: a job reads an overlap of preceding input and only the first writes the frame
 * header, only the last the end mark).  Chunks are the unit: partBegin must be a multiple of ZSTDB200_framePartAlignment()
 * (512 KiB), and the bytes [partBegin - ZSTDB200_framePartHalo(), partBegin + partSize) of the frame must be resident:
 * d_part points at frame offset partBegin - min(partBegin, halo).  The ranks' outputs, concatenated in order, are byte
 * for byte the frame one GPU would have produced. */
extern "C" size_t ZSTDB200_framePartAlignment(void) { return (size_t)ZB_CHUNK_BLOCKS * ZB_BLOCK_MAX; }
extern "C" size_t ZSTDB200_framePartHalo(void) { return ZB_PRIME_BYTES; }
extern "C" size_t ZSTDB200_compressFramePart(ZSTD_CCtx* c, void* d_dst, size_t dstCapacity, const void* d_part,
                                             size_t frameSize, size_t partBegin, size_t partSize, int level, void* stream)
{
    if (!c) return ZB_ERR(ZB_error_GENERIC);
    if (partBegin % ZSTDB200_framePartAlignment() || partBegin + partSize > frameSize || (partSize == 0 && frameSize != 0)) return ZB_ERR(ZB_error_srcSize_wrong);
    if (c->advChecksum) return ZB_ERR(ZB_error_parameter_unsupported);          /* a content checksum needs the whole content in one place */
    size_t const halo = partBegin < ZB_PRIME_BYTES ? partBegin : ZB_PRIME_BYTES;
    const u8* const frameBase = (const u8*)d_part + halo - partBegin;          /* address frame offset 0 would have; only offsets >= partBegin - halo are touched */
    size_t const off = 0;
    c->callPartBegin = partBegin; c->callPartEnd = partBegin + partSize; c->callNoDictID = (u32)c->advNoDictID;
    if (frameSize == 0) { c->callPartBegin = 0; c->callPartEnd = 1; }
    size_t const r = zb_compressFramesAny(c, d_dst, dstCapacity, frameBase, &off, &frameSize, 1, NULL, 0, NULL, NULL, level, 1, stream);
    c->callPartBegin = 0; c->callPartEnd = 0; c->callNoDictID = 0;
    return r;
}

extern "C" size_t ZSTDB200_compressDevice(ZSTD_CCtx* c, void* d_dst, size_t dstCapacity,
                                          const void* d_src, size_t srcSize, int level, void* stream)
{
    size_t const off = 0;
    return ZSTDB200_compressFrames(c, d_dst, dstCapacity, d_src, &off, &srcSize, 1, NULL, 0, NULL, level, 1, stream);
}

/* Seek table of a run of frames, in the reference's seekable format (contrib/seekable_format/
 * zstd_seekable_compression_format.md; writer: zstdseek_compress.c:268-360): a skippable frame holding one
 * (compressed size, decompressed size) pair per frame and the 9-byte footer.  Appended behind the frames a batch or a
 * multi-GPU call produced, it makes the output randomly accessible for the reference's ZSTD_seekable_* readers.  Host
 * code, no GPU.  Returns the number of bytes written (17 + 8 * nbFrames) or an error code. */
extern "C" size_t ZSTDB200_writeSeekTable(void* dstv, size_t dstCapacity, const size_t* cSizes, const size_t* dSizes, size_t nbFrames)
{
    u8* const dst = (u8*)dstv;
    if (nbFrames > 0x8000000u) return ZB_ERR(ZB_error_srcSize_wrong);                         /* ZSTD_SEEKABLE_MAXFRAMES, zstd_seekable.h:20 */
    size_t const need = 8 + 8 * nbFrames + 9;
    if (!dst || (nbFrames && (!cSizes || !dSizes))) return ZB_ERR(ZB_error_GENERIC);
    if (dstCapacity < need) return ZB_ERR(ZB_error_dstSize_tooSmall);
    auto w32 = [](u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); };
    for (size_t f = 0; f < nbFrames; f++)                           /* 32-bit fields; ZSTD_SEEKABLE_MAX_FRAME_DECOMPRESSED_SIZE = 1 GiB (zstd_seekable.h:19) */
        if (cSizes[f] > 0xFFFFFFFFull || dSizes[f] > 0x40000000ull) return ZB_ERR(ZB_error_srcSize_wrong);
    w32(dst, 0x184D2A5Eu);                                                                     /* Skippable_Magic_Number */
    w32(dst + 4, (u32)(need - 8));                                                             /* Frame_Size */
    u8* p = dst + 8;
    for (size_t f = 0; f < nbFrames; f++) { w32(p, (u32)cSizes[f]); w32(p + 4, (u32)dSizes[f]); p += 8; }
    w32(p, (u32)nbFrames); p[4] = 0;                                                           /* Number_Of_Frames, descriptor: no checksums */
    w32(p + 5, 0x8F92EAB1u);                                                                   /* Seekable_Magic_Number */
    return need;
}

/* the frame checksum's hash, exported for the CPU tests (compared with the reference's ZSTD_XXH64) */
extern "C" unsigned long long ZSTDB200_xxh64(const void* p, size_t len) { return zb_xxh64((const u8*)p, len); }

/* Host-side planning of one call, without touching a GPU (what the CPU tests compare with the oracle's plan).
 * Per frame, `out` receives 16 values: strategy, mls, tableN, tableNLong, stepSize, litDisabled, windowLog, insStep,
 * number of blocks, size of the first block, flags of the first block, history of the last block, dictionary part of
 * the first block's history, number of chunks, history walked by the last chunk, size of the last chunk.
 * Returns the total number of blocks. */
extern "C" size_t ZSTDB200_describePlan(const size_t* frameSizes, size_t nbFrames, int level, size_t dictSize, size_t dictTail, unsigned* out)
{
    std::vector<size_t> offs(nbFrames);
    size_t o = 0; for (size_t f = 0; f < nbFrames; f++) { offs[f] = o; o += frameSizes[f]; }
    ZbPlan P;
    zb_plan(P, 0, offs.data(), frameSizes, nbFrames, level, dictSize, dictTail, 0, NULL);
    for (size_t f = 0; f < nbFrames && out; f++) {
        ZbFrame const& fr = P.frames[f];
        const ZbParams* prm = NULL;
        for (size_t g = 0; g < P.groups.size(); g++) if (P.groups[g].b0 <= fr.firstBlock && fr.firstBlock < P.groups[g].b1) prm = &P.groups[g].prm;
        ZbBlock const& b0 = P.blocks[fr.firstBlock]; ZbBlock const& bl = P.blocks[fr.firstBlock + fr.nbBlocks - 1];
        u32 nc = 0; const ZbChunk* lastChunk = NULL;
        for (size_t ci = 0; ci < P.chunks.size(); ci++) if (P.chunks[ci].firstBlock >= fr.firstBlock && P.chunks[ci].firstBlock < fr.firstBlock + fr.nbBlocks) { nc++; lastChunk = &P.chunks[ci]; }
        unsigned* r = out + f * 16;
        r[0] = prm->strategy; r[1] = prm->mls; r[2] = prm->tableN; r[3] = prm->tableNLong; r[4] = prm->stepSize; r[5] = prm->litDisabled;
        r[6] = fr.windowLog; r[7] = prm->insStep; r[8] = fr.nbBlocks; r[9] = b0.size; r[10] = b0.flags;
        r[11] = bl.histLen; r[12] = b0.dictLen; r[13] = nc; r[14] = lastChunk ? lastChunk->histLen : 0; r[15] = lastChunk ? lastChunk->size : 0;
    }
    return P.blocks.size();
}

extern "C" void ZSTDB200_getLastStats(const ZSTD_CCtx* c, ZSTDB200_stats* out) { if (c && out) *out = c->stats; }


/* ------------------------------------------------------------------ advanced one-shot API (lib/zstd.h:337-603, :1088-1102)
 * ZSTD_CCtx_setParameter + ZSTD_compress2 is how python-zstandard, zstd-jni and the zstd CLI drive the library today.
 * Supported: compressionLevel, checksumFlag (XXH64 of the content on the host), contentSizeFlag (always written),
 * dictIDFlag, nbWorkers / jobSize / overlapLog (accepted and ignored: parallelism is the GPU's), the cParams only at
 * their default 0; everything else answers parameter_unsupported.  ZSTD_compressStream2 serves the one-shot form
 * (first call, ZSTD_e_end, output room >= ZSTD_compressBound: lib/zstd.h:787) — real streaming is out of scope. */
extern "C" size_t ZSTD_CCtx_setParameter(ZSTD_CCtx* c, ZSTD_cParameter paramE, int value)                 /* zstd_compress.c:720 */
{
    if (!c) return ZB_ERR(ZB_error_GENERIC);
    int const param = (int)paramE;
    switch (param) {
    case 100: c->advLevel = value == 0 ? 3 : (value > 22 ? 22 : (value < -(int)ZB_BLOCK_MAX ? -(int)ZB_BLOCK_MAX : value)); return 0;   /* ZSTD_c_compressionLevel */
    case 200: return 0;                                                                      /* ZSTD_c_contentSizeFlag: the size is always written */
    case 201: c->advChecksum = value != 0; return 0;                                         /* ZSTD_c_checksumFlag */
    case 202: c->advNoDictID = value == 0; return 0;                                         /* ZSTD_c_dictIDFlag */
    case 400: case 401: case 402: return 0;                                                  /* nbWorkers, jobSize, overlapLog */
    case 101: case 102: case 103: case 104: case 105: case 106: case 107:                    /* windowLog .. strategy: default only */
    case 160: case 161: case 162: case 163: case 164:                                        /* long distance matching: off only */
        return value == 0 ? 0 : ZB_ERR(ZB_error_parameter_unsupported);
    default: return ZB_ERR(ZB_error_parameter_unsupported);
    }
}

extern "C" size_t ZSTD_CCtx_setPledgedSrcSize(ZSTD_CCtx* c, unsigned long long) { return c ? 0 : ZB_ERR(ZB_error_GENERIC); }   /* one-shot calls know their size */

extern "C" size_t ZSTD_CCtx_reset(ZSTD_CCtx* c, ZSTD_ResetDirective reset)                                   /* zstd_compress.c:1390; 1 session, 2 parameters, 3 both */
{
    if (!c) return ZB_ERR(ZB_error_GENERIC);
    if (reset == 1 || reset == 3) { c->stInSize = 0; c->stOutSize = 0; c->stOutPos = 0; c->stFrames = 0; }   /* an unfinished stream is dropped */
    if (reset == 2 || reset == 3) {
        c->advLevel = 3; c->advChecksum = 0; c->advNoDictID = 0;
        ZSTD_freeCDict(c->advLocalDict); c->advLocalDict = NULL; c->advRefCDict = NULL;
    }
    return 0;
}

extern "C" size_t ZSTD_CCtx_loadDictionary(ZSTD_CCtx* c, const void* dict, size_t dictSize)  /* zstd_compress.c:1260: copied, sticky */
{
    if (!c) return ZB_ERR(ZB_error_GENERIC);
    ZSTD_freeCDict(c->advLocalDict); c->advLocalDict = NULL; c->advRefCDict = NULL;
    if (!dict || dictSize == 0) return 0;                                                    /* NULL / 0: back to no dictionary */
    c->advLocalDict = ZSTD_createCDict(dict, dictSize, c->advLevel);
    return c->advLocalDict ? 0 : ZB_ERR(ZB_error_dictionary_corrupted);
}

extern "C" size_t ZSTD_CCtx_refCDict(ZSTD_CCtx* c, const ZSTD_CDict* cdict)                  /* zstd_compress.c:1330: borrowed, sticky */
{
    if (!c) return ZB_ERR(ZB_error_GENERIC);
    ZSTD_freeCDict(c->advLocalDict); c->advLocalDict = NULL;
    c->advRefCDict = cdict;
    return 0;
}

extern "C" size_t ZSTD_compress2(ZSTD_CCtx* c, void* dst, size_t dstCapacity, const void* src, size_t srcSize)     /* zstd_compress.c:6365 */
{
    size_t const off = 0;
    if (!c) return ZB_ERR(ZB_error_GENERIC);
    if (dstCapacity && !dst) return ZB_ERR(ZB_error_dstBuffer_null);
    if (dstCapacity < 18) return ZB_ERR(ZB_error_dstSize_tooSmall);
    const ZSTD_CDict* const cd = c->advRefCDict ? c->advRefCDict : c->advLocalDict;
    int const level = c->advRefCDict ? c->advRefCDict->level : c->advLevel;                  /* a referenced CDict brings its own level (:5836) */
    c->callChecksum = (u32)c->advChecksum; c->callNoDictID = (u32)c->advNoDictID;
    size_t const r = zb_compressFramesAny(c, dst, dstCapacity, src, &off, &srcSize, 1, NULL, 0, cd, NULL, level, 0, NULL);
    c->callChecksum = 0; c->callNoDictID = 0;
    return r;
}

/* Streaming (lib/zstd.h:681-862).  The unit of GPU work is a whole frame, so the stream front end collects input on the
 * host and turns it into frames:
 *   ZSTD_e_continue  input is taken into the context's buffer; whenever ZB_STREAM_FRAME bytes are there they become a frame;
 *   ZSTD_e_flush     what is buffered becomes a frame now (the reference ends a block, here a frame ends: every byte given
 *                    so far is decodable from the output, which is what a flush promises);
 *   ZSTD_e_end       same, and the session is over once everything was handed out (an empty session yields an empty frame).
 * The result is a sequence of frames — a valid zstd stream that every decoder reads as the concatenation of their contents
 * (lib/zstd.h:160-162; contrib/pzstd writes the same shape) — not one frame as the reference's stream would be.
 * The first call of a session carrying everything with ZSTD_e_end and room for ZSTD_compressBound() bytes is served without
 * any buffering (the one-shot form, lib/zstd.h:787).  Return value: bytes still waiting to be handed out (0 = flushed). */
#define ZB_STREAM_FRAME ((size_t)256 << 20)
static size_t zb_streamHandOut(ZSTD_CCtx* c, ZSTD_outBuffer* out)
{
    size_t const have = c->stOutSize - c->stOutPos, room = out->size - out->pos;
    size_t const n = have < room ? have : room;
    if (n) { memcpy((u8*)out->dst + out->pos, c->stOut + c->stOutPos, n); out->pos += n; c->stOutPos += n; }
    if (c->stOutPos == c->stOutSize) { c->stOutPos = 0; c->stOutSize = 0; }
    return c->stOutSize - c->stOutPos;
}
static size_t zb_streamMakeFrame(ZSTD_CCtx* c)                        /* stIn -> one frame appended to stOut */
{
    size_t const bound = ZSTD_compressBound(c->stInSize) + 32;
    if (c->stOutSize + bound > c->stOutCap) {
        size_t const cap = c->stOutSize + bound;
        u8* const p = (u8*)realloc(c->stOut, cap);
        if (!p) return ZB_ERR(ZB_error_memory_allocation);
        c->stOut = p; c->stOutCap = cap;
    }
    size_t const r = ZSTD_compress2(c, c->stOut + c->stOutSize, bound, c->stIn ? c->stIn : (const u8*)"", c->stInSize);
    if (ZSTD_isError(r)) return r;
    c->stOutSize += r; c->stInSize = 0; c->stFrames++;
    return 0;
}
extern "C" size_t ZSTD_compressStream2(ZSTD_CCtx* c, ZSTD_outBuffer* out, ZSTD_inBuffer* in, ZSTD_EndDirective endOp)       /* zstd_compress.c:6176 */
{
    if (!c || !out || !in) return ZB_ERR(ZB_error_GENERIC);
    if (out->pos > out->size) return ZB_ERR(ZB_error_dstSize_tooSmall);
    if (in->pos > in->size) return ZB_ERR(ZB_error_srcSize_wrong);
    if ((int)endOp < 0 || (int)endOp > 2) return ZB_ERR(ZB_error_parameter_unsupported);
    size_t const n = in->size - in->pos, room = out->size - out->pos;
    bool const idle = c->stInSize == 0 && c->stOutSize == 0 && c->stFrames == 0;
    if (idle && endOp == ZSTD_e_end && room >= ZSTD_compressBound(n)) {      /* one-shot: straight from the caller's buffers */
        size_t const r = ZSTD_compress2(c, (u8*)out->dst + out->pos, room, (const u8*)in->src + in->pos, n);
        if (ZSTD_isError(r)) return r;
        in->pos = in->size; out->pos += r;
        return 0;
    }
    /* output produced earlier goes first; input is only taken while nothing is waiting */
    if (zb_streamHandOut(c, out) == 0) {
        size_t take = n;
        while (take) {
            size_t const space = ZB_STREAM_FRAME - c->stInSize;
            size_t const m = take < space ? take : space;
            if (c->stInSize + m > c->stInCap) {
                size_t cap = c->stInCap ? c->stInCap : ((size_t)1 << 20);
                while (cap < c->stInSize + m) cap *= 2;
                if (cap > ZB_STREAM_FRAME) cap = ZB_STREAM_FRAME;
                u8* const p = (u8*)realloc(c->stIn, cap);
                if (!p) return ZB_ERR(ZB_error_memory_allocation);
                c->stIn = p; c->stInCap = cap;
            }
            memcpy(c->stIn + c->stInSize, (const u8*)in->src + in->pos, m);
            c->stInSize += m; in->pos += m; take -= m;
            if (c->stInSize == ZB_STREAM_FRAME) {                           /* a full frame's worth: compress it, hand out what fits */
                size_t const e = zb_streamMakeFrame(c); if (ZSTD_isError(e)) return e;
                if (zb_streamHandOut(c, out) != 0) break;                    /* the caller has to make room before more input is taken */
            }
        }
        if (in->pos == in->size && endOp != ZSTD_e_continue && c->stOutSize == 0) {
            if (c->stInSize || (endOp == ZSTD_e_end && c->stFrames == 0)) { size_t const e = zb_streamMakeFrame(c); if (ZSTD_isError(e)) return e; }
            zb_streamHandOut(c, out);
        }
    }
    size_t const waiting = c->stOutSize - c->stOutPos;
    if (endOp == ZSTD_e_end && waiting == 0 && in->pos == in->size && c->stInSize == 0) c->stFrames = 0;   /* session over: the next call starts a new one */
    if (endOp == ZSTD_e_continue) return waiting ? waiting : (ZB_STREAM_FRAME - c->stInSize);              /* a hint for the next input size, as the reference gives one */
    return waiting + ((in->pos < in->size || c->stInSize) ? 1 : 0);                                        /* > 0 while the flush / end is incomplete */
}
/* the older streaming entry points are thin forms of the above (lib/zstd.h:832-862) */
extern "C" ZSTD_CStream* ZSTD_createCStream(void) { return ZSTD_createCCtx(); }
extern "C" size_t ZSTD_freeCStream(ZSTD_CStream* zcs) { return ZSTD_freeCCtx(zcs); }
extern "C" size_t ZSTD_initCStream(ZSTD_CStream* zcs, int level)
{
    if (!zcs) return ZB_ERR(ZB_error_GENERIC);
    ZSTD_CCtx_reset(zcs, ZSTD_reset_session_only);
    ZSTD_CCtx_refCDict(zcs, NULL);
    return ZSTD_CCtx_setParameter(zcs, ZSTD_c_compressionLevel, level);
}
extern "C" size_t ZSTD_compressStream(ZSTD_CStream* zcs, ZSTD_outBuffer* output, ZSTD_inBuffer* input) { return ZSTD_compressStream2(zcs, output, input, ZSTD_e_continue); }
extern "C" size_t ZSTD_flushStream(ZSTD_CStream* zcs, ZSTD_outBuffer* output) { ZSTD_inBuffer in = { NULL, 0, 0 }; return ZSTD_compressStream2(zcs, output, &in, ZSTD_e_flush); }
extern "C" size_t ZSTD_endStream(ZSTD_CStream* zcs, ZSTD_outBuffer* output) { ZSTD_inBuffer in = { NULL, 0, 0 }; return ZSTD_compressStream2(zcs, output, &in, ZSTD_e_end); }
extern "C" size_t ZSTD_CStreamInSize(void) { return ZB_BLOCK_MAX; }                                         /* lib/zstd.h:858 */
extern "C" size_t ZSTD_CStreamOutSize(void) { return ZSTD_compressBound(ZB_BLOCK_MAX) + 3 + 4; }            /* lib/zstd.h:859 */

/* ------------------------------------------------------------------ reference-identical entry points */
extern "C" size_t ZSTD_compress_usingDict(ZSTD_CCtx* c, void* dst, size_t dstCapacity, const void* src, size_t srcSize,
                                          const void* dict, size_t dictSize, int level)
{
    size_t const off = 0;
    if (!c) return ZB_ERR(ZB_error_GENERIC);
    if (dstCapacity && !dst) return ZB_ERR(ZB_error_dstBuffer_null);
    if (dstCapacity < 18) return ZB_ERR(ZB_error_dstSize_tooSmall);             /* ZSTD_FRAMEHEADERSIZE_MAX, zstd_compress.c:4643 */
    return zb_compressFramesAny(c, dst, dstCapacity, src, &off, &srcSize, 1, dict, dict ? dictSize : 0, NULL, NULL, level, 0, NULL);   /* the simple API ignores sticky parameters (lib/zstd.h:270-273) */
}
extern "C" size_t ZSTD_compressCCtx(ZSTD_CCtx* c, void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level)
{
    return ZSTD_compress_usingDict(c, dst, dstCapacity, src, srcSize, NULL, 0, level);
}
extern "C" size_t ZSTD_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level)
{
    ZSTD_CCtx* c = ZSTD_createCCtx();                                            /* zstd_compress.c:5423-5440: temporary context */
    if (!c) return ZB_ERR(ZB_error_memory_allocation);
    size_t const r = ZSTD_compressCCtx(c, dst, dstCapacity, src, srcSize, level);
    ZSTD_freeCCtx(c);
    return r;
}

extern "C" size_t ZSTD_compressBound(size_t srcSize)                             /* lib/zstd.h:235 */
{
    if (srcSize >= (sizeof(size_t) == 8 ? 0xFF00FF00FF00FF00ULL : 0xFF00FF00U)) return ZB_ERR(ZB_error_srcSize_wrong);
    return srcSize + (srcSize >> 8) + ((srcSize < (128u << 10)) ? (((128u << 10) - srcSize) >> 11) : 0);
}
extern "C" unsigned ZSTD_isError(size_t code) { return code > ZB_ERR(ZB_error_maxCode); }
extern "C" int ZSTD_getErrorCode(size_t code) { return ZSTD_isError(code) ? (int)(0 - code) : 0; }
extern "C" const char* ZSTD_getErrorName(size_t code)                            /* common/error_private.c:14-62 */
{
    switch (ZSTD_getErrorCode(code)) {
    case 0: return "No error detected";
    case 1: return "Error (generic)";
    case 10: return "Unknown frame descriptor";
    case 12: return "Version not supported";
    case 14: return "Unsupported frame parameter";
    case 16: return "Frame requires too much memory for decoding";
    case 20: return "Data corruption detected";
    case 22: return "Restored data doesn't match checksum";
    case 24: return "Header of Literals' block doesn't respect format specification";
    case 30: return "Dictionary is corrupted";
    case 32: return "Dictionary mismatch";
    case 34: return "Cannot create Dictionary from provided samples";
    case 40: return "Unsupported parameter";
    case 41: return "Unsupported combination of parameters";
    case 42: return "Parameter is out of bound";
    case 44: return "tableLog requires too much memory : unsupported";
    case 46: return "Unsupported max Symbol Value : too large";
    case 48: return "Specified maxSymbolValue is too small";
    case 50: return "pledged buffer stability condition is not respected";
    case 60: return "Operation not authorized at current processing stage";
    case 62: return "Context should be init first";
    case 64: return "Allocation error : not enough memory";
    case 66: return "workSpace buffer is not large enough";
    case 70: return "Destination buffer is too small";
    case 72: return "Src size is incorrect";
    case 74: return "Operation on NULL destination buffer";
    case 80: return "Operation made no progress over multiple calls, due to output buffer being full";
    case 82: return "Operation made no progress over multiple calls, due to input being empty";
    default: return "Unspecified error code";
    }
}
extern "C" int ZSTD_minCLevel(void) { return -(int)ZB_BLOCK_MAX; }                /* zstd_compress.c:7038 */
extern "C" int ZSTD_maxCLevel(void) { return 22; }
extern "C" int ZSTD_defaultCLevel(void) { return 3; }
extern "C" unsigned ZSTD_versionNumber(void) { return 10506; }                    /* lib/zstd.h:107-110 */
extern "C" const char* ZSTD_versionString(void) { return "1.5.6"; }
