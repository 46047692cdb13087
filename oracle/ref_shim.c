/* ref_shim.c — TEST INFRASTRUCTURE ONLY.
 * Compiles the UNMODIFIED reference translation unit lib/compress/zstd_compress.c (included from
 * where it lies under $(REF); nothing is copied) together with a few extern wrappers that expose
 * its file-static entropy stage, so tests can pin the oracle byte-for-byte against the reference.
 * Built into oracle/_ref/libzstd_ref.so by oracle/Makefile (only when the reference tree exists).
 */
#include "compress/zstd_compress.c"     /* resolved through -I$(REF)/lib */

/* Runs the reference's ZSTD_entropyCompressSeqStore (zstd_compress.c:3001) on a hand-built
 * seqStore with a FRESH entropy state (ZSTD_reset_compressedBlockState, :1911).
 * Returns the compressed block-body size, 0 for "emit raw", or a zstd error code. */
size_t ref_entropyCompressBlock(void* dst, size_t dstCapacity,
                                const unsigned* offBase, const unsigned* litLen, const unsigned* matchLen, size_t nbSeq,
                                const void* literals, size_t litSize, size_t blockSrcSize,
                                int strategy, unsigned targetLength)
{
    seqStore_t ss;
    ZSTD_compressedBlockState_t* prev = (ZSTD_compressedBlockState_t*)calloc(1, sizeof(*prev));
    ZSTD_compressedBlockState_t* next = (ZSTD_compressedBlockState_t*)calloc(1, sizeof(*next));
    ZSTD_CCtx_params params;
    void* wksp = malloc(TMP_WORKSPACE_SIZE);
    size_t r, i;
    memset(&ss, 0, sizeof(ss));
    memset(&params, 0, sizeof(params));
    ss.maxNbSeq = nbSeq + 1; ss.maxNbLit = litSize + 8;
    ss.sequencesStart = (seqDef*)malloc((nbSeq + 1) * sizeof(seqDef));
    ss.llCode = (BYTE*)malloc(nbSeq + 1); ss.mlCode = (BYTE*)malloc(nbSeq + 1); ss.ofCode = (BYTE*)malloc(nbSeq + 1);
    ss.litStart = (BYTE*)malloc(litSize + 8);
    memcpy(ss.litStart, literals, litSize);
    ss.lit = ss.litStart + litSize;
    ss.longLengthType = ZSTD_llt_none;
    for (i = 0; i < nbSeq; i++) {                       /* same bookkeeping as ZSTD_storeSeq, zstd_compress_internal.h:671-729 */
        unsigned const mlBase = matchLen[i] - MINMATCH;
        ss.sequencesStart[i].offBase = offBase[i];
        ss.sequencesStart[i].litLength = (U16)litLen[i];
        ss.sequencesStart[i].mlBase = (U16)mlBase;
        if (litLen[i] > 0xFFFF) { ss.longLengthType = ZSTD_llt_literalLength; ss.longLengthPos = (U32)i; }
        if (mlBase > 0xFFFF)    { ss.longLengthType = ZSTD_llt_matchLength;   ss.longLengthPos = (U32)i; }
    }
    ss.sequences = ss.sequencesStart + nbSeq;
    ZSTD_reset_compressedBlockState(prev);
    params.cParams.strategy = (ZSTD_strategy)strategy;
    params.cParams.targetLength = targetLength;
    params.literalCompressionMode = ZSTD_ps_auto;
    r = ZSTD_entropyCompressSeqStore(&ss, &prev->entropy, &next->entropy, &params,
                                     dst, dstCapacity, blockSrcSize, wksp, TMP_WORKSPACE_SIZE, 0 /* bmi2 */);
    free(ss.sequencesStart); free(ss.llCode); free(ss.mlCode); free(ss.ofCode); free(ss.litStart);
    free(prev); free(next); free(wksp);
    return r;
}

/* Same as ref_entropyCompressBlock, but the previous-block entropy state is what ZSTD_loadCEntropy
 * (zstd_compress.c:4987) installs from a zstd-format dictionary. */
size_t ref_entropyCompressBlock_dict(void* dst, size_t dstCapacity,
                                const unsigned* offBase, const unsigned* litLen, const unsigned* matchLen, size_t nbSeq,
                                const void* literals, size_t litSize, size_t blockSrcSize,
                                int strategy, unsigned targetLength, const void* dict, size_t dictSize)
{
    seqStore_t ss;
    ZSTD_compressedBlockState_t* prev = (ZSTD_compressedBlockState_t*)calloc(1, sizeof(*prev));
    ZSTD_compressedBlockState_t* next = (ZSTD_compressedBlockState_t*)calloc(1, sizeof(*next));
    ZSTD_CCtx_params params;
    void* wksp = malloc(TMP_WORKSPACE_SIZE);
    size_t r, i;
    memset(&ss, 0, sizeof(ss));
    memset(&params, 0, sizeof(params));
    ss.maxNbSeq = nbSeq + 1; ss.maxNbLit = litSize + 8;
    ss.sequencesStart = (seqDef*)malloc((nbSeq + 1) * sizeof(seqDef));
    ss.llCode = (BYTE*)malloc(nbSeq + 1); ss.mlCode = (BYTE*)malloc(nbSeq + 1); ss.ofCode = (BYTE*)malloc(nbSeq + 1);
    ss.litStart = (BYTE*)malloc(litSize + 8);
    memcpy(ss.litStart, literals, litSize);
    ss.lit = ss.litStart + litSize;
    ss.longLengthType = ZSTD_llt_none;
    for (i = 0; i < nbSeq; i++) {
        unsigned const mlBase = matchLen[i] - MINMATCH;
        ss.sequencesStart[i].offBase = offBase[i];
        ss.sequencesStart[i].litLength = (U16)litLen[i];
        ss.sequencesStart[i].mlBase = (U16)mlBase;
        if (litLen[i] > 0xFFFF) { ss.longLengthType = ZSTD_llt_literalLength; ss.longLengthPos = (U32)i; }
        if (mlBase > 0xFFFF)    { ss.longLengthType = ZSTD_llt_matchLength;   ss.longLengthPos = (U32)i; }
    }
    ss.sequences = ss.sequencesStart + nbSeq;
    ZSTD_reset_compressedBlockState(prev);
    r = ZSTD_loadCEntropy(prev, wksp, dict, dictSize);
    if (!ZSTD_isError(r)) {
        params.cParams.strategy = (ZSTD_strategy)strategy;
        params.cParams.targetLength = targetLength;
        params.literalCompressionMode = ZSTD_ps_auto;
        r = ZSTD_entropyCompressSeqStore(&ss, &prev->entropy, &next->entropy, &params,
                                         dst, dstCapacity, blockSrcSize, wksp, TMP_WORKSPACE_SIZE, 0);
    }
    free(ss.sequencesStart); free(ss.llCode); free(ss.mlCode); free(ss.ofCode); free(ss.litStart);
    free(prev); free(next); free(wksp);
    return r;
}

/* ZSTD_getCParams_internal (zstd_compress.c:7123) as the simple API calls it (:5405, cpm_noAttachDict) */
void ref_getCParams_simpleApi(int level, unsigned long long srcSize, size_t dictSize, unsigned out[7])
{
    ZSTD_compressionParameters const cp = ZSTD_getCParams_internal(level, srcSize, dictSize, ZSTD_cpm_noAttachDict);
    out[0] = cp.windowLog; out[1] = cp.chainLog; out[2] = cp.hashLog; out[3] = cp.searchLog;
    out[4] = cp.minMatch; out[5] = cp.targetLength; out[6] = (unsigned)cp.strategy;
}
