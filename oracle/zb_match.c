/* zb_match.c — oracle model of the block-parallel "warp-batch" greedy match-finder
 * (TEST INFRASTRUCTURE ONLY; the CUDA kernel in zstd_b200/csrc must reproduce it bit-for-bit).
 *
 * What it restates: the greedy single-probe LZ77 parse of ZSTD_compressBlock_fast
 * (/root/reference/lib/compress/zstd_fast.c:192-423): multiplicative hash of `mls` bytes
 * (zstd_compress_internal.h:821-861), one candidate per bucket, 4-byte verification, repcode-1
 * probe, backward catch-up (:387-391), forward count (:396), sparse post-match inserts
 * (:403-408), immediate repcode-2 loop (:410-420), step acceleration every 128 bytes without a
 * match (:234,342-347), position pairs (p, p+1) spaced by `step` (:225-229).
 *
 * What is different by design (block-parallel, data-parallel within a block):
 *   - every block is parsed independently: private table, primed from the `primeBytes` of input
 *     preceding the block (the ZSTDMT overlap idea, zstdmt_compress.c:726-731), encoder repcodes
 *     start invalid (zstdmt_compress.c:737-742);
 *   - 32 probe positions ("lanes") are evaluated per step against the table state at the start
 *     of the step; the lowest matching position wins; only lanes up to the winner insert;
 *   - table entries are 16-bit positions modulo 64 KiB (reach 65535 bytes), which is what lets
 *     7 blocks per SM keep their tables in shared memory.
 */
#include <string.h>
#include <stdlib.h>
#include "zb_oracle.h"

static inline u64 rd64(const u8* p) { u64 v; memcpy(&v, p, 8); return v; }
static inline u32 rd32(const u8* p) { u32 v; memcpy(&v, p, 4); return v; }

/* zstd_compress_internal.h:815-861 : hash of the low `mls` bytes of an 8-byte LE load */
static const u64 prime4 = 2654435761u;
static const u64 prime5 = 889523592379ull;
static const u64 prime6 = 227718039650203ull;
static const u64 prime7 = 58295818150454627ull;
static const u64 prime8 = 0xCF1BBCDCB7A56463ull;
static inline u32 zb_hash(u64 v, u32 mls, u32 hBits)
{
    switch (mls) {
    default:
    case 4: return (u32)(((u32)v * (u32)prime4) >> (32 - hBits));
    case 5: return (u32)(((v << 24) * prime5) >> (64 - hBits));
    case 6: return (u32)(((v << 16) * prime6) >> (64 - hBits));
    case 7: return (u32)(((v << 8) * prime7) >> (64 - hBits));
    case 8: return (u32)((v * prime8) >> (64 - hBits));
    }
}

/* zstd_compress_internal.h:771-795 */
static size_t zb_count(const u8* ip, const u8* match, const u8* iend)
{
    const u8* const start = ip;
    while (ip < iend && *ip == *match) { ip++; match++; }
    return (size_t)(ip - start);
}

void zbo_makePlan(zbo_plan* plan, const zbo_cparams* cp)
{
    memset(plan, 0, sizeof(*plan));
    plan->strategy = cp->strategy;
    plan->windowLog = cp->windowLog;
    plan->mls = cp->minMatch < 4 ? 4 : (cp->minMatch > 8 ? 8 : cp->minMatch);
    if (cp->strategy == 1) {
        plan->hashLog = cp->hashLog > 14 ? 14 : cp->hashLog;          /* 32 KiB of u16 per block */
        plan->longHashLog = 0;
        plan->stepSize = cp->targetLength + !cp->targetLength + 1;     /* zstd_fast.c:200 */
    } else {
        plan->hashLog = cp->chainLog;                                  /* short table, zstd_double_fast.c:113 */
        plan->longHashLog = cp->hashLog;
        plan->stepSize = 1;
    }
    plan->primeBytes = ZB_PRIME_DEFAULT;
    if (plan->primeBytes > (1u << cp->windowLog)) plan->primeBytes = 1u << cp->windowLog;
    /* zstd_compress_internal.h:621-633 */
    plan->litCompressionDisabled = (cp->strategy == 1) && (cp->targetLength > 0);
}

/* ---- 16-bit modular table ---- */
typedef struct { u16* t; u32 hashLog; size_t base; } ztable;   /* base = lowLimit (absolute) */

static inline void zt_put(ztable* z, u32 h, size_t pos) { z->t[h] = (u16)(pos - z->base); }
/* most recent position q < pos with (q - base) == stored (mod 65536); (size_t)-1 if none */
static inline size_t zt_get(const ztable* z, u32 h, size_t pos)
{
    u32 const rel = (u32)(pos - z->base);
    u32 const dist = (rel - z->t[h]) & 0xFFFFu;
    if (dist == 0 || dist > rel) return (size_t)-1;
    return pos - dist;
}

typedef struct { zbo_seq* seqs; size_t nbSeq; u8* lit; size_t litSize; const u8* frame; } emitter;
static void emit(emitter* e, size_t anchor, size_t litLen, size_t matchLen, u32 offBase)
{
    memcpy(e->lit + e->litSize, e->frame + anchor, litLen);
    e->litSize += litLen;
    e->seqs[e->nbSeq].offBase = offBase;
    e->seqs[e->nbSeq].litLen = (u32)litLen;
    e->seqs[e->nbSeq].matchLen = (u32)matchLen;
    e->nbSeq++;
}

#define ZB_FILL_STEP 1u       /* priming inserts every position (the reference primes every 3rd: zstd_fast.c:63) */

static size_t matchBlock_fast(const zbo_plan* plan, const u8* frame, size_t frameSize,
                              size_t bs, size_t blockSize, zbo_seq* seqs, u8* lit, size_t* litSizePtr)
{
    size_t const be = bs + blockSize;
    size_t const lowLimit = bs > plan->primeBytes ? bs - plan->primeBytes : 0;
    u32 const mls = plan->mls, hlog = plan->hashLog;
    ztable zt;
    emitter em = { seqs, 0, lit, 0, frame };
    size_t ip = bs, anchor = bs, searchStart = bs;
    u32 rep1 = 0, rep2 = 0;
    (void)frameSize;

    zt.t = (u16*)calloc((size_t)1 << hlog, sizeof(u16));
    zt.hashLog = hlog; zt.base = lowLimit;

    /* ---- prime the table from history [lowLimit, bs) ; later positions overwrite earlier ones ---- */
    for (size_t p = lowLimit; p < bs; p += ZB_FILL_STEP) {
        if (p + 8 > be) break;                                   /* hash reads 8 bytes, stay inside the block end */
        zt_put(&zt, zb_hash(rd64(frame + p), mls, hlog), p);
    }

    while (1) {
        u32 const step = plan->stepSize + (u32)((ip - searchStart) >> 7);     /* kSearchStrength = 8 */
        size_t p[ZB_WARP]; u32 h[ZB_WARP]; size_t cand[ZB_WARP]; int act[ZB_WARP], hit[ZB_WARP];
        int winner = -1, l, nact = 0;

        for (l = 0; l < (int)ZB_WARP; l++) {
            p[l] = ip + (size_t)(l >> 1) * step + (size_t)(l & 1);
            act[l] = (p[l] + 8 <= be);
            if (act[l]) nact = l + 1;
        }
        if (!act[0]) break;

        /* all lanes look at the table as it was when the step began */
        for (l = 0; l < nact; l++) {
            hit[l] = 0;
            if (!act[l]) continue;
            h[l] = zb_hash(rd64(frame + p[l]), mls, hlog);
            cand[l] = zt_get(&zt, h[l], p[l]);
            {   u32 const cur = rd32(frame + p[l]);
                if (rep1 && p[l] >= lowLimit + rep1 && rd32(frame + p[l] - rep1) == cur) hit[l] = 2;     /* repcode 1 */
                else if (cand[l] != (size_t)-1 && rd32(frame + cand[l]) == cur) hit[l] = 1;                 /* table hit */
            }
            if (hit[l] && winner < 0) winner = l;
        }

        /* inserts: lanes up to the winner (all active lanes when nobody matched); highest lane wins a bucket */
        {   int const last = winner >= 0 ? winner : nact - 1;
            for (l = 0; l <= last; l++) if (act[l]) zt_put(&zt, h[l], p[l]);
        }

        if (winner < 0) { ip += (size_t)(ZB_WARP / 2) * step; continue; }

        {   size_t const probe = p[winner];
            size_t ms = probe, mm;          /* match start / match source */
            size_t mlen;
            u32 offset, offBase;
            int const isRep = (hit[winner] == 2);
            offset = isRep ? rep1 : (u32)(probe - cand[winner]);
            mm = ms - offset;
            /* backward catch-up (zstd_fast.c:387-391) */
            while (ms > anchor && mm > lowLimit && frame[ms - 1] == frame[mm - 1]) { ms--; mm--; }
            mlen = (probe - ms) + 4 + zb_count(frame + probe + 4, frame + probe - offset + 4, frame + be);
            if (isRep && ms > anchor) offBase = 1;              /* REPCODE1_TO_OFFBASE, needs litLength > 0 */
            else { offBase = offset + 3; rep2 = rep1; rep1 = offset; }   /* decoder pushes every full offset */
            emit(&em, anchor, ms - anchor, mlen, offBase);
            ip = ms + mlen; anchor = ip;

            if (ip + 8 <= be) {
                /* sparse fill (zstd_fast.c:403-408) */
                zt_put(&zt, zb_hash(rd64(frame + probe + 2), mls, hlog), probe + 2);
                zt_put(&zt, zb_hash(rd64(frame + ip - 2), mls, hlog), ip - 2);
                /* immediate repcode-2 (zstd_fast.c:410-420) */
                while (ip + 8 <= be && rep2 && rd32(frame + ip) == rd32(frame + ip - rep2)) {
                    size_t const rlen = 4 + zb_count(frame + ip + 4, frame + ip + 4 - rep2, frame + be);
                    { u32 const t = rep2; rep2 = rep1; rep1 = t; }
                    zt_put(&zt, zb_hash(rd64(frame + ip), mls, hlog), ip);
                    emit(&em, anchor, 0, rlen, 1);
                    ip += rlen; anchor = ip;
                }
            }
            searchStart = ip;
        }
    }
    /* trailing literals (zstd_compress.c:3365-3366) */
    memcpy(em.lit + em.litSize, frame + anchor, be - anchor);
    em.litSize += be - anchor;
    *litSizePtr = em.litSize;
    free(zt.t);
    return em.nbSeq;
}

size_t zbo_matchBlock(const zbo_plan* plan, const u8* frame, size_t frameSize,
                      size_t blockStart, size_t blockSize, zbo_seq* seqs, u8* lit, size_t* litSizePtr)
{
    return matchBlock_fast(plan, frame, frameSize, blockStart, blockSize, seqs, lit, litSizePtr);
}
