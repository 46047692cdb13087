/* zb_match.cu — K1: LZ77 match-finder ("fast" and "doubleFast" strategies) = K1a candidate walk, K1b greedy parse, K1c merge.
 *
 * Replaces the CPU loops ZSTD_compressBlock_fast_noDict_generic / _extDict_generic
 * (/root/reference/lib/compress/zstd_fast.c:192-423, :709-960) and ZSTD_compressBlock_doubleFast_noDict_generic
 * (zstd_double_fast.c:105-323) with a data-parallel formulation:
 *   - K1a (one CTA per CHUNK of up to 4 blocks) keeps the chunk's hash table in shared memory — primed from
 *     the <=128 KiB in front of the chunk (zstd_fast.c:53-85 does this for dictionaries, zstdmt_compress.c:1182-1227 for
 *     job overlaps), then alive through all blocks of the chunk — and visits the positions in BATCHES of 1024: every
 *     position of a batch reads its bucket, positions that found no candidate are inserted (one atomicMax each: the
 *     lowest position of a batch wins a bucket), and a position that found nothing looks once more after the batch's
 *     insertions.  Two barriers per batch, no dependent global load: the walk does not depend on the parse;
 *   - K1b (one warp per 16 KiB segment of a block) does the greedy selection: 32 probe positions per step — pairs
 *     (p, p+1) spaced by `step` as in zstd_fast.c:225-229 — lowest matching lane wins (warp ballot); backward
 *     catch-up (:387-391) and forward extension (ZSTD_count, zstd_compress_internal.h:771) are warp-cooperative:
 *     32 x 8 bytes per round, first differing lane found by ballot.  A match may run past its segment's end;
 *   - K1c (one CTA per block) joins the segments: drops what lies under a match that ran over from an earlier segment,
 *     runs the repcode history over the block's sequences and gathers the literal bytes from the input.
 * Everything is deterministic: no result depends on the order in which threads reach an atomic.
 * The bit-exact CPU model of these kernels is oracle/zb_match.c (tests only).
 */
#include "zb_device.cuh"
#include "zb_kernels.h"

#ifndef WALK_FAST_BC
#define WALK_FAST_BC 1           /* development switch: 0 = entries and the second look through zb_walk_entry / zb_walk_cand */
#endif

/* matched bytes starting at rel positions (a, a - offset), never reading at or past `be` */
template <bool DICT>
__device__ __forceinline__ u32 zb_count_fwd(const ZbSeg& sg, u32 a, u32 offset, u32 be, u32 lane)
{
    u32 fwd = 0;
    while (true) {
        u32 const pa = a + fwd + 8u * lane;
        u32 m;
        if (pa + 8u <= be) {
            u64 const x = zb_seg_ld64x<DICT>(sg, pa) ^ zb_seg_ld64x<DICT>(sg, pa - offset);
            m = x ? (u32)((__ffsll((long long)x) - 1) >> 3) : 8u;
        } else {
            m = 0;
            while (pa + m < be && zb_seg_byte<DICT>(sg, pa + m) == zb_seg_byte<DICT>(sg, pa + m - offset)) m++;
        }
        u32 const inc = __ballot_sync(ZB_FULL, m != 8u);
        if (inc == 0) { fwd += 256u; continue; }
        int const f = __ffs((int)inc) - 1;
        fwd += 8u * (u32)f + __shfl_sync(ZB_FULL, m, f);
        return fwd;
    }
}

template <bool DICT>
__device__ __forceinline__ u32 zb_back_coop(const ZbSeg& sg, u32 probe, u32 offset, u32 anchor, u32 lane)
{
    u32 back = 0;
    while (true) {
        u32 const k = back + lane + 1u;                        /* compare bytes probe-k and probe-offset-k */
        bool const ok = (probe >= anchor + k) && (probe >= offset + k)
                     && (zb_seg_byte<DICT>(sg, probe - k) == zb_seg_byte<DICT>(sg, probe - offset - k));
        u32 const okb = __ballot_sync(ZB_FULL, ok);
        u32 const cnt = (okb == ZB_FULL) ? 32u : (u32)(__ffs((int)~okb) - 1);
        back += cnt;
        if (cnt < 32u) return back;
    }
}

/* raw sequence of the parse kernels: real offset, match length, match start relative to the block */
__device__ __forceinline__ u64 zb_pack_raw(u32 off, u32 mlen, u32 msRel) { return (u64)off | ((u64)mlen << 24) | ((u64)msRel << 42); }
#define ZB_RAW_OFF(r)  ((u32)(r) & 0xFFFFFFu)
#define ZB_RAW_MLEN(r) ((u32)((r) >> 24) & 0x3FFFFu)
#define ZB_RAW_MS(r)   ((u32)((r) >> 42))
/* final sequence, read by the sequences kernel */
__device__ __forceinline__ u64 zb_pack_seq(u32 offBase, u32 litLen, u32 matchLen)
{
    return (u64)offBase | ((u64)litLen << 24) | ((u64)matchLen << 42);
}

/* ------------------------------------------------------------------------------------------------
 * K1a — candidate walk (parse-independent).  One CTA per chunk, table in shared memory.
 * dist[p] = distance from p to its candidate: the latest earlier position that was inserted into p's bucket and has
 * p's tag, 0 if none.  Distances >= 0xFFFF go to the `far` array (dist16 = ZB_FAR).
 * ---------------------------------------------------------------------------------------------- */
/* A table entry is (key + 1) << 11 | tag.  key = walk coordinate x of the position with its offset inside the batch
 * reversed: of all insertions of one batch into a bucket the LOWEST position has the largest key, and every batch beats
 * the batches before it — so one shared-memory atomicMax per insertion arbitrates a batch, whatever the thread order. */
#define ZB_TAG_MASK ((1u << ZB_TAG_BITS) - 1u)
__device__ __forceinline__ u32 zb_walk_key(u32 x) { return (x | (ZB_BATCH - 1u)) - (x & (ZB_BATCH - 1u)); }   /* its own inverse */
__device__ __forceinline__ u32 zb_walk_entry(u32 h, u32 x) { return ((zb_walk_key(x) + 1u) << ZB_TAG_BITS) | (h & ZB_TAG_MASK); }
/* candidate distance of the position at walk coordinate x given a bucket's content c (0 = no candidate).
 * An empty bucket (c = 0) decodes to a coordinate far above x. */
__device__ __forceinline__ u32 zb_walk_cand(u32 c, u32 h, u32 x)
{
    u32 const px = zb_walk_key((c >> ZB_TAG_BITS) - 1u);
    return (((c ^ h) & ZB_TAG_MASK) == 0u && px < x) ? x - px : 0u;
}

/* which of the P consecutive positions starting at a position whose residue modulo `step` is r0 lie on the insertion
 * pattern (rel % step) < 2, as a bit mask.  step >= 3: the pair that began at or before the first position (bits 0,1
 * for r0 = 0; bit 0 for r0 = 1), then a pair every `step` positions from step - r0 on. */
template <int P>
__device__ __forceinline__ u32 zb_walk_pattern_res(u32 r0, u32 step)
{
    if (step <= 2u) return (1u << P) - 1u;
    u32 m = 3u >> (r0 < 2u ? r0 : 2u);
    u32 nxt = step - r0;
#pragma unroll
    for (int k = 0; k < (P + 2) / 3; k++) {                       /* at most ceil(P / 3) further pairs start inside P positions */
        m |= 3u << (nxt < 31u ? nxt : 31u);
        nxt += step;
    }
    return m & ((1u << P) - 1u);
}
/* residue of rel0 modulo step (rel0 < 2^22, step < 2^14: the float quotient is exact up to +-1, fixed below) */
__device__ __forceinline__ u32 zb_walk_residue(u32 rel0, u32 step)
{
    u32 const q = (u32)__fdividef((float)rel0, (float)step);
    int r = (int)rel0 - (int)(q * step);
    if (r < 0) r += (int)step; else if (r >= (int)step) r -= (int)step;
    return (u32)r;
}

/* One batch of the walk for one thread: P consecutive positions from walk coordinate xa.
 * INTERIOR: every position of the batch is walked, lies in the frame's own bytes and has its 8 bytes readable: no
 * activity predicates, bytes come from the words prefetched in wrd[].
 * Returns whether any position of the CTA found a candidate in phase A (the barrier between A and B carries the OR). */
#define ZB_WALK_NWR(P) ((P) == 1 ? 2 : ((P) + 7 + 3) / 4)        /* realigned words that hold a thread's P + 7 bytes */
template <int MLS, int P, bool INTERIOR>
__device__ __forceinline__ bool zb_walk_batch(u32* __restrict__ table, u32 xa, const u32 (&wrd)[ZB_WALK_NWR(P)], u32 pat,
                                              u32 N, u32 shift, u32 D, u32 total, u32 xLow, u32 xEnd,
                                              const u8* fbase, const u8* dbase, bool output, u16* __restrict__ distRow, u32* __restrict__ farRow)
{
    u32 h[P], bkt[P], dOld[P];
    bool act[P];
    /* ---- A: hash, read the bucket ---- */
    {   u32 const rel0 = xa - shift;
        bool const slow = !INTERIOR && ((xa < xLow) || (D != 0u && rel0 < D && rel0 + P + 7u > D) || (rel0 + P + 7u > total));
        u32 old[P];
#pragma unroll
        for (int i = 0; i < P; i++) {
            u32 const x = xa + (u32)i, rel = x - shift;
            act[i] = INTERIOR ? true : ((x >= xLow) && (rel + 8u <= total) && !(rel < D && rel + 8u > D));
            u64 v;
            if (slow) v = act[i] ? zb_ld64u((rel < D ? dbase : fbase) + rel) : 0ull;
            else {
                constexpr int NWR = ZB_WALK_NWR(P);
                int const wi = i >> 2; u32 const sh = 8u * (u32)(i & 3);
                int const w2 = wi + 2 < NWR ? wi + 2 : NWR - 1;           /* only read when i & 3: then wi + 2 < NWR */
                u32 const lo = (i & 3) ? __funnelshift_r(wrd[wi], wrd[wi + 1], sh) : wrd[wi];
                u32 const hi = (i & 3) ? __funnelshift_r(wrd[wi + 1], wrd[w2], sh) : wrd[wi + 1];
                v = ((u64)hi << 32) | lo;
            }
            h[i] = zb_hash(v, MLS, 32u);
            bkt[i] = __umulhi(h[i], N);
            old[i] = act[i] ? table[bkt[i]] : 0u;
        }
#pragma unroll
        for (int i = 0; i < P; i++) dOld[i] = zb_walk_cand(old[i], h[i], xa + (u32)i);
    }
    u32 anyOld = 0;
#pragma unroll
    for (int i = 0; i < P; i++) anyOld |= dOld[i];
    bool const anyHit = __syncthreads_or(anyOld != 0u) != 0;
    /* ---- B: insertions ---- */
#if WALK_FAST_BC
    /* a thread's first coordinate is a multiple of P, so the reversed in-batch offsets of its positions count down from
     * position 0's: key(xa + i) = key(xa) - i; an entry is (Y1 - i) << TAG_BITS | tag with Y1 = key(xa) + 1 */
    u32 const Y1 = zb_walk_key(xa) + 1u;
#pragma unroll
    for (int i = 0; i < P; i++)
        if (act[i] && dOld[i] == 0u && ((pat >> i) & 1u)) atomicMax(&table[bkt[i]], ((Y1 - (u32)i) << ZB_TAG_BITS) | (h[i] & ZB_TAG_MASK));
    __syncthreads();
    /* ---- C: second look, output.  A position without a candidate can only have gained one from this batch's insertions
     * (its bucket held no entry with its tag before, and what was there came from earlier batches): inside one batch the
     * distance is the difference of the keys, so the look is shift, add, tag test, sign test ---- */
    u32 d[P];
#pragma unroll
    for (int i = 0; i < P; i++) {
        d[i] = dOld[i];
        if (act[i] && d[i] == 0u) {
            u32 const c = table[bkt[i]];
            int const dd = (int)((c >> ZB_TAG_BITS) - Y1 + (u32)i);        /* key(candidate) - key(me) */
            d[i] = ((((c ^ h[i]) & ZB_TAG_MASK) == 0u) && dd > 0) ? (u32)dd : 0u;
        }
    }
#else
#pragma unroll
    for (int i = 0; i < P; i++)
        if (act[i] && dOld[i] == 0u && ((pat >> i) & 1u)) atomicMax(&table[bkt[i]], zb_walk_entry(h[i], xa + (u32)i));
    __syncthreads();
    /* ---- C: second look, output ---- */
    u32 d[P];
#pragma unroll
    for (int i = 0; i < P; i++) {
        d[i] = dOld[i];
        if (act[i] && d[i] == 0u) d[i] = zb_walk_cand(table[bkt[i]], h[i], xa + (u32)i);
    }
#endif
    if (output && (INTERIOR || xa < xEnd)) {
        u32 anyFar = 0;
#pragma unroll
        for (int i = 0; i < P; i++) anyFar |= d[i] >= ZB_FAR ? 1u : 0u;
        if (anyFar) {
#pragma unroll
            for (int i = 0; i < P; i++) if (d[i] >= ZB_FAR) { if (INTERIOR || xa + (u32)i < xEnd) farRow[i] = d[i]; d[i] = ZB_FAR; }
        }
        bool vec = false;
        if constexpr (P == 16) { if (INTERIOR || xa + 16u <= xEnd) { uint4* const o4 = reinterpret_cast<uint4*>(distRow);
                                     o4[0] = make_uint4(d[0] | (d[1] << 16), d[2] | (d[3] << 16), d[4] | (d[5] << 16), d[6] | (d[7] << 16));
                                     o4[1] = make_uint4(d[8] | (d[9] << 16), d[10] | (d[11] << 16), d[12] | (d[13] << 16), d[14] | (d[15] << 16)); vec = true; } }
        if constexpr (P == 8) { if (INTERIOR || xa + 8u <= xEnd) { *reinterpret_cast<uint4*>(distRow) = make_uint4(d[0] | (d[1] << 16), d[2] | (d[3] << 16), d[4] | (d[5] << 16), d[6] | (d[7] << 16)); vec = true; } }
        if constexpr (P == 4) { if (INTERIOR || xa + 4u <= xEnd) { *reinterpret_cast<uint2*>(distRow) = make_uint2(d[0] | (d[1] << 16), d[2] | (d[3] << 16)); vec = true; } }
        if constexpr (P == 2) { if (INTERIOR || xa + 2u <= xEnd) { *reinterpret_cast<u32*>(distRow) = d[0] | (d[1] << 16); vec = true; } }
        if (!vec) {
#pragma unroll
            for (int i = 0; i < P; i++) if (xa + (u32)i < xEnd) distRow[i] = (u16)d[i];
        }
    }
    return anyHit;
}

/* P consecutive positions per thread, ZB_BATCH / P threads per CTA.  Per batch:
 *   A  every position hashes its 8 bytes and reads its bucket (the table as the previous batch left it);
 *   B  positions that found no candidate and lie on the insertion pattern atomicMax their entry into the bucket;
 *   C  positions that found nothing in A look again: the batch's lowest insertion into their bucket may serve them.
 * Two barriers per batch (A|B, which also tells every thread whether the batch saw a hit, and B|C); C of one batch
 * and A of the next share a region.  The input bytes of a batch are loaded two batches ahead. */
#ifndef WALK_STEADY
#define WALK_STEADY 1            /* development switch: 0 = every batch takes the general path */
#endif
#ifndef WALK_P_SMALL
#define WALK_P_SMALL 8           /* positions per thread for tables <= 56 KiB: 128 threads per CTA (2.62 ms per GiB against 2.72 with 4; development knob: tools/build_variant.sh) */
#endif
#ifndef WALK_P_MID
#define WALK_P_MID 4             /* tables of 56 .. 113 KiB: two CTAs per SM (measured on config 4: 10.2 ms with 4 positions per thread, 11.2 with 2) */
#endif
#ifndef WALK_MINB_SMALL
#define WALK_MINB_SMALL 4        /* CTAs per SM the register allocation of that variant leaves room for */
#endif
template <int MLS, int P>
__global__ void __launch_bounds__(ZB_BATCH / P, P == WALK_P_SMALL ? WALK_MINB_SMALL : 1)
zb_walk_kernel(const u8* __restrict__ src, const u8* __restrict__ dictEnd, const ZbChunk* __restrict__ chunks, u32 insStep, u32 N, ZbStrides sd,
               u32 slotFirstBlock, u16* __restrict__ dist, u32* __restrict__ far,
               const u32* __restrict__ imageIn, u32* __restrict__ imageOut)
{
    constexpr u32 THREADS = ZB_BATCH / P;
    extern __shared__ __align__(16) u32 table[];
    u32 const t = threadIdx.x;
    ZbChunk const cd = chunks[blockIdx.x];
    bool const buildImage = imageOut != nullptr;
    u32 const D = cd.dictLen;                                     /* rel position of the frame's first byte when a dictionary is in front */
    u32 const H = cd.histLen;                                     /* rel position of the chunk's first byte */
    u32 const total = buildImage ? D : H + cd.size;               /* rel positions [0, total) are walked; no read at or past total */
    bool const fromImage = imageIn != nullptr && D != 0u && !buildImage;
    /* walk coordinate x = rel + shift: batch borders (frame positions that are multiples of the batch; dictionary
     * positions count backwards from the frame start) are the multiples of ZB_BATCH in x */
    u32 const shift = (ZB_BATCH - (D % ZB_BATCH)) % ZB_BATCH;
    const u8* const fbase = src + cd.srcOff - H;                  /* fbase + rel = the byte's address for rel >= D */
    const u8* const dbase = dictEnd - D;                          /* same for rel < D */

    if (fromImage) for (u32 i = t; i < N; i += THREADS) table[i] = __ldg(imageIn + i);
    else           for (u32 i = t; i < N; i += THREADS) table[i] = 0u;
    __syncthreads();

    u32 const xEnd = total + shift;
    u32 const xLow = (fromImage ? D : 0u) + shift;                /* first walked coordinate */
    u32 x0 = xLow & ~(ZB_BATCH - 1u);                             /* border of the first batch */
    /* interior batches: completely walked, completely in the frame's own bytes, all 8-byte reads inside [.., total) */
    u32 const xIntLo = ((D + shift) > xLow ? (D + shift) : xLow);
    u32 const xIntLoB = (xIntLo + ZB_BATCH - 1u) & ~(ZB_BATCH - 1u);
    /* a thread's P + 7 bytes lie in NW aligned words whatever its address modulo 4; the interior loads all NW of them */
    constexpr u32 NW = (P + 7u + 3u + 3u) / 4u;
    constexpr int NWR = ZB_WALK_NWR(P);
    u32 const xIntHi = xEnd >= (ZB_BATCH + 4u * NW) ? (xEnd + P - 4u * NW) & ~(ZB_BATCH - 1u) : 0u;     /* batches [x0, x0 + B) with x0 + B <= xIntHi are interior */

    /* the aligned words that hold the P + 7 bytes of a thread's positions, and the shift that realigns them: the words are
     * only touched (realigned) by the batch that uses them, two batches after the loads went out */
    auto fetch = [&](u32 xb, u32 (&a)[NW], u32& sh) {
        u32 const xa = xb + P * t;                                /* first coordinate of the thread */
#pragma unroll
        for (u32 k = 0; k < NW; k++) a[k] = 0u;
        sh = 0u;
        if (xb >= xIntLoB && xb + ZB_BATCH <= xIntHi) {           /* interior: no guards */
            const u8* const ad = fbase + (xa - shift);
            const u32* const p = reinterpret_cast<const u32*>((uintptr_t)ad & ~(uintptr_t)3);
            sh = ((u32)(uintptr_t)ad & 3u) * 8u;
#pragma unroll
            for (u32 k = 0; k < NW; k++) a[k] = __ldg(p + k);
            return;
        }
        if (xa + P <= xLow || xa >= xEnd) return;                 /* nothing of mine is walked */
        u32 const rel = xa - shift;
        bool const slow = (xa < xLow) || (D != 0u && rel < D && rel + P + 7u > D) || (rel + P + 7u > total);
        if (slow) return;                                         /* assembled byte-wise in the batch */
        const u8* const ad = (rel < D ? dbase : fbase) + rel;
        const u32* const p = reinterpret_cast<const u32*>((uintptr_t)ad & ~(uintptr_t)3);
        u32 const al = (u32)(uintptr_t)ad & 3u;
        sh = al * 8u;
        /* rel + P + 7 <= limit: a word is only touched when it holds one of the thread's P + 7 bytes */
#pragma unroll
        for (u32 k = 0; k < NW; k++) a[k] = (al + P + 7u > 4u * k) ? __ldg(p + k) : 0u;
    };
    u32 rawA[NW], rawB[NW], shA, shB;                             /* bytes of the next batch and of the one after it */
    fetch(x0, rawA, shA);
    fetch(x0 + ZB_BATCH, rawB, shB);
    u32 const blockMask = (1u << cd.blockLog) - 1u;
    /* insertion pattern: the residue of the thread's first position modulo insStep follows the walk by addition; only a
     * batch whose step was raised by the acceleration pays for a division */
    u32 const stepInc = ZB_BATCH % insStep;
    u32 r0 = (x0 + P * t + insStep * ZB_BATCH - shift) % insStep;
    u32 li = shift;                                               /* coordinate the acceleration counts from: the walk's start, then the end of the last batch with a hit */
    /* one batch: its words are realigned, the register set is refilled at once for the batch two ahead (no copies between
     * the sets: the loop below alternates them), then the three phases run */
    auto doBatch = [&](u32 (&raw)[NW], u32& sh) {
        u32 const xa = x0 + P * t;
        u32 cur[NWR];
#pragma unroll
        for (int k = 0; k < NWR; k++) cur[k] = __funnelshift_r(raw[k], (u32)k + 1u < NW ? raw[(u32)k + 1u < NW ? k + 1 : k] : 0u, sh);
        fetch(x0 + 2u * ZB_BATCH, raw, sh);                       /* in flight across two batches' barriers */
        if (D != 0u && x0 >= D + shift && li < D + shift) li = D + shift;   /* the frame starts with a fresh acceleration state behind a dictionary */
        u32 const sWalk = x0 > xLow ? x0 : xLow;                  /* first walked coordinate of the batch */
        u32 const step = insStep + ((sWalk - li) >> 7);
        u32 const pat = zb_walk_pattern_res<P>(step == insStep ? r0 : zb_walk_residue(xa - shift, step), step);
        bool const output = !buildImage && x0 >= H + shift;       /* H + shift is a batch border: the whole batch lies in the history or in the chunk */
        /* a batch never straddles two blocks (block sizes are multiples of the batch, or the frame is a single block) */
        u32 const qb0 = x0 - shift - H;                           /* offset of the batch in the chunk when output */
        size_t const idx = output ? (size_t)(cd.firstBlock - slotFirstBlock + (qb0 >> cd.blockLog)) * sd.dist + (qb0 & blockMask) + P * t : 0;
        bool hit;
        if (x0 >= xIntLoB && x0 + ZB_BATCH <= xIntHi)
            hit = zb_walk_batch<MLS, P, true>(table, xa, cur, pat, N, shift, D, total, xLow, xEnd, fbase, dbase, output, dist + idx, far + idx);
        else
            hit = zb_walk_batch<MLS, P, false>(table, xa, cur, pat, N, shift, D, total, xLow, xEnd, fbase, dbase, output, dist + idx, far + idx);
        if (hit) li = x0 + ZB_BATCH;
        r0 += stepInc; if (r0 >= insStep) r0 -= insStep;
        x0 += ZB_BATCH;
    };
    /* the same batch in the walk's steady state — this batch, the one whose words are requested and everything between are
     * interior batches of the frame's own bytes: none of doBatch's case distinctions apply, the words two batches ahead sit
     * 2 * ZB_BATCH bytes behind this batch's, the output row moves with the walk.  Same results as doBatch, about half
     * the instructions around the three phases. */
    const u8* const tbase = fbase + P * t - shift;                /* tbase + x0 = address of the thread's first byte in the batch at x0 */
    u16* const distT = dist + P * t;
    u32* const farT = far + P * t;
    auto steadyBatch = [&](u32 (&raw)[NW], u32 const sh) {
        u32 const xa = x0 + P * t;
        u32 cur[NWR];
#pragma unroll
        for (int k = 0; k < NWR; k++) cur[k] = __funnelshift_r(raw[k], (u32)k + 1u < NW ? raw[(u32)k + 1u < NW ? k + 1 : k] : 0u, sh);
        {   const u32* const pw = reinterpret_cast<const u32*>((uintptr_t)(tbase + x0 + 2u * ZB_BATCH) & ~(uintptr_t)3);
#pragma unroll
            for (u32 k = 0; k < NW; k++) raw[k] = __ldg(pw + k);  /* the alignment (sh) is the same in every batch: batches are 1024 bytes apart */
        }
        u32 const step = insStep + ((x0 - li) >> 7);
        u32 const pat = zb_walk_pattern_res<P>(step == insStep ? r0 : zb_walk_residue(xa - shift, step), step);
        bool const output = !buildImage && x0 >= H + shift;
        u32 const qb0 = x0 - shift - H;
        size_t const row = output ? (size_t)(cd.firstBlock - slotFirstBlock + (qb0 >> cd.blockLog)) * sd.dist + (qb0 & blockMask) : 0;
        if (zb_walk_batch<MLS, P, true>(table, xa, cur, pat, N, shift, D, total, xLow, xEnd, fbase, dbase, output, distT + row, farT + row)) li = x0 + ZB_BATCH;
        r0 += stepInc; if (r0 >= insStep) r0 -= insStep;
        x0 += ZB_BATCH;
    };
    while (x0 < xEnd) {
        /* pairs of steady-state batches (the two register sets keep their turns) */
        if (WALK_STEADY && x0 >= xIntLoB && x0 + 4u * ZB_BATCH <= xIntHi) {
            if (D != 0u && li < D + shift) li = D + shift;           /* as in doBatch: x0 >= xIntLoB >= D + shift */
            do { steadyBatch(rawA, shA); steadyBatch(rawB, shB); } while (x0 + 4u * ZB_BATCH <= xIntHi);
        }
        doBatch(rawA, shA);
        if (x0 >= xEnd) break;
        doBatch(rawB, shB);
    }
    if (buildImage) { __syncthreads(); for (u32 i = t; i < N; i += THREADS) imageOut[i] = table[i]; }
}

/* ------------------------------------------------------------------------------------------------
 * K1b — greedy selection + match extension.  One warp per 16 KiB segment, no shared memory (occupancy is
 * register-bound, so the L2 latency of the candidate checks is hidden by other warps).  Per step 32 probe
 * positions: pairs (p, p+1) spaced by `step` (zstd_fast.c:225-229, step acceleration :234,:342-347).  Hit
 * priority per lane: repcode-2 (lane 0, directly after a match, :410-420), repcode-1 (:281-297), table candidate
 * with 4-byte check (:102-141).  Lowest lane wins.  A segment owns the match starts inside it; a match may run
 * past the segment's end up to the block's end (the merge kernel resolves what that covers).
 * ---------------------------------------------------------------------------------------------- */
#ifndef PARSE_WARPS
#define PARSE_WARPS 8            /* = ZB_PARSE_SEGS: the eight segments of a full block share a CTA */
#endif
#ifndef PARSE_MIN_CTAS
#define PARSE_MIN_CTAS 6         /* 48 warps per SM at 40 registers (profiles/r1_history.md) */
#endif
__device__ __forceinline__ u32 zb_dist_at(const u16* __restrict__ d16, const u32* __restrict__ far, u32 i)
{
    u32 const d = d16[i];
    return d == ZB_FAR ? far[i] : d;
}

template <bool DICT>
__global__ void __launch_bounds__(32 * PARSE_WARPS, DICT ? (40 / PARSE_WARPS) : PARSE_MIN_CTAS)
zb_parse_kernel(const u8* __restrict__ src, const u8* __restrict__ dictEnd, const ZbBlock* __restrict__ blocks, u32 nbBlocks, ZbParams prm, ZbStrides sd,
                const u16* __restrict__ dist, const u32* __restrict__ far, u64* __restrict__ seqs, ZbBlockMeta* __restrict__ meta, ZbSegMeta* __restrict__ segmeta)
{
    u32 const lane = threadIdx.x & 31u;
    u32 const g = blockIdx.x * PARSE_WARPS + (threadIdx.x >> 5);  /* one warp per segment: the warps of a CTA share a block's history in L1/L2 */
    u32 const segs = (sd.dist + ZB_PARSE_SEG - 1u) / ZB_PARSE_SEG;   /* segments of the call's largest block (1 for calls of short frames) */
    u32 const b = g / segs, k = g % segs;
    if (b >= nbBlocks) return;
    ZbBlock const bd = blocks[b];
    u64* const myseq = seqs + (size_t)b * sd.seq + (size_t)k * (ZB_PARSE_SEG / 4u);
    const u16* const mydist = dist + (size_t)b * sd.dist;
    const u32* const myfar = far + (size_t)b * sd.dist;
    const u8* const base = src + bd.srcOff - bd.histLen;          /* base + rel addresses the frame's own bytes */
    u32 const bs = bd.histLen, blockEnd = bd.histLen + bd.size;
    ZbSeg sg; sg.hi = base; sg.lo = base; sg.split = 0;
    if (DICT && (bd.flags & ZB_FLAG_DICT)) { sg.lo = dictEnd - bd.dictLen; sg.split = bd.dictLen; }   /* oldest history = dictionary tail */

    if (bd.size < 7u) {                                        /* zstd_compress.c:3216 */
        if (lane == 0 && k == 0) {
            ZbBlockMeta m; m.nbSeq = 0; m.litSize = bd.size; m.litSecSize = 0; m.bodySize = bd.size;
            m.type = ZB_BT_RAW; m.forceRaw = 1; m.rleByte = 0; m.pad = 0;
            meta[b] = m;
        }
        return;
    }
    u32 const ss = bs + k * ZB_PARSE_SEG;                      /* this warp owns the match starts in [ss, se) */
    if (ss >= blockEnd) {
        if (lane == 0) { ZbSegMeta z; z.nbSeq = 0; z.pad[0] = z.pad[1] = z.pad[2] = 0; segmeta[(size_t)b * segs + k] = z; }
        return;
    }
    u32 const se = min(ss + ZB_PARSE_SEG, blockEnd);
    u32 const be = blockEnd;

    u32 ip = ss, anchor = ss, searchStart = ss;
    u32 rep1 = 0, rep2 = 0, nbSeq = 0;
    if ((bd.flags & ZB_FLAG_FIRST) && k == 0u) { rep1 = prm.startRep[0]; rep2 = prm.startRep[1]; }   /* a zstd-format dictionary's repcodes, zstd_compress.c:5054-5056 */

    while (ip < se && ip + 8u <= be) {
        u32 const step = prm.stepSize + ((ip - searchStart) >> 7);           /* kSearchStrength = 8 */
        u32 const p = ip + (lane >> 1) * step + (lane & 1u);
        bool const act = (p < se) && (p + 8u <= be);
        u32 const pp = act ? p : ip;                         /* a position every lane may load from (ip + 8 <= be) */
        /* the candidate row and the windows are independent loads: all of them go out before the first is looked at (the
         * rare far distance costs one more round trip, below) */
        u32 const d16 = act ? (u32)mydist[pp - bs] : 0u;
        bool const v3 = (lane == 0u) && (ip == anchor) && (rep2 != 0u);
        bool const v2 = act && rep1 != 0u && p >= rep1;
        /* dist[] only holds tag-verified candidates, so a step needs no random load: the current window and the
         * repcode window are contiguous across lanes.  The repcode window's 4 bytes in front of the position are only
         * needed by a lane whose repcode matched: asked for there */
        u32 pre, cur;
        zb_seg_pre_cur<DICT>(sg, pp, &pre, &cur);
        u32 const cur2 = DICT ? zb_seg_ld32<DICT>(sg, v2 ? pp - rep1 : pp) : zb_ld32w2(sg.hi + (v2 ? pp - rep1 : pp));
        u32 cur3 = ~cur;
        if (ip == anchor && rep2 != 0u) cur3 = (u32)zb_seg_ld64x<DICT>(sg, v3 ? pp - rep2 : pp);     /* warp-uniform condition */
        u32 const d = d16 == ZB_FAR ? myfar[pp - bs] : d16;
        bool const v1 = act && d != 0u && p >= d;
        u32 const hit = (v3 && cur3 == cur) ? 3u : ((v2 && cur2 == cur) ? 2u : (v1 ? 1u : 0u));
        /* backward catch-up (zstd_fast.c:387-391) of a repcode-1 hit: first 4 bytes in-lane from the windows */
        u32 myback = 0, mymore = 0;
        if (hit == 2u) {
            u32 pre2, unused;
            zb_seg_pre_cur<DICT>(sg, p - rep1, &pre2, &unused);
            u32 const x = pre ^ pre2;
            u32 const bm = x ? ((u32)__clz((int)x) >> 3) : 4u;
            u32 lim = p - anchor; lim = lim < 4u ? lim : 4u;
            u32 const src0 = p - rep1; lim = lim < src0 ? lim : src0;
            myback = bm < lim ? bm : lim;
            mymore = (bm == 4u && lim == 4u) ? 1u : 0u;
        }
        u32 tent = __ballot_sync(ZB_FULL, hit != 0u);
        /* lowest lane first.  A table hit (type 1) is only tag-verified by K1a: its bytes are checked while
         * the match is extended (one round trip, this lane only); a false positive drops out and the next
         * lane is tried — the result is "lowest lane whose hit is real", what the oracle computes. */
        u32 probe = 0, wtype = 0, offset = 0, back = 0, fwdFrom4 = 0;
        bool found = false;
        while (tent) {
            int const winner = __ffs((int)tent) - 1;
            probe = __shfl_sync(ZB_FULL, p, winner);
            wtype = __shfl_sync(ZB_FULL, hit, winner);
            u32 const wd = __shfl_sync(ZB_FULL, d, winner);
            back = __shfl_sync(ZB_FULL, myback, winner);
            u32 more = __shfl_sync(ZB_FULL, mymore, winner);
            offset = (wtype == 3u) ? rep2 : ((wtype == 2u) ? rep1 : wd);
            if (wtype == 1u) {
                u32 const f0 = zb_count_fwd<DICT>(sg, probe, offset, be, lane);          /* from the probe itself */
                if (f0 < 4u) { tent &= ~(1u << winner); continue; }                   /* tag collision */
                fwdFrom4 = f0 - 4u;
                more = 1u;
            } else {
                fwdFrom4 = zb_count_fwd<DICT>(sg, probe + 4u, offset, be, lane);
            }
            if (more) back += zb_back_coop<DICT>(sg, probe - back, offset, anchor, lane);   /* table hit, or a repcode hit with > 4 bytes of catch-up */
            found = true;
            break;
        }
        if (!found) { ip += 16u * step; continue; }
        u32 const ms = probe - back;
        u32 const mlen = back + 4u + fwdFrom4;
        if (wtype == 3u) { u32 const t = rep2; rep2 = rep1; rep1 = t; }
        else if (wtype == 1u) { rep2 = rep1; rep1 = offset; }
        if (lane == 0) myseq[nbSeq] = zb_pack_raw(offset, mlen, ms - bs);
        nbSeq++;
        ip = ms + mlen; anchor = ip; searchStart = ip;
    }
    if (lane == 0) { ZbSegMeta z; z.nbSeq = nbSeq; z.pad[0] = z.pad[1] = z.pad[2] = 0; segmeta[(size_t)b * segs + k] = z; }
}

/* ------------------------------------------------------------------------------------------------
 * K1b (doubleFast) — the greedy selection of ZSTD_compressBlock_doubleFast_noDict_generic
 * (zstd_double_fast.c:105-323) over two candidate arrays: distL (8-byte hash) and distS (mls-byte hash).
 * Per probe position p, in the reference's order: repcode-1 at p+1 (:190-195), long match at p
 * (:206-213), short match at p (:222-225) upgraded to the long match at p+1 when longer (:254-271);
 * probes are spaced by `step` (1, +1 per 256 bytes without a match, :131); lowest lane wins; immediate
 * repcode-2 at lane 0 right after a match (:302-316).  Table candidates are tag-verified only: the
 * winning lane's bytes are checked while the match is extended, a false positive drops out.
 * ---------------------------------------------------------------------------------------------- */
template <bool DICT>
__global__ void __launch_bounds__(32 * PARSE_WARPS)
zb_parse_dfast_kernel(const u8* __restrict__ src, const u8* __restrict__ dictEnd, const ZbBlock* __restrict__ blocks, u32 nbBlocks, ZbParams prm, ZbStrides sd,
                      const u16* __restrict__ distLong, const u32* __restrict__ farLong, const u16* __restrict__ distShort, const u32* __restrict__ farShort,
                      u64* __restrict__ seqs, ZbBlockMeta* __restrict__ meta, ZbSegMeta* __restrict__ segmeta)
{
    u32 const lane = threadIdx.x & 31u;
    u32 const g = blockIdx.x * PARSE_WARPS + (threadIdx.x >> 5);  /* one warp per segment, as in zb_parse_kernel */
    u32 const segs = (sd.dist + ZB_PARSE_SEG - 1u) / ZB_PARSE_SEG;
    u32 const b = g / segs, k = g % segs;
    if (b >= nbBlocks) return;
    ZbBlock const bd = blocks[b];
    u64* const myseq = seqs + (size_t)b * sd.seq + (size_t)k * (ZB_PARSE_SEG / 4u);
    const u16* const dLp = distLong + (size_t)b * sd.dist;
    const u32* const fLp = farLong + (size_t)b * sd.dist;
    const u16* const dSp = distShort + (size_t)b * sd.dist;
    const u32* const fSp = farShort + (size_t)b * sd.dist;
    const u8* const base = src + bd.srcOff - bd.histLen;
    u32 const bs = bd.histLen, blockEnd = bd.histLen + bd.size;
    ZbSeg sg; sg.hi = base; sg.lo = base; sg.split = 0;
    if (DICT && (bd.flags & ZB_FLAG_DICT)) { sg.lo = dictEnd - bd.dictLen; sg.split = bd.dictLen; }

    if (bd.size < 7u) {                                        /* zstd_compress.c:3216 */
        if (lane == 0 && k == 0) {
            ZbBlockMeta m; m.nbSeq = 0; m.litSize = bd.size; m.litSecSize = 0; m.bodySize = bd.size;
            m.type = ZB_BT_RAW; m.forceRaw = 1; m.rleByte = 0; m.pad = 0;
            meta[b] = m;
        }
        return;
    }
    u32 const ss = bs + k * ZB_PARSE_SEG;
    if (ss >= blockEnd) {
        if (lane == 0) { ZbSegMeta z; z.nbSeq = 0; z.pad[0] = z.pad[1] = z.pad[2] = 0; segmeta[(size_t)b * segs + k] = z; }
        return;
    }
    u32 const se = min(ss + ZB_PARSE_SEG, blockEnd);
    u32 const be = blockEnd;
    u32 ip = ss, anchor = ss, searchStart = ss;
    u32 rep1 = 0, rep2 = 0, nbSeq = 0;
    if ((bd.flags & ZB_FLAG_FIRST) && k == 0u) { rep1 = prm.startRep[0]; rep2 = prm.startRep[1]; }

    while (ip < se && ip + 9u <= be) {                        /* a lane reads 8 bytes at p and at p+1 */
        u32 const step = 1u + ((ip - searchStart) >> 8);                     /* kStepIncr = 1 << kSearchStrength */
        u32 const p = ip + lane * step;
        bool const act = (p < se) && (p + 9u <= be);
        u32 const pp = act ? p : ip;
        u32 dL = act ? zb_dist_at(dLp, fLp, pp - bs) : 0u;
        u32 dS = act ? zb_dist_at(dSp, fSp, pp - bs) : 0u;
        u32 dL1 = act ? zb_dist_at(dLp, fLp, pp + 1u - bs) : 0u;
        if (dL > pp) dL = 0u;                                  /* reaches past the visible history (window) */
        if (dS > pp) dS = 0u;
        if (dL1 > pp + 1u) dL1 = 0u;
        u64 const w = zb_seg_ld64x<DICT>(sg, pp);                            /* bytes p .. p+7 */
        u32 const cur = (u32)w, cur1 = (u32)(w >> 8);
        bool const v2 = act && rep1 != 0u && (p + 1u >= rep1);
        u32 const r2 = (u32)zb_seg_ld64x<DICT>(sg, v2 ? pp + 1u - rep1 : pp);
        u32 r3 = ~cur;
        bool const v3 = (lane == 0u) && (ip == anchor) && (rep2 != 0u);
        if (ip == anchor && rep2 != 0u) r3 = (u32)zb_seg_ld64x<DICT>(sg, v3 ? pp - rep2 : pp);
        /* 3 repcode-2, 2 repcode-1 (at p+1), 1 long candidate, 4 short candidate */
        u32 hit = (v3 && r3 == cur) ? 3u : ((v2 && r2 == cur1) ? 2u : (dL ? 1u : (dS ? 4u : 0u)));
        u32 tent = __ballot_sync(ZB_FULL, hit != 0u);
        u32 ms = 0, offset = 0, mlen = 0, wtype = 0;
        bool found = false;
        while (tent) {
            int const winner = __ffs((int)tent) - 1;
            u32 const probe = __shfl_sync(ZB_FULL, p, winner);
            wtype = __shfl_sync(ZB_FULL, hit, winner);
            u32 const wL = __shfl_sync(ZB_FULL, dL, winner), wS = __shfl_sync(ZB_FULL, dS, winner), wL1 = __shfl_sync(ZB_FULL, dL1, winner);
            if (wtype == 3u) { ms = probe; offset = rep2; mlen = 4u + zb_count_fwd<DICT>(sg, probe + 4u, rep2, be, lane); found = true; break; }
            if (wtype == 2u) { ms = probe + 1u; offset = rep1; mlen = 4u + zb_count_fwd<DICT>(sg, probe + 5u, rep1, be, lane); found = true; break; }
            if (wtype == 1u) {
                u32 const f0 = zb_count_fwd<DICT>(sg, probe, wL, be, lane);
                if (f0 >= 8u) {
                    u32 const back = zb_back_coop<DICT>(sg, probe, wL, anchor, lane);
                    ms = probe - back; offset = wL; mlen = back + f0; found = true; break;
                }
                /* tag collision on the long table: the lane may still have a short candidate */
                if (lane == (u32)winner) hit = dS ? 4u : 0u;
                if (wS == 0u) { tent &= ~(1u << winner); }
                continue;
            }
            /* short candidate */
            {   u32 const f0 = zb_count_fwd<DICT>(sg, probe, wS, be, lane);
                if (f0 < 4u) { tent &= ~(1u << winner); if (lane == (u32)winner) hit = 0u; continue; }
                u32 mp = probe, mo = wS, ml = f0;
                if (wL1) {
                    u32 const f1 = zb_count_fwd<DICT>(sg, probe + 1u, wL1, be, lane);
                    if (f1 >= 8u && f1 > ml) { mp = probe + 1u; mo = wL1; ml = f1; }
                }
                u32 const back = zb_back_coop<DICT>(sg, mp, mo, anchor, lane);
                ms = mp - back; offset = mo; mlen = back + ml; wtype = 1u; found = true; break;
            }
        }
        if (!found) { ip += 32u * step; continue; }
        if (wtype == 3u) { u32 const t = rep2; rep2 = rep1; rep1 = t; }
        else if (wtype == 1u) { rep2 = rep1; rep1 = offset; }
        if (lane == 0) myseq[nbSeq] = zb_pack_raw(offset, mlen, ms - bs);
        nbSeq++;
        ip = ms + mlen; anchor = ip; searchStart = ip;
    }
    if (lane == 0) { ZbSegMeta z; z.nbSeq = nbSeq; z.pad[0] = z.pad[1] = z.pad[2] = 0; segmeta[(size_t)b * segs + k] = z; }
}

/* ------------------------------------------------------------------------------------------------
 * K1c — joins a block's segments, assigns repcodes, materialises the literals.
 * `cur` = first byte of the block not yet covered by a sequence.  A raw sequence that ends at or before `cur` lies
 * under a match that ran over from an earlier segment and is dropped; one that straddles `cur` keeps its tail when
 * that is at least 3 bytes (MINMATCH, zstd_internal.h:102); the others take their literals from `cur`.
 * The repcode history (ZSTD_storeSeq / ZSTD_updateRep, zstd_compress_internal.h:671-760) then runs over the whole
 * block: it starts as {1,4,8} (or the dictionary's) in a frame's first block and unknown in every other block.
 * ---------------------------------------------------------------------------------------------- */
struct ZbRepHist { u32 r1, r2, r3; };
__device__ __forceinline__ u32 zb_rep_code(ZbRepHist& h, u32 off, u32 ll)
{
    u32 code;
    if (ll > 0u) {
        if (off == h.r1) return 1u;
        if (off == h.r2) { code = 2u; h.r2 = h.r1; h.r1 = off; return code; }
        if (off == h.r3) { code = 3u; h.r3 = h.r2; h.r2 = h.r1; h.r1 = off; return code; }
    } else {
        if (off == h.r2) { code = 1u; h.r2 = h.r1; h.r1 = off; return code; }
        if (off == h.r3) { code = 2u; h.r3 = h.r2; h.r2 = h.r1; h.r1 = off; return code; }
        if (h.r1 > 1u && off == h.r1 - 1u) { code = 3u; h.r3 = h.r2; h.r2 = h.r1; h.r1 = off; return code; }
    }
    h.r3 = h.r2; h.r2 = h.r1; h.r1 = off;
    return off + 3u;
}

#define MERGE_THREADS 256
#ifndef MERGE_MIN_CTAS
#define MERGE_MIN_CTAS 6           /* 40 registers (parse + merge 5.46 ms per GiB against 5.49 with 5 and 5.65 with 4) */
#endif
#define MERGE_TILE 1024u                       /* sequences scanned and gathered per round */
#define MERGE_PER (MERGE_TILE / MERGE_THREADS)  /* consecutive sequences of a tile owned by one thread */
#define SEG_SLOTS (ZB_PARSE_SEG / 4u)
__global__ void __launch_bounds__(MERGE_THREADS, MERGE_MIN_CTAS)
zb_merge_segments_kernel(const u8* __restrict__ src, const ZbBlock* __restrict__ blocks, ZbParams prm, ZbStrides sd, const ZbSegMeta* __restrict__ segmeta,
                         u64* __restrict__ seqs, u8* __restrict__ lits, ZbBlockMeta* __restrict__ meta)
{
    __shared__ u32 sPos[MERGE_TILE], sLit[MERGE_TILE], sLen[MERGE_TILE], sOff[MERGE_TILE];
    __shared__ u32 wsumL[MERGE_THREADS / 32], wsumA[MERGE_THREADS / 32], wmaxU[MERGE_THREADS / 32], wmaxK[MERGE_THREADS / 32];
    __shared__ u32 sR2[MERGE_TILE], sRep[3];
    __shared__ u32 baseL, baseA, carryEnd;
    __shared__ u32 gFirst[ZB_PARSE_SEGS], gCnt[ZB_PARSE_SEGS], gBase[ZB_PARSE_SEGS], gCur[ZB_PARSE_SEGS], gPm[ZB_PARSE_SEGS], gPl[ZB_PARSE_SEGS];
    __shared__ u32 gTotal;
    u32 const b = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    ZbBlock const bd = blocks[b];
    if (bd.size < 7u) return;                                    /* raw block: meta written by the parse kernel */
    u64* const myseq = seqs + (size_t)b * sd.seq;
    u8*  const mylit = lits + (size_t)b * sd.lit;
    const u8* const in = src + bd.srcOff;                        /* literals always lie inside the block itself */
    u32 const segs = (bd.size + ZB_PARSE_SEG - 1u) / ZB_PARSE_SEG;
    u32 const segStride = (sd.dist + ZB_PARSE_SEG - 1u) / ZB_PARSE_SEG;
    /* ---- 1. which raw sequences survive (warp 0; positions are relative to the block) ---- */
    if (warp == 0u) {
        u32 nb = 0; u64 r0 = 0, rl = 0;
        if (lane < segs) {
            nb = segmeta[(size_t)b * segStride + lane].nbSeq;
            if (nb) { r0 = myseq[(size_t)lane * SEG_SLOTS]; rl = myseq[(size_t)lane * SEG_SLOTS + nb - 1u]; }
        }
        u32 cur = 0, total = 0;
        for (u32 k = 0; k < segs; k++) {
            u32 const n = __shfl_sync(ZB_FULL, nb, (int)k);
            u64 r = __shfl_sync(ZB_FULL, r0, (int)k);
            u64 const last = __shfl_sync(ZB_FULL, rl, (int)k);
            u32 f = 0, pm = 0, pl = 0;
            while (f < n) {
                u32 const ms = ZB_RAW_MS(r), ml = ZB_RAW_MLEN(r);
                if (ms + ml <= cur || (ms < cur && ms + ml - cur < 3u)) { f++; if (f < n) r = myseq[(size_t)k * SEG_SLOTS + f]; continue; }
                if (ms < cur) { pm = cur; pl = ms + ml - cur; } else { pm = ms; pl = ml; }
                break;
            }
            if (lane == 0u) { gFirst[k] = f; gCnt[k] = n - f; gBase[k] = total; gCur[k] = cur; gPm[k] = pm; gPl[k] = pl; }
            if (f < n) { cur = (f == n - 1u) ? pm + pl : ZB_RAW_MS(last) + ZB_RAW_MLEN(last); total += n - f; }
        }
        if (lane == 0u) gTotal = total;
    }
    __syncthreads();
    /* ---- 2. survivors move down to be contiguous (in place: a destination never lies above its source) and become
     *         (offset, litLength, matchLength) ---- */
#pragma unroll 1
    for (u32 k = 0; k < segs; k++) {
        u32 const f = gFirst[k], cnt = gCnt[k];
        const u64* const sfrom = myseq + (size_t)k * SEG_SLOTS + f;
        u64* const sto = myseq + gBase[k];
        for (u32 c0 = 0; c0 < cnt; c0 += MERGE_THREADS) {
            u32 const i = c0 + tid;
            u64 r = 0; u32 ms = 0, ml = 0;
            if (i < cnt) { r = sfrom[i]; ms = ZB_RAW_MS(r); ml = ZB_RAW_MLEN(r); if (i == 0u) { ms = gPm[k]; ml = gPl[k]; } }
            u32 const myEnd = ms + ml;
            u32 prevEnd = __shfl_up_sync(ZB_FULL, myEnd, 1);
            if (lane == 31u) wsumL[warp] = myEnd;
            __syncthreads();
            if (lane == 0u) prevEnd = warp ? wsumL[warp - 1u] : (c0 ? carryEnd : gCur[k]);
            __syncthreads();
            if (i < cnt) sto[i] = zb_pack_seq(ZB_RAW_OFF(r), ms - prevEnd, ml);
            if (tid == MERGE_THREADS - 1u) carryEnd = myEnd;
            __syncthreads();
        }
    }
    u32 const nbSeq = gTotal;
    if (tid == 0) {
        baseL = 0; baseA = 0;
        bool const first = (bd.flags & ZB_FLAG_FIRST) != 0u;
        sRep[0] = first ? prm.codeRep[0] : 0u; sRep[1] = first ? prm.codeRep[1] : 0u; sRep[2] = first ? prm.codeRep[2] : 0u;
    }
    __syncthreads();
    /* ---- 3. literals + repcodes, a tile of sequences at a time ----
     * The repcode history (r1, r2, r3) is a serial recurrence in ZSTD_updateRep's form, but its solution is not:
     *   - after any sequence r1 is that sequence's offset, so "r1 before sequence i" is the offset of sequence i-1;
     *   - a sequence leaves the history alone (U) iff it has literals and repeats r1; every other sequence sets r2 to the r1
     *     it found: "r2 before i" is the r1 found by the last non-U sequence before i;
     *   - a non-U sequence whose offset is the r2 it found swaps r1 and r2 and keeps r3 (K = U or swap); every other sets r3
     *     to the r2 it found: "r3 before i" is the r2 found by the last non-K sequence before i.
     * Two "index of the last flagged element before me" scans (maximum scans) over the tile give every sequence the history
     * it meets; its code follows from ZSTD_storeSeq's rules.  The history passes from tile to tile through sRep. */
    for (u32 t0 = 0; t0 < nbSeq; t0 += MERGE_TILE) {
        u32 const n = min(MERGE_TILE, nbSeq - t0);
        /* every thread owns MERGE_PER consecutive sequences of the tile */
        u32 ll[MERGE_PER], adv[MERGE_PER], ml[MERGE_PER], off[MERGE_PER], myL = 0, myA = 0;
#pragma unroll
        for (u32 j = 0; j < MERGE_PER; j++) {
            u32 const i = tid * MERGE_PER + j;
            u64 const q = (i < n) ? myseq[t0 + i] : 0ull;
            ll[j] = (u32)((q >> 24) & 0x3FFFFu);
            ml[j] = (u32)(q >> 42);
            adv[j] = ll[j] + ml[j];
            off[j] = (u32)q & 0xFFFFFFu;
            if (i < n) sOff[i] = off[j];
            myL += ll[j]; myA += adv[j];
        }
        u32 inL = myL, inA = myA;                                /* inclusive scan over the warp, then over the warps */
#pragma unroll
        for (u32 o = 1; o < 32u; o <<= 1) {
            u32 const a = __shfl_up_sync(ZB_FULL, inL, o), c = __shfl_up_sync(ZB_FULL, inA, o);
            if (lane >= o) { inL += a; inA += c; }
        }
        if (lane == 31u) { wsumL[warp] = inL; wsumA[warp] = inA; }
        u32 const R1 = sRep[0], R2 = sRep[1], R3 = sRep[2];      /* history at the tile's start */
        __syncthreads();
        u32 offL = baseL + inL - myL, offA = baseA + inA - myA;
        for (u32 w = 0; w < warp; w++) { offL += wsumL[w]; offA += wsumA[w]; }
#pragma unroll
        for (u32 j = 0; j < MERGE_PER; j++) {
            u32 const i = tid * MERGE_PER + j;
            if (i < n) { sPos[i] = offA; sLit[i] = offL; sLen[i] = ll[j]; }
            offL += ll[j]; offA += adv[j];
        }
        /* r1 before each of my sequences, U flags, first scan */
        u32 prevOff[MERGE_PER], r2b[MERGE_PER], r3b[MERGE_PER], bef[MERGE_PER];
        bool U[MERGE_PER], K[MERGE_PER];
        u32 run = 0;
#pragma unroll
        for (u32 j = 0; j < MERGE_PER; j++) {
            u32 const i = tid * MERGE_PER + j;
            prevOff[j] = j ? off[j - 1u] : (i == 0u ? R1 : ((i < n) ? sOff[i - 1u] : 0u));
            U[j] = ll[j] > 0u && off[j] == prevOff[j];
            bef[j] = run;                                            /* 1 + index of the last non-U sequence of mine before this one */
            if (i < n && !U[j]) run = i + 1u;
        }
        {   u32 inc = run;
#pragma unroll
            for (u32 o = 1; o < 32u; o <<= 1) { u32 const x = __shfl_up_sync(ZB_FULL, inc, o); if (lane >= o) inc = max(inc, x); }
            if (lane == 31u) wmaxU[warp] = inc;
            u32 ex = __shfl_up_sync(ZB_FULL, inc, 1); if (lane == 0u) ex = 0u;
            __syncthreads();                                         /* also: sPos / sLit / sLen / sOff of the tile are complete */
            for (u32 w = 0; w < warp; w++) ex = max(ex, wmaxU[w]);
#pragma unroll
            for (u32 j = 0; j < MERGE_PER; j++) {
                u32 const i = tid * MERGE_PER + j;
                u32 const m = max(bef[j], ex);                       /* 1 + index of the last non-U sequence before i, 0: none in this tile */
                r2b[j] = m == 0u ? R2 : (m == 1u ? R1 : sOff[m - 2u]);   /* the r1 that sequence found */
                if (i < n) sR2[i] = r2b[j];
            }
        }
        if (tid == MERGE_THREADS - 1u) { baseL = offL; baseA = offA; }     /* totals up to the end of this tile */
        run = 0;
#pragma unroll
        for (u32 j = 0; j < MERGE_PER; j++) {
            u32 const i = tid * MERGE_PER + j;
            K[j] = U[j] || off[j] == r2b[j];
            bef[j] = run;
            if (i < n && !K[j]) run = i + 1u;
        }
        {   u32 inc = run;
#pragma unroll
            for (u32 o = 1; o < 32u; o <<= 1) { u32 const x = __shfl_up_sync(ZB_FULL, inc, o); if (lane >= o) inc = max(inc, x); }
            if (lane == 31u) wmaxK[warp] = inc;
            u32 ex = __shfl_up_sync(ZB_FULL, inc, 1); if (lane == 0u) ex = 0u;
            __syncthreads();                                         /* also: sR2 of the tile is complete, every thread holds R1..R3 */
            for (u32 w = 0; w < warp; w++) ex = max(ex, wmaxK[w]);
#pragma unroll
            for (u32 j = 0; j < MERGE_PER; j++) {
                u32 const i = tid * MERGE_PER + j;
                u32 const m = max(bef[j], ex);
                r3b[j] = m == 0u ? R3 : sR2[m - 1u];                 /* the r2 that sequence found */
                bool const swp = !U[j] && off[j] == r2b[j];
                u32 c = off[j] + 3u;
                if (ll[j] > 0u) { if (U[j]) c = 1u; else if (swp) c = 2u; else if (off[j] == r3b[j]) c = 3u; }
                else            { if (swp) c = 1u; else if (off[j] == r3b[j]) c = 2u; else if (prevOff[j] > 1u && off[j] == prevOff[j] - 1u) c = 3u; }
                if (i < n) myseq[t0 + i] = zb_pack_seq(c, ll[j], ml[j]);
                if (i == n - 1u) { sRep[0] = off[j]; sRep[1] = U[j] ? r2b[j] : prevOff[j]; sRep[2] = K[j] ? r3b[j] : r2b[j]; }   /* read again only behind the tile's last barrier */
            }
        }
        /* the tile's literal bytes [L0, L1) of the block's literal buffer, 8 at a time per thread: the run that holds
         * a group's first byte is found by bisection over the runs' start offsets, later bytes step to the next
         * non-empty run; all of a thread's loads are independent of one another.  Full groups leave as one 8-byte
         * store (the buffer is 16-byte aligned), the partial groups at the tile's edges byte by byte. */
        {   u32 const L0 = sLit[0], L1 = sLit[n - 1u] + sLen[n - 1u];
            for (u32 g = (L0 >> 3) + tid; (g << 3) < L1; g += MERGE_THREADS) {
                u32 const jb = g << 3;
                u32 const j0 = jb > L0 ? jb : L0, j1 = jb + 8u < L1 ? jb + 8u : L1;
                u32 sq = 0;                                              /* largest index with sLit[sq] <= j0 (sLit[0] = L0 <= j0) */
#pragma unroll
                for (u32 stp = MERGE_TILE / 2u; stp > 0u; stp >>= 1) {
                    u32 const c = sq + stp;
                    if (c < n && sLit[c] <= j0) sq = c;
                }
                u32 runEnd = sLit[sq] + sLen[sq];
                const u8* from = in + (sPos[sq] - sLit[sq]);             /* from[j] = the literal at buffer offset j while j lies in run sq */
                u64 v = 0;
#pragma unroll
                for (u32 k = 0; k < 8u; k++) {
                    u32 const j = jb + k;
                    if (j >= j0 && j < j1) {
                        while (j >= runEnd) { sq++; runEnd = sLit[sq] + sLen[sq]; from = in + (sPos[sq] - sLit[sq]); }
                        v |= (u64)from[j] << (8u * k);
                    }
                }
                if (j1 - j0 == 8u) *reinterpret_cast<u64*>(mylit + jb) = v;
                else for (u32 j = j0; j < j1; j++) mylit[j] = (u8)(v >> (8u * (j - jb)));
            }
        }
        __syncthreads();                                             /* the shared arrays are free for the next tile, sRep is its history */
    }
    u32 const litSeq = baseL, consumed = baseA;                   /* literals in sequences, bytes covered by sequences */
    u32 const lastLits = bd.size - consumed;
    for (u32 x = tid; x < lastLits; x += MERGE_THREADS) mylit[litSeq + x] = in[consumed + x];
    /* ---- 4. meta ---- */
    if (tid == 0) {
        ZbBlockMeta m; m.nbSeq = nbSeq; m.litSize = litSeq + lastLits; m.litSecSize = 0; m.bodySize = 0;
        m.type = ZB_BT_COMPRESSED; m.forceRaw = 0; m.rleByte = 0; m.pad = 0;
        meta[b] = m;
    }
}

/* K1c for calls made of short frames (one segment per block: nothing to join): one warp per block, 8 blocks per CTA;
 * every lane converts and gathers the literals of its own sequences, the repcode history runs through the warp. */
__global__ void __launch_bounds__(MERGE_THREADS)
zb_merge_small_kernel(const u8* __restrict__ src, const ZbBlock* __restrict__ blocks, u32 nbBlocks, ZbParams prm, ZbStrides sd, const ZbSegMeta* __restrict__ segmeta,
                      u64* __restrict__ seqs, u8* __restrict__ lits, ZbBlockMeta* __restrict__ meta)
{
    u32 const lane = threadIdx.x & 31u;
    u32 const b = blockIdx.x * (MERGE_THREADS / 32u) + (threadIdx.x >> 5);
    if (b >= nbBlocks) return;
    ZbBlock const bd = blocks[b];
    if (bd.size < 7u) return;
    u64* const myseq = seqs + (size_t)b * sd.seq;
    u8* const mylit = lits + (size_t)b * sd.lit;
    const u8* const in = src + bd.srcOff;
    u32 const nbSeq = segmeta[b].nbSeq;
    ZbRepHist hist; hist.r1 = 0; hist.r2 = 0; hist.r3 = 0;
    if (bd.flags & ZB_FLAG_FIRST) { hist.r1 = prm.codeRep[0]; hist.r2 = prm.codeRep[1]; hist.r3 = prm.codeRep[2]; }
    u32 posL = 0, posA = 0;
    for (u32 t0 = 0; t0 < nbSeq; t0 += 32u) {
        u32 const i = t0 + lane;
        u64 const r = (i < nbSeq) ? myseq[i] : 0ull;
        u32 const ms = ZB_RAW_MS(r), ml = ZB_RAW_MLEN(r), off = ZB_RAW_OFF(r);
        u32 const myEnd = (i < nbSeq) ? ms + ml : 0u;
        u32 prevEnd = __shfl_up_sync(ZB_FULL, myEnd, 1);
        if (lane == 0u) prevEnd = posA;
        u32 const ll = (i < nbSeq) ? ms - prevEnd : 0u;
        u32 const adv = ll + ((i < nbSeq) ? ml : 0u);
        u32 inL = ll;
#pragma unroll
        for (u32 o = 1; o < 32u; o <<= 1) { u32 const x = __shfl_up_sync(ZB_FULL, inL, o); if (lane >= o) inL += x; }
        /* repcodes: the history is uniform across the warp, lane j keeps sequence j's code */
        u32 code = 0;
        u32 const cnt = min(32u, nbSeq - t0);
        for (u32 j = 0; j < cnt; j++) {
            u32 const o = __shfl_sync(ZB_FULL, off, (int)j), l = __shfl_sync(ZB_FULL, ll, (int)j);
            u32 const c = zb_rep_code(hist, o, l);
            if (lane == j) code = c;
        }
        if (i < nbSeq) {
            const u8* const from = in + prevEnd;
            u8* const to = mylit + posL + inL - ll;
            for (u32 x = 0; x < ll; x++) to[x] = from[x];
            myseq[i] = zb_pack_seq(code, ll, ml);
        }
        posL += __shfl_sync(ZB_FULL, inL, 31);
        posA = __shfl_sync(ZB_FULL, myEnd, (int)(cnt - 1u));
        (void)adv;
    }
    u32 const lastLits = bd.size - posA;
    for (u32 x = lane; x < lastLits; x += 32u) mylit[posL + x] = in[posA + x];
    if (lane == 0) {
        ZbBlockMeta m; m.nbSeq = nbSeq; m.litSize = posL + lastLits; m.litSecSize = 0; m.bodySize = 0;
        m.type = ZB_BT_COMPRESSED; m.forceRaw = 0; m.rleByte = 0; m.pad = 0;
        meta[b] = m;
    }
}

/* ------------------------------------------------------------------------------------------------ launchers */
/* Positions per thread (P) follow the table size: the table decides how many walk CTAs fit an SM (227 KiB of shared
 * memory), and ZB_BATCH / P threads per CTA keep about 32 warps resident in every case — fast tables (<= 56 KiB): 4 CTAs of
 * 256 threads; <= 113 KiB: 2 CTAs of 512; the doubleFast tables (128 / 200 KiB): one CTA of 1024.  The result does not
 * depend on P (a batch is ZB_BATCH positions whatever the thread count). */
template <int MLS, int P>
static cudaError_t zb_launch_walk_p(const u8* d_src, const u8* d_dictEnd, const ZbChunk* d_chunks, u32 nbChunks, u32 N, u32 insStep, const ZbStrides& sd,
                                    u32 slotFirstBlock, u16* d_dist, u32* d_far, const u32* d_imageIn, u32* d_imageOut, cudaStream_t stream)
{
    cudaError_t const e = cudaFuncSetAttribute(zb_walk_kernel<MLS, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    zb_walk_kernel<MLS, P><<<nbChunks, ZB_BATCH / P, (size_t)N * 4u, stream>>>(d_src, d_dictEnd, d_chunks, insStep, N, sd, slotFirstBlock, d_dist, d_far, d_imageIn, d_imageOut);
    return cudaGetLastError();
}
template <int MLS>
static cudaError_t zb_launch_walk_m(const u8* d_src, const u8* d_dictEnd, const ZbChunk* d_chunks, u32 nbChunks, u32 N, u32 insStep, const ZbStrides& sd,
                                    u32 slotFirstBlock, u16* d_dist, u32* d_far, const u32* d_imageIn, u32* d_imageOut, cudaStream_t stream)
{
    size_t const smem = (size_t)N * 4u;
    if (smem <= 56u * 1024u)  return zb_launch_walk_p<MLS, WALK_P_SMALL>(d_src, d_dictEnd, d_chunks, nbChunks, N, insStep, sd, slotFirstBlock, d_dist, d_far, d_imageIn, d_imageOut, stream);
    if (smem <= 113u * 1024u) return zb_launch_walk_p<MLS, WALK_P_MID>(d_src, d_dictEnd, d_chunks, nbChunks, N, insStep, sd, slotFirstBlock, d_dist, d_far, d_imageIn, d_imageOut, stream);
    return zb_launch_walk_p<MLS, 1>(d_src, d_dictEnd, d_chunks, nbChunks, N, insStep, sd, slotFirstBlock, d_dist, d_far, d_imageIn, d_imageOut, stream);
}
static cudaError_t zb_launch_walk(const u8* d_src, const u8* d_dictEnd, const ZbChunk* d_chunks, u32 nbChunks, u32 mls, u32 N, u32 insStep, const ZbStrides& sd,
                                  u32 slotFirstBlock, u16* d_dist, u32* d_far, const u32* d_imageIn, u32* d_imageOut, cudaStream_t stream)
{
    switch (mls) {
    case 4: return zb_launch_walk_m<4>(d_src, d_dictEnd, d_chunks, nbChunks, N, insStep, sd, slotFirstBlock, d_dist, d_far, d_imageIn, d_imageOut, stream);
    case 5: return zb_launch_walk_m<5>(d_src, d_dictEnd, d_chunks, nbChunks, N, insStep, sd, slotFirstBlock, d_dist, d_far, d_imageIn, d_imageOut, stream);
    case 6: return zb_launch_walk_m<6>(d_src, d_dictEnd, d_chunks, nbChunks, N, insStep, sd, slotFirstBlock, d_dist, d_far, d_imageIn, d_imageOut, stream);
    case 7: return zb_launch_walk_m<7>(d_src, d_dictEnd, d_chunks, nbChunks, N, insStep, sd, slotFirstBlock, d_dist, d_far, d_imageIn, d_imageOut, stream);
    default: return zb_launch_walk_m<8>(d_src, d_dictEnd, d_chunks, nbChunks, N, insStep, sd, slotFirstBlock, d_dist, d_far, d_imageIn, d_imageOut, stream);
    }
}

/* one-CTA launches that walk the dictionary tail and store the table(s) in d_image: prm->tableN u32 of the (short) table,
 * followed for doubleFast by prm->tableNLong u32 of the 8-byte-hash table */
extern "C" cudaError_t zb_launch_dict_image(const u8* d_dictEnd, const ZbChunk* d_dictChunk, const ZbParams* prm, u32* d_image, cudaStream_t stream)
{
    ZbStrides sd; sd.dist = ZB_BLOCK_MAX; sd.seq = ZB_SEQ_STRIDE; sd.lit = ZB_LIT_STRIDE; sd.body = ZB_BODY_STRIDE; sd.state = ZB_STATE_STRIDE;   /* unused: no block is walked */
    cudaError_t e = zb_launch_walk(nullptr, d_dictEnd, d_dictChunk, 1, prm->mls, prm->tableN, prm->insStep, sd, 0, nullptr, nullptr, nullptr, d_image, stream);
    if (e == cudaSuccess && prm->strategy == 2)
        e = zb_launch_walk(nullptr, d_dictEnd, d_dictChunk, 1, 8, prm->tableNLong, prm->insStep, sd, 0, nullptr, nullptr, nullptr, d_image + prm->tableN, stream);
    return e;
}

extern "C" cudaError_t zb_launch_match(const u8* d_src, const u8* d_dictEnd, const u32* d_image, const ZbBlock* d_blocks, u32 nbBlocks,
                                       const ZbChunk* d_chunks, u32 nbChunks, u32 slotFirstBlock, const ZbParams* prm, const ZbStrides* sdp,
                                       u16* d_dist, u32* d_far, u16* d_dist2, u32* d_far2, u64* d_seqs, u8* d_lits, ZbBlockMeta* d_meta, ZbSegMeta* d_segmeta,
                                       cudaEvent_t evMid, cudaStream_t stream)
{
    if (nbBlocks == 0) return cudaSuccess;
    ZbStrides const sd = *sdp;
    cudaError_t e;
    u32 const segs = (sd.dist + ZB_PARSE_SEG - 1u) / ZB_PARSE_SEG;
    u32 const sgrid = (u32)(((u64)nbBlocks * segs + PARSE_WARPS - 1) / PARSE_WARPS);                       /* one warp per segment */
    if (prm->strategy == 2) {
        /* doubleFast: one candidate walk per table */
        e = zb_launch_walk(d_src, d_dictEnd, d_chunks, nbChunks, 8, prm->tableNLong, prm->insStep, sd, slotFirstBlock, d_dist, d_far, d_image ? d_image + prm->tableN : nullptr, nullptr, stream); if (e != cudaSuccess) return e;
        e = zb_launch_walk(d_src, d_dictEnd, d_chunks, nbChunks, prm->mls, prm->tableN, prm->insStep, sd, slotFirstBlock, d_dist2, d_far2, d_image, nullptr, stream); if (e != cudaSuccess) return e;
        if (evMid) cudaEventRecord(evMid, stream);
        if (d_dictEnd) zb_parse_dfast_kernel<true><<<sgrid, 32 * PARSE_WARPS, 0, stream>>>(d_src, d_dictEnd, d_blocks, nbBlocks, *prm, sd, d_dist, d_far, d_dist2, d_far2, d_seqs, d_meta, d_segmeta);
        else           zb_parse_dfast_kernel<false><<<sgrid, 32 * PARSE_WARPS, 0, stream>>>(d_src, nullptr, d_blocks, nbBlocks, *prm, sd, d_dist, d_far, d_dist2, d_far2, d_seqs, d_meta, d_segmeta);
    } else {
        e = zb_launch_walk(d_src, d_dictEnd, d_chunks, nbChunks, prm->mls, prm->tableN, prm->insStep, sd, slotFirstBlock, d_dist, d_far, d_image, nullptr, stream); if (e != cudaSuccess) return e;
        if (evMid) cudaEventRecord(evMid, stream);
        if (d_dictEnd) zb_parse_kernel<true><<<sgrid, 32 * PARSE_WARPS, 0, stream>>>(d_src, d_dictEnd, d_blocks, nbBlocks, *prm, sd, d_dist, d_far, d_seqs, d_meta, d_segmeta);
        else           zb_parse_kernel<false><<<sgrid, 32 * PARSE_WARPS, 0, stream>>>(d_src, nullptr, d_blocks, nbBlocks, *prm, sd, d_dist, d_far, d_seqs, d_meta, d_segmeta);
    }
    if (segs == 1u && sd.dist <= 8192u)
        zb_merge_small_kernel<<<(nbBlocks + MERGE_THREADS / 32u - 1u) / (MERGE_THREADS / 32u), MERGE_THREADS, 0, stream>>>(d_src, d_blocks, nbBlocks, *prm, sd, d_segmeta, d_seqs, d_lits, d_meta);
    else
        zb_merge_segments_kernel<<<nbBlocks, MERGE_THREADS, 0, stream>>>(d_src, d_blocks, *prm, sd, d_segmeta, d_seqs, d_lits, d_meta);
    return cudaGetLastError();
}
