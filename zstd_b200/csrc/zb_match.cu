/* zb_match.cu — K1: LZ77 match-finder ("fast" and "doubleFast" strategies) = K1a candidate walk, K1b greedy parse, K1c merge.
 *
 * Replaces the CPU loops ZSTD_compressBlock_fast_noDict_generic / _extDict_generic
 * (/root/reference/lib/compress/zstd_fast.c:192-423, :709-960) and ZSTD_compressBlock_doubleFast_noDict_generic
 * (zstd_double_fast.c:105-323) with a data-parallel formulation; a <=128 KiB block is the independent unit:
 *   - K1a (one warp per block) walks a private hash table in shared memory (2^hashLog buckets: 16-bit position
 *     modulo 64 KiB relative to the start of the visible history + 8-bit tag), primed from the <=64 KiB in front of
 *     the block (zstd_fast.c:53-85 does this for dictionaries, zstdmt_compress.c:726-731 for job overlaps), and
 *     records every position's candidate distance; insertion follows a fixed position pattern, so the walk does not
 *     depend on the parse;
 *   - K1b (one warp per 16 KiB segment of the block) does the greedy selection: 32 probe positions per step — pairs
 *     (p, p+1) spaced by `step` as in zstd_fast.c:225-229 — lowest matching lane wins (warp ballot); backward
 *     catch-up (:387-391) and forward extension (ZSTD_count, zstd_compress_internal.h:771) are warp-cooperative:
 *     32 x 8 bytes per round, first differing lane found by ballot.  It emits packed sequences only;
 *   - K1c (one CTA per block) joins the segments' sequences and gathers the literal bytes from the input.
 * Table writes are deterministic: the highest inserted lane of a step wins a bucket.
 * The bit-exact CPU model of these kernels is oracle/zb_match.c (tests only).
 */
#include <cuda_pipeline.h>
#include "zb_device.cuh"
#include "zb_kernels.h"

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

__device__ __forceinline__ u64 zb_pack_seq(u32 offBase, u32 litLen, u32 matchLen)
{
    return (u64)offBase | ((u64)litLen << 24) | ((u64)matchLen << 42);
}

/* ------------------------------------------------------------------------------------------------
 * K1a — candidate table walk (parse-independent).  One warp per block, table in shared memory.
 * For every position p of the block: dist[p] = distance to the most recent earlier position that was
 * inserted and has the same hash (0 = none).  32 consecutive positions per step; within a step the
 * sequential semantics are kept (a lane sees the inserted lanes below it, the highest inserted lane
 * of a hash group updates the table) with a write / read-back of the bucket; __match_any_sync would do
 * it in one instruction but costs ~400 cycles on sm_100a (profiles/r1_cand_match_any.txt).  No load depends on the table, so input
 * loads are issued one step ahead and the loop-carried chain is LDS -> STS only.
 * ---------------------------------------------------------------------------------------------- */
#define CAND_CHUNK 512u                      /* bytes staged per cp.async group: 32 lanes x 16 B */
#define CAND_STAGES 4u                       /* ring = 4 chunks = 2 KiB, 3 chunks in flight ahead of the consumer */
#define CAND_RING (CAND_CHUNK * CAND_STAGES)

/* One 16-step slice of the table walk (phase B of zb_cand_kernel).  INTERIOR: every lane of every
 * step is an in-block position with 8 readable bytes, so the activity predicates fold away.
 * table/tags are the warp's communication medium (lanes read what other lanes wrote in the same step):
 * volatile, never __restrict__. */
template <bool INTERIOR>
__device__ __forceinline__ void zb_cand_walk16(volatile u16* table, volatile u8* tags, const u32* hh,
                                               u32 q0, u32 o0, u32 pmin, u32 nPos, u32 bs, u32 lane,
                                               u32& ph, u32 inc, u32 period, u16* __restrict__ mydist)
{
    /* The bucket of step j+1 is read in the same shared-memory round as the read-back of step j (nothing writes
     * the table in between, except the rare peel below, which reads it again): one LDS round trip per step
     * instead of two on the loop-carried chain. */
    bool act_c = INTERIOR ? true : ((q0 + lane >= o0 + pmin) && (q0 + lane - o0 < nPos));
    u32 old_c = act_c ? table[hh[0] >> 8] : 0u;
    u32 oldtag_c = act_c ? tags[hh[0] >> 8] : 0x100u;
#pragma unroll
    for (u32 j = 0; j < CAND_CHUNK / 32u; j++) {
        u32 const q = q0 + 32u * j + lane;
        bool const act = act_c;
        u32 const p = q - o0;
        u32 const h = act ? (hh[j] >> 8) : 0u;
        u32 const tag = hh[j] & 0xFFu;
        bool const ins = act && ph < 2u;
        u32 const old = old_c, oldtag = oldtag_c;
        /* let every inserting lane write its bucket, read it back: when no two inserting lanes share a bucket
         * (the common case) the read-back alone resolves the step */
        __syncwarp();
        if (ins) { table[h] = (u16)p; tags[h] = (u8)tag; }
        __syncwarp();
        u32 const nw = act ? table[h] : 0u;
        u32 const nwtag = act ? tags[h] : 0x100u;
        u32 hn = 0; bool act_n = false;
        if (j + 1u < CAND_CHUNK / 32u) {
            u32 const qn = q + 32u;
            act_n = INTERIOR ? true : ((qn >= o0 + pmin) && (qn - o0 < nPos));
            hn = act_n ? (hh[(j + 1u) % (CAND_CHUNK / 32u)] >> 8) : 0u;
            old_c = act_n ? table[hn] : 0u;
            oldtag_c = act_n ? tags[hn] : 0x100u;
        }
        u32 losers = __ballot_sync(ZB_FULL, ins && nw != (p & 0xFFFFu));
        u32 d = 0;
        bool resolved = false;
        if (losers) {
            /* rare: some bucket has several inserting lanes.  Peel one hash group per round:
             * the highest inserting lane owns the bucket, every lane of the group takes the
             * nearest inserting lane below it as its candidate. */
            u32 const insmask = __ballot_sync(ZB_FULL, ins);
            while (losers) {
                int const L = __ffs((int)losers) - 1;
                u32 const hl = __shfl_sync(ZB_FULL, h, L);
                bool const mine = act && h == hl;
                u32 const grpAll = __ballot_sync(ZB_FULL, mine);
                u32 const grpIns = grpAll & insmask;
                u32 const lower = mine ? (grpIns & ((1u << lane) - 1u)) : 0u;
                u32 const lowLane = lower ? (31u - (u32)__clz((int)lower)) : lane;
                u32 const lowTag = __shfl_sync(ZB_FULL, tag, (int)lowLane);
                if (mine) {
                    if (lower) d = (lowTag == tag) ? lane - lowLane : 0u;
                    else { d = (p - old) & 0xFFFFu; if (d > p || oldtag != tag) d = 0u; }
                    resolved = true;
                    if ((31u - (u32)__clz((int)grpIns)) == lane) { table[h] = (u16)p; tags[h] = (u8)tag; }
                }
                losers &= ~grpAll;
            }
            __syncwarp();
            if (j + 1u < CAND_CHUNK / 32u) {                    /* the peel rewrote buckets: read the next step's again */
                old_c = act_n ? table[hn] : 0u;
                oldtag_c = act_n ? tags[hn] : 0x100u;
            }
        }
        if (!resolved) {
            u32 const dn = (p - nw) & 0xFFFFu;                  /* written by a lane below me in this step? */
            if (dn >= 1u && dn <= lane) d = (nwtag == tag) ? dn : 0u;
            else { d = (p - old) & 0xFFFFu; if (d > p || oldtag != tag) d = 0u; }
        }
        if (INTERIOR || (act && p >= bs)) mydist[p - bs] = (u16)d;
        ph += inc; if (ph >= period) ph -= period;
        act_c = act_n;
    }
}

template <int MLS>
__global__ void __launch_bounds__(32)
zb_cand_kernel(const u8* __restrict__ src, const u8* __restrict__ dictEnd, const ZbBlock* __restrict__ blocks, ZbParams prm, ZbStrides sd, u16* __restrict__ dist,
               const u8* __restrict__ imageIn, u8* __restrict__ imageOut)
{
    __shared__ __align__(16) u8 ring[CAND_RING];                  /* input staging */
    extern __shared__ __align__(16) u16 table[];                  /* 2^hashLog positions, followed by 2^hashLog tags */
    u32 const lane = threadIdx.x;
    ZbBlock bd = blocks[blockIdx.x];
    /* imageOut != NULL: this launch only walks the dictionary tail (positions whose 8 hashed bytes lie inside
     * the dictionary) and saves the table; imageIn != NULL: dictionary blocks start from that saved table and
     * walk only the last 7 dictionary positions (their hashes reach into the frame) and the block itself. */
    bool const buildImage = imageOut != nullptr;
    if (buildImage) bd.size = 0;
    if (!buildImage && bd.size < 7u) return;                      /* zstd_compress.c:3216 : block goes out raw */
    u16* const mydist = dist + (size_t)blockIdx.x * sd.dist;
    const u8* const base = buildImage ? dictEnd - bd.histLen : src + bd.srcOff - bd.histLen;   /* rel position 0 = oldest visible byte */
    u32 const bs = bd.histLen, be = bd.histLen + bd.size;
    bool const fromImage = imageIn != nullptr && (bd.flags & ZB_FLAG_DICT) && bd.histLen >= 8u;
    u32 const pmin = fromImage ? bd.histLen - 7u : 0u;            /* history positions below pmin are already in the image */
    u32 const hlog = prm.hashLog, period = prm.insPeriod;
    u8* const tags = reinterpret_cast<u8*>(table + ((size_t)1 << hlog));   /* 8 further hash bits per bucket */
    volatile u16* const vtable = table;     /* lanes communicate through the table inside a step: volatile accesses */
    volatile u8*  const vtags = tags;

    /* input is staged through shared memory in 16-byte aligned units: q = position relative to abase */
    u32 const o0 = (u32)((uintptr_t)base & 15u);
    const u8* const abase = base - o0;
    /* history of a dictionary block lives in the dictionary buffer: 16-byte units that are not entirely
     * frame bytes are assembled byte-wise (arbitrary mutual alignment), the rest goes through cp.async */
    bool const dictBlk = dictEnd != nullptr && (bd.flags & ZB_FLAG_DICT);
    const u8* const dlo = dictBlk ? dictEnd - bd.histLen : base;
    auto stage = [&](u32 q) {
        if (!dictBlk || q >= o0 + bd.histLen) { __pipeline_memcpy_async(ring + (q & (CAND_RING - 1u)), abase + q, 16); return; }
        u32 w[4] = { 0u, 0u, 0u, 0u };
#pragma unroll
        for (u32 k = 0; k < 16u; k++) {
            u32 const rel = q + k - o0;                       /* wraps for q + k < o0 : treated as out of range */
            u32 const byte = (q + k < o0 || rel >= bd.histLen + bd.size) ? 0u : (u32)(rel < bd.histLen ? dlo[rel] : base[rel]);
            w[k >> 2] |= byte << (8u * (k & 3u));
        }
        *reinterpret_cast<uint4*>(ring + (q & (CAND_RING - 1u))) = make_uint4(w[0], w[1], w[2], w[3]);
    };
    u32 const qEnd = o0 + be;                                     /* one past the last byte we may read */
    u32 const nChunks = (qEnd + CAND_CHUNK - 1u) / CAND_CHUNK;

    {   uint4* t4 = reinterpret_cast<uint4*>(table);
        u32 const n4 = (3u << hlog) / 16u;
        const uint4* im = reinterpret_cast<const uint4*>(imageIn);
        for (u32 i = lane; i < n4; i += 32) t4[i] = fromImage ? __ldg(im + i) : make_uint4(0, 0, 0, 0);
    }
    u32 const c0 = (o0 + pmin) / CAND_CHUNK;                      /* first chunk with a position to insert */
    /* prologue: chunks c0 .. c0+STAGES-2 in flight */
#pragma unroll
    for (u32 k = 0; k < CAND_STAGES - 1u; k++) {
        u32 const q = (c0 + k) * CAND_CHUNK + 16u * lane;
        if (c0 + k < nChunks && q < qEnd) stage(q);
        __pipeline_commit();
    }
    __syncwarp();

    u32 const nPos = be - 7u;                                     /* positions with 8 readable bytes inside the block */
    u32 const phase0 = (period - (o0 % period) + (prm.longPass ? bd.insPhaseLong : bd.insPhase)) % period;   /* pattern phase of q = 0 */
    u32 ph = (c0 * CAND_CHUNK + lane + phase0) % period;          /* pattern phase of this lane's q in the current step */
    u32 const inc = 32u % period;
    /* chunks that lie entirely inside the history only have to leave their inserted positions in the
     * table (nobody asks for their candidates): they are walked pair-wise, 16 pairs = 16*period
     * positions per step, without any look-up */
    u32 const nPrimeChunks = (o0 + bs) / CAND_CHUNK;
    for (u32 c = c0; c < nChunks; c++) {
        {   u32 const cn = c + CAND_STAGES - 1u;                  /* refill the slot consumed in the previous iteration */
            u32 const q = cn * CAND_CHUNK + 16u * lane;
            if (cn < nChunks && q < qEnd) stage(q);
            __pipeline_commit();
        }
        /* chunk c and (for the 8-byte reads that straddle its end) chunk c+1 must have landed */
        __pipeline_wait_prior(CAND_STAGES - 2u);
        __syncwarp();
        if (c < nPrimeChunks) {
            /* first pair start at or after the chunk start: q0 with (q0 + phase0) % period == 0 */
            u32 const cq = c * CAND_CHUNK;
            u32 const r = (cq + phase0) % period;
            u32 const first = cq + (r ? period - r : 0u);
            /* a pair that straddles the chunk start has its second element here: it precedes every
             * pair of this chunk, so it is written first */
            if (((cq + phase0) % period) == 1u && lane == 0u && cq >= o0 + pmin && cq - o0 < nPos) {
                u32 const q = cq, p = q - o0;
                u32 const w = (q & ~3u) & (CAND_RING - 1u);
                u32 const sh = (q & 3u) * 8u;
                u32 const a0 = *reinterpret_cast<const u32*>(ring + w);
                u32 const a1 = *reinterpret_cast<const u32*>(ring + ((w + 4u) & (CAND_RING - 1u)));
                u32 const a2 = *reinterpret_cast<const u32*>(ring + ((w + 8u) & (CAND_RING - 1u)));
                u64 const v = ((u64)__funnelshift_r(a1, a2, sh) << 32) | __funnelshift_r(a0, a1, sh);
                u32 const h24 = zb_hash(v, MLS, hlog + 8u);
                vtable[h24 >> 8] = (u16)p; vtags[h24 >> 8] = (u8)h24;
            }
            __syncwarp();
            for (u32 g0 = first; g0 < cq + CAND_CHUNK; g0 += 16u * period) {
                u32 const q = g0 + (lane >> 1) * period + (lane & 1u);
                bool const act = (q >= o0 + pmin) && (q < cq + CAND_CHUNK) && (q - o0 < nPos);
                u32 const p = q - o0;
                u32 const w = (q & ~3u) & (CAND_RING - 1u);
                u32 const sh = (q & 3u) * 8u;
                u32 const a0 = *reinterpret_cast<const u32*>(ring + w);
                u32 const a1 = *reinterpret_cast<const u32*>(ring + ((w + 4u) & (CAND_RING - 1u)));
                u32 const a2 = *reinterpret_cast<const u32*>(ring + ((w + 8u) & (CAND_RING - 1u)));
                u64 const v = ((u64)__funnelshift_r(a1, a2, sh) << 32) | __funnelshift_r(a0, a1, sh);
                u32 const h24 = zb_hash(v, MLS, hlog + 8u);
                u32 const h = h24 >> 8;
                if (act) vtable[h] = (u16)p;
                __syncwarp();
                /* the latest position must own the bucket: lanes that lost to an earlier one write again */
                while (true) {
                    u32 const diff = act ? ((p - (u32)vtable[h]) & 0xFFFFu) : 0u;      /* > 0 : an earlier position of this step is stored */
                    bool const again = diff != 0u && diff <= 16u * period;
                    if (!__any_sync(ZB_FULL, again)) break;
                    __syncwarp();
                    if (again) vtable[h] = (u16)p;
                    __syncwarp();
                }
                if (act && vtable[h] == (u16)p) vtags[h] = (u8)h24;       /* the bucket's owner sets its tag */
                __syncwarp();
            }
            __syncwarp();
            ph += (CAND_CHUNK % period); if (ph >= period) ph -= period;     /* keep the step phase in sync (CAND_CHUNK/32 steps skipped) */
            continue;
        }
        /* phase A: the 16 steps' hashes are independent of the table: compute them back to back */
        u32 hh[CAND_CHUNK / 32u];
#pragma unroll
        for (u32 j = 0; j < CAND_CHUNK / 32u; j++) {
            u32 const q = c * CAND_CHUNK + 32u * j + lane;
            u32 const w = (q & ~3u) & (CAND_RING - 1u);
            u32 const sh = (q & 3u) * 8u;
            u32 const a0 = *reinterpret_cast<const u32*>(ring + w);
            u32 const a1 = *reinterpret_cast<const u32*>(ring + ((w + 4u) & (CAND_RING - 1u)));
            u32 const a2 = *reinterpret_cast<const u32*>(ring + ((w + 8u) & (CAND_RING - 1u)));
            u64 const v = ((u64)__funnelshift_r(a1, a2, sh) << 32) | __funnelshift_r(a0, a1, sh);
            hh[j] = zb_hash(v, MLS, hlog + 8u);                 /* bucket << 8 | tag */
        }
        /* phase B: the table walk proper, one step after the other */
        {   u32 const q0 = c * CAND_CHUNK;
            bool const interior = (q0 >= o0 + bs) && (q0 + CAND_CHUNK <= o0 + nPos);
            if (interior) zb_cand_walk16<true>(table, tags, hh, q0, o0, pmin, nPos, bs, lane, ph, inc, period, mydist);
            else          zb_cand_walk16<false>(table, tags, hh, q0, o0, pmin, nPos, bs, lane, ph, inc, period, mydist);
        }
    }
    if (buildImage) {                                             /* save the primed table (positions + tags) */
        __syncwarp();
        const uint4* t4 = reinterpret_cast<const uint4*>(table);
        uint4* im = reinterpret_cast<uint4*>(imageOut);
        for (u32 i = lane; i < (3u << hlog) / 16u; i += 32) im[i] = t4[i];
        return;
    }
    for (u32 p = (nPos > bs ? nPos : bs) + lane; p < be; p += 32) mydist[p - bs] = 0;
}

/* ------------------------------------------------------------------------------------------------
 * K1b — greedy selection + match extension + sequence/literal emission.  One warp per block, no
 * shared memory (occupancy is register-bound, so the L2 latency of the candidate checks is hidden
 * by other warps).  Per step 32 probe positions: pairs (p, p+1) spaced by `step` (zstd_fast.c:225-229,
 * step acceleration :234,:342-347).  Hit priority per lane: repcode-2 (lane 0, directly after a match,
 * :410-420), repcode-1 (:281-297), table candidate with 4-byte check (:102-141).  Lowest lane wins.
 * ---------------------------------------------------------------------------------------------- */
#ifndef PARSE_WARPS
#define PARSE_WARPS 8            /* = ZB_PARSE_SEGS: the eight segments of a full block share a CTA */
#endif
#ifndef PARSE_MIN_CTAS
#define PARSE_MIN_CTAS 6           /* 48 warps per SM at 40 registers.  8 (= 32 registers, 64 warps) is 3.7 % faster when the parse
                                    * runs alone, but fills every warp slot: the candidate walk of the next wave no longer fits beside
                                    * it and a whole device-resident call gets 4 % slower (profiles/r1_history.md) */
#endif
template <bool DICT>
__global__ void __launch_bounds__(32 * PARSE_WARPS, DICT ? (40 / PARSE_WARPS) : PARSE_MIN_CTAS)
zb_parse_kernel(const u8* __restrict__ src, const u8* __restrict__ dictEnd, const ZbBlock* __restrict__ blocks, u32 nbBlocks, ZbParams prm, ZbStrides sd,
                const u16* __restrict__ dist, u64* __restrict__ seqs, ZbBlockMeta* __restrict__ meta, ZbSegMeta* __restrict__ segmeta)
{
    u32 const lane = threadIdx.x & 31u;
    u32 const g = blockIdx.x * PARSE_WARPS + (threadIdx.x >> 5);  /* one warp per segment: the warps of a CTA share a block's history in L1/L2 */
    u32 const segs = (sd.dist + ZB_PARSE_SEG - 1u) / ZB_PARSE_SEG;   /* segments of the call's largest block (1 for calls of short frames) */
    u32 const b = g / segs, k = g % segs;
    if (b >= nbBlocks) return;
    ZbBlock const bd = blocks[b];
    u64* const myseq = seqs + (size_t)b * sd.seq + (size_t)k * (ZB_PARSE_SEG / 4u);
    const u16* const mydist = dist + (size_t)b * sd.dist;
    const u8* const base = src + bd.srcOff - bd.histLen;          /* base + rel addresses the frame's own bytes */
    u32 const bs = bd.histLen, blockEnd = bd.histLen + bd.size;
    ZbSeg sg; sg.hi = base; sg.lo = base; sg.split = 0;
    if (DICT && (bd.flags & ZB_FLAG_DICT)) { sg.lo = dictEnd - bd.histLen; sg.split = bd.histLen; }   /* history = dictionary tail */

    if (bd.size < 7u) {                                        /* zstd_compress.c:3216 */
        if (lane == 0 && k == 0) {
            ZbBlockMeta m; m.nbSeq = 0; m.litSize = bd.size; m.litSecSize = 0; m.bodySize = bd.size;
            m.type = ZB_BT_RAW; m.forceRaw = 1; m.rleByte = 0; m.pad = 0;
            meta[b] = m;
        }
        return;
    }
    u32 const ss = bs + k * ZB_PARSE_SEG;                      /* this warp's segment [ss, be) */
    if (ss >= blockEnd) {
        if (lane == 0) { ZbSegMeta z; z.nbSeq = 0; z.litSize = 0; z.trail = 0; z.pad = 0; segmeta[(size_t)b * segs + k] = z; }
        return;
    }
    u32 const be = min(ss + ZB_PARSE_SEG, blockEnd);

    u32 ip = ss, anchor = ss, searchStart = ss;
    u32 rep1 = 0, rep2 = 0, nbSeq = 0;
    if (DICT && (bd.flags & ZB_FLAG_DICT) && k == 0u) { rep1 = prm.startRep[0]; rep2 = prm.startRep[1]; }   /* a zstd-format dictionary's repcodes, zstd_compress.c:5054-5056 */

    while (ip + 8u <= be) {
        u32 const step = prm.stepSize + ((ip - searchStart) >> 7);           /* kSearchStrength = 8 */
        u32 const p = ip + (lane >> 1) * step + (lane & 1u);
        bool const act = (p + 8u <= be);
        u32 const pp = act ? p : ip;                         /* a position every lane may load from (ip + 8 <= be) */
        u32 const d = act ? (u32)mydist[pp - bs] : 0u;
        bool const v3 = (lane == 0u) && (ip == anchor) && (rep2 != 0u);
        bool const v2 = act && rep1 != 0u && p >= rep1;
        bool const v1 = act && d != 0u;
        /* dist[] only holds candidates K1a has already verified (4 equal bytes), so a step needs no random
         * load: the current window and the repcode windows are contiguous across lanes */
        u32 pre, cur, pre2, cur2;
        zb_seg_pre_cur<DICT>(sg, pp, &pre, &cur);
        zb_seg_pre_cur<DICT>(sg, v2 ? pp - rep1 : pp, &pre2, &cur2);
        u32 cur3 = ~cur;
        if (ip == anchor && rep2 != 0u) cur3 = zb_seg_ld32<DICT>(sg, v3 ? pp - rep2 : pp);     /* warp-uniform condition */
        u32 const hit = (v3 && cur3 == cur) ? 3u : ((v2 && cur2 == cur) ? 2u : (v1 ? 1u : 0u));
        u32 tent = __ballot_sync(ZB_FULL, hit != 0u);
        /* backward catch-up (zstd_fast.c:387-391) of a repcode-1 hit: first 4 bytes in-lane from the windows */
        u32 myback = 0, mymore = 0;
        if (hit == 2u) {
            u32 const x = pre ^ pre2;
            u32 const bm = x ? ((u32)__clz((int)x) >> 3) : 4u;
            u32 lim = p - anchor; lim = lim < 4u ? lim : 4u;
            u32 const src0 = p - rep1; lim = lim < src0 ? lim : src0;
            myback = bm < lim ? bm : lim;
            mymore = (bm == 4u && lim == 4u) ? 1u : 0u;
        }
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
        u32 const litLen = ms - anchor;
        u32 offBase;
        if (wtype == 3u) { offBase = 1u; u32 const t = rep2; rep2 = rep1; rep1 = t; }   /* litLength 0: code 1 = repcode 2 */
        else if (wtype == 2u && litLen > 0u) offBase = 1u;                                 /* REPCODE1_TO_OFFBASE */
        else { offBase = offset + 3u; rep2 = rep1; rep1 = offset; }
        if (lane == 0) myseq[nbSeq] = zb_pack_seq(offBase, litLen, mlen);
        nbSeq++;                                               /* the literal bytes are gathered by the merge kernel */
        ip = ms + mlen; anchor = ip; searchStart = ip;
    }

    /* trailing literals (zstd_compress.c:3365-3366): the block's last literals, or the head of the next segment's first sequence */
    u32 const lastLits = be - anchor;
    if (lane == 0) { ZbSegMeta z; z.nbSeq = nbSeq; z.litSize = 0; z.trail = lastLits; z.pad = 0; segmeta[(size_t)b * segs + k] = z; }
}

/* K1c — joins a block's segments and materialises its literals.
 * 1. sequences move down to be contiguous (in place, ascending, every chunk read completely before it is written: the
 *    destination never lies above the source); the first sequence of a segment takes over the literals the segments
 *    before it left behind their last match;
 * 2. a scan over (litLength, litLength + matchLength) gives every sequence the block position its literals start at
 *    and their offset in the literal buffer; warps copy them straight from the input (the parse kernels emit no
 *    literal bytes at all), then the block's last literals;
 * 3. the block's meta record is written. */
#define MERGE_THREADS 256
#ifndef MERGE_TILE
#define MERGE_TILE 1024u                       /* sequences scanned and gathered per round */
#endif
#define MERGE_PER (MERGE_TILE / MERGE_THREADS)  /* consecutive sequences of a tile owned by one thread */
__global__ void __launch_bounds__(MERGE_THREADS)
zb_merge_segments_kernel(const u8* __restrict__ src, const ZbBlock* __restrict__ blocks, ZbStrides sd, const ZbSegMeta* __restrict__ segmeta,
                         u64* __restrict__ seqs, u8* __restrict__ lits, ZbBlockMeta* __restrict__ meta)
{
    __shared__ u32 sPos[MERGE_TILE], sLit[MERGE_TILE], sLen[MERGE_TILE];
    __shared__ u32 wsumL[MERGE_THREADS / 32], wsumA[MERGE_THREADS / 32];
    __shared__ u32 baseL, baseA;
    u32 const b = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    ZbBlock const bd = blocks[b];
    if (bd.size < 7u) return;                                    /* raw block: meta written by the parse kernel */
    u64* const myseq = seqs + (size_t)b * sd.seq;
    u8*  const mylit = lits + (size_t)b * sd.lit;
    const u8* const in = src + bd.srcOff;                        /* literals always lie inside the block itself */
    u32 const segs = (sd.dist + ZB_PARSE_SEG - 1u) / ZB_PARSE_SEG;
    ZbSegMeta sm[ZB_PARSE_SEGS];
#pragma unroll
    for (u32 k = 0; k < ZB_PARSE_SEGS; k++) if (k < segs) sm[k] = segmeta[(size_t)b * segs + k];
    /* ---- 1. sequences ---- */
    u32 seqOff = sm[0].nbSeq;
    u32 carry = sm[0].trail;                                     /* literals waiting for the next sequence */
#pragma unroll 1
    for (u32 k = 1; k < segs; k++) {
        u32 const ns = sm[k].nbSeq;
        u64* const sfrom = myseq + (size_t)k * (ZB_PARSE_SEG / 4u);
        for (u32 c0 = 0; c0 < ns; c0 += MERGE_THREADS) {
            u32 const i = c0 + tid;
            u64 v = 0;
            if (i < ns) { v = sfrom[i]; if (i == 0u) v += (u64)carry << 24; }        /* litLength field, zb_pack_seq */
            __syncthreads();
            if (i < ns) myseq[seqOff + i] = v;
            __syncthreads();
        }
        carry = ns ? sm[k].trail : carry + sm[k].trail;
        seqOff += ns;
    }
    u32 const nbSeq = seqOff;
    if (tid == 0) { baseL = 0; baseA = 0; }
    __syncthreads();
    /* ---- 2. literals ---- */
    for (u32 t0 = 0; t0 < nbSeq; t0 += MERGE_TILE) {
        u32 const n = min(MERGE_TILE, nbSeq - t0);
        /* every thread owns MERGE_PER consecutive sequences of the tile */
        u32 ll[MERGE_PER], adv[MERGE_PER], myL = 0, myA = 0;
#pragma unroll
        for (u32 j = 0; j < MERGE_PER; j++) {
            u32 const i = tid * MERGE_PER + j;
            u64 const q = (i < n) ? myseq[t0 + i] : 0ull;
            ll[j] = (u32)((q >> 24) & 0x3FFFFu);
            adv[j] = ll[j] + (u32)(q >> 42);
            myL += ll[j]; myA += adv[j];
        }
        u32 inL = myL, inA = myA;                                /* inclusive scan over the warp, then over the warps */
#pragma unroll
        for (u32 o = 1; o < 32u; o <<= 1) {
            u32 const a = __shfl_up_sync(ZB_FULL, inL, o), c = __shfl_up_sync(ZB_FULL, inA, o);
            if (lane >= o) { inL += a; inA += c; }
        }
        if (lane == 31u) { wsumL[warp] = inL; wsumA[warp] = inA; }
        __syncthreads();
        u32 offL = baseL + inL - myL, offA = baseA + inA - myA;
        for (u32 w = 0; w < warp; w++) { offL += wsumL[w]; offA += wsumA[w]; }
#pragma unroll
        for (u32 j = 0; j < MERGE_PER; j++) {
            u32 const i = tid * MERGE_PER + j;
            if (i < n) { sPos[i] = offA; sLit[i] = offL; sLen[i] = ll[j]; }
            offL += ll[j]; offA += adv[j];
        }
        __syncthreads();
        if (tid == MERGE_THREADS - 1u) { baseL = offL; baseA = offA; }     /* totals up to the end of this tile */
        for (u32 i = warp; i < n; i += MERGE_THREADS / 32u) {
            u32 const len = sLen[i];
            const u8* const from = in + sPos[i];
            u8* const to = mylit + sLit[i];
            for (u32 x = lane; x < len; x += 32u) to[x] = from[x];
        }
        __syncthreads();
    }
    u32 const litSeq = baseL, consumed = baseA;                   /* literals in sequences, bytes covered by sequences */
    u32 const lastLits = bd.size - consumed;
    for (u32 x = tid; x < lastLits; x += MERGE_THREADS) mylit[litSeq + x] = in[consumed + x];
    /* ---- 3. meta ---- */
    if (tid == 0) {
        ZbBlockMeta m; m.nbSeq = nbSeq; m.litSize = litSeq + lastLits; m.litSecSize = 0; m.bodySize = 0;
        m.type = ZB_BT_COMPRESSED; m.forceRaw = 0; m.rleByte = 0; m.pad = 0;
        meta[b] = m;
    }
}

/* K1c for calls made of short frames (one segment per block, so the sequences are already in place): one warp per
 * block, 8 blocks per CTA; every lane gathers the literals of its own sequences. */
__global__ void __launch_bounds__(MERGE_THREADS)
zb_merge_small_kernel(const u8* __restrict__ src, const ZbBlock* __restrict__ blocks, u32 nbBlocks, ZbStrides sd, const ZbSegMeta* __restrict__ segmeta,
                      const u64* __restrict__ seqs, u8* __restrict__ lits, ZbBlockMeta* __restrict__ meta)
{
    u32 const lane = threadIdx.x & 31u;
    u32 const b = blockIdx.x * (MERGE_THREADS / 32u) + (threadIdx.x >> 5);
    if (b >= nbBlocks) return;
    ZbBlock const bd = blocks[b];
    if (bd.size < 7u) return;
    const u64* const myseq = seqs + (size_t)b * sd.seq;
    u8* const mylit = lits + (size_t)b * sd.lit;
    const u8* const in = src + bd.srcOff;
    u32 const nbSeq = segmeta[b].nbSeq;
    u32 posL = 0, posA = 0;
    for (u32 t0 = 0; t0 < nbSeq; t0 += 32u) {
        u32 const i = t0 + lane;
        u64 const q = (i < nbSeq) ? myseq[i] : 0ull;
        u32 const ll = (u32)((q >> 24) & 0x3FFFFu), adv = ll + (u32)(q >> 42);
        u32 inL = ll, inA = adv;
#pragma unroll
        for (u32 o = 1; o < 32u; o <<= 1) {
            u32 const x = __shfl_up_sync(ZB_FULL, inL, o), y = __shfl_up_sync(ZB_FULL, inA, o);
            if (lane >= o) { inL += x; inA += y; }
        }
        const u8* const from = in + posA + inA - adv;
        u8* const to = mylit + posL + inL - ll;
        for (u32 x = 0; x < ll; x++) to[x] = from[x];
        posL += __shfl_sync(ZB_FULL, inL, 31); posA += __shfl_sync(ZB_FULL, inA, 31);
    }
    u32 const lastLits = bd.size - posA;
    for (u32 x = lane; x < lastLits; x += 32u) mylit[posL + x] = in[posA + x];
    if (lane == 0) {
        ZbBlockMeta m; m.nbSeq = nbSeq; m.litSize = posL + lastLits; m.litSecSize = 0; m.bodySize = 0;
        m.type = ZB_BT_COMPRESSED; m.forceRaw = 0; m.rleByte = 0; m.pad = 0;
        meta[b] = m;
    }
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
__global__ void __launch_bounds__(32 * PARSE_WARPS)
zb_parse_dfast_kernel(const u8* __restrict__ src, const ZbBlock* __restrict__ blocks, u32 nbBlocks, ZbParams prm, ZbStrides sd,
                      const u16* __restrict__ distLong, const u16* __restrict__ distShort,
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
    const u16* const dSp = distShort + (size_t)b * sd.dist;
    const u8* const base = src + bd.srcOff - bd.histLen;
    u32 const bs = bd.histLen, blockEnd = bd.histLen + bd.size;
    ZbSeg sg; sg.hi = base; sg.lo = base; sg.split = 0;

    if (bd.size < 7u) {                                        /* zstd_compress.c:3216 */
        if (lane == 0 && k == 0) {
            ZbBlockMeta m; m.nbSeq = 0; m.litSize = bd.size; m.litSecSize = 0; m.bodySize = bd.size;
            m.type = ZB_BT_RAW; m.forceRaw = 1; m.rleByte = 0; m.pad = 0;
            meta[b] = m;
        }
        return;
    }
    u32 const ss = bs + k * ZB_PARSE_SEG;                      /* this warp's segment [ss, be) */
    if (ss >= blockEnd) {
        if (lane == 0) { ZbSegMeta z; z.nbSeq = 0; z.litSize = 0; z.trail = 0; z.pad = 0; segmeta[(size_t)b * segs + k] = z; }
        return;
    }
    u32 const be = min(ss + ZB_PARSE_SEG, blockEnd);
    u32 ip = ss, anchor = ss, searchStart = ss;
    u32 rep1 = 0, rep2 = 0, nbSeq = 0;

    while (ip + 9u <= be) {                                   /* a lane reads 8 bytes at p and at p+1 */
        u32 const step = 1u + ((ip - searchStart) >> 8);                     /* kStepIncr = 1 << kSearchStrength */
        u32 const p = ip + lane * step;
        bool const act = (p + 9u <= be);
        u32 const pp = act ? p : ip;
        u32 const dL = act ? (u32)dLp[pp - bs] : 0u;
        u32 const dS = act ? (u32)dSp[pp - bs] : 0u;
        u32 const dL1 = act ? (u32)dLp[pp + 1u - bs] : 0u;
        u64 const w = zb_ld64w3(base + pp);                                  /* bytes p .. p+7 */
        u32 const cur = (u32)w, cur1 = (u32)(w >> 8);
        bool const v2 = act && rep1 != 0u && (p + 1u >= rep1);
        u32 const r2 = zb_ld32w2(base + (v2 ? pp + 1u - rep1 : pp));
        u32 r3 = ~cur;
        bool const v3 = (lane == 0u) && (ip == anchor) && (rep2 != 0u);
        if (ip == anchor && rep2 != 0u) r3 = zb_ld32w2(base + (v3 ? pp - rep2 : pp));
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
            if (wtype == 3u) { ms = probe; offset = rep2; mlen = 4u + zb_count_fwd<false>(sg, probe + 4u, rep2, be, lane); found = true; break; }
            if (wtype == 2u) { ms = probe + 1u; offset = rep1; mlen = 4u + zb_count_fwd<false>(sg, probe + 5u, rep1, be, lane); found = true; break; }
            if (wtype == 1u) {
                u32 const f0 = zb_count_fwd<false>(sg, probe, wL, be, lane);
                if (f0 >= 8u) {
                    u32 const back = zb_back_coop<false>(sg, probe, wL, anchor, lane);
                    ms = probe - back; offset = wL; mlen = back + f0; found = true; break;
                }
                /* tag collision on the long table: the lane may still have a short candidate */
                if (lane == (u32)winner) hit = dS ? 4u : 0u;
                if (wS == 0u) { tent &= ~(1u << winner); }
                continue;
            }
            /* short candidate */
            {   u32 const f0 = zb_count_fwd<false>(sg, probe, wS, be, lane);
                if (f0 < 4u) { tent &= ~(1u << winner); if (lane == (u32)winner) hit = 0u; continue; }
                u32 mp = probe, mo = wS, ml = f0;
                if (wL1) {
                    u32 const f1 = zb_count_fwd<false>(sg, probe + 1u, wL1, be, lane);
                    if (f1 >= 8u && f1 > ml) { mp = probe + 1u; mo = wL1; ml = f1; }
                }
                u32 const back = zb_back_coop<false>(sg, mp, mo, anchor, lane);
                ms = mp - back; offset = mo; mlen = back + ml; wtype = 1u; found = true; break;
            }
        }
        if (!found) { ip += 32u * step; continue; }
        u32 const litLen = ms - anchor;
        u32 offBase;
        if (wtype == 3u) { offBase = 1u; u32 const t = rep2; rep2 = rep1; rep1 = t; }
        else if (wtype == 2u && litLen > 0u) offBase = 1u;
        else { offBase = offset + 3u; rep2 = rep1; rep1 = offset; }
        if (lane == 0) myseq[nbSeq] = zb_pack_seq(offBase, litLen, mlen);
        nbSeq++;                                               /* the literal bytes are gathered by the merge kernel */
        ip = ms + mlen; anchor = ip; searchStart = ip;
    }
    u32 const lastLits = be - anchor;
    if (lane == 0) { ZbSegMeta z; z.nbSeq = nbSeq; z.litSize = 0; z.trail = lastLits; z.pad = 0; segmeta[(size_t)b * segs + k] = z; }
}

static void zb_launch_cand(const u8* d_src, const u8* d_dictEnd, const ZbBlock* d_blocks, u32 nbBlocks, const ZbParams& prm, const ZbStrides& sd, u16* d_dist,
                           const u8* d_imageIn, u8* d_imageOut, cudaStream_t stream)
{
    size_t const smem = (size_t)3 << prm.hashLog;       /* u16 positions + u8 tags */
    static bool optinDev[64];                                     /* the attribute is per device */
    int dev = 0; cudaGetDevice(&dev);
    bool& optin = optinDev[dev & 63];
    if (!optin) {             /* hashLog 14: 48 KiB of table + the 2 KiB static ring exceeds the default 48 KiB limit */
        cudaFuncSetAttribute(zb_cand_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(zb_cand_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(zb_cand_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(zb_cand_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(zb_cand_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        optin = true;
    }
    switch (prm.mls) {
    case 4:  zb_cand_kernel<4><<<nbBlocks, 32, smem, stream>>>(d_src, d_dictEnd, d_blocks, prm, sd, d_dist, d_imageIn, d_imageOut); break;
    case 5:  zb_cand_kernel<5><<<nbBlocks, 32, smem, stream>>>(d_src, d_dictEnd, d_blocks, prm, sd, d_dist, d_imageIn, d_imageOut); break;
    case 6:  zb_cand_kernel<6><<<nbBlocks, 32, smem, stream>>>(d_src, d_dictEnd, d_blocks, prm, sd, d_dist, d_imageIn, d_imageOut); break;
    case 7:  zb_cand_kernel<7><<<nbBlocks, 32, smem, stream>>>(d_src, d_dictEnd, d_blocks, prm, sd, d_dist, d_imageIn, d_imageOut); break;
    default: zb_cand_kernel<8><<<nbBlocks, 32, smem, stream>>>(d_src, d_dictEnd, d_blocks, prm, sd, d_dist, d_imageIn, d_imageOut); break;
    }
}

/* one-warp launch that primes a table from the dictionary tail and stores it (positions + tags) in d_image */
extern "C" cudaError_t zb_launch_dict_image(const u8* d_dictEnd, const ZbBlock* d_dictBlock, const ZbParams* prm, u8* d_image, cudaStream_t stream)
{
    ZbStrides sd; sd.dist = ZB_BLOCK_MAX; sd.seq = ZB_SEQ_STRIDE; sd.lit = ZB_LIT_STRIDE; sd.body = ZB_BODY_STRIDE; sd.state = ZB_STATE_STRIDE;   /* unused: no block is walked */
    zb_launch_cand(nullptr, d_dictEnd, d_dictBlock, 1, *prm, sd, nullptr, nullptr, d_image, stream);
    return cudaGetLastError();
}

extern "C" cudaError_t zb_launch_match(const u8* d_src, const u8* d_dictEnd, const u8* d_image, const ZbBlock* d_blocks, u32 nbBlocks, const ZbParams* prm, const ZbStrides* sdp,
                                       u16* d_dist, u16* d_dist2, u64* d_seqs, u8* d_lits, ZbBlockMeta* d_meta, ZbSegMeta* d_segmeta, cudaEvent_t evMid, cudaStream_t stream)
{
    if (nbBlocks == 0) return cudaSuccess;
    ZbStrides const sd = *sdp;
    if (prm->strategy == 2) {
        /* doubleFast: one candidate walk per table (both walks see the same per-block insertion phase) */
        ZbParams pl = *prm; pl.mls = 8; pl.hashLog = prm->longHashLog; pl.insPeriod = prm->insPeriodLong; pl.longPass = 1;
        zb_launch_cand(d_src, nullptr, d_blocks, nbBlocks, pl, sd, d_dist, nullptr, nullptr, stream);
        zb_launch_cand(d_src, nullptr, d_blocks, nbBlocks, *prm, sd, d_dist2, nullptr, nullptr, stream);
        if (evMid) cudaEventRecord(evMid, stream);
        u32 const segs = (sd.dist + ZB_PARSE_SEG - 1u) / ZB_PARSE_SEG;
        u32 const sgrid = (u32)(((u64)nbBlocks * segs + PARSE_WARPS - 1) / PARSE_WARPS);
        zb_parse_dfast_kernel<<<sgrid, 32 * PARSE_WARPS, 0, stream>>>(d_src, d_blocks, nbBlocks, *prm, sd, d_dist, d_dist2, d_seqs, d_meta, d_segmeta);
        if (segs == 1u && sd.dist <= 8192u)
            zb_merge_small_kernel<<<(nbBlocks + MERGE_THREADS / 32u - 1u) / (MERGE_THREADS / 32u), MERGE_THREADS, 0, stream>>>(d_src, d_blocks, nbBlocks, sd, d_segmeta, d_seqs, d_lits, d_meta);
        else
            zb_merge_segments_kernel<<<nbBlocks, MERGE_THREADS, 0, stream>>>(d_src, d_blocks, sd, d_segmeta, d_seqs, d_lits, d_meta);
    } else {
        zb_launch_cand(d_src, d_dictEnd, d_blocks, nbBlocks, *prm, sd, d_dist, d_image, nullptr, stream);
        if (evMid) cudaEventRecord(evMid, stream);
        u32 const segs = (sd.dist + ZB_PARSE_SEG - 1u) / ZB_PARSE_SEG;
        u32 const sgrid = (u32)(((u64)nbBlocks * segs + PARSE_WARPS - 1) / PARSE_WARPS);                   /* one warp per segment */
        if (d_dictEnd) zb_parse_kernel<true><<<sgrid, 32 * PARSE_WARPS, 0, stream>>>(d_src, d_dictEnd, d_blocks, nbBlocks, *prm, sd, d_dist, d_seqs, d_meta, d_segmeta);
        else           zb_parse_kernel<false><<<sgrid, 32 * PARSE_WARPS, 0, stream>>>(d_src, nullptr, d_blocks, nbBlocks, *prm, sd, d_dist, d_seqs, d_meta, d_segmeta);
        if (segs == 1u && sd.dist <= 8192u)
            zb_merge_small_kernel<<<(nbBlocks + MERGE_THREADS / 32u - 1u) / (MERGE_THREADS / 32u), MERGE_THREADS, 0, stream>>>(d_src, d_blocks, nbBlocks, sd, d_segmeta, d_seqs, d_lits, d_meta);
        else
            zb_merge_segments_kernel<<<nbBlocks, MERGE_THREADS, 0, stream>>>(d_src, d_blocks, sd, d_segmeta, d_seqs, d_lits, d_meta);
    }
    return cudaGetLastError();
}
