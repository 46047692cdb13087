/* zb_match.cu — K1: warp-per-block greedy LZ77 match-finder ("fast" strategy).
 *
 * Replaces the CPU loop ZSTD_compressBlock_fast_noDict_generic
 * (/root/reference/lib/compress/zstd_fast.c:192-423) with a data-parallel formulation:
 *   - one warp owns one <=128 KiB block; its hash table (2^hashLog u16 entries, positions modulo
 *     64 KiB relative to the start of the visible history) lives in shared memory, so 7 blocks
 *     (hashLog 14) are resident per SM;
 *   - the table is primed from the history bytes in front of the block (zstd_fast.c:53-85 does this
 *     for dictionaries, zstdmt_compress.c:726-731 for job overlaps);
 *   - each step probes 32 positions at once — position pairs (p, p+1) spaced by `step` as in
 *     zstd_fast.c:225-229 — against the table state at the start of the step; repcode-1 is probed
 *     at every lane; the lowest matching lane wins (warp ballot);
 *   - backward catch-up (:387-391) and forward extension (ZSTD_count, zstd_compress_internal.h:771)
 *     are warp-cooperative: 32 x 8 bytes per round, first differing lane found by ballot;
 *   - sparse post-match inserts (:403-408) and the immediate repcode-2 loop (:410-420) follow.
 * Table writes are made deterministic with __match_any_sync (highest lane wins a bucket).
 * The bit-exact CPU model of this kernel is oracle/zb_match.c (tests only).
 */
#include "zb_device.cuh"
#include "zb_kernels.h"

/* matched bytes starting at rel positions (a, a - offset), never reading at or past `be` */
__device__ __forceinline__ u32 zb_count_fwd(const u8* __restrict__ base, u32 a, u32 offset, u32 be, u32 lane)
{
    u32 fwd = 0;
    while (true) {
        u32 const pa = a + fwd + 8u * lane;
        u32 m;
        if (pa + 8u <= be) {
            u64 const x = zb_ld64u(base + pa) ^ zb_ld64u(base + pa - offset);
            m = x ? (u32)((__ffsll((long long)x) - 1) >> 3) : 8u;
        } else {
            m = 0;
            while (pa + m < be && base[pa + m] == base[pa + m - offset]) m++;
        }
        u32 const inc = __ballot_sync(ZB_FULL, m != 8u);
        if (inc == 0) { fwd += 256u; continue; }
        int const f = __ffs((int)inc) - 1;
        fwd += 8u * (u32)f + __shfl_sync(ZB_FULL, m, f);
        return fwd;
    }
}

__device__ __forceinline__ u64 zb_pack_seq(u32 offBase, u32 litLen, u32 matchLen)
{
    return (u64)offBase | ((u64)litLen << 24) | ((u64)matchLen << 42);
}

__global__ void __launch_bounds__(32)
zb_match_fast_kernel(const u8* __restrict__ src, const ZbBlock* __restrict__ blocks, ZbParams prm,
                     u64* __restrict__ seqs, u8* __restrict__ lits, ZbBlockMeta* __restrict__ meta)
{
    extern __shared__ u16 table[];
    u32 const lane = threadIdx.x;
    ZbBlock const bd = blocks[blockIdx.x];
    u64* const myseq = seqs + (size_t)blockIdx.x * ZB_SEQ_STRIDE;
    u8*  const mylit = lits + (size_t)blockIdx.x * ZB_LIT_STRIDE;
    const u8* const base = src + bd.srcOff - bd.histLen;      /* rel position 0 = oldest visible byte */
    u32 const bs = bd.histLen, be = bd.histLen + bd.size;
    u32 const mls = prm.mls, hlog = prm.hashLog;

    if (bd.size < 7u) {                                        /* zstd_compress.c:3216 */
        if (lane == 0) {
            ZbBlockMeta m; m.nbSeq = 0; m.litSize = bd.size; m.litSecSize = 0; m.bodySize = bd.size;
            m.type = ZB_BT_RAW; m.forceRaw = 1; m.rleByte = 0; m.pad = 0;
            meta[blockIdx.x] = m;
        }
        return;
    }

    /* ---- clear, then prime the table from the visible history ---- */
    {   uint4* t4 = reinterpret_cast<uint4*>(table);
        u32 const n4 = (2u << hlog) / 16u;
        for (u32 i = lane; i < n4; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    for (u32 p0 = 0; p0 < bs; p0 += 32) {
        u32 const p = p0 + lane;
        bool const act = (p < bs) && (p + 8u <= be);
        u32 const h = act ? zb_hash(zb_ld64u(base + p), mls, hlog) : (0xFFFF0000u | lane);
        u32 const mm = __match_any_sync(ZB_FULL, h);
        if (act && (31u - (u32)__clz((int)mm)) == lane) table[h] = (u16)p;
        __syncwarp();
    }

    u32 ip = bs, anchor = bs, searchStart = bs;
    u32 rep1 = 0, rep2 = 0, nbSeq = 0, litPos = 0;

    while (ip + 8u <= be) {
        u32 const step = prm.stepSize + ((ip - searchStart) >> 7);           /* kSearchStrength = 8 */
        u32 const p = ip + (lane >> 1) * step + (lane & 1u);
        bool const act = (p + 8u <= be);
        u64 const v = act ? zb_ld64u(base + p) : 0ull;
        u32 const cur = (u32)v;
        u32 const h = act ? zb_hash(v, mls, hlog) : (0xFFFF0000u | lane);
        u32 const stored = act ? table[h] : 0u;
        u32 const dist = (p - stored) & 0xFFFFu;
        bool const cvalid = act && dist != 0u && dist <= p;
        bool const rvalid = act && rep1 != 0u && p >= rep1;
        u32 const c4 = cvalid ? zb_ld32u(base + p - dist) : ~cur;
        u32 const r4 = rvalid ? zb_ld32u(base + p - rep1) : ~cur;
        u32 const hit = (r4 == cur) ? 2u : ((c4 == cur) ? 1u : 0u);
        u32 const bal = __ballot_sync(ZB_FULL, hit != 0u);
        int const winner = bal ? (__ffs((int)bal) - 1) : -1;

        /* inserts: lanes up to the winner; highest lane wins a bucket */
        {   bool const ins = act && (winner < 0 || (int)lane <= winner);
            u32 const insmask = __ballot_sync(ZB_FULL, ins);
            u32 const mm = __match_any_sync(ZB_FULL, h) & insmask;
            if (ins && (31u - (u32)__clz((int)mm)) == lane) table[h] = (u16)p;
        }
        __syncwarp();
        if (winner < 0) { ip += 16u * step; continue; }

        u32 const probe = __shfl_sync(ZB_FULL, p, winner);
        bool const isRep = __shfl_sync(ZB_FULL, hit, winner) == 2u;
        u32 const offset = isRep ? rep1 : __shfl_sync(ZB_FULL, dist, winner);

        /* backward catch-up */
        u32 back = 0;
        while (true) {
            u32 const k = back + lane + 1u;                    /* compare bytes probe-k and probe-offset-k */
            bool const ok = (probe >= anchor + k) && (probe >= offset + k)
                         && (base[probe - k] == base[probe - offset - k]);
            u32 const okb = __ballot_sync(ZB_FULL, ok);
            u32 const cnt = (okb == ZB_FULL) ? 32u : (u32)(__ffs((int)~okb) - 1);
            back += cnt;
            if (cnt < 32u) break;
        }
        u32 const ms = probe - back;
        u32 const mlen = back + 4u + zb_count_fwd(base, probe + 4u, offset, be, lane);
        u32 const litLen = ms - anchor;
        u32 offBase;
        if (isRep && litLen > 0u) offBase = 1u;                 /* REPCODE1_TO_OFFBASE */
        else { offBase = offset + 3u; rep2 = rep1; rep1 = offset; }
        if (lane == 0) myseq[nbSeq] = zb_pack_seq(offBase, litLen, mlen);
        for (u32 i = lane; i < litLen; i += 32) mylit[litPos + i] = base[anchor + i];
        litPos += litLen; nbSeq++;
        ip = ms + mlen; anchor = ip;

        if (ip + 8u <= be) {
            if (lane == 0) {                                    /* zstd_fast.c:403-408 */
                table[zb_hash(zb_ld64u(base + probe + 2u), mls, hlog)] = (u16)(probe + 2u);
                table[zb_hash(zb_ld64u(base + ip - 2u), mls, hlog)] = (u16)(ip - 2u);
            }
            __syncwarp();
            while (ip + 8u <= be && rep2 != 0u) {               /* zstd_fast.c:410-420 */
                if (zb_ld32u(base + ip) != zb_ld32u(base + ip - rep2)) break;
                u32 const rlen = 4u + zb_count_fwd(base, ip + 4u, rep2, be, lane);
                { u32 const t = rep2; rep2 = rep1; rep1 = t; }
                if (lane == 0) {
                    table[zb_hash(zb_ld64u(base + ip), mls, hlog)] = (u16)ip;
                    myseq[nbSeq] = zb_pack_seq(1u, 0u, rlen);
                }
                nbSeq++;
                ip += rlen; anchor = ip;
                __syncwarp();
            }
        }
        searchStart = ip;
    }

    /* trailing literals (zstd_compress.c:3365-3366) */
    {   u32 const lastLits = be - anchor;
        for (u32 i = lane; i < lastLits; i += 32) mylit[litPos + i] = base[anchor + i];
        litPos += lastLits;
    }
    if (lane == 0) {
        ZbBlockMeta m; m.nbSeq = nbSeq; m.litSize = litPos; m.litSecSize = 0; m.bodySize = 0;
        m.type = ZB_BT_COMPRESSED; m.forceRaw = 0; m.rleByte = 0; m.pad = 0;
        meta[blockIdx.x] = m;
    }
}

extern "C" cudaError_t zb_launch_match(const u8* d_src, const ZbBlock* d_blocks, u32 nbBlocks, const ZbParams* prm,
                                       u64* d_seqs, u8* d_lits, ZbBlockMeta* d_meta, cudaStream_t stream)
{
    if (nbBlocks == 0) return cudaSuccess;
    size_t const smem = (size_t)2 << prm->hashLog;
    static int configured = 0;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(zb_match_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        configured = 1;
    }
    zb_match_fast_kernel<<<nbBlocks, 32, smem, stream>>>(d_src, d_blocks, *prm, d_seqs, d_lits, d_meta);
    return cudaGetLastError();
}
