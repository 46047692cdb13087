/* zb_bitpack.cuh — parallel bit packing into a little-endian bit-stream.
 *
 * Each thread knows (from a prefix sum over bit counts) the absolute bit position at which its
 * run of fields starts, accumulates fields in a 64-bit register and emits whole 32-bit words.
 * Words that may be shared with a neighbouring thread (the first and the last one it touches) are
 * merged with atomicOr into pre-zeroed memory; interior words are plain stores.
 * Bit order is that of BIT_addBits (/root/reference/lib/common/bitstream.h:179-188): a field
 * occupies [pos, pos+nbBits), least significant bit first.
 */
#ifndef ZB_BITPACK_CUH
#define ZB_BITPACK_CUH
#include "zb_device.cuh"

struct ZbdParW { u32* words; u64 acc; u32 nacc; u32 widx; u32 first; };

__device__ __forceinline__ void zbd_pw_init(ZbdParW* w, u32* words, u64 bitPos)
{
    w->words = words; w->acc = 0; w->nacc = (u32)(bitPos & 31u); w->widx = (u32)(bitPos >> 5); w->first = 1;
}
/* value must already be < 2^nbBits ; nbBits <= 31 */
__device__ __forceinline__ void zbd_pw_add(ZbdParW* w, u32 value, u32 nbBits)
{
    w->acc |= (u64)value << w->nacc;
    w->nacc += nbBits;
    if (w->nacc >= 32u) {
        u32 const lo = (u32)w->acc;
        if (w->first) { if (lo) atomicOr(&w->words[w->widx], lo); w->first = 0; }
        else w->words[w->widx] = lo;
        w->widx++;
        w->acc >>= 32;
        w->nacc -= 32u;
    }
}
/* the same in two halves for short fields (Huffman codes, <= 12 bits): zbd_pw_put only accumulates, zbd_pw_flush — due at
 * least once per two fields: fewer than 32 bits stay behind a flush, 32 + 2 * 12 < 64 — emits a word when one is full.
 * Written without branches: lanes of a warp fill their words at different moments, a branch would make every lane's
 * iteration pay for the emit path. */
__device__ __forceinline__ void zbd_pw_put(ZbdParW* w, u32 value, u32 nbBits)
{
    w->acc |= (u64)value << w->nacc;
    w->nacc += nbBits;
}
__device__ __forceinline__ void zbd_pw_flush(ZbdParW* w)
{
    bool const full = w->nacc >= 32u;
    u32 const lo = (u32)w->acc;
    u32* const wp = w->words + w->widx;
    if (full && w->first == 0u) *wp = lo;
    if (full && w->first != 0u) atomicOr(wp, lo);
    w->first = full ? 0u : w->first;
    w->widx += full ? 1u : 0u;
    w->acc = full ? (w->acc >> 32) : w->acc;
    w->nacc -= full ? 32u : 0u;
}
__device__ __forceinline__ void zbd_pw_finish(ZbdParW* w)
{
    u32 const lo = (u32)w->acc;
    if (w->nacc && lo) atomicOr(&w->words[w->widx], lo);
}

#endif
