/* zb_device.cuh — small device helpers shared by the kernels. */
#ifndef ZB_DEVICE_CUH
#define ZB_DEVICE_CUH
#include <cuda_runtime.h>
#include "zb_common.h"

#define ZB_FULL 0xFFFFFFFFu

/* Unaligned little-endian loads built from aligned 32-bit words.  Only words that contain a
 * requested byte are dereferenced, so a load never leaves the 4-byte word of the last byte. */
__device__ __forceinline__ u32 zb_ld32u(const u8* p)
{
    const u32* q = (const u32*)((uintptr_t)p & ~(uintptr_t)3);
    u32 const sh = ((u32)(uintptr_t)p & 3u) * 8u;
    u32 const a = __ldg(q);
    u32 const b = sh ? __ldg(q + 1) : 0u;
    return __funnelshift_r(a, b, sh);
}
__device__ __forceinline__ u64 zb_ld64u(const u8* p)
{
    const u32* q = (const u32*)((uintptr_t)p & ~(uintptr_t)3);
    u32 const sh = ((u32)(uintptr_t)p & 3u) * 8u;
    u32 const a = __ldg(q);
    u32 const b = __ldg(q + 1);
    u32 const c = sh ? __ldg(q + 2) : 0u;
    u32 const lo = __funnelshift_r(a, b, sh);
    u32 const hi = __funnelshift_r(b, c, sh);
    return ((u64)hi << 32) | lo;
}

/* /root/reference/lib/compress/zstd_compress_internal.h:815-861 */
__device__ __forceinline__ u32 zb_hash(u64 v, u32 mls, u32 hBits)
{
    switch (mls) {
    default:
    case 4: return ((u32)v * 2654435761u) >> (32 - hBits);
    case 5: return (u32)(((v << 24) * 889523592379ull) >> (64 - hBits));
    case 6: return (u32)(((v << 16) * 227718039650203ull) >> (64 - hBits));
    case 7: return (u32)(((v << 8) * 58295818150454627ull) >> (64 - hBits));
    case 8: return (u32)((v * 0xCF1BBCDCB7A56463ull) >> (64 - hBits));
    }
}

__device__ __forceinline__ u32 zb_hb32(u32 v) { return 31u - (u32)__clz((int)v); }

#endif
