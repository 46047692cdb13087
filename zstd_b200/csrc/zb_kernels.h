/* zb_kernels.h — launch shims of the CUDA kernels (C linkage, called by the host driver zb_api.cu). */
#ifndef ZB_KERNELS_H
#define ZB_KERNELS_H
#include <cuda_runtime.h>
#include "zb_common.h"

#ifdef __cplusplus
extern "C" {
#endif

/* K1: match-finder = K1a candidate walk (one CTA per chunk) + K1b greedy parse (one warp per 16 KiB segment) + K1c merge
 * (d_segmeta: ZB_PARSE_SEGS records per block).  d_blocks / d_chunks point at the first block / chunk of the launch;
 * the launch's blocks use the workspace rows [0, nbBlocks) of the arrays passed in, slotFirstBlock = index (in the
 * call's block array) of the block that owns row 0.  Per-block workspace strides come in `sd` (ZbStrides).
 * d_dist / d_far: sd->dist u16 + u32 per block (dead after this call; K3 reuses d_dist for the FSE state records);
 * d_dist2 / d_far2: same, only used by the doubleFast strategy (short-hash candidates).
 * d_dictEnd: one past the dictionary content in device memory (NULL = no dictionary); chunks / blocks with dictLen > 0
 * take the oldest dictLen bytes of their history from in front of it.  d_image (may be NULL): table already walked
 * over that dictionary tail by zb_launch_dict_image (same ZbParams), prm->tableN u32. */
cudaError_t zb_launch_dict_image(const u8* d_dictEnd, const ZbChunk* d_dictChunk, const ZbParams* prm, u32* d_image, cudaStream_t stream);
cudaError_t zb_launch_match(const u8* d_src, const u8* d_dictEnd, const u32* d_image, const ZbBlock* d_blocks, u32 nbBlocks,
                            const ZbChunk* d_chunks, u32 nbChunks, u32 slotFirstBlock, const ZbParams* prm, const ZbStrides* sd,
                            u16* d_dist, u32* d_far, u16* d_dist2, u32* d_far2, u64* d_seqs, u8* d_lits, ZbBlockMeta* d_meta, ZbSegMeta* d_segmeta,
                            cudaEvent_t evMid, cudaStream_t stream);

/* host: the format's predefined FSE tables (zb_dict.cu), and their upload to the current device (zb_sequences.cu) */
void zb_buildDefaultTables(ZbdFseCTable* out3);
cudaError_t zb_upload_default_tables(const ZbdFseCTable* host3, cudaStream_t stream);

/* host: parse a dictionary (zb_dict.cu).  Returns the content offset, 0 for raw content, or an error code */
size_t zb_loadDictionary(ZbDictEntropy* de, const u8* dict, size_t dictSize);

/* K2: literals section (histogram, Huffman table, 1/4-stream encode).  One CTA per block.
 * d_de (may be NULL): dictionary entropy state used by ZB_FLAG_DICT blocks. */
cudaError_t zb_launch_literals(const ZbBlock* d_blocks, u32 nbBlocks, const ZbParams* prm, const ZbStrides* sd, const ZbDictEntropy* d_de,
                               const u8* d_lits, u8* d_body, ZbBlockMeta* d_meta, cudaStream_t stream);

/* K3: sequences section (codes, histograms, FSE tables, tANS bit-stream) + block-type decision. */
cudaError_t zb_launch_sequences(const u8* d_src, const ZbBlock* d_blocks, u32 nbBlocks, const ZbParams* prm, const ZbStrides* sd, const ZbDictEntropy* d_de,
                                const u64* d_seqs, u16* d_stateBits, u8* d_body, ZbBlockMeta* d_meta, cudaStream_t stream);

/* K4: stitch — per-block output sizes -> exclusive scan -> frame/block headers + payload copy, for
 * one wave of blocks.  d_blocks/d_meta/d_body/d_outOffsets point at the wave's first block;
 * d_outOffsets gets nbBlocks+1 absolute offsets, starting at *d_base (NULL = 0); *d_total receives
 * the running total after this wave (even past dstCapacity: nothing is written past dst+dstCapacity). */
cudaError_t zb_launch_stitch(const u8* d_src, const ZbBlock* d_blocks, u32 nbBlocks, const ZbFrame* d_frames,
                             const u8* d_body, u32 bodyStride, const ZbBlockMeta* d_meta,
                             u64* d_outOffsets, const u64* d_base, u64* d_total,
                             u8* d_dst, u64 dstCapacity, cudaStream_t stream);
cudaError_t zb_launch_checksums(const u8* d_src, const ZbFrame* d_frames, u32 nbFrames, const u64* d_outOffsets, u8* d_dst, u64 dstCapacity, cudaStream_t stream);
cudaError_t zb_launch_frame_sizes(const ZbFrame* d_frames, u32 nbFrames, const u64* d_outOffsets,
                                  u64* d_frameSizes, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif
