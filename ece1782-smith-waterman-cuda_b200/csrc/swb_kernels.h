// Internal C++ interface between the engine (swb_engine.cu) and the kernels (swb_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "swb_types.h"

// block shapes of the score kernel: SMALL when the staged profile leaves room for several blocks per SM,
// LARGE (one block per SM) when the profile of a long query fills most of the 227 KB of shared memory.
#define SWB_BLOCK_SMALL 0
#define SWB_BLOCK_LARGE 1
#ifndef SWB_NT_SMALL
#define SWB_NT_SMALL 256
#endif
#define SWB_MINB_SMALL 2
#ifndef SWB_NT_LARGE
#define SWB_NT_LARGE 512
#endif

// arithmetic policy of a score launch (swb_warp.cuh)
#define SWB_MODE_S16 0    // V16: two DB sequences per lane, one query
#define SWB_MODE_I32 1    // V32: exact recompute of flagged tiles
#define SWB_MODE_R16 2    // V16R: rebased s16x2, exact (K = 8, 16): recompute of flagged tiles, long-against-long tiles
#define SWB_MODE_S16A 3   // V16A: affine gaps, s16x2 (K = 8, 16)
#define SWB_MODE_I32A 4   // V32A: affine gaps, exact recompute (K = 8)

// K: query rows per lane (8, 16, 32; exact passes 8 or 16). split: the passes of a tile are separate, pipelined work
// items (very long sequences; SWB_MODE_S16 / SWB_MODE_R16 with K = 8, 16, SWB_MODE_I32 with K = 8).
cudaError_t swb_launch_score(int K, int mode, bool split, int block_cfg, const SwbScoreParams &p, int grid, size_t smem,
                             cudaStream_t st);
// zeroes the scores of every flagged tile (before an int32 pass with pipelined work items)
cudaError_t swb_launch_clear_flagged(const SwbTile *tiles, uint32_t ntiles, const uint8_t *flags, int32_t *scores,
                                     cudaStream_t st);
cudaError_t swb_score_occupancy(int K, int mode, bool split, int block_cfg, size_t smem, int *blocks_per_sm);
cudaError_t swb_launch_profile(const uint8_t *q, uint32_t qlen, const int8_t *mat, int bias, int8_t *prof,
                               uint32_t stride, uint32_t rows, cudaStream_t st);
cudaError_t swb_launch_pack(const SwbTile *tiles, uint32_t ntiles, const uint8_t *raw, const uint64_t *seq_off,
                            const uint32_t *seq_len, uint32_t nseq, uint8_t *residues, cudaStream_t st);
// k best (score descending, position ascending) of a score vector; ids: database id per position or NULL (position)
#define SWB_TOPK_MAX 1024
cudaError_t swb_launch_topk(const int32_t *scores, uint32_t n, const uint32_t *ids, uint32_t k, uint32_t *out_ids,
                            int32_t *out_scores, cudaStream_t st);
cudaError_t swb_launch_scatter(const int32_t *sorted, const uint32_t *dst, uint32_t n, int32_t *out,
                               cudaStream_t st);
// traceback alignments of a list of hits (cpu.cpp semantics; affine: Gotoh's three states), one block per job. hd_glob:
// rolling diagonals of the jobs whose swb_align_hd_ints(m, affine) ints exceed smem_ints (SwbAlignJob::hd_off); dir:
// packed directions, (m + 1) * swb_align_row_bytes(n, affine) bytes per job at dir_off; out_hdr: 5 ints per job {score,
// end_i, end_j, nops, ops_overflow}; out_ops: per job, from the END of the alignment back to its start
#define SWB_ALIGN_NT 256
#define SWB_ALIGN_SMEM_MAX (216u * 1024u)
cudaError_t swb_launch_align_batch(const SwbAlignJob *jobs, uint32_t njobs, const uint8_t *qbuf, const uint8_t *raw,
                                   const int8_t *mat, int gap_open, int gap_extend, bool affine, int32_t *hd_glob,
                                   uint8_t *dir, int32_t *out_hdr, uint8_t *out_ops, uint32_t smem_ints, cudaStream_t st);
