// Shared host/device types of the B200 Smith-Waterman scan engine.
#pragma once
#include <stdint.h>

#define SWB_ALPHA 32          // codes per residue (5 bits in a byte)
#define SWB_PAD 31            // padding code: zero row and column in every scoring matrix
#define SWB_STAR 24           // '*' / unknown (SWSolver.cu:41, 119 of the reference)
#define SWB_COLS_PER_CHUNK 4  // DB columns per residue / boundary chunk
#define SWB_MAX_LOGG 5        // lane-group sizes 1,2,4,8,16,32
#define SWB_MAX_RANGES 6      // tile ranges per launch (one per lane-group size)

#if defined(__CUDACC__)
#define SWB_HD __host__ __device__ __forceinline__
#else
#define SWB_HD inline
#endif

// One warp-tile: 32/G pairs of DB sequences of similar length, G lanes per pair.
// Residues: [chunk c][slot p][4 columns x (seqA, seqB)] bytes, 8 bytes per (c, p), slots = 32 >> logG.
// Boundary row scratch: width * slots elements (layout depends on logG, see swb_warp.cuh).
struct SwbTile {
    uint64_t res_off;     // byte offset into the packed residue buffer
    uint64_t bnd_off;     // element offset into the boundary buffer
    uint32_t first_pair;  // pair index (sorted order) of slot 0; sequences 2*pair, 2*pair+1
    uint32_t width;       // columns (multiple of 4) = longest sequence of the tile rounded up
    uint16_t npairs;      // occupied slots
    uint8_t logG;         // log2 of lanes per pair
    uint8_t reserved;
};

struct SwbScoreParams {
    const SwbTile *tiles;     // all tiles of the shard, grouped by lane-group size, longest first
    uint32_t ntiles;          // tiles covered by this launch (sum of its ranges)
    const uint8_t *residues;
    void *bnd;                // boundary scratch (uint32 per element for s16x2, 2x int32 for i32)
    const int8_t *profile;    // global query profile [32][prof_stride], one byte per entry = S(q_row, code) + gap
    uint32_t prof_stride;     // bytes per code row in global memory
    uint32_t row0;            // first query row of this launch (query chunk)
    uint32_t rows;            // query rows of this chunk that carry real residues or padding to use
    uint32_t smem_rows;       // rows staged per code in shared memory (multiple of 128)
    uint32_t first_chunk;     // 1: top boundary is zero
    uint32_t last_chunk;      // 1: bottom boundary is not stored
    int32_t *scores;          // per sequence, sorted order
    uint32_t *counter;        // dynamic tile counter (zeroed before launch)
    uint8_t *flags;           // per tile: s16 kernel sets 1 when a score may have wrapped
    uint32_t only_flagged;    // i32 recompute: skip tiles whose flag is 0
    uint32_t *recount;        // i32 recompute: number of tiles re-scored (may be null)
    int32_t gap;              // linear policies: gap penalty
    int32_t gap_open;         // affine policies: first residue of a gap
    int32_t gap_extend;       // affine policies: every further residue
    int32_t ovf_thr;          // s16: best > ovf_thr  =>  recompute in int32
    // tiles of this launch: positions [0, ntiles) of the concatenation of up to SWB_MAX_RANGES ranges of `tiles`
    uint32_t range_start[SWB_MAX_RANGES];
    uint32_t range_cum[SWB_MAX_RANGES];  // cumulative tile count up to and including range r
    // SPLIT launches (long sequences): ntiles counts (tile, pass) items, handed out tile by tile, pass by pass;
    // classes run from 32 lanes per pair (j = 0) down to 2 (j = 4): class j is a run of tiles starting at
    // split_tile_start[j], each owning ceil(rows / (K << (5 - j))) items; split_item_end[j] = items of classes 0..j
    uint32_t split_tile_start[SWB_MAX_LOGG];
    uint32_t split_item_end[SWB_MAX_LOGG];
    uint32_t warps_active;    // warps of a block that take work (0 = all): a launch with few tiles spreads them over the SMs
    uint32_t static_wave;     // bulk launches: work items of the first wave, one per working warp of the grid, taken by
                              // position (block-coherent); the counter hands out the items after them. 0 = all dynamic
    uint32_t *prog;           // [item] columns of its bottom row that a pass has published (zeroed per query)
    // V16R (rebased s16): columns per block = 1 << rebase_shift (chosen by the host from the scoring scheme so that a
    // pass over one block spans less than 2^15 score points); blog = base log of the boundary rows, two int32 per
    // (pass parity, slot, block), see swb_blog_offset
    uint32_t rebase_shift;
    void *blog;
    // one-lane-per-pair tiles: per resident warp, the row state of the passes of a group between two column blocks
    // (swb_run_tile); swb_colstate_elems(K) elements of the policy's type per warp slot
    void *colstate;
};

// Bytes per shared-memory profile load of the bulk kernels (= the padding between code rows: 4 puts the rows one bank
// apart, so the 32 lanes of an LDS.32 never conflict; 8 halves the LDS count -- 4.5 % fewer issue slots in the one-lane
// loop -- but rows that are 16 codes apart share banks. Measured (profiles/r2r_sweep_lds64_ab.txt): 8 is 0.3 % slower on
// the benchmark, 1 % with group_len 384, 2.6 % on the half-database rank workload -> 4)
#ifndef SWB_BULK_LDW
#define SWB_BULK_LDW 4
#endif

// One-lane tiles: passes that walk over the column blocks together (swb_run_tile). 1 = the straight order, every pass over
// the whole width -- the product default: the blocked order (4 passes x 16 chunks) cut the DRAM traffic of a 5,478-row
// launch from 140 GB to 76 GB and raised the L2 hit rate from 58 % to 76 %, but ran 3.5 % SLOWER (8,952 against 9,278
// GCUPS on the same GPU, profiles/r2n_*): the kernel is bound by the ALU pipe, not by memory, and the extra address
// arithmetic of the nested loops lands on that pipe. Kept behind this macro as the measured answer to "why not block it".
#ifndef SWB_PASS_GROUP
#define SWB_PASS_GROUP 1u
#endif
SWB_HD size_t swb_colstate_elems(int K) { return (size_t)SWB_PASS_GROUP * (size_t)(K + 4) * 32u; }

// one traceback alignment of swb_align_batch (one block of swb_align_batch_kernel)
struct SwbAlignJob {
    uint64_t q_off;    // query codes, byte offset into the uploaded query buffer
    uint64_t d_off;    // subject codes, byte offset into the raw residue buffer of the shard
    uint64_t dir_off;  // byte offset of the job's direction matrix
    uint64_t ops_off;  // byte offset of the job's ops in the output
    uint64_t hd_off;   // int offset of the job's H diagonals in the global scratch (only when they do not fit shared memory)
    uint32_t m, n;     // query rows, subject columns
    uint32_t cap;      // ops capacity
    uint32_t reserved;
};

// ints of working storage per alignment job: three H diagonals (affine: plus two each for E and F) of m + 2 entries and
// one direction byte per row; bytes of one row of packed directions (2 bits per cell, affine 4) for columns 0 .. n
SWB_HD uint64_t swb_align_hd_ints(uint32_t m, bool affine)
{
    return (affine ? 7ull : 3ull) * ((uint64_t)m + 2u) + (((uint64_t)m + 5u) >> 2);
}
SWB_HD uint32_t swb_align_row_bytes(uint32_t n, bool affine) { return affine ? (n + 2u) >> 1 : (n + 4u) >> 2; }

// base log (V16R): element offset (uint2) of a tile's region, 2 * (((width * slots) >> 6) + 33) elements long
SWB_HD size_t swb_blog_offset(uint64_t bnd_off, uint32_t tile_idx) { return 2u * ((size_t)(bnd_off >> 6) + 34u * (size_t)tile_idx); }
SWB_HD size_t swb_blog_elems(uint64_t bnd_elems, uint32_t ntiles) { return 2u * ((size_t)(bnd_elems >> 6) + 34u * (size_t)ntiles) + 68u; }

SWB_HD uint32_t swb_roundup(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }
