// Host-side database plan: length sort, residue-balanced sharding, pairing and tiling.
// Pure C++ (no CUDA) so it can be exercised on CPU-only machines through the C ABI (swb_plan_*).
#pragma once
#include <stdint.h>
#include <vector>
#include "swb_types.h"

struct SwbPlanOpts {
    uint32_t group_len;   // sequences up to this length run one lane per pair (G=1); each doubling of the
                          // length doubles G up to 32
    uint32_t xl_len;      // lane-group tiles wider than this can run their passes as pipelined work items (the
                          // engine decides per query whether they do); 0 = never
    SwbPlanOpts() : group_len(384), xl_len(3072) {}
};

struct SwbPlan {
    uint32_t n_total;                 // sequences in the whole database
    uint32_t n_local;                 // sequences of this shard
    uint32_t shard, nshards;
    std::vector<uint32_t> sorted_ids; // [n_local] DB ids, longest first, ties in DB order
    std::vector<uint32_t> out_pos;    // [n_local] position of sorted entry s in the shard's output order
    std::vector<uint32_t> sorted_of_out;  // [n_local] inverse of out_pos
    std::vector<uint32_t> shard_ids;  // [n_local] DB ids in output order (ascending)
    std::vector<uint64_t> seq_off;    // [n_local] offset of sorted entry s in the raw code buffer
    std::vector<uint32_t> seq_len;    // [n_local]
    std::vector<SwbTile> tiles;       // by group size (32 lanes per pair first), longest first inside a group size
    uint64_t res_bytes;               // packed residue buffer size
    uint64_t bnd_elems;               // boundary scratch elements
    uint64_t residues_local;          // true residues of this shard
    uint64_t residues_total;
    uint64_t padded_cols;             // sum over tiles of width * slots * 2 (cells per query row incl. padding)
    uint32_t max_len;                 // longest sequence of the shard
    uint32_t tiles_by_logg[SWB_MAX_LOGG + 1];
    uint32_t tile_start_by_logg[SWB_MAX_LOGG + 1];  // tiles are stored by group size, 32 lanes first
    uint32_t n_xl;                    // leading tiles wider than xl_len (lane-group tiles only: the long sequences)
    uint32_t xl_by_logg[SWB_MAX_LOGG + 1];  // of which per group size (a prefix of each group size's tiles)
    uint64_t cols_by_logg[SWB_MAX_LOGG + 1];  // padded sequence-columns (width * slots * 2) per group size
};

// offsets: n+1 entries. Returns 0 or a negative error (lengths above 2^31-8). sorted_order: the n ids by descending
// length as swb_sort_by_length returns them (several shards of one database share one sort), or NULL to sort here.
int swb_build_plan(const uint64_t *offsets, uint32_t n, uint32_t shard, uint32_t nshards, const SwbPlanOpts &o,
                   SwbPlan &plan, const uint32_t *sorted_order = nullptr);
// The same in two steps, for a sharded load that uploads while the tail is built: the head leaves sorted_ids, seq_len,
// seq_off, residues_local / _total and max_len; the tail adds the output order and the tiles.
int swb_build_plan_head(const uint64_t *offsets, uint32_t n, uint32_t shard, uint32_t nshards, SwbPlan &plan,
                        const uint32_t *sorted_order = nullptr);
void swb_build_plan_tail(const SwbPlanOpts &o, SwbPlan &plan);
// ids 0..n-1 by descending length, ties in id order (stable); -1 on decreasing offsets / a length above 2^31-16
int swb_sort_by_length(const uint64_t *offsets, uint32_t n, std::vector<uint32_t> &order);

// How one query is cut into score-kernel launches ("chunks" of query rows that fit shared memory), and how many
// query rows a lane holds for each group size.
struct SwbQueryChunk {
    uint32_t row0;       // first query row
    uint32_t rows;       // query rows of the chunk
    uint32_t smem_rows;  // rows staged per code (covers the padding of every group size, multiple of 128)
    uint32_t first, last;
};
struct SwbQueryPlan {
    int k_by_logg[SWB_MAX_LOGG + 1];  // rows per lane (8, 16 or 32) for tiles of 1 << l lanes per pair
    uint32_t prof_rows;               // rows the global profile must provide (rows beyond qlen score 0)
    std::vector<SwbQueryChunk> chunks;
};
// k_force: 0 = choose per group size (least padded rows, which is also the shortest tile time), else 8/16/32.
// k_max: 32 for the s16 passes, 16 for the int32 pass. logg_present: bit l set when the plan has tiles of that
// group size. chunk_rows must be a multiple of 1024 (every K << l divides it).
void swb_plan_query(uint32_t qlen, int k_force, int k_max, uint32_t logg_present, uint32_t chunk_rows,
                    SwbQueryPlan &qp);

// One score-kernel launch of a query pass: the tiles of all group sizes that share the same K.
struct SwbLaunchGroup {
    int K;
    bool split;          // pipelined passes: the work items are (tile, pass), one warp per block
    uint32_t xl_start[SWB_MAX_LOGG + 1];    // split group: first tile (index into the tile array) per group size ...
    uint32_t xl_by_logg[SWB_MAX_LOGG + 1];  // ... and how many
    uint32_t logg_mask;
    uint32_t ntiles;
    uint32_t range_start[SWB_MAX_RANGES];
    uint32_t range_cum[SWB_MAX_RANGES];
};
// The bulk groups, one per distinct K, over the group sizes in logg_mask; skip_by_logg[l] (may be NULL) leading tiles of
// group size l are left out (they run in split groups). longest_first puts the group owning the longest tiles first (a
// lone query), otherwise the group with most tiles first (batch); inside a group the ranges run from the largest
// group size (longest tiles) down. Appends to `groups`.
void swb_plan_bulk_groups(const SwbPlan &plan, const SwbQueryPlan &qp, bool longest_first, const uint32_t *skip_by_logg,
                          uint32_t logg_mask, std::vector<SwbLaunchGroup> &groups);
// A split group over count_by_logg[l] tiles of group size l >= 1 starting first_by_logg[l] tiles into that group
// size; K rows per lane (8 or 16). Returns false when the set is empty.
bool swb_plan_split_group(const SwbPlan &plan, const uint32_t *first_by_logg, const uint32_t *count_by_logg, int K,
                          SwbLaunchGroup &g);
// passes (work items) of a tile of 1 << l lanes per pair for `rows` query rows
inline uint32_t swb_split_passes(uint32_t rows, int l, int K)
{
    return (rows + ((uint32_t)K << l) - 1u) / ((uint32_t)K << l);
}
// work items of a split group for `rows` query rows; with p != NULL also fills p->ntiles and the class tables
uint32_t swb_split_items(uint32_t rows, const SwbLaunchGroup &g, SwbScoreParams *p);
// The launches of a group over the query rows: the chunks of the query plan, except for a split group: it stages per
// work item, shared memory does not limit it, so it covers all rows in ONE launch (no drain between chunks, several
// times the work items in flight).
void swb_group_chunks(const SwbQueryPlan &qp, const SwbLaunchGroup &g, std::vector<SwbQueryChunk> &out);
// rows a launch group must find in shared memory for a chunk of `rows` query rows (multiple of 128)
uint32_t swb_group_smem_rows(uint32_t rows, const SwbLaunchGroup &g);
// V16R: log2 of the columns per rebase block for a linear scheme (largest score max_s, smallest min_s, gap) with passes
// of up to `pass_rows` rows, or 0 when the scheme's steps are too large for the 16-bit window (use V32 then)
int swb_rebase_shift(int max_s, int min_s, int gap, uint32_t pass_rows);
