#include "swb_plan.h"
#include <algorithm>
#include <string.h>

// Stable LSD radix sort of ids by DESCENDING length (16-bit digits; the high pass is skipped when all
// lengths fit 16 bits, which is the case for protein databases up to titin's 35k).
static void sort_by_length_desc(const std::vector<uint32_t> &len, std::vector<uint32_t> &ids)
{
    const uint32_t n = (uint32_t)len.size();
    ids.resize(n);
    for (uint32_t i = 0; i < n; ++i) ids[i] = i;
    uint32_t maxlen = 0;
    for (uint32_t i = 0; i < n; ++i) maxlen = std::max(maxlen, len[i]);
    std::vector<uint32_t> tmp(n);
    std::vector<uint32_t> cnt(65537);
    const int passes = maxlen > 0xffffu ? 2 : 1;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = 16 * pass;
        std::fill(cnt.begin(), cnt.end(), 0u);
        // descending: bucket index = 0xffff - digit
        for (uint32_t i = 0; i < n; ++i) cnt[(0xffffu - ((len[ids[i]] >> shift) & 0xffffu)) + 1]++;
        for (uint32_t d = 0; d < 65536; ++d) cnt[d + 1] += cnt[d];
        for (uint32_t i = 0; i < n; ++i) tmp[cnt[0xffffu - ((len[ids[i]] >> shift) & 0xffffu)]++] = ids[i];
        ids.swap(tmp);
    }
}

static uint8_t classify_logg(uint32_t len, uint32_t group_len)
{
    uint8_t lg = 0;
    uint64_t cap = group_len ? group_len : 1;
    while (lg < SWB_MAX_LOGG && len > cap) { cap <<= 1; ++lg; }
    return lg;
}

int swb_sort_by_length(const uint64_t *offsets, uint32_t n, std::vector<uint32_t> &order)
{
    std::vector<uint32_t> len(n);
    for (uint32_t i = 0; i < n; ++i) {
        if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] > 0x7ffffff0ull) return -1;
        len[i] = (uint32_t)(offsets[i + 1] - offsets[i]);
    }
    sort_by_length_desc(len, order);
    return 0;
}

// Head of the plan: which sequences the shard owns, in sorted order, with their lengths and offsets in the caller's
// buffer. That is all a sharded load needs to start gathering and uploading residues; the tail (output order, tiles)
// can be built beside the upload.
int swb_build_plan_head(const uint64_t *offsets, uint32_t n, uint32_t shard, uint32_t nshards, SwbPlan &plan,
                        const uint32_t *sorted_order)
{
    if (nshards == 0 || shard >= nshards) return -1;
    plan = SwbPlan();
    plan.n_total = n;
    plan.shard = shard;
    plan.nshards = nshards;
    std::vector<uint32_t> len(n);
    for (uint32_t i = 0; i < n; ++i) {
        if (offsets[i + 1] < offsets[i]) return -2;
        const uint64_t l = offsets[i + 1] - offsets[i];
        if (l > 0x7ffffff0ull) return -3;
        len[i] = (uint32_t)l;
    }
    plan.residues_total = n ? offsets[n] - offsets[0] : 0;
    std::vector<uint32_t> own_order;
    if (!sorted_order) {
        sort_by_length_desc(len, own_order);
        sorted_order = own_order.data();
    }
    const uint32_t *order = sorted_order;

    // residue-balanced sharding: deal PAIRS of the length-sorted list round-robin, so every shard gets the
    // same length mix and its residue total differs from the others by at most one pair per round
    if (nshards == 1) {
        plan.sorted_ids.assign(order, order + n);
    } else {
        plan.sorted_ids.reserve(n / nshards + 2);
        for (uint32_t p = shard; 2 * (uint64_t)p < n; p += nshards) {
            plan.sorted_ids.push_back(order[2 * (size_t)p]);
            if (2 * (uint64_t)p + 1 < n) plan.sorted_ids.push_back(order[2 * (size_t)p + 1]);
        }
    }
    plan.n_local = (uint32_t)plan.sorted_ids.size();
    const uint32_t nl = plan.n_local;
    plan.seq_off.resize(nl);
    plan.seq_len.resize(nl);
    for (uint32_t s = 0; s < nl; ++s) {
        plan.seq_off[s] = offsets[plan.sorted_ids[s]];
        plan.seq_len[s] = len[plan.sorted_ids[s]];
        plan.residues_local += plan.seq_len[s];
    }
    plan.max_len = nl ? plan.seq_len[0] : 0;
    return 0;
}

int swb_build_plan(const uint64_t *offsets, uint32_t n, uint32_t shard, uint32_t nshards, const SwbPlanOpts &o,
                   SwbPlan &plan, const uint32_t *sorted_order)
{
    const int rc = swb_build_plan_head(offsets, n, shard, nshards, plan, sorted_order);
    if (rc != 0) return rc;
    swb_build_plan_tail(o, plan);
    return 0;
}

// Tail of the plan: output order and tiles. Reads only what the head left in `plan` (not seq_off, which a sharded load
// rewrites to offsets in the gathered stream while this runs on a helper thread).
void swb_build_plan_tail(const SwbPlanOpts &o, SwbPlan &plan)
{
    const uint32_t n = plan.n_total, nl = plan.n_local, nshards = plan.nshards;
    // output order = ascending DB id; rank_of[id] = position of id among the shard's ids (O(n), no sort)
    plan.shard_ids.resize(nl);
    plan.out_pos.resize(nl);
    if (nshards == 1) {
        for (uint32_t s = 0; s < nl; ++s) {
            plan.shard_ids[s] = s;
            plan.out_pos[s] = plan.sorted_ids[s];
        }
    } else {
        std::vector<uint32_t> rank_of(n, 0xffffffffu);
        for (uint32_t s = 0; s < nl; ++s) rank_of[plan.sorted_ids[s]] = 0;
        uint32_t r = 0;
        for (uint32_t id = 0; id < n; ++id)
            if (rank_of[id] == 0) {
                plan.shard_ids[r] = id;
                rank_of[id] = r++;
            }
        for (uint32_t s = 0; s < nl; ++s) plan.out_pos[s] = rank_of[plan.sorted_ids[s]];
    }
    plan.sorted_of_out.resize(nl);
    for (uint32_t s = 0; s < nl; ++s) plan.sorted_of_out[plan.out_pos[s]] = s;

    // tiles: consecutive pairs of the sorted list; the group size follows the tile's longest sequence
    const uint32_t npairs_total = (nl + 1) / 2;
    memset(plan.tiles_by_logg, 0, sizeof plan.tiles_by_logg);
    memset(plan.cols_by_logg, 0, sizeof plan.cols_by_logg);
    uint32_t pair = 0;
    while (pair < npairs_total) {
        const uint32_t head_len = plan.seq_len[2 * (size_t)pair];
        SwbTile t;
        memset(&t, 0, sizeof t);
        t.logG = classify_logg(head_len, o.group_len);
        const uint32_t slots = 32u >> t.logG;
        t.first_pair = pair;
        t.npairs = (uint16_t)std::min<uint32_t>(slots, npairs_total - pair);
        t.width = swb_roundup(head_len, SWB_COLS_PER_CHUNK);
        plan.tiles.push_back(t);
        plan.tiles_by_logg[t.logG]++;
        pair += t.npairs;
    }
    // The walk above already produced the order the kernels want: group sizes from 32 lanes down to 1, longest
    // tile first inside each, so a launch that takes its tiles in array order behaves like LPT list scheduling.
    {
        uint32_t at = 0;
        for (int l = SWB_MAX_LOGG; l >= 0; --l) {
            plan.tile_start_by_logg[l] = at;
            at += plan.tiles_by_logg[l];
        }
        // tiles are in descending width over the lane-group sizes, so "wider than xl_len" is a prefix of the array
        plan.n_xl = 0;
        memset(plan.xl_by_logg, 0, sizeof plan.xl_by_logg);
        for (int l = SWB_MAX_LOGG; l >= 1 && o.xl_len; --l) {
            const uint32_t s0 = plan.tile_start_by_logg[l];
            uint32_t k = 0;
            while (k < plan.tiles_by_logg[l] && plan.tiles[s0 + k].width > o.xl_len) ++k;
            plan.xl_by_logg[l] = k;
            plan.n_xl += k;
            if (k < plan.tiles_by_logg[l]) break;
        }
    }
    uint64_t res = 0, bnd = 0;
    for (size_t i = 0; i < plan.tiles.size(); ++i) {
        SwbTile &t = plan.tiles[i];
        const uint64_t slots = 32u >> t.logG;
        t.res_off = res;
        t.bnd_off = bnd;
        res += (uint64_t)t.width * slots * 2u;
        bnd += (uint64_t)t.width * slots;
        plan.padded_cols += (uint64_t)t.width * slots * 2u;
        plan.cols_by_logg[t.logG] += (uint64_t)t.width * slots * 2u;
    }
    plan.res_bytes = res;
    plan.bnd_elems = bnd;
}

void swb_plan_query(uint32_t qlen, int k_force, int k_max, uint32_t logg_present, uint32_t chunk_rows,
                    SwbQueryPlan &qp)
{
    qp.chunks.clear();
    qp.prof_rows = 0;
    const uint32_t nchunks = (qlen + chunk_rows - 1) / chunk_rows;
    // rows the dominant chunk has: all chunks but the last are chunk_rows long
    const uint32_t typical = nchunks > 1 ? chunk_rows : qlen;
    for (int l = 0; l <= SWB_MAX_LOGG; ++l) {
        int best_k = std::min(32, k_max);
        if (k_force) {
            best_k = std::min(k_force, k_max);
            while (best_k > 8 && nchunks > 1 && chunk_rows % ((uint32_t)best_k << l)) best_k >>= 1;
        } else {
            double best = 0;
            bool have = false;
            for (int K = std::min(32, k_max); K >= 8; K >>= 1) {
                // a chunk boundary must be a pass boundary of every group size
                if (nchunks > 1 && chunk_rows % ((uint32_t)K << l)) continue;
                // padded rows x relative cost of a padded cell with K rows per lane, measured on B200 with all group
                // sizes forced to one K (9024 / 7655 / 5228 GCUPS for K = 32 / 16 / 8): short strips pay the per-column
                // work (residue fetch, shuffles, boundary I/O, loop control) over fewer cells
                const double per_cell = K >= 32 ? 1.0 : (K == 16 ? 1.18 : 1.73);
                const double cost = (double)swb_roundup(typical, (uint32_t)K << l) * per_cell;
                if (!have || cost < best) { best = cost; best_k = K; have = true; }
            }
        }
        qp.k_by_logg[l] = best_k;
    }
    for (uint32_t c = 0; c < nchunks; ++c) {
        SwbQueryChunk ch;
        ch.row0 = c * chunk_rows;
        ch.rows = std::min(chunk_rows, qlen - ch.row0);
        uint32_t need = ch.rows;
        for (int l = 0; l <= SWB_MAX_LOGG; ++l)
            if (logg_present & (1u << l)) need = std::max(need, swb_roundup(ch.rows, (uint32_t)qp.k_by_logg[l] << l));
        ch.smem_rows = swb_roundup(need, 128);
        ch.first = c == 0;
        ch.last = c + 1 == nchunks;
        qp.chunks.push_back(ch);
        qp.prof_rows = std::max(qp.prof_rows, ch.row0 + ch.smem_rows);
    }
}

void swb_plan_bulk_groups(const SwbPlan &plan, const SwbQueryPlan &qp, bool longest_first, const uint32_t *skip_by_logg,
                          uint32_t logg_mask, std::vector<SwbLaunchGroup> &groups)
{
    const size_t first = groups.size();
    for (int K = 32; K >= 8; K >>= 1) {
        SwbLaunchGroup g;
        memset(&g, 0, sizeof g);
        g.K = K;
        uint32_t nr = 0;
        for (int l = SWB_MAX_LOGG; l >= 0; --l) {
            if (!(logg_mask & (1u << l)) || !plan.tiles_by_logg[l] || qp.k_by_logg[l] != K) continue;
            const uint32_t skip = skip_by_logg ? std::min(skip_by_logg[l], plan.tiles_by_logg[l]) : 0u;
            if (plan.tiles_by_logg[l] == skip) continue;
            g.logg_mask |= 1u << l;
            g.range_start[nr] = plan.tile_start_by_logg[l] + skip;
            g.ntiles += plan.tiles_by_logg[l] - skip;
            g.range_cum[nr] = g.ntiles;
            ++nr;
        }
        if (!nr) continue;
        for (uint32_t r = nr; r < SWB_MAX_RANGES; ++r) {  // unused ranges can never be selected
            g.range_start[r] = 0;
            g.range_cum[r] = g.ntiles;
        }
        groups.push_back(g);
    }
    // launch order. longest_first (a lone query): the group that owns the largest group size, i.e. the longest
    // tiles, goes first and the bulk last; a small group occupies few blocks, so the bulk still finds room and runs
    // beside it, and the long tiles are not left for the end. Otherwise (a batch: the next query fills the GPU
    // while this one drains) the bulk goes first.
    if (longest_first)
        std::stable_sort(groups.begin() + first, groups.end(),
                         [](const SwbLaunchGroup &a, const SwbLaunchGroup &b) { return a.logg_mask > b.logg_mask; });
    else
        std::stable_sort(groups.begin() + first, groups.end(),
                         [](const SwbLaunchGroup &a, const SwbLaunchGroup &b) { return a.ntiles > b.ntiles; });
}

bool swb_plan_split_group(const SwbPlan &plan, const uint32_t *first_by_logg, const uint32_t *count_by_logg, int K,
                          SwbLaunchGroup &g)
{
    memset(&g, 0, sizeof g);
    g.K = K;
    g.split = true;
    for (int l = 1; l <= SWB_MAX_LOGG; ++l) {
        const uint32_t first = first_by_logg ? first_by_logg[l] : 0u;
        if (first >= plan.tiles_by_logg[l]) continue;
        g.xl_start[l] = plan.tile_start_by_logg[l] + first;
        g.xl_by_logg[l] = std::min(count_by_logg[l], plan.tiles_by_logg[l] - first);
        if (g.xl_by_logg[l]) g.logg_mask |= 1u << l;
        g.ntiles += g.xl_by_logg[l];
    }
    for (uint32_t r = 0; r < SWB_MAX_RANGES; ++r) g.range_cum[r] = g.ntiles;
    return g.ntiles != 0;
}

void swb_group_chunks(const SwbQueryPlan &qp, const SwbLaunchGroup &g, std::vector<SwbQueryChunk> &out)
{
    out = qp.chunks;
    if (!g.split || out.size() < 2) return;
    SwbQueryChunk all = out.front();
    all.rows = out.back().row0 + out.back().rows;
    all.first = all.last = 1;
    out.assign(1, all);
}

uint32_t swb_split_items(uint32_t rows, const SwbLaunchGroup &g, SwbScoreParams *p)
{
    uint32_t items = 0;
    for (int j = 0; j < SWB_MAX_LOGG; ++j) {
        const int l = SWB_MAX_LOGG - j;
        items += g.xl_by_logg[l] * swb_split_passes(rows, l, g.K);
        if (p) {
            p->split_tile_start[j] = g.xl_start[l];
            p->split_item_end[j] = items;
        }
    }
    if (p) p->ntiles = items;
    return items;
}

uint32_t swb_group_smem_rows(uint32_t rows, const SwbLaunchGroup &g)
{
    uint32_t need = rows;
    for (int l = 0; l <= SWB_MAX_LOGG; ++l)
        if (g.logg_mask & (1u << l)) need = std::max(need, swb_roundup(rows, (uint32_t)g.K << l));
    return swb_roundup(need, 128);
}

int swb_rebase_shift(int max_s, int min_s, int gap, uint32_t pass_rows)
{
    // all values of a pass over one block lie within (pass_rows + cols + 2) * step of the block's base (swb_warp.cuh,
    // V16R); they, the transient diag + S and the clamped floors (-32000) must stay inside [-32768, 32767]
    const int step = std::max(1, max_s + gap);
    const int lowest = std::max(0, -min_s);
    const int span = 31000 - lowest - gap - std::max(0, max_s);
    const long cols = (long)span / step - (long)pass_rows - 2;
    int shift = 0;
    while (shift < 15 && (2L << shift) <= cols) ++shift;
    return shift >= 6 ? shift : 0;
}
