// TEST INFRASTRUCTURE (never linked into libswb.so): runs the SAME warp program the GPU runs
// (csrc/swb_warp.cuh: swb_warp_loop / swb_run_tile / swb_column / swb_pack_word) on the CPU, one fiber
// per lane, warp shuffles emulated by a lock-step exchange. It exists so that the indexing, wavefront,
// boundary-scratch, chunking and overflow-recompute logic can be checked against the oracle in the
// build container, which has no GPU. DPX intrinsics use the host implementations CUDA ships.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>
#include <ucontext.h>
#include <algorithm>
#include <vector>

#include "swb_plan.h"
#include "swb_warp.cuh"

namespace {

struct WarpSim;

struct HostBackend {
    WarpSim *w;
    int lane_id;
    int lane() const { return lane_id; }
    uint32_t warp_slot() const { return 0; }  // one warp at a time
    uint32_t static_item(uint32_t) const { return 0; }
    uint32_t shfl_up(uint32_t v, int d, int width);
    uint32_t shfl_xor(uint32_t v, int m, int width);
    void syncwarp();
    bool any(bool f);
    uint32_t next_tile(uint32_t *counter);
    void count(uint32_t *p) { (*p)++; }
    void atomic_max(int32_t *p, int32_t v) { if (v > *p) *p = v; }
    void publish(uint32_t *p, uint32_t v) { *p = v; }
    // the emulation runs the work items one after the other in hand-out order, so whatever a pass waits for must
    // already be there: a value that is too small here would be a deadlock on the GPU
    void stage_rows(int8_t *dst, uint32_t dstride, const int8_t *prof, uint32_t pstride, uint32_t row0, uint32_t rows)
    {
        if (lane_id == 0)
            for (uint32_t code = 0; code < SWB_ALPHA; ++code)
                memcpy(dst + (size_t)code * dstride, prof + (size_t)code * pstride + row0, rows);
        syncwarp();
    }
    uint32_t wait_progress(const uint32_t *p, uint32_t need, uint32_t)
    {
        if (*p < need) abort();  // on the GPU this would be a spin that never ends: hand-out order guarantee broken
        return *p;
    }

    void prefetch_l2(const void *) const {}
    uint8_t ld_flag(const uint8_t *p) const { return *p; }
    SwbTile ld_tile(const SwbTile *p) const { return *p; }
    uint32_t ld_code(const uint8_t *p) const { return *p; }
    uint32_t ld_cg(const uint32_t *p) const { return *p; }
    uint2 ld_cg2(const uint2 *p) const { return *p; }
    uint4 ld_cg4(const uint4 *p) const { return *p; }
    void st_cg(uint32_t *p, uint32_t v) const { *p = v; }
    void st_cg2(uint2 *p, uint2 v) const { *p = v; }
    void st_cg4(uint4 *p, uint4 v) const { *p = v; }
};

typedef void (*LaneFn)(HostBackend &, void *);

struct WarpSim {
    ucontext_t sched;
    ucontext_t ctx[32];
    std::vector<char> stacks[32];
    bool done[32];
    int current;
    uint32_t xchg[32];
    uint32_t snap[32];
    int arrived;
    uint32_t gen;
    LaneFn fn;
    void *arg;

    // Fibers are created with makecontext, but every later switch goes through _setjmp / _longjmp: swapcontext saves
    // and restores the signal mask with a system call per switch, which dominated the emulation time.
    jmp_buf sched_jb;
    jmp_buf lane_jb[32];
    bool started[32];
    void yield()
    {
        if (!_setjmp(lane_jb[current])) _longjmp(sched_jb, 1);
    }
    // all 32 lanes deposit, then all continue (lock-step point of a *_sync intrinsic)
    void rendezvous(int lane, uint32_t v)
    {
        xchg[lane] = v;
        const uint32_t my = gen;
        if (++arrived == 32) {
            arrived = 0;
            memcpy(snap, xchg, sizeof snap);
            ++gen;
        }
        while (gen == my) yield();
    }
    static void trampoline(unsigned lo, unsigned hi)
    {
        WarpSim *w = reinterpret_cast<WarpSim *>(((uintptr_t)hi << 32) | (uintptr_t)lo);
        HostBackend be;
        be.w = w;
        be.lane_id = w->current;
        w->fn(be, w->arg);
        w->done[be.lane_id] = true;
        _longjmp(w->sched_jb, 1);
    }
    void run(LaneFn f, void *a)
    {
        fn = f;
        arg = a;
        arrived = 0;
        gen = 0;
        for (int l = 0; l < 32; ++l) {
            stacks[l].resize(512 * 1024);
            done[l] = false;
            started[l] = false;
            getcontext(&ctx[l]);
            ctx[l].uc_stack.ss_sp = stacks[l].data();
            ctx[l].uc_stack.ss_size = stacks[l].size();
            ctx[l].uc_link = &sched;
            const uintptr_t p = (uintptr_t)this;
            makecontext(&ctx[l], (void (*)())trampoline, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
        }
        for (;;) {
            bool alive = false;
            for (int l = 0; l < 32; ++l) {
                if (done[l]) continue;
                alive = true;
                current = l;
                if (!_setjmp(sched_jb)) {
                    if (started[l]) {
                        _longjmp(lane_jb[l], 1);
                    } else {
                        started[l] = true;
                        setcontext(&ctx[l]);
                    }
                }
            }
            if (!alive) break;
        }
    }
};

uint32_t HostBackend::shfl_up(uint32_t v, int d, int width)
{
    w->rendezvous(lane_id, v);
    const uint32_t r = (lane_id % width) >= d ? w->snap[lane_id - d] : v;
    w->rendezvous(lane_id, 0);
    return r;
}
uint32_t HostBackend::shfl_xor(uint32_t v, int m, int width)
{
    w->rendezvous(lane_id, v);
    const int src = lane_id ^ m;
    const uint32_t r = (src / width == lane_id / width) ? w->snap[src] : v;
    w->rendezvous(lane_id, 0);
    return r;
}
void HostBackend::syncwarp() { w->rendezvous(lane_id, 0); }
bool HostBackend::any(bool f)
{
    w->rendezvous(lane_id, f ? 1u : 0u);
    bool r = false;
    for (int l = 0; l < 32; ++l) r = r || w->snap[l];
    w->rendezvous(lane_id, 0);
    return r;
}
uint32_t HostBackend::next_tile(uint32_t *counter)
{
    uint32_t v = 0;
    if (lane_id == 0) v = (*counter)++;
    w->rendezvous(lane_id, v);
    const uint32_t r = w->snap[0];
    w->rendezvous(lane_id, 0);
    return r;
}

struct LaneArgs {
    const SwbScoreParams *p;
    const int8_t *sprof;
    uint32_t sstride;
    int K;
    int mode;  // 0 V16, 1 V32, 2 V16R, 3 V16A, 4 V32A
    bool split;
};

void lane_main(HostBackend &be, void *a)
{
    const LaneArgs *la = static_cast<const LaneArgs *>(a);
    if (la->split) {
        if (la->mode == 1) {
            swb_warp_loop<8, V32, true>(be, *la->p, la->sprof, la->sstride);
        } else if (la->mode == 2) {
            if (la->K == 8) swb_warp_loop<8, V16R, true>(be, *la->p, la->sprof, la->sstride);
            else if (la->K == 16) swb_warp_loop<16, V16R, true>(be, *la->p, la->sprof, la->sstride);
            else swb_warp_loop<32, V16R, true>(be, *la->p, la->sprof, la->sstride);
        } else {
            if (la->K == 8) swb_warp_loop<8, V16, true>(be, *la->p, la->sprof, la->sstride);
            else if (la->K == 16) swb_warp_loop<16, V16, true>(be, *la->p, la->sprof, la->sstride);
            else swb_warp_loop<32, V16, true>(be, *la->p, la->sprof, la->sstride);
        }
    } else if (la->mode == 2) {
        if (la->K == 8) swb_warp_loop<8, V16R, false>(be, *la->p, la->sprof, la->sstride);
        else swb_warp_loop<16, V16R, false>(be, *la->p, la->sprof, la->sstride);
    } else if (la->mode == 3) {
        if (la->K == 8) swb_warp_loop<8, V16A, false>(be, *la->p, la->sprof, la->sstride);
        else if (la->K == 16) swb_warp_loop<16, V16A, false>(be, *la->p, la->sprof, la->sstride);
        else swb_warp_loop<32, V16A, false>(be, *la->p, la->sprof, la->sstride);
    } else if (la->mode == 4) {
        swb_warp_loop<8, V32A, false>(be, *la->p, la->sprof, la->sstride);
    } else if (la->mode == 0) {
        if (la->K == 8) swb_warp_loop<8, V16, false>(be, *la->p, la->sprof, la->sstride);
        else if (la->K == 16) swb_warp_loop<16, V16, false>(be, *la->p, la->sprof, la->sstride);
        else swb_warp_loop<32, V16, false>(be, *la->p, la->sprof, la->sstride);
    } else {
        if (la->K == 8) swb_warp_loop<8, V32, false>(be, *la->p, la->sprof, la->sstride);
        else swb_warp_loop<16, V32, false>(be, *la->p, la->sprof, la->sstride);
    }
}

}  // namespace

// Mirrors swb_db_load + one job of swb_search_batch of the engine (same plan, same chunking, same kernel parameters,
// same launch groups), with the kernels replaced by the fiber emulation. K: 0 = per-group choice of the planner (the
// product default), else 8/16/32 for every group size. force_exact: 0 = s16 pass then exact recompute of flagged tiles
// (the product flow), 1 = exact pass over every tile. ovf_thr_override >= 0 replaces the s16 overflow threshold (lets
// tests force the recompute path). xl_len: lane-group tiles wider than it run as pipelined passes, 0 = never.
// split_k: rows per lane of the pipelined groups (8 / 16). exact_i32: 1 = exact passes in int32 (V32) instead of the
// rebased s16 policy (V16R). direct_len: pipelined tiles at least this wide (and a query at least this long) are
// scored by V16R at once; 0 = never. rebase_shift: log2 of the columns per V16R block, 0 = what the engine computes.
namespace {

void run_pass(SwbScoreParams &p, int mode, const SwbQueryPlan &qp, const std::vector<SwbLaunchGroup> &groups,
              const std::vector<uint8_t> &profbytes, uint32_t prof_stride)
{
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const SwbLaunchGroup &g = groups[gi];
        for (int r = 0; r < SWB_MAX_RANGES; ++r) {
            p.range_start[r] = g.range_start[r];
            p.range_cum[r] = g.range_cum[r];
        }
        std::vector<SwbQueryChunk> chunks;
        swb_group_chunks(qp, g, chunks);
        for (size_t c = 0; c < chunks.size(); ++c) {
            const SwbQueryChunk &ch = chunks[c];
            p.ntiles = g.ntiles;
            std::vector<uint32_t> prog((g.split ? swb_split_items(ch.rows, g, &p) : 0u) + 1u, 0u);
            p.prog = prog.data();
            p.row0 = ch.row0;
            p.rows = ch.rows;
            p.smem_rows = g.split ? (uint32_t)g.K * 32u : swb_group_smem_rows(ch.rows, g);
            p.first_chunk = ch.first;
            p.last_chunk = ch.last;
            uint32_t counter = 0;
            p.counter = &counter;
            p.static_wave = g.split ? 0u : 1u;  // the one emulated warp takes item 0 by position, the rest from the counter
            // stage the chunk's profile exactly like swb_score_kernel (split groups stage per work item)
            const uint32_t sstride = p.smem_rows + (g.split ? 16u : (uint32_t)SWB_BULK_LDW);
            std::vector<uint4> sprof_words(((size_t)sstride * SWB_ALPHA + 64) / 16);  // 16-byte aligned like shared memory
            int8_t *sprof = reinterpret_cast<int8_t *>(sprof_words.data());
            if (!g.split)
                for (uint32_t code = 0; code < SWB_ALPHA; ++code)
                    memcpy(sprof + (size_t)code * sstride, profbytes.data() + (size_t)code * prof_stride + ch.row0,
                           (size_t)p.smem_rows);
            LaneArgs la;
            la.p = &p;
            la.sprof = sprof;
            la.sstride = sstride;
            la.K = g.K;
            la.mode = mode;
            la.split = g.split;
            WarpSim *w = new WarpSim();
            w->run(lane_main, &la);
            delete w;
        }
    }
}

}  // namespace

// gap_extend != gap: the affine policies (V16A, then V32A on flagged tiles), as enqueue_job of the engine plans them
extern "C" int swbemu_search(const uint8_t *codes, const uint64_t *offsets, uint32_t n, uint32_t shard, uint32_t nshards,
                             uint32_t group_len, const int8_t *mat32, int gap, int gap_extend, const uint8_t *q,
                             uint32_t qlen, int K, int force_exact, uint32_t chunk_rows, int ovf_thr_override,
                             uint32_t xl_len, int split_k, int exact_i32, uint32_t direct_len, int rebase_shift,
                             int32_t *scores_out, uint32_t *recomputed_tiles)
{
    const bool affine = gap_extend != gap;
    SwbPlanOpts o;
    if (group_len) o.group_len = group_len;
    o.xl_len = xl_len;
    SwbPlan pl;
    if (swb_build_plan(offsets, n, shard, nshards ? nshards : 1, o, pl) != 0) return -1;
    const uint32_t nl = pl.n_local;
    if (recomputed_tiles) *recomputed_tiles = 0;
    if (nl == 0) return 0;
    const uint32_t rows = qlen;
    if (rows == 0 || pl.tiles.empty() || pl.max_len == 0) {
        memset(scores_out, 0, sizeof(int32_t) * nl);
        return 0;
    }
    uint32_t present = 0;
    for (int l = 0; l <= SWB_MAX_LOGG; ++l)
        if (pl.tiles_by_logg[l]) present |= 1u << l;
    // pack (same function as the device pack kernel)
    std::vector<uint64_t> residues(pl.res_bytes / 8 + 1);
    for (size_t ti = 0; ti < pl.tiles.size(); ++ti) {
        const SwbTile &t = pl.tiles[ti];
        const uint32_t P = 32u >> t.logG;
        uint64_t *out = residues.data() + t.res_off / 8;
        for (uint32_t i = 0; i < (t.width >> 2) * P; ++i)
            out[i] = swb_pack_word(t, i / P, i % P, codes, pl.seq_off.data(), pl.seq_len.data(), nl);
    }
    int max_s = 0, min_s = 0;
    for (int i = 0; i < SWB_ALPHA * SWB_ALPHA; ++i) {
        max_s = std::max<int>(max_s, mat32[i]);
        min_s = std::min<int>(min_s, mat32[i]);
    }
    if (!chunk_rows) chunk_rows = 7168;
    if (!split_k) split_k = 8;
    const int eng_shift = affine ? 0 : swb_rebase_shift(max_s, min_s, gap, (split_k > 16 ? 32u : 16u) * 32u);
    const bool r16 = !affine && eng_shift > 0 && !exact_i32;
    if (r16 && rebase_shift >= 6) {
        if (rebase_shift > eng_shift) return -3;  // a larger block than the scheme allows would not be exact
    } else {
        rebase_shift = eng_shift;
    }

    std::vector<int32_t> sorted(2 * (size_t)((nl + 1) / 2), 0);
    std::vector<uint8_t> flags(pl.tiles.size(), 0);
    std::vector<uint32_t> bnd16(2 * pl.bnd_elems + 8);
    std::vector<uint64_t> bnd32((affine ? 2 : 1) * pl.bnd_elems + 4);
    std::vector<uint64_t> blog(swb_blog_elems(pl.bnd_elems, (uint32_t)pl.tiles.size()), 0x7f7f7f7f7f7f7f7full);
    uint32_t recount = 0;
    SwbScoreParams p;
    memset(&p, 0, sizeof p);
    p.tiles = pl.tiles.data();
    p.residues = reinterpret_cast<const uint8_t *>(residues.data());
    p.flags = flags.data();
    p.recount = &recount;
    p.gap = gap;
    p.gap_open = gap;
    p.gap_extend = gap_extend;
    p.ovf_thr = ovf_thr_override >= 0 ? ovf_thr_override : 32767 - max_s;
    p.rebase_shift = (uint32_t)rebase_shift;
    p.blog = blog.data();
    std::vector<uint64_t> colstate(swb_colstate_elems(32) * 2 + 8, 0x7f7f7f7f7f7f7f7full);  // up to 16 B per element
    p.colstate = colstate.data();
    p.scores = sorted.data();

    // the split set and its direct part, as enqueue_job
    uint32_t xl[SWB_MAX_LOGG + 1] = {}, direct[SWB_MAX_LOGG + 1] = {}, rest[SWB_MAX_LOGG + 1] = {};
    uint32_t n_direct = 0, n_rest = 0;
    const bool split = !affine;
    if (split)
        for (int l = 1; l <= SWB_MAX_LOGG; ++l) {
            xl[l] = pl.xl_by_logg[l];
            if (r16 && direct_len && rows >= direct_len && !force_exact)
                while (direct[l] < xl[l] && pl.tiles[pl.tile_start_by_logg[l] + direct[l]].width >= direct_len) ++direct[l];
            rest[l] = xl[l] - direct[l];
            n_direct += direct[l];
            n_rest += rest[l];
        }
    SwbQueryPlan qp0, qp1;
    swb_plan_query(rows, K, 32, present, chunk_rows, qp0);
    swb_plan_query(rows, K, affine ? 8 : 16, present, chunk_rows, qp1);
    uint32_t prof_rows = std::max(qp0.prof_rows, qp1.prof_rows);
    if (split) prof_rows = std::max(prof_rows, swb_roundup(rows, 32u << SWB_MAX_LOGG));
    const uint32_t stride = swb_roundup(std::max(prof_rows, 16u), 16);
    std::vector<uint8_t> prof((size_t)stride * SWB_ALPHA, 0);
    for (uint32_t r = 0; r < prof_rows; ++r) {
        const uint32_t qc = r < qlen ? (q[r] & 31u) : (uint32_t)SWB_PAD;
        for (uint32_t code = 0; code < SWB_ALPHA; ++code)
            prof[(size_t)code * stride + r] = (uint8_t)(int8_t)(mat32[qc * SWB_ALPHA + code] + gap);
    }
    p.profile = reinterpret_cast<const int8_t *>(prof.data());
    p.prof_stride = stride;

    if (!force_exact) {
        // pass 0: s16 over everything but the direct tiles, V16R over those
        std::vector<SwbLaunchGroup> g0, gd;
        SwbLaunchGroup g;
        if (n_rest && swb_plan_split_group(pl, direct, rest, split_k, g)) g0.push_back(g);
        swb_plan_bulk_groups(pl, qp0, true, split ? xl : nullptr, 0x3fu, g0);
        if (n_direct && swb_plan_split_group(pl, nullptr, direct, split_k, g)) gd.push_back(g);
        p.bnd = bnd16.data();
        p.only_flagged = 0;
        run_pass(p, affine ? 3 : 0, qp0, g0, prof, stride);
        p.recount = nullptr;
        run_pass(p, 2, qp0, gd, prof, stride);
        p.recount = &recount;
    }
    {
        // exact pass: over flagged tiles, or over every tile
        std::vector<SwbLaunchGroup> g1;
        SwbLaunchGroup g;
        uint32_t none[SWB_MAX_LOGG + 1] = {};
        const uint32_t *first = force_exact ? none : direct;
        const uint32_t *cnt = force_exact ? xl : rest;
        uint32_t any = 0;
        for (int l = 1; l <= SWB_MAX_LOGG; ++l) any += cnt[l];
        if (split && any && swb_plan_split_group(pl, first, cnt, r16 ? split_k : 8, g)) g1.push_back(g);
        swb_plan_bulk_groups(pl, qp1, true, split ? xl : nullptr, 0x3fu, g1);
        if (!g1.empty() && g1[0].split)  // swb_clear_flagged_kernel
            for (size_t ti = 0; ti < pl.tiles.size(); ++ti)
                if (flags[ti] || force_exact)
                    for (uint32_t sl = 0; sl < pl.tiles[ti].npairs; ++sl)
                        sorted[2 * ((size_t)pl.tiles[ti].first_pair + sl)] = sorted[2 * ((size_t)pl.tiles[ti].first_pair + sl) + 1] = 0;
        p.bnd = r16 ? (void *)bnd16.data() : (void *)bnd32.data();
        p.only_flagged = force_exact ? 0u : 1u;
        run_pass(p, affine ? 4 : (r16 ? 2 : 1), qp1, g1, prof, stride);
    }
    for (uint32_t s = 0; s < nl; ++s) scores_out[pl.out_pos[s]] = sorted[s];
    if (recomputed_tiles) *recomputed_tiles = recount;
    return 0;
}
