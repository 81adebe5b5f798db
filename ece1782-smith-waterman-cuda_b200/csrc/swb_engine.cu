// Engine behind the C ABI (include/swb.h): owns the device-resident packed database, the per-stream
// scratch and the launch logic. One engine = one GPU (one process per GPU under torchrun).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <algorithm>
#include <string>
#include <chrono>
#include <thread>
#include <vector>

#include "../../include/swb.h"
#include "swb_internal.h"
#include "swb_kernels.h"
#include "swb_plan.h"

#define SWB_MAX_SLOTS 24
#define SWB_MAX_COUNTERS 256
#define SWB_MAX_SUB 3  // extra streams per slot: up to three distinct K values plus the split group per query
#define SWB_CHUNK_ROWS 7168u          // query rows per launch when a query does not fit shared memory
#define SWB_SMALL_SMEM_LIMIT (100u * 1024u)
#define SWB_STAGE_BYTES (32u << 20)   // pinned staging buffers for the raw database upload

static thread_local std::string g_create_error;

static double wall_ms()
{
    timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec * 1e3 + tv.tv_usec * 1e-3;
}

struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t sub[SWB_MAX_SUB] = {};
    cudaEvent_t ev_fork = nullptr;
    cudaEvent_t ev_sub[SWB_MAX_SUB] = {};
    cudaEvent_t done = nullptr;
    bool busy = false;
    bool ready = false;  // scratch sized for the loaded database
    uint8_t *h_query = nullptr;  // pinned
    uint8_t *d_query = nullptr;
    uint32_t query_cap = 0;
    int8_t *d_prof = nullptr;
    size_t prof_cap = 0;
    uint8_t *d_state = nullptr;  // [counters | flags | sorted scores], zeroed per query
    size_t state_bytes = 0, state_cap = 0;
    uint32_t *d_counters = nullptr;
    uint32_t *d_recount = nullptr;  // tiles re-scored in int32, zeroed per batch
    uint32_t *d_prog = nullptr;     // progress counters of the split (pipelined-pass) launches, zeroed per query
    size_t prog_cap = 0;
    uint8_t *d_flags = nullptr;
    int32_t *d_sorted = nullptr;
    uint32_t *d_bnd16 = nullptr;
    size_t bnd16_cap = 0;
    void *d_bnd32 = nullptr;
    size_t bnd32_cap = 0;
    void *d_blog = nullptr;  // V16R: base log of the boundary rows
    size_t blog_cap = 0;
    void *d_colstate = nullptr;  // one-lane tiles: parked row state per resident warp (swb_run_tile)
    size_t colstate_cap = 0;
    int32_t *h_scores = nullptr;  // pinned, n_local
    size_t h_scores_cap = 0;
    int32_t *pending_dst = nullptr;   // caller memory the scores of the job in flight go to (finish_slot)
    bool pending_scatter = false;     // pending_dst is a full-database vector: entry shard_ids[k] receives score k
    uint32_t *d_topk = nullptr;       // device-side selection: [k ids | k scores]
    uint32_t *h_topk = nullptr;       // pinned copy
    size_t topk_cap = 0, h_topk_cap = 0;
    uint32_t *pending_ids = nullptr;  // caller memory of the job's hit list
    int32_t *pending_top = nullptr;
    uint32_t pending_k = 0;
};

struct swb_engine {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    std::string err;
    cudaStream_t own_stream = nullptr;
    cudaStream_t user_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_fork = nullptr;
    cudaEvent_t ev_join[SWB_MAX_SLOTS] = {};
    // scoring
    int8_t h_mat[SWB_ALPHA * SWB_ALPHA];
    int gap = 2;          // linear gap penalty (affine: the gap-open penalty)
    int gap_extend = 2;   // == gap for the linear model
    bool affine = false;  // gap_extend != gap: Gotoh recurrences (V16A / V32A policies)
    int max_s = 0;
    int min_s = 0;
    bool scoring_set = false;
    int8_t *d_mat = nullptr;
    // options
    SwbPlanOpts plan_opts;
    int opt_k = 0;
    bool group_len_auto = true;  // swb_db_load picks group_len from the shard size (below) unless the option sets it
    int opt_batch_order = 0;  // batches: 0 = longest query first, 1 = in the caller's order
    int rebase_shift = 0;     // V16R block size for the current scoring scheme (0: V16R cannot run, exact passes use V32)
    int rebase_shift32 = 0;   // the same for passes of 32 rows per lane (pipelined groups with split_k = 32)
    int opt_exact = 0;        // exact passes: 0 = V16R where the scheme allows it, 1 = always V32 (int32)
    int opt_split_k = 0;      // rows per lane of the pipelined-pass groups: 0 = auto, 8, 16
    uint32_t opt_direct_len = 16000;  // pipelined tiles at least this wide against queries at least this long skip the
                              // plain s16 pass and are scored by V16R at once (their true scores pass 32767 anyway:
                              // with the reference's gap of 2, random long sequences score ~2.7 per residue, so pairs from ~12,000 residues
                              // up overflow; measured on configs[3]: 10,000 / 14,000 -> 4,521 / 4,669 GCUPS, later
                              // 14,000 / 16,000 / 18,000 / 21,000 -> 5,045 / 5,120 / 4,990 / 4,690); 0 = never
    int opt_split = -1;       // pipelined passes for the very long tiles: 1 on, 0 off, -1 auto = on for small shards
                              // (fewer tiles than twice the GPU's warp slots), where a few long tiles are the critical
                              // path of a query (+10 % at 1/8 of Swiss-Prot, +3 % at 1/4); on a large shard the bulk
                              // hides them and the extra launches cost ~0.5 % (measured in profiles/)
    int opt_group_order = 0;  // 0 auto (lone query: longest tiles first; batch: bulk first), 1 longest first, 2 bulk first
    uint32_t cur_nq = 1;
    int nslots = 16;
    int load_threads = 4;  // host threads that gather a sharded load into the staging buffers
    int opt_static_wave = -1;  // bulk launches: block-coherent first wave (swb_warp_loop): -1 auto, 0 off, 1 on
    uint32_t chunk_rows = SWB_CHUNK_ROWS;  // query rows per launch for queries beyond shared memory
    bool chunk_rows_set = false;           // false: batches on small shards use 2048-row launches (below)
    // database
    bool db_loaded = false;
    SwbPlan plan;
    int max_logg = 0;
    SwbTile *d_tiles = nullptr;
    uint8_t *d_residues = nullptr;
    uint32_t *d_out_pos = nullptr;
    uint32_t *d_shard_ids = nullptr;  // database ids of the shard in output order (sharded loads; else position == id)
    uint8_t *d_raw = nullptr;       // raw concatenated codes (input of the pack kernel)
    uint64_t *d_seq_off = nullptr;
    uint32_t *d_seq_len = nullptr;
    size_t tiles_cap = 0, residues_cap = 0, out_pos_cap = 0, raw_cap = 0, seq_off_cap = 0, seq_len_cap = 0, shard_ids_cap = 0;
    uint8_t *h_stage[2] = {nullptr, nullptr};  // pinned staging for the raw upload
    size_t stage_cap[2] = {0, 0};
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    int32_t *d_out = nullptr;
    size_t out_cap = 0;  // bytes
    uint32_t last_nq = 0;
    Slot slots[SWB_MAX_SLOTS];
    // scratch of swb_align (grow-only)
    int32_t *d_align_h = nullptr;   // rolling H diagonals of the jobs too long for shared memory
    uint8_t *d_align_dir = nullptr; // 2-bit directions of a wave of jobs
    uint8_t *d_align_q = nullptr;   // the query codes of a swb_align_batch call
    SwbAlignJob *d_align_jobs = nullptr;
    size_t align_q_cap = 0, align_jobs_cap = 0;
    uint8_t *d_align_out = nullptr;  // [5 ints header | ops]
    uint8_t *h_align_out = nullptr;  // pinned
    size_t align_h_cap = 0, align_dir_cap = 0, align_out_cap = 0, align_hout_cap = 0;
    uint32_t *h_recount = nullptr;  // pinned, SWB_MAX_SLOTS
    swb_stats_t stats;
};

static int fail(swb_engine *e, int code, const std::string &msg)
{
    if (e) e->err = msg;
    return code;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            char _b[512];                                                                            \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                                      \
            return fail(e, SWB_ERR_CUDA, _b);                                                        \
        }                                                                                            \
    } while (0)

static cudaStream_t main_stream(swb_engine *e) { return e->user_stream ? e->user_stream : e->own_stream; }

static void free_slot_db(Slot &s)
{
    if (s.d_state) cudaFree(s.d_state);
    if (s.d_bnd16) cudaFree(s.d_bnd16);
    if (s.d_bnd32) cudaFree(s.d_bnd32);
    if (s.d_blog) cudaFree(s.d_blog);
    s.d_blog = nullptr;
    if (s.d_colstate) cudaFree(s.d_colstate);
    s.d_colstate = nullptr;
    s.colstate_cap = 0;
    s.blog_cap = 0;
    if (s.d_prog) cudaFree(s.d_prog);
    s.d_prog = nullptr;
    s.prog_cap = 0;
    if (s.d_topk) cudaFree(s.d_topk);
    if (s.h_topk) cudaFreeHost(s.h_topk);
    s.d_topk = nullptr;
    s.h_topk = nullptr;
    s.topk_cap = s.h_topk_cap = 0;
    s.busy = false;
    s.pending_dst = nullptr;
    s.pending_ids = nullptr;
    s.pending_top = nullptr;
    if (s.h_scores) cudaFreeHost(s.h_scores);
    s.d_state = nullptr;
    s.d_bnd16 = nullptr;
    s.d_bnd32 = nullptr;
    s.h_scores = nullptr;
    s.state_bytes = s.state_cap = s.bnd16_cap = s.bnd32_cap = s.h_scores_cap = 0;
}

static void free_db(swb_engine *e)
{
    void *dev[] = {e->d_tiles, e->d_residues, e->d_out_pos, e->d_out, e->d_raw, e->d_seq_off, e->d_seq_len, e->d_shard_ids};
    for (void *p : dev)
        if (p) cudaFree(p);
    e->d_tiles = nullptr;
    e->d_residues = nullptr;
    e->d_out_pos = nullptr;
    e->d_out = nullptr;
    e->d_raw = nullptr;
    e->d_seq_off = nullptr;
    e->d_seq_len = nullptr;
    e->d_shard_ids = nullptr;
    e->tiles_cap = e->residues_cap = e->out_pos_cap = e->raw_cap = e->seq_off_cap = e->seq_len_cap = e->shard_ids_cap = 0;
    e->out_cap = 0;
    if (e->d_align_h) cudaFree(e->d_align_h);
    if (e->d_align_q) cudaFree(e->d_align_q);
    if (e->d_align_jobs) cudaFree(e->d_align_jobs);
    e->d_align_q = nullptr;
    e->d_align_jobs = nullptr;
    e->align_q_cap = e->align_jobs_cap = 0;
    if (e->d_align_dir) cudaFree(e->d_align_dir);
    if (e->d_align_out) cudaFree(e->d_align_out);
    if (e->h_align_out) cudaFreeHost(e->h_align_out);
    e->d_align_h = nullptr;
    e->d_align_dir = nullptr;
    e->d_align_out = nullptr;
    e->h_align_out = nullptr;
    e->align_h_cap = e->align_dir_cap = e->align_out_cap = e->align_hout_cap = 0;
    for (int i = 0; i < 2; ++i) {
        if (e->h_stage[i]) cudaFreeHost(e->h_stage[i]);
        if (e->ev_stage[i]) cudaEventDestroy(e->ev_stage[i]);
        e->h_stage[i] = nullptr;
        e->ev_stage[i] = nullptr;
        e->stage_cap[i] = 0;
    }
    for (int i = 0; i < SWB_MAX_SLOTS; ++i) free_slot_db(e->slots[i]);
    e->db_loaded = false;
    e->last_nq = 0;
}

extern "C" int swb_create(swb_engine **out, int device)
{
    if (!out) return SWB_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(ce) +
                         " (this library has no CPU fallback)";
        return SWB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "device index out of range";
        return SWB_ERR_ARG;
    }
    swb_engine *e = new swb_engine();
    e->device = device;
    memset(&e->stats, 0, sizeof e->stats);
    // every failure path releases what was created so far (swb_destroy copes with a half-built engine)
    auto bail = [&](const char *what, cudaError_t err) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        swb_destroy(e);
        return SWB_ERR_CUDA;
    };
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", ce);
    cudaDeviceProp prop;
    if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", ce);
    if (prop.major < 10) {
        g_create_error = std::string("device ") + prop.name + " is not sm_100-class; this library targets B200 only";
        swb_destroy(e);
        return SWB_ERR_CUDA;
    }
    e->sm_count = prop.multiProcessorCount;
    e->smem_optin = prop.sharedMemPerBlockOptin;
    auto quiet_event = [&](cudaEvent_t *ev) { return cudaEventCreateWithFlags(ev, cudaEventDisableTiming); };
    if ((ce = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail("cudaStreamCreate", ce);
    if ((ce = cudaEventCreate(&e->ev_start)) != cudaSuccess) return bail("cudaEventCreate", ce);
    if ((ce = cudaEventCreate(&e->ev_stop)) != cudaSuccess) return bail("cudaEventCreate", ce);
    if ((ce = quiet_event(&e->ev_fork)) != cudaSuccess) return bail("cudaEventCreate", ce);
    for (int i = 0; i < SWB_MAX_SLOTS; ++i) {
        Slot &sl = e->slots[i];
        if ((ce = cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking)) != cudaSuccess)
            return bail("cudaStreamCreate", ce);
        if ((ce = quiet_event(&sl.done)) != cudaSuccess) return bail("cudaEventCreate", ce);
        if ((ce = quiet_event(&sl.ev_fork)) != cudaSuccess) return bail("cudaEventCreate", ce);
        for (int k = 0; k < SWB_MAX_SUB; ++k) {
            if ((ce = cudaStreamCreateWithFlags(&sl.sub[k], cudaStreamNonBlocking)) != cudaSuccess)
                return bail("cudaStreamCreate", ce);
            if ((ce = quiet_event(&sl.ev_sub[k])) != cudaSuccess) return bail("cudaEventCreate", ce);
        }
        if ((ce = quiet_event(&e->ev_join[i])) != cudaSuccess) return bail("cudaEventCreate", ce);
        if ((ce = cudaMalloc(&sl.d_recount, sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc", ce);
    }
    if ((ce = cudaMalloc(&e->d_mat, SWB_ALPHA * SWB_ALPHA)) != cudaSuccess) return bail("cudaMalloc", ce);
    if ((ce = cudaMallocHost(&e->h_recount, SWB_MAX_SLOTS * sizeof(uint32_t))) != cudaSuccess)
        return bail("cudaMallocHost", ce);
    e->stats.sm_count = (uint32_t)e->sm_count;
    *out = e;
    int rc = swb_set_scoring_preset(e, SWB_SCORING_BLOSUM50_REF);
    if (rc != SWB_OK) {
        g_create_error = e->err;
        swb_destroy(e);
        *out = nullptr;
        return rc;
    }
    return SWB_OK;
}

extern "C" void swb_destroy(swb_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    free_db(e);
    for (int i = 0; i < SWB_MAX_SLOTS; ++i) {
        Slot &s = e->slots[i];
        if (s.h_query) cudaFreeHost(s.h_query);
        if (s.d_query) cudaFree(s.d_query);
        if (s.d_prof) cudaFree(s.d_prof);
        if (s.d_recount) cudaFree(s.d_recount);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.ev_fork) cudaEventDestroy(s.ev_fork);
        for (int k = 0; k < SWB_MAX_SUB; ++k) {
            if (s.sub[k]) cudaStreamDestroy(s.sub[k]);
            if (s.ev_sub[k]) cudaEventDestroy(s.ev_sub[k]);
        }
        if (s.done) cudaEventDestroy(s.done);
        if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
    }
    if (e->d_mat) cudaFree(e->d_mat);
    if (e->h_recount) cudaFreeHost(e->h_recount);
    if (e->ev_start) cudaEventDestroy(e->ev_start);
    if (e->ev_stop) cudaEventDestroy(e->ev_stop);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

extern "C" const char *swb_last_error(const swb_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

extern "C" int swb_set_option(swb_engine *e, const char *key, int64_t value)
{
    if (!e || !key) return SWB_ERR_ARG;
    if (!strcmp(key, "group_len")) {
        if (value != 0 && (value < 8 || value > (1 << 30))) return fail(e, SWB_ERR_ARG, "group_len out of range");
        e->group_len_auto = value == 0;
        if (value) e->plan_opts.group_len = (uint32_t)value;
    } else if (!strcmp(key, "k")) {
        if (value != 0 && value != 8 && value != 16 && value != 32) return fail(e, SWB_ERR_ARG, "k must be 0, 8, 16 or 32");
        e->opt_k = (int)value;
    } else if (!strcmp(key, "streams")) {
        if (value < 1 || value > SWB_MAX_SLOTS) return fail(e, SWB_ERR_ARG, "streams must be 1..24");
        e->nslots = (int)value;
    } else if (!strcmp(key, "xl_len")) {
        if (value < 0 || value > (1ll << 31)) return fail(e, SWB_ERR_ARG, "xl_len out of range");
        e->plan_opts.xl_len = (uint32_t)value;
    } else if (!strcmp(key, "batch_order")) {
        if (value < 0 || value > 1) return fail(e, SWB_ERR_ARG, "batch_order must be 0 (longest query first) or 1 (as given)");
        e->opt_batch_order = (int)value;
    } else if (!strcmp(key, "split")) {
        if (value < -1 || value > 1) return fail(e, SWB_ERR_ARG, "split must be -1 (auto), 0 or 1");
        e->opt_split = (int)value;
    } else if (!strcmp(key, "group_order")) {
        if (value < 0 || value > 2) return fail(e, SWB_ERR_ARG, "group_order must be 0, 1 or 2");
        e->opt_group_order = (int)value;
    } else if (!strcmp(key, "exact")) {
        if (value < 0 || value > 1) return fail(e, SWB_ERR_ARG, "exact must be 0 (rebased s16 where possible) or 1 (int32)");
        e->opt_exact = (int)value;
    } else if (!strcmp(key, "split_k")) {
        if (value != 0 && value != 8 && value != 16 && value != 32) return fail(e, SWB_ERR_ARG, "split_k must be 0, 8, 16 or 32");
        e->opt_split_k = (int)value;
    } else if (!strcmp(key, "direct_len")) {
        if (value < 0 || value > (1ll << 31)) return fail(e, SWB_ERR_ARG, "direct_len out of range");
        e->opt_direct_len = (uint32_t)value;
    } else if (!strcmp(key, "static_wave")) {
        e->opt_static_wave = value < 0 ? -1 : (value != 0);
    } else if (!strcmp(key, "load_threads")) {
        if (value < 1 || value > 64) return fail(e, SWB_ERR_ARG, "load_threads must be 1..64");
        e->load_threads = (int)value;
    } else if (!strcmp(key, "chunk_rows")) {
        if (value < 1024 || value > SWB_CHUNK_ROWS || value % 1024) return fail(e, SWB_ERR_ARG, "chunk_rows must be a multiple of 1024 up to 7168");
        e->chunk_rows = (uint32_t)value;
        e->chunk_rows_set = true;
    } else {
        return fail(e, SWB_ERR_ARG, std::string("unknown option ") + key);
    }
    return SWB_OK;
}

extern "C" int swb_set_stream(swb_engine *e, void *cuda_stream)
{
    if (!e) return SWB_ERR_ARG;
    e->user_stream = (cudaStream_t)cuda_stream;
    return SWB_OK;
}

extern "C" int swb_set_scoring(swb_engine *e, const int8_t *matrix, int alpha, int gap)
{
    return swb_set_scoring_affine(e, matrix, alpha, gap, gap);
}

// gap_open: penalty of the first residue of a gap, gap_extend: of every further one (a gap of length L costs
// gap_open + (L-1) * gap_extend). gap_open == gap_extend is the reference's linear model and runs the linear kernels.
extern "C" int swb_set_scoring_affine(swb_engine *e, const int8_t *matrix, int alpha, int gap, int gap_extend)
{
    if (!e || !matrix) return SWB_ERR_ARG;
    if (alpha < 1 || alpha > SWB_ALPHA) return fail(e, SWB_ERR_ARG, "alpha must be 1..32");
    if (gap < 0 || gap > 64) return fail(e, SWB_ERR_ARG, "gap must be 0..64");
    if (gap_extend < 0 || gap_extend > gap) return fail(e, SWB_ERR_ARG, "gap_extend must be 0..gap_open");
    int8_t m[SWB_ALPHA * SWB_ALPHA];
    memset(m, 0, sizeof m);
    int mx = 0, mn = 0;
    for (int i = 0; i < alpha; ++i)
        for (int j = 0; j < alpha; ++j) {
            int v = matrix[i * alpha + j];
            if (i == SWB_PAD || j == SWB_PAD) v = 0;  // the padding code is score-neutral by construction
            if (v + gap > 127 || v + gap < -128) return fail(e, SWB_ERR_ARG, "matrix entry + gap does not fit int8");
            m[i * SWB_ALPHA + j] = (int8_t)v;
            mx = std::max(mx, v);
            mn = std::min(mn, v);
        }
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(main_stream(e)));
    memcpy(e->h_mat, m, sizeof m);
    e->gap = gap;
    e->gap_extend = gap_extend;
    e->affine = gap_extend != gap;
    e->max_s = mx;
    e->min_s = mn;
    e->rebase_shift = e->affine ? 0 : swb_rebase_shift(mx, mn, gap, 16u * 32u);
    e->rebase_shift32 = e->affine ? 0 : swb_rebase_shift(mx, mn, gap, 32u * 32u);
    CU(cudaMemcpy(e->d_mat, e->h_mat, sizeof m, cudaMemcpyHostToDevice));
    e->scoring_set = true;
    return SWB_OK;
}

extern "C" int swb_set_scoring_preset(swb_engine *e, int preset)
{
    if (!e) return SWB_ERR_ARG;
    int8_t m[SWB_ALPHA * SWB_ALPHA];
    int gap = 0;
    if (swb_scoring_matrix(preset, m, &gap) != SWB_OK) return fail(e, SWB_ERR_ARG, "unknown scoring preset");
    return swb_set_scoring(e, m, SWB_ALPHA, gap);
}

// ---------------------------------------------------------------------------------------------
// grow-only device / pinned buffers: a reload of a database of similar size reuses every allocation
static cudaError_t grow_dev(void **p, size_t *cap, size_t need)
{
    if (need <= *cap && *p) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    need = std::max<size_t>(need + need / 16, 256);
    cudaError_t e = cudaMalloc(p, need);
    if (e == cudaSuccess) *cap = need;
    return e;
}
static cudaError_t grow_host(void **p, size_t *cap, size_t need)
{
    if (need <= *cap && *p) return cudaSuccess;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    need = std::max<size_t>(need + need / 16, 256);
    cudaError_t e = cudaMallocHost(p, need);
    if (e == cudaSuccess) *cap = need;
    return e;
}
#define GROW_DEV(ptr, cap, need) grow_dev(reinterpret_cast<void **>(&(ptr)), &(cap), (need))
#define GROW_HOST(ptr, cap, need) grow_host(reinterpret_cast<void **>(&(ptr)), &(cap), (need))

// streams `bytes` of host memory to d_raw through the two pinned staging buffers: the host copy of slice i+1 overlaps the
// asynchronous H2D of slice i
static cudaError_t upload_staged(swb_engine *e, const uint8_t *src, uint64_t bytes, cudaStream_t st)
{
    const size_t stage = (size_t)std::min<uint64_t>(SWB_STAGE_BYTES, bytes);
    cudaError_t ce;
    for (int i = 0; i < 2; ++i) {
        if ((ce = GROW_HOST(e->h_stage[i], e->stage_cap[i], stage)) != cudaSuccess) return ce;
        if (!e->ev_stage[i] && (ce = cudaEventCreateWithFlags(&e->ev_stage[i], cudaEventDisableTiming)) != cudaSuccess)
            return ce;
    }
    int b = 0;
    for (uint64_t off = 0; off < bytes; off += stage, b ^= 1) {
        const size_t len = (size_t)std::min<uint64_t>(stage, bytes - off);
        if ((ce = cudaEventSynchronize(e->ev_stage[b])) != cudaSuccess) return ce;
        memcpy(e->h_stage[b], src + off, len);
        if ((ce = cudaMemcpyAsync(e->d_raw + off, e->h_stage[b], len, cudaMemcpyHostToDevice, st)) != cudaSuccess) return ce;
        if ((ce = cudaEventRecord(e->ev_stage[b], st)) != cudaSuccess) return ce;
    }
    return cudaSuccess;
}

// Fills dst with bytes [lo, hi) of the shard's residues in sorted order (sequence s sits at goff[s] of that stream): the
// gather of a sharded load, cut into equal byte ranges for `threads` host threads.
static void gather_range(const SwbPlan &pl, const std::vector<uint64_t> &goff, const uint8_t *codes, uint64_t lo,
                         uint64_t hi, uint8_t *dst, int threads)
{
    auto work = [&](uint64_t a, uint64_t b) {
        if (a >= b) return;
        size_t sq = (size_t)(std::upper_bound(goff.begin(), goff.end(), a) - goff.begin()) - 1;
        while (a < b) {
            const uint64_t within = a - goff[sq];
            const uint64_t take = std::min<uint64_t>(pl.seq_len[sq] - within, b - a);
            memcpy(dst + (a - lo), codes + pl.seq_off[sq] + within, (size_t)take);
            a += take;
            ++sq;
        }
    };
    if (threads <= 1 || hi - lo < (1u << 20)) return work(lo, hi);
    std::vector<std::thread> pool;
    const uint64_t step = (hi - lo + (uint64_t)threads - 1) / (uint64_t)threads;
    for (int t = 1; t < threads; ++t)
        pool.emplace_back(work, std::min(hi, lo + step * (uint64_t)t), std::min(hi, lo + step * (uint64_t)(t + 1)));
    work(lo, std::min(hi, lo + step));
    for (std::thread &th : pool) th.join();
}

// sorted_order: the ids of the WHOLE database by descending length (swb_sort_by_length), or NULL to sort here. An engine
// group sorts once and hands the same order to all of its engines.
int swb_db_load_sorted(swb_engine *e, const uint8_t *codes, const uint64_t *offsets, uint32_t n, uint32_t shard,
                       uint32_t nshards, const uint32_t *sorted_order)
{
    if (!e || !offsets || (!codes && n && offsets[n] != offsets[0])) return SWB_ERR_ARG;
    if (nshards == 0) nshards = 1;
    if (shard >= nshards) return fail(e, SWB_ERR_ARG, "shard >= nshards");
    const bool timing = getenv("SWB_TIMING") != nullptr;
    const double t0 = wall_ms();
    CU(cudaSetDevice(e->device));
    cudaStream_t st = main_stream(e);
    CU(cudaStreamSynchronize(st));
    for (int i = 0; i < SWB_MAX_SLOTS; ++i) CU(cudaStreamSynchronize(e->slots[i].stream));
    e->db_loaded = false;
    e->last_nq = 0;  // results of the previous database are gone (swb_fetch_scores must not index the new shard with them)
    // group_len (longest sequence that runs one lane per pair). One-lane tiles are the cheapest per cell (no shuffles,
    // conflict-free profile reads, least row padding), so a shard with plenty of tiles wants them for as many sequences
    // as possible; a small shard is the opposite case: with about as many tiles as warp slots, long one-lane tiles
    // become its critical path. Measured on fractions of the benchmark database, 20 reference queries, GCUPS for
    // group_len 384 / 768 / 1536 (profiles/r2a_sweep_group_len.txt):
    //   570 k sequences 8,868 / 9,115 / 9,153     285 k  8,834 / 8,960 / 8,882
    //   142 k           8,709 / 8,792 / 8,358      71 k  8,384 / 7,616 / 5,403
    if (e->group_len_auto) {
        const uint64_t slots = (uint64_t)e->sm_count * (SWB_NT_LARGE / 32);
        const uint64_t tiles64 = (uint64_t)(n / nshards) / 64u;  // one-lane tiles the shard would have
        e->plan_opts.group_len = tiles64 >= 3u * slots ? 1536u : (4u * tiles64 >= 3u * slots ? 768u : 384u);
    }
    // An unsharded load uploads the caller's buffer as it is, which does not need the plan: the plan (length sort, tiling;
    // ~20 ms for Swiss-Prot) is built on a helper thread while this one stages and uploads the residues.
    bool early_upload = nshards == 1 && n > 0 && offsets[n] > offsets[0];
    for (uint32_t i = 0; i < n && early_upload; ++i) early_upload = offsets[i + 1] >= offsets[i];  // else: fails below
    int rc = 0;
    std::thread planner;
    if (early_upload) {
        planner = std::thread(
            [&]() { rc = swb_build_plan(offsets, n, shard, nshards, e->plan_opts, e->plan, sorted_order); });
        const uint64_t bytes = offsets[n] - offsets[0];
        int urc = SWB_OK;
        cudaError_t ce = GROW_DEV(e->d_raw, e->raw_cap, bytes);
        if (ce == cudaSuccess) ce = upload_staged(e, codes + offsets[0], bytes, st);
        if (ce != cudaSuccess) urc = fail(e, SWB_ERR_CUDA, std::string("db upload: ") + cudaGetErrorString(ce));
        planner.join();
        if (urc != SWB_OK) return urc;
    } else {
        // A sharded load needs the head of the plan (the shard's sequences in sorted order) before it can gather; the
        // tail (output order, tiles) is built on the helper thread beside the gather and the upload.
        rc = swb_build_plan_head(offsets, n, shard, nshards, e->plan, sorted_order);
        if (rc == 0) {
            if (nshards > 1 && e->plan.n_local > 0)
                planner = std::thread([&]() { swb_build_plan_tail(e->plan_opts, e->plan); });
            else
                swb_build_plan_tail(e->plan_opts, e->plan);
        }
    }
    struct Joiner {  // every return below joins the helper first
        std::thread &t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } joiner{planner};
    if (rc != 0) return fail(e, SWB_ERR_ARG, "bad offsets (decreasing, or a sequence longer than 2^31-16)");
    const double t0b = wall_ms();
    // a sharded load uploads only the shard's residues (gathered in sorted order), a full load the caller's buffer
    const bool gather = nshards > 1;
    std::vector<uint64_t> goff;
    if (gather && e->plan.n_local > 0) {
        // positions in the gathered stream, then the stream itself through the two pinned staging buffers: the host
        // gather of slice i+1 (a few threads, byte-balanced) overlaps the asynchronous H2D of slice i. Reads seq_len and
        // seq_off of the plan, which the tail does not touch.
        const SwbPlan &ph = e->plan;
        const uint32_t nl0 = ph.n_local;
        const uint64_t raw_bytes0 = ph.residues_local;
        goff.resize((size_t)nl0 + 1);
        goff[0] = 0;
        for (uint32_t s = 0; s < nl0; ++s) goff[s + 1] = goff[s] + ph.seq_len[s];
        CU(GROW_DEV(e->d_raw, e->raw_cap, raw_bytes0));
        if (raw_bytes0) {
            const size_t stage = (size_t)std::min<uint64_t>(SWB_STAGE_BYTES, raw_bytes0);
            for (int i = 0; i < 2; ++i) {
                CU(GROW_HOST(e->h_stage[i], e->stage_cap[i], stage));
                if (!e->ev_stage[i]) CU(cudaEventCreateWithFlags(&e->ev_stage[i], cudaEventDisableTiming));
            }
            int b = 0;
            for (uint64_t done = 0; done < raw_bytes0; done += stage, b ^= 1) {
                const uint64_t len = std::min<uint64_t>(stage, raw_bytes0 - done);
                CU(cudaEventSynchronize(e->ev_stage[b]));
                gather_range(ph, goff, codes, done, done + len, e->h_stage[b], e->load_threads);
                CU(cudaMemcpyAsync(e->d_raw + done, e->h_stage[b], (size_t)len, cudaMemcpyHostToDevice, st));
                CU(cudaEventRecord(e->ev_stage[b], st));
            }
        }
    }
    if (planner.joinable()) planner.join();
    const double t1 = wall_ms();
    SwbPlan &pl = e->plan;
    e->max_logg = 0;
    for (int l = 0; l <= SWB_MAX_LOGG; ++l)
        if (pl.tiles_by_logg[l]) e->max_logg = l;
    const uint32_t nl = pl.n_local;
    const uint32_t ntiles = (uint32_t)pl.tiles.size();
    const uint64_t base = n ? offsets[0] : 0;
    const uint64_t raw_bytes = gather ? pl.residues_local : pl.residues_total;
    double t2 = t1, t3 = t1;

    if (nl > 0 && ntiles > 0) {
        CU(GROW_DEV(e->d_raw, e->raw_cap, raw_bytes));
        CU(GROW_DEV(e->d_seq_off, e->seq_off_cap, sizeof(uint64_t) * nl));
        CU(GROW_DEV(e->d_seq_len, e->seq_len_cap, sizeof(uint32_t) * nl));
        CU(GROW_DEV(e->d_tiles, e->tiles_cap, sizeof(SwbTile) * ntiles));
        CU(GROW_DEV(e->d_residues, e->residues_cap, pl.res_bytes));
        CU(GROW_DEV(e->d_out_pos, e->out_pos_cap, sizeof(uint32_t) * nl));
        if (gather) CU(GROW_DEV(e->d_shard_ids, e->shard_ids_cap, sizeof(uint32_t) * nl));
        for (int i = 0; i < SWB_MAX_SLOTS; ++i) e->slots[i].ready = false;  // per-stream scratch is (re)sized on first use
        t2 = wall_ms();
        if (gather) {
            for (uint32_t s = 0; s < nl; ++s) pl.seq_off[s] = goff[s];  // offsets into the uploaded buffer
            CU(cudaMemcpyAsync(e->d_shard_ids, pl.shard_ids.data(), sizeof(uint32_t) * nl, cudaMemcpyHostToDevice, st));
        } else {
            for (uint32_t s = 0; s < nl; ++s) pl.seq_off[s] -= base;
        }
        CU(cudaMemcpyAsync(e->d_seq_off, pl.seq_off.data(), sizeof(uint64_t) * nl, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(e->d_seq_len, pl.seq_len.data(), sizeof(uint32_t) * nl, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(e->d_tiles, pl.tiles.data(), sizeof(SwbTile) * ntiles, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(e->d_out_pos, pl.out_pos.data(), sizeof(uint32_t) * nl, cudaMemcpyHostToDevice, st));
        t3 = wall_ms();
        // the one HBM-bound kernel of the path, timed on the device (ev_start / ev_stop are free: no search is in flight)
        CU(cudaEventRecord(e->ev_start, st));
        CU(swb_launch_pack(e->d_tiles, ntiles, e->d_raw, e->d_seq_off, e->d_seq_len, nl, e->d_residues, st));
        CU(cudaEventRecord(e->ev_stop, st));
        CU(cudaStreamSynchronize(st));
        float pack_ms = 0.f;
        CU(cudaEventElapsedTime(&pack_ms, e->ev_start, e->ev_stop));
        e->stats.pack_us = (uint32_t)(pack_ms * 1000.f + 0.5f);
    }
    e->db_loaded = true;
    e->stats.load_ms = wall_ms() - t0;
    if (timing)
        fprintf(stderr, "[swb] db_load: plan (unsharded: beside the upload; sharded: its head) %.1f ms, sharded gather + h2d enqueue "
                        "beside the plan's tail %.1f ms, alloc %.1f ms, tables h2d enqueue %.1f ms, pack+sync %.1f ms, total %.1f ms\n",
                t0b - t0, t1 - t0b, t2 - t1, t3 - t2, wall_ms() - t3, e->stats.load_ms);
    e->stats.db_residues = pl.residues_local;
    e->stats.db_residues_total = pl.residues_total;
    e->stats.db_sequences = nl;
    e->stats.tiles = ntiles;
    for (int l = 0; l <= SWB_MAX_LOGG; ++l) e->stats.tiles_by_group[l] = pl.tiles_by_logg[l];
    return SWB_OK;
}

extern "C" int swb_db_load(swb_engine *e, const uint8_t *codes, const uint64_t *offsets, uint32_t n, uint32_t shard,
                           uint32_t nshards)
{
    return swb_db_load_sorted(e, codes, offsets, n, shard, nshards, nullptr);
}

extern "C" int swb_pack_time(swb_engine *e, int reps, double *us_per_launch, uint64_t *bytes_per_launch)
{
    if (!e || !us_per_launch || reps < 1 || reps > 1000) return SWB_ERR_ARG;
    if (!e->db_loaded) return fail(e, SWB_ERR_STATE, "swb_pack_time before swb_db_load");
    const SwbPlan &pl = e->plan;
    *us_per_launch = 0.0;
    if (bytes_per_launch) *bytes_per_launch = pl.residues_local + pl.res_bytes;
    if (pl.n_local == 0 || pl.tiles.empty()) return SWB_OK;
    CU(cudaSetDevice(e->device));
    cudaStream_t st = main_stream(e);
    CU(cudaStreamSynchronize(st));
    for (int i = 0; i < SWB_MAX_SLOTS; ++i) CU(cudaStreamSynchronize(e->slots[i].stream));
    for (int r = 0; r <= reps; ++r) {  // launch 0 is the warm-up
        if (r == 1) CU(cudaEventRecord(e->ev_start, st));
        CU(swb_launch_pack(e->d_tiles, (uint32_t)pl.tiles.size(), e->d_raw, e->d_seq_off, e->d_seq_len, pl.n_local,
                           e->d_residues, st));
    }
    CU(cudaEventRecord(e->ev_stop, st));
    CU(cudaStreamSynchronize(st));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e->ev_start, e->ev_stop));
    *us_per_launch = (double)ms * 1000.0 / reps;
    return SWB_OK;
}

extern "C" uint32_t swb_db_count(const swb_engine *e) { return (e && e->db_loaded) ? e->plan.n_local : 0; }

extern "C" int swb_db_ids(const swb_engine *e, uint32_t *ids)
{
    if (!e || !ids) return SWB_ERR_ARG;
    if (!e->db_loaded) return SWB_ERR_STATE;
    memcpy(ids, e->plan.shard_ids.data(), sizeof(uint32_t) * e->plan.n_local);
    return SWB_OK;
}

// ---------------------------------------------------------------------------------------------
struct LaunchShape {
    int block_cfg;
    int grid;
    uint32_t warps_active;  // 0 = every warp of a block takes tiles
    size_t smem;
    uint32_t smem_rows;
};

static int shape_for(swb_engine *e, int K, int mode, bool split, uint32_t smem_rows, uint32_t ntiles, LaunchShape &ls)
{
    const bool per_item = split;  // one warp per block, each work item stages the rows of its pass
    if (per_item) smem_rows = (uint32_t)K * 32u;
    ls.smem_rows = smem_rows;
    ls.smem = (size_t)SWB_ALPHA * (smem_rows + (split ? 16 : SWB_BULK_LDW));  // row stride as in swb_score_kernel
    if (ls.smem > e->smem_optin) return fail(e, SWB_ERR_ARG, "internal: query chunk does not fit shared memory");
    ls.block_cfg = ls.smem <= SWB_SMALL_SMEM_LIMIT ? SWB_BLOCK_SMALL : SWB_BLOCK_LARGE;
    int per_sm = 0;
    CU(swb_score_occupancy(K, mode, split, ls.block_cfg, ls.smem, &per_sm));
    if (per_sm < 1) return fail(e, SWB_ERR_CUDA, "score kernel does not fit on an SM");
    // Pipelined passes run one warp per block, and the passes of a tile advance at the pace of the slowest of them: an
    // SM with 13 such warps has one scheduler with four and three with three, and every chain that has a pass on the
    // crowded scheduler drags its other passes down with it (ncu: 40 % of the warp time spent waiting for the
    // predecessor). Keep the schedulers evenly loaded: a multiple of four blocks per SM.
    if (per_item && per_sm > 4 && !getenv("SWB_SPLIT_UNEVEN")) per_sm -= per_sm % 4;
    const int nt = per_item ? 32 : (ls.block_cfg == SWB_BLOCK_SMALL ? SWB_NT_SMALL : SWB_NT_LARGE);
    const int need = (int)((ntiles + nt / 32 - 1) / (nt / 32));
    const int slots = per_sm * e->sm_count;
    ls.grid = std::max(1, std::min(slots, need));
    ls.warps_active = 0;
    // A lone query whose launch has fewer tiles than the GPU has warps: one tile per warp packed into full blocks would
    // put 16 long-running warps on a few SMs and leave the others empty. Spread them: more blocks, fewer active warps
    // each. (Not for batches: there other queries fill the SMs, and half-empty blocks would only hold shared memory.)
    if (!per_item && e->cur_nq <= 1 && need < slots) {
        const uint32_t wpb = std::max<uint32_t>(1u, (ntiles + (uint32_t)slots - 1u) / (uint32_t)slots);
        if (wpb < (uint32_t)(nt / 32)) {
            ls.warps_active = wpb;
            ls.grid = std::max(1, (int)((ntiles + wpb - 1u) / wpb));
        }
    }
    return SWB_OK;
}

// Per-stream scratch of the loaded database, sized on the slot's first use after a load (grow-only buffers): a lone
// query touches one slot, a batch as many as it has jobs in flight. State layout: [launch counters | tile flags |
// scores in sorted order].
static int ensure_slot(swb_engine *e, Slot &s)
{
    if (s.ready) return SWB_OK;
    const SwbPlan &pl = e->plan;
    const uint32_t nl = pl.n_local;
    const size_t flags_bytes = swb_roundup((uint32_t)pl.tiles.size(), 16);
    const size_t sorted_bytes = sizeof(int32_t) * 2 * (size_t)((nl + 1) / 2);
    const size_t head = sizeof(uint32_t) * SWB_MAX_COUNTERS;
    s.state_bytes = head + flags_bytes + sorted_bytes;
    CU(GROW_DEV(s.d_state, s.state_cap, s.state_bytes));
    s.d_counters = reinterpret_cast<uint32_t *>(s.d_state);
    s.d_flags = s.d_state + head;
    s.d_sorted = reinterpret_cast<int32_t *>(s.d_state + head + flags_bytes);
    CU(GROW_DEV(s.d_bnd16, s.bnd16_cap, sizeof(uint32_t) * pl.bnd_elems));
    CU(GROW_HOST(s.h_scores, s.h_scores_cap, sizeof(int32_t) * (size_t)nl));
    s.ready = true;
    return SWB_OK;
}

// the launches of one pass (one arithmetic policy over one profile): every launch group on its own stream. `extra`
// (may be NULL) are groups of the V16R policy that run beside this pass on tiles of their own (the direct set).
static int enqueue_pass(swb_engine *e, Slot &s, int mode, SwbScoreParams &p, const SwbQueryPlan &qp,
                        const std::vector<SwbLaunchGroup> &groups, uint32_t &counter, uint32_t *prog,
                        const std::vector<SwbLaunchGroup> *extra, uint32_t *prog_extra)
{
    const size_t nmain = groups.size();
    const size_t ng = nmain + (extra ? extra->size() : 0);
    if (ng > 1 + SWB_MAX_SUB) return fail(e, SWB_ERR_ARG, "internal: too many launch groups");
    // fork: the launch groups of a pass are independent of each other (disjoint tiles)
    if (ng > 1) {
        CU(cudaEventRecord(s.ev_fork, s.stream));
        for (size_t gi = 1; gi < ng; ++gi) CU(cudaStreamWaitEvent(s.sub[gi - 1], s.ev_fork, 0));
    }
    for (size_t gi = 0; gi < ng; ++gi) {
        const bool is_extra = gi >= nmain;
        const SwbLaunchGroup &g = is_extra ? (*extra)[gi - nmain] : groups[gi];
        const int gmode = is_extra ? SWB_MODE_R16 : mode;
        cudaStream_t st = gi == 0 ? s.stream : s.sub[gi - 1];
        for (int r = 0; r < SWB_MAX_RANGES; ++r) {
            p.range_start[r] = g.range_start[r];
            p.range_cum[r] = g.range_cum[r];
        }
        size_t prog_at = 0;
        std::vector<SwbQueryChunk> chunks;
        swb_group_chunks(qp, g, chunks);
        uint32_t *const recount = p.recount;
        if (is_extra) p.recount = nullptr;  // tiles scored by V16R at once are no recomputes
        for (size_t c = 0; c < chunks.size(); ++c) {
            const SwbQueryChunk &ch = chunks[c];
            // work items of the launch: tiles, or (tile, pass) for split groups
            p.ntiles = g.ntiles;
            p.prog = g.split ? (is_extra ? prog_extra : prog) + prog_at : nullptr;
            if (g.split) prog_at += swb_split_items(ch.rows, g, &p);  // sets p.ntiles to the number of items
            LaunchShape ls;
            int rc = shape_for(e, g.K, gmode, g.split, swb_group_smem_rows(ch.rows, g), p.ntiles, ls);
            if (rc != SWB_OK) return rc;
            p.row0 = ch.row0;
            p.rows = ch.rows;
            p.rebase_shift = (uint32_t)(g.K > 16 ? e->rebase_shift32 : e->rebase_shift);
            p.smem_rows = ls.smem_rows;
            p.warps_active = ls.warps_active;
            {
                const uint32_t nt = ls.block_cfg == SWB_BLOCK_SMALL ? SWB_NT_SMALL : SWB_NT_LARGE;
                const uint32_t wpb = ls.warps_active ? ls.warps_active : nt / 32u;
                // Block-coherent first wave (swb_warp_loop). Measured per rank workload of the multi-GPU layouts
                // (profiles/r2o_sweep_static_wave.txt): +1.3 .. +1.6 % on half of Swiss-Prot with a quarter of the queries
                // (1.7 tiles per warp), neutral on 1/4 and 1/8 of it and on the whole database with all 20 queries,
                // -2.9 % on the whole database with a quarter of the queries (3.7 tiles per warp: there the widest tile is
                // the critical path of the launch, and a block full of the widest tiles never gets a scheduler to
                // itself). Auto: on when the launch has fewer than two tiles per working warp.
                const uint32_t wave = (uint32_t)ls.grid * wpb;
                const bool on = e->opt_static_wave > 0 || (e->opt_static_wave < 0 && p.ntiles < 2u * wave);
                p.static_wave = (g.split || !on) ? 0u : wave;
            }
            p.first_chunk = ch.first;
            p.last_chunk = ch.last;
            p.counter = s.d_counters + counter++;
            if (SWB_PASS_GROUP > 1u && !g.split && (g.logg_mask & 1u) && gmode == SWB_MODE_S16) {  // blocked builds only
                // One-lane tiles park the row state of a pass group between column blocks: one region per warp of the
                // launch. The one-lane tiles of a pass sit in exactly one launch group and the passes of a query follow
                // each other in stream order, so one buffer per slot serves them all.
                const size_t elem = sizeof(uint32_t);
                const size_t warps = (size_t)ls.grid * (size_t)((ls.block_cfg == SWB_BLOCK_SMALL ? SWB_NT_SMALL : SWB_NT_LARGE) / 32);
                CU(GROW_DEV(s.d_colstate, s.colstate_cap, warps * swb_colstate_elems(g.K) * elem));
                p.colstate = s.d_colstate;
            }
            CU(swb_launch_score(g.K, gmode, g.split, ls.block_cfg, p, ls.grid, ls.smem, st));
            e->stats.kernel_launches += 1;
        }
        p.recount = recount;
    }
    for (size_t gi = 1; gi < ng; ++gi) {  // join
        CU(cudaEventRecord(s.ev_sub[gi - 1], s.sub[gi - 1]));
        CU(cudaStreamWaitEvent(s.stream, s.ev_sub[gi - 1], 0));
    }
    return SWB_OK;
}

// Enqueues everything one query needs. Results land in d_out[qi].
// Stream layout per slot: profile build and clears on slot.stream, then one sub-stream per launch group (the tiles of
// the group sizes that share a K) so that the few long-sequence tiles run beside the bulk, then the scatter.
static int enqueue_job(swb_engine *e, Slot &s, uint32_t qi, const uint8_t *q, uint32_t qlen)
{
    SwbPlan &pl = e->plan;
    const uint32_t nl = pl.n_local;
    if (nl == 0) return SWB_OK;
    int rc = ensure_slot(e, s);
    if (rc != SWB_OK) return rc;
    int32_t *out = e->d_out + (size_t)qi * nl;
    const uint32_t rows = qlen;
    if (rows == 0 || pl.tiles.empty() || pl.max_len == 0) {
        CU(cudaMemsetAsync(out, 0, sizeof(int32_t) * nl, s.stream));
        return SWB_OK;
    }
    const int ovf_thr = 32767 - e->max_s;
    uint32_t present = 0;
    for (int l = 0; l <= SWB_MAX_LOGG; ++l)
        if (pl.tiles_by_logg[l]) present |= 1u << l;
    const bool longest_first = e->opt_group_order == 1 || (e->opt_group_order == 0 && e->cur_nq <= 1);
    const bool affine = e->affine;  // never split
    const int mode0 = affine ? SWB_MODE_S16A : SWB_MODE_S16;

    const bool small_shard = pl.tiles.size() < 2u * (size_t)e->sm_count * (SWB_NT_LARGE / 32);
    // batches on small shards: 2048-row launches keep the shared-memory footprint of a block small, so blocks of several
    // queries share an SM and fill each other's tails (+1..2 % at 1/8 and 1/4 of Swiss-Prot per GPU); queries beyond one
    // chunk keep the long chunks (several short launches in a row cost the titin-scale workload a factor of two)
    const uint32_t chunk_rows =
        !e->chunk_rows_set && small_shard && e->cur_nq > 1 && rows <= e->chunk_rows ? 2048u : e->chunk_rows;
    const bool split = !affine && (e->opt_split == 1 || (e->opt_split < 0 && small_shard));

    // Exact passes (scores beyond the s16 range): V16R where the scoring scheme allows it, else int32
    const bool r16 = !affine && e->rebase_shift > 0 && e->opt_exact == 0;
    const int mode1 = affine ? SWB_MODE_I32A : (r16 ? SWB_MODE_R16 : SWB_MODE_I32);
    // The tiles of the pipelined-pass ("split") set, per lane-group size: the leading `direct` of them -- at least
    // direct_len wide, against a query at least that long -- go to V16R at once, the others through the s16 pass.
    uint32_t xl[SWB_MAX_LOGG + 1] = {}, direct[SWB_MAX_LOGG + 1] = {}, rest[SWB_MAX_LOGG + 1] = {};
    uint32_t n_direct = 0, n_rest = 0;
    if (split)
        for (int l = 1; l <= SWB_MAX_LOGG; ++l) {
            xl[l] = pl.xl_by_logg[l];
            if (r16 && e->opt_direct_len && rows >= e->opt_direct_len)
                while (direct[l] < xl[l] && pl.tiles[pl.tile_start_by_logg[l] + direct[l]].width >= e->opt_direct_len)
                    ++direct[l];
            rest[l] = xl[l] - direct[l];
            n_direct += direct[l];
            n_rest += rest[l];
        }
    // rows per lane of the split groups: 16 when that still leaves more work items than one-warp blocks fit on the GPU
    // (13 per SM with 16.5 KB of staged rows each), else 8 (twice the passes in flight per tile)
    auto pick_split_k = [&](const uint32_t *cnt) {
        if (e->opt_split_k == 32 && r16 && !e->rebase_shift32) return 16;
        if (e->opt_split_k) return e->opt_split_k;
        uint64_t items16 = 0;
        for (int l = 1; l <= SWB_MAX_LOGG; ++l) items16 += (uint64_t)cnt[l] * swb_split_passes(rows, l, 16);
        return items16 >= 13ull * (uint64_t)e->sm_count ? 16 : 8;
    };

    // pass 0: the s16 pass over all tiles but the direct ones
    SwbQueryPlan qp0;
    std::vector<SwbLaunchGroup> g0, gd;
    swb_plan_query(rows, e->opt_k, 32, present, chunk_rows, qp0);
    if (n_rest) {
        SwbLaunchGroup g;
        if (swb_plan_split_group(pl, direct, rest, pick_split_k(rest), g)) g0.push_back(g);
    }
    swb_plan_bulk_groups(pl, qp0, longest_first, split ? xl : nullptr, 0x3fu, g0);
    if (n_direct) {
        SwbLaunchGroup g;
        if (swb_plan_split_group(pl, nullptr, direct, pick_split_k(direct), g)) gd.push_back(g);
    }
    // exact pass over flagged tiles, only when a score can exceed the s16 range at all
    // (affine lanes carry (H, E) per row: their int32 recompute stops at 8 rows per lane)
    const bool need_i32 = (int64_t)e->max_s * std::min<uint32_t>(qlen, pl.max_len) > ovf_thr;
    SwbQueryPlan qp1;
    std::vector<SwbLaunchGroup> g1;
    uint32_t prof8_rows = qp0.prof_rows;
    if (need_i32 || n_direct) {
        swb_plan_query(qlen, e->opt_k, affine ? 8 : 16, present, chunk_rows, qp1);
        prof8_rows = std::max(prof8_rows, qp1.prof_rows);
    }
    if (split) prof8_rows = std::max(prof8_rows, swb_roundup(rows, 32u << SWB_MAX_LOGG));  // the last pass of a split item
    if (need_i32) {
        if (n_rest) {
            SwbLaunchGroup g;
            // the int32 kernels of the split groups exist for 8 rows per lane only
            if (swb_plan_split_group(pl, direct, rest, r16 ? pick_split_k(rest) : 8, g)) g1.push_back(g);
        }
        swb_plan_bulk_groups(pl, qp1, longest_first, split ? xl : nullptr, 0x3fu, g1);
    }
    size_t nlaunch = g0.size() * qp0.chunks.size() + gd.size() + g1.size() * (need_i32 ? qp1.chunks.size() : 0);
    if (nlaunch > SWB_MAX_COUNTERS) return fail(e, SWB_ERR_ARG, "query too long");
    if (g0.size() + gd.size() > 1 + SWB_MAX_SUB) return fail(e, SWB_ERR_ARG, "internal: too many launch groups");
    const uint32_t prof8_stride = swb_roundup(std::max(prof8_rows, 16u), 16);

    // buffers
    if (qlen > s.query_cap) {
        if (s.h_query) cudaFreeHost(s.h_query);
        if (s.d_query) cudaFree(s.d_query);
        s.h_query = nullptr;
        s.d_query = nullptr;
        s.query_cap = 0;
        const uint32_t cap = swb_roundup(qlen, 4096);
        CU(cudaMallocHost(&s.h_query, cap));
        CU(cudaMalloc(&s.d_query, cap));
        s.query_cap = cap;
    }
    CU(GROW_DEV(s.d_prof, s.prof_cap, (size_t)prof8_stride * SWB_ALPHA));
    // boundary elements: 4 B (V16, V16R), 8 B (V32, V16A: H and F), 16 B (V32A)
    if (affine) CU(GROW_DEV(s.d_bnd16, s.bnd16_cap, 8ull * pl.bnd_elems));
    if (need_i32 && !r16) CU(GROW_DEV(s.d_bnd32, s.bnd32_cap, (affine ? 16ull : 8ull) * pl.bnd_elems));
    if (r16 && (need_i32 || n_direct))
        CU(GROW_DEV(s.d_blog, s.blog_cap, sizeof(uint2) * swb_blog_elems(pl.bnd_elems, (uint32_t)pl.tiles.size())));
    // progress counters of the split groups: one per (tile, pass); the direct group's follow those of the s16 group, and
    // the exact pass reuses the s16 group's after it (same stream order, cleared in between)
    size_t prog_words = 0, prog_direct = 0, prog_words1 = 0;
    if (!g0.empty() && g0[0].split) prog_words = swb_split_items(rows, g0[0], nullptr);
    if (!gd.empty()) prog_direct = swb_split_items(rows, gd[0], nullptr);
    if (need_i32 && !g1.empty() && g1[0].split) prog_words1 = swb_split_items(rows, g1[0], nullptr);
    const size_t prog_main = std::max(prog_words, prog_words1);
    if (prog_main + prog_direct) CU(GROW_DEV(s.d_prog, s.prog_cap, sizeof(uint32_t) * (prog_main + prog_direct)));

    memcpy(s.h_query, q, qlen);
    CU(cudaMemcpyAsync(s.d_query, s.h_query, qlen, cudaMemcpyHostToDevice, s.stream));
    CU(swb_launch_profile(s.d_query, qlen, e->d_mat, e->gap, s.d_prof, prof8_stride, prof8_rows, s.stream));
    CU(cudaMemsetAsync(s.d_state, 0, s.state_bytes, s.stream));
    if (prog_main + prog_direct) CU(cudaMemsetAsync(s.d_prog, 0, sizeof(uint32_t) * (prog_main + prog_direct), s.stream));
    e->stats.kernel_launches += 1;

    SwbScoreParams p;
    memset(&p, 0, sizeof p);
    p.tiles = e->d_tiles;
    p.residues = e->d_residues;
    p.flags = s.d_flags;
    p.recount = s.d_recount;
    p.gap = e->gap;
    p.gap_open = e->gap;
    p.gap_extend = e->gap_extend;
    p.ovf_thr = ovf_thr;
    p.rebase_shift = (uint32_t)e->rebase_shift;
    p.blog = s.d_blog;
    uint32_t counter = 0;
    p.profile = s.d_prof;
    p.prof_stride = prof8_stride;
    p.scores = s.d_sorted;
    p.bnd = s.d_bnd16;
    p.only_flagged = 0;
    // the direct group runs beside the s16 pass (disjoint tiles): it is enqueued as one more launch group of pass 0
    if ((rc = enqueue_pass(e, s, mode0, p, qp0, g0, counter, s.d_prog, &gd, s.d_prog + prog_main)) != SWB_OK) return rc;
    if (need_i32 && !g1.empty()) {
        p.bnd = r16 ? (void *)s.d_bnd16 : s.d_bnd32;
        p.only_flagged = 1;
        if (g1[0].split) {  // pipelined work items combine their scores with atomicMax
            CU(cudaMemsetAsync(s.d_prog, 0, sizeof(uint32_t) * prog_words1, s.stream));
            CU(swb_launch_clear_flagged(e->d_tiles, (uint32_t)pl.tiles.size(), s.d_flags, p.scores, s.stream));
            e->stats.kernel_launches += 1;
        }
        if ((rc = enqueue_pass(e, s, mode1, p, qp1, g1, counter, s.d_prog, nullptr, nullptr)) != SWB_OK) return rc;
    }
    CU(swb_launch_scatter(s.d_sorted, e->d_out_pos, nl, out, s.stream));
    e->stats.kernel_launches += 1;
    e->stats.last_k = (uint32_t)qp0.k_by_logg[0];
    for (int l = 0; l <= SWB_MAX_LOGG; ++l)
        e->stats.padded_cells += pl.cols_by_logg[l] * (uint64_t)swb_roundup(rows, (uint32_t)qp0.k_by_logg[l] << l);
    e->stats.cells += (uint64_t)qlen * pl.residues_local;
    return SWB_OK;
}

// waits for the slot's job and hands its results to the caller's memory
static int finish_slot(swb_engine *e, Slot &s)
{
    if (!s.busy) return SWB_OK;
    s.busy = false;
    int32_t *dst = s.pending_dst;
    uint32_t *ids = s.pending_ids;
    int32_t *top = s.pending_top;
    s.pending_dst = nullptr;
    s.pending_ids = nullptr;
    s.pending_top = nullptr;
    CU(cudaEventSynchronize(s.done));
    const size_t nl = e->plan.n_local;
    if (dst && !s.pending_scatter) {
        memcpy(dst, s.h_scores, sizeof(int32_t) * nl);
    } else if (dst) {  // a full-database vector shared with the engines of the other shards
        const uint32_t *sid = e->plan.shard_ids.data();
        for (size_t k = 0; k < nl; ++k) dst[sid[k]] = s.h_scores[k];
    }
    if (ids && top) {
        memcpy(ids, s.h_topk, sizeof(uint32_t) * s.pending_k);
        memcpy(top, s.h_topk + s.pending_k, sizeof(int32_t) * s.pending_k);
    }
    return SWB_OK;
}

// index of a slot among the first `ns` whose job has completed (polls the completion events)
static int wait_any_slot(swb_engine *e, int ns, int *out)
{
    for (;;) {
        for (int i = 0; i < ns; ++i) {
            Slot &s = e->slots[i];
            if (!s.busy) { *out = i; return SWB_OK; }
            const cudaError_t q = cudaEventQuery(s.done);
            if (q == cudaSuccess) { *out = i; return SWB_OK; }
            if (q != cudaErrorNotReady) return fail(e, SWB_ERR_CUDA, std::string("cudaEventQuery: ") + cudaGetErrorString(q));
        }
        std::this_thread::sleep_for(std::chrono::microseconds(30));
    }
}

// what a batch hands back per query
struct BatchOut {
    int32_t *scores = nullptr;   // nq x n_local (database order of the shard), or, with full_stride != 0,
    uint64_t full_stride = 0;    //   rows of full_stride entries: vectors of the WHOLE database, this shard fills its ids
    const uint32_t *row_of = nullptr;  // full_stride != 0: row of query q (NULL: q)
    uint32_t k = 0;              // device-side selection: k best per query
    uint32_t *ids = nullptr;     // nq x k
    int32_t *top = nullptr;      // nq x k
};

// the copies and the selection kernel that follow the scan of query qa on the slot's stream
static int enqueue_results(swb_engine *e, Slot &s, uint32_t qa, const BatchOut &bo)
{
    const uint32_t nl = e->plan.n_local;
    if (bo.k) {
        const size_t bytes = 2 * sizeof(uint32_t) * (size_t)bo.k;
        CU(GROW_DEV(s.d_topk, s.topk_cap, bytes));
        CU(GROW_HOST(s.h_topk, s.h_topk_cap, bytes));
        CU(swb_launch_topk(e->d_out + (size_t)qa * nl, nl, e->plan.nshards > 1 ? e->d_shard_ids : nullptr, bo.k, s.d_topk,
                           reinterpret_cast<int32_t *>(s.d_topk + bo.k), s.stream));
        e->stats.kernel_launches += 1;
        CU(cudaMemcpyAsync(s.h_topk, s.d_topk, bytes, cudaMemcpyDeviceToHost, s.stream));
        s.pending_ids = bo.ids + (size_t)qa * bo.k;
        s.pending_top = bo.top + (size_t)qa * bo.k;
        s.pending_k = bo.k;
    }
    if (bo.scores && nl > 0) {
        CU(cudaMemcpyAsync(s.h_scores, e->d_out + (size_t)qa * nl, sizeof(int32_t) * nl, cudaMemcpyDeviceToHost,
                           s.stream));
        s.pending_scatter = bo.full_stride != 0;
        s.pending_dst = bo.full_stride ? bo.scores + (size_t)(bo.row_of ? bo.row_of[qa] : qa) * bo.full_stride
                                       : bo.scores + (size_t)qa * nl;
    }
    return SWB_OK;
}

static int search_batch_impl(swb_engine *e, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq,
                             const BatchOut &bo)
{
    if (!e || !qoffsets || (!qcodes && nq && qoffsets[nq] != qoffsets[0])) return SWB_ERR_ARG;
    if (!e->db_loaded) return fail(e, SWB_ERR_STATE, "swb_search before swb_db_load");
    if (!e->scoring_set) return fail(e, SWB_ERR_STATE, "no scoring set");
    CU(cudaSetDevice(e->device));
    const uint32_t nl = e->plan.n_local;
    for (uint32_t i = 0; i < nq; ++i)
        if (qoffsets[i + 1] < qoffsets[i] || qoffsets[i + 1] - qoffsets[i] > 0x7fffffffull)
            return fail(e, SWB_ERR_ARG, "bad query offsets");
    cudaStream_t ms = main_stream(e);
    e->last_nq = 0;  // set again once every result of this batch is in place
    if (sizeof(int32_t) * (size_t)nq * nl > e->out_cap && nl > 0) {
        CU(cudaStreamSynchronize(ms));
        CU(GROW_DEV(e->d_out, e->out_cap, sizeof(int32_t) * (size_t)nq * nl));
    }
    e->stats.cells = 0;
    e->stats.padded_cells = 0;
    e->stats.kernel_launches = 0;
    e->stats.recomputed_tiles = 0;
    e->cur_nq = nq;
    // jobs: queries are taken longest first (the tiles of a long query are long-running work items: started last they
    // would leave the GPU half empty at the end of the batch; option batch_order = 1 keeps the caller's order)
    std::vector<uint32_t> order(nq);
    for (uint32_t i = 0; i < nq; ++i) order[i] = i;
    if (e->opt_batch_order == 0)
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            return qoffsets[a + 1] - qoffsets[a] > qoffsets[b + 1] - qoffsets[b];
        });
    const int ns = (int)std::min<uint32_t>((uint32_t)e->nslots, std::max<uint32_t>(1u, nq));  // streams in use
    // From the first enqueue on, a failure must not return before the common epilogue has waited for the slots and
    // cleared their pending destinations: they point into the caller's buffers of THIS call.
    int rc = SWB_OK;
    auto cu = [&](cudaError_t ce, const char *what) {
        if (ce == cudaSuccess || rc != SWB_OK) return ce == cudaSuccess;
        rc = fail(e, SWB_ERR_CUDA, std::string(what) + " failed: " + cudaGetErrorString(ce));
        return false;
    };
    cu(cudaEventRecord(e->ev_start, ms), "cudaEventRecord");
    cu(cudaEventRecord(e->ev_fork, ms), "cudaEventRecord");
    for (int i = 0; i < ns && rc == SWB_OK; ++i) {
        cu(cudaStreamWaitEvent(e->slots[i].stream, e->ev_fork, 0), "cudaStreamWaitEvent");
        cu(cudaMemsetAsync(e->slots[i].d_recount, 0, sizeof(uint32_t), e->slots[i].stream), "cudaMemsetAsync");
    }
    for (uint32_t j = 0; j < nq && rc == SWB_OK; ++j) {
        // the first `ns` jobs take the slots in turn; later ones take whichever slot finishes first (a slot that holds
        // a long query must not hold up the queue behind it)
        int si = (int)(j % (uint32_t)ns);
        if (j >= (uint32_t)ns && (rc = wait_any_slot(e, ns, &si)) != SWB_OK) break;
        Slot &s = e->slots[si];
        if ((rc = finish_slot(e, s)) != SWB_OK) break;
        const uint32_t qa = order[j];
        const uint32_t la = (uint32_t)(qoffsets[qa + 1] - qoffsets[qa]);
        if ((rc = enqueue_job(e, s, qa, qcodes + qoffsets[qa], la)) != SWB_OK) break;
        if ((rc = enqueue_results(e, s, qa, bo)) != SWB_OK) break;
        if (!cu(cudaEventRecord(s.done, s.stream), "cudaEventRecord")) break;
        s.busy = true;
    }
    for (int i = 0; i < ns && rc == SWB_OK; ++i) {
        cu(cudaMemcpyAsync(e->h_recount + i, e->slots[i].d_recount, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                           e->slots[i].stream), "cudaMemcpyAsync");
        cu(cudaEventRecord(e->ev_join[i], e->slots[i].stream), "cudaEventRecord");
        cu(cudaStreamWaitEvent(ms, e->ev_join[i], 0), "cudaStreamWaitEvent");
    }
    if (rc == SWB_OK) cu(cudaEventRecord(e->ev_stop, ms), "cudaEventRecord");
    if (rc == SWB_OK) {
        for (int i = 0; i < ns; ++i) {
            const int r2 = finish_slot(e, e->slots[i]);
            if (rc == SWB_OK) rc = r2;
        }
        if (rc == SWB_OK) cu(cudaStreamSynchronize(ms), "cudaStreamSynchronize");
    }
    if (rc != SWB_OK) {
        // drain whatever was enqueued and forget the destinations of this call
        const std::string msg = e->err;
        for (int i = 0; i < SWB_MAX_SLOTS; ++i) {
            Slot &s = e->slots[i];
            cudaStreamSynchronize(s.stream);
            s.busy = false;
            s.pending_dst = nullptr;
            s.pending_ids = nullptr;
            s.pending_top = nullptr;
        }
        cudaStreamSynchronize(ms);
        e->err = msg;
        return rc;
    }
    float ms_f = 0;
    CU(cudaEventElapsedTime(&ms_f, e->ev_start, e->ev_stop));
    e->stats.device_ms = ms_f;
    for (int i = 0; i < ns; ++i) e->stats.recomputed_tiles += e->h_recount[i];
    e->last_nq = nq;
    return SWB_OK;
}

extern "C" int swb_search_batch(swb_engine *e, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq,
                                int32_t *scores)
{
    BatchOut bo;
    bo.scores = scores;
    return search_batch_impl(e, qcodes, qoffsets, nq, bo);
}

int swb_search_batch_rows(swb_engine *e, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq,
                          int32_t *scores_full, uint64_t n_total, const uint32_t *row_of)
{
    if (!e || !scores_full) return SWB_ERR_ARG;
    if (e->db_loaded && n_total < e->plan.n_total) return fail(e, SWB_ERR_ARG, "n_total is smaller than the database");
    BatchOut bo;
    bo.scores = scores_full;
    bo.full_stride = n_total ? n_total : 1;
    bo.row_of = row_of;
    return search_batch_impl(e, qcodes, qoffsets, nq, bo);
}

extern "C" int swb_search_batch_scatter(swb_engine *e, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq,
                                        int32_t *scores_full, uint64_t n_total)
{
    return swb_search_batch_rows(e, qcodes, qoffsets, nq, scores_full, n_total, nullptr);
}

extern "C" int swb_search_batch_topk(swb_engine *e, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq,
                                     uint32_t k, uint32_t *ids, int32_t *top)
{
    if (!e || !ids || !top) return SWB_ERR_ARG;
    if (k < 1 || k > SWB_TOPK_MAX) return fail(e, SWB_ERR_ARG, "k must be 1..1024");
    BatchOut bo;
    bo.k = k;
    bo.ids = ids;
    bo.top = top;
    if (e->db_loaded && e->plan.n_local == 0) {  // an empty shard has no hits
        for (size_t i = 0; i < (size_t)nq * k; ++i) {
            ids[i] = 0xffffffffu;
            top[i] = -1;
        }
        bo.k = 0;
    }
    return search_batch_impl(e, qcodes, qoffsets, nq, bo);
}

extern "C" int swb_search(swb_engine *e, const uint8_t *query, uint32_t qlen, int32_t *scores)
{
    if (!e || (!query && qlen) || !scores) return SWB_ERR_ARG;
    const uint64_t offs[2] = {0, qlen};
    static const uint8_t dummy = 0;
    return swb_search_batch(e, query ? query : &dummy, offs, 1, scores);
}

extern "C" int swb_fetch_scores(swb_engine *e, uint32_t query_index, int32_t *scores)
{
    if (!e || !scores) return SWB_ERR_ARG;
    if (!e->db_loaded || query_index >= e->last_nq) return fail(e, SWB_ERR_STATE, "no such result");
    CU(cudaSetDevice(e->device));
    const uint32_t nl = e->plan.n_local;
    if (nl == 0) return SWB_OK;
    if (sizeof(int32_t) * ((size_t)query_index + 1) * nl > e->out_cap) return fail(e, SWB_ERR_STATE, "no such result");
    CU(cudaMemcpy(scores, e->d_out + (size_t)query_index * nl, sizeof(int32_t) * nl, cudaMemcpyDeviceToHost));
    return SWB_OK;
}

// Traceback alignments of a list of hits in one launch (one block per hit; swb_kernels.cu). Hits are processed in waves
// whose direction matrices (2 bits per cell) fit SWB_ALIGN_DIR_BUDGET; a wave is one kernel launch.
#define SWB_ALIGN_DIR_BUDGET (4ull << 30)
extern "C" int swb_align_batch(swb_engine *e, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq,
                               const uint32_t *hit_query, const uint32_t *hit_db_id, uint32_t nhits, int32_t *scores,
                               uint32_t *end_i, uint32_t *end_j, uint8_t *ops, const uint64_t *ops_offsets,
                               uint32_t *nops)
{
    if (!e || !qoffsets || (nhits && (!hit_query || !hit_db_id || !scores)) || (ops && !ops_offsets)) return SWB_ERR_ARG;
    if (!e->db_loaded) return fail(e, SWB_ERR_STATE, "swb_align before swb_db_load");
    const bool aff = e->affine;
    for (uint32_t i = 0; i < nq; ++i)
        if (qoffsets[i + 1] < qoffsets[i] || qoffsets[i + 1] - qoffsets[i] > 0x7ffffff0ull)
            return fail(e, SWB_ERR_ARG, "bad query offsets");
    const uint64_t qbytes = nq ? qoffsets[nq] - qoffsets[0] : 0;
    if (qbytes && !qcodes) return SWB_ERR_ARG;
    const SwbPlan &pl = e->plan;
    std::vector<SwbAlignJob> jobs(nhits);
    std::vector<uint32_t> live;  // hits with a non-empty matrix
    live.reserve(nhits);
    for (uint32_t h = 0; h < nhits; ++h) {
        if (hit_query[h] >= nq) return fail(e, SWB_ERR_ARG, "hit_query out of range");
        const std::vector<uint32_t>::const_iterator it =
            std::lower_bound(pl.shard_ids.begin(), pl.shard_ids.end(), hit_db_id[h]);
        if (it == pl.shard_ids.end() || *it != hit_db_id[h]) return fail(e, SWB_ERR_ARG, "db_id is not part of this shard");
        const uint32_t spos = pl.sorted_of_out[(size_t)(it - pl.shard_ids.begin())];
        SwbAlignJob &jb = jobs[h];
        memset(&jb, 0, sizeof jb);
        jb.q_off = qoffsets[hit_query[h]] - qoffsets[0];
        jb.d_off = pl.seq_off[spos];
        jb.m = (uint32_t)(qoffsets[hit_query[h] + 1] - qoffsets[hit_query[h]]);
        jb.n = pl.seq_len[spos];
        const uint64_t room = ops ? ops_offsets[h + 1] - ops_offsets[h] : 0;
        if (ops && ops_offsets[h + 1] < ops_offsets[h]) return fail(e, SWB_ERR_ARG, "bad ops offsets");
        jb.cap = (uint32_t)std::min<uint64_t>(room, 0xffffffffull);
        scores[h] = 0;
        if (end_i) end_i[h] = 0;
        if (end_j) end_j[h] = 0;
        if (nops) nops[h] = 0;
        if (jb.m == 0 || jb.n == 0) continue;
        if ((uint64_t)(jb.m + 1) * swb_align_row_bytes(jb.n, aff) > SWB_ALIGN_DIR_BUDGET)
            return fail(e, SWB_ERR_ARG, "alignment matrix larger than 2^34 cells");
        live.push_back(h);
    }
    if (live.empty()) return SWB_OK;
    CU(cudaSetDevice(e->device));
    cudaStream_t st = main_stream(e);
    CU(GROW_DEV(e->d_align_q, e->align_q_cap, (size_t)qbytes));
    if (qbytes) CU(cudaMemcpyAsync(e->d_align_q, qcodes + qoffsets[0], (size_t)qbytes, cudaMemcpyHostToDevice, st));
    const uint32_t smem_cap_ints = SWB_ALIGN_SMEM_MAX / sizeof(int32_t);
    size_t at = 0;
    while (at < live.size()) {
        // one wave: as many of the remaining hits as the direction budget holds (at least one)
        size_t end = at;
        uint64_t dir_bytes = 0, hd_ints = 0, ops_bytes = 0;
        uint32_t smem_ints = 0;
        std::vector<SwbAlignJob> wave;
        while (end < live.size()) {
            SwbAlignJob jb = jobs[live[end]];
            const uint64_t need = (uint64_t)(jb.m + 1) * swb_align_row_bytes(jb.n, aff);
            if (end > at && dir_bytes + need > SWB_ALIGN_DIR_BUDGET) break;
            jb.dir_off = dir_bytes;
            dir_bytes += (need + 15u) & ~15ull;
            jb.ops_off = ops_bytes;
            ops_bytes += jb.cap;
            const uint64_t hd = swb_align_hd_ints(jb.m, aff);
            if (hd <= smem_cap_ints) {
                smem_ints = std::max<uint32_t>(smem_ints, (uint32_t)hd);
            } else {
                jb.hd_off = hd_ints;
                hd_ints += hd;
            }
            wave.push_back(jb);
            ++end;
        }
        const size_t nw = wave.size();
        const size_t out_bytes = 5 * sizeof(int32_t) * nw + (size_t)ops_bytes;
        CU(GROW_DEV(e->d_align_jobs, e->align_jobs_cap, sizeof(SwbAlignJob) * nw));
        CU(GROW_DEV(e->d_align_h, e->align_h_cap, sizeof(int32_t) * (size_t)std::max<uint64_t>(hd_ints, 1)));
        CU(GROW_DEV(e->d_align_dir, e->align_dir_cap, (size_t)dir_bytes));
        CU(GROW_DEV(e->d_align_out, e->align_out_cap, out_bytes));
        CU(GROW_HOST(e->h_align_out, e->align_hout_cap, out_bytes));
        CU(cudaMemcpyAsync(e->d_align_jobs, wave.data(), sizeof(SwbAlignJob) * nw, cudaMemcpyHostToDevice, st));
        int32_t *d_hdr = reinterpret_cast<int32_t *>(e->d_align_out);
        uint8_t *d_ops = e->d_align_out + 5 * sizeof(int32_t) * nw;
        CU(swb_launch_align_batch(e->d_align_jobs, (uint32_t)nw, e->d_align_q, e->d_raw, e->d_mat, e->gap, e->gap_extend, aff,
                                  e->d_align_h, e->d_align_dir, d_hdr, d_ops, smem_ints, st));
        CU(cudaMemcpyAsync(e->h_align_out, e->d_align_out, out_bytes, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));  // the pageable `wave` vector was the source of an async copy: done by now
        const int32_t *hdr = reinterpret_cast<const int32_t *>(e->h_align_out);
        const uint8_t *hops = e->h_align_out + 5 * sizeof(int32_t) * nw;
        for (size_t w = 0; w < nw; ++w) {
            const uint32_t h = live[at + w];
            scores[h] = hdr[5 * w];
            if (end_i) end_i[h] = (uint32_t)hdr[5 * w + 1];
            if (end_j) end_j[h] = (uint32_t)hdr[5 * w + 2];
            const uint32_t cnt = (uint32_t)hdr[5 * w + 3];
            if (nops) nops[h] = cnt;
            if (!ops) continue;
            if (hdr[5 * w + 4] || cnt > wave[w].cap)
                return fail(e, SWB_ERR_ARG, "ops buffer too small (qlen + subject length always suffices)");
            // the kernel walks from the end of the alignment to its start (cpu.cpp:80-103); hand it out start to end
            uint8_t *dst = ops + ops_offsets[h];
            const uint8_t *src = hops + wave[w].ops_off;
            for (uint32_t k = 0; k < cnt; ++k) dst[k] = src[cnt - 1 - k];
        }
        at = end;
    }
    return SWB_OK;
}

extern "C" int swb_align(swb_engine *e, const uint8_t *query, uint32_t qlen, uint32_t db_id, int32_t *score,
                         uint32_t *end_i, uint32_t *end_j, uint8_t *ops, uint32_t cap, uint32_t *nops)
{
    if (!e || (!query && qlen) || !score) return SWB_ERR_ARG;
    const uint64_t qoff[2] = {0, qlen}, ooff[2] = {0, ops ? cap : 0u};
    const uint32_t hq = 0;
    return swb_align_batch(e, query, qoff, 1, &hq, &db_id, 1, score, end_i, end_j, ops, ooff, nops);
}

extern "C" int swb_topk(const swb_engine *e, const int32_t *scores, uint32_t k, uint32_t *ids, int32_t *top)
{
    if (!e || !scores || !ids || !top) return SWB_ERR_ARG;
    if (!e->db_loaded) return SWB_ERR_STATE;
    const uint32_t nl = e->plan.n_local;
    const uint32_t kk = std::min(k, nl);
    std::vector<uint32_t> idx(nl);
    for (uint32_t i = 0; i < nl; ++i) idx[i] = i;
    auto better = [&](uint32_t a, uint32_t b) { return scores[a] != scores[b] ? scores[a] > scores[b] : a < b; };
    std::partial_sort(idx.begin(), idx.begin() + kk, idx.end(), better);
    for (uint32_t i = 0; i < kk; ++i) {
        ids[i] = e->plan.shard_ids[idx[i]];
        top[i] = scores[idx[i]];
    }
    for (uint32_t i = kk; i < k; ++i) {
        ids[i] = 0xffffffffu;
        top[i] = -1;
    }
    return SWB_OK;
}

extern "C" int swb_stats(const swb_engine *e, swb_stats_t *out)
{
    if (!e || !out) return SWB_ERR_ARG;
    *out = e->stats;
    return SWB_OK;
}

extern "C" int swb_plan_describe(const uint64_t *offsets, uint32_t n, uint32_t shard, uint32_t nshards,
                                 uint32_t group_len, swb_plan_info_t *info, uint32_t *sorted_ids, uint32_t *shard_ids)
{
    if (!offsets || !info) return SWB_ERR_ARG;
    SwbPlanOpts o;
    if (group_len) o.group_len = group_len;
    SwbPlan pl;
    if (swb_build_plan(offsets, n, shard, nshards ? nshards : 1, o, pl) != 0) return SWB_ERR_ARG;
    memset(info, 0, sizeof *info);
    info->n_total = pl.n_total;
    info->n_local = pl.n_local;
    info->tiles = (uint32_t)pl.tiles.size();
    info->max_len = pl.max_len;
    info->residues_local = pl.residues_local;
    info->residues_total = pl.residues_total;
    info->res_bytes = pl.res_bytes;
    info->bnd_elems = pl.bnd_elems;
    info->padded_cols = pl.padded_cols;
    for (int l = 0; l <= SWB_MAX_LOGG; ++l) info->tiles_by_group[l] = pl.tiles_by_logg[l];
    if (sorted_ids) memcpy(sorted_ids, pl.sorted_ids.data(), sizeof(uint32_t) * pl.n_local);
    if (shard_ids) memcpy(shard_ids, pl.shard_ids.data(), sizeof(uint32_t) * pl.n_local);
    return SWB_OK;
}
