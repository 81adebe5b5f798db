// Engine group: all the GPUs of one box behind ONE handle in ONE process -- what a caller of the reference's single
// entry point gets (main.cpp:52-56 creates one FASTADatabase and calls smith_waterman_cuda once; there is no process
// per GPU in that world). One engine + one host worker thread + one stream set per GPU.
//
// Layout: the devices form a grid of P database parts x R query groups (P * R = devices). Device (p, r) keeps part p of
// the residue-balanced sharding resident (include/swb.h, swb_db_load) and scores the queries of group r against it, so
// every (query, sequence) pair is scored exactly once. P is chosen per load: a part should not get much smaller than a
// Swiss-Prot half (a small shard has fewer warp tiles than the GPU has warp slots and loses its tail), and what is left of
// the devices splits the QUERIES of a batch instead (longest-processing-time first, so the groups carry equal numbers of
// query rows). There is no data-path collective: scores are scattered straight into the caller's vector by database id,
// per-GPU hit lists are merged on the host.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/swb.h"
#include "swb_internal.h"
#include "swb_plan.h"

namespace {

// one host thread per GPU, alive as long as the group: a call hands every worker its job and waits for all of them
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, quit = false, done = true;
    void loop()
    {
        for (;;) {
            std::function<void()> j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return has_job || quit; });
                if (quit) return;
                j.swap(job);
                has_job = false;
            }
            j();
            {
                std::lock_guard<std::mutex> lk(mu);
                done = true;
            }
            cv.notify_all();
        }
    }
    void start(std::function<void()> j)
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = std::move(j);
            has_job = true;
            done = false;
        }
        cv.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
    }
};

}  // namespace

struct swb_group {
    std::vector<swb_engine *> eng;
    std::vector<int> devices;
    std::vector<Worker *> workers;
    std::string err;
    int parts = 0;          // database parts P of the loaded layout (0 = nothing loaded)
    int parts_forced = 0;   // option "db_parts": 0 = choose per load
    uint32_t min_part = SWB_MIN_PART_SEQUENCES;  // option "min_part_sequences"
    uint32_t n_total = 0;
    swb_stats_t stats;
    // runs f(device index) on every worker, returns the first error
    int run_all(const std::function<int(int)> &f)
    {
        const int nd = (int)eng.size();
        std::vector<int> rc(nd, SWB_OK);
        for (int i = 0; i < nd; ++i) workers[i]->start([&, i] { rc[i] = f(i); });
        for (int i = 0; i < nd; ++i) workers[i]->wait();
        for (int i = 0; i < nd; ++i)
            if (rc[i] != SWB_OK) {
                err = "device " + std::to_string(devices[i]) + ": " + swb_last_error(eng[i]);
                return rc[i];
            }
        return SWB_OK;
    }
};

static thread_local std::string g_group_create_error;

extern "C" int swb_group_create(swb_group **out, const int *devices, int ndev)
{
    if (!out || ndev < 0) return SWB_ERR_ARG;
    *out = nullptr;
    int visible = 0;
    const cudaError_t ce = cudaGetDeviceCount(&visible);
    if (ce != cudaSuccess || visible == 0) {
        g_group_create_error = std::string("no CUDA device: ") + cudaGetErrorString(ce) + " (this library has no CPU fallback)";
        return SWB_ERR_CUDA;
    }
    if (ndev == 0) ndev = visible;
    swb_group *g = new swb_group();
    memset(&g->stats, 0, sizeof g->stats);
    for (int i = 0; i < ndev; ++i) {
        const int dev = devices ? devices[i] : i;
        swb_engine *e = nullptr;
        const int rc = swb_create(&e, dev);
        if (rc != SWB_OK) {
            g_group_create_error = std::string("device ") + std::to_string(dev) + ": " + swb_last_error(nullptr);
            swb_group_destroy(g);
            return rc;
        }
        g->eng.push_back(e);
        g->devices.push_back(dev);
    }
    for (int i = 0; i < ndev; ++i) {
        Worker *w = new Worker();
        w->th = std::thread([w] { w->loop(); });
        g->workers.push_back(w);
    }
    *out = g;
    return SWB_OK;
}

// SWB_DEVICES=<i,j,...> names the devices (an index may repeat: several engines on one device, which is how the
// multi-device path is exercised on a one-GPU box), else SWB_GPUS=<n> takes the first n, else every visible device
extern "C" int swb_group_create_env(swb_group **out)
{
    std::vector<int> devs;
    if (const char *list = getenv("SWB_DEVICES")) {
        const char *p = list;
        while (*p) {
            char *end = nullptr;
            const long v = strtol(p, &end, 10);
            if (end == p) break;
            devs.push_back((int)v);
            p = *end == ',' ? end + 1 : end;
        }
    }
    if (!devs.empty()) return swb_group_create(out, devs.data(), (int)devs.size());
    const char *n = getenv("SWB_GPUS");
    return swb_group_create(out, nullptr, n ? std::max(0, atoi(n)) : 0);
}

extern "C" void swb_group_destroy(swb_group *g)
{
    if (!g) return;
    for (Worker *w : g->workers) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
        }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (swb_engine *e : g->eng) swb_destroy(e);
    delete g;
}

extern "C" const char *swb_group_last_error(const swb_group *g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }
extern "C" int swb_group_size(const swb_group *g) { return g ? (int)g->eng.size() : 0; }
extern "C" swb_engine *swb_group_engine(swb_group *g, int i)
{
    return (g && i >= 0 && i < (int)g->eng.size()) ? g->eng[i] : nullptr;
}
extern "C" int swb_group_db_parts(const swb_group *g) { return g ? g->parts : 0; }

extern "C" int swb_group_set_option(swb_group *g, const char *key, int64_t value)
{
    if (!g || !key) return SWB_ERR_ARG;
    if (!strcmp(key, "db_parts")) {
        if (value < 0 || (value > 0 && (int64_t)g->eng.size() % value != 0)) {
            g->err = "db_parts must be 0 (auto) or a divisor of the number of devices";
            return SWB_ERR_ARG;
        }
        g->parts_forced = (int)value;
        return SWB_OK;
    }
    if (!strcmp(key, "min_part_sequences")) {
        if (value < 1 || value > 0x7fffffff) return SWB_ERR_ARG;
        g->min_part = (uint32_t)value;
        return SWB_OK;
    }
    for (size_t i = 0; i < g->eng.size(); ++i) {
        const int rc = swb_set_option(g->eng[i], key, value);
        if (rc != SWB_OK) {
            g->err = swb_last_error(g->eng[i]);
            return rc;
        }
    }
    return SWB_OK;
}

extern "C" int swb_group_set_scoring_affine(swb_group *g, const int8_t *matrix, int alpha, int gap_open, int gap_extend)
{
    if (!g) return SWB_ERR_ARG;
    return g->run_all([&](int i) { return swb_set_scoring_affine(g->eng[i], matrix, alpha, gap_open, gap_extend); });
}
extern "C" int swb_group_set_scoring(swb_group *g, const int8_t *matrix, int alpha, int gap)
{
    return swb_group_set_scoring_affine(g, matrix, alpha, gap, gap);
}
extern "C" int swb_group_set_scoring_preset(swb_group *g, int preset)
{
    if (!g) return SWB_ERR_ARG;
    return g->run_all([&](int i) { return swb_set_scoring_preset(g->eng[i], preset); });
}

// P = the largest divisor of the device count whose parts keep at least `min_part` sequences (at least 1)
extern "C" int swb_layout_parts(uint32_t n, int ndev, uint32_t min_part)
{
    int best = 1;
    for (int p = 1; p <= ndev; ++p)
        if (ndev % p == 0 && (uint64_t)n / (uint64_t)p >= min_part) best = p;
    return best;
}

// Longest-processing-time-first split of nq queries into `groups` groups of (nearly) equal total length:
// group_of[q] for every query. Ties keep the caller's order, so the result is deterministic.
extern "C" int swb_layout_query_groups(const uint64_t *qoffsets, uint32_t nq, int groups, uint32_t *group_of)
{
    if (!qoffsets || !group_of || groups < 1) return SWB_ERR_ARG;
    std::vector<uint32_t> order(nq);
    for (uint32_t i = 0; i < nq; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return qoffsets[a + 1] - qoffsets[a] > qoffsets[b + 1] - qoffsets[b];
    });
    std::vector<uint64_t> load((size_t)groups, 0);
    for (uint32_t j = 0; j < nq; ++j) {
        int arg = 0;
        for (int r = 1; r < groups; ++r)
            if (load[r] < load[arg]) arg = r;
        group_of[order[j]] = (uint32_t)arg;
        load[arg] += qoffsets[order[j] + 1] - qoffsets[order[j]] + 1;  // + 1: empty queries still spread out
    }
    return SWB_OK;
}

// The same with the batch in view: with P parts the queries split into R = ndev / P groups. The database is split over
// all the devices instead (P = ndev, one query group) when the batch is too small or too uneven for R groups: fewer
// than SWB_MIN_GROUP_QUERIES queries per group, or the heaviest group more than 2 % above the mean total length.
//  - Few queries per group: a query's launch ends with a tail in which its longest lane-group tiles run alone, and only
//    other queries in flight fill it. Measured with the 20 reference queries on whole copies of Swiss-Prot: 10 per GPU
//    0.98 of the 20-query rate (2 GPUs, profiles/r2zf_*), 5 per GPU 0.87 (4 GPUs) -- while an eighth of the database
//    with all 20 queries runs at 0.89 of the whole (profiles/r2zd_*), a quarter necessarily above that.
//  - Layouts in between (P 2 x R 4 on 8 GPUs, profiles/r2k_*, r2w_*): the same device time as the pure split within
//    1.5 %, but every query group uploads its own copy of a part through the same host; end to end the pure split was
//    17 % faster twice (58.8 against 49.7 TCUPS).
#ifndef SWB_MIN_GROUP_QUERIES
#define SWB_MIN_GROUP_QUERIES 8u
#endif
extern "C" int swb_layout_parts_batch(uint32_t n, int ndev, uint32_t min_part, const uint64_t *qoffsets, uint32_t nq)
{
    if (ndev < 1) return 1;
    const int P = swb_layout_parts(n, ndev, min_part);
    if (!qoffsets || nq == 0 || P >= ndev) return P;
    const int R = ndev / P;
    if (nq < SWB_MIN_GROUP_QUERIES * (uint32_t)R) return ndev;
    std::vector<uint32_t> group_of(nq);
    swb_layout_query_groups(qoffsets, nq, R, group_of.data());
    std::vector<uint64_t> load((size_t)R, 0);
    uint64_t total = 0;
    for (uint32_t q = 0; q < nq; ++q) {
        load[group_of[q]] += qoffsets[q + 1] - qoffsets[q];
        total += qoffsets[q + 1] - qoffsets[q];
    }
    const uint64_t heaviest = *std::max_element(load.begin(), load.end());
    return (double)heaviest * R <= 1.02 * (double)total ? P : ndev;
}

extern "C" int swb_group_db_load(swb_group *g, const uint8_t *codes, const uint64_t *offsets, uint32_t n)
{
    if (!g || !offsets) return SWB_ERR_ARG;
    const int nd = (int)g->eng.size();
    const int P = g->parts_forced ? g->parts_forced : swb_layout_parts(n, nd, g->min_part);
    // one length sort for all the parts
    std::vector<uint32_t> order;
    if (swb_sort_by_length(offsets, n, order) != 0) {
        g->err = "bad offsets (decreasing, or a sequence longer than 2^31-16)";
        return SWB_ERR_ARG;
    }
    g->parts = 0;
    const int rc = g->run_all([&](int i) {
        return swb_db_load_sorted(g->eng[i], codes, offsets, n, (uint32_t)(i % P), (uint32_t)P, order.data());
    });
    if (rc != SWB_OK) return rc;
    g->parts = P;
    g->n_total = n;
    return SWB_OK;
}

namespace {

// the sub-batch of one query group: its queries' codes stay where they are, only the offsets are gathered
struct SubBatch {
    std::vector<uint32_t> index;     // positions in the caller's batch
    std::vector<uint8_t> codes;
    std::vector<uint64_t> offsets;
};

void make_sub_batches(const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq, int groups,
                      std::vector<SubBatch> &sub)
{
    std::vector<uint32_t> group_of(nq ? nq : 1);
    swb_layout_query_groups(qoffsets, nq, groups, group_of.data());
    sub.assign((size_t)groups, SubBatch());
    for (int r = 0; r < groups; ++r) sub[r].offsets.push_back(0);
    for (uint32_t q = 0; q < nq; ++q) {
        SubBatch &sb = sub[group_of[q]];
        sb.index.push_back(q);
        sb.codes.insert(sb.codes.end(), qcodes + qoffsets[q], qcodes + qoffsets[q + 1]);
        sb.offsets.push_back(sb.codes.size());
    }
    for (int r = 0; r < groups; ++r)
        if (sub[r].codes.empty()) sub[r].codes.push_back(0);
}

}  // namespace

static void collect_stats(swb_group *g)
{
    swb_stats_t t;
    memset(&t, 0, sizeof t);
    for (size_t i = 0; i < g->eng.size(); ++i) {
        swb_stats_t s;
        swb_stats(g->eng[i], &s);
        t.device_ms = std::max(t.device_ms, s.device_ms);
        t.load_ms = std::max(t.load_ms, s.load_ms);
        t.cells += s.cells;
        t.padded_cells += s.padded_cells;
        t.recomputed_tiles += s.recomputed_tiles;
        t.kernel_launches += s.kernel_launches;
        t.tiles += s.tiles;
        for (int l = 0; l < 6; ++l) t.tiles_by_group[l] += s.tiles_by_group[l];
        if ((int)i < g->parts) {  // every part once
            t.db_residues += s.db_residues;
            t.db_sequences += s.db_sequences;
        }
        t.db_residues_total = s.db_residues_total;
        t.last_k = s.last_k;
        t.sm_count += s.sm_count;
    }
    g->stats = t;
}

extern "C" int swb_group_search_batch(swb_group *g, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq,
                                      int32_t *scores)
{
    if (!g || !qoffsets || !scores || (!qcodes && nq && qoffsets[nq] != qoffsets[0])) return SWB_ERR_ARG;
    if (!g->parts) {
        g->err = "swb_group_search_batch before swb_group_db_load";
        return SWB_ERR_STATE;
    }
    const int nd = (int)g->eng.size(), P = g->parts, R = nd / P;
    std::vector<SubBatch> sub;
    make_sub_batches(qcodes, qoffsets, nq, R, sub);
    // device (p, r): part p, the queries of group r; their vectors are rows of the caller's matrix, which the engines of
    // the P parts fill side by side (disjoint database ids)
    const uint64_t n = g->n_total;
    const int rc = g->run_all([&](int i) {
        const SubBatch &sb = sub[(size_t)(i / P)];
        const uint32_t m = (uint32_t)sb.index.size();
        if (m == 0) return (int)SWB_OK;
        return swb_search_batch_rows(g->eng[i], sb.codes.data(), sb.offsets.data(), m, scores, n, sb.index.data());
    });
    collect_stats(g);
    return rc;
}

extern "C" int swb_group_search_batch_topk(swb_group *g, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq,
                                           uint32_t k, uint32_t *ids, int32_t *top)
{
    if (!g || !qoffsets || !ids || !top || (!qcodes && nq && qoffsets[nq] != qoffsets[0])) return SWB_ERR_ARG;
    if (!g->parts) {
        g->err = "swb_group_search_batch_topk before swb_group_db_load";
        return SWB_ERR_STATE;
    }
    const int nd = (int)g->eng.size(), P = g->parts, R = nd / P;
    std::vector<SubBatch> sub;
    make_sub_batches(qcodes, qoffsets, nq, R, sub);
    std::vector<std::vector<uint32_t> > pid((size_t)nd);
    std::vector<std::vector<int32_t> > ptop((size_t)nd);
    const int rc = g->run_all([&](int i) {
        const SubBatch &sb = sub[(size_t)(i / P)];
        const uint32_t m = (uint32_t)sb.index.size();
        if (m == 0) return (int)SWB_OK;
        pid[(size_t)i].resize((size_t)m * k);
        ptop[(size_t)i].resize((size_t)m * k);
        return swb_search_batch_topk(g->eng[i], sb.codes.data(), sb.offsets.data(), m, k, pid[(size_t)i].data(),
                                     ptop[(size_t)i].data());
    });
    collect_stats(g);
    if (rc != SWB_OK) return rc;
    // host merge of the P per-part lists of every query: score descending, id ascending
    std::vector<std::pair<int32_t, uint32_t> > hits;
    for (int r = 0; r < R; ++r) {
        const SubBatch &sb = sub[(size_t)r];
        for (size_t j = 0; j < sb.index.size(); ++j) {
            hits.clear();
            for (int p = 0; p < P; ++p) {
                const int dev = r * P + p;
                for (uint32_t t = 0; t < k; ++t) {
                    const uint32_t id = pid[(size_t)dev][j * k + t];
                    if (id != 0xffffffffu) hits.push_back(std::make_pair(ptop[(size_t)dev][j * k + t], id));
                }
            }
            std::sort(hits.begin(), hits.end(), [](const std::pair<int32_t, uint32_t> &a, const std::pair<int32_t, uint32_t> &b) {
                return a.first != b.first ? a.first > b.first : a.second < b.second;
            });
            uint32_t *oi = ids + (size_t)sb.index[j] * k;
            int32_t *ot = top + (size_t)sb.index[j] * k;
            for (uint32_t t = 0; t < k; ++t) {
                oi[t] = t < hits.size() ? hits[t].second : 0xffffffffu;
                ot[t] = t < hits.size() ? hits[t].first : -1;
            }
        }
    }
    return SWB_OK;
}

extern "C" int swb_group_stats(const swb_group *g, swb_stats_t *out)
{
    if (!g || !out) return SWB_ERR_ARG;
    *out = g->stats;
    return SWB_OK;
}
