// smith_waterman_cuda() of the reference (src/SWSolver.h:9, src/SWSolver.cu:266-404) on top of the C ABI.
// The reference encodes, packs, uploads, launches and gathers inside this one call, every call, on one GPU; here the
// packed database is loaded once per FASTADatabase content onto EVERY visible GPU (engine group, swb.h: one engine, host
// thread and stream set per device, the database residue-sharded across them) and the call is: encode query -> scan on
// all devices -> scores scattered into one vector in the reference's result order.
//
// Devices: all visible ones; SWB_GPUS=<n> limits the count, SWB_DEVICES=<i,j,...> names them (an index may repeat: two
// engines on one device, which is how the multi-device path is tested on a one-GPU box).
#include <stdlib.h>
#include <string.h>
#include <stdexcept>
#include <string>
#include <vector>

#include "SWSolver.h"
#include "swb.h"

namespace {

struct SolverState {
    swb_group *group = nullptr;
    const FASTADatabase *db = nullptr;
    unsigned long long fingerprint = 0;
    std::vector<int> ids;  // ids in the reference's result order
    ~SolverState()
    {
        if (group) swb_group_destroy(group);
    }
};

SolverState &state()
{
    static SolverState s;
    return s;
}

void check(int rc, swb_group *g, const char *what)
{
    if (rc != SWB_OK) throw std::runtime_error(std::string(what) + ": " + swb_group_last_error(g));
}

// Content fingerprint of the whole database: ids, lengths and EVERY residue (8 bytes per step). parsedDB is a public,
// mutable member, so a cache keyed on less could score a stale GPU copy after an edit in the middle of a sequence; the
// reference re-packs on every call and cannot go stale. ~20 ms per 100 MB; SWB_TRUST_DB=1 hashes only ids, lengths and
// the ends of every sequence for callers that never modify a database in place.
unsigned long long fingerprint_of(const FASTADatabase &db)
{
    const bool trust = getenv("SWB_TRUST_DB") != nullptr;
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&h](unsigned long long v) {
        h ^= v;
        h *= 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
    };
    mix((unsigned long long)db.numSubjects64);
    mix((unsigned long long)db.subjectLengthSum64);
    for (map<int, vector<subject_sequence> >::const_iterator it = db.parsedDB.begin(); it != db.parsedDB.end(); ++it)
        for (size_t i = 0; i < it->second.size(); ++i) {
            const string &q = it->second[i].sequence;
            mix((unsigned long long)(unsigned)it->second[i].id);
            mix(q.size());
            const char *d = q.data();
            const size_t n = q.size();
            if (trust && n > 32) {
                unsigned long long a, b;
                memcpy(&a, d, 8);
                memcpy(&b, d + n - 8, 8);
                mix(a);
                mix(b);
                continue;
            }
            size_t k = 0;
            for (; k + 8 <= n; k += 8) {
                unsigned long long w;
                memcpy(&w, d + k, 8);
                mix(w);
            }
            unsigned long long w = 0;
            memcpy(&w, d + k, n - k);
            mix(w);
        }
    return h;
}

}  // namespace

void smith_waterman_cuda(FASTAQuery &query, FASTADatabase &db, std::vector<seqid_score> &result)
{
    SolverState &st = state();
    if (!st.group) {
        int rc = swb_group_create_env(&st.group);
        if (rc != SWB_OK) throw std::runtime_error(std::string("swb_group_create: ") + swb_group_last_error(nullptr));
        check(swb_group_set_scoring_preset(st.group, SWB_SCORING_BLOSUM50_REF), st.group, "swb_group_set_scoring_preset");
        // this entry point scans ONE query per call: every device takes a part of the database
        check(swb_group_set_option(st.group, "db_parts", swb_group_size(st.group)), st.group, "swb_group_set_option");
    }
    const unsigned long long fp = fingerprint_of(db);
    if (st.db != &db || st.fingerprint != fp) {
        // database in the order the reference reports results: parsedDB.rbegin() .. rend() (SWSolver.cu:383-390)
        std::vector<uint8_t> codes;
        std::vector<uint64_t> offsets;
        codes.reserve((size_t)db.subjectLengthSum64);
        offsets.reserve((size_t)db.numSubjects64 + 1);
        st.ids.clear();
        st.ids.reserve((size_t)db.numSubjects64);
        offsets.push_back(0);
        for (map<int, vector<subject_sequence> >::reverse_iterator it = db.parsedDB.rbegin(); it != db.parsedDB.rend();
             ++it) {
            for (size_t i = 0; i < it->second.size(); ++i) {
                const string &s = it->second[i].sequence;
                const size_t at = codes.size();
                codes.resize(at + s.size());
                swb_encode(SWB_SCORING_BLOSUM50_REF, s.data(), s.size(), codes.data() + at);
                offsets.push_back(codes.size());
                st.ids.push_back(it->second[i].id);
            }
        }
        st.db = nullptr;
        check(swb_group_db_load(st.group, codes.data(), offsets.data(), (uint32_t)st.ids.size()), st.group,
              "swb_group_db_load");
        st.db = &db;
        st.fingerprint = fp;
    }
    const string q = query.get_buffer();
    std::vector<uint8_t> qcodes(q.size() ? q.size() : 1);
    swb_encode(SWB_SCORING_BLOSUM50_REF, q.data(), q.size(), qcodes.data());
    std::vector<int32_t> scores(st.ids.size() ? st.ids.size() : 1);
    const uint64_t qoffs[2] = {0, q.size()};
    check(swb_group_search_batch(st.group, qcodes.data(), qoffs, 1, scores.data()), st.group, "swb_group_search_batch");
    for (size_t k = 0; k < st.ids.size(); ++k) result.push_back(std::make_pair(st.ids[k], (int)scores[k]));
}
