// smith_waterman_cuda() of the reference (src/SWSolver.h:9, src/SWSolver.cu:266-404) on top of the C ABI.
// The reference encodes, packs, uploads, launches and gathers inside this one call, every call; here the
// packed database is loaded once per FASTADatabase object and the call is: encode query -> swb_search.
#include <stdexcept>
#include <string>
#include <vector>

#include "SWSolver.h"
#include "swb.h"

namespace {

struct SolverState {
    swb_engine *engine = nullptr;
    const FASTADatabase *db = nullptr;
    unsigned long long fingerprint = 0;
    std::vector<int> ids;  // ids in the reference's result order
    ~SolverState()
    {
        if (engine) swb_destroy(engine);
    }
};

SolverState &state()
{
    static SolverState s;
    return s;
}

void check(int rc, swb_engine *e, const char *what)
{
    if (rc != SWB_OK) throw std::runtime_error(std::string(what) + ": " + swb_last_error(e));
}

}  // namespace

void smith_waterman_cuda(FASTAQuery &query, FASTADatabase &db, std::vector<seqid_score> &result)
{
    SolverState &st = state();
    if (!st.engine) {
        int rc = swb_create(&st.engine, 0);
        if (rc != SWB_OK) throw std::runtime_error(std::string("swb_create: ") + swb_last_error(nullptr));
        check(swb_set_scoring_preset(st.engine, SWB_SCORING_BLOSUM50_REF), st.engine, "swb_set_scoring_preset");
    }
    // content fingerprint (ids, lengths and the ends of every sequence): a different database at the same address
    // must not hit the cache
    unsigned long long fp = 1469598103934665603ull;
    {
        auto mix = [&fp](unsigned long long v) { fp = (fp ^ v) * 1099511628211ull; };
        mix((unsigned long long)db.numSubjects64);
        mix((unsigned long long)db.subjectLengthSum64);
        for (map<int, vector<subject_sequence> >::const_iterator it = db.parsedDB.begin(); it != db.parsedDB.end(); ++it)
            for (size_t i = 0; i < it->second.size(); ++i) {
                const string &q = it->second[i].sequence;
                mix((unsigned long long)(unsigned)it->second[i].id);
                mix(q.size());
                for (size_t k = 0; k < q.size() && k < 12; ++k) mix((unsigned char)q[k]);
                for (size_t k = q.size() > 12 ? q.size() - 12 : 0; k < q.size(); ++k) mix((unsigned char)q[k]);
            }
    }
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
        check(swb_db_load(st.engine, codes.data(), offsets.data(), (uint32_t)st.ids.size(), 0, 1), st.engine,
              "swb_db_load");
        st.db = &db;
        st.fingerprint = fp;
    }
    const string q = query.get_buffer();
    std::vector<uint8_t> qcodes(q.size() ? q.size() : 1);
    swb_encode(SWB_SCORING_BLOSUM50_REF, q.data(), q.size(), qcodes.data());
    std::vector<int32_t> scores(st.ids.size() ? st.ids.size() : 1);
    check(swb_search(st.engine, qcodes.data(), (uint32_t)q.size(), scores.data()), st.engine, "swb_search");
    for (size_t k = 0; k < st.ids.size(); ++k) result.push_back(std::make_pair(st.ids[k], (int)scores[k]));
}
