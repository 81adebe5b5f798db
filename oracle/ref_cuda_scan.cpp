// Test infrastructure: a Boost-free driver around the REFERENCE's own smith_waterman_cuda
// (SWSolver.cu compiled unmodified from /root/reference/src, see Makefile).  Runs on a GPU box only.
//   ref_cuda_scan <query.fasta> <db.fasta> [repeat]
// stdout: one "id:score" line per subject in the reference's result order (main.cpp:58-60), then
// "#TIME solver_s=<best solver-call seconds over repeats> qlen=<n> padded_residues=<n> nsubj=<n>".
// Only meaningful for padded query length <= 1024 (SWSolver.cu:85).
#include <sys/time.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "FASTAParsers.h"
#include "SWSolver.h"

static double now()
{
    timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_usec * 1e-6 + tv.tv_sec;
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s query db [repeat]\n", argv[0]); return 2; }
    int rep = argc > 3 ? atoi(argv[3]) : 1;
    FASTAQuery query(argv[1], true);
    FASTADatabase db(argv[2]);
    std::vector<seqid_score> result;
    double best = 1e30;
    for (int r = 0; r < rep; ++r) {
        result.clear();
        result.reserve(db.numSubjects);
        double t0 = now();
        smith_waterman_cuda(query, db, result);
        double t1 = now();
        if (t1 - t0 < best) best = t1 - t0;
    }
    for (size_t k = 0; k < result.size(); ++k) printf("%d:%d\n", result[k].first, result[k].second);
    printf("#TIME solver_s=%.6f qlen=%zu padded_residues=%d nsubj=%d\n", best,
           query.get_buffer().length(), db.subjectLengthSum, db.numSubjects);
    return 0;
}
