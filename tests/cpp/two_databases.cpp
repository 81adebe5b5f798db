// smith_waterman_cuda() called for two different databases in one process (the shim caches the packed database per
// FASTADatabase object and content): database 2 holds the records of database 1 in reverse order, so its score for id k
// must be golden[n-1-k].
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>
#include "FASTAParsers.h"
#include "SWSolver.h"

static int check(FASTAQuery &query, const char *dbpath, const std::vector<int> &gold, bool reversed)
{
    FASTADatabase db(dbpath);
    std::vector<seqid_score> result;
    smith_waterman_cuda(query, db, result);
    int bad = 0;
    const int n = (int)gold.size();
    if ((int)result.size() != n) return 1000000;
    for (size_t k = 0; k < result.size(); ++k) {
        const int id = result[k].first;
        if (result[k].second != gold[reversed ? n - 1 - id : id]) ++bad;
    }
    return bad;
}

int main(int argc, char **argv)
{
    if (argc != 5) return 2;
    FASTAQuery query(argv[1], true);
    std::vector<int> gold;
    {
        std::ifstream f(argv[4]);
        int v;
        while (f >> v) gold.push_back(v);
    }
    int bad = 0;
    for (int round = 0; round < 2; ++round) {
        bad += check(query, argv[2], gold, false);
        bad += check(query, argv[3], gold, true);
    }
    printf("mismatches %d\n", bad);
    return bad == 0 ? 0 : 1;
}
