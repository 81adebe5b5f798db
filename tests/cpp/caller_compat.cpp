// A caller written the way the reference's own callers are (src/main.cpp:44-60, test/swissprot_tests.cpp:20-75):
// bare ifstream / map / cout from the header's `using namespace std`, the public fields of FASTADatabase, the
// seqid_score typedef and smith_waterman_cuda(). It must compile unchanged against include/ of this repo.
#include <sstream>
#include "FASTAParsers.h"
#include "SWSolver.h"

map<int, int> parse_golden_results(std::string filepath)
{
    ifstream filestream;
    filestream.open(filepath.c_str());
    map<int, int> parsed_results;
    string tmp;
    int idx = 0, score;
    while (getline(filestream, tmp)) {
        std::istringstream(tmp) >> score;
        parsed_results[idx++] = score;
    }
    return parsed_results;
}

int run(std::string querypath, std::string dbpath, std::string refpath)
{
    FASTAQuery query(querypath, true);
    FASTADatabase db(dbpath);
    std::vector<seqid_score> result;
    result.reserve(600000);
    smith_waterman_cuda(query, db, result);
    map<int, int> reference_results = parse_golden_results(refpath);
    int bad = 0;
    for (vector<seqid_score>::iterator it = result.begin(); it != result.end(); ++it)
        if ((*it).second != reference_results[(*it).first]) ++bad;
    cout << "Query " << querypath << " length " << query.get_buffer().length() << " subjects " << db.numSubjects
         << " residues " << db.subjectLengthSum << " largest " << db.largestSubjectLength << " buckets "
         << db.parsedDB.size() << " roundUp " << roundUp(13, TILE_SIZE) << " mismatches " << bad << endl;
    for (map<int, vector<subject_sequence> >::reverse_iterator it = db.parsedDB.rbegin(); it != db.parsedDB.rend(); ++it)
        for (size_t i = 0; i < it->second.size(); ++i)
            if (it->second[i].sequence.length() != (size_t)it->first) return -1;
    return bad;
}

int main(int argc, char **argv)
{
    if (argc != 4) return 2;
    return run(argv[1], argv[2], argv[3]) == 0 ? 0 : 1;
}
