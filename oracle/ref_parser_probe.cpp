// Test infrastructure: prints what the REFERENCE's FASTAParsers.h (included from /root/reference/src
// at build time, see Makefile) parses out of a file, so tests can compare include/FASTAParsers.h of
// this repo against it field by field.
//   ref_parser_probe db <path>     -> counters, then one "id paddedLen sequence" line per subject in
//                                     the order SWSolver.cu:383-390 reports them (parsedDB.rbegin())
//   ref_parser_probe query <path>  -> the query buffer
#include <cstdio>
#include <cstring>
#include "FASTAParsers.h"

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s db|query path\n", argv[0]); return 2; }
    if (!strcmp(argv[1], "query")) {
        FASTAQuery q(argv[2], true);
        printf("%s\n", q.get_buffer().c_str());
        return 0;
    }
    FASTADatabase db(argv[2]);
    printf("numSubjects %d\nlargestSubjectLength %d\nsubjectLengthSum %d\nbuckets %zu\n",
           db.numSubjects, db.largestSubjectLength, db.subjectLengthSum, db.parsedDB.size());
    for (map<int, vector<subject_sequence> >::reverse_iterator it = db.parsedDB.rbegin();
         it != db.parsedDB.rend(); ++it)
        for (size_t i = 0; i < it->second.size(); ++i)
            printf("%d %d %s\n", it->second[i].id, it->first, it->second[i].sequence.c_str());
    return 0;
}
