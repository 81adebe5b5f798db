// SWSolver.h -- drop-in replacement for the reference's solver header (src/SWSolver.h:1-11).
// Same typedef, same entry point, same result contract (src/SWSolver.cu:383-390): one (id, score) pair per
// database sequence is APPENDED to `result`, longest padded length first, file order inside a length
// bucket; scoring = BLOSUM50 with the '*' row/column zeroed, linear gap 2 (src/SWSolver.cu:7, 54-81).
// Differences: scores are exact int32 (the reference wraps at 32767, SWSolver.cu:263), queries of any
// length are valid (the reference stops at 1024 rows, SWSolver.cu:85), the packed database stays on the GPU
// between calls with the same FASTADatabase, and a CUDA failure throws std::runtime_error instead of
// returning garbage. Implemented in host/SWSolver.cpp on top of the C ABI in swb.h.
#ifndef SWSOLVER_H
#define SWSOLVER_H

#include <vector>
#include "FASTAParsers.h"

typedef std::pair<int, int> seqid_score;

void smith_waterman_cuda(FASTAQuery &query, FASTADatabase &db, std::vector<seqid_score> &result);

#endif /* SWSOLVER_H */
