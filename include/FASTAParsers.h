// FASTAParsers.h -- drop-in replacement for the reference's parser header (src/FASTAParsers.h).
//
// Same surface, so code written against the reference compiles unchanged:
//   TILE_SIZE, subject_sequence{id, sequence}, roundUp()                     (reference :12, :16-31)
//   FASTAQuery(path, isQuery), print_buffer(), get_buffer()                  (reference :33-63)
//   FASTADatabase(path): parsedDB, largestSubjectLength, numSubjects, subjectLengthSum   (reference :65-138)
//   `using namespace std;` is exported on purpose: the reference's callers rely on it
//   (test/swissprot_tests.cpp:21-26, 56 use bare ifstream / map / cout).
//
// Same parsing rules (checked against the reference header itself by tests/test_parsers.py):
//   * query: the first line is dropped, all other lines are concatenated
//   * database: a line whose first character is '>' starts a record; text before the first '>' is dropped
//     when a '>' follows, and forms the only record (id -1) when the file has no '>' at all; a missing or
//     empty file gives one empty record with id -1
//   * ids are 0-based record ordinals; header text is discarded; '\r' and blank lines are kept as they are
//   * every sequence is padded with '/' to a multiple of TILE_SIZE and filed under its padded length
//
// Different inside: the file is read with one bulk read and cut in place (the reference pays a getline,
// two string copies and a map lookup per line), and 64-bit counters sit beside the int ones, which
// overflow at UniProt scale (reference :69-71).
#ifndef FASTAPARSERS_H
#define FASTAPARSERS_H

#include <string>
#include <iostream>
#include <fstream>

#include <map>
#include <vector>

#define TILE_SIZE 8

using namespace std;

struct subject_sequence {
    int id;
    string sequence;
};

static inline int roundUp(int numToRound, int multiple)
{
    if (multiple == 0) return numToRound;
    const int over = numToRound % multiple;
    return over == 0 ? numToRound : numToRound + (multiple - over);
}

namespace swb_detail {
// whole file as one string; a file that cannot be opened reads as empty
inline string slurp(const std::string &path)
{
    string data;
    ifstream in(path.c_str(), ios::in | ios::binary);
    if (!in) return data;
    in.seekg(0, ios::end);
    const streamoff size = in.tellg();
    if (size > 0) {
        data.resize((size_t)size);
        in.seekg(0, ios::beg);
        in.read(&data[0], size);
        data.resize((size_t)in.gcount());
    }
    return data;
}
// calls f(begin, end) for every line the way std::getline sees them: split at '\n', no terminator
// needed on the last line, nothing after a trailing '\n'
template <class F> inline void for_each_line(const string &data, F f)
{
    size_t pos = 0;
    const size_t n = data.size();
    while (pos < n) {
        size_t nl = data.find('\n', pos);
        if (nl == string::npos) nl = n;
        f(pos, nl);
        pos = nl + 1;
    }
}
}  // namespace swb_detail

class FASTAQuery {
private:
    bool isQuery;
    string buffer;

public:
    FASTAQuery(std::string filepath, bool _isQuery) : isQuery(_isQuery)
    {
        const string data = swb_detail::slurp(filepath);
        bool header = true;
        buffer.reserve(data.size());
        swb_detail::for_each_line(data, [&](size_t b, size_t e) {
            if (header) header = false;
            else buffer.append(data, b, e - b);
        });
    }

    ~FASTAQuery() {}

    void print_buffer() { cout << buffer << endl; }

    string get_buffer() { return buffer; }
};

class FASTADatabase {
public:
    // key is the PADDED sequence length, value the records of that length in file order
    map<int, vector<subject_sequence> > parsedDB;
    int largestSubjectLength;
    int numSubjects;
    int subjectLengthSum;
    // same totals without the int overflow of the three fields above
    long long subjectLengthSum64;
    long long numSubjects64;

    FASTADatabase(std::string filepath)
        : largestSubjectLength(0), numSubjects(0), subjectLengthSum(0), subjectLengthSum64(0), numSubjects64(0)
    {
        const string data = swb_detail::slurp(filepath);
        string current;
        int id = -1;
        bool seen_header = false;
        swb_detail::for_each_line(data, [&](size_t b, size_t e) {
            if (e > b && data[b] == '>') {
                if (seen_header) file_record(id, current);
                seen_header = true;
                current.clear();
                ++id;
            } else {
                current.append(data, b, e - b);
            }
        });
        file_record(id, current);
    }

private:
    void file_record(int id, string &seq)
    {
        const size_t padded = (seq.size() + TILE_SIZE - 1) / TILE_SIZE * TILE_SIZE;
        seq.resize(padded, '/');
        vector<subject_sequence> &bucket = parsedDB[(int)padded];
        bucket.push_back(subject_sequence());
        bucket.back().id = id;
        bucket.back().sequence.swap(seq);
        seq.clear();
        subjectLengthSum += (int)padded;
        subjectLengthSum64 += (long long)padded;
        if ((int)padded > largestSubjectLength) largestSubjectLength = (int)padded;
        ++numSubjects;
        ++numSubjects64;
    }
};

#endif /* FASTAPARSERS_H */
