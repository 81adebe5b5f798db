// Command line of the reference (src/main.cpp:19-74) without Boost: same flags (--help, --query, --db, also
// as --flag=value and as unambiguous prefixes like boost::program_options accepts), same stdout text, same
// exit codes (1 for help / no arguments / a missing required option, 0 after a scan). An unknown option is
// an uncaught exception, as in the reference (main.cpp:38 only catches po::required_option).
// Extension that leaves the default output untouched: --gpus N scans on the first N visible GPUs (default: all of them;
// the database is residue-sharded across the devices inside this one process, include/swb.h swb_group_*).
#include <stdlib.h>
#include <sys/time.h>
#include <stdexcept>
#include <string>
#include <vector>

#include <algorithm>

#include "FASTAParsers.h"
#include "SWSolver.h"
#include "swb.h"

static double wall_seconds()
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return (double)tv.tv_usec / 1000000 + tv.tv_sec;
}

static void usage()
{
    // boost::program_options rendering of the reference's options_description (main.cpp:22, 26-29)
    cout << "Smith-Waterman CUDA Usage:\n"
            "  --help                Display this help message\n"
            "  --query arg           Path to query file (required)\n"
            "  --db arg              Path to database file (required)\n";
}

// 0 help, 1 query, 2 db, 3 gpus; throws on unknown / ambiguous names
static int match_option(const std::string &name)
{
    static const char *const names[4] = {"help", "query", "db", "gpus"};
    int hit = -1;
    for (int i = 0; i < 4; ++i) {
        const std::string full(names[i]);
        if (full == name) return i;
        if (!name.empty() && full.compare(0, name.size(), name) == 0) {
            if (hit >= 0) throw std::runtime_error("option '--" + name + "' is ambiguous");
            hit = i;
        }
    }
    if (hit < 0) throw std::runtime_error("unrecognised option '--" + name + "'");
    return hit;
}

// Extension: --db <file>.swbdb takes an encoded database written by bin/swb_mkdb (include/swb.h, swb_dbfile_*)
// instead of parsing text. Same stdout as the text path: ids, result order (descending padded length, file order
// inside a bucket, SWSolver.cu:383-390) and the METRICS block are recomputed from the stored offsets.
static int scan_encoded_db(const std::string &querypath, const std::string &datapath, double time_start)
{
    FASTAQuery query(querypath, true);
    cout << "Input buffer:";
    query.print_buffer();
    cout << endl;
    const string q = query.get_buffer();
    swb_dbfile *dbf = nullptr;
    if (swb_dbfile_open(datapath.c_str(), &dbf) != SWB_OK) throw std::runtime_error("cannot open encoded database " + datapath);
    const uint32_t n = swb_dbfile_count(dbf);
    const uint64_t *offsets = swb_dbfile_offsets(dbf);
    const int first_id = swb_dbfile_first_id(dbf);
    swb_group *eng = nullptr;
    if (swb_group_create_env(&eng) != SWB_OK)
        throw std::runtime_error(std::string("swb_group_create: ") + swb_group_last_error(nullptr));
    std::vector<uint8_t> qcodes(q.size() ? q.size() : 1);
    swb_encode(SWB_SCORING_BLOSUM50_REF, q.data(), q.size(), qcodes.data());
    std::vector<int32_t> scores(n ? n : 1);
    const uint64_t qoffs[2] = {0, q.size()};
    if (swb_group_set_option(eng, "db_parts", swb_group_size(eng)) != SWB_OK ||
        swb_group_db_load(eng, swb_dbfile_codes(dbf), offsets, n) != SWB_OK ||
        swb_group_search_batch(eng, qcodes.data(), qoffs, 1, scores.data()) != SWB_OK)
        throw std::runtime_error(std::string("scan failed: ") + swb_group_last_error(eng));
    std::vector<uint32_t> order(n);
    long long padded_sum = 0;
    std::vector<uint64_t> padded(n);
    for (uint32_t i = 0; i < n; ++i) {
        order[i] = i;
        padded[i] = (offsets[i + 1] - offsets[i] + TILE_SIZE - 1) / TILE_SIZE * TILE_SIZE;
        padded_sum += (long long)padded[i];
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return padded[a] > padded[b]; });
    std::string lines;
    lines.reserve((size_t)n * 12);
    for (uint32_t k = 0; k < n; ++k) {
        lines += std::to_string((long long)order[k] + first_id);
        lines += ':';
        lines += std::to_string(scores[order[k]]);
        lines += '\n';
    }
    cout << lines;
    const double seconds_elapsed = wall_seconds() - time_start;
    cout << std::string(80, '=') << endl;
    cout << "METRICS:" << endl;
    cout << "Query length: " << q.length() << " chars." << endl;
    cout << "Num subjects: " << n << endl;
    cout << "Sum of DB length: " << (int)padded_sum << " chars." << endl;
    cout << "Time elapsed: " << seconds_elapsed << " seconds." << endl;
    cout << "Performance: " << 1E-9 * ((double)q.length() * (double)padded_sum) / seconds_elapsed << " GCUPS." << endl;
    swb_group_destroy(eng);
    swb_dbfile_close(dbf);
    return 0;
}

int main(int argc, char *argv[])
{
    const double time_start = wall_seconds();
    bool help = false, have_query = false, have_db = false;
    std::string querypath, datapath;
    for (int i = 1; i < argc; ++i) {
        const std::string arg(argv[i]);
        if (arg.size() < 3 || arg.compare(0, 2, "--") != 0) throw std::runtime_error("too many positional options: " + arg);
        const size_t eq = arg.find('=');
        const int opt = match_option(arg.substr(2, eq == std::string::npos ? std::string::npos : eq - 2));
        if (opt == 0) {
            help = true;
            continue;
        }
        std::string value;
        if (eq != std::string::npos) value = arg.substr(eq + 1);
        else if (i + 1 < argc) value = argv[++i];
        else throw std::runtime_error("the required argument for option '" + arg + "' is missing");
        if (opt == 1) { querypath = value; have_query = true; }
        else if (opt == 2) { datapath = value; have_db = true; }
        else setenv("SWB_GPUS", value.c_str(), 1);  // read where the engine group is created
    }
    if (help || argc <= 1 || !have_query || !have_db) {
        usage();
        return 1;
    }

    if (datapath.size() > 6 && datapath.compare(datapath.size() - 6, 6, ".swbdb") == 0)
        return scan_encoded_db(querypath, datapath, time_start);

    FASTAQuery query(querypath, true);
    cout << "Input buffer:";
    query.print_buffer();
    cout << endl;
    string querySequence = query.get_buffer();

    FASTADatabase db(datapath);

    vector<seqid_score> result;
    result.reserve(db.numSubjects64 > 600000 ? (size_t)db.numSubjects64 : 600000);
    smith_waterman_cuda(query, db, result);

    // same text as main.cpp:58-60, one buffered write instead of a flush per line
    std::string lines;
    lines.reserve(result.size() * 12);
    for (vector<seqid_score>::iterator it = result.begin(); it != result.end(); ++it) {
        lines += std::to_string(it->first);
        lines += ':';
        lines += std::to_string(it->second);
        lines += '\n';
    }
    cout << lines;

    const double seconds_elapsed = wall_seconds() - time_start;
    cout << std::string(80, '=') << endl;
    cout << "METRICS:" << endl;
    cout << "Query length: " << querySequence.length() << " chars." << endl;
    cout << "Num subjects: " << db.numSubjects << endl;
    cout << "Sum of DB length: " << db.subjectLengthSum << " chars." << endl;
    cout << "Time elapsed: " << seconds_elapsed << " seconds." << endl;
    cout << "Performance: " << 1E-9 * ((double)querySequence.length() * (double)db.subjectLengthSum64) / seconds_elapsed
         << " GCUPS." << endl;
    return 0;
}
