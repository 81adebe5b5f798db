// Command line of the reference (src/main.cpp:19-74) without Boost: same flags (--help, --query, --db, also
// as --flag=value and as unambiguous prefixes like boost::program_options accepts), same stdout text, same
// exit codes (1 for help / no arguments / a missing required option, 0 after a scan). An unknown option is
// an uncaught exception, as in the reference (main.cpp:38 only catches po::required_option).
#include <sys/time.h>
#include <stdexcept>
#include <string>
#include <vector>

#include "FASTAParsers.h"
#include "SWSolver.h"

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

// 0 help, 1 query, 2 db; throws on unknown / ambiguous names
static int match_option(const std::string &name)
{
    static const char *const names[3] = {"help", "query", "db"};
    int hit = -1;
    for (int i = 0; i < 3; ++i) {
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
        else { datapath = value; have_db = true; }
    }
    if (help || argc <= 1 || !have_query || !have_db) {
        usage();
        return 1;
    }

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
