"""Drop-in surface: include/FASTAParsers.h (C++) and the Python mirror must parse exactly like the reference's
own header (expectations in tests/golden/parser_expect were printed by oracle/_ref/ref_parser_probe, i.e.
by /root/reference/src/FASTAParsers.h itself), and the Boost-free command line must keep the reference's
usage text and exit codes (src/main.cpp:26-41)."""
import os
import subprocess

import pytest

from conftest import GOLDEN, PKG, ROOT

EXPECT = os.path.join(GOLDEN, "parser_expect")
DB_CASES = ["uniprot_subset.fasta", "test.dat", "tricky.fasta", "empty.fasta", "noeol.fasta"]
QUERY_CASES = [("queries/P02232.fasta", "P02232.fasta"), ("tricky.fasta", "tricky.fasta"), ("empty.fasta", "empty.fasta"),
               ("noeol.fasta", "noeol.fasta")]


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    """oracle/ref_parser_probe.cpp compiled against include/ of THIS repo instead of the reference's src/"""
    out = str(tmp_path_factory.mktemp("probe") / "our_parser_probe")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", out,
                    os.path.join(ROOT, "oracle", "ref_parser_probe.cpp")], check=True)
    return out


def _dump_py(db):
    lines = ["numSubjects %d" % db.numSubjects, "largestSubjectLength %d" % db.largestSubjectLength,
             "subjectLengthSum %d" % db.subjectLengthSum, "buckets %d" % len(db.parsedDB)]
    for length in sorted(db.parsedDB, reverse=True):
        for sid, seq in db.parsedDB[length]:
            lines.append("%d %d %s" % (sid, length, seq))
    return ("\n".join(lines) + "\n").encode("latin-1")


@pytest.mark.parametrize("name", DB_CASES)
def test_cpp_database_parser(probe, name):
    got = subprocess.run([probe, "db", os.path.join(GOLDEN, name)], capture_output=True, check=True).stdout
    assert got == open(os.path.join(EXPECT, name + ".db.txt"), "rb").read()


def test_cpp_missing_file_and_queries(probe):
    got = subprocess.run([probe, "db", "/nonexistent/path"], capture_output=True, check=True).stdout
    assert got == open(os.path.join(EXPECT, "nonexistent.db.txt"), "rb").read()
    for path, key in QUERY_CASES:
        got = subprocess.run([probe, "query", os.path.join(GOLDEN, path)], capture_output=True, check=True).stdout
        assert got == open(os.path.join(EXPECT, key + ".query.txt"), "rb").read(), path


@pytest.mark.parametrize("name", DB_CASES)
def test_python_database_mirror(swb, name):
    assert _dump_py(swb.FASTADatabase(os.path.join(GOLDEN, name))) == open(os.path.join(EXPECT, name + ".db.txt"), "rb").read()


def test_python_query_mirror_and_missing_file(swb):
    for path, key in QUERY_CASES:
        q = swb.FASTAQuery(os.path.join(GOLDEN, path))
        assert (q.get_buffer() + "\n").encode("latin-1") == open(os.path.join(EXPECT, key + ".query.txt"), "rb").read()
    assert _dump_py(swb.FASTADatabase("/nonexistent/path")) == open(os.path.join(EXPECT, "nonexistent.db.txt"), "rb").read()
    assert swb.round_up(13, 8) == 16 and swb.round_up(16, 8) == 16 and swb.round_up(5, 0) == 5


def test_subset_facts(swb):
    """111 subjects, largest padded 1168, padded sum 26728, 50 buckets; result order starts 56, 34, 13"""
    db = swb.FASTADatabase(os.path.join(GOLDEN, "uniprot_subset.fasta"))
    assert (db.numSubjects, db.largestSubjectLength, db.subjectLengthSum, len(db.parsedDB)) == (111, 1168, 26728, 50)
    assert [sid for sid, _ in db.ordered()[:3]] == [56, 34, 13]


def test_reference_style_caller_compiles():
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "caller_compat.cpp")], check=True)


def test_cli_usage_and_exit_codes(swb):
    """no arguments / --help / a missing required option -> usage on stdout, exit 1 (main.cpp:34-41)"""
    main = os.path.join(ROOT, PKG, "bin", "main")
    if not os.path.exists(main):
        swb.build()
    usage = ("Smith-Waterman CUDA Usage:\n  --help                Display this help message\n"
             "  --query arg           Path to query file (required)\n  --db arg              Path to database file (required)\n")
    for args in ([], ["--help"], ["--query", "x.fasta"], ["--db=y.fasta"], ["--help", "--query", "a", "--db", "b"]):
        r = subprocess.run([main] + args, capture_output=True, text=True)
        assert r.returncode == 1 and r.stdout == usage, args
    r = subprocess.run([main, "--bogus"], capture_output=True, text=True)  # uncaught exception, like the reference
    assert r.returncode != 0 and r.returncode != 1
