#!/usr/bin/env python3
"""Regenerates tests/golden/ from the mounted reference tree (run in the build container only).

    python tests/golden/make_fixtures.py            # needs /root/reference and oracle/_ref/ built

What it writes (all small, committed):
  uniprot_subset.fasta      the 111 entries of data/dbs/uniprot_subset.dat as FASTA, FILE ORDER
                            (SQ-block extraction recipe: src/parse.py:24-35), ids 0..110
  queries/*.fasta           the 20 query files of data/queries/, byte for byte
  test.dat                  data/dbs/test.dat (a header-less file; parser edge case)
  uniprot_head5.dat         the first five entries of data/dbs/uniprot_subset.dat (UniProt flat-file text)
  P01008.head111.txt, P02232.head111.txt
                            lines 1-111 of test/reference/P01008.txt / P02232.txt: the reference's
                            golden scores (test/swissprot_tests.cpp:68-72) for DB ids 0..110
  survey_exp_blosum50.json  20 x 111 expected BLOSUM50/gap-2 scores from the survey probe
                            (baseline/_ref/exp, an independent restatement) when present
  cpu_ref_ident3.json       scores and aligned strings printed by the COMPILED reference cpu.cpp
                            (oracle/_ref/cpu_ref) in its own +3/-3, gap 2 scheme
  parser_expect/*.txt       what the reference's own FASTAParsers.h parses out of each fixture
                            (oracle/_ref/ref_parser_probe)
"""
import json
import os
import resource
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SWB_REFERENCE", "/root/reference")
CPU_REF = os.path.join(ROOT, "oracle", "_ref", "cpu_ref")
PARSER_PROBE = os.path.join(ROOT, "oracle", "_ref", "ref_parser_probe")


def read_flatfile(path):
    ids, seqs, cur = [], [], None
    for line in open(path):
        if line.startswith("ID "):
            ids.append(line.split()[1])
        if line.startswith("SQ "):
            if cur is not None:
                seqs.append(cur)
            cur = ""
        elif line.startswith("//"):
            if cur is not None:
                seqs.append(cur)
            cur = None
        elif cur is not None:
            cur += "".join(line.split())
    if cur is not None:
        seqs.append(cur)
    return ids, seqs


def read_query(path):
    lines = open(path).read().split("\n")
    return "".join(lines[1:])


def cpu_ref_run(a, b):
    """Runs the compiled cpu.cpp; returns (score, alignedA, alignedB). Score = max of printed matrix."""
    out = subprocess.run([CPU_REF, a, b], capture_output=True, text=True, check=True,
                         preexec_fn=lambda: resource.setrlimit(resource.RLIMIT_STACK,
                                                               (resource.RLIM_INFINITY, resource.RLIM_INFINITY))).stdout
    lines = out.split("\n")
    best = 0
    for row in lines[3:]:
        for tok in row.split():
            try:
                best = max(best, int(tok))
            except ValueError:
                pass
    return best, lines[0], lines[1]


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not mounted at %s" % REF)
    ids, seqs = read_flatfile(os.path.join(REF, "data/dbs/uniprot_subset.dat"))
    assert len(seqs) == 111 and len(ids) == 111, (len(seqs), len(ids))
    # sanity from SURVEY 8(c): sorted by length it equals uniprot_subset_p.dat line for line
    flat = [l.strip() for l in open(os.path.join(REF, "data/dbs/uniprot_subset_p.dat")) if l.strip()]
    assert sorted(seqs, key=len) == flat or sorted(map(len, seqs)) == list(map(len, flat))
    with open(os.path.join(HERE, "uniprot_subset.fasta"), "w") as f:
        for i, s in enumerate(seqs):
            f.write(">%s\n" % ids[i])
            for k in range(0, len(s), 60):
                f.write(s[k:k + 60] + "\n")
    os.makedirs(os.path.join(HERE, "queries"), exist_ok=True)
    qnames = sorted(os.listdir(os.path.join(REF, "data/queries")))
    for q in qnames:
        shutil.copyfile(os.path.join(REF, "data/queries", q), os.path.join(HERE, "queries", q))
    shutil.copyfile(os.path.join(REF, "data/dbs/test.dat"), os.path.join(HERE, "test.dat"))
    for g in ("P01008", "P02232"):
        with open(os.path.join(REF, "test/reference", g + ".txt")) as f:
            head = [next(f) for _ in range(111)]
        open(os.path.join(HERE, g + ".head111.txt"), "w").writelines(head)

    # first five entries of the flat file (the last one without its "//"), for the flat-file reader
    lines, seen = [], 0
    for line in open(os.path.join(REF, "data/dbs/uniprot_subset.dat")):
        lines.append(line)
        if line.startswith("//"):
            seen += 1
            if seen == 5:
                break
    open(os.path.join(HERE, "uniprot_head5.dat"), "w").writelines(lines[:-1])

    exp_dir = os.path.join(ROOT, "baseline", "_ref", "exp")
    if os.path.isdir(exp_dir):
        exp = {}
        for fn in sorted(os.listdir(exp_dir)):
            exp[fn[:-4]] = [int(x) for x in open(os.path.join(exp_dir, fn)).read().split()]
        json.dump(exp, open(os.path.join(HERE, "survey_exp_blosum50.json"), "w"))

    # compiled cpu.cpp, +3/-3: three short queries x all 111 subjects, plus small known pairs
    cpu = {"pairs": [], "scans": {}}
    for a, b in [("GGTTGACTA", "TGTTACGG"), ("TGTTACGG", "GGTTGACTA"), ("AAAA", "AAAA"), ("ACGT", "TGCA"),
                 ("MKV", "W"), ("HEAGAWGHEE", "PAWHEAE")]:
        s, x, y = cpu_ref_run(a, b)
        cpu["pairs"].append({"a": a, "b": b, "score": s, "aligned_a": x, "aligned_b": y})
    for q in ("P02232", "P05013", "P14942"):
        qs = read_query(os.path.join(REF, "data/queries", q + ".fasta"))
        cpu["scans"][q] = [cpu_ref_run(qs, s)[0] for s in seqs]
    # a few alignments on real proteins (traceback parity)
    qs = read_query(os.path.join(REF, "data/queries", "P02232.fasta"))
    for k in (0, 16, 110):
        s, x, y = cpu_ref_run(qs, seqs[k])
        cpu["pairs"].append({"a": qs, "b": seqs[k], "score": s, "aligned_a": x, "aligned_b": y})
    json.dump(cpu, open(os.path.join(HERE, "cpu_ref_ident3.json"), "w"))

    # reference parser behaviour on the fixtures
    os.makedirs(os.path.join(HERE, "parser_expect"), exist_ok=True)
    tricky = os.path.join(HERE, "tricky.fasta")
    open(tricky, "wb").write(b">a\nACD\n\n>b\n\n>c\nAC\r\nD\n")
    open(os.path.join(HERE, "empty.fasta"), "wb").write(b"")
    open(os.path.join(HERE, "noeol.fasta"), "wb").write(b">x y z\nMKVLAAGIW\nWW\n>second\nACDEFGHIKLMNPQRSTVWY")
    for name in ("uniprot_subset.fasta", "test.dat", "tricky.fasta", "empty.fasta", "noeol.fasta"):
        out = subprocess.run([PARSER_PROBE, "db", os.path.join(HERE, name)], capture_output=True, check=True).stdout
        open(os.path.join(HERE, "parser_expect", name + ".db.txt"), "wb").write(out)
    for name in ("queries/P02232.fasta", "tricky.fasta", "empty.fasta", "noeol.fasta"):
        out = subprocess.run([PARSER_PROBE, "query", os.path.join(HERE, name)], capture_output=True, check=True).stdout
        open(os.path.join(HERE, "parser_expect", os.path.basename(name) + ".query.txt"), "wb").write(out)
    out = subprocess.run([PARSER_PROBE, "db", "/nonexistent/path"], capture_output=True, check=True).stdout
    open(os.path.join(HERE, "parser_expect", "nonexistent.db.txt"), "wb").write(out)
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
