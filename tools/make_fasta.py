#!/usr/bin/env python3
"""Writes the synthetic benchmark database (bench.synth_db) as a FASTA file, for programs that read files
(the reference's CUDA solver behind oracle/_ref/ref_cuda_scan, and bin/main).  usage: make_fasta.py OUT [scale]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_db  # noqa: E402

LETTERS = np.frombuffer(b"ARNDCQEGHILKMFPSTWYVBJZXU", dtype=np.uint8)


def main():
    out = sys.argv[1]
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    codes, offsets = synth_db(scale=scale)
    text = LETTERS[np.minimum(codes, 24)].tobytes()
    with open(out, "wb") as f:
        for i in range(len(offsets) - 1):
            f.write(b">s%d\n" % i)
            f.write(text[int(offsets[i]):int(offsets[i + 1])])
            f.write(b"\n")
    print(len(offsets) - 1, int(offsets[-1]))


if __name__ == "__main__":
    main()
