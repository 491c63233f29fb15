#!/usr/bin/env python
"""Regenerates tests/golden/queries.json: answers of the UNMODIFIED reference (TextCollection::load + LF / getL,
driven by oracle/ref_driver.cpp `query`) to seeded queries on the golden .fmi files.  Run in the build container:

    make -C oracle ref && python tests/golden/make_query_golden.py
"""
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402

CASES = ["single", "duplicates", "multiline_and_blank", "poly_a", "two_letter", "colour_space", "mixed_alphabet",
         "small_random", "reads100", "one_base_reads", "empty"]
ULONG_MAX = 2**64 - 1


def queries_for(name, n, symbols, rng):
    """(op, c, i): every symbol at the boundaries (i = -1, 0, n-1) and at random positions; getL everywhere small."""
    qs = []
    for c in symbols:
        for i in (ULONG_MAX, 0, n - 1):
            qs.append(("L", c, i))
        for _ in range(24):
            qs.append(("L", c, rng.randrange(n)))
    for c in (1, ord("Z"), 254):  # symbols that do not occur
        qs.append(("L", c, rng.randrange(n)))
    pos = range(n) if n <= 64 else [0, n - 1] + [rng.randrange(n) for _ in range(96)]
    for i in pos:
        qs.append(("G", 0, i))
    return qs


def main():
    assert oracle.have_reference(), "build the reference first: make -C oracle ref"
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    out = {"reference": "TextCollection::load + LF / getL of the unmodified reference (oracle/ref_driver.cpp query)", "cases": {}}
    rng = random.Random(1234)
    with tempfile.TemporaryDirectory() as tmp:
        for name in CASES:
            fmi = open(os.path.join(HERE, name + ".fmi"), "rb").read()
            import struct
            parsed = oracle.parse_fmi(fmi)
            n = struct.unpack("<Q", parsed["n"])[0]
            symbols = [c for c in range(256) if struct.unpack_from("<Q", parsed["codetable"], 16 * c)[0] > 0]
            qs = queries_for(name, n, symbols, rng)
            qf, af = os.path.join(tmp, "q.txt"), os.path.join(tmp, "a.txt")
            with open(qf, "w") as f:
                for op, c, i in qs:
                    f.write("L %d %d\n" % (c, i) if op == "L" else "G %d\n" % i)
            subprocess.run([exe, "query", os.path.join(HERE, name + ".fmi"), qf, af], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            ans = [int(x) for x in open(af).read().split()]
            assert len(ans) == len(qs)
            out["cases"][name] = {"n": n, "lf": [[c, i, a] for (op, c, i), a in zip(qs, ans) if op == "L"],
                                  "getl": [[i, a] for (op, c, i), a in zip(qs, ans) if op == "G"]}
    with open(os.path.join(HERE, "queries.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", sum(len(v["lf"]) + len(v["getl"]) for v in out["cases"].values()), "answers for", len(CASES), "indexes")


if __name__ == "__main__":
    main()
