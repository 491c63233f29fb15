"""Seeded small inputs shared by the golden-file generator and the parity tests."""
import random


def rnd_fasta(seed, nreads, maxlen, alpha="ACGT", dup=0.2, pn=0.02, genome=2000, minlen=1, lower=0.0, wrap=0):
    rng = random.Random(seed)
    g = "".join(rng.choice(alpha) for _ in range(genome))
    reads, out = [], []
    for i in range(nreads):
        if reads and rng.random() < dup:
            r = rng.choice(reads)
        else:
            ln = rng.randint(minlen, maxlen)
            s = rng.randint(0, genome - ln)
            r = "".join("N" if rng.random() < pn else c for c in g[s:s + ln])
            if lower and rng.random() < lower:
                r = r.lower()
        reads.append(r)
        body = r if not wrap else "\n".join(r[j:j + wrap] for j in range(0, len(r), wrap))
        out.append(">r%d some description\n%s\n" % (i, body))
    return "".join(out).encode()


NASTY_CASES = [(1, 5, 10), (2, 40, 30), (3, 300, 90), (4, 2000, 200), (5, 50, 9000)]  # seed, rows, longest row


def nasty_fasta(seed, nlines, maxlen):
    """Everything the record loop has an opinion about: rows in front of the first header, empty rows, runs of
    headers, CRLF, lower case, foreign symbols, '>' inside a row, long rows, blank-padded names."""
    rng = random.Random(seed)
    out = []
    if rng.random() < 0.5:
        out.append("ACGTTGCA"[: rng.randint(0, 8)] + "\n")
    for i in range(nlines):
        r = rng.random()
        if r < 0.25:
            out.append(">%sr%d%s\n" % (" \t"[rng.randint(0, 1)] * rng.randint(0, 2), i, rng.choice(["", " desc", "\tx y"])))
        elif r < 0.30:
            out.append("\n")
        else:
            ln = rng.randint(1, maxlen)
            alpha = rng.choice(["ACGT", "ACGTN", "acgtn", "ACGTacgtRYKM", "ACGT>", "AC GT"])
            row = "".join(rng.choice(alpha) for _ in range(ln))
            if row[0] == ">":
                row = "A" + row[1:]
            out.append(row + rng.choice(["\n", "\n", "\n", "\r\n"]))
    s = "".join(out)
    if rng.random() < 0.3:
        s += "ACGTAC"  # unterminated last row: dropped
    return s.encode()


# name -> FASTA bytes.  Small enough for the plain-C oracle and for committing the reference's output.
def golden_cases():
    c = {}
    c["empty"] = b""
    c["single"] = b">a\nACGT\n"
    c["no_trailing_newline"] = b">a\nACGT\n>b\nGGCA"          # last line dropped (builder.cpp:211)
    c["multiline_and_blank"] = b">a\nAC\nGT\n>b\n\n>c\nacgtnxyz0123.\n>d \tname cut\nTTTT\n"
    c["crlf"] = b">a\r\nACGT\r\n>b\r\nGG\r\n"                  # '\r' is an invalid symbol -> N
    c["no_header_first"] = b"ACGT\n>b\nGGA\n"
    c["duplicates"] = (b">x\nACGTACGTAC\n" * 7) + b">y\nACGTACGTAC\n>z\nCGTACGTACG\n"
    c["one_base_reads"] = b"".join(b">r%d\n%s\n" % (i, b"ACGTN"[i % 5:i % 5 + 1]) for i in range(23))
    c["poly_a"] = rnd_fasta(3, 200, 40, alpha="A", genome=100)          # huge tie groups
    c["two_letter"] = rnd_fasta(4, 300, 50, alpha="AC", genome=100)
    c["colour_space"] = rnd_fasta(5, 150, 35, alpha="0123.", genome=400, pn=0.0)   # 4-bit alphabet
    c["mixed_alphabet"] = rnd_fasta(6, 150, 35, alpha="ACGT0123.", genome=400)     # sigma = 11
    c["small_random"] = rnd_fasta(7, 50, 30)
    c["reads100"] = rnd_fasta(8, 400, 100, minlen=100, genome=3000, lower=0.3, wrap=60)
    return c


# Larger seeded cases: only their SHA-256 is committed (tests/golden/manifest.json).
def digest_cases():
    c = {}
    c["reads100_3k"] = rnd_fasta(9, 3000, 100, minlen=100, genome=5000)
    c["ragged_2k"] = rnd_fasta(10, 2000, 150, genome=4000, dup=0.3)
    c["high_coverage"] = rnd_fasta(11, 4000, 60, minlen=60, genome=600, dup=0.0, pn=0.0)
    return c


# (golden case, sample rate) pairs whose `.sa` file (FMIndex::saveSamples) is committed as <case>.s<rate>.sa
SA_CASES = [("reads100", 124), ("reads100", 16), ("small_random", 32), ("small_random", 3), ("duplicates", 5),
            ("one_base_reads", 2), ("poly_a", 7), ("mixed_alphabet", 16), ("single", 3), ("single", 124),
            ("multiline_and_blank", 4), ("two_letter", 9), ("colour_space", 11)]
# digest-only
SA_DIGEST_CASES = [("reads100_3k", 124), ("ragged_2k", 40), ("high_coverage", 13)]
