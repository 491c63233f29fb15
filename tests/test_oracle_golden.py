"""CPU suite, part 1: the oracle (oracle/dsm_oracle.c) is pinned against the golden files the
UNMODIFIED reference builder produced (tests/golden/, made by tests/golden/make_golden.py) and,
when the compiled reference is present (oracle/_ref), against the reference run live."""
import hashlib
import json
import os

import pytest

import cases
import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLDEN, "manifest.json")))


def _golden(name, ext):
    with open(os.path.join(GOLDEN, name + ext), "rb") as f:
        return f.read()


@pytest.mark.parametrize("name", sorted(MANIFEST["files"]))
def test_oracle_matches_golden_file(name):
    fasta = _golden(name, ".fasta")
    assert fasta == cases.golden_cases()[name], "tests/cases.py drifted from the committed fixture"
    want = _golden(name, ".fmi")
    got = oracle.build(fasta)
    assert oracle.diff_fmi(got, want) == []
    assert got == want


def test_oracle_samplerate_field():
    got = oracle.build(_golden("small_random", ".fasta"), samplerate=32)
    assert got == _golden("small_random", ".s32.fmi")
    # -s only changes the 4-byte header field
    assert oracle.diff_fmi(got, _golden("small_random", ".fmi")) == ["samplerate"]


@pytest.mark.parametrize("name", sorted(MANIFEST["digests"]))
def test_oracle_matches_golden_digest(name):
    fasta = cases.digest_cases()[name]
    assert hashlib.sha256(fasta).hexdigest() == MANIFEST["digests"][name]["fasta_sha256"]
    got = oracle.build(fasta)
    assert len(got) == MANIFEST["digests"][name]["fmi_bytes"]
    assert hashlib.sha256(got).hexdigest() == MANIFEST["digests"][name]["fmi_sha256"]


def test_oracle_matches_golden_generated_case():
    import dsmgen
    g = MANIFEST["generated"]["gen_20k"]
    fasta = dsmgen.fasta(**g["params"]).tobytes()
    assert hashlib.sha256(fasta).hexdigest() == g["fasta_sha256"], "generator drifted"
    got = oracle.build(fasta)
    assert hashlib.sha256(got).hexdigest() == g["fmi_sha256"]


def test_generator_docs_equal_front_end():
    """dsmgen.docs (used to feed bench.py) == FASTA -> builder front end -> documents."""
    import dsmgen
    kw = dict(seed=5, pool_seed=6, pool_size=3, n_genomes=2, genome_len=3000, n_reads=500, read_len=37,
              sub=0.02, pn=0.01)
    docs, nd = oracle.fasta_to_docs(dsmgen.fasta(**kw).tobytes())
    assert nd == 500
    assert docs == dsmgen.docs(**kw).tobytes()


def test_transform_definition():
    # doc = complement(read) + '-' + reverse(read), only A<->T, C<->G complemented (builder.cpp:35-55,183-201)
    assert oracle.transform(b"ACGTN") == b"TGCAN-NTGCA"
    assert oracle.transform(b"acgtnxR0123.") == b"TGCANNN0123.-.3210NNNTGCA"


def test_bwt_definition_small():
    # two identical documents: ties are broken by document order, terminators sort first in document order
    text = b"AC\0AC\0"
    b, sa = oracle.bwt(text, want_sa=True)
    assert list(sa) == [2, 5, 0, 3, 1, 4]
    assert b == b"CC\0\0AA"


@pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [101, 102, 103])
def test_oracle_matches_live_reference(seed, tmp_path):
    fasta = cases.rnd_fasta(seed, 300, 80, genome=1500, dup=0.25, lower=0.2, wrap=50)
    assert oracle.build(fasta) == oracle.reference_build(fasta, tmp_path)


# ---- the dormant SA sampling: `.sa` files written by the reference's FMIndex::saveSamples -------------

@pytest.mark.parametrize("fn", sorted(MANIFEST["sa"]))
def test_oracle_sa_file_matches_reference(fn):
    e = MANIFEST["sa"][fn]
    docs, _ = oracle.fasta_to_docs(_golden(e["case"], ".fasta"))
    want = _golden(fn, "")
    assert hashlib.sha256(want).hexdigest() == e["sha256"]
    assert oracle.sa_file_from_docs(docs, e["samplerate"]) == want


@pytest.mark.parametrize("key", sorted(MANIFEST["sa_digests"]))
def test_oracle_sa_file_matches_reference_digest(key):
    e = MANIFEST["sa_digests"][key]
    docs, _ = oracle.fasta_to_docs(cases.digest_cases()[e["case"]])
    got = oracle.sa_file_from_docs(docs, e["samplerate"])
    assert len(got) == e["bytes"] and hashlib.sha256(got).hexdigest() == e["sha256"]


def test_oracle_sa_file_generated_case():
    import dsmgen
    g = MANIFEST["generated"]["gen_20k"]
    got = oracle.sa_file_from_docs(dsmgen.docs(**g["params"]).tobytes())
    assert len(got) == g["sa_bytes"] and hashlib.sha256(got).hexdigest() == g["sa_sha256"]


@pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not present")
def test_oracle_sa_file_matches_live_reference(tmp_path):
    fa = cases.rnd_fasta(77, 300, 90, genome=1500, dup=0.3)
    docs, _ = oracle.fasta_to_docs(fa)
    for rate in (6, 31):
        assert oracle.sa_file_from_docs(docs, rate) == oracle.reference_sa(fa, tmp_path, samplerate=rate)


# ---- the query side: FMIndex::LF / getL as the mining client calls them ----------------

QUERIES = json.load(open(os.path.join(GOLDEN, "queries.json")))


@pytest.mark.parametrize("name", sorted(QUERIES["cases"]))
def test_oracle_queries_match_the_reference(name):
    """dsm_oracle_lf / dsm_oracle_access against the answers of the unmodified reference (tests/golden/queries.json,
    made by tests/golden/make_query_golden.py through TextCollection::load + LF / getL)."""
    q = QUERIES["cases"][name]
    x = oracle.Index(_golden(name, ".fmi"))
    assert x.n == q["n"]
    for c, i, want in q["lf"]:
        assert x.lf(c, i) == want, (c, i)
    for i, want in q["getl"]:
        assert x.access(i)[0] == want, i
    x.close()


@pytest.mark.parametrize("name", ["small_random", "mixed_alphabet", "poly_a", "duplicates"])
def test_oracle_queries_match_plain_counting_over_the_bwt(name):
    """rank(c, i) is the number of c in BWT[0..i]; access(i) is BWT[i] together with its rank."""
    docs, _ = oracle.fasta_to_docs(_golden(name, ".fasta"))
    bwt = oracle.bwt(docs)
    x = oracle.Index(_golden(name, ".fmi"))
    assert x.n == len(bwt)
    seen = {}
    for i, c in enumerate(bwt):
        seen[c] = seen.get(c, 0) + 1
        assert x.access(i) == (c, seen[c])
        if i % 7 == 0 or i + 1 == len(bwt):
            for s in set(bwt):
                assert x.rank(s, i) == seen.get(s, 0)
    for s in set(bwt):
        assert x.rank(s, -1) == 0
    x.close()


@pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not present")
def test_oracle_queries_match_live_reference(tmp_path):
    import random
    import subprocess
    fasta = cases.rnd_fasta(77, 300, 80)
    fmi = oracle.build(fasta)
    (tmp_path / "x.fmi").write_bytes(fmi)
    x = oracle.Index(fmi)
    rng = random.Random(5)
    qs = [("L", rng.choice(b"\0-ACGNT"), rng.randrange(x.n)) for _ in range(400)] + [("G", 0, rng.randrange(x.n)) for _ in range(200)]
    (tmp_path / "q.txt").write_text("".join("L %d %d\n" % (c, i) if op == "L" else "G %d\n" % i for op, c, i in qs))
    exe = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "ref_driver")
    subprocess.run([exe, "query", str(tmp_path / "x.fmi"), str(tmp_path / "q.txt"), str(tmp_path / "a.txt")], check=True,
                   capture_output=True)
    ans = [int(v) for v in (tmp_path / "a.txt").read_text().split()]
    for (op, c, i), a in zip(qs, ans):
        assert (x.lf(c, i) if op == "L" else x.access(i)[0]) == a
    x.close()


# ---- the record loop on adversarial FASTA (rows before the first header, runs of headers, CRLF, '>' inside a row ...) ----

NASTY = json.load(open(os.path.join(GOLDEN, "nasty.json")))


@pytest.mark.parametrize("seed,rows,longest", cases.NASTY_CASES)
def test_oracle_front_end_matches_the_reference_on_adversarial_fasta(seed, rows, longest):
    """tests/golden/nasty.json: digests of what the unmodified reference builder wrote for cases.nasty_fasta."""
    e = NASTY["cases"]["nasty_%d" % seed]
    fasta = cases.nasty_fasta(seed, rows, longest)
    assert hashlib.sha256(fasta).hexdigest() == e["fasta_sha256"], "tests/cases.py drifted from the committed digest"
    got = oracle.build(fasta)
    assert len(got) == e["fmi_bytes"] and hashlib.sha256(got).hexdigest() == e["fmi_sha256"]


@pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not present")
def test_oracle_front_end_matches_live_reference_on_adversarial_fasta(tmp_path):
    for seed in (11, 12, 13):
        fasta = cases.nasty_fasta(seed, 200, 150)
        assert oracle.build(fasta) == oracle.reference_build(fasta, tmp_path)
