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
