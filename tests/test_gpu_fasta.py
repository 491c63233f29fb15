"""GPU suite: the FASTA front end on the device (dsmfm_append_fasta) against the oracle's restatement of the
reference CLI's record loop (builder.cpp:203-262) and against the golden files of the unmodified reference."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MANIFEST = json.load(open(os.path.join(GOLDEN, "manifest.json")))
NASTY = json.load(open(os.path.join(GOLDEN, "nasty.json")))


def _golden(name, ext):
    with open(os.path.join(GOLDEN, name + ext), "rb") as f:
        return f.read()


def _fmi_from_fasta(fasta, chunk=None, **kw):
    """Build through dsmfm_append_fasta; chunk = feed that many new bytes per call (streaming)."""
    import dsmfm
    infos = []
    with dsmfm.Builder(device=0, **kw) as b:
        if chunk is None:
            infos.append(b.append_fasta(fasta, final=True))
        else:
            pending = b""
            pos = 0
            while pos < len(fasta):
                pending += fasta[pos:pos + chunk]
                pos += chunk
                last = pos >= len(fasta)
                info = b.append_fasta(pending, final=last)
                infos.append(info)
                pending = pending[info["consumed"]:]
            if not fasta:
                infos.append(b.append_fasta(b"", final=True))
        b.finish()
        return b.fmi(), infos


def _expect(fasta):
    docs, nd = oracle.fasta_to_docs(fasta)
    return oracle.fmi_from_docs(docs), docs, nd


def _assert_same_fmi(got, want):
    assert oracle.diff_fmi(got, want) == []
    assert got == want


nasty_fasta = cases.nasty_fasta


@pytest.mark.parametrize("name", sorted(MANIFEST["files"]))
def test_golden_fasta_through_the_gpu_front_end(name):
    fasta = _golden(name, ".fasta")
    got, infos = _fmi_from_fasta(fasta)
    docs, nd = oracle.fasta_to_docs(fasta)
    assert infos[0]["documents"] == nd and infos[0]["doc_bytes"] == len(docs)
    _assert_same_fmi(got, _golden(name, ".fmi"))
    if name == "small_random":
        assert _fmi_from_fasta(fasta, samplerate=32)[0] == _golden(name, ".s32.fmi")


@pytest.mark.parametrize("seed,nlines,maxlen", [(1, 5, 10), (2, 40, 30), (3, 300, 90), (4, 2000, 200), (5, 50, 9000),
                                                (6, 4000, 70), (7, 20000, 120), (8, 300, 5000)])
def test_nasty_fasta_matches_the_oracle(seed, nlines, maxlen):
    fasta = nasty_fasta(seed, nlines, maxlen)
    want, docs, nd = _expect(fasta)
    got, infos = _fmi_from_fasta(fasta)
    i = infos[0]
    assert (i["documents"], i["doc_bytes"], i["bases"]) == (nd, len(docs), (len(docs) - 2 * nd) // 2)
    assert i["records"] == sum(1 for row in fasta.split(b"\n")[:-1] if row[:1] == b">")
    assert i["bad_headers"] == 0
    _assert_same_fmi(got, want)
    pinned = NASTY["cases"].get("nasty_%d" % seed)  # what the unmodified reference wrote for this input
    if pinned and (seed, nlines, maxlen) in cases.NASTY_CASES:
        assert hashlib.sha256(got).hexdigest() == pinned["fmi_sha256"]


@pytest.mark.parametrize("seed,chunk", [(3, 7), (3, 100), (4, 4096), (6, 5000), (7, 65536), (8, 1 << 20)])
def test_streaming_calls_equal_one_call(seed, chunk):
    nl, ml = {3: (300, 90), 4: (2000, 200), 6: (4000, 70), 7: (20000, 120), 8: (300, 5000)}[seed]
    fasta = nasty_fasta(seed, nl, ml)
    want, docs, nd = _expect(fasta)
    got, infos = _fmi_from_fasta(fasta, chunk=chunk)
    assert sum(i["documents"] for i in infos) == nd
    assert sum(i["doc_bytes"] for i in infos) == len(docs)
    _assert_same_fmi(got, want)


def test_invalid_symbols_and_blank_headers_are_reported():
    import dsmfm
    fasta = b">a\nACGT\n>b\nACXT\n>c\nRRRR\nAC\n>d\nAC\r\n"
    with dsmfm.Builder(device=0) as b:
        i = b.append_fasta(fasta)
        assert i["invalid_records"] == 3 and i["first_invalid_offset"] == fasta.index(b"X")
        assert i["records"] == 4 and i["documents"] == 4 and i["bad_headers"] == 0
    for bad in (b">\nACGT\n", b">a\nAC\n>  \t\nGG\n", b">a\nAC\n>"):
        with dsmfm.Builder(device=0) as b:
            i = b.append_fasta(bad)
            assert i["bad_headers"] == (0 if bad.endswith(b">") else 1)  # an unterminated last row is not a row


def test_fasta_and_insert_text_can_be_mixed():
    import dsmfm
    f1, f2 = cases.rnd_fasta(21, 200, 60), cases.rnd_fasta(22, 150, 80)
    d1, _ = oracle.fasta_to_docs(f1)
    d2, _ = oracle.fasta_to_docs(f2)
    with dsmfm.Builder(device=0) as b:
        b.append_fasta(f1)
        for doc in d2.split(b"\0")[:-1]:
            b.insert_text(doc)
        b.append_fasta(f1)
        b.finish()
        got = b.fmi()
    _assert_same_fmi(got, oracle.fmi_from_docs(d1 + d2 + d1))


def test_generated_sample_fasta_equals_generated_documents():
    """25 Mbp (BASELINE configs[0] shape): the generator's FASTA through the GPU front end gives the same index
    as the generator's documents through append_batch."""
    import dsmfm
    import dsmgen
    kw = dict(dsmgen.CONFIGS["C1"])
    fasta = dsmgen.fasta(**kw)
    docs = dsmgen.docs(**kw)
    with dsmfm.Builder(device=0) as b:
        info = b.append_fasta(fasta)
        b.finish()
        got = b.fmi()
    assert info["documents"] == kw["n_reads"] and info["doc_bytes"] == docs.size and info["invalid_records"] == 0
    assert got == dsmfm.build_fmi(docs)
