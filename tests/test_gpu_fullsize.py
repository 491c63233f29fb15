"""GPU suite at BASELINE.json's sizes, where neither the oracle nor the reference finishes in test
time: size-independent properties of the result (sortedness of sampled suffix-array windows under the
reference's comparison rule, BWT/text consistency, rank-directory self-consistency)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _suffix_less_equal(text, a, b):
    """Reference order: compare up to the first terminator; identical up to it -> text order."""
    i = 0
    while True:
        x, y = text[a + i], text[b + i]
        if x != y:
            return x < y
        if x == 0:
            return a < b
        i += 1


def _check_index_properties(docs, nreads, doc_len, b, idx, rng, windows=24, window=2048):
    n = idx.n
    assert n == docs.size and idx.number_of_texts == nreads and idx.max_text_length == doc_len
    bwt = np.frombuffer(b.bwt(), dtype=np.uint8)
    # (1) the BWT is a permutation of the text
    assert np.array_equal(np.bincount(bwt, minlength=256), np.bincount(docs, minlength=256))
    # (2) BWT[k] for k < D is the last symbol of document k; C[1] = D
    last = docs.reshape(nreads, doc_len)[:, doc_len - 2]
    assert np.array_equal(bwt[:nreads], last)
    assert idx.C[1] == nreads and idx.C[0] == 0
    # (3) sampled windows of the suffix array are sorted under the reference rule and agree with the BWT
    for _ in range(windows):
        first = int(rng.integers(0, n - window))
        sa = b.suffix_array(first, window).astype(np.int64)
        prev = np.where(sa > 0, docs[np.maximum(sa - 1, 0)], 0)
        assert np.array_equal(prev, bwt[first:first + window])
        for j in range(0, window - 1, 7):
            assert _suffix_less_equal(docs, int(sa[j]), int(sa[j + 1])), "order violated at rank %d" % (first + j)
    # (4) the suffix array is a permutation: first D entries are the terminators in document order
    sa0 = b.suffix_array(0, min(nreads, 100000)).astype(np.int64)
    assert np.array_equal(sa0, np.arange(sa0.size, dtype=np.int64) * doc_len + doc_len - 1)
    # (5) root bitvector: its rank directory is consistent with its bits, and its length is n
    root = idx.nodes[0]
    assert not root.leaf and root.nbits == n and root.ch == bwt[0]
    words = np.ctypeslib.as_array(root.data, shape=(root.integers,))
    rs = np.ctypeslib.as_array(root.Rs, shape=(n // 256 + 1,))
    pc = np.bitwise_count(words) if hasattr(np, "bitwise_count") else np.array([bin(int(x)).count("1") for x in words])
    cum = np.concatenate([[0], np.cumsum(pc.astype(np.uint64))])
    assert np.array_equal(rs, cum[0:4 * (n // 256) + 1:4])


@pytest.mark.parametrize("config,scale", [("C1", 1.0), ("C3", 1.0)])
def test_full_size_build_properties(config, scale):
    import dsmfm
    import dsmgen
    kw = dict(dsmgen.CONFIGS[config])
    if os.environ.get("DSMFM_TEST_SMALL"):
        kw["n_reads"] //= 10
    nreads, L = kw["n_reads"], kw["read_len"]
    docs = dsmgen.docs(**kw)
    rng = np.random.default_rng(3)
    with dsmfm.Builder(flags=dsmfm.FLAG_KEEP_BWT | dsmfm.FLAG_KEEP_SA, expected_bytes=docs.size) as b:
        b.append_batch(docs)
        idx = b.finish()
        _check_index_properties(docs, nreads, 2 * L + 2, b, idx, rng)
        s = b.stats()
        assert s.fallback_elems == 0 or config != "C1"
