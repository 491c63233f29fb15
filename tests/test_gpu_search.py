"""GPU suite: the query half of the index (dsmfm_searcher_*, csrc/search.cu) against the answers of the unmodified
reference (tests/golden/queries.json), the oracle's restatement of HuffWT::rank / access and FMIndex::LF, and plain
counting over the BWT.  Bit-exact (integer work)."""
import json
import os

import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
QUERIES = json.load(open(os.path.join(GOLDEN, "queries.json")))
NEG1 = 2**64 - 1


def _golden(name, ext):
    with open(os.path.join(GOLDEN, name + ext), "rb") as f:
        return f.read()


@pytest.mark.parametrize("name", sorted(QUERIES["cases"]))
def test_answers_of_the_reference_from_the_golden_fmi_file(name):
    import dsmfm
    q = QUERIES["cases"][name]
    with dsmfm.Searcher(os.path.join(GOLDEN, name + ".fmi")) as s:
        assert s.n == q["n"]
        lf = np.array(q["lf"], dtype=np.uint64).reshape(-1, 3)
        got = s.lf(lf[:, 0].astype(np.uint8), lf[:, 1])
        assert np.array_equal(got, lf[:, 2])
        gl = np.array(q["getl"], dtype=np.uint64).reshape(-1, 2)
        sym, _ = s.access(gl[:, 0])
        assert np.array_equal(sym, gl[:, 1].astype(np.uint8))


@pytest.mark.parametrize("seed,nreads,maxlen,alpha", [(31, 400, 60, "ACGT"), (32, 3000, 100, "ACGT"), (33, 500, 40, "ACGT0123."),
                                                      (34, 300, 30, "A")])
def test_built_index_answers_like_the_oracle_and_like_counting(seed, nreads, maxlen, alpha):
    """Searcher created from a finished builder (sections uploaded), every position and every symbol."""
    import dsmfm
    fasta = cases.rnd_fasta(seed, nreads, maxlen, alpha=alpha, genome=1500)
    docs, _ = oracle.fasta_to_docs(fasta)
    bwt = np.frombuffer(oracle.bwt(docs), dtype=np.uint8)
    n = bwt.size
    with dsmfm.Builder(device=0) as b:
        b.append_batch(docs)
        b.finish()
        fmi = b.fmi()
        with dsmfm.Searcher(b) as s:
            assert s.n == n
            pos = np.arange(n, dtype=np.uint64)
            sym, rank = s.access(pos)
            assert np.array_equal(sym, bwt)
            symbols = sorted(set(bwt.tolist()))
            C = {c: int((bwt < c).sum()) for c in range(257)}
            for c in symbols + [1, ord("Z"), 255]:
                occ = np.cumsum(bwt == c).astype(np.uint64)
                assert np.array_equal(s.rank(c, pos), occ)
                assert s.rank(c, [NEG1])[0] == 0
                want_lf = occ + np.uint64(C[c]) if C[c + 1] != C[c] else np.full(n, C[c], dtype=np.uint64)
                assert np.array_equal(s.lf(c, pos), want_lf)
                assert np.array_equal(rank[bwt == c], occ[bwt == c])
            # the oracle's tree walk over the same .fmi bytes
            x = oracle.Index(fmi)
            rng = np.random.default_rng(seed)
            qi = rng.integers(0, n, size=500, dtype=np.uint64)
            qc = rng.choice(np.array(symbols, dtype=np.uint8), size=500)
            assert s.lf(qc, qi).tolist() == [x.lf(int(c), int(i)) for c, i in zip(qc, qi)]
            x.close()


def test_trie_walk_like_enumerate_query():
    """Level-synchronous walk of the ACGT suffix trie with dsmfm_searcher_extend, checked against substring counts:
    what EnumerateQuery does one Query::pushChar at a time (Query.h:37-45), and pattern counts by backward search."""
    import dsmfm
    fasta = cases.rnd_fasta(41, 300, 50, genome=400, pn=0.0)
    docs, _ = oracle.fasta_to_docs(fasta)
    text = bytes(docs)
    with dsmfm.Builder(device=0) as b:
        b.append_batch(docs)
        b.finish()
        with dsmfm.Searcher(b) as s:
            frontier = {b"": (0, s.n - 1)}
            for depth in range(1, 5):
                keys = sorted(frontier)
                sp = np.array([frontier[k][0] for k in keys], dtype=np.uint64)
                ep = np.array([frontier[k][1] for k in keys], dtype=np.uint64)
                so, eo = s.extend(sp, ep, b"ACGT")
                nxt = {}
                for r, k in enumerate(keys):
                    for j, c in enumerate(b"ACGT"):
                        pat = bytes([c]) + k  # pushChar extends to the left
                        lo, hi = int(so[r, j]), int(eo[r, j])
                        occ = hi - lo + 1 if lo <= hi else 0
                        want = sum(1 for p in range(len(text)) if text.startswith(pat, p))
                        assert occ == want, (pat, lo, hi)
                        if occ:
                            nxt[pat] = (lo, hi)
                frontier = nxt
            pats = list(frontier)[:200] + [b"ACGTACGTTTTTTTTTTTT", b"-", b"A-T", b"N"]
            sp, ep = s.count(pats)
            for p, lo, hi in zip(pats, sp.tolist(), ep.tolist()):
                occ = hi - lo + 1 if lo <= hi else 0
                assert occ == sum(1 for q in range(len(text)) if text.startswith(p, q)), p
            # empty intervals pass through extend unchanged (EnumerateQuery.cpp:45-55 pushes them as they are)
            so, eo = s.extend([5, 9], [4, 3], b"AC")
            assert so.tolist() == [[5, 5], [9, 9]] and eo.tolist() == [[4, 4], [3, 3]]


def test_searcher_errors():
    import dsmfm
    with pytest.raises(dsmfm.DsmfmError):
        dsmfm.Searcher("/nonexistent/file.fmi")
    with dsmfm.Searcher(os.path.join(GOLDEN, "single.fmi")) as s:
        with pytest.raises(dsmfm.DsmfmError):
            s.access([s.n])


def test_large_index_lf_on_device_matches_counting():
    """25 Mbp sample (C1 shape): 4M LF queries answered on device arrays, checked against counting over the BWT."""
    import torch
    import dsmfm
    import dsmgen
    docs = dsmgen.docs(**dsmgen.CONFIGS["C1"])
    with dsmfm.Builder(device=0, flags=dsmfm.FLAG_KEEP_BWT) as b:
        b.append_batch(docs)
        b.finish()
        bwt = np.frombuffer(b.bwt(), dtype=np.uint8)
        with dsmfm.Searcher(b) as s:
            rng = np.random.default_rng(3)
            m = 4_000_000
            qi = np.sort(rng.integers(0, s.n, size=m, dtype=np.int64))
            qc = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=m)
            out = torch.empty(m, dtype=torch.int64, device="cuda")
            s.lf_device(torch.from_numpy(qc).cuda(), torch.from_numpy(qi).cuda(), out)
            got = out.cpu().numpy().astype(np.uint64)
            for c in b"ACGT":
                occ = np.cumsum(bwt == c, dtype=np.uint64)
                sel = qc == c
                want = occ[qi[sel]] + np.uint64(int((bwt < c).sum()))
                assert np.array_equal(got[sel], want)


# ---- the mining client's byte stream -------------------------------------------------------

def _stream_cases():
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "streams.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("key", sorted(_stream_cases()))
def test_enumerate_stream_equals_the_reference_clients(key):
    """tests/golden/streams.json holds what the UNMODIFIED reference client (metaenumerate: EnumerateQuery::enumerate
    over ClientSocket) sent a recording server for a golden index, an enforced path, fmin and maxdepth
    (tests/golden/make_stream_golden.py).  dsmfm_searcher_enumerate walks the same trie level by level on the GPU
    and must produce the same bytes behind the handshake 'S' name '.'."""
    import hashlib
    import dsmfm
    case = _stream_cases()[key]
    fmi = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", case["index"] + ".fmi")
    with dsmfm.Searcher(fmi) as s:
        got = b"S" + case["index"].encode() + b"." + s.enumerate(case["path"].encode(), fmin=case["fmin"], maxdepth=case["maxdepth"])
    assert len(got) == case["bytes"]
    assert hashlib.sha256(got).hexdigest() == case["sha256"]
    if "hex" in case:
        assert got == bytes.fromhex(case["hex"])


def test_enumerate_stream_to_a_socket_and_errors():
    import socket
    import threading
    import dsmfm
    case = _stream_cases()["reads100:G:f10:m0"]
    fmi = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reads100.fmi")
    a, b = socket.socketpair()
    got = []
    t = threading.Thread(target=lambda: [got.append(x) for x in iter(lambda: b.recv(1 << 16), b"")])
    t.start()
    with dsmfm.Searcher(fmi) as s:
        n = s.enumerate_to_fd(b"G", a.fileno(), fmin=10)
        a.close()
        t.join()
        assert n == case["bytes"] - len(b"Sreads100.") and len(b"".join(got)) == n
        with pytest.raises(dsmfm.DsmfmError):
            s.enumerate(b"A", fmin=1)  # unary-path following is not rebuilt
        # an enforced path that stops occurring half way: the nodes that were reached open and close, nothing below
        part = s.enumerate(b"A" * 64, fmin=2)
        assert part == b"" or (part.startswith(b"(A") and part.count(b"(") == part.count(b")"))
