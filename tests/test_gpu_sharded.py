"""The multi-GPU build on ONE GPU: every rank's device work (dsmfm_block_stats / block_pack / build_packed /
pieces_build / pieces_merge / pieces_write through the C ABI) runs for all ranks one after the other in one
process (multigpu.build_sharded_local), the exchanges become copies.  Everything but NCCL itself is exercised,
so the path bench.py --gpus N and `builder --gpus N` take is checked bit for bit on the one-GPU test box:
against the oracle, against the committed golden files, against the reference digests of tests/golden/fullsize.json."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import cases
import oracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _blocks(docs, world, device=None):
    import multigpu
    doc_list = docs.split(b"\0")[:-1]
    out = []
    for r in range(world):
        b, e = multigpu.block_of(len(doc_list), r, world)
        local = b"".join(d + b"\0" for d in doc_list[b:e])
        t = torch.frombuffer(bytearray(local), dtype=torch.uint8) if local else torch.empty(0, dtype=torch.uint8)
        out.append(t.cuda() if device == "cuda" and local else (t.pin_memory() if local else t))
    return out


def _build_file(blocks, tmp_path, name, ranges=1, order=None):
    import multigpu
    world = len(blocks)
    sbs = multigpu.build_sharded_local(blocks, [multigpu.CudaEngine(0) for _ in range(world)], ranges_per_gpu=ranges)
    prefix = str(tmp_path / name)
    for r in (order or range(world)):
        sbs[r].write(prefix)
    held = [sb.section_bytes for sb in sbs]
    for sb in sbs:
        sb.close()
    return open(prefix + ".fmi", "rb").read(), held


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("name", ["reads100", "duplicates", "poly_a", "colour_space", "mixed_alphabet", "one_base_reads", "single"])
def test_ranks_on_one_gpu_reproduce_the_golden_files(world, name, tmp_path):
    fa = open(os.path.join(HERE, "golden", name + ".fasta"), "rb").read()
    want = open(os.path.join(HERE, "golden", name + ".fmi"), "rb").read()
    docs, _ = oracle.fasta_to_docs(fa)
    got, _ = _build_file(_blocks(docs, world), tmp_path, name, ranges=1 + world % 2, order=list(reversed(range(world))))
    assert got == want, oracle.diff_fmi(got, want)


@pytest.mark.parametrize("world,ranges,lo_bits", [(2, 1, None), (4, 2, None), (3, 1, "14"), (7, 3, "12")])
def test_ranks_on_one_gpu_match_the_oracle(world, ranges, lo_bits, tmp_path, monkeypatch):
    """Random collections incl. 8-bit alphabets; DSMFM_POS_LO_BITS puts the high position bits into play."""
    if lo_bits:
        monkeypatch.setenv("DSMFM_POS_LO_BITS", lo_bits)
    rng = np.random.default_rng(world)
    for case in range(3):
        if case == 0:
            docs, _ = oracle.fasta_to_docs(cases.rnd_fasta(40 + world, 3000, 100, minlen=20, genome=3000))
        elif case == 1:
            docs, _ = oracle.fasta_to_docs(cases.rnd_fasta(50 + world, 2500, 60, minlen=60, genome=300, dup=0.0, pn=0.0))
        else:  # sigma = 40: 8 bits per symbol
            parts = [bytes(rng.integers(33, 73, size=int(rng.integers(1, 80)), dtype=np.uint8)) for _ in range(700)]
            docs = b"".join(p + b"\0" for p in parts)
        got, _ = _build_file(_blocks(docs, world, "cuda" if case == 1 else None), tmp_path, "c%d" % case, ranges=ranges)
        want = oracle.fmi_from_docs(docs)
        assert got == want, "case %d: %r" % (case, oracle.diff_fmi(got, want))


def test_a_rank_without_documents(tmp_path):
    docs, _ = oracle.fasta_to_docs(cases.rnd_fasta(7, 5, 30))
    blocks = _blocks(docs, 8)  # five documents over eight ranks: three ranks hold nothing
    assert sum(b.numel() == 0 for b in blocks) == 3
    got, held = _build_file(blocks, tmp_path, "sparse")
    assert got == oracle.fmi_from_docs(docs)


@pytest.mark.parametrize("config,world", [("C4s", 4), ("C5s", 2), ("C1", 8)])
def test_ranks_on_one_gpu_reproduce_the_reference_digest(config, world, tmp_path):
    """Reduced C4 / C5 shapes and C1 (1M / 1M / 250k reads): the file the ranks write together has the SHA-256 of
    the file the unmodified reference wrote for the same documents (tests/golden/fullsize.json)."""
    import dsmgen
    with open(os.path.join(HERE, "golden", "fullsize.json")) as f:
        want = json.load(f)[config]
    docs = dsmgen.docs(**want["params"])
    L = 2 * want["params"]["read_len"] + 2
    nreads = want["params"]["n_reads"]
    blocks = []
    import multigpu
    for r in range(world):
        b, e = multigpu.block_of(nreads, r, world)
        blocks.append(torch.from_numpy(docs[b * L:e * L]).pin_memory())
    got, held = _build_file(blocks, tmp_path, config, ranges=2)
    assert len(got) == want["fmi_bytes"]
    assert hashlib.sha256(got).hexdigest() == want["fmi_sha256"]
    # the sections are spread over the ranks: nobody holds the whole index
    assert sum(held) <= want["fmi_bytes"] and max(held) < 0.6 * want["fmi_bytes"]
