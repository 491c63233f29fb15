"""Host logic of the multi-GPU build (dsm-framework_b200/multigpu.py) on CPU: world_size 2 and 3 with the gloo
backend, and all ranks in one process.  The device work is replaced by tests/cpu_engine.py (a numpy model over
the oracle), so what is exercised is everything around it: uneven and empty document blocks, the exchanges of
block records / packed slots / slice histograms / edge records, slice tiling checks, and the `.fmi` file that
every rank writes its share of -- which must equal the oracle's byte for byte, including the words and
directory entries that straddle slice boundaries."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import oracle
from cpu_engine import CpuEngine


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _block(doc_list, rank, world):
    import multigpu
    b, e = multigpu.block_of(len(doc_list), rank, world)
    local = b"".join(d + b"\0" for d in doc_list[b:e])
    return torch.frombuffer(bytearray(local), dtype=torch.uint8) if local else torch.empty(0, dtype=torch.uint8)


def _worker(rank, world, port, doc_list, ranges_per_gpu, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import multigpu
        t = _block(doc_list, rank, world)
        sb = multigpu.build_sharded(dist, t, CpuEngine(), ranges_per_gpu=ranges_per_gpu)
        assert sb.n_total == sum(len(d) + 1 for d in doc_list)
        sb.write(os.path.join(out_dir, "out"))
        sb.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ndocs,ranges", [(2, 41, 1), (2, 40, 3), (3, 50, 2), (3, 2, 1)])
def test_sharded_build_plumbing_over_gloo(world, ndocs, ranges, tmp_path):
    docs, _ = oracle.fasta_to_docs(cases.rnd_fasta(21 + ndocs, ndocs, 40))
    doc_list = docs.split(b"\0")[:-1]
    assert len(doc_list) == ndocs
    mp.spawn(_worker, args=(world, _free_port(), doc_list, ranges, str(tmp_path)), nprocs=world, join=True)
    got = open(tmp_path / "out.fmi", "rb").read()
    want = oracle.fmi_from_docs(docs)
    assert got == want, oracle.diff_fmi(got, want)


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("seed,kw", [(3, dict(nreads=37, maxlen=30)), (4, dict(nreads=300, maxlen=50, alpha="AC", genome=100)),
                                     (5, dict(nreads=9, maxlen=3, minlen=1)), (6, dict(nreads=150, maxlen=35, alpha="ACGT0123.", genome=400))])
def test_every_rank_writes_its_share_of_the_file(world, seed, kw, tmp_path):
    """All ranks in one process (multigpu.build_sharded_local): slices of very different sizes -- down to slices
    that fit inside one 64-bit word of a node, so that several ranks meet in one word and in one superblock --
    must still add up to the oracle's file, whatever the order in which the ranks write."""
    import multigpu
    docs, _ = oracle.fasta_to_docs(cases.rnd_fasta(seed, **kw))
    doc_list = docs.split(b"\0")[:-1]
    blocks = [_block(doc_list, r, world) for r in range(world)]
    sbs = multigpu.build_sharded_local(blocks, [CpuEngine() for _ in range(world)], ranges_per_gpu=1 + seed % 2)
    prefix = str(tmp_path / "local")
    for sb in reversed(sbs):  # header last: the file must not depend on the order
        sb.write(prefix)
        sb.close()
    got = open(prefix + ".fmi", "rb").read()
    want = oracle.fmi_from_docs(docs)
    assert got == want, oracle.diff_fmi(got, want)


def test_block_of_partitions_exactly():
    import multigpu
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [multigpu.block_of(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_slices_that_do_not_tile_are_rejected():
    import multigpu
    multigpu.check_tiling([0, 5], [5, 5], 10)
    with pytest.raises(RuntimeError):
        multigpu.check_tiling([0, 6], [5, 5], 10)  # a hole between the slices
    with pytest.raises(RuntimeError):
        multigpu.check_tiling([0, 5], [5, 4], 10)  # one suffix short


def test_text_plan_is_host_arithmetic():
    """dsmfm_text_plan_make needs no GPU: alphabet -> bits per symbol, equal slots of whole 128-byte lines."""
    import ctypes as C
    import dsmfm
    eng = CpuEngine()
    infos = [eng.block_stats(eng.open(torch.frombuffer(bytearray(d), dtype=torch.uint8), r, 3, 1))
             for r, d in enumerate([b"ACGT\0GG\0", b"T-A\0", b"NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN\0"])]
    plan = dsmfm.text_plan(infos)
    assert (plan.world, plan.bits, plan.n, plan.documents, plan.max_text_length) == (3, 3, 55, 4, 43)
    assert plan.slot_words == 16 and plan.text_bytes == (3 * 16 + 8) * 8
    assert [plan.block_bytes[r] for r in range(3)] == [8, 4, 43]
    bad = eng.block_stats(eng.open(torch.frombuffer(bytearray(b"AC\0\0"), dtype=torch.uint8), 0, 1, 1))
    with pytest.raises(dsmfm.DsmfmError):
        dsmfm.text_plan([bad])
