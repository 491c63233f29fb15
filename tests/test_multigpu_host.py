"""Host logic of the multi-GPU build (dsm-framework_b200/multigpu.py) on CPU: world_size 2 and 3 with the
gloo backend.  The device work is replaced by an engine backed by the oracle, so what is exercised is the
plumbing around it: uneven document blocks, text all-gather, slice order and tiling checks, assembly on
the root rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import oracle


class OracleEngine:
    """Stands in for the CUDA engine: suffix order and index come from oracle/ (test infrastructure)."""

    def __init__(self, shuffle_ranges=False):
        self.shuffle = shuffle_ranges

    def tensor_device(self):
        return torch.device("cpu")

    def sort_slice(self, text, shard_index, shard_count, shard_span):
        docs = text.numpy().tobytes()
        bwt = oracle.bwt(docs)
        n = len(docs)
        # any cut of [0, n) into shard_count contiguous ranges is a valid sharding
        cuts = [n * i // shard_count for i in range(shard_count + 1)]
        lo, hi = cuts[shard_index], cuts[shard_index + shard_span]
        return {"docs": docs, "bwt": bwt[lo:hi]}, lo, hi - lo

    def export_bwt(self, handle, out):
        out.copy_(torch.frombuffer(bytearray(handle["bwt"]), dtype=torch.uint8))

    def assemble(self, handle, bwt, n_total):
        assert bwt.numel() == n_total
        nd, maxlen = oracle.doc_stats(handle["docs"])
        handle["fmi"] = oracle.fmi_from_bwt(bwt.numpy().tobytes(), 124, nd, maxlen)
        return handle

    def close(self, handle):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, doc_list, ranges_per_gpu, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import multigpu
        b, e = multigpu.block_of(len(doc_list), rank, world)
        local = b"".join(d + b"\0" for d in doc_list[b:e])
        t = torch.frombuffer(bytearray(local), dtype=torch.uint8) if local else torch.empty(0, dtype=torch.uint8)
        handle, info = multigpu.build_sharded(dist, t, OracleEngine(), ranges_per_gpu=ranges_per_gpu)
        assert info["n_total"] == sum(len(d) + 1 for d in doc_list)
        assert info["block_bytes"][rank] == len(local)
        if rank == 0:
            with open(os.path.join(out_dir, "out.fmi"), "wb") as f:
                f.write(handle["fmi"])
        else:
            assert handle is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ndocs,ranges", [(2, 41, 1), (2, 40, 3), (3, 50, 2), (3, 2, 1)])
def test_sharded_build_plumbing_over_gloo(world, ndocs, ranges, tmp_path):
    docs, _ = oracle.fasta_to_docs(cases.rnd_fasta(21 + ndocs, ndocs, 40))
    doc_list = docs.split(b"\0")[:-1]
    assert len(doc_list) == ndocs
    mp.spawn(_worker, args=(world, _free_port(), doc_list, ranges, str(tmp_path)), nprocs=world, join=True)
    got = open(tmp_path / "out.fmi", "rb").read()
    assert got == oracle.fmi_from_docs(docs)


def test_block_of_partitions_exactly():
    import multigpu
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [multigpu.block_of(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1


def _gap_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import multigpu
        piece = torch.zeros(5, dtype=torch.uint8)
        # rank 1 claims to begin at 6: slices [0,5) and [6,11) leave a hole
        with pytest.raises(RuntimeError):
            multigpu.gather_slices(dist, piece, 0 if rank == 0 else 6, 10, torch.device("cpu"))
    finally:
        dist.destroy_process_group()


def test_slices_that_do_not_tile_are_rejected():
    mp.spawn(_gap_worker, args=(2, _free_port()), nprocs=2, join=True)
