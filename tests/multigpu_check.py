"""Run under torchrun (one rank per GPU): builds ONE index over documents spread across the ranks with the
CUDA engine -- every rank writes its share of the `.fmi` file -- and checks the file bit-for-bit against the
oracle on rank 0.  Fresh data in every iteration (nothing can be right by reusing a buffer of the iteration
before).  Used by test_gpu_multi.py; DSMFM_CHECK_LARGE=1 adds a case of 400k reads per rank whose reference
digest is unknown but which must equal the same collection built on ONE GPU (rank 0)."""
import hashlib
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dsm-framework_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import cases
    import multigpu
    import oracle
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lo_bits = os.environ.get("DSMFM_POS_LO_BITS")
    tmp = tempfile.mkdtemp(prefix="dsmfm_mg_") if rank == 0 else None
    box = [tmp]
    dist.broadcast_object_list(box, src=0)
    tmp = box[0]
    try:
        it = 0
        for seed, kw, ranges in [(1, dict(nreads=501, maxlen=100, genome=2000), 1),
                                 (2, dict(nreads=3000, maxlen=60, minlen=60, genome=500, dup=0.0, pn=0.0), 2),
                                 (4, dict(nreads=700, maxlen=45, alpha="A", genome=64), 1),
                                 (5, dict(nreads=300, maxlen=80, alpha="ACGT0123.", genome=900), 3),
                                 (6, dict(nreads=3, maxlen=5), 1)]:
            docs, nd = oracle.fasta_to_docs(cases.rnd_fasta(seed, **kw))
            doc_list = docs.split(b"\0")[:-1]
            b, e = multigpu.block_of(len(doc_list), rank, world)
            mine = b"".join(d + b"\0" for d in doc_list[b:e])
            host = torch.frombuffer(bytearray(mine), dtype=torch.uint8).pin_memory() if mine else torch.empty(0, dtype=torch.uint8)
            for src in (host, host.cuda()):
                engine = multigpu.CudaEngine(local)
                sb = multigpu.build_sharded(dist, src, engine, ranges_per_gpu=ranges)
                assert sb.n_total == len(docs)
                prefix = os.path.join(tmp, "case%d" % it)
                it += 1
                sb.write(prefix)
                sb.close()
                dist.barrier()
                if rank == 0:
                    got = open(prefix + ".fmi", "rb").read()
                    want = oracle.fmi_from_docs(docs)
                    assert got == want, "seed %d: sections %r differ" % (seed, oracle.diff_fmi(got, want))
        if os.environ.get("DSMFM_CHECK_LARGE"):
            import dsmfm
            import dsmgen
            kw = dict(dsmgen.CONFIGS["C3"], n_reads=400_000, genome_len=40_000)
            blocks = []
            for r in range(world):
                k = dict(kw, seed=kw["seed"] + 1000 * r, pool_seed=kw["pool_seed"] + 1000 * r)
                blocks.append(k)
            mine = torch.from_numpy(dsmgen.docs(**blocks[rank])).pin_memory()
            sb = multigpu.build_sharded(dist, mine, multigpu.CudaEngine(local), ranges_per_gpu=2)
            prefix = os.path.join(tmp, "large")
            sb.write(prefix)
            sb.close()
            dist.barrier()
            if rank == 0:
                h = hashlib.sha256(open(prefix + ".fmi", "rb").read()).hexdigest()
                with dsmfm.Builder(device=local) as b1:
                    for r in range(world):
                        b1.append_batch(dsmgen.docs(**blocks[r]))
                    b1.finish()
                    want = hashlib.sha256(b1.fmi()).hexdigest()
                assert h == want, "large case: the %d-GPU index differs from the 1-GPU index" % world
        dist.barrier()
        if rank == 0:
            print("MULTIGPU_CHECK_OK world=%d lo_bits=%s" % (world, lo_bits))
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
