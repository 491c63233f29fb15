"""Run under torchrun (one rank per GPU): builds ONE index over documents spread across the ranks with the
CUDA engine and checks it bit-for-bit against the oracle on rank 0.  Used by test_gpu_multi.py."""
import os
import sys

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
    try:
        for seed, kw, ranges in [(1, dict(nreads=501, maxlen=100, genome=2000), 1),
                                 (2, dict(nreads=3000, maxlen=60, minlen=60, genome=500, dup=0.0, pn=0.0), 2),
                                 (4, dict(nreads=700, maxlen=45, alpha="A", genome=64), 1),
                                 (5, dict(nreads=300, maxlen=80, alpha="ACGT0123.", genome=900), 3)]:
            docs, nd = oracle.fasta_to_docs(cases.rnd_fasta(seed, **kw))
            doc_list = docs.split(b"\0")[:-1]
            b, e = multigpu.block_of(len(doc_list), rank, world)
            mine = b"".join(d + b"\0" for d in doc_list[b:e])
            host = torch.frombuffer(bytearray(mine), dtype=torch.uint8).pin_memory()
            engine = multigpu.CudaEngine(local, stream=torch.cuda.current_stream().cuda_stream)
            for src, wavelet in ((host, "distributed"), (host.cuda(), "distributed"), (host, "root")):
                handle, info = multigpu.build_sharded(dist, src, engine, ranges_per_gpu=ranges, wavelet=wavelet)
                assert info["n_total"] == len(docs)
                if rank == 0:
                    handle.fetch()
                    got = handle.fmi()
                    handle.close()
                    want = oracle.fmi_from_docs(docs)
                    assert got == want, "seed %d: sections %r differ" % (seed, oracle.diff_fmi(got, want))
        dist.barrier()
        if rank == 0:
            print("MULTIGPU_CHECK_OK world=%d lo_bits=%s" % (world, lo_bits))
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
