"""Profiling helper (not product code): builds ONE key-range shard of a G-GPU collection on a single GPU,
so that `ncu` can list the launches of the sharded path.  usage: profile_shard.py <gpus> [shard]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dsm-framework_b200"))
import torch
import dsmfm
import dsmgen

G = int(sys.argv[1])
shard = int(sys.argv[2]) if len(sys.argv) > 2 else 0
parts = []
for r in range(G):
    kw = dict(dsmgen.CONFIGS["C3"])
    kw["seed"] += 1000 * r
    kw["pool_seed"] += 1000 * r
    t = torch.empty(kw["n_reads"] * 202, dtype=torch.uint8, pin_memory=True)
    dsmgen.docs(out=t, **kw)
    parts.append(t.cuda())
full = torch.cat(parts)
del parts
for it in range(2):
    b = dsmfm.Builder(device=0, stream=torch.cuda.current_stream().cuda_stream, expected_bytes=full.numel(),
                      shard_index=shard, shard_count=G)
    b.append_batch_device(full)
    b.build_device()
    s = b.stats()
    h = b.slice_hist()
    import numpy as np
    print("n=%d count=%d pack %.1f sort %.1f refine %.1f total %.1f launches %d" % (
        s.n, b.shard_info().count, s.ms_pack, s.ms_sort, s.ms_refine, s.ms_total, s.kernel_launches))
    b.close()
