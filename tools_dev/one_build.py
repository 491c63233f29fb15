"""Profiling helper (not product code): one or two unsharded builds of a workload, for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dsm-framework_b200"))
import torch
import dsmfm
import dsmgen

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kw = dict(dsmgen.CONFIGS[name])
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0  # fewer reads from proportionally shorter genomes (same coverage)
if scale != 1.0:
    kw["n_reads"] = int(kw["n_reads"] * scale)
    kw["genome_len"] = int(kw["genome_len"] * scale)
t = torch.empty(kw["n_reads"] * (2 * kw["read_len"] + 2), dtype=torch.uint8, pin_memory=True)
dsmgen.docs(out=t, **kw)
d = t.cuda()
for it in range(reps):
    b = dsmfm.Builder(device=0, stream=torch.cuda.current_stream().cuda_stream, expected_bytes=d.numel())
    b.append_batch_device(d)
    b.build_device()
    s = b.stats()
    print("n=%d pack %.1f sort %.1f (pass %.2f) refine %.1f wt %.1f total %.1f launches %d active %s members %.3f fetches %.3f" % (
        s.n, s.ms_pack, s.ms_sort, s.ms_sort_pass, s.ms_refine, s.ms_wt, s.ms_total, s.kernel_launches,
        [round(s.active[r] / s.n, 3) for r in range(s.rounds)], s.refine_members / s.n, s.refine_key_fetches / s.n))
    b.close()
