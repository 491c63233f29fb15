"""Profiling helper (not product code): builds through the GPU FASTA front end, for ncu / DSMFM_TRACE."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dsm-framework_b200"))
import torch
import dsmfm
import dsmgen

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kw = dict(dsmgen.CONFIGS[name])
fa = dsmgen.fasta(**kw)
host = torch.empty(fa.size, dtype=torch.uint8, pin_memory=True)
host.numpy()[:] = fa
for it in range(reps):
    b = dsmfm.Builder(device=0, stream=torch.cuda.current_stream().cuda_stream)
    info = b.append_fasta(host)
    b.build_device()
    s = b.stats()
    print("docs %d n=%d total %.1f launches %d" % (info["documents"], s.n, s.ms_total, s.kernel_launches))
    b.close()
