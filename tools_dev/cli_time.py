"""Timing helper (not product code): whole-CLI wall time of `builder` on a generated FASTA file, GPU front end
vs the per-read host loop.  usage: cli_time.py [config]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dsm-framework_b200"))
import dsmgen

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
kw = dict(dsmgen.CONFIGS[name])
fa = "/tmp/cli_%s.fasta" % name
t0 = time.time()
dsmgen.fasta(**kw).tofile(fa)
print("generated %s: %.1f MB in %.1f s" % (fa, os.path.getsize(fa) / 1e6, time.time() - t0), flush=True)
exe = os.path.join(ROOT, "dsm-framework_b200", "builder")
for args in (["-v"], ["-v"], ["-v", "--host-parse"]):
    t0 = time.time()
    r = subprocess.run([exe] + args + [fa, "/tmp/cli_out"], capture_output=True, text=True)
    dt = time.time() - t0
    rep = [l for l in r.stderr.splitlines() if "GPU build" in l or "Creating new index" in l or "Saving" in l or "Skipping" in l]
    print("builder %s: rc %d wall %.2f s  (%.0f Mbp/s)" % (" ".join(args), r.returncode, dt, kw["n_reads"] * kw["read_len"] / dt / 1e6), flush=True)
    for l in rep:
        print("    " + l)
    print("    .fmi bytes", os.path.getsize("/tmp/cli_out.fmi"))
