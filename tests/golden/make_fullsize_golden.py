#!/usr/bin/env python
"""Reference digests of the FULL-SIZE workloads (BASELINE.json configs / SURVEY.md section 8d):
runs the UNMODIFIED reference classes (oracle/_ref/ref_driver `fmi`: RLCSABuilder -> FMIndex -> save, with
incbwt's OpenMP sort; byte-identical to the stock `builder`) on the documents of a
dsmgen configuration and records size and SHA-256 of the `.fmi` it wrote in tests/golden/fullsize.json.

Run in the build container (needs oracle/_ref, i.e. /root/reference):

    python tests/golden/make_fullsize_golden.py C1 C3 [C5s ...]

C3 (1 Gbp, four 512 MiB batches merged by backward search) takes the reference tens of minutes.
The GPU tests (tests/test_gpu_fullsize.py) and bench.py compare the index they build with these digests.
"""
import hashlib
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "dsm-framework_b200"))
import dsmgen  # noqa: E402

OUT = os.path.join(HERE, "fullsize.json")
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")

# reduced shapes of the big configurations (same generator, fewer reads) the CPU finishes in minutes
EXTRA = {
    # C5 shape (4 genomes, error free, 200x coverage) at 1/16 of the reads and genome length
    "C5s": dict(seed=13, pool_seed=13, pool_size=4, n_genomes=4, genome_len=125_000, n_reads=1_000_000,
                read_len=100, sub=0.0, pn=0.0),
    # C4 shape (many genomes, 8x) at 1/160 of the reads: what the multi-GPU preflight builds
    "C4s": dict(seed=12, pool_seed=12, pool_size=2000, n_genomes=12, genome_len=1_000_000, n_reads=1_000_000,
                read_len=100, sub=0.005, pn=0.001),
}


def sha_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest(), os.path.getsize(path)


def main():
    names = sys.argv[1:] or ["C1"]
    assert os.path.exists(REF_DRIVER), "build the reference first: make -f oracle/Makefile.ref"
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    tmp = os.environ.get("TMPDIR", "/tmp")
    threads = os.cpu_count() or 1
    for name in names:
        kw = dsmgen.CONFIGS.get(name) or EXTRA[name]
        docs = dsmgen.docs(**kw)
        dpath = os.path.join(tmp, "fullsize_%s.docs" % name)
        docs.tofile(dpath)
        docs_sha = hashlib.sha256(docs).hexdigest()
        nbytes = docs.nbytes
        del docs
        prefix = os.path.join(tmp, "fullsize_%s" % name)
        t0 = time.time()
        out = subprocess.run([REF_DRIVER, "fmi", dpath, prefix, str(threads)], check=True, capture_output=True, text=True)
        dt = time.time() - t0
        sha, size = sha_file(prefix + ".fmi")
        res[name] = {"params": kw, "docs_bytes": nbytes, "docs_sha256": docs_sha, "fmi_bytes": size, "fmi_sha256": sha,
                     "reference": "oracle/_ref/ref_driver fmi (unmodified reference classes, OpenMP sort, %d threads)" % threads,
                     "reference_seconds": round(dt, 1), "reference_stdout": out.stdout.strip().splitlines()[-1]}
        os.remove(dpath)
        os.remove(prefix + ".fmi")
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)
        print(name, res[name]["fmi_sha256"], size, "bytes,", round(dt, 1), "s", flush=True)


if __name__ == "__main__":
    main()
