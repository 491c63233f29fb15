#!/usr/bin/env python
"""Regenerates tests/golden/ by running the UNMODIFIED reference `builder`
(oracle/_ref/builder, compiled in place by oracle/Makefile.ref) on the seeded
inputs of tests/cases.py.  Run in the build container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

Small cases are stored whole (<name>.fasta, <name>.fmi); larger ones only as
SHA-256 digests in manifest.json, together with the synthetic-generator case
(dsmgen parameters) used by the GPU tests.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "dsm-framework_b200"))
import cases  # noqa: E402
import oracle  # noqa: E402

GEN_CASES = {
    # a toydata-shaped sample small enough for the CPU suite: 20k x 100 bp
    "gen_20k": dict(seed=21, pool_seed=21, pool_size=4, n_genomes=4, genome_len=50_000, n_reads=20_000,
                    read_len=100, sub=0.005, pn=0.001),
    # high repetition: 100x coverage, error free
    "gen_rep_20k": dict(seed=22, pool_seed=22, pool_size=2, n_genomes=2, genome_len=10_000, n_reads=20_000,
                        read_len=100, sub=0.0, pn=0.0),
}


def main():
    assert oracle.have_reference(), "build the reference first: make -C oracle ref"
    manifest = {"reference": "HIITMetagenomics/dsm-framework (oracle/_ref/builder, stock flags)",
                "files": {}, "digests": {}, "generated": {}}
    with tempfile.TemporaryDirectory() as tmp:
        for name, fa in cases.golden_cases().items():
            fmi = oracle.reference_build(fa, tmp)
            with open(os.path.join(HERE, name + ".fasta"), "wb") as f:
                f.write(fa)
            with open(os.path.join(HERE, name + ".fmi"), "wb") as f:
                f.write(fmi)
            manifest["files"][name] = {"fasta_bytes": len(fa), "fmi_bytes": len(fmi),
                                       "fmi_sha256": hashlib.sha256(fmi).hexdigest()}
        # one case through `-s 32`: only the header's samplerate field changes
        fa = cases.golden_cases()["small_random"]
        fmi = oracle.reference_build(fa, tmp, samplerate=32)
        with open(os.path.join(HERE, "small_random.s32.fmi"), "wb") as f:
            f.write(fmi)
        for name, fa in cases.digest_cases().items():
            fmi = oracle.reference_build(fa, tmp)
            manifest["digests"][name] = {"fasta_sha256": hashlib.sha256(fa).hexdigest(), "fmi_bytes": len(fmi),
                                         "fmi_sha256": hashlib.sha256(fmi).hexdigest()}
        # the dormant SA sampling (FMIndex::saveSamples), driven by oracle/ref_driver.cpp over the reference classes
        manifest["sa"], manifest["sa_digests"] = {}, {}
        for name, rate in cases.SA_CASES:
            sa = oracle.reference_sa(cases.golden_cases()[name], tmp, samplerate=rate)
            fn = "%s.s%d.sa" % (name, rate)
            with open(os.path.join(HERE, fn), "wb") as f:
                f.write(sa)
            manifest["sa"][fn] = {"case": name, "samplerate": rate, "bytes": len(sa), "sha256": hashlib.sha256(sa).hexdigest()}
        for name, rate in cases.SA_DIGEST_CASES:
            sa = oracle.reference_sa(cases.digest_cases()[name], tmp, samplerate=rate)
            manifest["sa_digests"]["%s.s%d" % (name, rate)] = {"case": name, "samplerate": rate, "bytes": len(sa),
                                                            "sha256": hashlib.sha256(sa).hexdigest()}
        import dsmgen
        for name, kw in GEN_CASES.items():
            fa = dsmgen.fasta(**kw).tobytes()
            fmi = oracle.reference_build(fa, tmp)
            sa = oracle.reference_sa(fa, tmp)
            manifest["generated"][name] = {"params": kw, "fasta_sha256": hashlib.sha256(fa).hexdigest(),
                                           "fmi_bytes": len(fmi), "fmi_sha256": hashlib.sha256(fmi).hexdigest(),
                                           "sa_bytes": len(sa), "sa_sha256": hashlib.sha256(sa).hexdigest()}
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", len(manifest["files"]), "file cases,", len(manifest["digests"]), "digest cases,",
          len(manifest["generated"]), "generated cases")


if __name__ == "__main__":
    main()
