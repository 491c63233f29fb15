#!/bin/bash
# Timing helper (not product code): one cold `builder -v` run of a generated C3 FASTA with DSMFM_TRACE milestones.
cd "$(dirname "$0")/.."
python - <<'PY'
import sys, os
sys.path.insert(0, "dsm-framework_b200")
import dsmgen
kw = dict(dsmgen.CONFIGS["C3"])
dsmgen.fasta(**kw).tofile("/tmp/c3.fasta")
PY
for i in 1 2; do
  ( time env DSMFM_TRACE=1 ./dsm-framework_b200/builder -v /tmp/c3.fasta /tmp/c3out ) 2>&1 | grep -v "^Warning"
done
