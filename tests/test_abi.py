"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/dsmfm.h declares, its
struct layouts match the Python mirror, and it fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dsmfm.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"DSMFM_API\s+[\w\s\*]+?\b(dsmfm_\w+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for s in ["dsmfm_create", "dsmfm_append", "dsmfm_append_batch", "dsmfm_finish", "dsmfm_write_fmi",
              "dsmfm_last_error", "dsmfm_destroy", "dsmfm_get_stats"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    import dsmfm
    L = dsmfm.lib()
    for s in declared_symbols():
        assert hasattr(L, s), "libdsmfm.so does not export %s" % s
    assert L.dsmfm_version() == 1


def test_struct_layouts_match_the_header(tmp_path):
    """Compile a tiny C program against include/dsmfm.h and compare sizeof/offsetof with ctypes."""
    import dsmfm
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "dsmfm.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(dsmfm_options), sizeof(dsmfm_code), sizeof(dsmfm_node),
         sizeof(dsmfm_index), sizeof(dsmfm_stats), offsetof(dsmfm_index, codetable), offsetof(dsmfm_index, nodes),
         offsetof(dsmfm_stats, ms_total), offsetof(dsmfm_stats, sort_pass_bytes), sizeof(dsmfm_fasta_info),
         offsetof(dsmfm_fasta_info, bad_headers));
  return 0; }
'''
    src = tmp_path / "layout.c"
    src.write_text(prog)
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(dsmfm.Options), C.sizeof(dsmfm.Code), C.sizeof(dsmfm.Node), C.sizeof(dsmfm.Index),
            C.sizeof(dsmfm.Stats), dsmfm.Index.codetable.offset, dsmfm.Index.nodes.offset,
            dsmfm.Stats.ms_total.offset, dsmfm.Stats.sort_pass_bytes.offset, C.sizeof(dsmfm.FastaInfo),
            dsmfm.FastaInfo.bad_headers.offset]
    assert got == want


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    import dsmfm
    with pytest.raises(dsmfm.DsmfmError) as e:
        dsmfm.Builder()
    assert e.value.code == dsmfm.ECUDA
    assert "no CPU fallback" in str(e.value)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_builder_cli_fails_loudly_without_gpu(tmp_path):
    exe = os.path.join(ROOT, "dsm-framework_b200", "builder")
    if not os.path.exists(exe):
        pytest.skip("builder CLI not built")
    fa = tmp_path / "x.fasta"
    fa.write_bytes(b">a\nACGT\n")
    r = subprocess.run([exe, str(fa)], capture_output=True, text=True)
    assert r.returncode == 1
    assert "no CPU fallback" in r.stderr
    assert not os.path.exists(str(fa) + ".fmi")


def test_product_does_not_reference_the_oracle():
    """The product tree must not include, link or import anything under oracle/."""
    prod = os.path.join(ROOT, "dsm-framework_b200")
    for dirpath, _, files in os.walk(prod):
        for fn in files:
            if fn.endswith((".cu", ".cuh", ".cpp", ".h", ".c", ".py")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle" not in text.lower(), "%s mentions the oracle" % os.path.join(dirpath, fn)
    out = subprocess.run(["ldd", os.path.join(prod, "libdsmfm.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
