"""ctypes binding of libdsmgen.so: the seeded synthetic read generator.
Input preparation only -- nothing here is on the measured path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdsmgen.so")


class Params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("pool_seed", C.c_uint64), ("pool_size", C.c_uint32),
                ("n_genomes", C.c_uint32), ("genome_len", C.c_uint64), ("n_reads", C.c_uint64),
                ("read_len", C.c_uint32), ("reserved", C.c_uint32), ("sub", C.c_double), ("pn", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libdsmgen.so is missing: run `make -C dsm-framework_b200 libdsmgen.so`")
        L = C.CDLL(LIB_PATH)
        L.dsmgen_fasta_size.argtypes = [C.POINTER(Params)]
        L.dsmgen_fasta_size.restype = C.c_uint64
        L.dsmgen_fasta.argtypes = [C.POINTER(Params), C.c_void_p, C.c_uint64]
        L.dsmgen_reads.argtypes = [C.POINTER(Params), C.c_void_p]
        L.dsmgen_docs.argtypes = [C.POINTER(Params), C.c_void_p]
        _lib = L
    return _lib


# The workloads of BASELINE.json `configs` / SURVEY.md section 8(d).
CONFIGS = {
    # name: seed, pool_seed, pool_size, n_genomes, genome_len, n_reads, read_len, sub, pn
    "C1": dict(seed=1, pool_seed=1, pool_size=10, n_genomes=10, genome_len=250_000, n_reads=250_000, read_len=100, sub=0.005, pn=0.001),
    "C3": dict(seed=11, pool_seed=11, pool_size=200, n_genomes=200, genome_len=1_000_000, n_reads=10_000_000, read_len=100, sub=0.005, pn=0.001),
    "C4": dict(seed=12, pool_seed=12, pool_size=2000, n_genomes=2000, genome_len=1_000_000, n_reads=160_000_000, read_len=100, sub=0.005, pn=0.001),
    "C5": dict(seed=13, pool_seed=13, pool_size=4, n_genomes=4, genome_len=2_000_000, n_reads=16_000_000, read_len=100, sub=0.0, pn=0.0),
}
for _i in range(1, 6):  # C2: five samples drawing 10 of a shared pool of 16 genomes
    CONFIGS["C2-%d" % _i] = dict(seed=_i, pool_seed=99, pool_size=16, n_genomes=10, genome_len=250_000,
                                 n_reads=250_000, read_len=100, sub=0.005, pn=0.001)


def params(**kw):
    return Params(reserved=0, **kw)


def fasta(**kw):
    p = params(**kw)
    n = lib().dsmgen_fasta_size(C.byref(p))
    if n == 0:
        raise ValueError("bad generator parameters")
    out = np.empty(n, dtype=np.uint8)
    if lib().dsmgen_fasta(C.byref(p), out.ctypes.data, n) != 0:
        raise ValueError("bad generator parameters")
    return out


def reads(**kw):
    p = params(**kw)
    out = np.empty((p.n_reads, p.read_len), dtype=np.uint8)
    if lib().dsmgen_reads(C.byref(p), out.ctypes.data) != 0:
        raise ValueError("bad generator parameters")
    return out


def docs(out=None, **kw):
    """Documents as InsertText receives them ('\\0'-terminated, 2L+2 bytes each), into `out` if given."""
    p = params(**kw)
    nbytes = p.n_reads * (2 * p.read_len + 2)
    if out is None:
        out = np.empty(nbytes, dtype=np.uint8)
        addr = out.ctypes.data
    elif hasattr(out, "data_ptr"):
        assert out.numel() * out.element_size() >= nbytes
        addr = out.data_ptr()
    else:
        assert out.nbytes >= nbytes
        addr = out.ctypes.data
    if lib().dsmgen_docs(C.byref(p), addr) != 0:
        raise ValueError("bad generator parameters")
    return out
