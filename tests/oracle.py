"""ctypes binding of oracle/liboracle.so -- the CPU checker.  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


class Code(C.Structure):
    _fields_ = [("count", C.c_uint64), ("bits", C.c_uint32), ("code", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
        L = C.CDLL(LIB_PATH)
        L.dsm_oracle_transform.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        L.dsm_oracle_transform.restype = C.c_size_t
        L.dsm_oracle_fasta_to_docs.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                               C.POINTER(C.c_uint64)]
        L.dsm_oracle_bwt.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.dsm_oracle_codetable.argtypes = [C.POINTER(C.c_uint64), C.POINTER(Code)]
        L.dsm_oracle_codetable.restype = None
        L.dsm_oracle_fmi.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64,
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.dsm_oracle_build.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_size_t)]
        L.dsm_oracle_sa_file.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.dsm_oracle_free.argtypes = [C.c_void_p]
        L.dsm_oracle_free.restype = None
        L.dsm_oracle_index_open.argtypes = [C.c_void_p, C.c_size_t]
        L.dsm_oracle_index_open.restype = C.c_void_p
        L.dsm_oracle_index_close.argtypes = [C.c_void_p]
        L.dsm_oracle_index_close.restype = None
        L.dsm_oracle_index_length.argtypes = [C.c_void_p]
        L.dsm_oracle_index_length.restype = C.c_uint64
        L.dsm_oracle_rank.argtypes = [C.c_void_p, C.c_uint8, C.c_uint64]
        L.dsm_oracle_rank.restype = C.c_uint64
        L.dsm_oracle_lf.argtypes = [C.c_void_p, C.c_uint8, C.c_uint64]
        L.dsm_oracle_lf.restype = C.c_uint64
        L.dsm_oracle_access.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.dsm_oracle_access.restype = C.c_uint8
        _lib = L
    return _lib


def _take(out, n):
    b = C.string_at(out, n.value)
    lib().dsm_oracle_free(out)
    return b


def transform(read):
    out = C.create_string_buffer(2 * len(read) + 1)
    n = lib().dsm_oracle_transform(read, len(read), out)
    return out.raw[:n]


def fasta_to_docs(fasta):
    fasta = bytes(fasta)
    out, n, nd = C.c_void_p(), C.c_size_t(), C.c_uint64()
    rc = lib().dsm_oracle_fasta_to_docs(fasta, len(fasta), C.byref(out), C.byref(n), C.byref(nd))
    assert rc == 0
    return _take(out, n), nd.value


class Index:
    """The query side of a loaded .fmi as the consumers see it: FMIndex::LF / getL over HuffWT::rank / access."""

    def __init__(self, fmi):
        self._fmi = bytes(fmi)
        self._h = lib().dsm_oracle_index_open(self._fmi, len(self._fmi))
        assert self._h, "not an .fmi image"
        self.n = lib().dsm_oracle_index_length(self._h)

    def rank(self, c, i):
        return lib().dsm_oracle_rank(self._h, c, i % 2**64)

    def lf(self, c, i):
        return lib().dsm_oracle_lf(self._h, c, i % 2**64)

    def access(self, i):
        r = C.c_uint64()
        c = lib().dsm_oracle_access(self._h, i, C.byref(r))
        return c, r.value

    def close(self):
        if self._h:
            lib().dsm_oracle_index_close(self._h)
            self._h = None

    def __del__(self):
        self.close()


def bwt(text, want_sa=False):
    """text: '\\0'-terminated documents.  Returns the BWT (and the suffix array)."""
    import numpy as np
    text = bytes(text)
    n = len(text)
    out = C.create_string_buffer(n if n else 1)
    sa = np.empty(n, dtype=np.uint64) if want_sa else None
    rc = lib().dsm_oracle_bwt(text, n, out, sa.ctypes.data if want_sa else None)
    assert rc == 0, rc
    return (out.raw[:n], sa) if want_sa else out.raw[:n]


def fmi_from_bwt(bwt_bytes, samplerate, ntexts, maxlen):
    bwt_bytes = bytes(bwt_bytes)
    out, n = C.c_void_p(), C.c_size_t()
    rc = lib().dsm_oracle_fmi(bwt_bytes, len(bwt_bytes), samplerate, ntexts, maxlen, C.byref(out), C.byref(n))
    assert rc == 0
    return _take(out, n)


def doc_stats(docs):
    docs = bytes(docs)
    parts = docs.split(b"\0")[:-1]
    return len(parts), (max(len(p) for p in parts) + 1 if parts else 0)


def fmi_from_docs(docs, samplerate=124):
    """The InsertText..save path for already-transformed documents."""
    docs = bytes(docs)
    if len(docs) == 0:
        return fmi_from_bwt(b"\0", samplerate, 1, 1)  # TextCollectionBuilder.cpp:111-119
    nd, maxlen = doc_stats(docs)
    return fmi_from_bwt(bwt(docs), samplerate, nd, maxlen)


def sa_file_from_bwt(bwt_bytes, samplerate, ntexts, maxlen):
    """Bytes of the `.sa` file FMIndex::saveSamples writes for this index."""
    bwt_bytes = bytes(bwt_bytes)
    out, n = C.c_void_p(), C.c_size_t()
    rc = lib().dsm_oracle_sa_file(bwt_bytes, len(bwt_bytes), samplerate, ntexts, maxlen, C.byref(out), C.byref(n))
    assert rc == 0, rc
    return _take(out, n)


def sa_file_from_docs(docs, samplerate=124):
    docs = bytes(docs)
    if len(docs) == 0:
        return sa_file_from_bwt(b"\0", samplerate, 1, 1)
    nd, maxlen = doc_stats(docs)
    return sa_file_from_bwt(bwt(docs), samplerate, nd, maxlen)


def reference_sa(fasta, tmpdir, samplerate=None):
    """`.sa` bytes from the UNMODIFIED reference: builder [-s R], then its dormant FMIndex::saveSamples
    driven by oracle/_ref/ref_driver (oracle/ref_driver.cpp)."""
    reference_build(fasta, tmpdir, samplerate)
    base = os.path.join(str(tmpdir), "ref_input.fasta")
    subprocess.run([os.path.join(REF_DIR, "ref_driver"), "sa", base + ".fmi", base], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with open(base + ".sa", "rb") as f:
        return f.read()


def build(fasta, samplerate=0):
    """FASTA bytes -> .fmi bytes, the whole `builder <input>` path."""
    fasta = bytes(fasta)
    out, n = C.c_void_p(), C.c_size_t()
    rc = lib().dsm_oracle_build(fasta, len(fasta), samplerate, C.byref(out), C.byref(n))
    assert rc == 0
    return _take(out, n)


def have_reference():
    return os.path.exists(os.path.join(REF_DIR, "builder"))


def reference_build(fasta, tmpdir, samplerate=None):
    """Run the UNMODIFIED reference builder (oracle/_ref/builder) on FASTA bytes."""
    path = os.path.join(str(tmpdir), "ref_input.fasta")
    with open(path, "wb") as f:
        f.write(bytes(fasta))
    cmd = [os.path.join(REF_DIR, "builder")]
    if samplerate:
        cmd += ["-s", str(samplerate)]
    subprocess.run(cmd + [path], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with open(path + ".fmi", "rb") as f:
        return f.read()


def parse_fmi(b):
    """Split .fmi bytes into named sections (FMIndex.cpp:155-217) for readable diffs."""
    import struct
    o = 0
    sec = {}

    def take(name, n):
        nonlocal o
        sec[name] = b[o:o + n]
        o += n
    take("version", 1); take("n", 8); take("samplerate", 4); take("C", 2048); take("bwtEndPos", 8)
    take("codetable", 256 * 16)
    k = [0]

    def node():
        i = k[0]
        k[0] += 1
        leaf = b[o]
        take("node%d.leaf" % i, 1); take("node%d.ch" % i, 1)
        if leaf:
            return
        n, integers = struct.unpack_from("<QQ", b, o)
        take("node%d.hdr" % i, 24)
        take("node%d.data" % i, 8 * integers); take("node%d.Rs" % i, 8 * (n // 256 + 1)); take("node%d.Rb" % i, n // 64 + 1)
        node(); node()
    node()
    take("numberOfTexts", 4); take("maxTextLength", 8); take("flags", 3); take("rotationLength", 4)
    sec["trailing"] = b[o:]
    return sec


def diff_fmi(a, b):
    """Names of the sections in which two .fmi images differ."""
    if a == b:
        return []
    try:
        sa, sb = parse_fmi(a), parse_fmi(b)
    except Exception as e:  # malformed
        return ["unparseable: %r" % e]
    return [k for k in sa if sa.get(k) != sb.get(k)] + [k for k in sb if k not in sa]
