#!/usr/bin/env python
"""Golden client streams of the mining step: what the UNMODIFIED reference
client `metaenumerate` (oracle/_ref/metaenumerate: EnumerateQuery::enumerate over ClientSocket) sends to a server
for a golden index, an enforced path and an fmin.  A recording TCP server stands in for `metaserver`.

    python tests/golden/make_stream_golden.py        # needs oracle/_ref (i.e. /root/reference at build time)

Writes tests/golden/streams.json: per case the SHA-256 and size of the byte stream (handshake 'S' name '.'
included) and, for the small cases, the stream itself (hex)."""
import hashlib
import json
import os
import socket
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CLIENT = os.path.join(ROOT, "oracle", "_ref", "metaenumerate")

CASES = [  # (golden index, enforced path, fmin, maxdepth or None)
    ("small_random", "A", 2, None), ("small_random", "C", 2, None), ("small_random", "GT", 2, None), ("small_random", "T", 3, 5),
    ("duplicates", "A", 2, None), ("duplicates", "CG", 2, None), ("poly_a", "A", 2, None), ("poly_a", "AAAA", 5, 30),
    ("two_letter", "A", 2, None), ("two_letter", "C", 4, None), ("reads100", "A", 2, None), ("reads100", "C", 2, None),
    ("reads100", "G", 10, None), ("reads100", "TT", 2, 40), ("one_base_reads", "A", 2, None), ("mixed_alphabet", "A", 2, None),
]
KEEP_BYTES_BELOW = 4096


def record(index, path, fmin, maxdepth):
    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]
    got = []

    def serve():
        conn, _ = srv.accept()
        while True:
            b = conn.recv(1 << 16)
            if not b:
                break
            got.append(b)
        conn.close()

    t = threading.Thread(target=serve)
    t.start()
    cmd = [CLIENT, "--fmin", str(fmin)] + (["--maxdepth", str(maxdepth)] if maxdepth else []) + [index]
    subprocess.run(cmd, input=("127.0.0.1\t%d\t%s\n" % (port, path)).encode(), check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    t.join()
    srv.close()
    return b"".join(got)


def main():
    out = {}
    for name, path, fmin, maxdepth in CASES:
        stream = record(os.path.join(HERE, name + ".fmi"), path, fmin, maxdepth)
        key = "%s:%s:f%d:m%s" % (name, path, fmin, maxdepth or 0)
        out[key] = {"index": name, "path": path, "fmin": fmin, "maxdepth": maxdepth or 0, "bytes": len(stream),
                    "sha256": hashlib.sha256(stream).hexdigest()}
        if len(stream) < KEEP_BYTES_BELOW:
            out[key]["hex"] = stream.hex()
        print(key, len(stream), "bytes")
    with open(os.path.join(HERE, "streams.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
