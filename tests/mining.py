"""Runs the UNMODIFIED reference mining pipeline (oracle/_ref/metaserver x4 + metaenumerate per sample) on
loopback over a set of `.fmi` files and returns the servers' outputs.  TEST INFRASTRUCTURE (SURVEY.md section 4 /
appendix C): the consumers of the index this path writes, used as an end-to-end parity check."""
import os
import socket
import subprocess
import time

import oracle


def _free_ports(k):
    socks = []
    for _ in range(k):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        socks.append(s)
    ports = [s.getsockname()[1] for s in socks]
    for s in socks:
        s.close()
    return ports


def _gpu_streams(fmi_paths, hashes, fmin):
    """What metaenumerate would send, made on the GPU: {(sample, hash): handshake + dsmfm_searcher_enumerate stream}."""
    import dsmfm
    out = {}
    for n, path in fmi_paths.items():
        with dsmfm.Searcher(path) as s:
            for h in hashes:
                out[(n, h)] = b"S" + n.encode() + b"." + s.enumerate(h.encode(), fmin=int(fmin))
    return out


def mine(fmi_paths, workdir, emax="1.2", fmin="2", timeout=900, client="reference"):
    """fmi_paths: {sample name: path of <name>.<...>.fmi}.  The file's basename up to the first '.' must be the
    sample name (metaenumerate.cpp:79-88).  Returns {hash prefix: server stdout bytes}.
    client="gpu": the reference's metaenumerate processes are replaced by the GPU trie walk (dsmfm_searcher_enumerate);
    one thread per connection sends its stream, the unmodified metaserver processes do the rest."""
    ref = oracle.REF_DIR
    os.makedirs(workdir, exist_ok=True)
    names = sorted(fmi_paths)
    for n in names:
        assert os.path.basename(fmi_paths[n]).split(".")[0] == n
    names_txt = ("\n".join(names) + "\n").encode()
    hashes = ["A", "C", "G", "T"]
    ports = _free_ports(len(hashes))
    # a numeric address: the box's hostname / "localhost" need not resolve
    hosts = "".join("127.0.0.1\t%d\t%s\n" % (p, h) for p, h in zip(ports, hashes)).encode()
    servers, clients, outs = [], [], {}
    try:
        for p, h in zip(ports, hashes):
            out = open(os.path.join(workdir, "out.%s.txt" % h), "wb")
            err = open(os.path.join(workdir, "srv.%s.log" % h), "wb")
            sp = subprocess.Popen([os.path.join(ref, "metaserver"), "-p", str(p), "--emax", emax],
                                  stdin=subprocess.PIPE, stdout=out, stderr=err)
            sp.stdin.write(names_txt)
            sp.stdin.close()
            servers.append((sp, out, err))
        time.sleep(1.0)  # servers listen before the clients connect (wrapper-simple does the same)
        if client == "gpu":
            import threading
            streams = _gpu_streams(fmi_paths, hashes, fmin)
            errors = []

            def send(port, blob):
                try:
                    with socket.create_connection(("127.0.0.1", port)) as c:
                        c.sendall(blob)
                except Exception as e:  # noqa: BLE001
                    errors.append(e)
            threads = [threading.Thread(target=send, args=(p, streams[(n, h)])) for n in names for p, h in zip(ports, hashes)]
            for t in threads:
                t.start()
            for t in threads:
                t.join(timeout)
            assert not errors, errors
            names = []
        for n in names:
            log = open(os.path.join(workdir, "cli.%s.log" % n), "wb")
            cp = subprocess.Popen([os.path.join(ref, "metaenumerate"), "--fmin", fmin, fmi_paths[n]],
                                  stdin=subprocess.PIPE, stdout=log, stderr=subprocess.STDOUT)
            cp.stdin.write(hosts)
            cp.stdin.close()
            clients.append((cp, log))
        deadline = time.time() + timeout
        for cp, log in clients:
            rc = cp.wait(timeout=max(1, deadline - time.time()))
            log.close()
            assert rc == 0, "metaenumerate failed: " + open(log.name, "rb").read()[-2000:].decode(errors="replace")
        for sp, out, err in servers:
            rc = sp.wait(timeout=max(1, deadline - time.time()))
            out.close()
            err.close()
            assert rc == 0, "metaserver failed"
    finally:
        for p, *files in servers + clients:
            if p.poll() is None:
                p.kill()
    for h in hashes:
        with open(os.path.join(workdir, "out.%s.txt" % h), "rb") as f:
            outs[h] = f.read()
    if not any(outs.values()):  # nothing mined: show what the processes said
        logs = []
        for fn in sorted(os.listdir(workdir)):
            if fn.endswith(".log"):
                with open(os.path.join(workdir, fn), "rb") as f:
                    logs.append("== %s ==\n%s" % (fn, f.read()[-1500:].decode(errors="replace")))
        raise AssertionError("the mining run printed nothing\n" + "\n".join(logs))
    return outs
