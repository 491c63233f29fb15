"""TEST INFRASTRUCTURE: a numpy / big-integer MODEL of the device work of the multi-GPU build, so that the host
logic of dsm-framework_b200/multigpu.py (phases, exchanges, slice tiling, file writing by every rank) runs on
CPU tensors under the gloo backend.  The suffix order comes from oracle/ (the CPU checker); the rules by which a
slice of the BWT becomes its share of the wavelet tree's bit vectors and BitRank directories -- who owns which
words, what travels in the edge records, how words and directory entries next to a slice boundary are completed --
are restated here independently of csrc/api.cu (dsmfm_pieces_build / merge / write), on Python integers.
Never imported by the product."""
import ctypes as C
import os
import struct

import numpy as np
import torch

import oracle

MASK64 = (1 << 64) - 1


def _codetable(counts):
    tab = (oracle.Code * 256)()
    arr = (C.c_uint64 * 256)(*[int(x) for x in counts])
    oracle.lib().dsm_oracle_codetable(arr, tab)
    return [(tab[c].count, tab[c].bits, tab[c].code) for c in range(256)]


def _shape(tab):
    """Pre-order node list of the Huffman-shaped tree (HuffWT.cpp:5-55, 73-86): dicts with leaf, ch, nbits,
    members (symbol -> branch bit) for internal nodes."""
    nodes = []

    def rec(prefix, level):
        mem = [c for c in range(256) if tab[c][0] and tab[c][1] >= level and (tab[c][2] & ((1 << level) - 1)) == prefix]
        if len(mem) == 1 and tab[mem[0]][1] == level:
            nodes.append({"leaf": 1, "ch": mem[0]})
            return
        nodes.append({"leaf": 0, "ch": 0, "nbits": sum(tab[c][0] for c in mem),
                      "members": {c: (tab[c][2] >> level) & 1 for c in mem}})
        rec(prefix, level + 1)
        rec(prefix | (1 << level), level + 1)
    rec(0, 0)
    return nodes


def _ceil(a, b):
    return -(-a // b)


class CpuEngine:
    def tensor_device(self):
        return torch.device("cpu")

    def open(self, local_docs, rank, world, ranges_per_gpu):
        return {"docs": local_docs.numpy().tobytes(), "rank": rank, "world": world, "k": max(1, int(ranges_per_gpu))}

    def block_stats(self, h):
        import dsmfm
        docs = h["docs"]
        info = dsmfm.BlockInfo()
        cnt = np.bincount(np.frombuffer(docs, dtype=np.uint8), minlength=256)
        for c in range(256):
            info.counts[c] = int(cnt[c])
        info.bytes = len(docs)
        parts = docs.split(b"\0")[:-1]
        assert not docs or docs[-1] == 0
        info.documents = len(parts)
        info.max_text_length = max((len(p) + 1 for p in parts), default=0)
        info.empty_document = 1 if any(len(p) == 0 for p in parts) else 0
        return bytes(info)

    def plan(self, infos):
        import dsmfm
        p = dsmfm.text_plan(infos)  # pure host arithmetic of the library
        slot = max(max(int(p.block_bytes[r]) for r in range(p.world)), 1)
        return p, slot * p.world, slot

    def new_text(self, text_bytes):
        return torch.zeros(text_bytes, dtype=torch.uint8)

    def block_pack(self, h, plan, rank, text):
        slot = text.numel() // plan.world
        d = h["docs"]
        text[rank * slot:rank * slot + len(d)] = torch.frombuffer(bytearray(d), dtype=torch.uint8) if d else torch.empty(0, dtype=torch.uint8)
        top = np.zeros(4096, dtype=np.uint64)
        top[0] = len(d)  # the model cuts the order itself; only the total is checked
        return top

    def build_packed(self, h, plan, text, top_sum):
        assert int(np.sum(top_sum)) == plan.n
        slot = text.numel() // plan.world
        t = text.numpy()
        docs = b"".join(t[r * slot:r * slot + int(plan.block_bytes[r])].tobytes() for r in range(plan.world))
        bwt = oracle.bwt(docs)
        n, sc = len(docs), h["world"] * h["k"]
        cuts = [n * i // sc for i in range(sc + 1)]  # any cut into contiguous ranges is a valid sharding
        lo, hi = cuts[h["rank"] * h["k"]], cuts[(h["rank"] + 1) * h["k"]]
        h.update(slice=bwt[lo:hi], n=n, counts=[int(plan.counts[c]) for c in range(256)], ntexts=int(plan.documents),
                 maxlen=int(plan.max_text_length))
        return lo, hi - lo

    def slice_hist(self, h):
        return np.bincount(np.frombuffer(h["slice"], dtype=np.uint8), minlength=256).astype(np.uint64)

    # ---- the share of the wavelet tree ----
    def pieces_build(self, h, hist_all, rank):
        import dsmfm
        world = hist_all.shape[0]
        tab = _codetable(h["counts"])
        nodes = _shape(tab)
        h["tab"], h["nodes"] = tab, nodes
        sl = np.frombuffer(h["slice"], dtype=np.uint8)
        pieces, edges = [], []
        for nd in nodes:
            if nd["leaf"]:
                continue
            mem = nd["members"]
            cnt = [sum(int(hist_all[q][c]) for c in mem) for q in range(world)]
            ones = [sum(int(hist_all[q][c]) for c in mem if mem[c]) for q in range(world)]
            o, c, B = sum(cnt[:rank]), cnt[rank], sum(ones[:rank])
            is_member = np.isin(sl, list(mem.keys()))
            seq = sl[is_member]
            bits = 0
            for i, s in enumerate(seq):  # own bits at their global positions
                if mem[int(s)]:
                    bits |= 1 << (o + i)
            word = lambda k: (bits >> (64 * k)) & MASK64
            before = lambda k: B + bin(bits & ((1 << (64 * k)) - 1)).count("1")  # ones in front of word k (k*64 >= o)
            last = c > 0 and not any(cnt[rank + 1:])
            nbits = nd["nbits"]
            pc = {"o": o, "c": c, "nbits": nbits, "ch": int(seq[0]) if c else 0}
            e = dsmfm.PieceEdge()
            if c:
                w0, w1 = _ceil(o, 64), (nbits // 64 + 1 if last else _ceil(o + c, 64))
                s0, s1 = _ceil(o, 256), (nbits // 256 + 1 if last else _ceil(o + c, 256))
                pc["w0"], pc["data"] = w0, [word(k) for k in range(w0, max(w0, w1))]
                pc["s0"], pc["Rs"] = s0, [before(4 * j) for j in range(s0, max(s0, s1))]
                # Rb relative to the superblock start; entries whose superblock starts in front of the slice are
                # completed in pieces_merge
                pc["Rb"] = [(before(k) - before(4 * (k // 4))) if 256 * (k // 4) >= o else 0 for k in range(w0, max(w0, w1))]
                e.count, e.first_word, e.last_word, e.ch = c, o // 64, (o + c - 1) // 64, pc["ch"]
                for t in range(4):
                    e.first[t] = word(o // 64 + t)
                    k = (o + c - 1) // 64 - 3 + t
                    e.last[t] = word(k) if k >= 0 else 0
            pieces.append(pc)
            edges.append(bytes(e))
        h["pieces"] = pieces
        held = sum(8 * len(p.get("data", [])) + 8 * len(p.get("Rs", [])) + len(p.get("Rb", [])) for p in pieces)
        return b"".join(edges), held

    def pieces_merge(self, h, edges_all, world):
        import dsmfm
        m = len(h["pieces"])
        rec = C.sizeof(dsmfm.PieceEdge)
        E = [[dsmfm.PieceEdge.from_buffer_copy(edges_all[(q * m + v) * rec:(q * m + v + 1) * rec]) for v in range(m)]
             for q in range(world)]
        rank = h["rank"]
        for v, pc in enumerate(h["pieces"]):
            contrib = {}  # word -> bits, from everybody's edge records
            for q in range(world):
                e = E[q][v]
                if not e.count:
                    continue
                for t in range(4):
                    contrib[e.first_word + t] = contrib.get(e.first_word + t, 0) | e.first[t]
                    if e.last_word - 3 + t >= 0:
                        contrib[e.last_word - 3 + t] = contrib.get(e.last_word - 3 + t, 0) | e.last[t]
            pc["node_ch"] = next((E[q][v].ch for q in range(world) if E[q][v].count), 0)
            if not pc["c"]:
                continue
            w0 = pc["w0"]
            for i in range(len(pc["data"])):
                others = 0
                for q in range(world):
                    e = E[q][v]
                    if q == rank or not e.count:
                        continue
                    for t in range(4):
                        if e.first_word + t == w0 + i:
                            others |= e.first[t]
                        if e.last_word - 3 + t == w0 + i:
                            others |= e.last[t]
                pc["data"][i] |= others
            merged = lambda w: pc["data"][w - w0] if w0 <= w < w0 + len(pc["data"]) else contrib.get(w, 0)
            for i in range(len(pc["Rb"])):
                k = w0 + i
                if 256 * (k // 4) < pc["o"]:
                    pc["Rb"][i] = sum(bin(merged(w)).count("1") for w in range(4 * (k // 4), k))
        h["merged"] = True

    def pieces_write(self, h, prefix, header):
        fd = os.open(prefix + ".fmi", os.O_WRONLY | os.O_CREAT, 0o644)
        try:
            pos = 1 + 8 + 4 + 2048 + 8 + 256 * 16
            it = iter(h["pieces"])
            for nd in h["nodes"]:
                if header:
                    ch = nd["ch"] if nd["leaf"] else None
                if nd["leaf"]:
                    if header:
                        os.pwrite(fd, bytes([1, nd["ch"]]), pos)
                    pos += 2
                    continue
                pc = next(it)
                nbits = nd["nbits"]
                if header:
                    os.pwrite(fd, bytes([0, pc["node_ch"]]) + struct.pack("<QQII", nbits, nbits // 64 + 1, 64, 256), pos)
                pos += 2 + 24
                if pc["c"]:
                    os.pwrite(fd, struct.pack("<%dQ" % len(pc["data"]), *pc["data"]), pos + 8 * pc["w0"])
                pos += 8 * (nbits // 64 + 1)
                if pc["c"]:
                    os.pwrite(fd, struct.pack("<%dQ" % len(pc["Rs"]), *pc["Rs"]), pos + 8 * pc["s0"])
                pos += 8 * (nbits // 256 + 1)
                if pc["c"]:
                    os.pwrite(fd, bytes(pc["Rb"]), pos + pc["w0"])
                pos += nbits // 64 + 1
            if header:
                C_tab, run = [], 0
                for c in range(256):
                    C_tab.append(run)
                    run += h["counts"][c]
                head = bytes([17]) + struct.pack("<QI", h["n"], 124) + struct.pack("<256Q", *C_tab) + struct.pack("<Q", 0)
                head += b"".join(struct.pack("<QII", *h["tab"][c]) for c in range(256))
                os.pwrite(fd, head, 0)
                os.pwrite(fd, struct.pack("<IQBBBI", h["ntexts"], h["maxlen"], 0, 0, 0, 0), pos)
                os.ftruncate(fd, pos + 19)
        finally:
            os.close(fd)

    def close(self, h):
        pass
