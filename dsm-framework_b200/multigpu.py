"""One FM-index over reads held by several GPUs of one box (BASELINE.json configs[3]).

One process per GPU (torchrun).  Rank r holds a contiguous block of the documents (document ids are
global: all documents of rank r precede those of rank r+1, exactly as if they had been inserted one after
the other into one TextCollectionBuilder).  The build:

  1. the raw blocks are all-gathered over NVLink (NCCL), so every GPU holds the whole text (1 byte per
     symbol; 180 GB of HBM holds the 32 GB of the 16 Gbp configuration many times over);
  2. every GPU packs the text and suffix-sorts ITS key ranges (dsmfm_options.shard_*): the suffixes are cut
     by their first 16 symbols into world x ranges_per_gpu ranges of equal population, which every GPU
     derives from the same histogram without talking to the others.  The refinement keys come from the
     replicated text, so no rank exchange is needed -- this replaces incbwt's merge by backward search
     (rlcsa_builder.cpp:245-318) and yields the same order, because the order is a property of the text;
  3. the BWT slices are contiguous pieces of the global BWT, in rank order.  Every GPU turns its slice into
     the bits it contributes to each node of the Huffman-shaped wavelet tree, already shifted to their
     global bit offset (known from an all-gather of the slices' 256-bin histograms), and sends these
     pieces (0.28 bytes per symbol) to rank 0 (NCCL point-to-point), which copies them into place, builds
     the BitRank directories and owns the finished index.  (Engines without piece support -- and
     wavelet="root" -- ship the BWT slices themselves and rank 0 builds the whole tree, dsmfm_assemble.)

torch.distributed is plumbing only.  `engine` abstracts the device work so that the host logic (block
offsets, uneven sizes, slice order) is testable with the gloo backend on CPU tensors.
"""
import os
import time

import torch


class CudaEngine:
    """The device work through the C ABI (libdsmfm.so)."""

    def __init__(self, device, stream=None, flags=0, samplerate=0):
        import dsmfm
        self._dsmfm = dsmfm
        self.device = device
        self.stream = stream
        self.flags = flags
        self.samplerate = samplerate

    def tensor_device(self):
        return torch.device("cuda", self.device)

    def open(self, text, shard_index, shard_count, shard_span):
        """text: uint8 CUDA tensor with the whole collection; it is copied into the builder (the caller may drop
        its tensor afterwards, which matters when the collection is tens of GB)."""
        b = self._dsmfm.Builder(device=self.device, stream=self.stream, expected_bytes=text.numel(), flags=self.flags,
                                samplerate=self.samplerate, shard_index=shard_index, shard_count=shard_count,
                                shard_span=shard_span)
        try:
            t0 = time.perf_counter()
            b.append_batch_device(text)
            torch.cuda.current_stream().synchronize()
            self._t_append = 1000 * (time.perf_counter() - t0)
        except Exception:
            b.close()
            raise
        return b

    def sort(self, b):
        """Returns (handle, rank_begin, count) of this builder's slice of the global suffix order."""
        try:
            t1 = time.perf_counter()
            b.build_device()
            t2 = time.perf_counter()
            info = b.shard_info()
        except Exception:
            b.close()
            raise
        self.last_walls = (self._t_append, 1000 * (t2 - t1))  # append, build (host wall, for DSMFM_MG_TRACE)
        return b, info.rank_begin, info.count

    def sort_slice(self, text, shard_index, shard_count, shard_span):
        return self.sort(self.open(text, shard_index, shard_count, shard_span))

    def export_bwt(self, handle, out):
        handle.shard_export(bwt_dst=out)

    def assemble(self, handle, bwt, n_total):
        handle.assemble(bwt, n_total)
        return handle

    def slice_hist(self, handle):
        return handle.slice_hist()

    def pieces_bytes(self, handle, hist_all, rank):
        return handle.pieces_bytes(hist_all, rank)

    def build_pieces(self, handle, hist_all, rank, out):
        handle.build_pieces(hist_all, rank, out)

    def assemble_pieces(self, handle, hist_all, pieces):
        handle.assemble_pieces(hist_all, pieces)
        return handle

    def stats(self, handle):
        return handle.stats()

    def close(self, handle):
        handle.close()


def block_of(n_items, rank, world):
    """Contiguous block [begin, end) of n_items for `rank` of `world` (sizes differ by at most one)."""
    q, r = divmod(n_items, world)
    begin = rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


def _all_gather_sizes(dist, value, device):
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    out = torch.empty(dist.get_world_size(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t)
    return [int(x) for x in out.cpu()]


def gather_text(dist, local, device):
    """All-gathers the ranks' raw blocks (uint8, possibly of different sizes) into the whole text."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = _all_gather_sizes(dist, local.numel(), device)
    n = sum(sizes)
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    full = torch.empty(n, dtype=torch.uint8, device=device)
    mine = full[offs[rank]:offs[rank + 1]]
    mine.copy_(local, non_blocking=True)  # host->device for a pinned host block, device->device otherwise
    if world > 1:
        if len(set(sizes)) == 1:
            dist.all_gather_into_tensor(full, mine)
        else:
            for r in range(world):
                if sizes[r]:
                    dist.broadcast(full[offs[r]:offs[r + 1]], src=r)
    return full, sizes


def check_tiling(dist, rank_begin, count, n_total, device):
    """The ranks' slices must tile the suffix order [0, n_total) in rank order, without gaps."""
    counts = _all_gather_sizes(dist, count, device)
    begins = _all_gather_sizes(dist, rank_begin, device)
    pos = 0
    for r in range(dist.get_world_size()):
        if counts[r] and begins[r] != pos:
            raise RuntimeError("BWT slices do not tile the suffix order: rank %d begins at %d, expected %d"
                               % (r, begins[r], pos))
        pos += counts[r]
    if pos != n_total:
        raise RuntimeError("BWT slices cover %d of %d suffixes" % (pos, n_total))
    return counts


def gather_to_root(dist, piece, sizes, device, root=0):
    """Concatenates the ranks' uint8 buffers (sizes[r] bytes each) on `root` in rank order; None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    assert piece.numel() == sizes[rank]
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    if rank == root:
        full = torch.empty(offs[-1], dtype=torch.uint8, device=device)
        full[offs[rank]:offs[rank + 1]].copy_(piece)
        ops = [dist.P2POp(dist.irecv, full[offs[r]:offs[r + 1]], r) for r in range(world) if r != root and sizes[r]]
    else:
        full = None
        ops = [dist.P2POp(dist.isend, piece, root)] if sizes[rank] else []
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return full


def gather_slices(dist, piece, rank_begin, n_total, device, root=0):
    """Collects the ranks' BWT slices on `root` in global rank order.  Returns the whole BWT there, None elsewhere."""
    counts = check_tiling(dist, rank_begin, piece.numel(), n_total, device)
    return gather_to_root(dist, piece, counts, device, root)


def all_gather_hist(dist, hist, device):
    """hist: numpy uint64[256] of this rank -> numpy uint64[world, 256]."""
    import numpy as np
    t = torch.from_numpy(hist.astype(np.int64)).to(device)
    out = torch.empty(dist.get_world_size() * 256, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy().astype(np.uint64).reshape(dist.get_world_size(), 256)


def build_sharded(dist, local_docs, engine, ranges_per_gpu=1, root=0, wavelet="distributed"):
    """Builds ONE index over the documents of all ranks (rank order = document order).

    local_docs: uint8 tensor (pinned host or device) with this rank's '\\0'-terminated documents.
    wavelet: "distributed" (every GPU builds its pieces of the wavelet tree) or "root" (BWT slices go to
    the root, which builds the whole tree).
    Returns (handle, info): on `root` handle is the engine's builder holding the assembled index
    (fetch()/fmi()/save() as for a single-GPU build), None elsewhere; info has the slice layout.
    """
    world, rank = dist.get_world_size(), dist.get_rank()
    device = engine.tensor_device()
    trace = _Trace(device) if os.environ.get("DSMFM_MG_TRACE") else None
    full, sizes = gather_text(dist, local_docs, device)
    if trace: trace.mark("gather_text")
    n = full.numel()
    k = max(1, int(ranges_per_gpu))
    if hasattr(engine, "open"):
        opened = engine.open(full, rank * k, world * k, k)
        big = full.numel() > (8 << 30)
        del full  # the builder holds its own copy now
        if big and device.type == "cuda":
            torch.cuda.empty_cache()  # hand the gathered text's memory back before the sort buffers are allocated
        handle, rank_begin, count = engine.sort(opened)
    else:
        handle, rank_begin, count = engine.sort_slice(full, rank * k, world * k, k)
        del full
    if trace: trace.mark("sort_slice")
    if trace and hasattr(engine, "last_walls"): trace.marks.append(("(append %.1f build %.1f)" % engine.last_walls, 0.0))
    info = {"n_total": n, "block_bytes": sizes, "rank_begin": rank_begin, "count": count}
    distributed = wavelet == "distributed" and hasattr(engine, "build_pieces")
    if distributed:
        check_tiling(dist, rank_begin, count, n, device)
        hist_all = all_gather_hist(dist, engine.slice_hist(handle), device)
        if int(hist_all[rank].sum()) != count:
            raise RuntimeError("slice histogram of rank %d does not match its slice" % rank)
        psizes = [engine.pieces_bytes(handle, hist_all, r) for r in range(world)]
        piece = torch.empty(psizes[rank], dtype=torch.uint8, device=device)
        engine.build_pieces(handle, hist_all, rank, piece)
        if trace: trace.mark("build_pieces")
        gathered = gather_to_root(dist, piece, psizes, device, root)
    else:
        piece = torch.empty(count, dtype=torch.uint8, device=device)
        engine.export_bwt(handle, piece)
        gathered = gather_slices(dist, piece, rank_begin, n, device, root)
    if trace: trace.mark("gather")
    if rank != root:
        if hasattr(engine, "stats"):
            info["stats"] = engine.stats(handle)
        engine.close(handle)
        if trace: trace.report(rank, info)
        return None, info
    if distributed:
        engine.assemble_pieces(handle, hist_all, gathered)
    else:
        engine.assemble(handle, gathered, n)
    if trace: trace.mark("assemble")
    if hasattr(engine, "stats"):
        info["stats"] = engine.stats(handle)
    if trace: trace.report(rank, info)
    return handle, info


class _Trace:
    """DSMFM_MG_TRACE=1: wall time of every phase (with a device synchronize in between) on stderr."""

    def __init__(self, device):
        import time
        self.time = time
        self.cuda = device.type == "cuda"
        self.t = self._now()
        self.marks = []

    def _now(self):
        if self.cuda:
            torch.cuda.synchronize()
        return self.time.perf_counter()

    def mark(self, name):
        t = self._now()
        self.marks.append((name, 1000 * (t - self.t)))
        self.t = t

    def report(self, rank, info):
        import sys
        s = info.get("stats")
        extra = ""
        if s is not None:
            extra = " | build: pack %.1f sort %.1f refine %.1f wt %.1f total %.1f, wall %.1f of which alloc %.1f, count %d" % (
                s.ms_pack, s.ms_sort, s.ms_refine, s.ms_wt, s.ms_total, s.ms_wall_build, s.ms_wall_alloc, info["count"])
        print("[multigpu rank %d] " % rank + " ".join("%s %.1f ms" % m for m in self.marks) + extra, file=sys.stderr)
