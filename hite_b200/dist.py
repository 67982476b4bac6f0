"""Multi-GPU plumbing for the path: candidates shard embarrassingly (reference fan-out: one worker per
candidate, module/Util.py:8141-8147; one Nextflow process per chunk, main.nf:627-648); the genome and
its index are replicated per GPU; the ONLY exchange is a variable-length all-gather of the emitted
(name, consensus) records in rank order before the merge that feeds get_nonRedundant_lib
(reference main.py:652-664, `cat >>`).  torch.distributed is plumbing: NCCL on GPUs, gloo in CPU tests.
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world, lengths=None):
    """Deal items to ranks.  With ``lengths`` the deal is longest-first round-robin (the reference's own
    longest-processing-time heuristic, module/Util.py:1771-1798); output order inside a rank keeps the
    original index order so results are deterministic."""
    if lengths is None:
        return list(range(rank, n_items, world))
    order = sorted(range(n_items), key=lambda i: (-lengths[i], i))
    return sorted(order[rank::world])


def allgather_records(records, device=None):
    """``records``: list of (name:str, seq:str).  Returns the concatenation over ranks, rank order."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(records)
    world = dist.get_world_size()
    payload = "".join("%s\t%s\n" % (n, s) for n, s in records).encode("ascii")
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    n = torch.tensor([len(payload)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    mx = max(1, max(int(s.item()) for s in sizes))
    buf = torch.zeros(mx, dtype=torch.uint8, device=dev)
    if payload:
        buf[:len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    merged = []
    for o, s in zip(out, sizes):
        text = bytes(o[:int(s.item())].cpu().numpy()).decode("ascii")
        for line in text.splitlines():
            name, seq = line.split("\t")
            merged.append((name, seq))
    return merged


def shard_by_cost(costs, rank, world):
    """Longest-processing-time deal on measured costs (e.g. loci x length of a first round: a stage runs the same candidates
    through several rounds, judge_TIR_transposons.py:29-38, and a candidate of a 5 000-copy family costs 60 x one of a 5-copy
    family of the same length): heaviest first, each to the currently least loaded rank.  Deterministic and identical on
    every rank; returns this rank's indices in ascending order."""
    import heapq
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0, r) for r in range(world)]
    mine = []
    for i in order:
        ld, r = heapq.heappop(heap)
        if r == rank:
            mine.append(i)
        heapq.heappush(heap, (ld + costs[i], r))
    return sorted(mine)


def shard_by_chunk(chunk_of, lengths, rank, world):
    """SURVEY 8(e) partitioning for a multi-chunk genome (reference main.nf:627-648: one process per chunk): chunk c goes to
    rank c % world, then the candidates are rebalanced -- every rank above the mean base count puts its shortest candidates
    into a pool until it is at the mean, and the pool is dealt longest-first to the currently least loaded rank.
    Deterministic, identical on every rank; returns this rank's candidate indices in ascending order."""
    import heapq
    n = len(lengths)
    by_rank = [[] for _ in range(world)]
    for i in range(n):
        by_rank[int(chunk_of[i]) % world].append(i)
    load = [sum(lengths[i] for i in b) for b in by_rank]
    mean = sum(load) / float(world)
    pool = []
    for r in range(world):
        by_rank[r].sort(key=lambda i: (-lengths[i], i))
        while by_rank[r] and load[r] - lengths[by_rank[r][-1]] >= mean:
            i = by_rank[r].pop()
            load[r] -= lengths[i]
            pool.append(i)
    pool.sort(key=lambda i: (-lengths[i], i))
    heap = [(load[r], r) for r in range(world)]
    heapq.heapify(heap)
    for i in pool:
        ld, r = heapq.heappop(heap)
        by_rank[r].append(i)
        heapq.heappush(heap, (ld + lengths[i], r))
    return sorted(by_rank[rank])


class RecordGather:
    """The path's one exchange, fused: every step each rank contributes ONE fixed-capacity buffer
    ``[n_records, n_bytes | (candidate index, consensus length) x n_records | consensus bytes]`` to a single
    ``all_gather_into_tensor`` -- the sizes travel in the buffer's own header, so there is no size exchange and no host
    synchronisation between collectives.  The capacity is a bound every rank derives by itself from the (deterministic)
    shards.  The collective is launched asynchronously and finished one step later (``finish``), so it overlaps the next
    step's sketch / seed kernels; rank 0 -- the host of the CPU merge that consumes the records (get_nonRedundant_lib,
    reference main.py:652-725) -- then reads the headers and only the used part of every rank's payload back.
    Works on NCCL (device tensors) and gloo (CPU tensors, the world-size-2 tests)."""

    HDR = 16

    def __init__(self, shard_sizes, shard_bases, device, read_back_rank=0):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())     # (a helper thread has no current device of its own)
        self.nmax = max(shard_sizes) if shard_sizes else 0
        self.rec_cap = 8 * self.nmax
        body = max(int(b * 1.25) + 64 * n for n, b in zip(shard_sizes, shard_bases)) if shard_sizes else 0
        self.cap = (self.HDR + self.rec_cap + body + 255) // 256 * 256
        self.buf = torch.zeros(self.cap, dtype=torch.uint8, device=self.device)
        self.out = torch.zeros(self.cap * self.world, dtype=torch.uint8, device=self.device)
        pin = self.device.type == "cuda"
        self.h_hdr = torch.zeros(self.HDR + self.rec_cap, dtype=torch.uint8, pin_memory=pin)
        self.h_out = torch.zeros(self.cap * self.world, dtype=torch.uint8, pin_memory=pin) if self.rank == read_back_rank else None
        self.read_back_rank = read_back_rank
        self.work = None
        self.bytes_sent = 0

    def start(self, cand_index, cons_len, cons):
        """``cand_index`` / ``cons_len``: int32 arrays of this rank's emitted records; ``cons``: their consensus bytes as a
        uint8 tensor on ``device`` (or anything ``torch.as_tensor`` wraps without a copy).  Launches the collective."""
        import numpy as np
        n = len(cand_index)
        body = torch.as_tensor(cons, device=self.device) if cons is not None else torch.zeros(0, dtype=torch.uint8, device=self.device)
        nb = int(body.numel())
        if n > self.nmax or self.HDR + self.rec_cap + nb > self.cap:
            raise RuntimeError("RecordGather: %d records / %d bytes exceed the capacity derived from the shards" % (n, nb))
        hdr = self.h_hdr.numpy()
        hdr[:16].view(np.int64)[:] = (n, nb)
        rec = hdr[16:16 + 8 * n].view(np.int32).reshape(n, 2)
        rec[:, 0] = cand_index
        rec[:, 1] = cons_len
        self.buf[:self.HDR + 8 * n].copy_(self.h_hdr[:self.HDR + 8 * n], non_blocking=True)
        if nb:
            self.buf[self.HDR + self.rec_cap:self.HDR + self.rec_cap + nb].copy_(body.view(torch.uint8).reshape(-1))
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()     # the library reuses its consensus buffer in the next hb_adjust
        self.bytes_sent = self.HDR + 8 * n + nb
        self._own = (np.array(cand_index, dtype=np.int32), np.array(cons_len, dtype=np.int32), n, nb)
        if self.world > 1:
            self.work = dist.all_gather_into_tensor(self.out, self.buf, async_op=True)
        else:
            self.work = None                              # one rank: the records are already where they are needed

    def finish_async(self, concat=False):
        """``finish`` on a helper thread, so that the read-back rank's host keeps launching the next round while the collective
        completes and its payload crosses PCIe; ``result()`` joins.  Nothing may call ``start`` before ``result``."""
        import threading
        box = {}

        def run():
            try:
                if self.device.type == "cuda":
                    torch.cuda.set_device(self.device)
                    if getattr(self, "_side", None) is None:
                        self._side = torch.cuda.Stream(device=self.device)
                    with torch.cuda.stream(self._side):   # the wait and the copies stay off the streams the next round uses
                        box["v"] = self.finish(concat)
                else:
                    box["v"] = self.finish(concat)
            except BaseException as e:                    # surfaced by result()
                box["e"] = e
        self._thread = (threading.Thread(target=run, daemon=True), box)
        self._thread[0].start()

    def result(self):
        th, box = self._thread
        th.join()
        self._thread = None
        if "e" in box:
            raise box["e"]
        return box["v"]

    def finish(self, concat=True):
        """Wait for the collective launched by ``start``; on the read-back rank returns (indices, lengths, consensus bytes)
        of ALL ranks' records in rank order as numpy arrays (``concat=False``: per-rank lists of zero-copy views into the
        pinned staging buffer, valid until the next ``finish``), elsewhere None.  With one rank nothing is exchanged or copied
        and the bytes come back as None (they are the caller's own hb_adjust result)."""
        import numpy as np
        if self.work is not None:
            self.work.wait()
            self.work = None
        if self.h_out is None:
            return None
        cap, hdr_all = self.cap, self.HDR + self.rec_cap
        if self.world == 1:                               # nothing was exchanged: hand the rank's own records back
            idx, ln, n, nb = self._own                    # (the consensus bytes are the caller's own hb_adjust result)
            return (idx, ln, None) if concat else ([idx], [ln], None)
        for r in range(self.world):                       # headers + record tables first (small) ...
            self.h_out[r * cap:r * cap + hdr_all].copy_(self.out[r * cap:r * cap + hdr_all], non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        ho = self.h_out.numpy()
        sizes = [ho[r * cap:r * cap + 16].view(np.int64).copy() for r in range(self.world)]
        for r in range(self.world):                       # ... then only the used consensus bytes
            nb = int(sizes[r][1])
            if nb:
                a = r * cap + hdr_all
                self.h_out[a:a + nb].copy_(self.out[a:a + nb], non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        idx, ln, cons = [], [], []
        for r in range(self.world):
            n, nb = int(sizes[r][0]), int(sizes[r][1])
            rec = ho[r * cap + 16:r * cap + 16 + 8 * n].view(np.int32).reshape(n, 2)
            idx.append(rec[:, 0]); ln.append(rec[:, 1])
            a = r * cap + hdr_all
            cons.append(ho[a:a + nb])                       # views into the pinned staging buffer: valid until the next finish()
        if not concat:
            return idx, ln, cons
        return np.concatenate(idx), np.concatenate(ln), np.concatenate(cons)


def sharded_adjust(ctx, names, seqs, **kw):
    """Run one adjust round over this rank's shard and all-gather the accepted (name, consensus) pairs."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    mine = shard_indices(len(seqs), rank, world, [len(s) for s in seqs])
    res = ctx.adjust([seqs[i] for i in mine], **kw)
    recs = [(names[i], r["cons"]) for i, r in zip(mine, res) if r["is_te"]]
    return allgather_records(recs)
