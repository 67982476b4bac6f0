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


def sharded_adjust(ctx, names, seqs, **kw):
    """Run one adjust round over this rank's shard and all-gather the accepted (name, consensus) pairs."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    mine = shard_indices(len(seqs), rank, world, [len(s) for s in seqs])
    res = ctx.adjust([seqs[i] for i in mine], **kw)
    recs = [(names[i], r["cons"]) for i, r in zip(mine, res) if r["is_te"]]
    return allgather_records(recs)
