"""GPU, 2 ranks over NCCL (gpurun --gpus 2; skipped on a one-GPU box): hite_b200.dist on the real backend -- the candidates of a
round sharded over two processes with one context each, the emitted records exchanged by the fused RecordGather and by
allgather_records -- must give exactly what one GPU gives; the sharded coarse stage's parts concatenate to the whole."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch
import torch.distributed as dist
import hite_b200
from hite_b200 import dist as hd
from workloads import synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = synth.make_workload(genome_bp=3_000_000, n_contigs=3, n_families=24, cands_per_family=4, seed=29)
names, seqs = wl.cand_names, [s.decode() for s in wl.cand_seqs]
ctx = hite_b200.Context(local)
ctx.load_genome(wl.contig_seqs, wl.contig_names)
ctx.build_index(15, 10)
# (1) the Python-level exchange
merged = hd.sharded_adjust(ctx, names, seqs, te_type="tir", plant=1)
# (2) the fused exchange on this rank's shard
lens = [len(s) for s in seqs]
shards = [hd.shard_indices(len(seqs), r, world, lens) for r in range(world)]
mine = shards[rank]
g = hd.RecordGather([len(s) for s in shards], [sum(lens[i] for i in s) for s in shards], "cuda")
calls, cons = ctx.adjust_raw([seqs[i] for i in mine], "tir", 1)
te = calls["is_te"] != 0
g.start(np.asarray(mine, dtype=np.int32)[te], calls["cons_len"][te], ctx.last_cons_device())
got = g.finish()
# (3) the coarse stage, one part per rank, intervals gathered through the same Python-level exchange
part = ctx.coarse_boundary(seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64, threads=2, part=rank, n_parts=world)
iv = hd.allgather_records([("%%d:%%d-%%d" %% (int(o["chr"]), int(o["start"]), int(o["end"])), "x") for o in part])
if rank == 0:
    single = ctx.adjust(seqs, "tir", 1)
    exp = [(names[i], single[i]["cons"]) for r in range(world) for i in shards[r] if single[i]["is_te"]]
    assert merged == exp and len(exp) > 40, (len(merged), len(exp))
    idx, ln, body = got
    assert idx.tolist() == [i for r in range(world) for i in shards[r] if single[i]["is_te"]]
    assert body.tobytes().decode() == "".join(c for _, c in exp) and ln.tolist() == [len(c) for _, c in exp]
    whole = ctx.coarse_boundary(seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64, threads=2)
    assert [n for n, _ in iv] == ["%%d:%%d-%%d" %% (int(o["chr"]), int(o["start"]), int(o["end"])) for o in whole] and len(whole) > 20
ctx.close()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharded_round_over_nccl_equals_one_gpu(tmp_path):
    from hite_b200 import _lib
    if _lib.load().hb_device_count() < 2:
        pytest.skip("one GPU visible")
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
