import sys, time
sys.path.insert(0,'/root/repo')
import hite_b200
from workloads import synth
wl = synth.make_workload(genome_bp=100_000_000, n_contigs=10, n_families=1000, cands_per_family=10, copies=(10,40), seed=1)
ctx = hite_b200.Context(0)
ctx.load_genome(wl.contig_seqs); ctx.build_index()
for _ in range(3):
    t=time.perf_counter(); res = ctx.adjust(wl.cand_seqs)
    print('adjust total', time.perf_counter()-t, ctx.timing()['kernel_ms'], file=sys.stderr)
