"""GPU: HB-cdhit end to end through the drop-in's function (sequence set as the genome, hb_query_hsps as the all-vs-all):
same clusters as with the oracle's alignments, planted families intact."""
import pytest

from test_cluster_cpu import OracleBackedCtx, _families

pytestmark = pytest.mark.gpu


def test_cd_hit_est_gpu_equals_oracle_backed(ctx, tmp_path):
    from hite_b200 import cluster
    recs, fam_of = _families(seed=9)
    inp = str(tmp_path / "in.fa")
    with open(inp, "w") as f:
        for n, s in recs:
            f.write(">%s\n%s\n" % (n, s))
    a, b = str(tmp_path / "gpu.fa"), str(tmp_path / "cpu.fa")
    n_gpu = cluster.cd_hit_est(inp, a, ctx=ctx)
    n_cpu = cluster.cd_hit_est(inp, b, ctx=OracleBackedCtx())
    assert n_gpu == n_cpu == 12
    assert open(a).read() == open(b).read() and open(a + ".clstr").read() == open(b + ".clstr").read()
