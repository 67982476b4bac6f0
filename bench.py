#!/usr/bin/env python
"""bench.py -- HiTE dynamic boundary adjustment on B200: candidates boundary-adjusted per second.

One "step" = one ``flank_region_align_v5``-equivalent round (reference module/Util.py:8032) over the
workload's candidate batch: copy search (seed + banded SW), star MSA, sparse-column removal, column
vote and boundary call.  Workload at N=1 = BASELINE.json configs[1] ("synthetic 100 Mbp genome, 10k
planted TIR candidates, 1xB200"); at N>1 the genome and candidate set scale with N (weak scaling:
N x 100 Mbp replicated on every GPU, 10k candidates per rank, one NCCL all-gather of the emitted
consensus records per step -- the only collective on the path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--small]

Prints ONE JSON line (rank 0).  ``--impl reference`` times the CPU implementation of the same path
(the plain-C port under oracle/, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def workload_for(world, small=False):
    from workloads import synth
    if small:
        return synth.make_workload(genome_bp=4_000_000 * world, n_contigs=2, n_families=40 * world, cands_per_family=10,
                                   copies=(10, 40), seed=1), "synthetic 4 Mbp x N, 400 TIR candidates per GPU (--small)"
    wl = synth.make_workload(genome_bp=100_000_000 * world, n_contigs=10, n_families=1000 * world, cands_per_family=10,
                             copies=(10, 40), length=(200, 2000), seed=1)
    return wl, "synthetic %d Mbp genome, %d planted TIR candidates (10k per GPU), %dxB200" % (
        100 * world, len(wl.cand_seqs), world)


def algo_bytes_pair(L):
    """Algorithmic bytes of one candidate x copy banded extension:
    2-bit+N query and target window in, 4 B per query position of edit record out, 56 B result."""
    return (2 * L + 64) * 3.0 / 8.0 + 4.0 * L + 56.0


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def usable_cores():
    """Host cores this process may actually use: min(cpu_count, affinity, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max",):
        try:
            q, per = open(path).read().split()
            if q != "max":
                n = min(n, max(1, int(int(q) / int(per))))
        except Exception:
            pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0:
            n = min(n, max(1, q // per))
    except Exception:
        pass
    return n


def cpu_reference_run(wl, cand_idx, threads, G=None):
    """The CPU implementation of the path (plain-C port, oracle/): per candidate copy search + star MSA
    + sparse-column removal + judge on ``threads`` host threads (pthreads inside the port, no Python in
    the loop).  Index build is outside the timed region (as for the GPU arm)."""
    from oracle import oracle
    if G is None:
        G = oracle.Genome(wl.contig_seqs)
    cands = [wl.cand_seqs[i] for i in cand_idx]
    G.run_batch(cands[:max(1, threads)], threads)
    t0 = time.perf_counter()
    out = G.run_batch(cands, threads)
    dt = time.perf_counter() - t0
    return dt, int(out[:, 0].sum()), G


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hite_b200")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        print(json.dumps({"error": "launch with torch.distributed.run for --gpus > 1"}))
        return 2
    cores = usable_cores()

    # ------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        wl, wname = workload_for(1, args.small)
        per_step = min(len(wl.cand_seqs), max(512, 128 * cores))
        times = []
        G = None
        for step in range(args.warmup + args.steps):
            idx = [(step * per_step + i) % len(wl.cand_seqs) for i in range(per_step)]
            dt, _, G = cpu_reference_run(wl, idx, cores, G)
            if step >= args.warmup:
                times.append(dt)
        dt = float(np.mean(times))
        val = per_step / dt
        line = {"impl": "reference", "metric": "candidate TEs boundary-adjusted/sec", "value": val, "unit": "candidates/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": wname, "sample": "%d candidates per step" % per_step},
                "cpu_baseline": {"value": val, "unit": "candidates/s", "cores": cores, "kind": "port",
                                 "sample": "%d candidates of the same workload per step, %d threads (plain-C port of the path, "
                                           "oracle/; the reference's own aligners minimap2/mafft are not installed)" % (per_step, cores)},
                "e2e": {"value": val, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------ GPU arm
    import torch
    import torch.distributed as dist
    import hite_b200
    if not torch.cuda.is_available():
        raise hite_b200.HiteB200Error("bench.py needs a CUDA device: there is no CPU fallback on the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl, wname = workload_for(world, args.small)
    my = list(range(rank, len(wl.cand_seqs), world))            # embarrassingly sharded candidates
    my_cands = [wl.cand_seqs[i] for i in my]
    ctx = hite_b200.Context(local)
    t0 = time.perf_counter()
    t_load = ctx.load_genome(wl.contig_seqs)
    t_idx = ctx.build_index(15, 10)
    setup_s = time.perf_counter() - t0
    info = ctx.genome_info()
    # coarse-boundary stage on the same genome (module/coarse_boundary.py): all-vs-all self-join + FMEA
    coarse = None
    if rank == 0:
        tc = time.perf_counter()
        iv = ctx.coarse_boundary(seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64, threads=max(1, cores))
        tmc = ctx.timing()
        tcw = time.perf_counter() - tc
        # GPU stages (event time) + host FMEA (wall) of the one call
        one = (tmc["kernel_ms"] + tmc["k_ms"][3]) * 1e-3
        coarse = {"intervals": int(len(iv)), "hsps": int(tmc["k_ms"][6]), "seed_pairs_M": float(tmc["k_ms"][7]),
                  "gpu_ms": tmc["kernel_ms"], "fmea_host_ms": tmc["k_ms"][3], "wall_s": tcw,
                  "Mbp_per_s": wl.genome_bp / 1e6 / max(1e-9, one)}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    my_arr = np.asarray(my, dtype=np.int64)

    def gather_results(res):
        """The path's only exchange: all-gather of the emitted records (candidate index, consensus length,
        consensus bytes) before the get_nonRedundant_lib merge (reference main.py:652-664 cat >>).  The
        consensus bytes are gathered straight from the library's device buffer over NVLink; rank 0 (the
        CPU merge's host) then reads the merged payload back."""
        calls, cons = res
        te = calls["is_te"] != 0
        if world == 1:
            return int(te.sum())
        head = torch.from_numpy(np.stack([my_arr[te], calls["cons_len"][te].astype(np.int64)], axis=1).reshape(-1))
        dev = ctx.last_cons_device()
        body = torch.as_tensor(dev, device="cuda") if dev is not None else torch.zeros(0, dtype=torch.uint8, device="cuda")
        n = torch.tensor([head.numel(), body.numel()], device="cuda", dtype=torch.int64)
        sz_out = gbuf("sz", world * 2, torch.int64)
        dist.all_gather_into_tensor(sz_out, n)
        sz = sz_out.cpu().numpy().reshape(world, 2)
        hmax = (int(sz[:, 0].max()) + 1023) // 1024 * 1024
        bmax = (int(sz[:, 1].max()) + (1 << 20) - 1) >> 20 << 20
        hbuf, bbuf = gbuf("h", hmax, torch.int64), gbuf("b", bmax, torch.uint8)
        hbuf[:head.numel()].copy_(head.pin_memory(), non_blocking=True)
        bbuf[:body.numel()].copy_(body)
        hout, bout = gbuf("ho", world * hmax, torch.int64), gbuf("bo", world * bmax, torch.uint8)
        dist.all_gather_into_tensor(hout, hbuf)
        dist.all_gather_into_tensor(bout, bbuf)
        if rank == 0:                                   # the merge consumer is a CPU stage: pinned read-back
            hh, bh = gbuf("hh", world * hmax, torch.int64, host=True), gbuf("bh", world * bmax, torch.uint8, host=True)
            hh.copy_(hout, non_blocking=True)
            bh.copy_(bout, non_blocking=True)
            torch.cuda.synchronize()
            lens = sum(int(hh[r * hmax:r * hmax + int(sz[r, 0])][1::2].sum()) for r in range(world))
            assert lens == int(sz[:, 1].sum())
        return int(sz[:, 0].sum() // 2)

    _bufs = {}

    def gbuf(name, n, dtype, host=False):
        """grow-only scratch tensors (device, or pinned host) so a step allocates nothing"""
        t = _bufs.get(name)
        if t is None or t.numel() < n:
            t = torch.empty(max(n, 1), dtype=dtype, pin_memory=True) if host else torch.empty(max(n, 1), dtype=dtype, device="cuda")
            _bufs[name] = t
        return t[:n]

    flat = np.frombuffer(b"".join(my_cands) + b"\0", dtype=np.uint8)
    offs = np.zeros(len(my_cands) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(c) for c in my_cands])
    host_cands = (flat, offs)                                   # the step's input, in host memory
    for _ in range(args.warmup):
        res = ctx.adjust_raw(host_cands, "tir", 1)
        gather_results(res)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    dev_ms, ext_ms, stage = [], [], np.zeros(8)
    h2d = d2h = launches = 0
    n_te_total = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = ctx.adjust_raw(host_cands, "tir", 1)
        tm = ctx.timing()
        n_te_total = gather_results(res)
        dev_ms.append(tm["kernel_ms"]); ext_ms.append(tm["k_ms"][2]); stage += np.array(tm["k_ms"])
        h2d, d2h, launches = tm["h2d_bytes"], tm["d2h_bytes"], tm["launches"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    # max over ranks
    stats = torch.tensor([wall, float(np.sum(dev_ms))], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    wall_max, dev_max = float(stats[0]), float(stats[1]) * 1e-3
    n_total = len(wl.cand_seqs)
    copies = ctx.last_copies()
    lens = np.array([len(c) for c in my_cands])
    pair_L = lens[copies["cand"]]
    algo_bytes = float(np.sum(algo_bytes_pair(pair_L)))
    ext_s = float(np.mean(ext_ms)) * 1e-3
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic, alu_pct = None, None     # dram read+write bytes / ALU-pipe utilisation of the dominant kernel, one `ncu --set full` capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            if not args.small and world == 1:
                traffic = tj["extend16_kernel<16>"]["dram_bytes_per_launch"]
                alu_pct = tj["extend16_kernel<16>"].get("alu_pipe_pct_of_peak")
        except Exception:
            pass
        achieved = algo_bytes / ext_s / 1e9 if ext_s > 0 else 0.0
        dp_cells = float(stage[7] / args.steps) * 1e6 * 32
        line = {
            "metric": "candidate TEs boundary-adjusted/sec", "value": n_total * args.steps / dev_max, "unit": "candidates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_max / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": wname, "genome_bp": wl.genome_bp, "candidates": n_total, "k": 15, "w": 10,
                       "max_loci": 128, "band": 64, "l2": "inputs larger than L2 (packed genome+index %.0f MB, traceback stream %.1f GB/step)"
                       % ((info["total_padded"] * 0.375 + info["n_minimizers"] * 8) / 1e6, float(stage[7] / args.steps) * 1e6 * 16 / 1e9),
                       "parallelism": "candidates sharded over %d GPU(s), genome+index replicated, 1 NCCL all-gather/step" % world},
            "e2e": {"value": n_total * args.steps / wall_max, "unit": "candidates/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches) * args.steps,
            "clocks": clocks,
            "roofline": {"kernel": "extend16_kernel + traceback_kernel (banded SW, two pairs per lane; hb_extend16.cuh)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "note": "integer-ALU-bound DP, not HBM-bound: %.2f GCUPS, ALU pipe %s %% of peak (ncu, profiles/); "
                                 "algorithmic bytes/pair=(2L+64)*3/8+4L+56"
                                 % (dp_cells / ext_s / 1e9 if ext_s else 0, "%.1f" % alu_pct if alu_pct else "n/a"),
                         "kernel_ms": ext_s * 1e3, "pairs": int(len(copies)),
                         "binding_resource": {"name": "integer ALU pipe (sm__inst_executed_pipe_alu)", "pct_of_peak": alu_pct,
                                              "source": "ncu --set full, profiles/r1_g_full_extend16_C2.md"}},
            "stages_ms": {k: float(v / args.steps) for k, v in zip(
                ("sketch_cands", "seed_sort_cluster", "extend", "copy_rules_msa", "clean_judge", "compact"), stage[:6])},
            "setup": {"genome_load_ms": t_load["kernel_ms"] + t_load["h2d_ms"], "index_build_ms": t_idx["kernel_ms"],
                      "index_Mbp_per_s": wl.genome_bp / 1e6 / max(1e-9, t_idx["kernel_ms"] * 1e-3), "setup_wall_s": setup_s,
                      "n_minimizers": int(info["n_minimizers"])},
            "calls": {"is_te": int(n_te_total)},
            "coarse_boundary": coarse,
            "mbp_per_s": coarse["Mbp_per_s"] if coarse else None,
        }
        if world == 1 and not args.small:
            # the metric is reported per TE type.  Helitron / non-LTR candidates take the same
            # copy search + MSA and their own judges (judge_boundary_v6 / v9); 30 Mbp genome, 3000 candidates each.
            from workloads import synth
            per_type = {"tir": float(n_total * args.steps / dev_max)}
            for kind in ("helitron", "non_ltr"):
                wk = synth.make_workload(genome_bp=30_000_000, n_contigs=3, n_families=300, cands_per_family=10, copies=(10, 40),
                                         length=(200, 2000), seed=2, kind=kind)
                ctx.load_genome(wk.contig_seqs)
                ctx.build_index(15, 10)
                fk = np.frombuffer(b"".join(wk.cand_seqs) + b"\0", dtype=np.uint8)
                ok_ = np.zeros(len(wk.cand_seqs) + 1, dtype=np.uint64)
                ok_[1:] = np.cumsum([len(c) for c in wk.cand_seqs])
                ms = []
                for it in range(2 + 5):
                    ck, _ = ctx.adjust_raw((fk, ok_), kind, 1)
                    if it >= 2:
                        ms.append(ctx.timing()["kernel_ms"])
                tmk = ctx.timing()
                per_type[kind] = float(len(wk.cand_seqs) / (np.mean(ms) * 1e-3))
                per_type[kind + "_stages_ms"] = [round(float(x), 3) for x in tmk["k_ms"][:6]]
                per_type[kind + "_is_te"] = int((ck["is_te"] != 0).sum())
            line["per_type_candidates_per_s"] = per_type
            # the TSD-bounded variant generation that runs right before the path in the TIR module
            # (search_confident_tir_v4, module/Util.py:7734) on this step's flanked candidates, through the C-ABI
            tc = [c for c in my_cands if len(c) > 100]
            ms, wl_s, nrec = [], [], 0
            for it in range(2 + 5):
                tw = time.perf_counter()
                _, cnt = ctx.tsd_candidates(tc, [51] * len(tc), [len(c) - 50 for c in tc], 50, 1)
                if it >= 2:
                    wl_s.append(time.perf_counter() - tw)
                    ms.append(ctx.timing()["kernel_ms"])
                    nrec = int(cnt.sum())
            line["tir_prefilter"] = {"candidates": len(tc), "variants": nrec, "kernel_ms": float(np.mean(ms)),
                                     "candidates_per_s": float(len(tc) / (np.mean(ms) * 1e-3)),
                                     "e2e_candidates_per_s": float(len(tc) / np.mean(wl_s))}
            # the non-LTR candidate generator's scan (search_polyA_TSD, module/Util.py:10915)
            ms, wl_s = [], []
            for it in range(2 + 5):
                tw = time.perf_counter()
                pr = ctx.nonltr_candidates(tc, 50)
                if it >= 2:
                    wl_s.append(time.perf_counter() - tw)
                    ms.append(ctx.timing()["kernel_ms"])
            line["nonltr_prefilter"] = {"candidates": len(tc), "with_tsd": int(pr["found"].sum()), "kernel_ms": float(np.mean(ms)),
                                        "candidates_per_s": float(len(tc) / (np.mean(ms) * 1e-3)),
                                        "e2e_candidates_per_s": float(len(tc) / np.mean(wl_s))}
            # known-TE masking of the genome (mask_genome_intactTE, module/Util.py:6389) with
            # the planted families' consensus sequences as the library, and the homology copy search of
            # judge_Other_transposons.py on the same records (get_query_copies chaining)
            from hite_b200 import boundary as hb_boundary, masking as hb_masking, other as hb_other
            fams = {}
            for t in wl.truth:
                fams.setdefault("fam%d#DNA/TIR" % t["family"], t["consensus"].decode())
            ctx.load_genome(wl.contig_seqs)
            ctx.build_index(15, 10)
            ctx.contig_names = list(wl.contig_names)
            lib_names = list(fams.keys())
            tq = time.perf_counter()
            rec = hb_other.query_hsp_records(ctx, lib_names, fams)
            t_hsp = time.perf_counter() - tq
            hsp_tm = dict(hb_other.query_hsp_records.last)
            tq = time.perf_counter()
            copies = hb_other.copies_from_records(rec, lib_names, fams, ctx.contig_names, 0.95, max(1, cores))
            t_chain_other = time.perf_counter() - tq
            tq = time.perf_counter()
            iv = hb_masking.intervals_from_hsp_records(rec, lib_names, fams, 0.95, max(1, cores))
            t_chain = time.perf_counter() - tq
            fresh = ctx.mask_genome([i[1] for i in iv], [i[2] for i in iv], [i[3] for i in iv])
            t_fill = ctx.timing()["kernel_ms"]
            t_re = ctx.build_index(15, 10)["kernel_ms"]
            line["known_te_masking"] = {
                "library_seqs": len(fams), "library_bp": int(sum(len(s) for s in fams.values())), "hsps": hsp_tm["hsps"],
                "pairs_extended": hsp_tm["pairs"], "hsp_device_ms": hsp_tm["kernel_ms"], "hsp_wall_ms": t_hsp * 1e3,
                "full_length_copies": len(iv), "chain_host_ms": t_chain * 1e3, "mask_fill_kernel_ms": t_fill,
                "masked_bp": int(fresh), "reindex_ms": t_re,
                "genome_Mbp_per_s": wl.genome_bp / 1e6 / max(1e-9, t_hsp + t_chain + (t_fill + t_re) * 1e-3)}
            line["homology_copy_search"] = {"queries": len(fams), "copies": int(sum(len(v) for v in copies.values())),
                                            "hsp_device_ms": hsp_tm["kernel_ms"], "chain_host_ms": t_chain_other * 1e3,
                                            "queries_per_s": len(fams) / max(1e-9, t_hsp + t_chain_other)}
        if not args.no_cpu_baseline and world == 1:
            per = min(len(wl.cand_seqs), max(512, 128 * cores))
            dt, _, _ = cpu_reference_run(wl, list(range(per)), cores)
            line["cpu_baseline"] = {"value": per / dt, "unit": "candidates/s", "cores": cores, "kind": "port",
                                    "sample": "first %d candidates of the same workload, %d threads, %.1f s (plain-C port of the "
                                              "path under oracle/)" % (per, cores, dt)}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
