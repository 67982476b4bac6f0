#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into committed text under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md "<command>"
  python tools/ncu_summary.py full gpurun_out/prof_r1.ncu-rep profiles/r1_full.md "<command>"
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "launch__grid_size", "launch__block_size", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]


def launches(path, out, cmd):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v * 1e3 if u in ("s", "second") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write("command: `%s`\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n" % cmd)
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| %s | %d | %.3f | %.1f%% |\n" % (k, v[0], v[1], 100 * v[1] / tot))
        f.write("| **total** | %d | %.3f | 100%% |\n" % (sum(v[0] for v in agg.values()), tot))


def full(path, out, cmd):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none (one capture per kernel; per launch)\n\ncommand: `%s`\n\n" % cmd)
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "")
            f.write("## %s\n\n| metric | value | unit |\n|---|---:|---|\n" % name)
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("| %s | %s | %s |\n" % (k, r[i], units[i]))
            try:
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
                ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                rd = float(r[ir].replace(",", "")) * mult[units[ir]]
                wr = float(r[iw].replace(",", "")) * mult[units[iw]]
                f.write("| traffic (dram read+write) | %.0f | byte |\n" % (rd + wr))
            except Exception:
                pass
            f.write("\n")


def table(path, out, cmd, peak=6540.5):
    """One row per captured launch from `ncu -i x.ncu-rep --page raw --csv` output (a .csv, made on the GPU box so the .ncu-rep
    need not travel): duration, DRAM bytes, achieved DRAM GB/s and its fraction of the MEASURED HBM peak, pipe / issue / occupancy."""
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tm = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}

    def g(r, k):
        try:
            i = hdr.index(k)
            return float(r[i].replace(",", "")), units[i]
        except Exception:
            return float("nan"), ""
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none, one row per captured launch\n\ncommand: `%s`\n\n" % cmd)
        f.write("achieved = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum; peak = %.1f GB/s (MEASURED_PEAKS.json hbm_gbs)\n\n" % peak)
        f.write("| kernel | ms | DRAM read MB | DRAM write MB | achieved GB/s | of measured HBM | DRAM % | SM % | ALU pipe % | issue % | warps active % | regs | grid |\n")
        f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
            t, u = g(r, "gpu__time_duration.sum"); t *= tm.get(u, 1)
            rd, u1 = g(r, "dram__bytes_read.sum"); rd *= mult.get(u1, 1)
            wr, u2 = g(r, "dram__bytes_write.sum"); wr *= mult.get(u2, 1)
            gbs = (rd + wr) / 1e9 / (t * 1e-3) if t else 0.0
            f.write("| %s | %.3f | %.1f | %.1f | %.0f | %.3f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %d |\n" % (
                name, t, rd / 1e6, wr / 1e6, gbs, gbs / peak, g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")[0],
                g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")[0],
                g(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")[0],
                g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")[0],
                g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")[0], g(r, "launch__registers_per_thread")[0],
                g(r, "launch__grid_size")[0]))


if __name__ == "__main__":
    {"launches": launches, "full": full, "table": table}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
