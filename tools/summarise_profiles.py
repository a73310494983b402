"""Turns the raw ncu outputs in gpurun_out/ into the small, committed summaries under profiles/.

  python tools/summarise_profiles.py <ncu-rep> <out.txt>            # one --set full capture -> key metrics
  python tools/summarise_profiles.py --launches <launches.csv> <out.txt>   # gpu__time_duration launch list -> shares
"""
import csv, collections, json, os, re, subprocess, sys


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
    lines = ["# ncu --set full --clock-control none  (%s)" % os.path.basename(rep)]
    traffic = None
    for r in rows[2:]:
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                lines.append("%-85s %s %s" % (w, r[i], units[i]))
        def val(name):
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        traffic = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        lines.append("%-85s %.0f byte" % ("traffic = dram read + write per launch", traffic))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    return traffic


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"\(mgcmt::LevelDev.*|\(LevelDev.*", "", row["Kernel Name"]).replace("void ", "").replace("mgcmt::", "")
        key = (name[:70], row["Grid Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    o = ["# ncu --metrics gpu__time_duration.sum --clock-control none  (%s): per-launch times are cold-cache and" % os.path.basename(path),
         "# serialised -- read the SHARES.  total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())),
         "%-72s %-16s %6s %10s %9s %7s" % ("kernel", "grid", "n", "total_us", "avg_us", "share")]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        o.append("%-72s %-16s %6d %10.1f %9.2f %6.1f%%" % (k[0], k[1], c, t, t / c, 100 * t / tot))
    open(out, "w").write("\n".join(o) + "\n")
    print("\n".join(o[:30]))


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        t = full(sys.argv[1], sys.argv[2])
        if len(sys.argv) > 3:
            key = sys.argv[3]
            p = os.path.join(os.path.dirname(sys.argv[2]), "traffic.json")
            d = json.load(open(p)) if os.path.exists(p) else {}
            d[key] = t
            json.dump(d, open(p, "w"), indent=1)
