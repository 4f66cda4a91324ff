"""Turn an ncu launch list (csv) and a --set full report into the text summaries committed under profiles/."""
import csv, subprocess, sys
from collections import defaultdict

def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = defaultdict(float), defaultdict(int)
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(row["Metric Unit"], 1e-6)
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_ms,avg_ms,share_pct\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write("%s,%d,%.4f,%.4f,%.2f\n" % (k, cnt[k], v, v / cnt[k], 100 * v / T))

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]

def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("metric,unit," + ",".join(r[idx["Kernel Name"]].split("(")[0].replace("void ", "") for r in rows[2:]) + "\n")
        for w in WANT:
            if w in idx:
                f.write(w + "," + units[idx[w]] + "," + ",".join(r[idx[w]] for r in rows[2:]) + "\n")

if __name__ == "__main__":
    launches(sys.argv[1], sys.argv[2])
    full(sys.argv[3], sys.argv[4])
