"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the small text summaries committed here.
    python profiles/summarize.py <tag> <launches.csv> <full.ncu-rep> [<more.ncu-rep> ...]
"""
import collections
import csv
import subprocess
import sys

tag, launches = sys.argv[1:3]
reps = sys.argv[3:]
out = open(f"profiles/{tag}_ncu_summary.md", "w")
# ---- launch list: per-kernel share of the (serialised, cold-cache) device time
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
H = {n: i for i, n in enumerate(rows[hdr])}
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows[hdr + 1:]:
    if r[H["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[H["Kernel Name"]].split("(")[0].replace("void ", "").replace("admmnet::", "")
    v = float(r[H["Metric Value"]].replace(",", ""))
    unit = r[H["Metric Unit"]]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)      # -> ms
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
out.write(f"# {tag}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, serialised launches)\n\n")
out.write("| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|\n")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    out.write(f"| {k} | {cnt[k]} | {v:.1f} | {v / cnt[k]:.3f} | {100 * v / T:.1f} % |\n")
# ---- full capture: key metrics per kernel
def load(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    return rr[0], rr[1], rr[2:]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
# one launch per kernel name: the longest one (layer 0 launches the general kernels too, but there they return at
# once for every signal the arrowhead shortcut handled)
best = {}
for rep in reps:
    h, units, data = load(rep)
    I = {n: i for i, n in enumerate(h)}
    for r in data:
        name = r[I["Kernel Name"]].split("(")[0].replace("void ", "")
        dur = float(r[I["gpu__time_duration.sum"]].replace(",", ""))
        if name not in best or dur > best[name][0]:
            best[name] = (dur, r, I, units)
out.write(f"\n# {tag}: `ncu --set full --clock-control none` (one launch per kernel; layer kernels 4096 signals per launch, k_dc 2368, k_classic_p 65536)\n")
for name, (_, r, I, units) in best.items():
    out.write(f"\n## {name}\n\n| metric | value | unit |\n|---|---|---|\n")
    for w in want:
        if w in I:
            out.write(f"| {w} | {r[I[w]]} | {units[I[w]]} |\n")
out.close()
print(open(f"profiles/{tag}_ncu_summary.md").read()[:3000])
