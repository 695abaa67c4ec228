"""Turn ncu outputs (read here, without a GPU) into the small text summaries kept under profiles/.

    python scripts/summarize_ncu.py launches  gpurun_out/launches.csv          > profiles/...launches.txt
    python scripts/summarize_ncu.py kernel    gpurun_out/prof.ncu-rep          > profiles/...kernel.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
        "sass__inst_executed_global_loads", "sass__inst_executed_global_stores", "inst_executed"]
STALLS = ['stall_barrier', 'stall_long_sb', 'stall_math', 'stall_mio', 'stall_short_sb', 'stall_wait',
          'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_branch_resolving', 'stall_lg', 'stall_no_inst']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split('(')[0][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', ''))
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none : per-kernel totals (cold-cache, serialised)")
    print("# %d launches, %.3f ms total" % (sum(a[0] for a in agg.values()), tot / 1e6))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-92s n=%4d %12.3f ms %6.2f%%" % (n, c, t / 1e6, 100 * t / tot))


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        print("## %s" % v[h.index("Kernel Name")][:120])
        for k in KEYS:
            if k in h:
                print("%-80s %-12s %s" % (k, u[h.index(k)], v[h.index(k)]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    if hi:
        h = rows[hi[0]]
        ix = {c: i for i, c in enumerate(h)}
        data = [r for r in rows[hi[0] + 1:] if len(r) == len(h)]
        agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in STALLS if s in ix}
        tot = sum(agg.values()) or 1
        print("## warp-state samples over the whole kernel (source page)")
        for s, c in sorted(agg.items(), key=lambda x: -x[1]):
            print("%-28s %10d %5.1f%%" % (s, c, 100.0 * c / tot))
        mx = max(int(r[ix['Instructions Executed']]) for r in data)
        print("## hot loop (instructions executed >= 40%% of the maximum)")
        for r in data:
            if int(r[ix['Instructions Executed']]) >= 0.4 * mx:
                print("%-52s samples=%s" % (r[ix['Source']].strip()[:52], r[ix['# Samples']]))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
