"""Turns the raw ncu outputs of a measurement run (tools/measure_round.sh) into the text summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/<tag>_launches.csv  > profiles/<tag>_ncu_launches_summary.txt
    python tools/summarize_ncu.py full     gpurun_out/<tag>_tc_prof.ncu-rep > profiles/<tag>_lift_step_tc_ncu.txt
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active
sm__inst_executed.avg.per_cycle_elapsed smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_pipe_tc_wavefronts_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
sm__throughput.avg.pct_of_peak_sustained_elapsed""".split()


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    col = {n: i for i, n in enumerate(rows[start])}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) < len(col) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        a = agg[r[col["Kernel Name"]][:70]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("%-70s %9s %12s %7s" % ("kernel", "launches", "total_us", "share"))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-70s %9d %12.1f %6.2f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    print("%-70s %9d %12.1f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")] for r in data]
    print("# kernels:", " | ".join(names))
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
    for m in METRICS + stall:
        if m in hdr:
            i = hdr.index(m)
            print("%-92s %s  [%s]" % (m, " | ".join(r[i] for r in data), units[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
