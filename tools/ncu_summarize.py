"""Turn the CSVs of tools/ncu_capture.sh into the markdown kept under profiles/ (run locally)."""
import csv
import sys
from collections import OrderedDict

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def launch_list():
    rows = list(csv.reader(l for l in open("%s/launches_%s.csv" % (src, tag)) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | mean us | total us | share of all captured launches |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("| `%s` | %d | %.2f | %.1f | %.1f%% |" % (k[:110], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))


def full():
    rows = list(csv.reader(l for l in open("%s/full_%s.csv" % (src, tag)) if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen = set()
    for r in rows[2:]:
        if r[ki] in seen:
            continue
        seen.add(r[ki])
        print("\n## `%s`  grid %s  block %s\n" % (r[ki][:120], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        print("| metric | unit | value |\n|---|---|---|")
        for m in KEEP:
            if m in hdr:
                j = hdr.index(m)
                print("| %s | %s | %s |" % (m, units[j], r[j]))


if __name__ == "__main__":
    print("# ncu launch list (%s): per-launch times are cold-cache and serialised -- compare SHARES, not absolutes\n" % tag)
    launch_list()
    print("\n# ncu --set full, first captured launch of each of our kernels (%s)" % tag)
    full()
