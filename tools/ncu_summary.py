"""Summarise an `ncu --set full` report for profiles/: selected raw metrics (one column per captured launch), the SASS opcode
histogram weighted by executed warp instructions, and the source lines with the most stall samples.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_ncu_full.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|"
    r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"l1tex__m_xbar2l1tex_read_bytes\.sum(\.per_second)?|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__t_sector_hit_rate\.pct|"
    r"launch__(grid_size|block_size|cluster_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_\w+|waves_per_multiprocessor)|"
    r"sm__cycles_elapsed\.max(\.per_second)?|sm__cycles_active\.avg|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__inst_executed_pipe_(alu|fma|fmaheavy|xu|lsu|tensor\w*|uniform)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_(tensor|tc|fma|fmaheavy|alu|shared)_cycles_active\w*\.avg\.pct_of_peak_sustained_(active|elapsed)|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__sass_thread_inst_executed_op_(ffma|fadd|fmul)_pred_on\.sum|sm__sass_inst_executed_op_shared_(ld|st)\.sum)$")


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True, check=True).stdout


def main(rep, top=25):
    raw = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, launches = raw[0], raw[1], raw[2:]
    print("# %s : selected metrics of `ncu --set full --clock-control none` (one column per captured launch)" % rep.split("/")[-1])
    ik = hdr.index("Kernel Name")
    print("Kernel Name [] " + " | ".join(r[ik][:120] for r in launches))
    for i, h in sorted(enumerate(hdr), key=lambda t: t[1]):
        if KEEP.match(h):
            print("%s [%s] %s" % (h, units[i], " | ".join(r[i] for r in launches)))
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    # the source page repeats a 2-line header per kernel; summarise the first kernel only
    h = src[1]
    ia, ie, iss = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    data = []
    for r in src[2:]:
        if len(r) <= max(ia, ie, iss) or not r[ie].isdigit():
            break
        data.append(r)
    tot = sum(int(r[ie]) for r in data) or 1
    tots = sum(int(r[iss]) for r in data) or 1
    ops, st = Counter(), Counter()
    for r in data:
        t = r[ia].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += int(r[ie]); st[op] += int(r[iss])
    print("\n# SASS opcode histogram of the first captured launch (%d warp instructions, %d stall samples)" % (tot, tots))
    for k, v in ops.most_common(18):
        print("%-10s %11d %5.1f%% of instructions  %5.1f%% of stall samples" % (k, v, 100.0 * v / tot, 100.0 * st[k] / tots))
    print("\n# SASS lines with the most stall samples (index, samples, executed, instruction)")
    for i, r in sorted(enumerate(data), key=lambda t: -int(t[1][iss]))[:top]:
        print("%5d %6d %9s  %s" % (i, int(r[iss]), r[ie], r[ia].strip()[:100]))


if __name__ == "__main__":
    main(sys.argv[1])
