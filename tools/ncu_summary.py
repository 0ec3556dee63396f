#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + per-region SASS execution counts) into text.
usage: python tools/ncu_summary.py report.ncu-rep [bucket]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'l1tex__t_sectors.sum', 'lts__t_sectors.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('--- kernel', r[hdr.index('Kernel Name')][:70])
        for k in KEYS:
            if k in hdr:
                print('   %-82s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))

        def num(k):
            try:
                return float(r[hdr.index(k)].replace(',', '')), units[hdr.index(k)]
            except (ValueError, IndexError):
                return None, None
        dur, du = num('gpu__time_duration.sum')
        if dur:
            sec = dur * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}.get(du, 1e-9)
            scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            for k, label in (('l1tex__t_bytes.sum', 'achieved L1TEX GB/s'), ('lts__t_bytes.sum', 'achieved L2 GB/s')):
                v, u = num(k)
                if v is not None:
                    print('   %-82s %.1f' % (label + ' (= %s / duration)' % k, v * scale.get(u, 1.0) / sec / 1e9))
            rd, ru = num('dram__bytes_read.sum')
            wr, wu = num('dram__bytes_write.sum')
            if rd is not None and wr is not None:
                print('   %-82s %.1f' % ('achieved HBM GB/s (= dram read + write / duration)', (rd * scale.get(ru, 1.0) + wr * scale.get(wu, 1.0)) / sec / 1e9))
    if not bucket:
        return
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    kern, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and len(r) > 10:
            cur["rows"].append(r)
    for k in kern:
        h = k["hdr"]
        iI, iT, iS, iSrc = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples"), h.index("Source")
        R = k["rows"]
        tot = sum(int(r[iI]) for r in R) or 1
        tott = sum(int(r[iT]) for r in R)
        tots = sum(int(r[iS]) for r in R) or 1
        print("=== %s  sass %d  warp-inst %d  avg threads %.2f" % (k["name"][:60], len(R), tot, tott / tot))
        for s in range(0, len(R), bucket):
            seg = R[s:s + bucket]
            wi = sum(int(r[iI]) for r in seg)
            if wi / tot < 0.005:
                continue
            ti = sum(int(r[iT]) for r in seg)
            sm = sum(int(r[iS]) for r in seg)
            ops = {}
            for r in seg:
                t = r[iSrc].split()
                op = t[1] if t[0].startswith('@') else t[0]
                ops[op] = ops.get(op, 0) + 1
            top = sorted(ops.items(), key=lambda x: -x[1])[:5]
            print(f"{s:5d} warp-inst {wi / tot * 100:5.1f}%  avg thr {ti / max(wi, 1):5.1f}  samples {sm / tots * 100:5.1f}%  {top}")


if __name__ == "__main__":
    main()
