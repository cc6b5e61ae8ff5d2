#!/usr/bin/env python
"""Condense an .ncu-rep (read here, no GPU needed) into (a) a key-metric csv and (b) a basic-block table of the SASS
source page: instructions executed, share of issue slots, share of stall samples, average active threads."""
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'lts__t_sectors.sum', 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_static']


def page(rep, name):
    return list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout.splitlines()))


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    rows = page(rep, 'raw')
    hdr, units = rows[0], rows[1]
    print(f"# {title}")
    for vals in rows[2:]:
        print("metric,unit,value")
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                print(f"{h},{u},{v}")
    rows = page(rep, 'source')
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    tot_s = sum(int(r[ix['# Samples']]) for r in data) or 1
    tot_i = sum(int(r[ix['Instructions Executed']]) for r in data) or 1
    print(f"# SASS basic blocks (>=0.4% of issue slots or samples); total warp-instructions {tot_i}, samples {tot_s}")
    print("first_instr,last_instr,n_instr,exec_per_instr_M,issue_pct,sample_pct,avg_threads,first_sass")
    cur = None
    segs = []
    for k, r in enumerate(data):
        s = int(r[ix['# Samples']]); i = int(r[ix['Instructions Executed']]); t = float(r[ix['Avg. Threads Executed']] or 0)
        if cur and abs(i - cur['i']) <= 0.02 * max(cur['i'], 1):
            cur['n'] += 1; cur['s'] += s; cur['ti'] += i; cur['tt'] += t * i; cur['end'] = k
        else:
            if cur:
                segs.append(cur)
            cur = dict(start=k, end=k, i=i, n=1, s=s, ti=i, tt=t * i, first=r[ix['Source']].strip()[:48])
    segs.append(cur)
    for g in segs:
        if g['ti'] / tot_i > 0.004 or g['s'] / tot_s > 0.004:
            print(f"{g['start']},{g['end']},{g['n']},{g['i'] / 1e6:.1f},{100 * g['ti'] / tot_i:.2f},{100 * g['s'] / tot_s:.2f},{g['tt'] / max(g['ti'], 1):.1f},\"{g['first']}\"")


if __name__ == '__main__':
    main()
