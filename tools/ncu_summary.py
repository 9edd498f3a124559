#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics
DESIGN.md / bench.py quote.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_write.sum.per_second',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
    'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sectors_op_write.sum',
]
STALL = 'smsp__average_warps_issue_stalled_'


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        print('kernel:', rec.get('Kernel Name', '?')[:160])
        for k in KEYS:
            if k in rec:
                print(f'  {k:82s} {rec[k]:>16s} {units[hdr.index(k)]}')
        stalls = sorted(((float(rec[h]), h[len(STALL):-len('_per_issue_active.ratio')]) for h in hdr
                         if h.startswith(STALL) and h.endswith('_per_issue_active.ratio')), reverse=True)
        print('  warp stall reasons (warps stalled per issued instruction):')
        for v, name in stalls[:8]:
            print(f'    {name:28s} {v:8.3f}')


if __name__ == '__main__':
    main(sys.argv[1])
