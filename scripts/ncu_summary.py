"""Compact text summary of an `ncu --set full` report: one block per profiled launch with the metrics DESIGN.md cites.
    python scripts/ncu_summary.py REPORT.ncu-rep > profiles/NAME.txt"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max']


def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f'# {rep}: {len(data)} profiled launches (ncu --set full --clock-control none; cold caches, serialised)')
    for r in data:
        print(f"\n== [{r[idx['ID']]}] {r[idx['Kernel Name']][:110]}")
        for w in WANT:
            if w in idx:
                print(f'   {w:72s} {r[idx[w]]:>18s} {units[idx[w]]}')


if __name__ == '__main__':
    main(sys.argv[1])
