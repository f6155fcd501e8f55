"""Compact text summary of an `ncu --set full` report: one block per profiled launch with the metrics DESIGN.md cites.
    python scripts/ncu_summary.py REPORT.ncu-rep [TRAFFIC.json] > profiles/NAME.txt
With a second argument, also writes {kernel name: {dram_bytes, ms, launches}} (the LAST profiled launch of each kernel name:
targets launch twice, the second is warm) -- bench.py reports it as `roofline.traffic`."""
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


def main(rep, traffic_json=None):
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
    if traffic_json:
        import json
        import re

        def num(r, key):
            v = float(r[idx[key]].replace(',', ''))
            u = units[idx[key]].lower()
            return v * {'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'usecond': 1e-3, 'us': 1e-3, 'nsecond': 1e-6, 'ns': 1e-6, 'msecond': 1.0, 'ms': 1.0,
                        'second': 1e3, 's': 1e3}.get(u, 1.0)
        out, order = {}, []
        for r in data:
            name = re.sub(r'\(.*', '', r[idx['Kernel Name']]).replace('void ', '')
            key = f"{name} grid {r[idx['launch__grid_size']]}"
            if key not in out:
                order.append(key)
            prev = out.get(key, {}).get('launches', 0)
            out[key] = dict(dram_bytes=int(num(r, 'dram__bytes_read.sum') + num(r, 'dram__bytes_write.sum')),
                            ms=round(num(r, 'gpu__time_duration.sum'), 4), launches=prev + 1)
        json.dump(dict(source=rep.split('/')[-1], how='ncu --set full --clock-control none, last launch of each (kernel, grid)',
                       kernels={k: out[k] for k in order}), open(traffic_json, 'w'), indent=1)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
