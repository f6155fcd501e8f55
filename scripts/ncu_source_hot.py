"""Per-source-line hot spots of one kernel from an .ncu-rep (source page, CUDA view):
    python scripts/ncu_source_hot.py REPORT KERNEL_REGEX [launch-index]
Prints, per source line: stall samples, instructions executed, shared wavefronts (ideal / excessive)."""
import csv
import io
import subprocess
import sys


def main(rep, kre, which=None, top=28):
    cmd = ['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kre}', '--print-source', 'cuda,sass']
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    # the output is a sequence of blocks: "File Path", "Function Name", header, rows
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == 'File Path':
            cur = dict(file=row[1], rows=[], hdr=None, fn=None)
            blocks.append(cur)
        elif row[0] == 'Function Name' and cur is not None:
            cur['fn'] = row[1]
        elif row[0] == 'Line No' and cur is not None:
            cur['hdr'] = row
        elif cur is not None and cur['hdr'] is not None:
            cur['rows'].append(row)
    fns = []
    for b in blocks:
        if b['fn'] not in fns:
            fns.append(b['fn'])
    lines = []
    for b in blocks:
        h = {n: i for i, n in enumerate(b['hdr'])}
        for r in b['rows']:
            if not r[0].strip().isdigit():          # SASS rows carry an empty line number
                continue
            try:
                lines.append(dict(file=b['file'].split('/')[-1], fn=b['fn'], line=r[h['Line No']], src=r[1].strip()[:90],
                                  samples=int(r[h['# Samples']] or 0), inst=int(r[h['Instructions Executed']] or 0),
                                  wave=int(r[h['L1 Wavefronts Shared']] or 0), ideal=int(r[h['L1 Wavefronts Shared Ideal']] or 0)))
            except (ValueError, KeyError):
                pass
    tot = sum(l['samples'] for l in lines) or 1
    print(f'# {rep} {kre}: {len(blocks)} source blocks, {tot} stall samples')
    for l in sorted(lines, key=lambda l: -l['samples'])[:top]:
        print(f"{100 * l['samples'] / tot:5.1f}% smp  inst {l['inst']:>10d}  smem wave {l['wave']:>9d} (ideal {l['ideal']:>9d})  {l['file']}:{l['line']}  {l['src']}")


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
