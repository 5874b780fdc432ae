"""Per-source-line view of an .ncu-rep captured with --import-source on (kernels built with -lineinfo):
executed warp instructions and stall samples per CUDA source line, largest first.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [warp_frames] [top]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
path, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        path = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = {h: i for i, h in enumerate(r)}
        first_source = r.index('Source')
    elif r[0] not in ('', 'Function Name') and hdr:
        try:
            lines.append((path, int(r[0]), r[first_source].strip(), int(r[hdr['# Samples']] or 0), int(r[hdr['Instructions Executed']] or 0)))
        except ValueError:
            pass
tot_s = sum(l[3] for l in lines) or 1
tot_i = sum(l[4] for l in lines) or 1
print('total samples %d, warp instructions %d%s' % (tot_s, tot_i, (' (%.0f per unit)' % (tot_i / units)) if units else ''))
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    per = (' %7.1f/unit' % (l[4] / units)) if units else ''
    print('%5.1f%% inst %5.1f%% samples%s  %s:%d  %s' % (100.0 * l[4] / tot_i, 100.0 * l[3] / tot_s, per, l[0], l[1], l[2][:110]))
