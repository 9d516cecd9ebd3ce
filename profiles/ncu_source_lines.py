"""Aggregates an ncu source page (``--page source --csv --print-source sass,cuda``) per CUDA source line:
executed warp instructions and stall samples.  usage: ncu_source_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main(path, kernel, top=40):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name',
                          f'regex:{kernel}'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr = None, None
    agg = defaultdict(lambda: [0, 0, ''])
    total_inst = total_samp = 0
    first_kernel = None
    seen_files = set()
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            if fname in seen_files:
                break                 # second launch of the kernel: stop
            seen_files.add(fname)
            continue
        if r[0] == 'Function Name':
            first_kernel = first_kernel or r[1]
            continue
        if r[0] == 'Line No':
            hdr = r
            i_inst = hdr.index('Instructions Executed')
            i_samp = hdr.index('# Samples')
            continue
        if fname is None or hdr is None or len(r) <= i_inst:
            continue
        try:
            inst = int(r[i_inst] or 0)
            samp = int(r[i_samp] or 0)
        except ValueError:
            continue
        if not r[0].isdigit():
            continue                  # SASS rows (already aggregated in their source-line row)
        key = (fname, int(r[0]))
        a = agg[key]
        a[0] += inst
        a[1] += samp
        if r[1]:
            a[2] = r[1].strip()
        total_inst += inst
        total_samp += samp
    print(f'# {first_kernel}\n# total warp instructions {total_inst}, stall samples {total_samp}')
    print(f'{"file:line":34s} {"inst":>10s} {"inst%":>6s} {"samples%":>8s}  source')
    by = 1 if '--by-samples' in sys.argv else 0
    for (f, ln), (inst, samp, src) in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
        print(f'{f + ":" + str(ln):34s} {inst:10d} {100.0 * inst / max(total_inst, 1):6.2f} {100.0 * samp / max(total_samp, 1):8.2f}  {src[:110]}')


if __name__ == '__main__':
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    main(args[0], args[1], int(args[2]) if len(args) > 2 else 40)
