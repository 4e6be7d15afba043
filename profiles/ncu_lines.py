"""Map an ncu SASS source page onto CUDA source lines (needs -lineinfo): per-line warp-instruction
counts and stall samples for one profiled kernel.
usage: python profiles/ncu_lines.py <rep> <cubin> <kernel-substring> [min_pct]"""
import csv
import re
import subprocess
import sys


def main():
    rep, cubin, kern = sys.argv[1:4]
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # first kernel block
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or not r[0].startswith('0x'):
            break
        body.append(r)
    base = int(body[0][0], 16)
    ie, ss = hdr.index('Instructions Executed'), hdr.index('# Samples')
    per_off = {int(r[0], 16) - base: (int(float(r[ie] or 0)), int(float(r[ss] or 0)), r[1].strip()) for r in body}
    dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    line_of = {}
    cur_line, in_fn = None, False
    for ln in dis.splitlines():
        m = re.match(r'\s*\.text\.(\S+):', ln)
        if m:
            in_fn = kern in m.group(1)
            continue
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
        if m:
            line_of[int(m.group(1), 16)] = cur_line
    agg = {}
    tot_i = sum(v[0] for v in per_off.values())
    tot_s = sum(v[1] for v in per_off.values())
    for off, (n, smp, _) in per_off.items():
        k = line_of.get(off)
        a = agg.setdefault(k, [0, 0])
        a[0] += n
        a[1] += smp
    print(f'total warp-instructions {tot_i}, samples {tot_s}')
    for k, (n, smp) in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
        if n >= tot_i * min_pct / 100 or smp >= tot_s * min_pct / 100:
            print(f'{100 * n / tot_i:6.2f}% inst  {100 * smp / max(tot_s, 1):6.2f}% samples  {k}')


if __name__ == '__main__':
    main()
