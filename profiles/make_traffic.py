"""profiles/traffic.json (read by bench.py: roofline.traffic) from --set full captures of the kernel launches of one
bench step: mean dram__bytes_read.sum + dram__bytes_write.sum per launch, per kernel family.
usage: python profiles/make_traffic.py <out.txt> <rep> [<rep> ...]      (reports from profiles/run_bench_ncu.sh)"""
import csv, json, os, re, subprocess, sys

txt, reps = sys.argv[1], sys.argv[2:]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size']
mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
fam = {}
with open(txt, 'w') as f:
    for rep in reps:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')]
            f.write('--- ' + name + '\n')
            for k in keys:
                if k in hdr:
                    f.write(f"  {k:72s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}\n")
            b = sum(float(r[hdr.index(k)]) * mult[units[hdr.index(k)]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
            key = name.split('<')[0].replace('void ', '').strip()
            if key == 'string_pack_kernel' and re.search(r'string_pack_kernel<\d+, \d+, \d+, 0', name):
                key = 'string_pack_kernel<no heap>'  # the lean form for columns without a heap: a family of its own in bench.py
            fam.setdefault(key, []).append(b)
out = {"rows": 60000000,
       "source": "ncu --set full --clock-control none -k regex:<family> python bench.py --steps 2 --warmup 1 --no-cpu (" + os.path.basename(txt) +
                 "): dram__bytes_read.sum + dram__bytes_write.sum, mean over the family's launches of one step",
       "kernels": {k: {"dram_bytes_per_launch": int(sum(v) / len(v)), "launches": len(v)} for k, v in fam.items()}}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'traffic.json'), 'w'), indent=1)
print(open(txt).read())
print(json.dumps(out, indent=1))
