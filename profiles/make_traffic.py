"""profiles/traffic.json (read by bench.py: roofline.traffic) from the --set full capture of the
string_pack_kernel launches of one bench step.
usage: python profiles/make_traffic.py gpurun_out/<tag>_bench_string_full.ncu-rep profiles/<tag>_bench_string_kernel_ncu_full.txt"""
import csv, json, os, subprocess, sys

rep, txt = sys.argv[1:3]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size']
mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
tot = []
with open(txt, 'w') as f:
    for r in rows[2:]:
        f.write('--- ' + r[hdr.index('Kernel Name')] + '\n')
        for k in keys:
            if k in hdr:
                f.write(f"  {k:72s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}\n")
        b = sum(float(r[hdr.index(k)]) * mult[units[hdr.index(k)]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        tot.append(b)
json.dump({"rows": 60000000, "kernel": "string_pack_kernel", "dram_bytes_per_launch": int(sum(tot) / len(tot)),
           "source": f"ncu --set full --clock-control none -k regex:string_pack -s 2 -c 2 python bench.py --steps 2 --warmup 1 --no-cpu "
                     f"({os.path.basename(txt)}): dram__bytes_read.sum + dram__bytes_write.sum, mean over the string_pack_kernel launches of one step (l_shipinstruct, l_comment)",
           "launches": len(tot)}, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'traffic.json'), 'w'), indent=1)
print(open(txt).read())
