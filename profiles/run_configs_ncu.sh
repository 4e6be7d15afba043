# profiles/bench_configs.py (C3 forward VARCHAR-heavy, C5 reverse) under ncu, after a plain run has exited 0:
# full captures of one launch of string_pack_kernel (C3), rev_fixed_kernel and rev_string_kernel (C5).
# usage: bash profiles/run_configs_ncu.sh <tag>
tag=${1:-r01w}
set -x
python profiles/bench_configs.py --configs c3,c5 --scale 0.4 > gpurun_out/${tag}_configs_plain.jsonl 2> gpurun_out/${tag}_configs_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:'string_pack' -s 3 -c 1 -o gpurun_out/${tag}_c3_string_pack_full -f \
    python profiles/bench_configs.py --configs c3 --scale 0.4 > gpurun_out/ncu_${tag}_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'rev_' -s 6 -c 2 -o gpurun_out/${tag}_c5_rev_full -f \
    python profiles/bench_configs.py --configs c5 --scale 0.4 > gpurun_out/ncu_${tag}_c5.log 2>&1
ls -la gpurun_out/${tag}_*
