# round-2 (third session) captures: LIST and ENUM configs of profiles/bench_configs.py, plain run first, then --set full of one launch each
# usage: bash profiles/run_prof_r03.sh <tag> [list-kernel-regex] [enum-kernel-regex]
tag=${1:-r03b}
lk=${2:-list_emit}
ek=${3:-enum_pack}
set -x
python profiles/bench_configs.py --configs list,enum > gpurun_out/${tag}_configs_plain.jsonl 2> gpurun_out/${tag}_configs_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:$lk -s 3 -c 1 -o gpurun_out/${tag}_list_full -f \
    python profiles/bench_configs.py --configs list > gpurun_out/ncu_${tag}_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$ek -s 3 -c 1 -o gpurun_out/${tag}_enum_full -f \
    python profiles/bench_configs.py --configs enum > gpurun_out/ncu_${tag}_enum.log 2>&1
ls -la gpurun_out/${tag}_*
