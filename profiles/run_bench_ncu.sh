# (round 2: --no-parity --no-configs --no-layout-leg keep the launch order of the value leg: warm-up step, then the timed steps)
# bench.py under ncu (after a plain run has exited 0): the launch list of one bench run and full captures of the
# launches of one step, per kernel family.  usage: bash profiles/run_bench_ncu.sh <tag>
tag=${1:-r01n}
set -x
python bench.py --steps 2 --warmup 1 --no-cpu --no-parity --no-configs --no-layout-leg > gpurun_out/bench_plain_${tag}.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'batch_kernel|pack_kernel|short_kernel|render|rev_' -c 400 --csv \
    --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-parity --no-configs --no-layout-leg > gpurun_out/ncu_launches_${tag}.log 2>&1
# value leg: warm-up step, then two timed steps; capture the launches of the second step (-s = launches of that family in one step)
# (five string_pack_kernel launches per step since the heap-less columns take the pipeline's lean form)
ncu --set full --clock-control none --import-source on -k regex:string_pack -s 5 -c 5 -o gpurun_out/${tag}_bench_string_pack_full -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-parity --no-configs --no-layout-leg > gpurun_out/ncu_full_pack_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:string_short -s 3 -c 3 -o gpurun_out/${tag}_bench_string_short_full -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-parity --no-configs --no-layout-leg > gpurun_out/ncu_full_short_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fixed_batch -s 3 -c 3 -o gpurun_out/${tag}_bench_fixed_full -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-parity --no-configs --no-layout-leg > gpurun_out/ncu_full_fixed_${tag}.log 2>&1
ls -la gpurun_out/${tag}_*
