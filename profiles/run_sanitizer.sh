#!/bin/bash
# compute-sanitizer over the GPU parity suite (round-1 verdict item 8).  Run under gpurun:
#   gpurun --timeout 2400 -- 'bash profiles/run_sanitizer.sh r02a'
# memcheck over the whole -m gpu suite except the bench-scale file (tests/test_gpu_scale.py: 100 M-row inputs under a
# 20-50x slowdown would take hours); racecheck (shared-memory hazards: the kernels' stages, mbarrier pipelines, named
# barriers) over the L0 kernel tests + the host API file.  Summaries land in gpurun_out/ and are copied to profiles/.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
export DMB_SANITIZER=1
SAN=/usr/local/cuda/bin/compute-sanitizer
cd "$(dirname "$0")/.."
timeout 1200 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 --log-file $out/${tag}_memcheck.log \
  python -m pytest tests -x -q -m gpu --ignore=tests/test_gpu_scale.py -p no:cacheprovider > $out/${tag}_memcheck_pytest.log 2>&1
echo "memcheck exit: $?" | tee -a $out/${tag}_memcheck_pytest.log
tail -5 $out/${tag}_memcheck.log
timeout 900 $SAN --tool racecheck --racecheck-report all --error-exitcode 9 --print-limit 20 --log-file $out/${tag}_racecheck.log \
  python -m pytest tests/test_gpu_l0_parity.py tests/test_gpu_l0_reverse.py tests/test_gpu_l0_list.py tests/test_gpu_enum.py tests/test_gpu_render.py \
  -x -q -m gpu -p no:cacheprovider > $out/${tag}_racecheck_pytest.log 2>&1
echo "racecheck exit: $?" | tee -a $out/${tag}_racecheck_pytest.log
tail -5 $out/${tag}_racecheck.log
