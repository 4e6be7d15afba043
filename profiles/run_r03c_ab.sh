# same-box A/B of the list / ENUM kernels: variants built by profiles/build_one_variant.sh, selected with DMB_LIB_PATH
set -x
timeout 300 python -m pytest tests/test_gpu_l0_list.py tests/test_gpu_enum.py tests/test_gpu_nested.py -x -q > gpurun_out/${1}_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${1}_tests.log
V=duckdb.mbt_b200/csrc/variants
for rep in 1 2; do
  DMB_LIB_PATH=$V/lib_base.so timeout 200 python profiles/bench_configs.py --configs list,enum > gpurun_out/${1}_base_$rep.jsonl 2>gpurun_out/${1}_base.err
  timeout 200 python profiles/bench_configs.py --configs list,enum > gpurun_out/${1}_new_$rep.jsonl 2>gpurun_out/${1}_new.err
  for v in $2; do
    DMB_LIB_PATH=$V/lib_$v.so timeout 200 python profiles/bench_configs.py --configs list > gpurun_out/${1}_${v}_$rep.jsonl 2>gpurun_out/${1}_$v.err
  done
done
