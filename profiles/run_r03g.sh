set -x
timeout 300 python -m pytest tests/test_gpu_enum.py -x -q > gpurun_out/r03g_tests.log 2>&1; echo "tests rc=$?"
V=duckdb.mbt_b200/csrc/variants
for rep in 1 2; do
  DMB_LIB_PATH=$V/lib_base.so timeout 200 python profiles/bench_configs.py --configs enum > gpurun_out/r03g_base_$rep.jsonl 2>gpurun_out/r03g_base.err
  timeout 200 python profiles/bench_configs.py --configs enum > gpurun_out/r03g_new_$rep.jsonl 2>gpurun_out/r03g_new.err
  DMB_LIB_PATH=$V/lib_enumlbf.so timeout 200 python profiles/bench_configs.py --configs enum > gpurun_out/r03g_enumlbf_$rep.jsonl 2>gpurun_out/r03g_enumlbf.err
done
