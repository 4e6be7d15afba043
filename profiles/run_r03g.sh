# same-box A/B of the ENUM kernels: lib_base (string_short_kernel ENUM form), lib_enum1 (enum_pack_kernel, one CTA per tile), current build
set -x
timeout 300 python -m pytest tests/test_gpu_enum.py tests/test_gpu_nested.py -x -q > gpurun_out/$1_tests.log 2>&1; echo "tests rc=$?"
V=duckdb.mbt_b200/csrc/variants
for rep in 1 2; do
  DMB_LIB_PATH=$V/lib_base.so timeout 200 python profiles/bench_configs.py --configs enum > gpurun_out/$1_base_$rep.jsonl 2>gpurun_out/$1_base.err
  DMB_LIB_PATH=$V/lib_enum1.so timeout 200 python profiles/bench_configs.py --configs enum > gpurun_out/$1_enum1_$rep.jsonl 2>gpurun_out/$1_enum1.err
  timeout 200 python profiles/bench_configs.py --configs enum > gpurun_out/$1_new_$rep.jsonl 2>gpurun_out/$1_new.err
done
