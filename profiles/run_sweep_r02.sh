# round-2 A/B of the string_pack_kernel knobs: parity first, then CUDA-event times per variant, then (NCU=1) instruction and
# shared-memory wavefront counts of one launch per shape (16 M rows)
# usage: [SHAPES=comment,mixed] [NCU=1] bash profiles/run_sweep_r02.sh <tag> <variant> [<variant> ...]   (variants: duckdb.mbt_b200/csrc/variants/lib_<name>.so)
tag=$1; shift
V=duckdb.mbt_b200/csrc/variants
for v in "$@"; do
  DMB_LIB_PATH=$PWD/$V/lib_$v.so timeout 900 python -m pytest tests/test_gpu_l0_parity.py -m gpu -x -q > gpurun_out/${tag}_parity_$v.log 2>&1
  echo "parity $v rc=$? $(tail -1 gpurun_out/${tag}_parity_$v.log)" | tee -a gpurun_out/${tag}_sweep.jsonl
  DMB_LIB_PATH=$PWD/$V/lib_$v.so timeout 600 python profiles/sweep_string.py --shapes ${SHAPES:-comment,mixed,c3,run} 2>gpurun_out/${tag}_sweep_$v.err | tee -a gpurun_out/${tag}_sweep.jsonl
  if [ -n "$NCU" ]; then
    for sh in ${NCU_SHAPES:-comment mixed}; do
      DMB_LIB_PATH=$PWD/$V/lib_$v.so timeout 600 ncu --metrics smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum \
        --clock-control none -k regex:string_pack -s 2 -c 1 --csv python profiles/sweep_string.py --shapes $sh --rows 16000000 --iters 1 2>/dev/null \
        | grep string_pack | awk -F'","' -v v=$v -v sh=$sh '{gsub(/"/,"",$NF); print "ncu", v, sh, $(NF-2), $NF}' | tee -a gpurun_out/${tag}_sweep.jsonl
    done
  fi
done
