# round-2 ncu captures (one launch each, after warm-up) of the kernels VERDICT r1 names: string_pack_kernel on the l_comment and
# l_shipinstruct shapes, string_short_kernel on the l_shipmode shape.  usage: bash profiles/run_prof_r02.sh <tag>
tag=${1:-r02}
for spec in "string:string_pack" "string_mixed:string_pack" "string_mode:string_short" "string_short:string_short"; do
  w=${spec%%:*}; k=${spec##*:}
  python profiles/prof_kernels.py --which $w > gpurun_out/${tag}_plain_$w.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/${tag}_${w} -f python profiles/prof_kernels.py --which $w > gpurun_out/${tag}_ncu_$w.log 2>&1
done
ls -la gpurun_out/${tag}_*.ncu-rep
