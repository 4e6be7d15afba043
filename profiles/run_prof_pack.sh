# ncu captures of string_pack_kernel on the four bench-shaped columns (one launch each, after warm-up)
tag=${1:-r01k}
for w in string string_mixed string_short string_c3; do
  ncu --set full --clock-control none --import-source on -k regex:string_pack -s 2 -c 1 -o gpurun_out/prof_pack_${w}_${tag} -f python profiles/prof_kernels.py --which $w > gpurun_out/ncu_pack_$w.log 2>&1
done
ls -la gpurun_out/*_${tag}.ncu-rep
