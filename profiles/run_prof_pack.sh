set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in string string_mixed string_short; do
  ncu --set full --clock-control none --import-source on -k regex:string_pack -s 2 -c 1 -o gpurun_out/prof_pack_${w}_r01j -f python profiles/prof_kernels.py --which $w > gpurun_out/ncu_pack_$w.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -4
