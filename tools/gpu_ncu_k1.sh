mkdir -p gpurun_out
for v in ${K1_VARIANTS:-0}; do
TC_SAMPLE_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:sample_kernel -s 6 -c 1 -f -o gpurun_out/prof_k1_v$v python tools/k1_bench.py > gpurun_out/ncu_k1_v$v.log 2>&1; echo "rc=$?"
done
ls -la gpurun_out/*.ncu-rep
