export SHAPES="512,3072,12288;128,3072,12288;32,3072,12288;256,3072,3072;2048,512,2048;1000,3072,3072"
echo "--- cold"; python profiles/tools/shape_sweep.py --cfgs 2>&1 | tail -7
echo "--- hot (SETS=1)"; SETS=1 python profiles/tools/shape_sweep.py --cfgs 2>&1 | tail -7
python profiles/run_kernels.py mm:512,3072,12288 4 > gpurun_out/plain_mm512.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fp8_gemm_tcgen05 -s 2 -c 1 -f -o gpurun_out/r2_mm512 python profiles/run_kernels.py mm:512,3072,12288 4 > gpurun_out/ncu_mm512.log 2>&1; echo "ncu rc=$?"
