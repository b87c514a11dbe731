# per-tile clock stamps of the tcgen05 kernel on small / medium shapes (profiling build; see dbg_gemm.py)
export FP8B_LIB=profiles/tools/bin/libfp8_b200_profile.so
for spec in "512,3072,12288 5" "128,3072,12288 2" "2048,512,2048 3" "256,3072,3072 2" "4096,3072,12288 3"; do
  set -- $spec
  echo "=== SHAPE $1 cfg $2"
  SHAPE=$1 FP8B_GEMM_CFG=$2 python profiles/tools/dbg_gemm.py 2>&1 | head -6
done
