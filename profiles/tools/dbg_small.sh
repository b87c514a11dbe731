# per-tile clock stamps of the tcgen05 kernels on small / medium shapes (profiling build; see dbg_gemm.py)
export FP8B_LIB=profiles/tools/bin/libfp8_b200_profile.so
for spec in "256,3072,3072 2" "256,3072,3072 1" "32,3072,3072 4" "32,3072,3072 2" "32,3072,3072 1" "100,4096,4096 4"; do
  set -- $spec
  echo "=== SHAPE $1 splitk $2"
  SHAPE=$1 FP8B_GEMM_SPLITK=$2 python profiles/tools/dbg_gemm.py 2>&1 | head -6
done
