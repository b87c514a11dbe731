// PROFILING-ONLY entry points (compiled into profiles/tools/bin/libfp8_b200_profile.so by FP8B_BUILD_PROFILE=1, never into
// the shipped library): raw store-bandwidth probes used to find out what a peer GPU's memory accepts over NVLink from
// SM-issued traffic, so that the fused GEMM + exchange kernel can be judged against a measured floor.
#ifdef FP8B_PROFILE
#include "fp8_common.cuh"
#include "fp8_async.cuh"

namespace fp8b {

// mode 0: st.global.v4 from registers, grid-stride, fully coalesced 512 B per warp instruction
__global__ void prof_fill_stg_kernel(uint4* __restrict__ dst, size_t nvec)
{
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 0x3C3C3C3Cu, 0x3C3C3C3Cu);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// mode 1: cp.async.bulk shared -> global, `chunk` bytes per copy, `depth` copies in flight per CTA.
// row_bytes > 0: the destination is a strided matrix (row pitch `pitch` bytes) and every copy writes ONE row segment of
// row_bytes (the access pattern of a GEMM tile); row_bytes == 0: contiguous chunks.
__global__ void prof_fill_bulk_kernel(uint8_t* __restrict__ dst, size_t bytes, int chunk, int depth, int row_bytes, size_t pitch)
{
    extern __shared__ __align__(128) uint8_t prof_smem[];
    for (int i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(prof_smem)[i] = 0x3C3C3C3Cu;
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t nchunks = bytes / chunk;
    int inflight = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        uint8_t* p;
        if (row_bytes) {
            const size_t per_row = pitch / row_bytes;                    // segments per matrix row
            p = dst + (c / per_row) * pitch + (c % per_row) * (size_t)row_bytes;
        } else {
            p = dst + c * (size_t)chunk;
        }
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(p), "r"(smem_u32(prof_smem)), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (++inflight >= depth) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); inflight = 0; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace fp8b

using namespace fp8b;

extern "C" __attribute__((visibility("default")))
int fp8b_prof_fill(void* dst, size_t bytes, int mode, int chunk, int depth, int row_bytes, size_t pitch, int ctas_per_sm, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = device_info().sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    if (mode == 0) {
        prof_fill_stg_kernel<<<grid, 512, 0, st>>>(static_cast<uint4*>(dst), bytes / 16);
    } else {
        static bool once = false;
        if (!once) { cudaFuncSetAttribute(prof_fill_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); once = true; }
        prof_fill_bulk_kernel<<<grid, 128, chunk, st>>>(static_cast<uint8_t*>(dst), bytes, chunk, depth, row_bytes, pitch);
    }
    return after_launch();
}
#endif
