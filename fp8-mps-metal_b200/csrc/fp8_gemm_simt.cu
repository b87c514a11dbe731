// CUDA-core tiled FP8 GEMM for ANY shape and alignment.
//
// This is the device path for problems the TMA/tcgen05 kernel cannot take (K % 16 != 0,
// unaligned base pointers) -- the library has no CPU fallback, so something on the GPU must serve
// them -- and an independent on-device cross-check of the tcgen05 kernel in the tests.
// Same arithmetic as fp8_scaled_matmul_kernel (fp8_matmul.metal:99-147): masked decode (NaN -> 0; e5m2 operands IEEE),
// fp32 FMA accumulation, fused epilogue.  64x64 output tile per CTA, 4x4 outputs per thread,
// 32-wide K panels staged in shared memory as fp32.
#include "fp8_mm.cuh"

namespace fp8b {

constexpr int kSimtTile = 64;
constexpr int kSimtK = 32;
constexpr int kSimtThreads = 256;

__global__ void __launch_bounds__(kSimtThreads)
fp8_gemm_simt_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B, int M, int N, int K, const Epi epi,
                     int a_fmt, int b_fmt)
{
    __shared__ float As[kSimtK][kSimtTile + 4];
    __shared__ float Bs[kSimtK][kSimtTile + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m_base = blockIdx.y * kSimtTile, n_base = blockIdx.x * kSimtTile;
    const int lrow = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 8;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < K; k0 += kSimtK) {
        const int am = m_base + lrow, bn = n_base + lrow;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + lk + j;
            As[lk + j][lrow] = (am < M && k < K) ? dec1_fmt_f32(A[(size_t)am * K + k], a_fmt) : 0.0f;
            Bs[lk + j][lrow] = (bn < N && k < K) ? dec1_fmt_f32(B[(size_t)bn * K + k], b_fmt) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kSimtK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m_base + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n_base + tx * 4 + j;
            if (n < N) epi_store(epi, m, n, epi_apply(epi, acc[i][j], m, n));
        }
    }
}

int launch_gemm_simt(const MMArgs& a)
{
    const Epi epi = make_epi(a);
    dim3 grid((a.N + kSimtTile - 1) / kSimtTile, (a.M + kSimtTile - 1) / kSimtTile, 1);
    if (grid.y > 65535) return FP8B_ERR_UNSUPPORTED;
    fp8_gemm_simt_kernel<<<grid, kSimtThreads, 0, a.st>>>(a.A, a.B, a.M, a.N, a.K, epi, a.a_fmt, a.b_fmt);
    return after_launch();
}

}  // namespace fp8b
