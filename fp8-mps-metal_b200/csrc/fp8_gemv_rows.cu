// FP8 GEMV, M = 1, SM-balanced persistent variant:  out[n] = epi( sum_k dec(x[k]) * dec(W[n,k]) ).
//
// Replaces fp8_scaled_vecmat_kernel (fp8_matmul.metal:155-210: one 32-lane simdgroup per output
// row, x re-read from cache for every row).  HBM-bound: W is streamed exactly once.
//
// Work split chosen for the memory system rather than for the output:
//   * grid = one CTA per SM; CTA b owns the contiguous row range [b*N/G, (b+1)*N/G) -- for N = 4096
//     on 148 SMs that is 27 or 28 rows each, so every SM moves the same number of bytes (the
//     warp-per-row kernel put 3 or 4 eight-row CTAs on an SM: up to 15 % imbalance);
//   * inside a CTA the 16 warps split K, not the rows: warp w owns K-slice w for EVERY row of the
//     CTA, so each lane's 16-byte vectors of x are decoded once into registers (fp16 pairs) and
//     never touched again -- no shared memory, no barrier before the stream starts;
//   * a lane walks down the rows with kRowsUnroll independent ld.global.nc.L1::no_allocate.v4
//     per step (adjacent warps read adjacent slices of the same row, so each row is still one
//     contiguous burst) and keeps one fp32 accumulator per row in registers (FHFMA: exact fp16
//     product, fp32 accumulate);
//   * after a pass of 32 rows the accumulators are reduced with warp shuffles, then across the 16
//     warps through shared memory in warp order (deterministic), and the fused epilogue (scales,
//     bias, scale_result, out dtype) writes the outputs.
// NaN bytes: detected as a NaN sum and repaired with the masked scalar loop (metal:21).
#include "fp8_mm.cuh"

namespace fp8b {

constexpr int kRowsThreads = 512;
constexpr int kRowsWarps = kRowsThreads / 32;
constexpr int kRowsPerPass = 32;
constexpr int kRowsUnroll = 4;

struct GemvRowsParams {
    const uint8_t* A;
    const uint8_t* B;
    int N, K;
    int vpw;               // 16-byte vectors of K per warp
    Epi epi;
};

__device__ __forceinline__ uint4 ldg_rows_v4(const uint8_t* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void fhfma2_rows(float& acc, uint32_t w2, uint32_t x2) {
    asm("{\n\t.reg .b16 a0, a1, b0, b1;\n\t"
        "mov.b32 {a0, a1}, %1;\n\tmov.b32 {b0, b1}, %2;\n\t"
        "fma.rn.f32.f16 %0, a0, b0, %0;\n\tfma.rn.f32.f16 %0, a1, b1, %0;\n\t}"
        : "+f"(acc) : "r"(w2), "r"(x2));
}

__device__ __forceinline__ void rows_consume(const uint4& w, const uint32_t (&x)[8], float& acc) {
    uint32_t lo, hi;
    dec4_f16x2_raw(w.x, lo, hi); fhfma2_rows(acc, lo, x[0]); fhfma2_rows(acc, hi, x[1]);
    dec4_f16x2_raw(w.y, lo, hi); fhfma2_rows(acc, lo, x[2]); fhfma2_rows(acc, hi, x[3]);
    dec4_f16x2_raw(w.z, lo, hi); fhfma2_rows(acc, lo, x[4]); fhfma2_rows(acc, hi, x[5]);
    dec4_f16x2_raw(w.w, lo, hi); fhfma2_rows(acc, lo, x[6]); fhfma2_rows(acc, hi, x[7]);
}

template <int NV>                                   // 16-byte vectors of x held per lane
__global__ void __launch_bounds__(kRowsThreads, 1)
fp8_gemv_rows_kernel(const GemvRowsParams p)
{
    __shared__ float part[kRowsWarps][kRowsPerPass + 1];
    constexpr int U = (NV <= 2) ? kRowsUnroll : 2;      // row groups in flight; bounded by the 128-register budget
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = p.K;
    const int nvecK = K >> 4;
    const int v_begin = warp * p.vpw;
    const int v_end = min(nvecK, v_begin + p.vpw);

    // this lane's slice of x, decoded once (raw hardware decode: NaN bytes stay NaN on purpose)
    uint32_t xr[NV][8];
    bool vok[NV];
    uint32_t voff[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int v = v_begin + lane + 32 * j;
        vok[j] = v < v_end;
        voff[j] = (uint32_t)(vok[j] ? v : 0) * 16u;
        uint4 xb = make_uint4(0u, 0u, 0u, 0u);
        if (vok[j]) xb = *reinterpret_cast<const uint4*>(p.A + voff[j]);
        dec4_f16x2_raw(xb.x, xr[j][0], xr[j][1]);
        dec4_f16x2_raw(xb.y, xr[j][2], xr[j][3]);
        dec4_f16x2_raw(xb.z, xr[j][4], xr[j][5]);
        dec4_f16x2_raw(xb.w, xr[j][6], xr[j][7]);
    }

    const int row_begin = (int)(((long long)blockIdx.x * p.N) / gridDim.x);
    const int row_end = (int)(((long long)(blockIdx.x + 1) * p.N) / gridDim.x);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

    for (int pass = row_begin; pass < row_end; pass += kRowsPerPass) {
        const int R = min(kRowsPerPass, row_end - pass);
        const uint8_t* wbase = p.B + (size_t)pass * K;
        float acc[kRowsPerPass];
#pragma unroll
        for (int i = 0; i < kRowsPerPass; ++i) acc[i] = 0.0f;

        uint4 cur[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < NV; ++j)
                cur[u][j] = (u < R && vok[j]) ? ldg_rows_v4(wbase + (size_t)u * K + voff[j]) : zero4;

#pragma unroll
        for (int g = 0; g < kRowsPerPass / U; ++g) {
            if (g * U < R) {                              // warp-uniform
                uint4 nxt[U][NV];
#pragma unroll
                for (int u = 0; u < U; ++u) {             // next row group in flight while this one is consumed
                    const int row = (g + 1) * U + u;
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        nxt[u][j] = (row < R && vok[j]) ? ldg_rows_v4(wbase + (size_t)row * K + voff[j]) : zero4;
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < NV; ++j) rows_consume(cur[u][j], xr[j], acc[g * U + u]);
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < NV; ++j) cur[u][j] = nxt[u][j];
            }
        }

        // lanes -> warp partial -> shared
#pragma unroll
        for (int i = 0; i < kRowsPerPass; ++i) {
            float t = acc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
            if (lane == 0) part[warp][i] = t;
        }
        __syncthreads();
        if ((int)threadIdx.x < R) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < kRowsWarps; ++w) v += part[w][threadIdx.x];
            const int n = pass + threadIdx.x;
            if (v != v) v = slow_dot_masked(p.A, p.B + (size_t)n * K, K);
            epi_store(p.epi, 0, n, epi_apply(p.epi, v, 0, n));
        }
        __syncthreads();
    }
}

bool gemv_rows_supported(const MMArgs& a)
{
    if (a.M != 1 || a.K < 16 || (a.K % 16) != 0 || !aligned(a.A, 16) || !aligned(a.B, 16)) return false;
    const int nvecK = a.K / 16;
    const int vpw = (nvecK + kRowsWarps - 1) / kRowsWarps;
    return vpw <= 32 * 4 && a.N >= 2 * device_info().sm_count;
}

int launch_gemv_rows(const MMArgs& a)
{
    if (!gemv_rows_supported(a)) return FP8B_ERR_UNSUPPORTED;
    GemvRowsParams p;
    p.A = a.A; p.B = a.B; p.N = a.N; p.K = a.K;
    const int nvecK = a.K / 16;
    p.vpw = (nvecK + kRowsWarps - 1) / kRowsWarps;
    p.epi = make_epi(a);
    const int nv = (p.vpw + 31) / 32;
    const int grid = device_info().sm_count;
    switch (nv) {
        case 1: fp8_gemv_rows_kernel<1><<<grid, kRowsThreads, 0, a.st>>>(p); break;
        case 2: fp8_gemv_rows_kernel<2><<<grid, kRowsThreads, 0, a.st>>>(p); break;
        case 3: fp8_gemv_rows_kernel<3><<<grid, kRowsThreads, 0, a.st>>>(p); break;
        default: fp8_gemv_rows_kernel<4><<<grid, kRowsThreads, 0, a.st>>>(p); break;
    }
    return after_launch();
}

}  // namespace fp8b
