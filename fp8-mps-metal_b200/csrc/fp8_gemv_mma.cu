// FP8 small-batch GEMV on the warp-level tensor-core path, M = 1..16:
//     C[m, n] = epi( sum_k dec(A[m,k]) * dec(B[n,k]) )
//
// Replaces the M = 2..16 use of fp8_scaled_matmul_kernel (fp8_matmul.metal:99-147, selected by
// fp8_mps_native.py:208): in the reference every output element re-reads its weight row, so M rows
// of activations cost M passes over the weights.  Here the weights stream from HBM exactly ONCE for
// all M rows, which keeps the kernel HBM-bound up to M = 16 -- on CUDA cores M FMAs per weight byte
// would make it issue-bound from M = 4 (measured: 28 % of HBM bandwidth at M = 4).
//
// mma.sync.m16n8k32 (e4m3 x e4m3 -> f32) with the roles swapped: the 16-row operand is a 16 x 32
// tile of WEIGHTS, the 8-column operand is 8 rows of activations, so D[i][j] = out[j][n0 + i].
// The dot product is invariant under any permutation of k applied to both operands, so each lane
// loads one contiguous 16-byte vector per weight row (rows g and g+8 of the fragment, k offset
// 16*t) and one 16-byte vector of x, and feeds the four 32-bit words to two MMAs: 64 values of k
// per step with three 128-bit loads, no shared memory, no prologue, no barrier before the main
// loop.  The 8 warps of a CTA split K and reduce their 16 x 16 partial tiles through shared memory
// in warp order (deterministic); scales, bias, scale_result and the output cast are fused into that
// final step.  NaN bytes (0x7F/0xFF) produce a NaN accumulator where the reference decodes 0
// (metal:21); exactly those outputs are recomputed with the masked scalar loop.
#include "fp8_mm.cuh"

namespace fp8b {

constexpr int kMmaThreads = 256;
constexpr int kMmaWarps = kMmaThreads / 32;
constexpr int kMmaRows = 16;           // weight rows per CTA (the m16 of the MMA)

struct GemvMmaParams {
    const uint8_t* A;
    const uint8_t* B;
    int M, N, K;
    int k_per_warp;                    // multiple of 64
    int static_b;                      // FP8B_OPT_STATIC_WEIGHTS: B may be read before the predecessor completes
    Epi epi;
};

__device__ __forceinline__ uint4 ldg_stream16(const uint8_t* p) {
    uint4 r;
    // coherent (no .nc): every launch uses programmatic dependent launch, and ptxas hoists non-coherent loads above
    // griddepcontrol.wait -- B may have been written by the kernel just before (an encode, a transfer)
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_cached16(const uint8_t* p) {
    uint4 r;
    // NOT .nc: x is typically written by the predecessor kernel, and ptxas moves non-coherent loads above
    // griddepcontrol.wait (seen in SASS: LDG.CONSTANT ahead of ACQBULK), which reads stale activations.
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// WF / XF: format of the 16-row operand (weights) and of the 8-column operand (activations); 0 = e4m3fn, 1 = e5m2.
template <int WF, int XF>
__device__ __forceinline__ void mma_f8_m16n8k32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                uint32_t b0, uint32_t b1) {
#define FP8B_MMA(TA, TB) \
    asm volatile("mma.sync.aligned.m16n8k32.row.col.f32." TA "." TB ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) \
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1))
    if (WF == 0 && XF == 0) FP8B_MMA("e4m3", "e4m3");
    else if (WF == 0 && XF == 1) FP8B_MMA("e4m3", "e5m2");
    else if (WF == 1 && XF == 0) FP8B_MMA("e5m2", "e4m3");
    else FP8B_MMA("e5m2", "e5m2");
#undef FP8B_MMA
}

template <int NB, int BATCH, int WF = 0, int XF = 0>   // NB = 1: M <= 8, NB = 2: M <= 16; BATCH 64-byte k-steps per load group
__global__ void __launch_bounds__(kMmaThreads)
fp8_gemv_mma_kernel(const GemvMmaParams p)
{
    __shared__ float part[kMmaWarps][kMmaRows][8 * NB + 1];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int n0 = blockIdx.x * kMmaRows;
    const int K = p.K;
    const int r0 = min(n0 + g, p.N - 1), r1 = min(n0 + g + 8, p.N - 1);
    const uint8_t* w0 = p.B + (size_t)r0 * K + 16 * t;
    const uint8_t* w1 = p.B + (size_t)r1 * K + 16 * t;
    const bool xa_ok = g < p.M, xb_ok = (NB == 2) && (g + 8 < p.M);
    const uint8_t* x0 = p.A + (size_t)(xa_ok ? g : 0) * K + 16 * t;
    const uint8_t* x1 = p.A + (size_t)(xb_ok ? g + 8 : 0) * K + 16 * t;

    // M <= 8: two accumulator sets, so the two MMAs of a k-step are independent -- halves the chain of dependent
    // tensor-core instructions each warp has to get through once its loads have landed (C3: 6.08 -> 5.83 us)
    float c[NB][4], d[NB][4];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int i = 0; i < 4; ++i) { c[b][i] = 0.0f; d[b][i] = 0.0f; }

    const int k_lo = warp * p.k_per_warp;
    const int k_hi = min(K, k_lo + p.k_per_warp);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

    uint4 wa[BATCH], wb[BATCH], xa[BATCH], xb[BATCH];
    auto load_w = [&](uint4 (&a)[BATCH], uint4 (&b)[BATCH], int kb) {
#pragma unroll
        for (int s = 0; s < BATCH; ++s) {
            const int off = kb + 64 * s;
            const bool ok = off + 16 * t < k_hi;                 // k_hi is a multiple of 16
            a[s] = ok ? ldg_stream16(w0 + off) : zero4;
            b[s] = ok ? ldg_stream16(w1 + off) : zero4;
        }
    };
    auto load_x = [&](uint4 (&a)[BATCH], uint4 (&b)[BATCH], int kb) {
#pragma unroll
        for (int s = 0; s < BATCH; ++s) {
            const int off = kb + 64 * s;
            const bool ok = off + 16 * t < k_hi;
            a[s] = (ok && xa_ok) ? ldg_cached16(x0 + off) : zero4;
            if (NB == 2) b[s] = (ok && xb_ok) ? ldg_cached16(x1 + off) : zero4;
        }
    };

    // Programmatic dependent launch: let the next kernel on the stream start scheduling now.  With
    // static weights the first group of B goes in flight BEFORE we wait for the predecessor (whose
    // output is typically our x); everything the predecessor may have written is read after the wait.
    pdl_launch_dependents();
    if (!p.static_b) pdl_wait();
    load_w(wa, wb, k_lo);
    if (p.static_b) pdl_wait();
    load_x(xa, xb, k_lo);

    for (int kb = k_lo; kb < k_hi; kb += 64 * BATCH) {
        // next group in flight while this one is multiplied
        uint4 nwa[BATCH], nwb[BATCH], nxa[BATCH], nxb[BATCH];
        load_w(nwa, nwb, kb + 64 * BATCH);
        load_x(nxa, nxb, kb + 64 * BATCH);
#pragma unroll
        for (int s = 0; s < BATCH; ++s) {
            if (NB == 1) {
                mma_f8_m16n8k32<WF, XF>(c[0], wa[s].x, wb[s].x, wa[s].y, wb[s].y, xa[s].x, xa[s].y);
                mma_f8_m16n8k32<WF, XF>(d[0], wa[s].z, wb[s].z, wa[s].w, wb[s].w, xa[s].z, xa[s].w);
            } else {                     // two activation tiles are two independent chains already (and four sets spill occupancy)
                mma_f8_m16n8k32<WF, XF>(c[0], wa[s].x, wb[s].x, wa[s].y, wb[s].y, xa[s].x, xa[s].y);
                mma_f8_m16n8k32<WF, XF>(c[NB - 1], wa[s].x, wb[s].x, wa[s].y, wb[s].y, xb[s].x, xb[s].y);
                mma_f8_m16n8k32<WF, XF>(c[0], wa[s].z, wb[s].z, wa[s].w, wb[s].w, xa[s].z, xa[s].w);
                mma_f8_m16n8k32<WF, XF>(c[NB - 1], wa[s].z, wb[s].z, wa[s].w, wb[s].w, xb[s].z, xb[s].w);
            }
        }
#pragma unroll
        for (int s = 0; s < BATCH; ++s) {
            wa[s] = nwa[s]; wb[s] = nwb[s]; xa[s] = nxa[s];
            if (NB == 2) xb[s] = nxb[s];
        }
    }

    // fragment -> shared: c0:(g, 2t) c1:(g, 2t+1) c2:(g+8, 2t) c3:(g+8, 2t+1)
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (NB == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) c[b][i] += d[b][i];
        }
        part[warp][g][8 * b + 2 * t] = c[b][0];
        part[warp][g][8 * b + 2 * t + 1] = c[b][1];
        part[warp][g + 8][8 * b + 2 * t] = c[b][2];
        part[warp][g + 8][8 * b + 2 * t + 1] = c[b][3];
    }
    __syncthreads();

    // 256 threads = 16 weight rows x 16 activation rows; consecutive threads -> consecutive n
    const int row = threadIdx.x & 15, m = threadIdx.x >> 4;
    const int n = n0 + row;
    if (m < p.M && m < 8 * NB && n < p.N) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < kMmaWarps; ++w) v += part[w][row][m];
        if (v != v) v = (WF | XF) ? slow_dot_fmt(p.A + (size_t)m * K, p.B + (size_t)n * K, K, XF, WF)
                              : slow_dot_masked(p.A + (size_t)m * K, p.B + (size_t)n * K, K);
        epi_store(p.epi, m, n, epi_apply(p.epi, v, m, n));
    }
}

bool gemv_mma_supported(const MMArgs& a)
{
    return a.M >= 1 && a.M <= 16 && a.K >= 16 && (a.K % 16 == 0) && aligned(a.A, 16) && aligned(a.B, 16);
}

int launch_gemv_mma(const MMArgs& a)
{
    if (!gemv_mma_supported(a)) return FP8B_ERR_UNSUPPORTED;
    GemvMmaParams p;
    p.A = a.A; p.B = a.B; p.M = a.M; p.N = a.N; p.K = a.K;
    p.k_per_warp = (((a.K + kMmaWarps - 1) / kMmaWarps) + 63) & ~63;
    p.epi = make_epi(a);
    p.static_b = (g_opt_pdl.load(std::memory_order_relaxed) && (a.chain_pdl || g_opt_static_weights.load(std::memory_order_relaxed))) ? 1 : 0;
    const bool pdl = g_opt_pdl.load(std::memory_order_relaxed) != 0;   // always: resident while the predecessor drains
    const int grid = (a.N + kMmaRows - 1) / kMmaRows;
    const int batch = tune(kTuneGemvBatch, 4);
    if (a.a_fmt | a.b_fmt) {                // an e5m2 operand: same kernel, other MMA types (WF = weights, XF = activations)
        const int f = a.b_fmt * 2 + a.a_fmt;
        if (a.M <= 8) {
            if (f == 1) return launch_ex(fp8_gemv_mma_kernel<1, 4, 0, 1>, dim3(grid), dim3(kMmaThreads), 0, a.st, 1, 1, pdl, p);
            if (f == 2) return launch_ex(fp8_gemv_mma_kernel<1, 4, 1, 0>, dim3(grid), dim3(kMmaThreads), 0, a.st, 1, 1, pdl, p);
            return launch_ex(fp8_gemv_mma_kernel<1, 4, 1, 1>, dim3(grid), dim3(kMmaThreads), 0, a.st, 1, 1, pdl, p);
        }
        if (f == 1) return launch_ex(fp8_gemv_mma_kernel<2, 4, 0, 1>, dim3(grid), dim3(kMmaThreads), 0, a.st, 1, 1, pdl, p);
        if (f == 2) return launch_ex(fp8_gemv_mma_kernel<2, 4, 1, 0>, dim3(grid), dim3(kMmaThreads), 0, a.st, 1, 1, pdl, p);
        return launch_ex(fp8_gemv_mma_kernel<2, 4, 1, 1>, dim3(grid), dim3(kMmaThreads), 0, a.st, 1, 1, pdl, p);
    }
    if (a.M <= 8) {
        if (batch == 4) return launch_ex(fp8_gemv_mma_kernel<1, 4>, dim3(grid), dim3(kMmaThreads), 0, a.st, 1, 1, pdl, p);
        if (batch == 1) return launch_ex(fp8_gemv_mma_kernel<1, 1>, dim3(grid), dim3(kMmaThreads), 0, a.st, 1, 1, pdl, p);
        return launch_ex(fp8_gemv_mma_kernel<1, 2>, dim3(grid), dim3(kMmaThreads), 0, a.st, 1, 1, pdl, p);
    }
    return launch_ex(fp8_gemv_mma_kernel<2, 4>, dim3(grid), dim3(kMmaThreads), 0, a.st, 1, 1, pdl, p);
}

}  // namespace fp8b
