// FP8 GEMV, M = 1..16, as ONE persistent kernel: every SM owns an equal share of the weight rows and pulls them
// through a shared-memory ring with TMA bulk copies.
//     C[m, n] = epi( sum_k dec(A[m,k]) * dec(B[n,k]) )
//
// Replaces fp8_scaled_vecmat_kernel (fp8_matmul.metal:155-210) and the M = 2..16 use of fp8_scaled_matmul_kernel
// (:99-147, dispatch rule fp8_mps_native.py:208).  HBM-bound: B (N*K bytes) is read exactly once.
//
// Why another GEMV kernel.  The round-1 kernels (fp8_gemv.cu, fp8_gemv_mma.cu) launch one CTA per 8 / 16 rows:
// 256..512 CTAs on 148 SMs, i.e. 3 or 4 CTAs per SM (a built-in 13 % imbalance), each of which starts its loads
// only after it has been scheduled and has computed its addresses -- half of a 5 us launch was ramp, tail and
// imbalance (ncu: SMs busy 50-54 % of the launch).  Here:
//   * grid = one CTA per SM; CTA c owns floor(N/G) or ceil(N/G) CONSECUTIVE rows (byte shares equal to within one
//     row, < 4 %), processed as tiles of 16 rows (the m16 of the MMA; the last tile of a CTA may be partial);
//   * warp 8 is the PRODUCER: one cp.async.bulk (SASS UBLKCP) per weight row and K-segment into a ring of
//     STAGES x 16 rows x <=2 KB, completion counted on an mbarrier -- the whole ring (>= 100 KB per SM, 15 MB
//     per GPU) is in flight a few hundred cycles after launch, before any thread has touched an address;
//   * warps 0..7 are CONSUMERS: per 64-byte k-chunk of a 16-row tile a lane reads 16 bytes of rows g and g+8
//     (conflict-free LDS.128: rows are padded by 64 bytes), decodes them with cvt.rn.f16x2.e4m3x2 (SASS F2FP, exact) and
//     issues four mma.sync.m16n8k16 (f16 x f16 -> f32) with the roles swapped -- the 16-row operand is the WEIGHT
//     tile, the 8-column operand is up to 8 activation rows, so M = 4 costs what M = 1 costs.  (The dot product is
//     invariant under a permutation of k applied to both operands, which is what lets a lane feed contiguous
//     bytes.)  The F2FP conversions are the kernel's dominant instruction count; the e4m3 form of mma.sync would
//     also convert the activation operand again for every weight tile (+50 % F2FP, measured 9 % slower), so the
//     activations are decoded ONCE per CTA to fp16 in shared memory.  Two accumulator sets halve the dependent-MMA
//     chain.  The 8 warps take the chunks of a stage round-robin and keep their 16 x 8 partial tiles in registers
//     across the K-segments of a row tile; one shared-memory reduction in warp order (deterministic) per row tile,
//     with the scale / bias / scale_result / cast epilogue fused into it;
// Programmatic dependent launch as in the other GEMVs: launch_dependents first thing; the producer waits for the
// predecessor grid before its first copy unless the caller declared the weights static (FP8B_OPT_STATIC_WEIGHTS),
// the consumers always wait before they stage x or store.
// NaN bytes (0x7F/0xFF) produce a NaN accumulator where the reference decodes 0 (metal:21): exactly those outputs
// are recomputed with the masked scalar loop.
#include "fp8_mm.cuh"
#include "fp8_async.cuh"

namespace fp8b {

constexpr int kRingConsumerWarps = 8;
constexpr int kRingThreads = 32 * (kRingConsumerWarps + 1);
constexpr int kRingRows = 16;               // weight rows per tile = the m16 of the MMA
constexpr int kRingMaxSeg = 2048;           // bytes of K per stage and row
constexpr int kRingRowPad = 64;             // row pitch = segment + 64: rows g, g+1 land in different bank groups
constexpr int kRingMaxStages = 8;
constexpr int kRingSmemBudget = 220 * 1024;

struct RingParams {
    const uint8_t* A;
    const uint8_t* B;
    int M, N, K;
    int rows_lo;            // every CTA owns rows_lo rows, the first rows_rem CTAs one more
    int rows_rem;
    int kseg;               // bytes of K per stage (multiple of 128, <= kRingMaxSeg)
    int nseg;
    int stages;
    int x_pitch;            // bytes between activation rows (fp16) in shared memory: 2K + 16
    int static_b;           // FP8B_OPT_STATIC_WEIGHTS
    int dbg;                // -DFP8B_PROFILE builds only (FP8B_RING_DEBUG): 1 = consumers only wait/arrive, 2 = + LDS,
                            // 3 = + F2FP (no MMA).  Results are garbage; 0 in every shipped build.
    Epi epi;
};

// D += A (16x16 f16: a0 = row g k-pair 0, a1 = row g+8 pair 0, a2 = row g pair 1, a3 = row g+8 pair 1) * B (16x8 f16)
__device__ __forceinline__ void ring_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 16 e4m3 bytes -> 8 f16x2 words (hardware decode: NaN codes stay NaN on purpose, see the fix-up in the epilogue)
__device__ __forceinline__ void ring_dec16(const uint4& w, uint32_t (&h)[8]) {
    dec4_f16x2_raw(w.x, h[0], h[1]);
    dec4_f16x2_raw(w.y, h[2], h[3]);
    dec4_f16x2_raw(w.z, h[4], h[5]);
    dec4_f16x2_raw(w.w, h[6], h[7]);
}
// one 64-byte k-chunk of a 16-row tile against one 8-column activation tile: four MMAs, alternating accumulators
__device__ __forceinline__ void ring_chunk(float (&c0)[4], float (&c1)[4], const uint32_t (&wa)[8], const uint32_t (&wb)[8],
                                           const uint4& x_lo, const uint4& x_hi) {
    ring_mma(c0, wa[0], wb[0], wa[1], wb[1], x_lo.x, x_lo.y);
    ring_mma(c1, wa[2], wb[2], wa[3], wb[3], x_lo.z, x_lo.w);
    ring_mma(c0, wa[4], wb[4], wa[5], wb[5], x_hi.x, x_hi.y);
    ring_mma(c1, wa[6], wb[6], wa[7], wb[7], x_hi.z, x_hi.w);
}
__device__ __forceinline__ uint4 ring_lds16(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void ring_consumer_sync() {       // the 8 consumer warps only (named barrier 1)
    asm volatile("bar.sync 1, %0;" :: "n"(32 * kRingConsumerWarps) : "memory");
}

template <int NB>        // NB = 1: M <= 8; NB = 2: M <= 16
__global__ void __launch_bounds__(kRingThreads)
fp8_gemv_ring_kernel(const RingParams p)
{
    extern __shared__ __align__(128) uint8_t ring_smem[];
    // layout: [x: M rows x x_pitch][ring: stages x 16 x (kseg + pad)][red: 2 x 8 warps x 16 x 8*NB floats][barriers]
    const int pitch = p.kseg + kRingRowPad;
    const int stage_bytes = kRingRows * pitch;
    const int x_bytes = (p.M * p.x_pitch + 127) & ~127;
    uint8_t* xs = ring_smem;
    uint8_t* ring = ring_smem + x_bytes;
    float* red = reinterpret_cast<float*>(ring + p.stages * stage_bytes);
    constexpr int kRedFloats = kRingConsumerWarps * kRingRows * 8 * NB;
    uint8_t* bar_mem = reinterpret_cast<uint8_t*>(red + 2 * kRedFloats);
    const uint32_t bar_base = smem_u32(bar_mem);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kRingMaxStages + s); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int my_rows = p.rows_lo + (cta < p.rows_rem ? 1 : 0);
    const int row0 = cta * p.rows_lo + min(cta, p.rows_rem);
    const int ntiles = (my_rows + kRingRows - 1) / kRingRows;
    const int K = p.K;

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), kRingConsumerWarps); }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kRingConsumerWarps) {
        // ===================== producer =====================
        if (!p.static_b) pdl_wait();                 // B may have been written by the predecessor (an encode, a transfer)
        uint32_t seq = 0;
        for (int tile = 0; tile < ntiles; ++tile) {
            const int rows = min(kRingRows, my_rows - tile * kRingRows);
            const uint8_t* src_row = p.B + (size_t)(row0 + tile * kRingRows + lane) * K;
            for (int s = 0; s < p.nseg; ++s, ++seq) {
                const int len = min(p.kseg, K - s * p.kseg);
                const int stage = seq % p.stages;
                mbar_wait(empty_bar(stage), ((seq / p.stages) & 1) ^ 1);
                if (lane == 0) mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(rows * len));
                __syncwarp();
                if (lane < rows)
                    bulk_load_1d(smem_u32(ring + stage * stage_bytes + lane * pitch), src_row + (size_t)s * p.kseg,
                                 (uint32_t)len, full_bar(stage));
            }
        }
    } else {
        // ===================== consumers =====================
        const int g = lane >> 2, t = lane & 3;
        pdl_wait();                                  // x, scales, bias and the output buffer belong to the stream order
        // stage the activations once per CTA: e4m3 bytes -> fp16 (exact), row pitch 2K + 16 bytes
        {
            const int vec_per_row = K >> 4;
            for (int i = threadIdx.x; i < p.M * vec_per_row; i += 32 * kRingConsumerWarps) {
                const int m = i / vec_per_row, v = i - m * vec_per_row;
                uint4 xb;
                asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"      // coherent: x is the predecessor's output
                             : "=r"(xb.x), "=r"(xb.y), "=r"(xb.z), "=r"(xb.w) : "l"(p.A + (size_t)m * K + (size_t)v * 16));
                uint32_t h[8];
                ring_dec16(xb, h);
                uint4* dst = reinterpret_cast<uint4*>(xs + (size_t)m * p.x_pitch + (size_t)v * 32);
                dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
                dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
            }
        }
        ring_consumer_sync();
        const bool xa_ok = g < p.M, xb_ok = (NB == 2) && (g + 8 < p.M);
        const uint32_t xa_base = smem_u32(xs + (xa_ok ? g : 0) * p.x_pitch + 32 * t);
        const uint32_t xb_base = smem_u32(xs + (xb_ok ? g + 8 : 0) * p.x_pitch + 32 * t);
        const uint32_t w_lane = (uint32_t)(g * pitch + 16 * t);
        const uint32_t ring_base = smem_u32(ring);
        const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

        uint32_t seq = 0;
        for (int tile = 0; tile < ntiles; ++tile) {
            float c[NB][2][4];
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int i = 0; i < 4; ++i) { c[b][0][i] = 0.0f; c[b][1][i] = 0.0f; }
            for (int s = 0; s < p.nseg; ++s, ++seq) {
                const int len = min(p.kseg, K - s * p.kseg);
                const int chunks = len >> 6;
                const int stage = seq % p.stages;
                mbar_wait(full_bar(stage), (seq / p.stages) & 1);
                const uint32_t wbase = ring_base + (uint32_t)(stage * stage_bytes) + w_lane;
                const uint32_t koff = (uint32_t)(s * p.kseg) * 2u;
#pragma unroll 2
                for (int ch = warp; ch < chunks; ch += kRingConsumerWarps) {
#ifdef FP8B_PROFILE
                    if (p.dbg == 1) continue;
#endif
                    const uint4 w0 = ring_lds16(wbase + 64u * ch);
                    const uint4 w1 = ring_lds16(wbase + 8u * pitch + 64u * ch);
#ifdef FP8B_PROFILE
                    if (p.dbg == 2) { c[0][0][0] += __uint_as_float(w0.x ^ w1.y ^ w0.z ^ w1.w); continue; }
#endif
                    uint4 xl = zero4, xh = zero4;
                    if (xa_ok) { xl = ring_lds16(xa_base + koff + 128u * ch); xh = ring_lds16(xa_base + koff + 128u * ch + 16u); }
                    uint32_t wa[8], wb[8];
                    ring_dec16(w0, wa);
                    ring_dec16(w1, wb);
#ifdef FP8B_PROFILE
                    if (p.dbg == 3) {
                        uint32_t acc = xl.x;
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc ^= wa[j] ^ wb[j];
                        c[0][0][0] += __uint_as_float(acc);
                        continue;
                    }
#endif
                    ring_chunk(c[0][0], c[0][1], wa, wb, xl, xh);
                    if (NB == 2) {
                        uint4 yl = zero4, yh = zero4;
                        if (xb_ok) { yl = ring_lds16(xb_base + koff + 128u * ch); yh = ring_lds16(xb_base + koff + 128u * ch + 16u); }
                        ring_chunk(c[NB - 1][0], c[NB - 1][1], wa, wb, yl, yh);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar(stage));       // this warp is done with the slot
            }
            // reduce the 8 warps' partial tiles (warp order), then the fused epilogue.  Double-buffered: the barrier
            // of tile i+1 separates the readers of buffer (i & 1) from its next writers.
            float* rbuf = red + (tile & 1) * kRedFloats;
            constexpr int MC = 8 * NB;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                float* r = rbuf + (warp * kRingRows) * MC + 8 * b + 2 * t;
                r[g * MC] = c[b][0][0] + c[b][1][0]; r[g * MC + 1] = c[b][0][1] + c[b][1][1];
                r[(g + 8) * MC] = c[b][0][2] + c[b][1][2]; r[(g + 8) * MC + 1] = c[b][0][3] + c[b][1][3];
            }
            ring_consumer_sync();
            const int rows = min(kRingRows, my_rows - tile * kRingRows);
            for (int idx = threadIdx.x; idx < kRingRows * p.M; idx += 32 * kRingConsumerWarps) {
                const int r = idx & (kRingRows - 1), m = idx >> 4;
                if (r < rows) {
                    float v = 0.0f;
#pragma unroll
                    for (int w = 0; w < kRingConsumerWarps; ++w) v += rbuf[(w * kRingRows + r) * MC + m];
                    const int n = row0 + tile * kRingRows + r;
                    if (v != v) v = slow_dot_masked(p.A + (size_t)m * K, p.B + (size_t)n * K, K);
                    epi_store(p.epi, m, n, epi_apply(p.epi, v, m, n));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ host side

static bool ring_plan(const MMArgs& a, RingParams& p, int& smem_bytes)
{
    const DeviceInfo& di = device_info();
    if (a.M < 1 || a.M > 16 || a.K < 64 || (a.K % 64) != 0) return false;
    if (!aligned(a.A, 16) || !aligned(a.B, 16)) return false;
    if (a.a_fmt | a.b_fmt) return false;                                  // e4m3fn x e4m3fn only (e5m2: fp8_gemv_mma.cu)
    if (a.N < 8 * di.sm_count) return false;                              // fewer than 8 rows per SM: the split-K kernels
    const int G = di.sm_count * (tune(kTuneGemvUnroll, 1) == 2 ? 2 : 1);
    p.A = a.A; p.B = a.B; p.M = a.M; p.N = a.N; p.K = a.K;
    p.rows_lo = a.N / G; p.rows_rem = a.N % G;
    p.nseg = (a.K + kRingMaxSeg - 1) / kRingMaxSeg;
    p.kseg = ((a.K + p.nseg - 1) / p.nseg + 127) & ~127;                  // multiple of 128: pitch/16 = 4 (mod 8), so the
    p.nseg = (a.K + p.kseg - 1) / p.kseg;                                 // LDS.128 of rows g, g+1 hit different bank groups
    p.x_pitch = 2 * a.K + 16;                                             // fp16 activations; pitch/16 odd: conflict-free B fragments
    const int nb = a.M > 8 ? 2 : 1;
    const int x_bytes = (a.M * p.x_pitch + 127) & ~127;
    const int red_bytes = 2 * kRingConsumerWarps * kRingRows * 8 * nb * (int)sizeof(float);
    const int bar_bytes = 2 * kRingMaxStages * 8;
    const int stage_bytes = kRingRows * (p.kseg + kRingRowPad);
    const int ctas_per_sm = tune(kTuneGemvUnroll, 1) == 2 ? 2 : 1;         // profiling knob: two half-size rings per SM
    int stages = (kRingSmemBudget / ctas_per_sm - x_bytes - red_bytes - bar_bytes) / stage_bytes;
    if (stages > kRingMaxStages) stages = kRingMaxStages;
    if (stages < 3) return false;                                         // activations too large for a useful ring
    p.stages = stages;
    p.static_b = g_opt_static_weights.load(std::memory_order_relaxed) ? 1 : 0;
    p.dbg = 0;
#ifdef FP8B_PROFILE
    p.dbg = tune_int("FP8B_RING_DEBUG", 0);
#endif
    p.epi = make_epi(a);
    smem_bytes = x_bytes + stages * stage_bytes + red_bytes + bar_bytes;
    return true;
}

bool gemv_ring_supported(const MMArgs& a)
{
    RingParams p;
    int smem = 0;
    return ring_plan(a, p, smem);
}

int launch_gemv_ring(const MMArgs& a)
{
    RingParams p;
    int smem = 0;
    if (!ring_plan(a, p, smem)) return FP8B_ERR_UNSUPPORTED;
    static std::atomic<int> attr1[64], attr2[64];
    const bool pdl = g_opt_pdl.load(std::memory_order_relaxed) != 0;
    const dim3 grid(device_info().sm_count * (tune(kTuneGemvUnroll, 1) == 2 ? 2 : 1)), block(kRingThreads);
    if (a.M <= 8) {
        if (int rc = ensure_max_smem(fp8_gemv_ring_kernel<1>, kRingSmemBudget, attr1)) return rc;
        return launch_ex(fp8_gemv_ring_kernel<1>, grid, block, (size_t)smem, a.st, 1, 1, pdl, p);
    }
    if (int rc = ensure_max_smem(fp8_gemv_ring_kernel<2>, kRingSmemBudget, attr2)) return rc;
    return launch_ex(fp8_gemv_ring_kernel<2>, grid, block, (size_t)smem, a.st, 1, 1, pdl, p);
}

}  // namespace fp8b
