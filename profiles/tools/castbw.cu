// Experiment: does a 256-bit (ld.global.v8.b32) access pattern with L2::evict_first hints move a bf16 -> fp8 cast
// closer to the copy roofline than the library's 128-bit pattern?  Same codec arithmetic (fp8_codec.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fp8-mps-metal_b200/csrc -o profiles/tools/castbw profiles/tools/castbw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "fp8_codec.cuh"
using namespace fp8b;

template <int MODE, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) k(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t nunits)
{
    // unit = 16 elements (32 B in, 16 B out) for MODE 1/2; 8 elements (16 B in, 8 B out) for MODE 0
    const size_t tile = (size_t)THREADS * UNROLL;
    for (size_t t = blockIdx.x; t * tile < nunits; t += gridDim.x) {
        const size_t v0 = t * tile + threadIdx.x;
        if (MODE == 0) {
            uint4 w[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t v = v0 + (size_t)u * THREADS;
                if (v < nunits) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[u].x), "=r"(w[u].y), "=r"(w[u].z), "=r"(w[u].w) : "l"(in + v * 16));
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t v = v0 + (size_t)u * THREADS;
                if (v < nunits) {
                    uint32_t o0 = (uint32_t)enc2_bf16x2(w[u].x) | ((uint32_t)enc2_bf16x2(w[u].y) << 16);
                    uint32_t o1 = (uint32_t)enc2_bf16x2(w[u].z) | ((uint32_t)enc2_bf16x2(w[u].w) << 16);
                    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(out + v * 8), "r"(o0), "r"(o1) : "memory");
                }
            }
        } else {
            uint32_t w[UNROLL][8];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t v = v0 + (size_t)u * THREADS;
                if (v < nunits) {
                    if (MODE == 1)
                        asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]), "=r"(w[u][4]), "=r"(w[u][5]), "=r"(w[u][6]), "=r"(w[u][7]) : "l"(in + v * 32));
                    else
                        asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]), "=r"(w[u][4]), "=r"(w[u][5]), "=r"(w[u][6]), "=r"(w[u][7]) : "l"(in + v * 32));
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t v = v0 + (size_t)u * THREADS;
                if (v < nunits) {
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = (uint32_t)enc2_bf16x2(w[u][2 * j]) | ((uint32_t)enc2_bf16x2(w[u][2 * j + 1]) << 16);
                    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out + v * 16), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                }
            }
        }
    }
}

template <int MODE, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) kd(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t nunits)
{
    const size_t tile = (size_t)THREADS * UNROLL;
    for (size_t t = blockIdx.x; t * tile < nunits; t += gridDim.x) {
        const size_t v0 = t * tile + threadIdx.x;
        if (MODE == 0) {
            uint2 w[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t v = v0 + (size_t)u * THREADS;
                if (v < nunits) asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(w[u].x), "=r"(w[u].y) : "l"(in + v * 8));
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t v = v0 + (size_t)u * THREADS;
                if (v < nunits) {
                    uint4 o;
                    dec4_f16x2(w[u].x, o.x, o.y); dec4_f16x2(w[u].y, o.z, o.w);
                    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out + v * 16), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                }
            }
        } else {
            uint4 w[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t v = v0 + (size_t)u * THREADS;
                if (v < nunits) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[u].x), "=r"(w[u].y), "=r"(w[u].z), "=r"(w[u].w) : "l"(in + v * 16));
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t v = v0 + (size_t)u * THREADS;
                if (v < nunits) {
                    uint32_t o[8];
                    dec4_f16x2(w[u].x, o[0], o[1]); dec4_f16x2(w[u].y, o[2], o[3]);
                    dec4_f16x2(w[u].z, o[4], o[5]); dec4_f16x2(w[u].w, o[6], o[7]);
                    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                                 :: "l"(out + v * 32), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
                }
            }
        }
    }
}

template <int MODE, int THREADS, int UNROLL>
static void rund(const char* name, const uint8_t* in, uint8_t* out, size_t n, int ctas_per_sm)
{
    const size_t nunits = n / (MODE == 0 ? 8 : 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) kd<MODE, THREADS, UNROLL><<<148 * ctas_per_sm, THREADS>>>(in, out, nunits);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) kd<MODE, THREADS, UNROLL><<<148 * ctas_per_sm, THREADS>>>(in, out, nunits);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("DEQUANT %-36s T%-4d U%d x%d/SM: %7.0f GB/s  (%s)\n", name, THREADS, UNROLL, ctas_per_sm, 3.0 * n / (ms / 5 * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
}

template <int MODE, int THREADS, int UNROLL>
static void run(const char* name, const uint8_t* in, uint8_t* out, size_t n, int ctas_per_sm)
{
    const size_t nunits = n / (MODE == 0 ? 8 : 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) k<MODE, THREADS, UNROLL><<<148 * ctas_per_sm, THREADS>>>(in, out, nunits);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k<MODE, THREADS, UNROLL><<<148 * ctas_per_sm, THREADS>>>(in, out, nunits);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s T%-4d U%d x%d/SM: %7.0f GB/s  (%s)\n", name, THREADS, UNROLL, ctas_per_sm, 3.0 * n / (ms / 5 * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const size_t n = 1ull << 32;
    uint8_t *in, *out;
    cudaMalloc(&in, n * 2); cudaMalloc(&out, n);
    cudaMemset(in, 0x3C, n * 2);
    run<0, 1024, 4>("128-bit loads (library shape)", in, out, n, 1);
    run<1, 512, 4>("256-bit loads", in, out, n, 1);
    run<1, 512, 8>("256-bit loads", in, out, n, 1);
    run<1, 256, 8>("256-bit loads", in, out, n, 1);
    run<1, 256, 4>("256-bit loads", in, out, n, 2);
    run<1, 384, 4>("256-bit loads", in, out, n, 1);
    run<1, 640, 4>("256-bit loads", in, out, n, 1);
    run<1, 768, 4>("256-bit loads", in, out, n, 1);
    run<1, 128, 4>("256-bit loads", in, out, n, 4);
    run<1, 512, 4>("256-bit loads", in, out, n, 1);
    // decode: `out` (n bytes) is the fp8 source, `in` (2n bytes) the fp16 destination
    rund<0, 512, 8>("64-bit loads / 128-bit stores (library)", out, const_cast<uint8_t*>(in), n, 1);
    rund<1, 512, 4>("128-bit loads / 256-bit stores", out, const_cast<uint8_t*>(in), n, 1);
    rund<1, 512, 8>("128-bit loads / 256-bit stores", out, const_cast<uint8_t*>(in), n, 1);
    rund<1, 256, 4>("128-bit loads / 256-bit stores", out, const_cast<uint8_t*>(in), n, 2);
    rund<1, 256, 8>("128-bit loads / 256-bit stores", out, const_cast<uint8_t*>(in), n, 1);
    rund<1, 1024, 4>("128-bit loads / 256-bit stores", out, const_cast<uint8_t*>(in), n, 1);
    rund<1, 1024, 2>("128-bit loads / 256-bit stores", out, const_cast<uint8_t*>(in), n, 1);
    rund<0, 512, 8>("64-bit loads / 128-bit stores (library)", out, const_cast<uint8_t*>(in), n, 1);
    return 0;
}
