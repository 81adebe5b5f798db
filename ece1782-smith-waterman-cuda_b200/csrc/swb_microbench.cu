// Measurement support: issue-rate micro-benchmarks of the integer SIMD instructions the score kernel is
// made of. bench.py uses the "mix" figure as the denominator of the integer-pipe roofline (SURVEY 8(d):
// MEASURED_PEAKS.json only has HBM and bf16 numbers). Not part of the scoring path.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/swb.h"

#define MB_CHAINS 8

// max of two s16x2 words through the fp16x2 comparator: exact for non-negative halves below 0x7C00 (31744), whose bit
// patterns order like the integers
__device__ __forceinline__ uint32_t swb_hmax2_bits(uint32_t a, uint32_t b)
{
    __half2 x = *reinterpret_cast<__half2 *>(&a), y = *reinterpret_cast<__half2 *>(&b);
    __half2 r = __hmax2(x, y);
    return *reinterpret_cast<uint32_t *>(&r);
}

template <int KIND>
__global__ void __launch_bounds__(256) swb_mb_kernel(uint32_t *out, uint32_t seed, int iters)
{
    uint32_t x[MB_CHAINS], y = 0u, dprev = 0u;
    const uint32_t a = seed * 0x9E3779B9u + threadIdx.x, b = seed ^ 0x00010001u, one = (seed >> 31) + 1u;
#pragma unroll
    for (int c = 0; c < MB_CHAINS; ++c) x[c] = a + c * 0x00030005u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < MB_CHAINS; ++c) {
            if (KIND == 0) x[c] = __viaddmax_s16x2_relu(x[c], b, a);
            if (KIND == 1) x[c] = __vimax3_s16x2(x[c], b, a);
            if (KIND == 2) x[c] = __vadd2(x[c], b);
            if (KIND == 3) asm volatile("prmt.b32 %0, %1, %2, 0xC480;" : "=r"(x[c]) : "r"(x[c]), "r"(b));
            if (KIND == 4) {  // the score kernel's per-cell-pair mix: 2 viaddmax, 1 vadd, 1 prmt, 1/2 vimax3
                uint32_t s;
                asm volatile("prmt.b32 %0, %1, %2, 0xD591;" : "=r"(s) : "r"(x[c]), "r"(b));
                const uint32_t cc = __viaddmax_s16x2_relu(x[c], s, a);
                x[c] = __viaddmax_s16x2(x[c], b, cc);
                x[c] = __vadd2(x[c], b);
                if (c & 1) x[c] = __vimax3_s16x2(x[c], x[c - 1], cc);
            }
            if (KIND == 5) {  // does an IMAD (fma pipe) issue beside the DPX op?
                x[c] = __viaddmax_s16x2_relu(x[c], b, a);
                x[c] = x[c] * one + b;
            }
            if (KIND == 6) x[c] = x[c] * one + b;  // IMAD alone
            if (KIND == 7) x[c] = max(x[c] + b, a);  // scalar add+max (what the compiler makes of int32 cells)
            if (KIND == 9) x[c] = swb_hmax2_bits(x[c] & 0x3fff3fffu, b & 0x3fff3fffu);  // HMNMX2 (+ LOP3)
            if (KIND == 10) {  // does HMNMX2 issue beside the DPX op?
                x[c] = __viaddmax_s16x2_relu(x[c], b, a);
                x[c] = swb_hmax2_bits(x[c], a);
            }
            // which of the kernel's instructions issue beside VIADDMNMX (i.e. not on the same 64-lane pipe)?
            if (KIND == 11) { x[c] = __viaddmax_s16x2_relu(x[c], b, a); x[c] = __vadd2(x[c], b); }
            if (KIND == 12) {
                x[c] = __viaddmax_s16x2_relu(x[c], b, a);
                asm volatile("prmt.b32 %0, %1, %2, 0xC480;" : "=r"(x[c]) : "r"(x[c]), "r"(b));
            }
            if (KIND == 13) { x[c] = __viaddmax_s16x2_relu(x[c], b, a); x[c] = __vimax3_s16x2(x[c], b, a); }
            if (KIND == 14) {  // the score kernel's per-cell-pair mix since r2x: prmt, 2 vadd2, vimax3.relu, 1/2 vimax3
                uint32_t sc;
                asm volatile("prmt.b32 %0, %1, %2, 0xD591;" : "=r"(sc) : "r"(x[c]), "r"(b));
                const uint32_t dd = __vadd2(x[c], sc);  // used twice (cell and running maximum): stays a vadd2
                const uint32_t hh = __vimax3_s16x2_relu(dd, a, x[c]);
                x[c] = __vadd2(hh, b);
                if (c & 1) y = __vimax3_s16x2(y, dprev, dd);
                dprev = dd;
            }
            if (KIND == 8) {  // the biased policy's mix: prmt, vimax3, viaddmax, 1/2 vimax3 (ALU) + 3 imad (FMA)
                uint32_t s, ds, l3;
                asm volatile("prmt.b32 %0, %1, %2, 0xD591;" : "=r"(s) : "r"(x[c]), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ds) : "r"(x[c]), "r"(one), "r"(s));
                asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(l3) : "r"(x[c]), "r"(one), "r"(b));
                const uint32_t cc = __vimax3_s16x2(ds, l3, a);
                x[c] = __viaddmax_s16x2(x[c], b, cc);
                asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(x[c]) : "r"(x[c]), "r"(one), "r"(b));
                if (c & 1) x[c] = __vimax3_s16x2(x[c], x[c - 1], cc);
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < MB_CHAINS; ++c) r ^= x[c];
    r ^= y;
    if (r == 0x12345678u) out[0] = r;
}

static const double kInstrPerIter[15] = {MB_CHAINS, MB_CHAINS, MB_CHAINS, MB_CHAINS, MB_CHAINS * 4.5, MB_CHAINS * 2.0,
                                         MB_CHAINS, MB_CHAINS, MB_CHAINS * 6.5, MB_CHAINS * 3.0, MB_CHAINS * 2.0,
                                         MB_CHAINS * 2.0, MB_CHAINS * 2.0, MB_CHAINS * 2.0, MB_CHAINS * 4.5};

// kind 0 viaddmax.relu, 1 vimax3, 2 vadd2, 3 prmt, 4 V16 mix of round 1, 14 V16 mix now, 5 viaddmax+imad, 6 imad, 7 scalar add/max, 8 V16B mix.
// Returns giga lane-instructions per second (warp instructions x 32) over the whole GPU.
extern "C" int swb_microbench(int device, int kind, double *glane_instr_per_s, double *ms_out)
{
    if (kind < 0 || kind > 14 || !glane_instr_per_s) return SWB_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return SWB_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SWB_ERR_CUDA;
    uint32_t *d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return SWB_ERR_CUDA;
    const int grid = prop.multiProcessorCount * 8, block = 256, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        switch (kind) {
        case 0: swb_mb_kernel<0><<<grid, block>>>(d, 7u + rep, iters); break;
        case 1: swb_mb_kernel<1><<<grid, block>>>(d, 7u + rep, iters); break;
        case 2: swb_mb_kernel<2><<<grid, block>>>(d, 7u + rep, iters); break;
        case 3: swb_mb_kernel<3><<<grid, block>>>(d, 7u + rep, iters); break;
        case 4: swb_mb_kernel<4><<<grid, block>>>(d, 7u + rep, iters); break;
        case 5: swb_mb_kernel<5><<<grid, block>>>(d, 7u + rep, iters); break;
        case 6: swb_mb_kernel<6><<<grid, block>>>(d, 7u + rep, iters); break;
        case 7: swb_mb_kernel<7><<<grid, block>>>(d, 7u + rep, iters); break;
        case 8: swb_mb_kernel<8><<<grid, block>>>(d, 7u + rep, iters); break;
        case 9: swb_mb_kernel<9><<<grid, block>>>(d, 7u + rep, iters); break;
        case 11: swb_mb_kernel<11><<<grid, block>>>(d, 7u + rep, iters); break;
        case 12: swb_mb_kernel<12><<<grid, block>>>(d, 7u + rep, iters); break;
        case 13: swb_mb_kernel<13><<<grid, block>>>(d, 7u + rep, iters); break;
        case 14: swb_mb_kernel<14><<<grid, block>>>(d, 7u + rep, iters); break;
        default: swb_mb_kernel<10><<<grid, block>>>(d, 7u + rep, iters); break;
        }
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return SWB_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    const double lane_instr = (double)grid * block * iters * kInstrPerIter[kind];
    *glane_instr_per_s = lane_instr / (best * 1e-3) * 1e-9;
    if (ms_out) *ms_out = best;
    return SWB_OK;
}
