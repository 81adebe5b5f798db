// sm_100a kernels of the scan engine: score kernels (instantiations of swb_warp.cuh), query-profile
// builder, device-side DB packer and score scatter. Host-side launch wrappers at the bottom.
#include <atomic>
#include "swb_kernels.h"
#include "swb_warp.cuh"

// ---------------------------------------------------------------------------------------------
struct DevBackend {
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31u); }
    // first work item of this warp: its position among the working warps of the grid (wa = working warps per block, 0 = all)
    __device__ __forceinline__ uint32_t static_item(uint32_t wa) const
    {
        return blockIdx.x * (wa ? wa : (blockDim.x >> 5)) + (threadIdx.x >> 5);
    }
    // index of this warp among the warps of the launch (its colstate region)
    __device__ __forceinline__ uint32_t warp_slot() const { return blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); }
    __device__ __forceinline__ uint32_t shfl_up(uint32_t v, int d, int w) const
    {
        return __shfl_up_sync(0xffffffffu, v, (unsigned)d, w);
    }
    __device__ __forceinline__ uint32_t shfl_xor(uint32_t v, int m, int w) const
    {
        return __shfl_xor_sync(0xffffffffu, v, m, w);
    }
    __device__ __forceinline__ void syncwarp() const { __syncwarp(); }
    __device__ __forceinline__ bool any(bool f) const { return __any_sync(0xffffffffu, f) != 0; }
    __device__ __forceinline__ uint32_t next_tile(uint32_t *counter) const
    {
        uint32_t v = 0;
        if ((threadIdx.x & 31u) == 0) v = atomicAdd(counter, 1u);
        return __shfl_sync(0xffffffffu, v, 0);
    }
    __device__ __forceinline__ void count(uint32_t *p) const { atomicAdd(p, 1u); }
    __device__ __forceinline__ void atomic_max(int32_t *p, int32_t v) const { atomicMax(p, v); }
    // progress counters of SPLIT launches. The boundary stores of all lanes are ordered before the publishing lane's
    // release store by the __syncwarp the caller executes first (barrier synchronisation is part of causality order,
    // and a release is cumulative); the reader polls with acquire loads, so what it reads afterwards is at least as
    // new. No full fence (MEMBAR.SC + cache invalidate) on either side.
    __device__ __forceinline__ void publish(uint32_t *p, uint32_t v) const
    {
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
    // SPLIT launches: the warp copies `rows` profile rows of all 32 codes into its block's shared memory
    __device__ __forceinline__ void stage_rows(int8_t *dst, uint32_t dstride, const int8_t *prof, uint32_t pstride,
                                               uint32_t row0, uint32_t rows) const
    {
        __syncwarp();
        const uint32_t wpr = rows >> 2;
        for (uint32_t i = threadIdx.x & 31u; i < wpr * SWB_ALPHA; i += 32u) {
            const uint32_t code = i / wpr, w = i - code * wpr;
            reinterpret_cast<uint32_t *>(dst + (size_t)code * dstride)[w] =
                __ldg(reinterpret_cast<const uint32_t *>(prof + (size_t)code * pstride + row0) + w);
        }
        __syncwarp();
    }
    // returns the published column count once it is at least `need`; if it was not on the first look, waits until the
    // predecessor is SWB_SPLIT_HYST columns further (or done: `total`), so that the next poll is that far away
    __device__ __forceinline__ uint32_t wait_progress(const uint32_t *p, uint32_t need, uint32_t total) const
    {
        uint32_t v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if (v >= need) return v;
        const uint32_t want = need + SWB_SPLIT_HYST < total ? need + SWB_SPLIT_HYST : total;
        do {
            __nanosleep(200);
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        } while (v < want);
        return v;
    }
    __device__ __forceinline__ void prefetch_l2(const void *p) const
    {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
    __device__ __forceinline__ uint8_t ld_flag(const uint8_t *p) const { return __ldcg(p); }
    __device__ __forceinline__ SwbTile ld_tile(const SwbTile *p) const
    {
        union { SwbTile t; uint4 v[2]; } u;
        u.v[0] = __ldg(reinterpret_cast<const uint4 *>(p));
        u.v[1] = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
        return u.t;
    }
    __device__ __forceinline__ uint32_t ld_code(const uint8_t *p) const { return (uint32_t)__ldg(p); }
    __device__ __forceinline__ uint32_t ld_cg(const uint32_t *p) const { return __ldcg(p); }
    __device__ __forceinline__ uint2 ld_cg2(const uint2 *p) const { return __ldcg(p); }
    __device__ __forceinline__ uint4 ld_cg4(const uint4 *p) const { return __ldcg(p); }
    __device__ __forceinline__ void st_cg(uint32_t *p, uint32_t v) const { __stcg(p, v); }
    __device__ __forceinline__ void st_cg2(uint2 *p, uint2 v) const { __stcg(p, v); }
    __device__ __forceinline__ void st_cg4(uint4 *p, uint4 v) const { __stcg(p, v); }
};

static_assert(sizeof(SwbTile) == 32, "SwbTile must be 32 bytes");

// Persistent blocks: stage the query profile of this chunk in shared memory once, then every warp
// pulls tiles from the shared counter until the list is empty.
// Shared profile layout: [32 codes][smem_rows + 4] int8. The +4 makes consecutive code rows start one
// bank apart, so the 32 lanes of an LDS.32 (same row offset, per-lane code) never conflict: equal
// codes broadcast, different codes hit different banks.
template <int K, class V, int NT, int MINB, bool SPLIT>
__global__ void __launch_bounds__(NT, MINB) swb_score_kernel(const SwbScoreParams p)
{
    extern __shared__ __align__(16) int8_t sprof[];
    // code rows one bank apart (+4) for the word loads of the bulk kernels; SPLIT kernels (one warp per block, mostly
    // one pair per warp: every lane reads the same code's row at its own offset) keep them 16-byte aligned for LDS.128
    const uint32_t sstride = p.smem_rows + (SPLIT ? 16u : (uint32_t)SWB_BULK_LDW);
    if (!SPLIT) {  // SPLIT: every work item stages the rows of its own pass (swb_warp_loop)
        const uint32_t wpr = p.smem_rows >> 2;  // words per code row
        for (uint32_t i = threadIdx.x; i < wpr * SWB_ALPHA; i += NT) {
            const uint32_t code = i / wpr, w = i - code * wpr;
            reinterpret_cast<uint32_t *>(sprof + (size_t)code * sstride)[w] =
                __ldg(reinterpret_cast<const uint32_t *>(p.profile + (size_t)code * p.prof_stride + p.row0) + w);
        }
        __syncthreads();
    }
    if (p.warps_active && (threadIdx.x >> 5) >= p.warps_active) return;
    DevBackend be;
    swb_warp_loop<K, V, SPLIT>(be, p, sprof, sstride);
}

// profile[code][r] = S(q_r, code) + bias for r < qlen, bias (score 0) for the padding rows; bias = gap + t0.
__global__ void swb_profile_kernel(const uint8_t *__restrict__ q, uint32_t qlen, const int8_t *__restrict__ mat,
                                   int bias, int8_t *__restrict__ prof, uint32_t stride, uint32_t rows)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const uint32_t qc = r < qlen ? (uint32_t)(q[r] & 31u) : (uint32_t)SWB_PAD;
#pragma unroll 4
    for (uint32_t code = 0; code < SWB_ALPHA; ++code)
        prof[(size_t)code * stride + r] = (int8_t)(mat[qc * SWB_ALPHA + code] + bias);
}

// One block per tile (grid-stride): gathers the tile's sequences from the raw concatenated codes into the interleaved
// 8-byte words the score kernel streams ([chunk of 4 columns][slot][4 x (seqA, seqB)], same bytes as swb_pack_word,
// which the host emulation uses). The only HBM-bound kernel of the path: one read and one write of the database per
// load. Three versions were measured on the benchmark database (403 MB per launch, swb_pack_time):
//   byte gathers, one output word per thread (16 LDG.U8 each)            282 us  1.4 TB/s  0.22 of the copy rate
//   128-column strips staged through shared memory, block barriers        slower than the first (0.35 vs 0.29 ms, ncu)
//   this one: 16 columns per thread, aligned 32-bit loads + funnel shifts  121 us  3.3 TB/s  0.50
// The byte version was bound by L1 wavefronts (every lane of a load hits another sequence's sector and uses one byte
// of it); here every loaded word is used whole and the stores of a warp are 32 consecutive words.
// 16 consecutive residues of one sequence starting at column col0 as four words (columns past the end hold PAD): five
// aligned 32-bit loads and four funnel shifts instead of sixteen byte loads. raw is a cudaMalloc'ed buffer with slack
// behind the last residue (grow_dev), so the aligned word that holds the last byte is always readable.
__device__ __forceinline__ void swb_pack_load16(const uint8_t *__restrict__ raw, uint64_t off, uint32_t len, uint32_t col0,
                                                uint32_t (&v)[4])
{
    const uint32_t PADW = 0x01010101u * (uint32_t)SWB_PAD;
    if (col0 >= len) {
        v[0] = v[1] = v[2] = v[3] = PADW;
        return;
    }
    const uint64_t a = off + col0;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(raw + (a & ~3ull));
    const uint32_t sh = (uint32_t)(a & 3ull) * 8u;
    const uint32_t left = len - col0;  // residues from col0 on
    uint32_t x[5];
#pragma unroll
    const uint32_t span = (uint32_t)(a & 3ull) + (left < 16u ? left : 16u);  // bytes from w[0] to the last one needed
#pragma unroll
    for (int k = 0; k < 5; ++k) x[k] = (4u * (uint32_t)k < span) ? __ldg(w + k) : 0u;  // never past the word that holds it
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t t = __funnelshift_r(x[k], x[k + 1], sh) & 0x1f1f1f1fu;
        const uint32_t have = left > 4u * (uint32_t)k ? left - 4u * (uint32_t)k : 0u;  // valid bytes of this word
        if (have < 4u) {
            const uint32_t keep = have ? (0xffffffffu >> (8u * (4u - have))) : 0u;
            t = (t & keep) | (PADW & ~keep);
        }
        v[k] = t;
    }
}

__global__ void __launch_bounds__(256) swb_pack_kernel(const SwbTile *__restrict__ tiles, uint32_t ntiles,
                                                       const uint8_t *__restrict__ raw,
                                                       const uint64_t *__restrict__ seq_off,
                                                       const uint32_t *__restrict__ seq_len, uint32_t nseq,
                                                       uint8_t *__restrict__ residues)
{
    for (uint32_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const SwbTile t = tiles[ti];
        const uint32_t P = 32u >> t.logG;
        const uint32_t nch = t.width >> 2;
        const uint32_t ngr = (nch + 3u) >> 2;  // groups of four chunks = 16 columns
        uint64_t *out = reinterpret_cast<uint64_t *>(residues + t.res_off);
        // one thread = 16 columns of one slot (both sequences of the pair): the 32 lanes of a warp write 32 consecutive
        // 8-byte words per chunk, and every load is a 4-byte word of which all bytes are used
        for (uint32_t i = threadIdx.x; i < ngr * P; i += blockDim.x) {
            const uint32_t gq = i >> (5u - t.logG), slot = i & (P - 1u);
            uint32_t va[4], vb[4];
            uint32_t lenA = 0, lenB = 0;
            uint64_t offA = 0, offB = 0;
            const uint64_t sA = 2ull * ((uint64_t)t.first_pair + slot);
            if (slot < t.npairs && sA < nseq) { lenA = seq_len[sA]; offA = seq_off[sA]; }
            if (slot < t.npairs && sA + 1 < nseq) { lenB = seq_len[sA + 1]; offB = seq_off[sA + 1]; }
            swb_pack_load16(raw, offA, lenA, gq * 16u, va);
            swb_pack_load16(raw, offB, lenB, gq * 16u, vb);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t c = gq * 4u + (uint32_t)k;
                if (c < nch) {
                    const uint32_t lo = __byte_perm(va[k], vb[k], 0x5140), hi = __byte_perm(va[k], vb[k], 0x7362);
                    out[(size_t)c * P + slot] = (uint64_t)lo | ((uint64_t)hi << 32);  // A0 B0 A1 B1 | A2 B2 A3 B3
                }
            }
        }
    }
}

// scores in sorted order -> caller order
__global__ void swb_scatter_kernel(const int32_t *__restrict__ sorted, const uint32_t *__restrict__ dst,
                                   uint32_t n, int32_t *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[dst[i]] = sorted[i];
}

// scores of the flagged tiles back to 0 before the int32 pass: its pipelined work items combine with atomicMax, and what
// the s16 pass left there may be larger than the true score (wrapped values)
__global__ void swb_clear_flagged_kernel(const SwbTile *__restrict__ tiles, uint32_t ntiles,
                                         const uint8_t *__restrict__ flags, int32_t *__restrict__ scores)
{
    const uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= ntiles || !flags[ti]) return;
    const SwbTile t = tiles[ti];
    for (uint32_t s = 0; s < t.npairs; ++s) {
        scores[2 * ((size_t)t.first_pair + s)] = 0;
        scores[2 * ((size_t)t.first_pair + s) + 1] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// The dynamic shared-memory limit of a kernel is per-device state shared by every engine (and host thread) of the
// process: raising it to what one launch needs and lowering it for the next would race between the engines of a group.
// It is set once per kernel and device to the device's opt-in maximum instead.
template <int K, class V, int NT, int MINB, bool SPLIT>
static cudaError_t allow_max_smem()
{
    static std::atomic<int> done[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && done[dev].load(std::memory_order_acquire)) return cudaSuccess;
    int optin = 0;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(swb_score_kernel<K, V, NT, MINB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  optin)) != cudaSuccess)
        return e;
    if (dev < 64) done[dev].store(1, std::memory_order_release);
    return cudaSuccess;
}

template <int K, class V, int NT, int MINB, bool SPLIT>
static cudaError_t launch_one(const SwbScoreParams &p, int grid, size_t smem, cudaStream_t st)
{
    cudaError_t e = allow_max_smem<K, V, NT, MINB, SPLIT>();
    if (e != cudaSuccess) return e;
    swb_score_kernel<K, V, NT, MINB, SPLIT><<<grid, NT, smem, st>>>(p);
    return cudaGetLastError();
}

template <int K, class V, int NT, int MINB, bool SPLIT>
static cudaError_t occ_one(size_t smem, int *blocks)
{
    cudaError_t e = allow_max_smem<K, V, NT, MINB, SPLIT>();
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, swb_score_kernel<K, V, NT, MINB, SPLIT>, NT, smem);
}

// op: 0 = launch, 1 = occupancy query
template <int K, class V, bool SPLIT>
static cudaError_t dispatch_cfg(int op, int block_cfg, const SwbScoreParams *p, int grid, size_t smem, cudaStream_t st,
                                int *blocks)
{
    // short strips (K = 8) have less work per column to hide latency with: cap their registers so that three blocks
    // (24 warps) fit on an SM instead of two
    constexpr int kMinBlocks = (K == 8 && V::is16) ? 3 : SWB_MINB_SMALL;
    if (block_cfg == SWB_BLOCK_SMALL)
        return op == 0 ? launch_one<K, V, SWB_NT_SMALL, kMinBlocks, SPLIT>(*p, grid, smem, st)
                       : occ_one<K, V, SWB_NT_SMALL, kMinBlocks, SPLIT>(smem, blocks);
    return op == 0 ? launch_one<K, V, SWB_NT_LARGE, 1, SPLIT>(*p, grid, smem, st)
                   : occ_one<K, V, SWB_NT_LARGE, 1, SPLIT>(smem, blocks);
}

static cudaError_t dispatch(int op, int K, int mode, bool split, int block_cfg, const SwbScoreParams *p, int grid,
                            size_t smem, cudaStream_t st, int *blocks)
{
    const bool i32 = mode == SWB_MODE_I32;
    if (mode == SWB_MODE_S16A || mode == SWB_MODE_I32A) {  // affine gaps: twice the row state, so shorter strips
        if (split) return cudaErrorInvalidValue;
        if (mode == SWB_MODE_S16A) {
            switch (K) {
            case 8: return dispatch_cfg<8, V16A, false>(op, block_cfg, p, grid, smem, st, blocks);
            case 16: return dispatch_cfg<16, V16A, false>(op, block_cfg, p, grid, smem, st, blocks);
            case 32: return dispatch_cfg<32, V16A, false>(op, block_cfg, p, grid, smem, st, blocks);
            }
        } else if (K == 8) {
            return dispatch_cfg<8, V32A, false>(op, block_cfg, p, grid, smem, st, blocks);
        }
        return cudaErrorInvalidValue;
    }
    const bool r16 = mode == SWB_MODE_R16;
    if (split) {  // pipelined passes: one warp per block, the profile rows of a pass staged per work item
        if (K == 8) {
            if (i32)
                return op == 0 ? launch_one<8, V32, 32, 20, true>(*p, grid, smem, st)
                               : occ_one<8, V32, 32, 20, true>(smem, blocks);
            if (r16)
                return op == 0 ? launch_one<8, V16R, 32, 24, true>(*p, grid, smem, st)
                               : occ_one<8, V16R, 32, 24, true>(smem, blocks);
            return op == 0 ? launch_one<8, V16, 32, 28, true>(*p, grid, smem, st)
                           : occ_one<8, V16, 32, 28, true>(smem, blocks);
        }
        if (K == 16 && !i32) {  // 16.5 KB of staged rows per work item: 13 blocks per SM
            if (r16)
                return op == 0 ? launch_one<16, V16R, 32, 13, true>(*p, grid, smem, st)
                               : occ_one<16, V16R, 32, 13, true>(smem, blocks);
            return op == 0 ? launch_one<16, V16, 32, 13, true>(*p, grid, smem, st)
                           : occ_one<16, V16, 32, 13, true>(smem, blocks);
        }
        if (K == 32 && !i32) {  // 33 KB of staged rows per work item: 6 blocks per SM
            if (r16)
                return op == 0 ? launch_one<32, V16R, 32, 6, true>(*p, grid, smem, st)
                               : occ_one<32, V16R, 32, 6, true>(smem, blocks);
            return op == 0 ? launch_one<32, V16, 32, 6, true>(*p, grid, smem, st)
                           : occ_one<32, V16, 32, 6, true>(smem, blocks);
        }
        return cudaErrorInvalidValue;
    }
    if (r16) {
        switch (K) {
        case 8: return dispatch_cfg<8, V16R, false>(op, block_cfg, p, grid, smem, st, blocks);
        case 16: return dispatch_cfg<16, V16R, false>(op, block_cfg, p, grid, smem, st, blocks);
        }
        return cudaErrorInvalidValue;
    }
    if (!i32) {
        switch (K) {
        case 8: return dispatch_cfg<8, V16, false>(op, block_cfg, p, grid, smem, st, blocks);
        case 16: return dispatch_cfg<16, V16, false>(op, block_cfg, p, grid, smem, st, blocks);
        case 32: return dispatch_cfg<32, V16, false>(op, block_cfg, p, grid, smem, st, blocks);
        }
    } else {
        switch (K) {
        case 8: return dispatch_cfg<8, V32, false>(op, block_cfg, p, grid, smem, st, blocks);
        case 16: return dispatch_cfg<16, V32, false>(op, block_cfg, p, grid, smem, st, blocks);
        }
    }
    return cudaErrorInvalidValue;
}

cudaError_t swb_launch_score(int K, int mode, bool split, int block_cfg, const SwbScoreParams &p, int grid, size_t smem,
                             cudaStream_t st)
{
    return dispatch(0, K, mode, split, block_cfg, &p, grid, smem, st, nullptr);
}

cudaError_t swb_score_occupancy(int K, int mode, bool split, int block_cfg, size_t smem, int *blocks)
{
    return dispatch(1, K, mode, split, block_cfg, nullptr, 0, smem, nullptr, blocks);
}

cudaError_t swb_launch_clear_flagged(const SwbTile *tiles, uint32_t ntiles, const uint8_t *flags, int32_t *scores,
                                     cudaStream_t st)
{
    if (!ntiles) return cudaSuccess;
    swb_clear_flagged_kernel<<<(ntiles + 255) / 256, 256, 0, st>>>(tiles, ntiles, flags, scores);
    return cudaGetLastError();
}

cudaError_t swb_launch_profile(const uint8_t *q, uint32_t qlen, const int8_t *mat, int bias, int8_t *prof,
                               uint32_t stride, uint32_t rows, cudaStream_t st)
{
    swb_profile_kernel<<<(rows + 255) / 256, 256, 0, st>>>(q, qlen, mat, bias, prof, stride, rows);
    return cudaGetLastError();
}

cudaError_t swb_launch_pack(const SwbTile *tiles, uint32_t ntiles, const uint8_t *raw, const uint64_t *seq_off,
                            const uint32_t *seq_len, uint32_t nseq, uint8_t *residues, cudaStream_t st)
{
    if (ntiles == 0) return cudaSuccess;
    const uint32_t grid = ntiles < 65535u * 16u ? ntiles : 65535u * 16u;
    swb_pack_kernel<<<grid, 256, 0, st>>>(tiles, ntiles, raw, seq_off, seq_len, nseq, residues);
    return cudaGetLastError();
}

cudaError_t swb_launch_scatter(const int32_t *sorted, const uint32_t *dst, uint32_t n, int32_t *out, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    swb_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(sorted, dst, n, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// The k best entries of one score vector, on the device: score descending, position ascending (the positions of a
// shard's vector are its database ids in ascending order, so this is "score desc, id asc"). What the reference does
// instead is hand all n scores back (SWSolver.cu:383-390); a scan of many queries then returns 8 k bytes per query and
// GPU instead of 4 n.
// One block per vector. All keys score << pbits | (n - 1 - position) are distinct, so the k best are the k largest
// keys: radix select, 11 bits per pass, on the significant bits only (scores are >= 0: the bits of the largest score plus
// the bits of n - 1, typically 3 passes), then one pass that collects the keys at or above the threshold and a bitonic
// sort of those k in shared memory.
#define SWB_TOPK_NT 512
#define SWB_TOPK_BINS 2048
__global__ void __launch_bounds__(SWB_TOPK_NT) swb_topk_kernel(const int32_t *__restrict__ scores, uint32_t n,
                                                                const uint32_t *__restrict__ ids, uint32_t k,
                                                                uint32_t *__restrict__ out_ids,
                                                                int32_t *__restrict__ out_scores)
{
    __shared__ uint32_t hist[SWB_TOPK_BINS];
    __shared__ unsigned long long sel[SWB_TOPK_MAX];
    __shared__ uint32_t s_red[SWB_TOPK_NT / 32];
    __shared__ uint32_t s_bin, s_above, s_count;
    const uint32_t tid = threadIdx.x;
    const uint32_t kk = k < n ? k : n;
    // largest score -> number of significant score bits
    uint32_t mx = 0;
    for (uint32_t i = tid; i < n; i += SWB_TOPK_NT) {
        const int32_t v = scores[i];
        mx = max(mx, (uint32_t)(v < 0 ? 0 : v));
    }
    for (int m = 16; m >= 1; m >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    if ((tid & 31u) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    mx = 0;
    for (uint32_t w = 0; w < SWB_TOPK_NT / 32; ++w) mx = max(mx, s_red[w]);
    const uint32_t pbits = n > 1 ? 32u - (uint32_t)__clz((int)(n - 1)) : 0u;
    const uint32_t sbits = mx ? 32u - (uint32_t)__clz((int)mx) : 0u;
    const uint32_t bits = pbits + sbits;  // <= 63
    const unsigned long long pmask = (1ull << pbits) - 1ull;
    auto key_of = [&](uint32_t i) {
        const int32_t v = scores[i];
        return ((unsigned long long)(uint32_t)(v < 0 ? 0 : v) << pbits) | (unsigned long long)(n - 1u - i);
    };
    unsigned long long prefix = 0;  // the digits chosen so far
    uint32_t remaining = kk;
    int shift = (int)bits;          // bits below the chosen prefix
    unsigned long long threshold = 0;
    while (kk > 0 && shift > 0) {
        const int width = shift >= 11 ? 11 : shift;
        const int lo = shift - width;
        for (uint32_t b = tid; b < SWB_TOPK_BINS; b += SWB_TOPK_NT) hist[b] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < n; i += SWB_TOPK_NT) {
            const unsigned long long key = key_of(i);
            if ((key >> shift) == prefix) atomicAdd(&hist[(uint32_t)(key >> lo) & ((1u << width) - 1u)], 1u);
        }
        __syncthreads();
        // the bin that holds the remaining-th largest key: suffix sums over the bins, 4 bins per thread
        uint32_t c[4], mine = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            c[j] = hist[SWB_TOPK_BINS - 1u - (tid * 4u + j)];  // bins in descending order
            mine += c[j];
        }
        uint32_t incl = mine;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31u) >= (uint32_t)d) incl += t;
        }
        if ((tid & 31u) == 31u) s_red[tid >> 5] = incl;
        __syncthreads();
        uint32_t before = incl - mine;  // keys in larger bins than this thread's four
        for (uint32_t w = 0; w < (tid >> 5); ++w) before += s_red[w];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (before < remaining && before + c[j] >= remaining) {
                s_bin = SWB_TOPK_BINS - 1u - (tid * 4u + j);
                s_above = before;
                s_count = c[j];
            }
            before += c[j];
        }
        __syncthreads();
        prefix = (prefix << width) | s_bin;
        remaining -= s_above;
        shift = lo;
        const uint32_t in_bin = s_count;
        __syncthreads();
        if (in_bin == remaining) break;  // every key with this prefix is among the k best
    }
    threshold = prefix << shift;
    // collect the kk keys at or above the threshold, then sort them
    if (tid == 0) s_count = 0;
    uint32_t p2 = 1;
    while (p2 < kk) p2 <<= 1;
    for (uint32_t j = tid; j < p2; j += SWB_TOPK_NT) sel[j] = 0ull;
    __syncthreads();
    if (kk > 0)
        for (uint32_t i = tid; i < n; i += SWB_TOPK_NT) {
            const unsigned long long key = key_of(i);
            if (key >= threshold) {
                const uint32_t at = atomicAdd(&s_count, 1u);
                if (at < SWB_TOPK_MAX) sel[at] = key + 1ull;  // + 1: a real key never equals the padding value 0
            }
        }
    __syncthreads();
    for (uint32_t size = 2; size <= p2; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t j = tid; j < p2; j += SWB_TOPK_NT) {
                const uint32_t partner = j ^ stride;
                if (partner > j) {
                    const bool desc = (j & size) == 0;
                    const unsigned long long a = sel[j], b = sel[partner];
                    if (desc ? a < b : a > b) {
                        sel[j] = b;
                        sel[partner] = a;
                    }
                }
            }
            __syncthreads();
        }
    for (uint32_t j = tid; j < k; j += SWB_TOPK_NT) {
        if (j < kk) {
            const unsigned long long key = sel[j] - 1ull;
            const uint32_t pos = n - 1u - (uint32_t)(key & pmask);
            out_ids[j] = ids ? ids[pos] : pos;
            out_scores[j] = (int32_t)(key >> pbits);
        } else {
            out_ids[j] = 0xffffffffu;
            out_scores[j] = -1;
        }
    }
}

cudaError_t swb_launch_topk(const int32_t *scores, uint32_t n, const uint32_t *ids, uint32_t k, uint32_t *out_ids,
                            int32_t *out_scores, cudaStream_t st)
{
    if (k == 0 || k > SWB_TOPK_MAX) return cudaErrorInvalidValue;
    swb_topk_kernel<<<1, SWB_TOPK_NT, 0, st>>>(scores, n, ids, k, out_ids, out_scores);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Alignments with traceback for a list of (query, subject) hits -- what the reference's cpu.cpp prints for two strings
// (cpu.cpp:39-103): fill with the update order LEFT, TOP, DIAG and strict '>' (cpu.cpp:47-64), keep the first
// row-major maximum (cpu.cpp:66-70), walk back until a cell with no direction, i.e. H == 0 (cpu.cpp:80-103).
// One block per hit, all hits of a list in one launch. Anti-diagonal wavefront; the three rolling H diagonals live in
// shared memory (in global scratch only for queries beyond ~18,000 rows); directions are packed 2 bits per cell,
// rows of ceil((n + 1) / 4) bytes. A row always belongs to the same thread (row i -> thread (i - 1) mod block), so
// the four cells of a direction byte are written by one thread on four consecutive diagonals: it collects them in a byte
// per row next to the diagonals and stores each finished byte once -- no atomics, no read-modify-write of global
// memory, no clearing pass.
// AFF: Gotoh's three states (the reference has no affine code, so the tie-breaks are this repo's
// definition, mirrored by the oracle's swo_align_affine): H takes its sources in cpu.cpp's order LEFT (E), TOP (F),
// DIAG with strict '>'; a gap state prefers OPEN over EXTEND on a tie; 4 bits per cell (source of H, "E extended",
// "F extended"), two more rolling diagonals each for E and F. With open == extend every choice coincides with the
// linear walk.
template <bool AFF>
__global__ void __launch_bounds__(SWB_ALIGN_NT) swb_align_batch_kernel(const SwbAlignJob *__restrict__ jobs,
                                                                       const uint8_t *__restrict__ qbuf,
                                                                       const uint8_t *__restrict__ raw,
                                                                       const int8_t *__restrict__ mat, int gap, int gap_ext,
                                                                       int32_t *hd_glob, uint8_t *dirbuf,
                                                                       int32_t *out_hdr, uint8_t *out_ops,
                                                                       uint32_t smem_ints)
{
    extern __shared__ int32_t s_hd[];
    __shared__ int8_t s_mat[SWB_ALPHA * SWB_ALPHA];
    __shared__ int s_best[SWB_ALIGN_NT];
    __shared__ uint32_t s_bi[SWB_ALIGN_NT], s_bj[SWB_ALIGN_NT];
    constexpr uint32_t BITS = AFF ? 4u : 2u;          // direction bits per cell
    constexpr uint32_t PERB = 8u / BITS;              // cells per direction byte
    constexpr uint32_t NDIAG = AFF ? 7u : 3u;         // rolling diagonals: 3 x H (+ 2 x E + 2 x F)
    constexpr int NEG = -(1 << 28);
    const SwbAlignJob jb = jobs[blockIdx.x];
    const uint32_t m = jb.m, n = jb.n;
    const uint8_t *q = qbuf + jb.q_off, *d = raw + jb.d_off;
    uint8_t *dir = dirbuf + jb.dir_off;
    const uint32_t Wb = swb_align_row_bytes(n, AFF);  // bytes per direction row (columns 0 .. n)
    const uint32_t tid = threadIdx.x;
    const size_t L = (size_t)m + 2;
    int32_t *hd = swb_align_hd_ints(m, AFF) <= smem_ints ? s_hd : hd_glob + jb.hd_off;
    int32_t *h0 = hd, *h1 = hd + L, *h2 = hd + 2 * L;
    int32_t *e1 = hd + 3 * L, *e2 = hd + 4 * L, *f1 = hd + 5 * L, *f2 = hd + 6 * L;  // AFF only: previous / current
    uint8_t *cur = reinterpret_cast<uint8_t *>(hd + NDIAG * L);  // per row: the direction byte being filled
    for (uint32_t i = tid; i < 3 * L; i += SWB_ALIGN_NT) hd[i] = 0;
    if (AFF)
        for (uint32_t i = tid; i < 4 * L; i += SWB_ALIGN_NT) e1[i] = NEG;
    for (uint32_t i = tid; i < SWB_ALPHA * SWB_ALPHA; i += SWB_ALIGN_NT) s_mat[i] = mat[i];
    int best = 0;
    uint32_t bi = 0, bj = 0;
    __syncthreads();
    // diagonal dd = i + j; h2 = current, h1 = dd-1, h0 = dd-2, all indexed by i
    for (uint32_t dd = 2; dd <= m + n; ++dd) {
        const uint32_t ilo = dd > n ? dd - n : 1u;
        const uint32_t ihi = dd - 1 < m ? dd - 1 : m;
        // cells outside [ilo, ihi] of the new diagonal are borders (H = 0, no gap state) for the next two diagonals
        if (tid == 0) {
            h2[ilo - 1] = 0;
            h2[ihi + 1] = 0;
            if (AFF) {
                e2[ilo - 1] = NEG; e2[ihi + 1] = NEG;
                f2[ilo - 1] = NEG; f2[ihi + 1] = NEG;
            }
        }
        // first row of this thread at or after ilo
        uint32_t i = ilo + (tid + SWB_ALIGN_NT - ((ilo - 1u) % SWB_ALIGN_NT)) % SWB_ALIGN_NT;
        for (; i <= ihi; i += SWB_ALIGN_NT) {
            const uint32_t j = dd - i;
            int h = 0;
            uint32_t t = 0;
            int left = h1[i] - gap;      // H(i, j-1) - open
            int up = h1[i - 1] - gap;    // H(i-1, j) - open
            if (AFF) {
                const int ee = e1[i] - gap_ext, fe = f1[i - 1] - gap_ext;
                if (ee > left) { left = ee; t |= 4u; }
                if (fe > up) { up = fe; t |= 8u; }
                e2[i] = left;
                f2[i] = up;
            }
            const int dg = h0[i - 1] + s_mat[(uint32_t)(q[i - 1] & 31u) * SWB_ALPHA + (d[j - 1] & 31u)];
            if (left > h) { h = left; t = (t & 12u) | 1u; }
            if (up > h) { h = up; t = (t & 12u) | 2u; }
            if (dg > h) { h = dg; t = (t & 12u) | 3u; }
            h2[i] = h;
            // the cells of a row that share a direction byte are collected next to the diagonals, stored when complete
            const uint32_t k = j & (PERB - 1u);
            const uint32_t b = ((k == 0u || j == 1u) ? 0u : (uint32_t)cur[i]) | (t << (BITS * k));
            cur[i] = (uint8_t)b;
            if (k == PERB - 1u || j == n) dir[(size_t)i * Wb + j / PERB] = (uint8_t)b;
            // first row-major maximum: larger value, else smaller i, else smaller j
            if (h > best || (h == best && h > 0 && (i < bi || (i == bi && j < bj)))) { best = h; bi = i; bj = j; }
        }
        int32_t *tmp = h0; h0 = h1; h1 = h2; h2 = tmp;
        if (AFF) {
            tmp = e1; e1 = e2; e2 = tmp;
            tmp = f1; f1 = f2; f2 = tmp;
        }
        __syncthreads();
    }
    s_best[tid] = best;
    s_bi[tid] = bi;
    s_bj[tid] = bj;
    __syncthreads();
    if (tid == 0) {
        for (uint32_t k = 1; k < SWB_ALIGN_NT; ++k) {
            const int v = s_best[k];
            if (v > best || (v == best && v > 0 && (s_bi[k] < bi || (s_bi[k] == bi && s_bj[k] < bj)))) {
                best = v; bi = s_bi[k]; bj = s_bj[k];
            }
        }
        uint8_t *ops = out_ops + jb.ops_off;
        uint32_t i = bi, j = bj, nops = 0, state = 0;  // state: 0 in H, 1 in E (gap in the query), 2 in F
        bool overflow = false;
        while (best > 0 && i > 0 && j > 0) {  // row 0 / column 0 are the H == 0 border (never written)
            const uint32_t t = ((uint32_t)dir[(size_t)i * Wb + j / PERB] >> (BITS * (j & (PERB - 1u)))) & (AFF ? 15u : 3u);
            uint32_t op;
            if (state == 0) {
                op = t & 3u;
                if (op == 0) break;
                if (AFF && op != 3u) { state = op; continue; }  // enter the gap state of this very cell
            } else {
                op = state;
                if (!(t & (state == 1u ? 4u : 8u))) state = 0;  // the gap was opened here: back to H after this column
            }
            if (nops < jb.cap) ops[nops] = (uint8_t)op; else overflow = true;
            ++nops;
            if (op == 1) --j;
            else if (op == 2) --i;
            else { --i; --j; }
        }
        int32_t *hdr = out_hdr + 5 * (size_t)blockIdx.x;
        hdr[0] = best;
        hdr[1] = (int32_t)bi;
        hdr[2] = (int32_t)bj;
        hdr[3] = (int32_t)nops;
        hdr[4] = overflow ? 1 : 0;
    }
}

cudaError_t swb_launch_align_batch(const SwbAlignJob *jobs, uint32_t njobs, const uint8_t *qbuf, const uint8_t *raw,
                                   const int8_t *mat, int gap_open, int gap_extend, bool affine, int32_t *hd_glob,
                                   uint8_t *dir, int32_t *out_hdr, uint8_t *out_ops, uint32_t smem_ints, cudaStream_t st)
{
    if (njobs == 0) return cudaSuccess;
    const size_t smem = (size_t)smem_ints * sizeof(int32_t);
    static std::atomic<int> done[64];  // the limit is per-device state shared by all engines: set once (see allow_max_smem)
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return ce;
    if (dev >= 64 || !done[dev].load(std::memory_order_acquire)) {
        ce = cudaFuncSetAttribute(swb_align_batch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)SWB_ALIGN_SMEM_MAX);
        if (ce == cudaSuccess)
            ce = cudaFuncSetAttribute(swb_align_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SWB_ALIGN_SMEM_MAX);
        if (ce != cudaSuccess) return ce;
        if (dev < 64) done[dev].store(1, std::memory_order_release);
    }
    if (affine)
        swb_align_batch_kernel<true><<<njobs, SWB_ALIGN_NT, smem, st>>>(jobs, qbuf, raw, mat, gap_open, gap_extend, hd_glob,
                                                                        dir, out_hdr, out_ops, smem_ints);
    else
        swb_align_batch_kernel<false><<<njobs, SWB_ALIGN_NT, smem, st>>>(jobs, qbuf, raw, mat, gap_open, gap_open, hd_glob,
                                                                         dir, out_hdr, out_ops, smem_ints);
    return cudaGetLastError();
}
