// Stable LSD radix sort of (u32 key, u32 value) pairs whose count lives on the device.
//
// Used twice per ACS iteration, both times because the reference's arithmetic depends on an
// ORDER that a parallel machine must reproduce exactly:
//   * ranking the colony by (steps, ant index)  — replaces the unstable std::sort at
//     ACSRank_3D.hpp:273-274 with the total order the oracle uses (stability = tie-break by
//     ant index, because the input is in ant order);
//   * ordering deposit records by pheromone slot while keeping rank order inside a slot, so
//     that the float additions of update_pheromone (:209-211) happen in the reference's order.
//
// Digits of up to kMaxDigitBits bits, chosen per sort so that the key width splits into the fewest
// equal passes (27-bit slot ids of a 256^3 grid: 3 x 9 instead of 4 x 8; 14-bit step counts: 2 x 7).
// One warp owns one tile of kTile consecutive items: pass A counts digits per tile, pass B scans
// each digit's row of the (digit, tile) matrix, pass C adds the digit bases, re-reads the tile in
// order and scatters with warp-match ranking, which keeps equal digits in input order (stable).
#include "wr_internal.cuh"

namespace wr {

constexpr int kTile = 512;           // items per warp: short tiles = many warps, the passes are latency-bound per warp
constexpr int kSortWarps = 4;        // warps per CTA
constexpr int kSortThreads = kSortWarps * 32;
constexpr int kMaxDigitBits = 10;
constexpr int kMaxDigits = 1 << kMaxDigitBits;

// digit of a key: bits [shift, shift+bits) — or, in PARTITION mode (span > 0), 0 for keys in [lo, lo+span) and 1 for the
// rest: one stable pass that moves a rank's slot slice to the front of the list (sharded colonies, acs.cu)
__device__ __forceinline__ unsigned digit_of(uint32_t key, int shift, uint32_t dmask, uint32_t lo, uint32_t span)
{
    return span ? ((key - lo) < span ? 0u : 1u) : ((key >> shift) & dmask);
}

__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const uint32_t* __restrict__ keys, const int* __restrict__ d_n, int shift, int bits,
                                                             uint32_t* __restrict__ hist, uint32_t lo, uint32_t span)
{
    __shared__ uint32_t cnt[kSortWarps][kMaxDigits];
    const int nd = 1 << bits;
    const uint32_t dmask = (uint32_t)nd - 1u;
    const int n = *d_n;
    const int ntiles = (n + kTile - 1) / kTile;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * kSortWarps + w;
    if (tile >= ntiles) return;
    for (int d = lane; d < nd; d += 32) cnt[w][d] = 0;
    __syncwarp();
    const int t_lo = tile * kTile, t_hi = min(t_lo + kTile, n);
#pragma unroll 4
    for (int i = t_lo + lane; i < t_hi; i += 32) atomicAdd(&cnt[w][digit_of(keys[i], shift, dmask, lo, span)], 1u);
    __syncwarp();
    for (int d = lane; d < nd; d += 32) hist[(size_t)d * ntiles + tile] = cnt[w][d];
}

// Row scans: CTA d turns hist[d][0..ntiles) into exclusive prefixes (in place) and publishes the row total.
// The digit bases (exclusive scan of the 256 totals) are folded into the scatter kernel, so the former
// single-CTA scan of the whole 256 x ntiles matrix (27 us per pass at 10^6 keys) becomes 256 short ones.
__global__ void __launch_bounds__(128) k_sort_scan_rows(uint32_t* __restrict__ hist, uint32_t* __restrict__ totals, const int* __restrict__ d_n)
{
    __shared__ uint32_t warp_sum[4];
    __shared__ uint32_t carry;
    const int n = *d_n;
    const int ntiles = (n + kTile - 1) / kTile;
    uint32_t* row = hist + (size_t)blockIdx.x * ntiles;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 128) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < ntiles ? row[i] : 0u;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) warp_sum[w] = incl;
        __syncthreads();
        uint32_t before = carry;
        for (int j = 0; j < w; j++) before += warp_sum[j];
        if (i < ntiles) row[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 127) carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                const int* __restrict__ d_n, int shift, int bits, const uint32_t* __restrict__ hist,
                                                                const uint32_t* __restrict__ totals, uint32_t part_lo, uint32_t part_span)
{
    __shared__ uint32_t base[kSortWarps][kMaxDigits];
    __shared__ uint32_t digit_base[kMaxDigits];
    __shared__ uint32_t wsum[kSortWarps];
    const int nd = 1 << bits;
    const uint32_t dmask = (uint32_t)nd - 1u;
    const int n = *d_n;
    const int ntiles = (n + kTile - 1) / kTile;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x * kSortWarps >= ntiles) return;   // whole CTA idle
    {   // exclusive scan of the digit totals: thread t owns digits [t*per, (t+1)*per)
        const int per = (nd + kSortThreads - 1) / kSortThreads;   // <= 8
        uint32_t loc[kMaxDigits / kSortThreads];
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < kMaxDigits / kSortThreads; j++) {
            const int d = threadIdx.x * per + j;
            loc[j] = (j < per && d < nd) ? totals[d] : 0u;
            sum += loc[j];
        }
        uint32_t incl = sum;
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int j = 0; j < w; j++) run += wsum[j];
#pragma unroll
        for (int j = 0; j < kMaxDigits / kSortThreads; j++) {
            const int d = threadIdx.x * per + j;
            if (j < per && d < nd) { digit_base[d] = run; run += loc[j]; }
        }
        __syncthreads();
    }
    const int tile = blockIdx.x * kSortWarps + w;
    if (tile >= ntiles) return;
    for (int d = lane; d < nd; d += 32) base[w][d] = digit_base[d] + hist[(size_t)d * ntiles + tile];
    __syncwarp();
    const int lo = tile * kTile, hi = min(lo + kTile, n);
    const unsigned lt = (1u << lane) - 1;
    // the chunk after the current one is already in flight while the current one is ranked and scattered
    uint32_t nk = 0, nv = 0;
    if (lo + lane < hi) { nk = keys_in[lo + lane]; nv = vals_in[lo + lane]; }
    for (int i0 = lo; i0 < hi; i0 += 32) {
        const int i = i0 + lane;
        const bool act = i < hi;
        const uint32_t k = nk, v = nv;
        if (i + 32 < hi) { nk = keys_in[i + 32]; nv = vals_in[i + 32]; }
        const unsigned d = act ? digit_of(k, shift, dmask, part_lo, part_span) : (unsigned)nd;  // inactive lanes form their own group
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (act) {
            const uint32_t pos = base[w][d] + __popc(peers & lt);
            keys_out[pos] = k; vals_out[pos] = v;
        }
        __syncwarp();
        if (act && (peers & lt) == 0) base[w][d] += __popc(peers);  // group leader advances the digit cursor
        __syncwarp();
    }
}

int sort_plan_create(SortPlan* p, size_t max_n, cudaStream_t s)
{
    if (max_n < 1) max_n = 1;
    p->max_n = max_n;
    p->max_tiles = (int)((max_n + kTile - 1) / kTile);
    WR_CUDA(dmalloc(&p->keys_a, max_n * sizeof(uint32_t), s));
    WR_CUDA(dmalloc(&p->keys_b, max_n * sizeof(uint32_t), s));
    WR_CUDA(dmalloc(&p->vals_a, max_n * sizeof(uint32_t), s));
    WR_CUDA(dmalloc(&p->vals_b, max_n * sizeof(uint32_t), s));
    WR_CUDA(dmalloc(&p->hist, ((size_t)kMaxDigits * p->max_tiles + kMaxDigits) * sizeof(uint32_t), s));   // + the digit totals
    return WR_OK;
}

void sort_plan_destroy(SortPlan* p, cudaStream_t s)
{
    pool_free(p->keys_a, s); pool_free(p->keys_b, s); pool_free(p->vals_a, s); pool_free(p->vals_b, s); pool_free(p->hist, s);
    *p = SortPlan();
}

int sort_pairs(SortPlan* p, const int* d_n, int key_bits, cudaStream_t s, bool* result_in_b, bool input_in_b)
{
    if (key_bits < 1) key_bits = 1;
    const int passes = (key_bits + kMaxDigitBits - 1) / kMaxDigitBits;
    const int bits = (key_bits + passes - 1) / passes;   // fewest passes, then the narrowest digit that covers the key
    const int blocks = (p->max_tiles + kSortWarps - 1) / kSortWarps;
    uint32_t *ki = p->keys_a, *vi = p->vals_a, *ko = p->keys_b, *vo = p->vals_b;
    if (input_in_b) { std::swap(ki, ko); std::swap(vi, vo); }
    uint32_t* totals = p->hist + (size_t)kMaxDigits * p->max_tiles;
    for (int pass = 0; pass < passes; pass++) {
        k_sort_hist<<<blocks, kSortThreads, 0, s>>>(ki, d_n, pass * bits, bits, p->hist, 0u, 0u);
        k_sort_scan_rows<<<1 << bits, 128, 0, s>>>(p->hist, totals, d_n);
        k_sort_scatter<<<blocks, kSortThreads, 0, s>>>(ki, vi, ko, vo, d_n, pass * bits, bits, p->hist, totals, 0u, 0u);
        std::swap(ki, ko); std::swap(vi, vo);
    }
    WR_CUDA(cudaGetLastError());
    *result_in_b = ((passes & 1) != 0) != input_in_b;
    return WR_OK;
}

__global__ void k_copy_count(const uint32_t* __restrict__ src, int* __restrict__ dst) { *dst = (int)*src; }

// Stable partition of keys_a/vals_a: pairs with key in [lo, lo+span) first (count -> *d_count_out), the rest behind them.
// Result in keys_b/vals_b.
int sort_partition(SortPlan* p, const int* d_n, uint32_t lo, uint32_t span, cudaStream_t s, int* d_count_out)
{
    const int blocks = (p->max_tiles + kSortWarps - 1) / kSortWarps;
    uint32_t* totals = p->hist + (size_t)kMaxDigits * p->max_tiles;
    k_sort_hist<<<blocks, kSortThreads, 0, s>>>(p->keys_a, d_n, 0, 1, p->hist, lo, span);
    k_sort_scan_rows<<<2, 128, 0, s>>>(p->hist, totals, d_n);
    k_sort_scatter<<<blocks, kSortThreads, 0, s>>>(p->keys_a, p->vals_a, p->keys_b, p->vals_b, d_n, 0, 1, p->hist, totals, lo, span);
    k_copy_count<<<1, 1, 0, s>>>(totals, d_count_out);
    WR_CUDA(cudaGetLastError());
    return WR_OK;
}

// force the lazy loading of this file's kernels (acs.cu: preload_iteration_kernels)
int sort_preload()
{
    cudaFuncAttributes at;
    WR_CUDA(cudaFuncGetAttributes(&at, k_sort_hist));
    WR_CUDA(cudaFuncGetAttributes(&at, k_sort_scan_rows));
    WR_CUDA(cudaFuncGetAttributes(&at, k_sort_scatter));
    WR_CUDA(cudaFuncGetAttributes(&at, k_copy_count));
    return WR_OK;
}

}  // namespace wr
