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
// 8-bit digits.  One warp owns one tile of kTile consecutive items: pass A counts digits per
// tile, pass B scans the digit-major (digit, tile) matrix, pass C re-reads the tile in order and
// scatters with warp-match ranking, which keeps equal digits in input order (stable).
#include "wr_internal.cuh"

namespace wr {

constexpr int kTile = 2048;          // items per warp
constexpr int kSortWarps = 4;        // warps per CTA
constexpr int kSortThreads = kSortWarps * 32;

__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const uint32_t* __restrict__ keys, const int* __restrict__ d_n, int shift,
                                                             uint32_t* __restrict__ hist)
{
    __shared__ uint32_t cnt[kSortWarps][256];
    const int n = *d_n;
    const int ntiles = (n + kTile - 1) / kTile;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * kSortWarps + w;
    for (int d = lane; d < 256; d += 32) cnt[w][d] = 0;
    __syncwarp();
    if (tile < ntiles) {
        const int lo = tile * kTile, hi = min(lo + kTile, n);
        for (int i = lo + lane; i < hi; i += 32) atomicAdd(&cnt[w][(keys[i] >> shift) & 255u], 1u);
        __syncwarp();
        for (int d = lane; d < 256; d += 32) hist[(size_t)d * ntiles + tile] = cnt[w][d];
    }
}

// exclusive scan of hist[0 .. 256*ntiles) in place (single CTA; the matrix is small: 256 x n/2048)
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t* __restrict__ hist, const int* __restrict__ d_n)
{
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    const int n = *d_n;
    const int ntiles = (n + kTile - 1) / kTile;
    const int total = 256 * ntiles;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < total; base += 1024 * 4) {
        // 4 consecutive items per thread
        int i0 = base + threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (i0 + j < total) ? hist[i0 + j] : 0u;
        uint32_t t = v[0] + v[1] + v[2] + v[3];
        uint32_t incl = t;
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) warp_sum[w] = incl;
        __syncthreads();
        if (w == 0) {
            uint32_t s = warp_sum[lane], si = s;
            for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += u; }
            warp_sum[lane] = si - s;  // exclusive
        }
        __syncthreads();
        uint32_t excl = carry + warp_sum[w] + (incl - t);
#pragma unroll
        for (int j = 0; j < 4; j++) { if (i0 + j < total) hist[i0 + j] = excl; excl += v[j]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                const int* __restrict__ d_n, int shift, const uint32_t* __restrict__ hist)
{
    __shared__ uint32_t base[kSortWarps][256];
    const int n = *d_n;
    const int ntiles = (n + kTile - 1) / kTile;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * kSortWarps + w;
    if (tile >= ntiles) return;
    for (int d = lane; d < 256; d += 32) base[w][d] = hist[(size_t)d * ntiles + tile];
    __syncwarp();
    const int lo = tile * kTile, hi = min(lo + kTile, n);
    const unsigned lt = (1u << lane) - 1;
    for (int i0 = lo; i0 < hi; i0 += 32) {
        const int i = i0 + lane;
        const bool act = i < hi;
        uint32_t k = 0, v = 0;
        if (act) { k = keys_in[i]; v = vals_in[i]; }
        const unsigned d = act ? ((k >> shift) & 255u) : 256u;  // inactive lanes form their own group
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

int sort_plan_create(SortPlan* p, size_t max_n)
{
    if (max_n < 1) max_n = 1;
    p->max_n = max_n;
    p->max_tiles = (int)((max_n + kTile - 1) / kTile);
    WR_CUDA(cudaMalloc(&p->keys_a, max_n * sizeof(uint32_t)));
    WR_CUDA(cudaMalloc(&p->keys_b, max_n * sizeof(uint32_t)));
    WR_CUDA(cudaMalloc(&p->vals_a, max_n * sizeof(uint32_t)));
    WR_CUDA(cudaMalloc(&p->vals_b, max_n * sizeof(uint32_t)));
    WR_CUDA(cudaMalloc(&p->hist, (size_t)256 * p->max_tiles * sizeof(uint32_t)));
    return WR_OK;
}

void sort_plan_destroy(SortPlan* p)
{
    cudaFree(p->keys_a); cudaFree(p->keys_b); cudaFree(p->vals_a); cudaFree(p->vals_b); cudaFree(p->hist);
    *p = SortPlan();
}

int sort_pairs(SortPlan* p, const int* d_n, int key_bits, cudaStream_t s, bool* result_in_b)
{
    const int passes = (key_bits + 7) / 8;
    const int blocks = (p->max_tiles + kSortWarps - 1) / kSortWarps;
    uint32_t *ki = p->keys_a, *vi = p->vals_a, *ko = p->keys_b, *vo = p->vals_b;
    for (int pass = 0; pass < passes; pass++) {
        k_sort_hist<<<blocks, kSortThreads, 0, s>>>(ki, d_n, pass * 8, p->hist);
        k_sort_scan<<<1, 1024, 0, s>>>(p->hist, d_n);
        k_sort_scatter<<<blocks, kSortThreads, 0, s>>>(ki, vi, ko, vo, d_n, pass * 8, p->hist);
        std::swap(ki, ko); std::swap(vi, vo);
    }
    WR_CUDA(cudaGetLastError());
    *result_in_b = (passes & 1) != 0;
    return WR_OK;
}

}  // namespace wr
