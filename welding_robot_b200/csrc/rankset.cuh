// Rank-set deposits (WR_UPDATE_RANKSET): update_pheromone (core/ACSRank_3D.hpp:198-215) without sorting records.
//
// The reference adds, per directed slot, the deposits of the ants that crossed it in RANK order (the sorted colony is
// walked front to back, :275-280, and a self-avoiding ant crosses a slot at most once):
//     tau[s] = ((tau[s]*rho + d(r1, s)) + d(r2, s)) + ...      r1 < r2 < ... the ranks of the ants that crossed s
//     d(r, s) = (lambda - order_r)*Q/L_r + float(s's edge lies on the best path)*lambda*Q/L_best            (:209-211)
// so the value depends on the rank and on ONE bit of the slot.  What has to be known per slot is therefore the SET of
// ranks that crossed it — a bitmask over the <= 0.2*colony eligible ranks — and a set is built with atomicOr in any
// order: no (slot, rank) sort, no record list.  Then one thread (or, for a slot most of the colony crossed, one warp)
// walks the set bits in ascending order and performs the same additions, in the same order, as the reference.
//
//   k_rankset_gen     one CTA per eligible rank: value pair of the rank -> vtab; every step of its trail: find or claim
//                     the slot's entry in an open-addressed table (atomicCAS on the key), set bit `rank` in its row
//   k_evaporate       (acs_kernels.cuh) the dense tau *= rho pass, :268-272 — the one HBM stream of the update
//   k_rankset_apply   entries claimed this iteration (a compact list): ordered additions, then the row words that were
//                     used and the entry are cleared for the next iteration; the slot is left in `touched` for the next
//                     walk's L2 warm-up
// Table entry (8 bytes, one sector access): low word = slot + 1 (0 = free); high word = summary: bit w (< 31) <=> word w
// of the entry's row is non-zero, bit 31 = the slot's edge lies on the best path.  A wandering colony gives ~10^6 entries
// with one rank each: the summary lets gen, apply and the clean-up touch one row word instead of all of them.
// Row: nwords = ceil(w_max / 32) <= 31 words of rank bits (colonies of up to 4955 ants; larger ones take WR_UPDATE_FUSED).
#pragma once
#include "acs_kernels.cuh"

namespace wr {

constexpr int kRankSetMaxWords = 31;

struct RankSet {
    unsigned long long* ent;   // [T]  key | summary << 32
    uint32_t* rows;            // [T][nwords]
    uint32_t* list;            // [list_cap] entries claimed this iteration
    uint32_t* touched;         // [list_cap] their slots (written by apply, read by the next warm-up)
    uint32_t* count;           // [0] entries in `list` this iteration, [1] entries in `touched` (previous iteration), [2] = 0
    float* vtab;               // [w_max][2]  d(r, s) without / with the elitist term
    uint32_t tmask;            // T - 1
    int shift;                 // 32 - log2(T): entry = (slot * golden) >> shift
    int nwords;
};

__device__ __forceinline__ uint32_t rankset_hash(uint32_t slot) { return slot * 2654435761u; }

__global__ void __launch_bounds__(128) k_rankset_gen(const IterState* st, const uint32_t* __restrict__ rank_keys, const uint32_t* __restrict__ rank_vals,
                                                      const uint32_t* __restrict__ path_ids, const uint8_t* __restrict__ path_dirs, int cap, int goal,
                                                      const float* __restrict__ Ltab, const uint32_t* __restrict__ onbest, RankSet rs, int K,
                                                      const int* __restrict__ steps26)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int r = blockIdx.x;
    if (r >= st->n_eligible || !st->use_rankset) return;
    const int ant = (int)rank_vals[r];
    const int steps = steps26 ? steps26[ant] : (int)rank_keys[r];
    if (threadIdx.x == 0) {   // the same expressions, in the same order, as k_deposit_gen
        const float L_ant = steps26 ? __uint_as_float(rank_keys[r]) : Ltab[steps];
        const float lambda = st->lambda, Q = st->Q;
        const float base = __fdiv_rn(__fmul_rn(__fsub_rn(lambda, (float)(r + 1)), Q), L_ant);
        const float elite = __fdiv_rn(__fmul_rn(__fmul_rn(1.0f, lambda), Q), st->best_L);
        rs.vtab[2 * r] = __fadd_rn(base, 0.0f);
        rs.vtab[2 * r + 1] = __fadd_rn(base, elite);
    }
    const uint32_t* pid = path_ids + (size_t)ant * cap;
    const uint8_t* pdir = path_dirs + (size_t)ant * cap;
    const uint32_t bit = 1u << (r & 31);
    const int word = r >> 5;
    const uint32_t wbit = 1u << word;
    const int lane = threadIdx.x & 31;
    uint32_t* ent32 = reinterpret_cast<uint32_t*>(rs.ent);   // [2h] key, [2h+1] summary
    for (int i0 = 0; i0 < steps; i0 += blockDim.x) {         // CTA-uniform trips: the list append below is warp-aggregated
        const int i = i0 + threadIdx.x;
        const bool valid = i < steps;
        uint32_t h = 0;
        bool claimed = false;
        uint32_t sum_bits = wbit;
        if (valid) {
            const uint32_t node = pid[i];
            const uint32_t slot = node * (uint32_t)K + pdir[i];
            const uint32_t want = slot + 1u;
            h = rankset_hash(slot) >> rs.shift;
            while (true) {
                uint32_t k = __ldcg(ent32 + 2 * (size_t)h);
                if (k == 0u) {
                    k = atomicCAS(ent32 + 2 * (size_t)h, 0u, want);
                    if (k == 0u) { claimed = true; break; }
                }
                if (k == want) break;
                h = (h + 1) & rs.tmask;
            }
            if (claimed) {   // the slot's on-best bit travels in the summary (every ant on this slot would compute the same bit)
                const uint32_t next = (i + 1 < steps) ? pid[i + 1] : (uint32_t)goal;
                const bool onb = ((onbest[node >> 5] >> (node & 31)) & 1u) && ((onbest[next >> 5] >> (next & 31)) & 1u);
                sum_bits |= onb ? 0x80000000u : 0u;
            }
        }
        const unsigned cm = __ballot_sync(FULL, claimed);
        if (cm) {   // one counter update per warp: a wandering colony claims ~10^6 entries per iteration
            uint32_t at = 0;
            if (lane == __ffs(cm) - 1) at = atomicAdd(rs.count, (uint32_t)__popc(cm));
            at = __shfl_sync(FULL, at, __ffs(cm) - 1);
            if (claimed) rs.list[at + __popc(cm & ((1u << lane) - 1u))] = h;
        }
        if (valid) {
            if (claimed || (__ldcg(ent32 + 2 * (size_t)h + 1) & sum_bits) != sum_bits) atomicOr(ent32 + 2 * (size_t)h + 1, sum_bits);
            atomicOr(rs.rows + (size_t)h * rs.nwords + word, bit);
        }
    }
}

// Lane per entry; an entry with more than two non-empty row words is handed to the whole warp: the 32 values of a word are
// fetched by the lanes at once and the dependent FADD chain runs on values exchanged by shuffle (a converged colony puts
// ~0.2*colony ranks on every slot of the best path).  Absent ranks contribute +0, the identity of the chain.
__global__ void __launch_bounds__(256) k_rankset_apply(const IterState* st, float* tau, RankSet rs, float rho, uint8_t* dirty)
{
    if (!st->use_rankset) return;
    const float base_new = __fmul_rn(st->base, rho);   // clean-tile field: a slot that still holds the sentinel is worth this after the evaporation
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t n = rs.count[0];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nwords = rs.nwords;
    // Consecutive list entries go to consecutive WARPS (entry = base + lane*nwarps + warp): the first ant to run claims
    // the whole best path in one stretch of the list, and those are exactly the heavy entries — one per warp, not 32.
    for (uint32_t base = 0; base < n; base += nwarps * 32u) {   // warp-uniform trips
        const uint32_t e = base + (uint32_t)lane * nwarps + warp;
        const bool has = e < n;
        uint32_t h = 0, slot = 0, flag = 0, wm = 0;
        uint32_t* row = rs.rows;
        if (has) {
            h = rs.list[e];
            const unsigned long long ent = rs.ent[h];
            slot = (uint32_t)ent - 1u;
            wm = (uint32_t)(ent >> 32) & 0x7FFFFFFFu;
            flag = (uint32_t)(ent >> 63);
            row = rs.rows + (size_t)h * nwords;
        }
        const bool heavy = has && __popc(wm) > 2;
        if (has && !heavy) {
            float x = tau_or_base(tau[slot], base_new);
            for (uint32_t ws = wm; ws; ws &= ws - 1) {
                const int w = __ffs(ws) - 1;
                uint32_t m = row[w];
                row[w] = 0u;
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    x = __fadd_rn(x, rs.vtab[2 * (w * 32 + b) + flag]);
                }
            }
            tau[slot] = x;
            dirty[slot / (uint32_t)kUpdTile] = 1;
        }
        unsigned hm = __ballot_sync(FULL, heavy);
        while (hm) {
            const int src = __ffs(hm) - 1;
            hm &= hm - 1;
            const uint32_t hs = __shfl_sync(FULL, h, src);
            const uint32_t ss = __shfl_sync(FULL, slot, src);
            const uint32_t fs = __shfl_sync(FULL, flag, src);
            uint32_t* rrow = rs.rows + (size_t)hs * nwords;
            float x = tau_or_base(tau[ss], base_new);
            const uint32_t mine = lane < nwords ? rrow[lane] : 0u;   // the row: one coalesced load, lane j holds word j (nwords <= 31)
            if (lane < nwords) rrow[lane] = 0u;
            for (int j0 = 0; j0 < nwords; j0 += 4) {   // four words per round: their value loads are in flight together
                uint32_t m[4];
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    m[j] = __shfl_sync(FULL, mine, (j0 + j) & 31);   // words >= nwords are 0
                    v[j] = ((m[j] >> lane) & 1u) ? rs.vtab[2 * ((j0 + j) * 32 + lane) + fs] : 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (m[j] == 0u) continue;   // warp-uniform
#pragma unroll
                    for (int i = 0; i < 32; i++) x = __fadd_rn(x, __shfl_sync(FULL, v[j], i));
                }
            }
            if (lane == 0) { tau[ss] = x; dirty[ss / (uint32_t)kUpdTile] = 1; }
        }
        __syncwarp();
        if (has) {   // leave the table empty for the next iteration; remember the slot for the walk's L2 warm-up
            rs.ent[h] = 0ull;
            rs.touched[e] = slot;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) rs.count[1] = n;
}

// L2 warm-up for the next walk from the slots that just received deposits (cf. k_path_warm, which reads the sorted records).
__global__ void __launch_bounds__(256) k_rankset_warm(const uint32_t* __restrict__ touched, const uint32_t* __restrict__ count, const float* tau, const float* heur,
                                                       IterState* st)
{
    if (!st->use_rankset) return;   // runs before k_iter_begin: the flag still describes the iteration that just ended
    const uint32_t n = count[1];
    uint32_t acc = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t node = touched[i] / 6u;
        const uint32_t* pt = reinterpret_cast<const uint32_t*>(tau) + (size_t)node * 6;
        const uint32_t* ph = reinterpret_cast<const uint32_t*>(heur) + (size_t)node * 6;
        acc ^= __ldcg(pt) ^ __ldcg(pt + 5) ^ __ldcg(ph) ^ __ldcg(ph + 5);
    }
    if (acc == 0x9E3779B9u && n == 0xFFFFFFFFu) st->cnt[8] = 0;   // keeps the loads alive; never true
}

}  // namespace wr
