/* TEST INFRASTRUCTURE (oracle/) — not part of the product.
 *
 * Philox4x32-10 counter-based generator (Salmon et al., "Parallel random numbers:
 * as easy as 1, 2, 3", SC'11; Random123 reference constants).  It replaces the
 * reference's glibc rand() (ACSRank_3D.hpp:169, ACS_GTSP.hpp:126) so that a draw
 * is a pure function of (seed; iteration, ant, step) and the GPU can reproduce it.
 *
 * The reference's conversion is kept so the inclusive-1.0 edge case survives:
 *   3-D : (float)rand()  / (float)RAND_MAX    (ACSRank_3D.hpp:169)
 *   GTSP: (double)rand() / (double)RAND_MAX   (ACS_GTSP.hpp:126)
 * with rand() := philox(...)[0] >> 1  (a 31-bit value, RAND_MAX = 2^31-1).
 */
#ifndef WR_ORACLE_PHILOX_H
#define WR_ORACLE_PHILOX_H
#include <stdint.h>

#define WR_PHILOX_M0 0xD2511F53u
#define WR_PHILOX_M1 0xCD9E8D57u
#define WR_PHILOX_W0 0x9E3779B9u
#define WR_PHILOX_W1 0xBB67AE85u

static inline void wr_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)WR_PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)WR_PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += WR_PHILOX_W0; k1 += WR_PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream tags (counter word 3) so that the 3-D search, the seam-ordering colonies and
 * the synthetic-input generators never share a counter. */
#define WR_STREAM_ACS3D 0x3D3D0000u
#define WR_STREAM_GTSP  0x65700000u
#define WR_STREAM_SEQ   0x5E900000u /* sequential n-th-call stream (reference pinning) */
#define WR_STREAM_SYNTH 0x51170000u

/* rand() replacement: 31-bit draw for (seed; a, b, c) on a stream. */
static inline uint32_t wr_rand31(uint64_t seed, uint32_t a, uint32_t b, uint32_t c, uint32_t stream)
{
    uint32_t ctr[4] = {a, b, c, stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    wr_philox4x32_10(ctr, key, out);
    return out[0] >> 1;
}

/* The 3-D search's keyed stream: one Philox block serves 4 consecutive steps of an ant —
 * draw(search, iteration, ant, step) = word (step & 3) of
 *     Philox(counter = (iteration, ant, (step >> 2) | (search >> 16) << 16, stream + (search & 0xFFFF))) >> 1.
 * `search` counts the computeSolution calls on one ACS_Rank object (0 for the first): the reference draws all its
 * searches from ONE continuous rand() stream (ACSRank_3D.hpp:169, seeded once at :327), so successive searches are
 * statistically independent; keying the counter by the search index keeps that property while a draw stays a pure
 * function of (seed; search, iteration, ant, step).  step >> 2 < 2^14 (step cap 65532), so the halves do not overlap. */
static inline uint32_t wr_rand31_step(uint64_t seed, uint32_t search, uint32_t iteration, uint32_t ant, uint32_t step, uint32_t stream)
{
    uint32_t ctr[4] = {iteration, ant, (step >> 2) | ((search >> 16) << 16), stream + (search & 0xFFFFu)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    wr_philox4x32_10(ctr, key, out);
    return out[step & 3] >> 1;
}

#endif
