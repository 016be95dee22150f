// TEST INFRASTRUCTURE (oracle/) — not part of the product.  See wr_oracle.h for the rules
// on who may load this and for how it is pinned to the unmodified reference.
//
// A CPU restatement of the reference's hot path, written from the reference's behaviour
// (file:line cited per function, paths relative to /root/reference).  It keeps the
// reference's arithmetic exactly — x86-64 SSE2 scalar float, no FMA (-ffp-contract=off),
// same operation order, same float/double promotions — and replaces only what the
// reference leaves undefined or unreproducible:
//   (1) rand() -> Philox4x32-10, keyed (seed; iteration, ant, step), or the n-th-call
//       sequential stream used to pin this file against the reference itself;
//   (2) the J.back()-on-empty UB at ACSRank_3D.hpp:174 -> "ant dies" (counted);
//   (3) unstable std::sort at :273 -> total order (L, ant index); WRO_SORT_STD keeps
//       std::sort with the reference's comparator for pinning runs;
//   (4) optional, flag-controlled deviations mirrored by the GPU: fixed colony size,
//       step cap, K = 26 neighbourhood.
// Containers differ (flat arrays instead of node objects / std::set) — results do not.
#include "wr_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <limits>
#include <vector>

#include "philox.h"

#define WRO_INF_FLOAT (1.0 / 0.0) /* ACSRank_3D.hpp:22 (a double, as there) */
#define wro_abs(x) ((x) > 0 ? (x) : -(x)) /* model_grid_map.hpp:23 */

// ------------------------------------------------------------------------------------------
// STL: read_STL.hpp:131-156 (binary branch; 80-byte header, u32 count, 50 B per triangle)
// ------------------------------------------------------------------------------------------
extern "C" int wro_stl_parse(const uint8_t* buf, size_t len, float* t12, int cap)
{
    if (len < 84) return -1;
    if (buf[79] != '\0') return -2; /* ASCII branch (read_STL.hpp:65-67) is not supported */
    uint32_t n;
    memcpy(&n, buf + 80, 4);
    if (84 + (size_t)n * 50 > len) return -3;
    const uint8_t* p = buf + 84;
    for (uint32_t i = 0; i < n && (int)i < cap; i++, p += 50) memcpy(t12 + 12 * (size_t)i, p, 48); /* +2 attr bytes skipped */
    return (int)n;
}

// ------------------------------------------------------------------------------------------
// Grid: model_grid_map.hpp:151-273
// ------------------------------------------------------------------------------------------
struct wro_grid {
    int rx, ry, rz, wall;
    float precision;
    std::vector<float> xs, ys, zs;  // separable node coordinates (:204-211)
    std::vector<uint8_t> isfree;    // z,y,x order, id = z*ry*rx + y*rx + x (:214)
    float gmin[3], gmax[3];         // global bbox (:165-181)
    float lmin[3], lmax[3];         // last triangle's box +-precision (what :279 really writes)
    uint64_t tests;
};

static float axis_coord(int i, int range, int wall, float mn, float mx, float precision)
{   /* model_grid_map.hpp:204-205 (same expression for y, z) */
    return i < wall ? mn - (wall - i) * precision : (i >= (range - wall) ? mx + (i - range + wall) * precision : mn + (i - wall) * precision);
}

extern "C" wro_grid* wro_grid_from_triangles(const float* t12, int ntri, float precision, int wall, int mode)
{
    if (ntri <= 0) return nullptr;
    wro_grid* g = new wro_grid();
    g->precision = precision; g->wall = wall; g->tests = 0;
    /* :165-181 global bbox, seeded with vertex 0 of triangle 0 */
    float min_x = t12[3], min_y = t12[4], min_z = t12[5];
    float max_x = min_x, max_y = min_y, max_z = min_z;
    for (int t = 0; t < ntri; t++)
        for (int i = 0; i < 3; i++) {
            const float* v = t12 + 12 * (size_t)t + 3 + 3 * i;
            max_x = v[0] > max_x ? v[0] : max_x; max_y = v[1] > max_y ? v[1] : max_y; max_z = v[2] > max_z ? v[2] : max_z;
            min_x = v[0] < min_x ? v[0] : min_x; min_y = v[1] < min_y ? v[1] : min_y; min_z = v[2] < min_z ? v[2] : min_z;
        }
    g->gmin[0] = min_x; g->gmin[1] = min_y; g->gmin[2] = min_z;
    g->gmax[0] = max_x; g->gmax[1] = max_y; g->gmax[2] = max_z;
    /* :198-200 */
    g->rx = (int)((max_x - min_x) / precision) + 1 + 2 * wall;
    g->ry = (int)((max_y - min_y) / precision) + 1 + 2 * wall;
    g->rz = (int)((max_z - min_z) / precision) + 1 + 2 * wall;
    const int rx = g->rx, ry = g->ry, rz = g->rz;
    g->xs.resize(rx); g->ys.resize(ry); g->zs.resize(rz);
    for (int x = 0; x < rx; x++) g->xs[x] = axis_coord(x, rx, wall, min_x, max_x, precision);
    for (int y = 0; y < ry; y++) g->ys[y] = axis_coord(y, ry, wall, min_y, max_y, precision);
    for (int z = 0; z < rz; z++) g->zs[z] = axis_coord(z, rz, wall, min_z, max_z, precision);
    g->isfree.assign((size_t)rx * ry * rz, 1);

    std::vector<int> cx, cy, cz;
    for (int t = 0; t < ntri; t++) {
        const float* n = t12 + 12 * (size_t)t;
        const float* v0 = n + 3;
        /* :224-226 */
        float D = -(v0[0] * n[0] + v0[1] * n[1] + v0[2] * n[2]);
        /* :228-248 */
        min_x = v0[0]; min_y = v0[1]; min_z = v0[2];
        max_x = min_x; max_y = min_y; max_z = min_z;
        for (int i = 0; i < 3; i++) {
            const float* v = n + 3 + 3 * i;
            max_x = v[0] > max_x ? v[0] : max_x; max_y = v[1] > max_y ? v[1] : max_y; max_z = v[2] > max_z ? v[2] : max_z;
            min_x = v[0] < min_x ? v[0] : min_x; min_y = v[1] < min_y ? v[1] : min_y; min_z = v[2] < min_z ? v[2] : min_z;
        }
        min_x -= precision; min_y -= precision; min_z -= precision;
        max_x += precision; max_y += precision; max_z += precision;
        if (mode == WRO_VOX_BRUTE) {
            /* :251-268, every node */
            size_t id = 0;
            for (int z = 0; z < rz; z++)
                for (int y = 0; y < ry; y++)
                    for (int x = 0; x < rx; x++, id++) {
                        float px = g->xs[x], py = g->ys[y], pz = g->zs[z];
                        float distance = px * n[0] + py * n[1] + pz * n[2] + D;
                        g->tests++;
                        if (wro_abs(distance) < 1.2 * precision) {
                            if (min_x <= px && px <= max_x && min_y <= py && py <= max_y && min_z <= pz && pz <= max_z) g->isfree[id] = 0;
                        }
                    }
        } else {
            /* Same predicate, restricted to the nodes that can pass the box test of :258-260.
             * The box test is separable, so the surviving set is a product of per-axis index
             * lists (lists, not ranges: coordinates are not monotone across the max-side seam). */
            cx.clear(); cy.clear(); cz.clear();
            for (int x = 0; x < rx; x++) if (min_x <= g->xs[x] && g->xs[x] <= max_x) cx.push_back(x);
            for (int y = 0; y < ry; y++) if (min_y <= g->ys[y] && g->ys[y] <= max_y) cy.push_back(y);
            for (int z = 0; z < rz; z++) if (min_z <= g->zs[z] && g->zs[z] <= max_z) cz.push_back(z);
            for (int z : cz)
                for (int y : cy)
                    for (int x : cx) {
                        float px = g->xs[x], py = g->ys[y], pz = g->zs[z];
                        float distance = px * n[0] + py * n[1] + pz * n[2] + D;
                        g->tests++;
                        if (wro_abs(distance) < 1.2 * precision) g->isfree[((size_t)z * ry + y) * rx + x] = 0;
                    }
        }
    }
    g->lmin[0] = min_x; g->lmin[1] = min_y; g->lmin[2] = min_z;
    g->lmax[0] = max_x; g->lmax[1] = max_y; g->lmax[2] = max_z;
    return g;
}

extern "C" wro_grid* wro_grid_from_occupancy(const uint8_t* isfree, int rx, int ry, int rz, const float* xs,
                                             const float* ys, const float* zs, float precision)
{
    wro_grid* g = new wro_grid();
    g->rx = rx; g->ry = ry; g->rz = rz; g->wall = 0; g->precision = precision; g->tests = 0;
    g->xs.assign(xs, xs + rx); g->ys.assign(ys, ys + ry); g->zs.assign(zs, zs + rz);
    g->isfree.assign(isfree, isfree + (size_t)rx * ry * rz);
    for (int k = 0; k < 3; k++) g->gmin[k] = g->gmax[k] = g->lmin[k] = g->lmax[k] = 0;
    return g;
}
extern "C" void wro_grid_destroy(wro_grid* g) { delete g; }
extern "C" void wro_grid_dims(const wro_grid* g, int d[3]) { d[0] = g->rx; d[1] = g->ry; d[2] = g->rz; }
extern "C" float wro_grid_precision(const wro_grid* g) { return g->precision; }
extern "C" void wro_grid_isfree(const wro_grid* g, uint8_t* out) { memcpy(out, g->isfree.data(), g->isfree.size()); }
extern "C" void wro_grid_coords(const wro_grid* g, float* xs, float* ys, float* zs)
{
    memcpy(xs, g->xs.data(), 4 * g->xs.size()); memcpy(ys, g->ys.data(), 4 * g->ys.size()); memcpy(zs, g->zs.data(), 4 * g->zs.size());
}
extern "C" uint64_t wro_grid_tests(const wro_grid* g) { return g->tests; }

/* model_grid_map.hpp:275-294 */
extern "C" int wro_grid_write_file(const wro_grid* g, const char* path, int compat)
{
    FILE* fp = fopen(path, "w");
    if (!fp) return -1;
    const float* mn = compat ? g->lmin : g->gmin;
    const float* mx = compat ? g->lmax : g->gmax;
    fprintf(fp, "%d %d %d %d %f %d\n", (int)g->isfree.size(), g->rx, g->ry, g->rz, g->precision, g->wall);
    fprintf(fp, "%f %f %f %f %f %f\n", mn[0], mn[1], mn[2], mx[0], mx[1], mx[2]);
    size_t id = 0;
    for (int i = 0; i < g->rz; i++)
        for (int j = 0; j < g->ry; j++) {
            for (int k = 0; k < g->rx; k++) fprintf(fp, "%d ", (int)g->isfree[id++]);
            fprintf(fp, "\n");
        }
    fclose(fp);
    return 0;
}
/* model_grid_map.hpp:300-356 (coordinates are rebuilt from whatever box the header holds) */
extern "C" wro_grid* wro_grid_read_file(const char* path)
{
    FILE* fp = fopen(path, "r");
    if (!fp) return nullptr;
    wro_grid* g = new wro_grid();
    int map_size = 0;
    float mn[3], mx[3];
    if (fscanf(fp, "%d %d %d %d %f %d", &map_size, &g->rx, &g->ry, &g->rz, &g->precision, &g->wall) != 6 ||
        fscanf(fp, "%f %f %f %f %f %f", &mn[0], &mn[1], &mn[2], &mx[0], &mx[1], &mx[2]) != 6) { fclose(fp); delete g; return nullptr; }
    g->tests = 0;
    for (int k = 0; k < 3; k++) { g->gmin[k] = g->lmin[k] = mn[k]; g->gmax[k] = g->lmax[k] = mx[k]; }
    g->xs.resize(g->rx); g->ys.resize(g->ry); g->zs.resize(g->rz);
    for (int x = 0; x < g->rx; x++) g->xs[x] = axis_coord(x, g->rx, g->wall, mn[0], mx[0], g->precision);
    for (int y = 0; y < g->ry; y++) g->ys[y] = axis_coord(y, g->ry, g->wall, mn[1], mx[1], g->precision);
    for (int z = 0; z < g->rz; z++) g->zs[z] = axis_coord(z, g->rz, g->wall, mn[2], mx[2], g->precision);
    size_t n = (size_t)g->rx * g->ry * g->rz;
    g->isfree.resize(n);
    for (size_t i = 0; i < n; i++) { int v = 1; if (fscanf(fp, "%d", &v) != 1) v = 1; g->isfree[i] = v != 0; }
    fclose(fp);
    return g;
}

// ------------------------------------------------------------------------------------------
// Rank-based 3-D ACS: ACSRank_3D.hpp
// ------------------------------------------------------------------------------------------
/* :48-60 */
template <class T> static T wro_power(T x, int y)
{
    T ans = 1;
    while (y) { if (y & 1) ans *= x; x *= x; y >>= 1; }
    return ans;
}

struct wro_ant {
    std::vector<int32_t> ids;   // path (node ids), ids[0] = start
    std::vector<uint8_t> dirs;  // chosen slot per step
    float L;
    int order;                  // 1-based rank after the sort (0 = not ranked yet)
};

struct wro_acs {
    const wro_grid* g;
    wro_acs_params p;
    int rx, ry, rz, K;
    size_t N;
    std::vector<float> tau;       // N*K, node-major, slot order of :355-359
    int dx[26], dy[26], dz[26];
    float dist[26];
    int64_t start, goal;
    float predict;
    int iter;                     // iterations since begin()
    uint64_t seq_calls;           // SEQUENTIAL stream position
    uint32_t search, next_search; // keyed stream: index of the current search (begin() calls so far - 1) / of the next one
    wro_ant best;
    std::vector<uint8_t> onbest;  // node-membership of the best path (findPathNode :101-108)
    std::vector<uint32_t> stamp;  // tabu (std::set in the reference, :70)
    uint32_t serial;
    std::vector<wro_ant> ants;    // last iteration's colony, in ant-index order
    int colony; float lambda, Q;
    uint64_t cnt[9];
    double phase[3];
};

extern "C" void wro_acs_default_params(wro_acs_params* p)
{
    p->alpha = 1; p->beta = 0.6; p->rho = 0.8; p->tau0 = 1; /* :319-324 */
    p->fixed_colony = 0; p->step_cap = 0; p->K = 6; p->seed = 0;
    p->rng_mode = WRO_RNG_KEYED; p->sort_mode = WRO_SORT_TOTAL;
}

/* :317-410 — slot order and per-slot distance; out-of-bounds slots start at 0 (:396) */
extern "C" wro_acs* wro_acs_create(const wro_grid* g, const wro_acs_params* p)
{
    if (p->K != 6 && p->K != 26) return nullptr;
    wro_acs* a = new wro_acs();
    a->g = g; a->p = *p; a->rx = g->rx; a->ry = g->ry; a->rz = g->rz; a->K = p->K;
    a->N = (size_t)g->rx * g->ry * g->rz;
    int s = 0;
    const float precision = g->precision;
    for (int i = -1; i <= 1; i++)
        for (int j = -1; j <= 1; j++)
            for (int k = -1; k <= 1; k++) {
                int type = i * j * k != 0 ? 3 : ((i == 0 && j * k != 0) || (j == 0 && i * k != 0) || (k == 0 && i * j != 0)) ? 2 : (i == 0 && j == 0 && k == 0) ? 0 : 1;
                float distance = type == 1 ? precision : type == 2 ? (a->K == 26 ? precision * 1.414f : 0) : type == 3 ? (a->K == 26 ? precision * 1.732f : 0) : 0;
                if (distance != 0) { a->dz[s] = i; a->dy[s] = j; a->dx[s] = k; a->dist[s] = distance; s++; }
            }
    a->tau.resize(a->N * a->K);
    size_t id = 0;
    for (int z = 0; z < a->rz; z++)
        for (int y = 0; y < a->ry; y++)
            for (int x = 0; x < a->rx; x++, id++)
                for (int k = 0; k < a->K; k++) {
                    int nx = x + a->dx[k], ny = y + a->dy[k], nz = z + a->dz[k];
                    bool oob = nx >= a->rx || nx < 0 || ny >= a->ry || ny < 0 || nz >= a->rz || nz < 0;
                    a->tau[id * a->K + k] = oob ? 0.f : p->tau0;
                }
    a->start = a->goal = -1; a->predict = 0; a->iter = 0; a->seq_calls = 0;
    a->search = 0; a->next_search = 0;
    a->best.L = WRO_INF_FLOAT; a->best.order = 0;
    a->onbest.assign(a->N, 0);
    a->stamp.assign(a->N, 0); a->serial = 0;
    a->colony = 0; a->lambda = 0; a->Q = 0;
    memset(a->cnt, 0, sizeof a->cnt); memset(a->phase, 0, sizeof a->phase);
    return a;
}
extern "C" void wro_acs_destroy(wro_acs* a) { delete a; }

/* :537-565, literal: scan every node in z,y,x order, last match wins */
extern "C" int wro_acs_set_points_scan(wro_acs* a, const float s[3], const float e[3], int64_t ids[2])
{
    const wro_grid* g = a->g;
    int findx = 0;
    int64_t sn = -1, en = -1, id = 0;
    for (int z = 0; z < a->rz; z++)
        for (int y = 0; y < a->ry; y++)
            for (int x = 0; x < a->rx; x++, id++) {
                float t = 1.2 * g->precision;
                if (wro_abs(s[0] - g->xs[x]) < t && wro_abs(s[1] - g->ys[y]) < t && wro_abs(s[2] - g->zs[z]) < t && g->isfree[id]) { sn = id; findx++; }
                if (wro_abs(e[0] - g->xs[x]) < t && wro_abs(e[1] - g->ys[y]) < t && wro_abs(e[2] - g->zs[z]) < t && g->isfree[id]) { en = id; findx++; }
            }
    a->start = sn; a->goal = en;
    ids[0] = sn; ids[1] = en;
    return (findx >= 2 && sn >= 0 && en >= 0) ? 1 : 0;
}

/* Same rule evaluated on the separable per-axis candidate lists (for big grids). */
static int64_t snap_one(const wro_acs* a, const float p[3])
{
    const wro_grid* g = a->g;
    float t = 1.2 * g->precision;
    std::vector<int> cx, cy, cz;
    for (int x = 0; x < a->rx; x++) if (wro_abs(p[0] - g->xs[x]) < t) cx.push_back(x);
    for (int y = 0; y < a->ry; y++) if (wro_abs(p[1] - g->ys[y]) < t) cy.push_back(y);
    for (int z = 0; z < a->rz; z++) if (wro_abs(p[2] - g->zs[z]) < t) cz.push_back(z);
    int64_t last = -1;
    for (int z : cz) for (int y : cy) for (int x : cx) {
        int64_t id = ((int64_t)z * a->ry + y) * a->rx + x;
        if (g->isfree[id] && id > last) last = id;
    }
    return last;
}
extern "C" int wro_acs_set_points(wro_acs* a, const float s[3], const float e[3], int64_t ids[2])
{
    a->start = snap_one(a, s); a->goal = snap_one(a, e);
    ids[0] = a->start; ids[1] = a->goal;
    return (a->start >= 0 && a->goal >= 0) ? 1 : 0;
}
extern "C" int wro_acs_set_endpoints(wro_acs* a, int64_t s, int64_t e)
{
    if (s < 0 || e < 0 || (size_t)s >= a->N || (size_t)e >= a->N) return 0;
    a->start = s; a->goal = e;
    return 1;
}

/* :229-233 */
extern "C" void wro_acs_begin(wro_acs* a, float predict)
{
    a->best.L = WRO_INF_FLOAT; /* best.path is left as is, like the reference */
    a->predict = predict; a->iter = 0; a->seq_calls = 0;
    a->search = a->next_search++;
}
/* index the NEXT begin() takes in the keyed stream (queries of a batch are numbered by the caller) */
extern "C" void wro_acs_set_next_search(wro_acs* a, uint32_t idx) { a->next_search = idx; }
/* position of the sequential (n-th call) stream: the genuine all-pairs driver (:472-499) never
 * reseeds between pairs, so a pinning run seeks to the cumulative draw count before each pair */
extern "C" void wro_acs_seq_seek(wro_acs* a, uint64_t pos) { a->seq_calls = pos; }
extern "C" uint64_t wro_acs_seq_tell(const wro_acs* a) { return a->seq_calls; }

/* :307-315 — every slot, including the out-of-bounds ones, becomes tau0 */
extern "C" void wro_acs_reset(wro_acs* a)
{
    for (size_t i = 0; i < a->tau.size(); i++) a->tau[i] = a->p.tau0;
}

static inline uint32_t draw31(wro_acs* a, uint32_t iter, uint32_t ant, uint32_t step)
{
    a->cnt[8]++;
    if (a->p.rng_mode == WRO_RNG_SEQUENTIAL) {
        uint64_t n = a->seq_calls++;
        return wr_rand31(a->p.seed, (uint32_t)n, (uint32_t)(n >> 32), 0, WR_STREAM_SEQ);
    }
    return wr_rand31_step(a->p.seed, a->search, iter, ant, step, WR_STREAM_ACS3D);
}

/* One construction step — :134-193.  Returns 1: moved and not at the goal, 0: stop.
 * `forced` >= 0 supplies the 31-bit draw (single-step KATs). */
static int select_next(wro_acs* a, wro_ant& ant, int64_t& cur, int cx, int cy, int cz, int* ncx, int* ncy, int* ncz,
                       uint32_t iter, uint32_t ant_idx, int64_t forced, float* infos_out)
{
    const wro_grid* g = a->g;
    const int K = a->K;
    if (a->p.step_cap > 0 && (int)ant.dirs.size() >= a->p.step_cap) { ant.L = WRO_INF_FLOAT; a->cnt[5]++; return 0; }
    int gz = (int)(a->goal / ((int64_t)a->rx * a->ry)), gr = (int)(a->goal % ((int64_t)a->rx * a->ry));
    int gy = gr / a->rx, gx = gr % a->rx;
    /* :137 vector_a = end - cur */
    float ax = g->xs[gx] - g->xs[cx], ay = g->ys[gy] - g->ys[cy], az = g->zs[gz] - g->zs[cz];
    float prob_sum = 0, total = 0;
    float info[26];
    bool cand[26];
    int ncand = 0;
    for (int i = 0; i < K; i++) { /* :142-159 */
        cand[i] = false;
        int nx = cx + a->dx[i], ny = cy + a->dy[i], nz = cz + a->dz[i];
        if (nx >= a->rx || nx < 0 || ny >= a->ry || ny < 0 || nz >= a->rz || nz < 0) continue; /* self slot: in tabu */
        int64_t nid = ((int64_t)nz * a->ry + ny) * a->rx + nx;
        if (a->stamp[nid] == a->serial) continue; /* tabu */
        if (!g->isfree[nid]) continue;
        float bx = g->xs[nx] - g->xs[cx], by = g->ys[ny] - g->ys[cy], bz = g->zs[nz] - g->zs[cz];
        float na = sqrtf(ax * ax + ay * ay + az * az), nb = sqrtf(bx * bx + by * by + bz * bz); /* :51-54 */
        float cos = (ax * bx + ay * by + az * bz) / (na * nb); /* :152 */
        info[i] = wro_power(a->tau[cur * K + i], a->p.alpha) * (1 + a->p.beta * cos); /* :154 */
        total += info[i];
        cand[i] = true; ncand++;
    }
    if (infos_out) for (int i = 0; i < K; i++) infos_out[i] = cand[i] ? info[i] : -12345.0f;
    if (ncand == 0) { ant.L = WRO_INF_FLOAT; a->cnt[3]++; return 0; } /* :162-166 */
    uint32_t r31 = forced >= 0 ? (uint32_t)forced : draw31(a, iter, ant_idx, (uint32_t)ant.dirs.size());
    float rnd = (float)(int)r31 / (float)2147483647; /* :169 */
    rnd *= total;
    for (int i = K - 1; i >= 0; i--) { /* :172-189, with the empty-J guard */
        if (cand[i]) {
            prob_sum += info[i];
            if (prob_sum >= rnd) {
                int nx = cx + a->dx[i], ny = cy + a->dy[i], nz = cz + a->dz[i];
                int64_t nid = ((int64_t)nz * a->ry + ny) * a->rx + nx;
                a->stamp[nid] = a->serial;      /* :75 */
                ant.ids.push_back((int32_t)nid);
                ant.dirs.push_back((uint8_t)i);
                ant.L += a->dist[i];            /* :78 */
                a->cnt[0]++;
                cur = nid; *ncx = nx; *ncy = ny; *ncz = nz;
                return nid != a->goal ? 1 : 0;  /* :182-186 */
            }
        }
    }
    /* fall-through: NaN (total or rnd), or rounding with rnd ~ total.  The reference reads
     * J.back() on an empty vector here (:174); de-facto outcome: dead end (:191-192). */
    if (total == total && rnd == rnd) a->cnt[6]++;
    ant.L = WRO_INF_FLOAT; a->cnt[4]++;
    return 0;
}

static void next_serial(wro_acs* a)
{
    if (++a->serial == 0) { std::fill(a->stamp.begin(), a->stamp.end(), 0u); a->serial = 1; }
}

static double now_s()
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

struct SortRec { float L; int idx; };

/* n passes of the loop body :237-299 */
extern "C" int wro_acs_iterate(wro_acs* a, int n)
{
    if (a->start < 0 || a->goal < 0) return -1;
    const wro_grid* g = a->g;
    const int K = a->K;
    const float precision = g->precision;
    const float predict_path_len = a->predict;
    const int sz0 = (int)(a->start / ((int64_t)a->rx * a->ry)), sr0 = (int)(a->start % ((int64_t)a->rx * a->ry));
    const int sy0 = sr0 / a->rx, sx0 = sr0 % a->rx;
    for (int it = 0; it < n; it++, a->iter++) {
        /* :247-250 */
        int colony_num = a->p.fixed_colony > 0 ? a->p.fixed_colony
                                               : (int)(0.35 * (a->best.L < predict_path_len ? a->best.L : predict_path_len) / precision);
        if (colony_num < 0) colony_num = 0;
        float lambda = 0.2 * colony_num;
        float Q = a->p.tau0 / lambda * (a->best.L == WRO_INF_FLOAT ? predict_path_len : a->best.L);
        a->colony = colony_num; a->lambda = lambda; a->Q = Q;
        a->ants.assign(colony_num, wro_ant());
        double t0 = now_s();
        /* :252-265 */
        for (int k = 0; k < colony_num; k++) {
            wro_ant& ant = a->ants[k];
            next_serial(a);
            a->stamp[a->start] = a->serial; /* addStartNode :81-86 */
            ant.ids.push_back((int32_t)a->start);
            ant.L = 0; ant.order = 0;
            int64_t cur = a->start;
            int cx = sx0, cy = sy0, cz = sz0;
            while (select_next(a, ant, cur, cx, cy, cz, &cx, &cy, &cz, (uint32_t)a->iter, (uint32_t)k, -1, nullptr)) {}
            a->cnt[1]++;
            if (ant.L != WRO_INF_FLOAT) a->cnt[2]++;
            if (ant.L < a->best.L) { /* :263-264 */
                for (int32_t id : a->best.ids) a->onbest[id] = 0;
                a->best = ant;
                for (int32_t id : a->best.ids) a->onbest[id] = 1;
            }
        }
        double t1 = now_s();
        /* :268-272 (the reference hard-codes k<6; the K=26 extension evaporates every slot) */
        { float rho = a->p.rho; float* t = a->tau.data(); size_t m = a->tau.size(); for (size_t i = 0; i < m; i++) t[i] *= rho; }
        double t2 = now_s();
        /* :273-274 */
        std::vector<SortRec> v(colony_num);
        for (int k = 0; k < colony_num; k++) { v[k].L = a->ants[k].L; v[k].idx = k; }
        if (a->p.sort_mode == WRO_SORT_STD)
            std::sort(v.begin(), v.end(), [](SortRec& x, SortRec& y) -> bool { return x.L < y.L; });
        else
            std::sort(v.begin(), v.end(), [](const SortRec& x, const SortRec& y) -> bool { return x.L < y.L || (x.L == y.L && x.idx < y.idx); });
        /* :275-280 with update_pheromone :198-215 */
        int agent_order = 1;
        for (int r = 0; r < colony_num; r++, agent_order++) {
            wro_ant& agentK = a->ants[v[r].idx];
            agentK.order = agent_order;
            int order = agent_order;
            if (agentK.L == WRO_INF_FLOAT || order > lambda - 1) continue;
            int _size = (int)agentK.ids.size();
            for (int i = 0; i < _size - 1; i++) {
                int32_t node = agentK.ids[i];
                int32_t nxt = agentK.ids[i + 1]; /* == adjacency_nodes[next_select[i]] */
                bool isOnBestPath = a->onbest[node] && a->onbest[nxt];
                a->tau[(size_t)node * K + agentK.dirs[i]] +=
                    (lambda - order) * Q / agentK.L + static_cast<float>(isOnBestPath) * lambda * Q / a->best.L;
            }
        }
        double t3 = now_s();
        a->phase[0] += t1 - t0; a->phase[1] += t2 - t1; a->phase[2] += t3 - t2;
        a->cnt[7]++;
    }
    return 0;
}

extern "C" int wro_acs_best(const wro_acs* a, int64_t* ids, int* dirs, int cap, float* L)
{
    *L = a->best.L;
    int n = (int)a->best.ids.size();
    for (int i = 0; i < n && i < cap; i++) ids[i] = a->best.ids[i];
    for (int i = 0; i < (int)a->best.dirs.size() && i < cap; i++) dirs[i] = a->best.dirs[i];
    return n;
}
extern "C" void wro_acs_pheromone(const wro_acs* a, float* out) { memcpy(out, a->tau.data(), 4 * a->tau.size()); }
extern "C" void wro_acs_set_pheromone(wro_acs* a, const float* in) { memcpy(a->tau.data(), in, 4 * a->tau.size()); }
extern "C" int wro_acs_last_colony(const wro_acs* a, int* colony, float* lambda, float* Q)
{
    *colony = a->colony; *lambda = a->lambda; *Q = a->Q;
    return (int)a->ants.size();
}
extern "C" int wro_acs_last_ant(const wro_acs* a, int k, int64_t* ids, int* dirs, int cap, float* L, int* order)
{
    if (k < 0 || k >= (int)a->ants.size()) return -1;
    const wro_ant& ant = a->ants[k];
    *L = ant.L; *order = ant.order;
    int n = (int)ant.ids.size();
    for (int i = 0; i < n && i < cap; i++) ids[i] = ant.ids[i];
    if (dirs) for (int i = 0; i < (int)ant.dirs.size() && i < cap; i++) dirs[i] = ant.dirs[i];
    return n;
}
extern "C" void wro_acs_counters(const wro_acs* a, uint64_t out[9]) { memcpy(out, a->cnt, sizeof a->cnt); }
extern "C" void wro_acs_phase_seconds(const wro_acs* a, double out[3]) { memcpy(out, a->phase, sizeof a->phase); }

extern "C" int wro_acs_select_step(wro_acs* a, int64_t cur_id, int64_t goal_id, const int64_t* tabu, int ntabu, uint32_t r31,
                                   float* infos, int* dir, int64_t* next_id, float* L_after)
{
    int64_t save_goal = a->goal;
    a->goal = goal_id;
    wro_ant ant;
    next_serial(a);
    a->stamp[cur_id] = a->serial;
    for (int i = 0; i < ntabu; i++) a->stamp[tabu[i]] = a->serial;
    ant.ids.push_back((int32_t)cur_id); ant.L = 0; ant.order = 0;
    int cz = (int)(cur_id / ((int64_t)a->rx * a->ry)), r = (int)(cur_id % ((int64_t)a->rx * a->ry));
    int cy = r / a->rx, cx = r % a->rx;
    int64_t cur = cur_id;
    int more = select_next(a, ant, cur, cx, cy, cz, &cx, &cy, &cz, 0, 0, (int64_t)r31, infos);
    *dir = ant.dirs.empty() ? -1 : ant.dirs[0];
    *next_id = ant.dirs.empty() ? -1 : ant.ids[1];
    *L_after = ant.L;
    a->goal = save_goal;
    return more;
}

// ------------------------------------------------------------------------------------------
// Seam ordering: ACS_GTSP.hpp
// ------------------------------------------------------------------------------------------
struct wro_gtsp {
    int n, colony_id, rng_mode;
    uint64_t seed, seq_calls, steps;
    std::vector<double> dis, ph, heur, info;
    double tau0, alpha;
    int delta, beta;
    int index_itera;
    std::vector<int> best_path; /* 2 ints per edge */
    double best_L;
};

/* readFromGraphFile :224-253 + init_param :187-218 (matrix passed in memory) */
extern "C" wro_gtsp* wro_gtsp_create(const double* dis, int n, int cnt, uint64_t seed, int colony_id, int rng_mode)
{
    wro_gtsp* g = new wro_gtsp();
    g->n = n; g->seed = seed; g->colony_id = colony_id; g->rng_mode = rng_mode; g->seq_calls = 0; g->steps = 0;
    g->dis.assign(dis, dis + (size_t)n * n);
    double tmp = 0;
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) tmp += g->dis[(size_t)i * n + j]; /* :246 */
    g->tau0 = (double)cnt / (tmp * n); /* :249 */
    g->alpha = 0.1; g->delta = 1; g->beta = 6; /* :189-191 */
    g->ph.assign((size_t)n * n, g->tau0);
    g->heur.resize((size_t)n * n); g->info.resize((size_t)n * n);
    for (size_t i = 0; i < (size_t)n * n; i++) g->heur[i] = 1 / (g->dis[i] + 1e-8); /* :211 */
    g->best_L = 0x3f3f3f3f; /* clean() :29-34 */
    g->index_itera = 0;
    return g;
}
extern "C" void wro_gtsp_destroy(wro_gtsp* g) { delete g; }

extern "C" int wro_gtsp_iterate(wro_gtsp* g, int iters, int early_stop)
{
    const int n = g->n;
    double last = 0x3f3f3f3f; /* :259 */
    int bad_times = 0, ran = 0;
    std::vector<std::vector<int>> tour(n);
    std::vector<uint8_t> inJ((size_t)n * n);
    std::vector<int> r1(n), r(n), left(n);
    for (int itc = 0; itc < iters; itc++, g->index_itera++) {
        if (early_stop && bad_times > n) break; /* :263-264 */
        /* reset :103-120 */
        for (int i = 0; i < n; i++) {
            tour[i].clear(); r1[i] = i; r[i] = i; left[i] = n - 1;
            for (int c = 0; c < n; c++) inJ[(size_t)i * n + c] = c != i;
        }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++)
                g->info[(size_t)i * n + j] = wro_power(g->ph[(size_t)i * n + j], g->delta) * wro_power(g->heur[(size_t)i * n + j], g->beta);
        /* construct_solution :146-159 (step-major, ant-minor) */
        for (int i = 0; i < n; i++)
            for (int k = 0; k < n; k++) {
                int next;
                if (left[k] == 0) next = r1[k]; /* select_next :124-125 */
                else {
                    uint32_t r31;
                    if (g->rng_mode == WRO_RNG_SEQUENTIAL) { uint64_t c = g->seq_calls++; r31 = wr_rand31(g->seed, (uint32_t)c, (uint32_t)(c >> 32), 0, WR_STREAM_SEQ); }
                    else r31 = wr_rand31(g->seed, (uint32_t)g->index_itera, (uint32_t)k, (uint32_t)i, WR_STREAM_GTSP + (uint32_t)g->colony_id);
                    double rnd = (double)(int)r31 / (double)2147483647; /* :126 */
                    const double* row = &g->info[(size_t)r[k] * n];
                    const uint8_t* J = &inJ[(size_t)k * n];
                    double sum_prob = 0, sum = 0;
                    for (int c = 0; c < n; c++) if (J[c]) sum += row[c]; /* :129-132 */
                    rnd *= sum;
                    next = r1[k]; /* :143 */
                    for (int c = 0; c < n; c++) if (J[c]) { sum_prob += row[c]; if (sum_prob >= rnd) { next = c; break; } }
                }
                if (inJ[(size_t)k * n + next]) { inJ[(size_t)k * n + next] = 0; left[k]--; } /* J.erase :153 */
                tour[k].push_back(r[k]); tour[k].push_back(next); /* :155 */
                r[k] = next;
                g->steps++;
            }
        /* update_pheromone :161-185 */
        double now_L = 0x3f3f3f3f; int now = -1;
        for (int k = 0; k < n; k++) {
            double L = 0; int sz = (int)tour[k].size() / 2;
            for (int e = 0; e < sz - 1; e++) L += g->dis[(size_t)tour[k][2 * e] * n + tour[k][2 * e + 1]]; /* calc :36-44 */
            if (L < now_L) { now_L = L; now = k; }
        }
        if (now >= 0 && now_L < g->best_L) { g->best_L = now_L; g->best_path = tour[now]; }
        for (size_t i = 0; i < (size_t)n * n; i++) g->ph[i] *= (1 - g->alpha); /* :175-177 */
        if (now >= 0) {
            int sz = (int)tour[now].size() / 2;
            for (int e = 0; e < sz; e++) {
                int rr = tour[now][2 * e], ss = tour[now][2 * e + 1];
                g->ph[(size_t)rr * n + ss] += 1. / (double)now_L; /* :182 */
                g->ph[(size_t)ss * n + rr] = g->ph[(size_t)rr * n + ss]; /* :183 */
            }
        }
        ran++;
        if (last > g->best_L) { last = g->best_L; bad_times = 0; } else bad_times++; /* :269-275 */
    }
    return ran;
}
extern "C" int wro_gtsp_best(const wro_gtsp* g, int* tour, double* L)
{
    *L = g->best_L;
    memcpy(tour, g->best_path.data(), sizeof(int) * g->best_path.size());
    return (int)g->best_path.size() / 2;
}
extern "C" void wro_gtsp_pheromone(const wro_gtsp* g, double* out) { memcpy(out, g->ph.data(), 8 * g->ph.size()); }
extern "C" double wro_gtsp_tau0(const wro_gtsp* g) { return g->tau0; }
extern "C" uint64_t wro_gtsp_steps(const wro_gtsp* g) { return g->steps; }

extern "C" void wro_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { wr_philox4x32_10(ctr, key, out); }

// ---- trajectory smoothing: BS_Basic<float, 3, DEGREE, CI, CF> (core/BSplineBasic.h) ---------------------------------------------
// SetParam (:70-76) + getCurvePoint (:85-111) for m times, with run-time degree / constraint levels.  Float arithmetic in the
// reference's order (compiled with -ffp-contract=off).  The one element the reference reads without ever writing it
// (c_mat[idx][CF + 2 - h] for DEGREE < CF + 1, :403) is 0 here, as in oracle/ref_harness.cpp's zeroing `new`; the bisection of
// _findSpan is bounded (64 trips -> the call fails) where the reference would not terminate.
namespace {
struct Spline {
    int degree, ci, cf, nk, ncp;
    std::vector<float> K, C;   // knots, control points [ncp][3]
    float left(int i, int j, float u) const { return u - K[i + 1 - j]; }     // _Left :339
    float right(int i, int j, float u) const { return K[i + j] - u; }        // _Right :341
    bool find_span(int& ret, float u) const
    {   // :345-379
        if (u < K[0] || K[nk - 1] < u) return false;
        const float dd = u - K[nk - 1];
        const float sq = dd * dd;
        if ((double)sq < 1.e-10) {   // SP_IS_EQUAL :8
            for (int i = nk - 2; i > -1; --i)
                if (K[i] < u && u <= K[i + 1]) { ret = i; return true; }
            return false;
        }
        int low = 0, high = nk - 1, mid = (low + high) >> 1, trips = 0;
        while (u < K[mid] || u >= K[mid + 1]) {
            if (++trips > 64) return false;
            if (u < K[mid]) high = mid; else low = mid;
            mid = (low + high) >> 1;
        }
        ret = mid;
        return true;
    }
    // _BasisFunsDers(ders, u, n) :195-201 + :203-291; ders rows have `cols` entries, zero where the reference writes nothing
    void basis_ders(std::vector<float>& ders, int cols, float u, int n) const
    {
        int span;
        if (!find_span(span, u)) return;
        const int p = degree;
        std::vector<float> ndu((size_t)(p + 1) * (p + 1), 0.0f), a((size_t)2 * (p + 1), 0.0f);
        auto ND = [&](int i, int j) -> float& { return ndu[(size_t)i * (p + 1) + j]; };
        auto A = [&](int i, int j) -> float& { return a[(size_t)i * (p + 1) + j]; };
        ND(0, 0) = 1.0f;
        for (int j = 1; j <= p; ++j) {
            float saved = 0.0f;
            for (int r = 0; r < j; ++r) {
                const float l = left(span, j - r, u), rr = right(span, r + 1, u);
                ND(j, r) = rr + l;
                const float temp = ND(r, j - 1) / ND(j, r);
                ND(r, j) = saved + rr * temp;
                saved = l * temp;
            }
            ND(j, j) = saved;
        }
        for (int j = 0; j <= p; ++j) ders[j] = ND(j, p);
        for (int r = 0; r <= p; ++r) {
            int s1 = 0, s2 = 1;
            A(0, 0) = 1.0f;
            for (int k = 1; k <= n; ++k) {
                float d = 0.0f;
                const int rk = r - k, pk = p - k;
                if (r >= k) { A(s2, 0) = A(s1, 0) / ND(pk + 1, rk); d = A(s2, 0) * ND(rk, pk); }
                const int j1 = rk >= -1 ? 1 : -rk, j2 = (r - 1 <= pk) ? k - 1 : p - r;
                for (int j = j1; j <= j2; ++j) {
                    A(s2, j) = (A(s1, j) - A(s1, j - 1)) / ND(pk + 1, rk + j);
                    d += A(s2, j) * ND(rk + j, pk);
                }
                if (r <= pk) { A(s2, k) = -A(s1, k - 1) / ND(pk + 1, r); d += A(s2, k) * ND(r, pk); }
                ders[(size_t)k * cols + r] = d;
                const int t = s1; s1 = s2; s2 = t;
            }
        }
        int r = p;
        for (int k = 1; k <= n; ++k) {
            for (int j = 0; j <= p; ++j) ders[(size_t)k * cols + j] *= r;
            r *= (p - k);
        }
    }
};
}  // namespace

extern "C" int wro_bspline(int degree, int ci, int cf, const float* init, const float* fin, const float* middle, int n_mid, int mid_stride, float tf,
                           const float* u, int m, float* out, unsigned char* ok, float* knots, float* cps)
{
    if (degree < 0 || degree > 8 || ci < 0 || cf < 0 || ci > degree || cf > degree) return -1;
    Spline sp;
    sp.degree = degree; sp.ci = ci; sp.cf = cf;
    sp.nk = degree + n_mid + 2 + ci + cf + 1;
    sp.ncp = n_mid + 2 + ci + cf;
    sp.K.assign(sp.nk, 0.0f);
    sp.C.assign((size_t)sp.ncp * 3, 0.0f);
    {   // _CalcKnot :149-164
        int i = 0;
        const int nmid = sp.nk - 2 * degree - 2;
        const float step = tf / (nmid + 1);
        for (int j = 0; j < degree + 1; ++j) sp.K[i++] = 0.0f;
        for (int j = 0; j < nmid; ++j) { sp.K[i] = sp.K[i - 1] + step; ++i; }
        for (int j = 0; j < degree + 1; ++j) sp.K[i++] = tf;
    }
    {   // _CalcConstrainedCPoints :381-427
        for (int d = 0; d < 3; ++d) { sp.C[d] = init[d]; sp.C[(size_t)(sp.ncp - 1) * 3 + d] = fin[d]; }
        const int cols = std::max(std::max(ci, cf) + 2, degree + 1);
        std::vector<float> dm((size_t)(ci + 1) * cols, 0.0f);
        sp.basis_ders(dm, cols, 0.0f, ci);
        for (int j = 1; j < ci + 1; ++j)
            for (int k = 0; k < 3; ++k) {
                float c = init[j * 3 + k];
                for (int h = j; h > 0; --h) c -= dm[(size_t)j * cols + h - 1] * sp.C[(size_t)(h - 1) * 3 + k];
                sp.C[(size_t)j * 3 + k] = c / dm[(size_t)j * cols + j];
            }
        std::vector<float> cm((size_t)(cf + 1) * cols, 0.0f);
        sp.basis_ders(cm, cols, tf, cf);
        int idx = 1;
        for (int j = sp.ncp - 2; j > sp.ncp - 2 - cf; --j) {
            for (int k = 0; k < 3; ++k) {
                float c = fin[idx * 3 + k];
                for (int h = idx; h > 0; --h) c -= cm[(size_t)idx * cols + cf + 2 - h] * sp.C[(size_t)(sp.ncp - h) * 3 + k];
                sp.C[(size_t)j * 3 + k] = c / cm[(size_t)idx * cols + cf + 1 - idx];
            }
            ++idx;
        }
    }
    for (int i = 0; i < n_mid; ++i)   // _CalcCPoints :441-447
        for (int d = 0; d < 3; ++d) sp.C[(size_t)(ci + 1 + i) * 3 + d] = middle[(size_t)i * mid_stride + d];
    if (knots) memcpy(knots, sp.K.data(), sizeof(float) * sp.K.size());
    if (cps) memcpy(cps, sp.C.data(), sizeof(float) * sp.C.size());
    for (int i = 0; i < m; ++i) {   // getCurvePoint :85-111
        float t = u[i];
        if (t < sp.K[0]) t = sp.K[0];
        else if (t > sp.K[sp.nk - 1]) t = sp.K[sp.nk - 1];
        int span;
        if (!sp.find_span(span, t)) { if (ok) ok[i] = 0; continue; }
        std::vector<float> N(degree + 1, 0.0f);
        float temp = 0.0f;
        N[0] = 1.0f;
        for (int j = 1; j <= degree; ++j) {   // _BasisFuns :316-338
            float saved = 0.0f;
            for (int r = 0; r < j; ++r) {
                const float l = sp.left(span, j - r, t), rr = sp.right(span, r + 1, t);
                if ((rr + l) != 0) temp = N[r] / (rr + l);
                N[r] = saved + rr * temp;
                saved = l * temp;
            }
            N[j] = saved;
        }
        for (int d = 0; d < 3; ++d) {
            float c = 0.0f;
            for (int q = 0; q <= degree; ++q) c += N[q] * sp.C[(size_t)(span - degree + q) * 3 + d];
            out[3 * i + d] = c;
        }
        if (ok) ok[i] = 1;
    }
    return 0;
}
