/* TEST INFRASTRUCTURE (oracle/) — not part of the product.
 *
 * C ABI of the CPU restatement of the reference's hot path (wr_oracle.cpp).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it;
 * the product (welding_robot_b200/) never does.
 *
 * PARITY PINNING: the restatement is checked against the UNMODIFIED reference compiled
 * from /root/reference into oracle/_ref/libwrref.so (ref_harness.cpp) — voxel grids
 * bit-exact, single selectNext steps bit-exact, and whole searches (best path, best
 * length, full pheromone field after 1/2/10/150 iterations) bit-exact under a shared
 * sequential Philox stream (rng_mode = WRO_RNG_SEQUENTIAL, sort_mode = WRO_SORT_STD) —
 * see tests/test_oracle_vs_reference.py, and against the committed fixtures in
 * tests/golden/ that the same comparison produced (tests/golden/make_golden.py).
 */
#ifndef WR_ORACLE_H
#define WR_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct wro_grid wro_grid;
typedef struct wro_acs wro_acs;
typedef struct wro_gtsp wro_gtsp;

enum { WRO_RNG_KEYED = 0, WRO_RNG_SEQUENTIAL = 1 };
enum { WRO_SORT_TOTAL = 0, WRO_SORT_STD = 1 };
enum { WRO_VOX_BRUTE = 0, WRO_VOX_AABB = 1 };

typedef struct {
    int alpha;          /* ACSRank_3D.hpp:319  (1)   */
    float beta;         /* :320 (0.6)  */
    float rho;          /* :321 (0.8)  */
    float tau0;         /* :324 (1)    */
    int fixed_colony;   /* 0 = adaptive rule of :247 */
    int step_cap;       /* 0 = unbounded (reference); else an ant may take at most this many steps */
    int K;              /* 6 (reference) or 26 (extension, :367-385) */
    uint64_t seed;
    int rng_mode;       /* WRO_RNG_*  */
    int sort_mode;      /* WRO_SORT_* */
} wro_acs_params;

void wro_acs_default_params(wro_acs_params* p);

/* ---- STL (read_STL.hpp:131-174) ---- */
int wro_stl_parse(const uint8_t* buf, size_t len, float* tris12, int cap);

/* ---- grid (model_grid_map.hpp:151-273) ---- */
wro_grid* wro_grid_from_triangles(const float* tris12, int ntri, float precision, int wall, int mode);
wro_grid* wro_grid_from_occupancy(const uint8_t* isfree, int rx, int ry, int rz, const float* xs, const float* ys,
                                  const float* zs, float precision);
void wro_grid_destroy(wro_grid* g);
void wro_grid_dims(const wro_grid* g, int dims[3]);
float wro_grid_precision(const wro_grid* g);
void wro_grid_isfree(const wro_grid* g, uint8_t* out);
void wro_grid_coords(const wro_grid* g, float* xs, float* ys, float* zs);
uint64_t wro_grid_tests(const wro_grid* g); /* triangle-node predicate evaluations performed */
/* text dump / reload in the reference's format (model_grid_map.hpp:275-356); compat!=0 keeps the
 * header-clobber bug (:279 writes the last triangle's box), compat==0 writes the global box. */
int wro_grid_write_file(const wro_grid* g, const char* path, int compat);
wro_grid* wro_grid_read_file(const char* path);

/* ---- rank-based 3-D ACS (ACSRank_3D.hpp) ---- */
wro_acs* wro_acs_create(const wro_grid* g, const wro_acs_params* p);
void wro_acs_destroy(wro_acs* a);
int wro_acs_set_points(wro_acs* a, const float s[3], const float e[3], int64_t ids[2]); /* :537-565 */
int wro_acs_set_points_scan(wro_acs* a, const float s[3], const float e[3], int64_t ids[2]); /* literal full scan */
int wro_acs_set_endpoints(wro_acs* a, int64_t start_id, int64_t goal_id);
void wro_acs_begin(wro_acs* a, float predict_path_len);   /* :229-233; takes the next search index of the keyed stream */
void wro_acs_set_next_search(wro_acs* a, uint32_t idx);    /* keyed stream: search index the next begin() uses (default: begin() calls so far) */
void wro_acs_seq_seek(wro_acs* a, uint64_t pos);           /* sequential-stream position (pinning runs) */
uint64_t wro_acs_seq_tell(const wro_acs* a);
int wro_acs_iterate(wro_acs* a, int n);                    /* n passes of the loop body :237-299 */
void wro_acs_reset(wro_acs* a);                            /* :307-315 */
int wro_acs_best(const wro_acs* a, int64_t* ids, int* dirs, int cap, float* L); /* node count */
void wro_acs_pheromone(const wro_acs* a, float* out);      /* N*K floats */
void wro_acs_set_pheromone(wro_acs* a, const float* in);
/* last iteration's colony */
int wro_acs_last_colony(const wro_acs* a, int* colony, float* lambda, float* Q);
int wro_acs_last_ant(const wro_acs* a, int ant, int64_t* ids, int* dirs, int cap, float* L, int* order);
/* counters: [0] ant-steps [1] ants [2] arrived [3] dead (no candidate) [4] dead (fall-through/NaN)
 *           [5] dead (step cap) [6] finite fall-through (the :174 UB case) [7] iterations [8] rng draws */
void wro_acs_counters(const wro_acs* a, uint64_t out[9]);
/* phase seconds: [0] walk [1] evaporate [2] sort+deposit */
void wro_acs_phase_seconds(const wro_acs* a, double out[3]);
/* single selectNext (:134-193) on a fresh ant; mirrors wrref_acs_select_step */
int wro_acs_select_step(wro_acs* a, int64_t cur_id, int64_t goal_id, const int64_t* tabu, int ntabu, uint32_t r31,
                        float* infos, int* dir, int64_t* next_id, float* L_after);

/* ---- seam ordering (ACS_GTSP.hpp) ---- */
wro_gtsp* wro_gtsp_create(const double* dis, int n, int cnt, uint64_t seed, int colony_id, int rng_mode);
void wro_gtsp_destroy(wro_gtsp* g);
int wro_gtsp_iterate(wro_gtsp* g, int iters, int early_stop); /* returns iterations run (:261-276) */
int wro_gtsp_best(const wro_gtsp* g, int* tour_pairs, double* L);
void wro_gtsp_pheromone(const wro_gtsp* g, double* out);
double wro_gtsp_tau0(const wro_gtsp* g);
uint64_t wro_gtsp_steps(const wro_gtsp* g);

/* ---- trajectory smoothing: BS_Basic<float, 3, degree, ci, cf> SetParam + getCurvePoint at m times (BSplineBasic.h:70-111);
 * mirrors wrref_bspline.  Returns 0, -1 for an unsupported setup. */
int wro_bspline(int degree, int ci, int cf, const float* init, const float* fin, const float* middle, int n_mid, int mid_stride, float tf,
                const float* u, int m, float* out, unsigned char* ok, float* knots, float* cps);

/* ---- Philox KAT hook ---- */
void wro_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
