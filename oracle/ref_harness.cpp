// TEST INFRASTRUCTURE (oracle/) — not part of the product.
//
// Headless driver around the UNMODIFIED reference headers, compiled from where they lie
// (/root/reference/core/*.hpp) into oracle/_ref/libwrref.so by oracle/Makefile.  It exists
// to pin oracle/wr_oracle.cpp (the CPU restatement the GPU is checked against) to the real
// reference, and to serve as the "reference" CPU baseline in bench.py.
//
// Nothing in the reference is edited.  Three preprocessor hooks, applied in THIS
// translation unit only, make it drivable and reproducible:
//   * rand  -> wrref_rand_hook : the n-th call returns Philox(seed; n) >> 1 instead of
//     glibc's additive-feedback generator (ACSRank_3D.hpp:169, ACS_GTSP.hpp:126);
//   * srand -> wrref_srand_hook: swallows srand(time(0)) (ACSRank_3D.hpp:327);
//   * private/protected -> public: lets the driver call computeSolution / selectNext /
//     reset directly and read the pheromone field, exactly as searchBestPathOfPoints
//     (ACSRank_3D.hpp:472-499) would.
// The plotting header is replaced by oracle/stub/matplotlibcpp.h (include-path order).
//
// Known reference UB that this build inherits (documented in DESIGN.md): J.back() on an
// empty vector (ACSRank_3D.hpp:174).  glibc's chunk header makes that read 0, so the only
// observable effect is on a roulette fall-through, which is counted by the oracle
// restatement and asserted absent in every pinned run.
#include <bits/stdc++.h>
#include <fcntl.h>
#include <stdint.h>
#include <unistd.h>

#include "philox.h"

static uint64_t g_seed = 0;
static uint64_t g_calls = 0;
static int g_fixed = -1;  // >=0: every call returns this value (single-step KATs)

static int wrref_rand_hook()
{
    if (g_fixed >= 0) { g_calls++; return g_fixed; }
    uint32_t r = wr_rand31(g_seed, (uint32_t)g_calls, (uint32_t)(g_calls >> 32), 0, WR_STREAM_SEQ);
    g_calls++;
    return (int)r;
}
static void wrref_srand_hook(unsigned) {}

#define rand wrref_rand_hook
#define srand wrref_srand_hook
#define private public
#define protected public
#include "core/ACSRank_3D.hpp"
#include "core/read_STL.hpp"
#include "core/ACS_GTSP.hpp"
#undef private
#undef protected
#undef rand
#undef srand

// BSplineBasic.h (trajectory smoothing, main.cpp:287-352).  One more hook, for this header only: `new T[n]` hands out ZEROED
// memory.  BS_Basic<T, DIM, 2, 2, 2> — the demo's second curve, main.cpp:337 — reads c_mat[1][3] in _CalcConstrainedCPoints
// (BSplineBasic.h:403-404), an element _BasisFunsDers never writes for DEGREE = 2 (it fills DEGREE + 1 = 3 entries of a row of
// CONST_LEVEL_FIN + 2 = 4): an uninitialised heap read.  In a fresh process glibc returns zero pages, so 0 is what the demo sees;
// the hook makes that outcome deterministic, and the restatement and the GPU path define the element as 0.
struct wrref_zero_t {};
static wrref_zero_t wrref_zero;
inline void* operator new[](size_t n, wrref_zero_t&) { return calloc(1, n ? n : 1); }
#define new new (wrref_zero)
#define private public
#include "core/BSplineBasic.h"
#undef private
#undef new

namespace {
struct Quiet {  // the reference prints progress with printf/cout; keep test logs clean
    int saved = -1;
    Quiet()
    {
        if (getenv("WR_REF_VERBOSE")) return;
        fflush(stdout); std::cout.flush();
        saved = dup(1);
        int n = open("/dev/null", O_WRONLY);
        dup2(n, 1); close(n);
    }
    ~Quiet()
    {
        if (saved < 0) return;
        fflush(stdout); std::cout.flush();
        dup2(saved, 1); close(saved);
    }
};

std::vector<Triangles<float>> mesh_from(const float* t12, int n)
{
    std::vector<Triangles<float>> m(n);
    for (int i = 0; i < n; i++) {
        const float* p = t12 + 12 * i;
        m[i].nor_vec = Point3<float>(p[0], p[1], p[2]);
        for (int j = 0; j < 3; j++) m[i].vertex[j] = Point3<float>(p[3 + 3 * j], p[4 + 3 * j], p[5 + 3 * j]);
        m[i].trait = 0;
    }
    return m;
}
}  // namespace

extern "C" {

// read_STL.hpp:26-77.  Returns the triangle count; fills min(count, cap) triangles as
// 12 floats each (normal, v0, v1, v2).
int wrref_stl_read(const char* path, float* t12, int cap)
{
    Quiet q;
    STLReader r;
    r.readFile(path);
    const std::vector<Triangles<float>>& m = r.TriangleList();
    int n = (int)m.size();
    for (int i = 0; i < n && i < cap; i++) {
        float* p = t12 + 12 * i;
        p[0] = m[i].nor_vec.x; p[1] = m[i].nor_vec.y; p[2] = m[i].nor_vec.z;
        for (int j = 0; j < 3; j++) { p[3 + 3 * j] = m[i].vertex[j].x; p[4 + 3 * j] = m[i].vertex[j].y; p[5 + 3 * j] = m[i].vertex[j].z; }
    }
    return r.NumTri();
}

void* wrref_create() { return new ACS_Rank(); }
void wrref_destroy(void* h) { delete (ACS_Rank*)h; }  // leaks what the reference leaks

// model_grid_map.hpp:151-298 (no dump file).
int wrref_voxelize(void* h, const float* t12, int ntri, float precision, int wall, int dims[3])
{
    Quiet q;
    ACS_Rank* a = (ACS_Rank*)h;
    std::vector<Triangles<float>> m = mesh_from(t12, ntri);
    a->creatGridMap(m, precision, wall, "");
    dims[0] = a->rangeX; dims[1] = a->rangeY; dims[2] = a->rangeZ;
    return a->size_of_map();
}

// Dump / reload through the reference's text format (model_grid_map.hpp:275-356).
int wrref_voxelize_to_file(void* h, const float* t12, int ntri, float precision, int wall, const char* file, int dims[3])
{
    Quiet q;
    ACS_Rank* a = (ACS_Rank*)h;
    std::vector<Triangles<float>> m = mesh_from(t12, ntri);
    a->creatGridMap(m, precision, wall, file);
    dims[0] = a->rangeX; dims[1] = a->rangeY; dims[2] = a->rangeZ;
    return a->size_of_map();
}
int wrref_read_grid_file(void* h, const char* file, int dims[3])
{
    Quiet q;
    ACS_Rank* a = (ACS_Rank*)h;
    a->readGridMap(file);
    dims[0] = a->rangeX; dims[1] = a->rangeY; dims[2] = a->rangeZ;
    return a->size_of_map();
}

// isfree: N bytes in z,y,x order; xs/ys/zs: the per-axis coordinates (they are separable,
// model_grid_map.hpp:204-211) read back from the node cuboid.
int wrref_grid_read(void* h, uint8_t* isfree, float* xs, float* ys, float* zs)
{
    ACS_Rank* a = (ACS_Rank*)h;
    Vertex3<float>*** g = a->ptr_grid_map();
    if (!g) return -1;
    size_t n = 0;
    for (int z = 0; z < a->rangeZ; z++)
        for (int y = 0; y < a->rangeY; y++)
            for (int x = 0; x < a->rangeX; x++) isfree[n++] = g[z][y][x].isFree ? 1 : 0;
    if (xs) for (int x = 0; x < a->rangeX; x++) xs[x] = g[0][0][x].pt.x;
    if (ys) for (int y = 0; y < a->rangeY; y++) ys[y] = g[0][y][0].pt.y;
    if (zs) for (int z = 0; z < a->rangeZ; z++) zs[z] = g[z][0][0].pt.z;
    return 0;
}

// Overwrite occupancy (synthetic grids): keeps the reference's coordinates.
int wrref_grid_set_free(void* h, const uint8_t* isfree)
{
    ACS_Rank* a = (ACS_Rank*)h;
    Vertex3<float>*** g = a->ptr_grid_map();
    if (!g) return -1;
    size_t n = 0;
    for (int z = 0; z < a->rangeZ; z++)
        for (int y = 0; y < a->rangeY; y++)
            for (int x = 0; x < a->rangeX; x++) g[z][y][x].isFree = isfree[n++] != 0;
    return 0;
}

// ACSRank_3D.hpp:317-410
int wrref_acs_init(void* h)
{
    Quiet q;
    ((ACS_Rank*)h)->initFromGridMap();
    return 0;
}

// ACSRank_3D.hpp:537-565.  ids[0]=start id, ids[1]=end id (-1 when not found).
int wrref_acs_set_points(void* h, const float s[3], const float e[3], int64_t ids[2])
{
    ACS_Rank* a = (ACS_Rank*)h;
    Point3<float> ps(s[0], s[1], s[2]), pe(e[0], e[1], e[2]);
    bool ok = a->setPoints(ps, pe);
    ids[0] = a->start_node ? (int64_t)a->start_node->id : -1;
    ids[1] = a->end_node ? (int64_t)a->end_node->id : -1;
    return ok ? 1 : 0;
}

// choose the endpoints by node id (what setPoints would leave in start_node / end_node)
int wrref_acs_set_endpoints(void* h, int64_t start_id, int64_t goal_id)
{
    ACS_Rank* a = (ACS_Rank*)h;
    int rx = a->rangeX, ry = a->rangeY;
    auto node = [&](int64_t id) { int z = id / ((int64_t)rx * ry); int r = id % ((int64_t)rx * ry); return &a->nodes[z][r / rx][r % rx]; };
    a->start_node = node(start_id);
    a->end_node = node(goal_id);
    return (a->start_node->isFree && a->end_node->isFree) ? 1 : 0;
}

// ACSRank_3D.hpp:220-305 with max_iteration overridden and rand() = Philox SEQ stream.
int wrref_acs_compute(void* h, float predict, int max_iter, uint64_t seed, uint64_t* rand_calls)
{
    Quiet q;
    ACS_Rank* a = (ACS_Rank*)h;
    a->max_iteration = max_iter;
    g_seed = seed; g_calls = 0; g_fixed = -1;
    a->computeSolution(predict);
    if (rand_calls) *rand_calls = g_calls;
    return 0;
}

// best path read-out (ACSRank_3D.hpp:506-509, Agent::getPath/nodeIndex :93-100).
// Returns node count; ids gets min(count,cap) ids, dirs count-1 directions.
int wrref_acs_best(void* h, int64_t* ids, int* dirs, int cap, float* L)
{
    ACS_Rank* a = (ACS_Rank*)h;
    const Agent<float>* b = a->getSolution();
    *L = b->L;
    const std::vector<ACS_Node<float>*>* p = b->getPath();
    const std::vector<int>* d = b->nodeIndex();
    int n = (int)p->size();
    for (int i = 0; i < n && i < cap; i++) ids[i] = (int64_t)(*p)[i]->id;
    for (int i = 0; i < (int)d->size() && i < cap; i++) dirs[i] = (*d)[i];
    return n;
}

// N*6 floats, node-major, slot order as ACSRank_3D.hpp:355-359.
int wrref_acs_pheromone(void* h, float* out)
{
    ACS_Rank* a = (ACS_Rank*)h;
    size_t n = 0;
    for (int z = 0; z < a->rangeZ; z++)
        for (int y = 0; y < a->rangeY; y++)
            for (int x = 0; x < a->rangeX; x++)
                for (int k = 0; k < 6; k++) out[n++] = a->nodes[z][y][x].adjacency_infos[k].pheromone;
    return 0;
}

void wrref_acs_reset(void* h) { ((ACS_Rank*)h)->reset(); }  // ACSRank_3D.hpp:307-315

// One selectNext call (ACSRank_3D.hpp:134-193) on a fresh ant standing on cur_id with the
// given tabu ids, goal goal_id, and rand() forced to r31.  Reports the six info values the
// call left in the node, the chosen direction (-1: none) and the ant's L afterwards.
int wrref_acs_select_step(void* h, int64_t cur_id, int64_t goal_id, const int64_t* tabu, int ntabu, int r31,
                          float infos[6], int* dir, int64_t* next_id, float* L_after)
{
    ACS_Rank* a = (ACS_Rank*)h;
    int rx = a->rangeX, ry = a->rangeY;
    auto node = [&](int64_t id) { int z = id / ((int64_t)rx * ry); int r = id % ((int64_t)rx * ry); return &a->nodes[z][r / rx][r % rx]; };
    ACS_Node<float>* cur = node(cur_id);
    a->end_node = node(goal_id);
    Agent<float> ant;
    ant.addStartNode(cur);
    for (int i = 0; i < ntabu; i++) ant.tabu_list.insert((unsigned long)tabu[i]);
    for (int k = 0; k < 6; k++) cur->adjacency_infos[k].info = -12345.0f;
    ACS_Node<float>* next = nullptr;
    g_fixed = r31;
    bool more = a->selectNext(ant, cur, next);
    g_fixed = -1;
    for (int k = 0; k < 6; k++) infos[k] = cur->adjacency_infos[k].info;
    const std::vector<int>* d = ant.nodeIndex();
    *dir = d->empty() ? -1 : (*d)[0];
    *next_id = d->empty() ? -1 : (int64_t)next->id;
    *L_after = ant.L;
    return more ? 1 : 0;
}

// The genuine all-pairs driver (ACSRank_3D.hpp:427-504) through its file interface.
// lens: npts*npts floats (best_matrix[i][j].L); returns the number of pairs written.
int wrref_acs_search_all(void* h, const float* pts, int npts, float predict, uint64_t seed,
                         const char* tmp_points, const char* tmp_graph, float* lens)
{
    Quiet q;
    ACS_Rank* a = (ACS_Rank*)h;
    FILE* fp = fopen(tmp_points, "w");
    if (!fp) return -1;
    fprintf(fp, "%d\n", npts);
    for (int i = 0; i < npts; i++) fprintf(fp, "%.9g %.9g %.9g\n", pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    fclose(fp);
    g_seed = seed; g_calls = 0; g_fixed = -1;
    a->searchBestPathOfPoints(predict, tmp_points, tmp_graph);
    int cnt = 0;
    for (int i = 0; i < npts; i++)
        for (int j = 0; j < npts; j++) {
            lens[i * npts + j] = (i == j) ? 0.f : a->best_matrix[i][j].L;
            if (i < j) cnt++;
        }
    return cnt;
}
int wrref_acs_pair_best(void* h, int i, int j, int64_t* ids, int cap, float* L)
{
    ACS_Rank* a = (ACS_Rank*)h;
    const Agent<float>& b = a->best_matrix[i][j];
    *L = b.L;
    const std::vector<ACS_Node<float>*>* p = b.getPath();
    int n = (int)p->size();
    for (int k = 0; k < n && k < cap; k++) ids[k] = (int64_t)(*p)[k]->id;
    return n;
}

// ---- seam ordering (ACS_GTSP.hpp) -------------------------------------------------------
// dis: full N*N symmetric matrix; written to a temp graph file with %.17g (exact round trip)
// because readFromGraphFile (ACS_GTSP.hpp:224-253) is the only way in.
void* wrref_gtsp_create(const double* dis, int n, const char* tmp_graph)
{
    Quiet q;
    FILE* fp = fopen(tmp_graph, "w");
    if (!fp) return nullptr;
    fprintf(fp, "%d %d\n", n, n * (n - 1) / 2);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) fprintf(fp, "%.17g\n", dis[i * n + j]);
    fclose(fp);
    ACS_GTSP* g = new ACS_GTSP();
    g->readFromGraphFile(tmp_graph);
    return g;
}

// ACS_GTSP.hpp:255-284 with MAX_itera overridden; returns iterations actually run.
int wrref_gtsp_run(void* h, int iters, uint64_t seed, uint64_t* rand_calls)
{
    Quiet q;
    ACS_GTSP* g = (ACS_GTSP*)h;
    g->MAX_itera = g->index_itera + iters;
    int before = g->index_itera;
    g_seed = seed; g_calls = 0; g_fixed = -1;
    g->computeSolution();
    if (rand_calls) *rand_calls = g_calls;
    return g->index_itera - before;
}

// tour: 2*N ints (r,s per edge); returns the edge count.
int wrref_gtsp_best(void* h, int* tour, double* L)
{
    ACS_GTSP* g = (ACS_GTSP*)h;
    *L = g->best.L;
    int n = g->best.size();
    for (int i = 0; i < n; i++) { tour[2 * i] = g->best.r(i); tour[2 * i + 1] = g->best.s(i); }
    return n;
}
int wrref_gtsp_pheromone(void* h, double* out)
{
    ACS_GTSP* g = (ACS_GTSP*)h;
    int n = g->city_num;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[i * n + j] = g->pheromone[i][j];
    return n;
}
double wrref_gtsp_tau0(void* h) { return ((ACS_GTSP*)h)->pheromone_0; }

// BS_Basic<float, 3, DEGREE, CI, CF>::SetParam + getCurvePoint (BSplineBasic.h:70-120) at m times.  init / fin hold
// 3 * (CI + 1) / 3 * (CF + 1) floats (position, velocity, acceleration); middle points are rows of `mid_stride` floats of which the
// first three are used (main.cpp:325-334 passes rows of nine).  ok[i] = getCurvePoint's return value; out rows of a failed call keep
// their previous contents, as `res` does in the demo.  Returns 0, or -1 for a (DEGREE, CI, CF) this harness does not instantiate.
}  // extern "C"
template <int DEG, int CI, int CF>
static int bspline_run(const float* init, const float* fin, const float* middle, int n_mid, int mid_stride, float tf, const float* u, int m, float* out,
                       unsigned char* ok, float* knots, float* cps)
{
    std::vector<float> a(init, init + 3 * (CI + 1)), b(fin, fin + 3 * (CF + 1));
    std::vector<std::vector<float>> rows(n_mid);
    std::vector<float*> mp(n_mid);
    for (int i = 0; i < n_mid; i++) { rows[i].assign(middle + (size_t)i * mid_stride, middle + (size_t)i * mid_stride + 3); mp[i] = rows[i].data(); }
    BS_Basic<float, 3, DEG, CI, CF> c(n_mid);
    c.SetParam(a.data(), b.data(), mp.data(), tf);
    if (knots) for (int i = 0; i < c.NumKnots_; i++) knots[i] = c.Knots_[i];
    if (cps) for (int i = 0; i < c.NumCPs_; i++) for (int j = 0; j < 3; j++) cps[3 * i + j] = c.CPoints_[i][j];
    for (int i = 0; i < m; i++) {
        const bool r = c.getCurvePoint(u[i], out + 3 * i);
        if (ok) ok[i] = r ? 1 : 0;
    }
    return 0;
}
extern "C" {
int wrref_bspline(int degree, int ci, int cf, const float* init, const float* fin, const float* middle, int n_mid, int mid_stride, float tf,
                  const float* u, int m, float* out, unsigned char* ok, float* knots, float* cps)
{
    Quiet q;
#define WRREF_BS(D, A, B) if (degree == D && ci == A && cf == B) return bspline_run<D, A, B>(init, fin, middle, n_mid, mid_stride, tf, u, m, out, ok, knots, cps)
    WRREF_BS(0, 0, 0); WRREF_BS(1, 0, 0); WRREF_BS(2, 0, 0); WRREF_BS(3, 0, 0); WRREF_BS(2, 1, 1); WRREF_BS(3, 1, 1);
    WRREF_BS(2, 2, 2); WRREF_BS(3, 2, 2); WRREF_BS(4, 2, 2); WRREF_BS(5, 2, 2); WRREF_BS(3, 2, 1); WRREF_BS(3, 0, 2);
#undef WRREF_BS
    return -1;
}

}  // extern "C"
