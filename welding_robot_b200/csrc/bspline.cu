// Trajectory smoothing of the stitched path: the clamped B-spline of core/BSplineBasic.h, as the demo uses it
// (main.cpp:299-300 BS_Basic<float, 3, 0, 0, 0>, :337-338 BS_Basic<float, 3, 2, 2, 2>).
//
//   SetParam (BSplineBasic.h:70-76)      host, O(#points): knots accumulated in float (:149-164), end control points from the
//                                        position / velocity / acceleration constraints (_CalcConstrainedCPoints :381-427 with the
//                                        derivative basis functions of _BasisFunsDers :203-291), middle points copied (:441-447)
//   getCurvePoint (:85-111)              k_bspline_eval: one thread per sample time — clamp, _findSpan (:345-379), _BasisFuns
//                                        (:311-338, incl. the `_temp` that survives a zero denominator), DEGREE + 1 products per
//                                        coordinate added in the reference's order
// All arithmetic is float with explicit round-to-nearest intrinsics on the device and -ffp-contract=off on the host, so a curve
// point equals the reference's bit for bit.  Two reference defects are pinned rather than copied:
//   * BS_Basic<T, DIM, 2, 2, 2> reads c_mat[1][3], an element _BasisFunsDers never writes (BSplineBasic.h:403-404): it is 0 here
//     (what a fresh heap gives the demo; oracle/ref_harness.cpp makes the reference deterministic the same way);
//   * _findSpan's bisection does not terminate on a non-monotone knot vector (float accumulation can push the last inner knot
//     past fin_time): bounded here at 64 trips, the sample is then reported as failed like an out-of-range time.
// The demo samples the curve at wall-clock times (main.cpp:309-320), which is not reproducible; here the caller passes the times.
#include <math.h>

#include <vector>

#include "wr_internal.cuh"

namespace wr {

constexpr int kBsMaxDegree = 5;
constexpr int kBsDim = 3;

struct BsplineArgs {
    const float* knots;   // [nknots]
    const float* cps;     // [ncp][3]
    const float* u;       // [m]
    float* out;           // [m][3]
    unsigned char* ok;    // [m] getCurvePoint's return value
    int nknots, ncp, degree, m;
};

// SP_IS_EQUAL (BSplineBasic.h:8): the float product is compared with a double literal
__host__ __device__ __forceinline__ bool bs_is_equal(float x, float y)
{
#ifdef __CUDA_ARCH__
    const float d = __fsub_rn(x, y);
    return (double)__fmul_rn(d, d) < 1.e-10;
#else
    const float d = x - y;
    const float p = d * d;
    return (double)p < 1.e-10;
#endif
}

// _findSpan :345-379
__host__ __device__ inline bool bs_find_span(const float* K, int nknots, float u, int& ret)
{
    if (u < K[0] || K[nknots - 1] < u) return false;
    if (bs_is_equal(u, K[nknots - 1])) {
        for (int i = nknots - 2; i > -1; --i)
            if (K[i] < u && u <= K[i + 1]) { ret = i; return true; }
        return false;
    }
    int low = 0, high = nknots - 1, mid = (low + high) >> 1;
    for (int trips = 0; u < K[mid] || u >= K[mid + 1]; trips++) {
        if (trips >= 64) return false;   // the reference would spin (see the header)
        if (u < K[mid]) high = mid; else low = mid;
        mid = (low + high) >> 1;
    }
    ret = mid;
    return true;
}

__global__ void __launch_bounds__(256) k_bspline_eval(BsplineArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.m) return;
    const float* K = a.knots;
    float u = a.u[i];
    if (u < K[0]) u = K[0];
    else if (u > K[a.nknots - 1]) u = K[a.nknots - 1];
    int span = 0;
    if (!bs_find_span(K, a.nknots, u, span)) { a.ok[i] = 0; return; }   // `ret` keeps its previous contents (:96)
    // _BasisFuns :316-338
    float N[kBsMaxDegree + 1];
    float temp = 0.0f;
    N[0] = 1.0f;
#pragma unroll
    for (int j = 1; j <= kBsMaxDegree; ++j) {
        if (j <= a.degree) {
            float saved = 0.0f;
#pragma unroll
            for (int r = 0; r < j; ++r) {
                const float left = __fsub_rn(u, K[span + 1 - (j - r)]);
                const float right = __fsub_rn(K[span + r + 1], u);
                const float den = __fadd_rn(right, left);
                if (den != 0.0f) temp = __fdiv_rn(N[r], den);
                N[r] = __fadd_rn(saved, __fmul_rn(right, temp));
                saved = __fmul_rn(left, temp);
            }
            N[j] = saved;
        }
    }
    // :100-105
#pragma unroll
    for (int j = 0; j < kBsDim; ++j) {
        float c = 0.0f;
#pragma unroll
        for (int q = 0; q <= kBsMaxDegree; ++q)
            if (q <= a.degree) c = __fadd_rn(c, __fmul_rn(N[q], a.cps[(size_t)(span - a.degree + q) * kBsDim + j]));
        a.out[(size_t)i * kBsDim + j] = c;
    }
    a.ok[i] = 1;
}

// ---- host side of SetParam ---------------------------------------------------------------------------------------------------
// _BasisFunsDers(ders, span, u, n) :203-291 (algorithm A2.3 of the NURBS book, in float); ders is (n + 1) rows of degree + 1
static void bs_basis_ders(const std::vector<float>& K, int degree, int span, float u, int n, std::vector<std::vector<float>>& ders)
{
    const int p = degree;
    std::vector<std::vector<float>> ndu(p + 1, std::vector<float>(p + 1, 0.0f)), a(2, std::vector<float>(p + 1, 0.0f));
    auto Left = [&](int i, int j) { return u - K[i + 1 - j]; };     // :339
    auto Right = [&](int i, int j) { return K[i + j] - u; };        // :341
    ndu[0][0] = 1.0f;
    for (int j = 1; j <= p; ++j) {
        float saved = 0.0f;
        for (int r = 0; r < j; ++r) {
            const float left = Left(span, j - r), right = Right(span, r + 1);
            ndu[j][r] = right + left;
            const float temp = ndu[r][j - 1] / ndu[j][r];
            ndu[r][j] = saved + right * temp;
            saved = left * temp;
        }
        ndu[j][j] = saved;
    }
    for (int j = 0; j <= p; ++j) ders[0][j] = ndu[j][p];
    for (int r = 0; r <= p; ++r) {
        int s1 = 0, s2 = 1;
        a[0][0] = 1.0f;
        for (int k = 1; k <= n; ++k) {
            float d = 0.0f;
            const int rk = r - k, pk = p - k;
            if (r >= k) {
                a[s2][0] = a[s1][0] / ndu[pk + 1][rk];
                d = a[s2][0] * ndu[rk][pk];
            }
            const int j1 = rk >= -1 ? 1 : -rk;
            const int j2 = (r - 1 <= pk) ? k - 1 : p - r;
            for (int j = j1; j <= j2; ++j) {
                a[s2][j] = (a[s1][j] - a[s1][j - 1]) / ndu[pk + 1][rk + j];
                d += a[s2][j] * ndu[rk + j][pk];
            }
            if (r <= pk) {
                a[s2][k] = -a[s1][k - 1] / ndu[pk + 1][r];
                d += a[s2][k] * ndu[r][pk];
            }
            ders[k][r] = d;
            std::swap(s1, s2);
        }
    }
    int r = p;
    for (int k = 1; k <= n; ++k) {
        for (int j = 0; j <= p; ++j) ders[k][j] *= (float)r;   // `ders[_k][_j] *= _r` with an int _r (:282)
        r *= (p - k);
    }
}

}  // namespace wr

using namespace wr;

// BS_Basic<float, 3, degree, ci, cf>(n_middle).SetParam(init, fin, middle, fin_time), then getCurvePoint(u[i], out + 3 i) for every
// i (BSplineBasic.h:36-111).  Host pointers; `middle` rows are `middle_stride` floats of which the first three are used.
extern "C" int wr_bspline_eval(int degree, int ci, int cf, const float* init, const float* fin, const float* middle, int n_middle, int middle_stride,
                               float fin_time, const float* u, int m, float* out, unsigned char* ok, float* knots_out, float* cps_out)
{
    WR_REQUIRE(degree >= 0 && degree <= kBsMaxDegree && ci >= 0 && ci <= 2 && cf >= 0 && cf <= 2, WR_ERR_INVALID,
               "wr_bspline_eval: degree 0..5, constraint levels 0..2 (position, + velocity, + acceleration)");
    WR_REQUIRE(init && fin && n_middle >= 0 && (n_middle == 0 || (middle && middle_stride >= kBsDim)) && m >= 0 && (m == 0 || (u && out)), WR_ERR_INVALID,
               "wr_bspline_eval: bad argument");
    const int nknots = degree + n_middle + 2 + ci + cf + 1;      // :39-40
    const int ncp = n_middle + 2 + ci + cf;                      // :41
    WR_REQUIRE(nknots >= 2 * (degree + 1), WR_ERR_INVALID, "wr_bspline_eval: invalid setup (num_knots < 2 * (degree + 1), BSplineBasic.h:54-56)");
    WR_REQUIRE(ci <= degree && cf <= degree, WR_ERR_INVALID, "wr_bspline_eval: a constraint level above the degree divides by a zero basis derivative");
    // _CalcKnot :149-164
    std::vector<float> K(nknots, 0.0f);
    {
        int i = 0;
        const int nmid = nknots - 2 * degree - 2;
        const float step = fin_time / (float)(nmid + 1);
        for (int j = 0; j < degree + 1; ++j) K[i++] = 0.0f;
        for (int j = 0; j < nmid; ++j) { K[i] = K[i - 1] + step; ++i; }
        for (int j = 0; j < degree + 1; ++j) K[i++] = fin_time;
    }
    std::vector<float> C((size_t)ncp * kBsDim, 0.0f);
    // _CalcConstrainedCPoints :381-427
    for (int d = 0; d < kBsDim; ++d) { C[d] = init[d]; C[(size_t)(ncp - 1) * kBsDim + d] = fin[d]; }
    {
        std::vector<std::vector<float>> dm(ci + 1, std::vector<float>(std::max(ci + 2, degree + 1), 0.0f));
        int span = 0;
        if (bs_find_span(K.data(), nknots, 0.0f, span)) bs_basis_ders(K, degree, span, 0.0f, ci, dm);
        for (int j = 1; j < ci + 1; ++j)
            for (int k = 0; k < kBsDim; ++k) {
                float c = init[j * kBsDim + k];
                for (int h = j; h > 0; --h) c -= dm[j][h - 1] * C[(size_t)(h - 1) * kBsDim + k];
                C[(size_t)j * kBsDim + k] = c / dm[j][j];
            }
    }
    {
        // rows of max(cf + 2, degree + 1) entries, zero-filled: the reference allocates cf + 2 and fills degree + 1, so for
        // degree < cf + 1 the element [cf + 2 - h] it reads (:403) was never written — 0, see the header
        std::vector<std::vector<float>> cm(cf + 1, std::vector<float>(std::max(cf + 2, degree + 1), 0.0f));
        int span = 0;
        if (bs_find_span(K.data(), nknots, fin_time, span)) bs_basis_ders(K, degree, span, fin_time, cf, cm);
        int idx = 1;
        for (int j = ncp - 2; j > ncp - 2 - cf; --j) {
            for (int k = 0; k < kBsDim; ++k) {
                float c = fin[idx * kBsDim + k];
                for (int h = idx; h > 0; --h) c -= cm[idx][cf + 2 - h] * C[(size_t)(ncp - h) * kBsDim + k];
                C[(size_t)j * kBsDim + k] = c / cm[idx][cf + 1 - idx];
            }
            ++idx;
        }
    }
    // _CalcCPoints :441-447
    for (int i = 0; i < n_middle; ++i)
        for (int d = 0; d < kBsDim; ++d) C[(size_t)(ci + 1 + i) * kBsDim + d] = middle[(size_t)i * middle_stride + d];
    if (knots_out) std::copy(K.begin(), K.end(), knots_out);
    if (cps_out) std::copy(C.begin(), C.end(), cps_out);
    if (m == 0) return WR_OK;

    float *d_k = nullptr, *d_c = nullptr, *d_u = nullptr, *d_o = nullptr;
    unsigned char* d_ok = nullptr;
    cudaStream_t s = nullptr;
    auto cleanup = [&]() { pool_free(d_k, s); pool_free(d_c, s); pool_free(d_u, s); pool_free(d_o, s); pool_free(d_ok, s); };
#define WR_CUDA_B(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); cleanup(); return WR_ERR_CUDA; } } while (0)
    WR_CUDA_B(dmalloc(&d_k, K.size() * sizeof(float), s));
    WR_CUDA_B(dmalloc(&d_c, C.size() * sizeof(float), s));
    WR_CUDA_B(dmalloc(&d_u, (size_t)m * sizeof(float), s));
    WR_CUDA_B(dmalloc(&d_o, (size_t)m * kBsDim * sizeof(float), s));
    WR_CUDA_B(dmalloc(&d_ok, (size_t)m, s));
    WR_CUDA_B(cudaMemcpyAsync(d_k, K.data(), K.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    WR_CUDA_B(cudaMemcpyAsync(d_c, C.data(), C.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    WR_CUDA_B(cudaMemcpyAsync(d_u, u, (size_t)m * sizeof(float), cudaMemcpyHostToDevice, s));
    WR_CUDA_B(cudaMemcpyAsync(d_o, out, (size_t)m * kBsDim * sizeof(float), cudaMemcpyHostToDevice, s));   // failed samples keep the caller's values
    BsplineArgs a;
    a.knots = d_k; a.cps = d_c; a.u = d_u; a.out = d_o; a.ok = d_ok; a.nknots = nknots; a.ncp = ncp; a.degree = degree; a.m = m;
    k_bspline_eval<<<(m + 255) / 256, 256, 0, s>>>(a);
    WR_CUDA_B(cudaGetLastError());
    std::vector<unsigned char> hok((size_t)m);
    WR_CUDA_B(cudaMemcpyAsync(out, d_o, (size_t)m * kBsDim * sizeof(float), cudaMemcpyDeviceToHost, s));
    WR_CUDA_B(cudaMemcpyAsync(hok.data(), d_ok, (size_t)m, cudaMemcpyDeviceToHost, s));
    WR_CUDA_B(cudaStreamSynchronize(s));
#undef WR_CUDA_B
    if (ok) std::copy(hok.begin(), hok.end(), ok);
    cleanup();
    return WR_OK;
}
