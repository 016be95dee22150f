// Headless planning pipeline — the reference's "demo 3" (main.cpp:268-283, and its commented-out
// headless twin :500-525) without CoppeliaSim, on the B200-native facade:
//   STL -> voxel grid -> all-pairs ant-colony path search -> seam ordering -> stitched path.
// Usage: headless_main <file.stl> [precision=0.005] [wall=10] [predict=0.5] [workdir=/tmp]
// The reference's weld-point file is not shipped (.gitignore:43); six points on the free plane
// x = min_x - 2*precision of the mesh box are synthesised.
#include <stdlib.h>

#include "ACSRank_3D.hpp"
#include "read_STL.hpp"
#include "ACS_GTSP.hpp"

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <file.stl> [precision] [wall] [predict] [workdir]\n", argv[0]); return 2; }
    const float precision = argc > 2 ? atof(argv[2]) : 0.005f;
    const int wall = argc > 3 ? atoi(argv[3]) : 10;
    const float predict = argc > 4 ? atof(argv[4]) : 0.5f;
    const std::string dir = argc > 5 ? argv[5] : "/tmp";
    try {
        STLReader model;
        ACS_Rank SearchPath;
        ACS_GTSP GlobalRoute;
        model.readFile(argv[1]);
        const std::vector<Triangles<float>> meshes = model.TriangleList();
        SearchPath.creatGridMap(meshes, precision, wall, dir + "/grid_map.in");

        float mn[3] = {meshes[0].vertex[0].x, meshes[0].vertex[0].y, meshes[0].vertex[0].z}, mx[3] = {mn[0], mn[1], mn[2]};
        for (const auto& t : meshes)
            for (int i = 0; i < 3; i++) {
                const float v[3] = {t.vertex[i].x, t.vertex[i].y, t.vertex[i].z};
                for (int k = 0; k < 3; k++) { mn[k] = v[k] < mn[k] ? v[k] : mn[k]; mx[k] = v[k] > mx[k] ? v[k] : mx[k]; }
            }
        const std::string points = dir + "/weld_points.in", graph = dir + "/graph.in";
        FILE* fp = fopen(points.c_str(), "w");
        if (!fp) { fprintf(stderr, "cannot write %s\n", points.c_str()); return 1; }
        const float px = mn[0] - 2 * precision;
        const float ys[3] = {mn[1] - 2 * precision, 0.5f * (mn[1] + mx[1]), mx[1] - 3 * precision};
        const float zs[2] = {mn[2] - 2 * precision, mx[2] - 3 * precision};
        fprintf(fp, "6\n");
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 2; b++) fprintf(fp, "%.6f %.6f %.6f\n", px, ys[a], zs[b]);
        fclose(fp);

        SearchPath.searchBestPathOfPoints(predict, points, graph);
        GlobalRoute.readFromGraphFile(graph);
        GlobalRoute.computeSolution();
        GlobalRoute.read_all_segments(SearchPath.best_matrix);
        uint64_t c[9];
        SearchPath.counters(c);
        printf("[headless] stitched path: %zu points over %d segments; last search: %llu ant-steps, %llu ants\n", GlobalRoute.g_path_x.size(),
               GlobalRoute.path_segment_nums(), (unsigned long long)c[0], (unsigned long long)c[1]);
        // ---- curve smoothing (main.cpp:287-352): degree-0 spline over the stitched path, then the constrained degree-2 spline over
        //      its samples.  The demo samples both at wall-clock times (every >= 10 / >= 50 clock ticks); here at fixed times.
        const int pt_num = (int)GlobalRoute.g_path_x.size();
        if (pt_num >= 2) {
            float start_pt[3] = {GlobalRoute.g_path_x[0], GlobalRoute.g_path_y[0], GlobalRoute.g_path_z[0]};
            float end_pt[3] = {GlobalRoute.g_path_x[pt_num - 1], GlobalRoute.g_path_y[pt_num - 1], GlobalRoute.g_path_z[pt_num - 1]};
            std::vector<float> rows((size_t)pt_num * 3);
            std::vector<float*> ctrl(pt_num);
            for (int i = 0; i < pt_num; i++) {
                rows[3 * i] = GlobalRoute.g_path_x[i]; rows[3 * i + 1] = GlobalRoute.g_path_y[i]; rows[3 * i + 2] = GlobalRoute.g_path_z[i];
                ctrl[i] = &rows[3 * i];
            }
            BS_Basic<float, 3, 0, 0, 0> smooth_curve(pt_num);
            smooth_curve.SetParam(start_pt, end_pt, ctrl.data(), 150);
            std::vector<float> t1, first;
            for (int t = 10; t <= 160; t += 10) t1.push_back((float)t);
            smooth_curve.getCurvePoints(t1, first);
            const int n2 = (int)t1.size();
            const float constrain = 0.05f;
            float s2[9] = {start_pt[0], start_pt[1], start_pt[2], 0, 0, 0, 0, 0, 0}, e2[9] = {end_pt[0], end_pt[1], end_pt[2], 0, 0, 0, 0, 0, 0};
            (void)constrain;   // main.cpp:329-331 fills columns 3..8 of the middle points with it; _CalcCPoints never reads them
            std::vector<float*> ctrl2(n2);
            for (int i = 0; i < n2; i++) ctrl2[i] = &first[3 * i];
            BS_Basic<float, 3, 2, 2, 2> second_curve(n2);
            second_curve.SetParam(s2, e2, ctrl2.data(), 6000);
            std::vector<float> t2, smooth;
            for (int t = 50; t <= 6050; t += 50) t2.push_back((float)t);
            second_curve.getCurvePoints(t2, smooth);
            printf("[headless] smoothed path: %zu points; first (%.6f, %.6f, %.6f), last (%.6f, %.6f, %.6f)\n", smooth.size() / 3, smooth[0], smooth[1], smooth[2],
                   smooth[smooth.size() - 3], smooth[smooth.size() - 2], smooth[smooth.size() - 1]);
        }
    } catch (const wr::Error& e) {
        fprintf(stderr, "[headless] %s (status %d)\n", e.what(), e.status);
        return 1;
    }
    return 0;
}
