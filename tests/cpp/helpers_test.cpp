// CPU test of the node-cuboid helper templates (model_grid_map.hpp:90-138 contract) of the C++ facade.
#include <stdio.h>

#include <vector>

#include "welding_robot.hpp"

int main()
{
    const int rx = 5, ry = 3, rz = 4;
    Vertex3<float>*** m = NULL;
    std::vector<int> order;
    creat_all_nodes(m, rx, ry, rz, [&](int z, int y, int x) {
        m[z][y][x].id = (unsigned long)((z * ry + y) * rx + x);
        m[z][y][x].isFree = (x + y + z) % 2 == 0;
        order.push_back((z * ry + y) * rx + x);
    });
    for (size_t i = 0; i < order.size(); i++) if (order[i] != (int)i) { printf("creat order\n"); return 1; }
    if (order.size() != (size_t)rx * ry * rz) { printf("creat count\n"); return 1; }
    unsigned long sum = 0; int visited = 0, last = -1; bool mono = true;
    for_each_nodes(m, rx, ry, rz, [&](int z, int y, int x) {
        sum += m[z][y][x].id; visited++;
        const int id = (z * ry + y) * rx + x;
        mono = mono && id == last + 1; last = id;
    });
    if (!mono || visited != rx * ry * rz || sum != (unsigned long)(rx * ry * rz) * (rx * ry * rz - 1) / 2) { printf("for_each\n"); return 1; }
    if (&m[1][0][0] != &m[0][ry - 1][rx - 1] + 1) { printf("not contiguous\n"); return 1; }
    bool threw = false;
    try { creat_all_nodes(m, rx, ry, rz, [](int, int, int) {}); } catch (const wr::Error&) { threw = true; }
    if (!threw) { printf("double create accepted\n"); return 1; }
    delete_all_nodes(m, rx, ry, rz);
    if (m != NULL) { printf("delete leaves pointer\n"); return 1; }
    threw = false;
    try { for_each_nodes(m, rx, ry, rz, [](int, int, int) {}); } catch (const wr::Error&) { threw = true; }
    if (!threw) { printf("for_each on NULL accepted\n"); return 1; }
    printf("ok\n");
    return 0;
}
