// CPU: the C++ facade's BS_Basic<float, 3, D, CI, CF>::SetParam (BSplineBasic.h:70-76) through the C ABI — the host half of the
// trajectory smoothing (knots, control points); no kernel runs when no sample time is passed.  Prints the values as hex words so that
// the Python test can compare them bit for bit with the reference fixture.
#include <stdio.h>
#include <string.h>

#include <vector>

#include "compat/BSplineBasic.h"

template <class Curve>
static void dump(Curve& c)
{
    for (float v : c.knots()) { unsigned u; memcpy(&u, &v, 4); printf("%08x ", u); }
    printf("\n");
    for (float v : c.controlPoints()) { unsigned u; memcpy(&u, &v, 4); printf("%08x ", u); }
    printf("\n");
}

int main(int argc, char** argv)
{
    // stdin: degree ci cf n_mid fin_time, then 3*(ci+1) init, 3*(cf+1) fin, n_mid*3 middle floats (as hex words)
    int d, ci, cf, n;
    unsigned tfw;
    if (scanf("%d %d %d %d %x", &d, &ci, &cf, &n, &tfw) != 5) return 2;
    float tf; memcpy(&tf, &tfw, 4);
    auto rd = [](int k) { std::vector<float> v(k); for (float& x : v) { unsigned u; if (scanf("%x", &u) != 1) u = 0; memcpy(&x, &u, 4); } return v; };
    std::vector<float> init = rd(3 * (ci + 1)), fin = rd(3 * (cf + 1)), mid = rd(3 * n);
    std::vector<float*> rows(n);
    for (int i = 0; i < n; i++) rows[i] = &mid[3 * i];
    try {
        if (d == 0 && ci == 0 && cf == 0) { BS_Basic<float, 3, 0, 0, 0> c(n); c.SetParam(init.data(), fin.data(), rows.data(), tf); dump(c); }
        else if (d == 2 && ci == 2 && cf == 2) { BS_Basic<float, 3, 2, 2, 2> c(n); c.SetParam(init.data(), fin.data(), rows.data(), tf); dump(c); }
        else if (d == 3 && ci == 2 && cf == 1) { BS_Basic<float, 3, 3, 2, 1> c(n); c.SetParam(init.data(), fin.data(), rows.data(), tf); dump(c); }
        else return 3;
    } catch (const wr::Error& e) { fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}
