// TEST INFRASTRUCTURE (oracle/): no-op stand-in for common/matplotlibcpp.h.
// The reference's model_grid_map.hpp:13 includes "matplotlibcpp.h" only for its
// plot_* helpers (model_grid_map.hpp:368-379, ACSRank_3D.hpp:567-598,
// ACS_GTSP.hpp:319-326), none of which is on the compute path.  Putting this
// directory first on the include path lets the UNMODIFIED reference headers
// compile without libpython / numpy headers.
#pragma once
#include <map>
#include <string>
#include <vector>
namespace matplotlibcpp {
template <class... A> inline bool scatter(const A&...) { return true; }
template <class... A> inline bool plot3(const A&...) { return true; }
template <class... A> inline bool plot(const A&...) { return true; }
inline void show() {}
inline void cla() {}
}  // namespace matplotlibcpp
