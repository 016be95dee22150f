// Seam ordering (ACS_GTSP) — placeholder until K4 lands; exports keep the ABI complete.
#include "wr_internal.cuh"
using namespace wr;
struct wr_gtsp { int n; };
extern "C" int wr_gtsp_create(const double*, int, int, int, int, uint64_t, wr_gtsp**) { set_error("wr_gtsp: not implemented yet"); return WR_ERR_STATE; }
extern "C" int wr_gtsp_destroy(wr_gtsp*) { return WR_OK; }
extern "C" int wr_gtsp_iterate(wr_gtsp*, int) { set_error("wr_gtsp: not implemented yet"); return WR_ERR_STATE; }
extern "C" int wr_gtsp_sync(wr_gtsp*) { set_error("wr_gtsp: not implemented yet"); return WR_ERR_STATE; }
extern "C" int wr_gtsp_best(wr_gtsp*, int, int*, int*, double*) { set_error("wr_gtsp: not implemented yet"); return WR_ERR_STATE; }
extern "C" int wr_gtsp_download_pheromone(wr_gtsp*, int, double*) { set_error("wr_gtsp: not implemented yet"); return WR_ERR_STATE; }
extern "C" int wr_gtsp_tau0(wr_gtsp*, double*) { set_error("wr_gtsp: not implemented yet"); return WR_ERR_STATE; }
extern "C" int wr_gtsp_kernel_ms(wr_gtsp*, float*) { set_error("wr_gtsp: not implemented yet"); return WR_ERR_STATE; }
