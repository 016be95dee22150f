// One ant-colony search sharded over N GPUs from C++ — no Python, no torch: the reference's class surface
// (STLReader, ACS_Rank) on the B200-native facade, one process per GPU, NCCL for the rendezvous only
// (wr_comm_unique_id / ACS_Rank::commInit), the iteration loop inside libwrgpu.so (wr_acs_iterate).
//
//   sharded_main <file.stl> <ranks> [precision=0.012] [wall=4] [ants=4096] [iterations=12] [workdir=/tmp]
//
// The parent starts `ranks` workers (fork + exec of itself; worker r drives GPU r) plus one single-GPU run of the same
// colony, then compares what they report: digest of the whole pheromone field, best length, best path.  Exit code 0 =
// every rank of the sharded search reproduced the single-GPU result bit for bit.
#include <stdlib.h>
#include <sys/wait.h>
#include <unistd.h>

#include "ACSRank_3D.hpp"
#include "read_STL.hpp"

static unsigned long long digest(const std::vector<float>& tau)
{
    unsigned long long h = 0;
    for (size_t i = 0; i < tau.size(); i++) {
        unsigned int w;
        memcpy(&w, &tau[i], 4);
        h += (unsigned long long)w * ((i * 0x9E3779B97F4A7C15ull) | 1ull);
    }
    return h;
}

static int worker(int argc, char** argv)
{   // --worker <rank> <ranks> <device> <tag> <file.stl> <precision> <wall> <ants> <iterations> <workdir>
    const int rank = atoi(argv[2]), ranks = atoi(argv[3]), device = atoi(argv[4]);
    const std::string tag = argv[5], stl = argv[6], dir = argv[11];
    const float precision = atof(argv[7]);
    const int wall = atoi(argv[8]), ants = atoi(argv[9]), iterations = atoi(argv[10]);
    try {
        wr::check(wr_set_device(device));
        STLReader model;
        ACS_Rank search;
        model.readFile(stl);
        search.materialise_cuboid = false;
        search.params.fixed_colony = ants; search.params.step_cap = 1200; search.params.seed = 21;
        search.creatGridMap(model.TriangleList(), precision, wall);
        search.initFromGridMap();
        const std::vector<uint8_t> fr = search.isFreeArray();
        long long s = -1, e = -1, seen = 0;
        for (size_t i = 0; i < fr.size(); i++) if (fr[i]) { if (seen++ == 11) s = (long long)i; }
        seen = 0;
        for (size_t i = fr.size(); i-- > 0;) if (fr[i]) { if (seen++ == 11) { e = (long long)i; break; } }
        search.setEndpoints(s, e);
        if (ranks > 1) {
            unsigned char id[WR_COMM_ID_BYTES];
            const std::string idfile = dir + "/wr_nccl_id." + tag;
            if (rank == 0) {
                wr::check(wr_comm_unique_id(id));
                FILE* fp = fopen((idfile + ".tmp").c_str(), "wb");
                if (!fp || fwrite(id, 1, sizeof id, fp) != sizeof id) throw wr::Error(WR_ERR_FORMAT, "cannot write " + idfile);
                fclose(fp);
                rename((idfile + ".tmp").c_str(), idfile.c_str());
            } else {
                FILE* fp = nullptr;
                for (int t = 0; t < 600 && !(fp = fopen(idfile.c_str(), "rb")); t++) usleep(100000);
                if (!fp || fread(id, 1, sizeof id, fp) != sizeof id) throw wr::Error(WR_ERR_FORMAT, "cannot read " + idfile);
                fclose(fp);
            }
            search.commInit(id, rank, ranks);
        }
        search.max_iteration = iterations;
        search.computeSolution(1.0f);            // begin (+ exchange of the peer slabs) + iterations, ACSRank_3D.hpp:220-305
        search.sync();
        const Agent<float>* best = search.getSolution();
        unsigned long long pd = 0;
        for (size_t i = 0; i < best->getPath()->size(); i++) pd = pd * 1000003ull + (*best->getPath())[i]->id;
        uint64_t c[9];
        search.counters(c);
        FILE* fp = fopen((dir + "/wr_result." + tag + "." + std::to_string(rank)).c_str(), "w");
        if (!fp) throw wr::Error(WR_ERR_FORMAT, "cannot write result");
        fprintf(fp, "%016llx %.9g %016llx %llu\n", digest(search.pheromone()), (double)best->L, pd, (unsigned long long)c[0]);
        fclose(fp);
    } catch (const wr::Error& err) {
        fprintf(stderr, "[sharded rank %d/%d] %s (status %d)\n", rank, ranks, err.what(), err.status);
        return 1;
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc >= 12 && std::string(argv[1]) == "--worker") return worker(argc, argv);
    if (argc < 3) { fprintf(stderr, "usage: %s <file.stl> <ranks> [precision] [wall] [ants] [iterations] [workdir]\n", argv[0]); return 2; }
    const int ranks = atoi(argv[2]);
    const std::string precision = argc > 3 ? argv[3] : "0.012", wall = argc > 4 ? argv[4] : "4", ants = argc > 5 ? argv[5] : "4096",
                      iterations = argc > 6 ? argv[6] : "12", dir = argc > 7 ? argv[7] : "/tmp";
    const std::string tag = std::to_string((long long)getpid());
    auto spawn = [&](int rank, int nranks, int device, const std::string& t) {
        pid_t pid = fork();
        if (pid == 0) {
            const std::string r = std::to_string(rank), n = std::to_string(nranks), d = std::to_string(device);
            execl(argv[0], argv[0], "--worker", r.c_str(), n.c_str(), d.c_str(), t.c_str(), argv[1], precision.c_str(), wall.c_str(), ants.c_str(), iterations.c_str(),
                  dir.c_str(), (char*)nullptr);
            _exit(127);
        }
        return pid;
    };
    auto wait_all = [](std::vector<pid_t>& pids) {
        bool ok = true;
        for (pid_t p : pids) { int st = 0; waitpid(p, &st, 0); ok = ok && WIFEXITED(st) && WEXITSTATUS(st) == 0; }
        return ok;
    };
    auto read_result = [&](const std::string& t, int rank) {
        std::string line;
        std::ifstream f(dir + "/wr_result." + t + "." + std::to_string(rank));
        std::getline(f, line);
        return line;
    };
    std::vector<pid_t> one = {spawn(0, 1, 0, tag + "s")};
    if (!wait_all(one)) { fprintf(stderr, "[sharded] the single-GPU run failed\n"); return 1; }
    const std::string want = read_result(tag + "s", 0);
    std::vector<pid_t> pids;
    for (int r = 0; r < ranks; r++) pids.push_back(spawn(r, ranks, r, tag));
    if (!wait_all(pids)) { fprintf(stderr, "[sharded] a rank failed\n"); return 1; }
    bool same = true;
    for (int r = 0; r < ranks; r++) {
        const std::string got = read_result(tag, r);
        // ant-steps (last field) are per rank; the field digest, best length and best path must agree
        same = same && got.substr(0, got.rfind(' ')) == want.substr(0, want.rfind(' '));
        printf("[sharded] rank %d/%d: %s\n", r, ranks, got.c_str());
    }
    printf("[sharded] 1 GPU     : %s\n[sharded] %s\n", want.c_str(), same ? "sharded == single GPU, bit for bit" : "MISMATCH");
    return same && !want.empty() ? 0 : 1;
}
