// tests/cpp_mirror.cpp — compiles the C++ host mirror (include/scpp_b200.hpp) against libscpp_b200.so and exercises what can be
// exercised without a GPU: parameter loading, the reference's error behaviour, and the loud failure when there is no CUDA device.
// With a GPU (argv[2] == "gpu") it runs SC and SCvx for one instance.
#include <cstdio>
#include <cmath>
#include <cstring>
#include "scpp_b200.hpp"

int main(int argc, char **argv)
{
    const std::string cfgdir = argv[1];
    const bool gpu = argc > 2 && !strcmp(argv[2], "gpu");
    try {
        scpp_b200::SCAlgorithm a(SCPP_B200_MODEL_ROCKETQUAT, cfgdir + "/RocketQuat", 1);
        try { a.solve(); printf("FAIL: solve before initialize did not throw\n"); return 1; }
        catch (const std::runtime_error &) {}
        a.loadParameters();
        if (a.cfg.K != 15 || a.cfg.max_iterations != 15 || a.cfg.weight_virtual_control != 1000.) { printf("FAIL: SC.info values\n"); return 1; }
        scpp_b200::SCvxAlgorithm v(SCPP_B200_MODEL_ROCKETQUAT, cfgdir + "/RocketQuat", 1);
        v.loadParameters();
        if (v.cfg.algorithm != 1 || v.cfg.K != 30 || v.cfg.scvx_trust_region != 5.) { printf("FAIL: SCvx.info values\n"); return 1; }
        try { scpp_b200::SCAlgorithm bad(SCPP_B200_MODEL_ROCKETQUAT, cfgdir + "/does_not_exist", 1); bad.initialize(); printf("FAIL: missing folder did not throw\n"); return 1; }
        catch (const std::runtime_error &) {}
        if (!gpu) {
            try { a.initialize(); printf("FAIL: initialize without a CUDA device did not throw\n"); return 1; }
            catch (const std::runtime_error &e) { if (!strstr(e.what(), "no CUDA device")) { printf("FAIL: unexpected message: %s\n", e.what()); return 1; } }
            printf("ok (no GPU)\n");
            return 0;
        }
        a.initialize(); a.cfg.K = 15; a.solve();
        scpp_b200::trajectory_data_t td; a.getSolution(td);
        std::vector<scpp_b200::trajectory_data_t> all; a.getAllSolutions(all);
        // getAllSolutions redimensionalises every iterate (SCAlgorithm.cpp:217-232): the last one equals getSolution, the first starts at x_init
        for (size_t k = 0; k < td.n_X(); k++) for (int i = 0; i < a.state_dim(); i++) if (all.back().X[k][i] != td.X[k][i]) { printf("FAIL: last iterate != solution\n"); return 1; }
        if (std::fabs(all.front().X[0][0] - 24000.) > 1e-6) { printf("FAIL: iterates are not redimensionalised (m0 = %g)\n", all.front().X[0][0]); return 1; }
        v.initialize(); v.solve();
        std::vector<int> it, fl; v.getStatus(it, fl);
        printf("ok (GPU): SC K=%zu iterates=%zu t=%.3f ; SCvx iterations=%d flag=%d\n", td.n_X(), all.size(), td.t, it[0], fl[0]);
    } catch (const std::exception &e) { printf("FAIL: %s\n", e.what()); return 1; }
    return 0;
}
