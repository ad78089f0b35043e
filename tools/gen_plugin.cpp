// tools/gen_plugin.cpp — build-time generator for models written against the plugin surface (scpp_b200/plugins/*.hpp):
// records the model's addApplicationConstraints through the cvx:: shim, lowers it to the stage-wise table and writes
// scpp_b200/csrc/gen/<model>.inc (row table, cone dimensions, constant-slot recipe, pinned-variable lists).
// Built and run by scpp_b200/build.py before nvcc:  g++ -std=c++17 -Iinclude tools/gen_plugin.cpp -o gen && ./gen scpp_b200/csrc/gen
#define SCPP_PLUGIN_HOST
#define SCPP_PLUGIN_GENERATE
#include "scpp_plugin.hpp"
#include "../scpp_b200/csrc/models.cuh"
#include <fstream>
#include <iostream>

template <class M>
static int generate(const std::string &dir, const std::string &file, const std::string &NAME)
{
    const int K = 5;                                     // the structure is per node; any horizon >= 3 shows first / interior / last
    double constants[M::NCONST], x_init[M::NX], x_final[M::NX];
    for (int i = 0; i < M::NCONST; i++) constants[i] = 1.5 + i;     // values are irrelevant: the lowering traces addresses
    for (int i = 0; i < M::NX; i++) { x_init[i] = 10. + i; x_final[i] = 20. + i; }
    cvx::OptimizationProblem socp;
    socp.addVariable("X", M::NX, K);
    socp.addVariable("U", M::NU, K);
    std::vector<double> node_array(3 * K, 0.);
    scpp_plugin::Lowering L;
    if constexpr (M::uses_node_array) {
        M::addApplicationConstraints(socp, constants, x_init, x_final, node_array.data());
        L.node_array = {node_array.data(), 3 * K}; L.node_rows = 3;
    } else
        M::addApplicationConstraints(socp, constants, x_init, x_final);
    L.constants = {constants, M::NCONST}; L.x_init = {x_init, M::NX}; L.x_final = {x_final, M::NX};
    const scpp_plugin::StageTable t = L.lower(socp);
    const scpp_plugin::Emitted e = scpp_plugin::emit_inc(t, NAME, scpp::MAX_CST);
    const std::string path = dir + "/" + file;
    { std::ifstream in(path); std::string old((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>()); if (in && old == e.text) return 0; }   // unchanged: keep the time stamp
    std::ofstream out(path);
    out << e.text;
    std::cerr << "gen_plugin: wrote " << path << " (" << e.nlp << " LP rows, " << e.ncone << " cones / " << e.ncr << " rows, " << e.ncst << " constant slots)\n";
    return out ? 0 : 1;
}

int main(int argc, char **argv)
{
    const std::string dir = argc > 1 ? argv[1] : ".";
    try {
        int rc = generate<scpp::Rocket2dPlugin>(dir, "rocket2d_plugin.inc", "ROCKET2D_PLUGIN");
        if (!rc) rc = generate<scpp::RocketQuatRollPlugin>(dir, "rocketquat_roll_plugin.inc", "ROCKETQUAT_ROLL_PLUGIN");
        return rc;
    } catch (const std::exception &ex) {
        std::cerr << "gen_plugin: " << ex.what() << "\n";
        return 2;
    }
}
