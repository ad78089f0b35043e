// include/scpp_b200.hpp — C++17 host mirror of the reference's algorithm classes on top of the C-ABI (scpp_b200.h).
//
// Same method names, argument meaning and error behaviour as scpp::SCAlgorithm (scpp_core/include/SCAlgorithm.hpp:9-45) and
// scpp::SCvxAlgorithm (scpp_core/include/SCvxAlgorithm.hpp): constructor from a parameter folder (the reference takes the model, whose
// getParameterFolder() names it, systemModel.hpp:147-155), initialize(), solve(bool warm_start), getSolution(), getAllSolutions();
// configuration / plugin misuse throws std::runtime_error as the reference does.  Batched: every method acts on all N instances; the
// trajectory type mirrors TrajectoryData (scpp_core/include/trajectoryData.hpp:8-32) with std::vector storage (Eigen is not required).
// Header-only; link with -lscpp_b200.  The ctypes mirror in scpp_b200/__init__.py is the Python equivalent.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "scpp_b200.h"

namespace scpp_b200 {

// TrajectoryData of ONE instance: X[k] (state_dim), U[k] (input_dim, first-order hold: K nodes), t
struct trajectory_data_t {
    std::vector<std::vector<double>> X, U;
    double t = 0.;
    size_t n_X() const { return X.size(); }
    size_t n_U() const { return U.size(); }
    bool interpolatedInput() const { return X.size() == U.size(); }          // trajectoryData.hpp:34-38
};

inline void check(int rc)
{
    if (rc) throw std::runtime_error(scpp_b200_last_error());
}

class SCAlgorithm {
public:
    // model_id: SCPP_B200_MODEL_*; folder holds model.info and SC.info (SCvx.info for the SCvx subclass)
    SCAlgorithm(int model_id, std::string parameter_folder, int n_instances = 1, int device = 0)
        : model(model_id), folder(std::move(parameter_folder)), N(n_instances), dev(device)
    {
        check(scpp_b200_model_dims(model, &nx, &nu, &np));
        if (N < 1) throw std::runtime_error("SCAlgorithm: n_instances must be positive");
    }
    virtual ~SCAlgorithm() { scpp_b200_destroy(engine); }
    SCAlgorithm(const SCAlgorithm &) = delete;
    SCAlgorithm &operator=(const SCAlgorithm &) = delete;

    virtual void loadParameters()                       // SCAlgorithm::loadParameters, SCAlgorithm.cpp:22-46
    {
        scpp_b200_default_config(model, &cfg);
        check(scpp_b200_load_sc_info((folder + "/SC.info").c_str(), &cfg));
    }
    void initialize()                                   // SCAlgorithm::initialize, SCAlgorithm.cpp:48-64
    {
        x_init.assign(nx, 0.); x_final.assign(nx, 0.);
        check(scpp_b200_load_model_info((folder + "/model.info").c_str(), model, &params, x_init.data(), x_final.data()));
        loadParameters();
        cfg.keep_history = 1;
        scpp_b200_destroy(engine); engine = nullptr;
        check(scpp_b200_create(model, &params, &cfg, N, dev, &engine));
        initialized = true;
    }
    // per-instance boundary states [N][nx] (the Monte-Carlo use, rocketQuat.cpp:203-227 / SC_sim.cpp:36); default: model.info for all
    void setBoundaryStates(const std::vector<double> &xi, const std::vector<double> &xf)
    {
        require_init();
        if (xi.size() != size_t(N) * nx || xf.size() != size_t(N) * nx) throw std::runtime_error("setBoundaryStates: need N * state_dim values");
        check(scpp_b200_set_boundary_states(engine, xi.data(), xf.data()));
        have_states = true;
    }
    void solve(bool warm_start = false)                 // SCAlgorithm::solve, SCAlgorithm.cpp:134-189
    {
        require_init();
        if (!have_states) {
            std::vector<double> a, b;
            for (int n = 0; n < N; n++) { a.insert(a.end(), x_init.begin(), x_init.end()); b.insert(b.end(), x_final.begin(), x_final.end()); }
            setBoundaryStates(a, b);
        }
        check(scpp_b200_solve(engine, warm_start ? 1 : 0));
        solved = true;
    }
    void getSolution(trajectory_data_t &td, int instance = 0) const      // SCAlgorithm::getSolution, :212-215 (redimensionalised)
    {
        require_solved(instance);
        const int K = cfg.K;
        std::vector<double> X(size_t(N) * K * nx), U(size_t(N) * K * nu), t(N);
        check(scpp_b200_get_solution(engine, X.data(), U.data(), t.data(), nullptr, nullptr));
        fill(td, X.data() + size_t(instance) * K * nx, U.data() + size_t(instance) * K * nu, t[instance]);
    }
    // every iterate of one instance, redimensionalised like the reference's (model->redimensionalizeTrajectory on a copy of every iterate,
    // SCAlgorithm.cpp:217-232)
    void getAllSolutions(std::vector<trajectory_data_t> &all, int instance = 0) const
    {
        require_solved(instance);
        const int K = cfg.K;
        std::vector<int> its(N);
        check(scpp_b200_get_solution(engine, nullptr, nullptr, nullptr, its.data(), nullptr));
        std::vector<double> X(size_t(N) * K * nx), U(size_t(N) * K * nu), t(N);
        all.clear();
        for (int it = 0; it <= its[instance]; it++) {
            check(scpp_b200_get_iterate_dimensional(engine, it, X.data(), U.data(), t.data()));
            all.emplace_back();
            fill(all.back(), X.data() + size_t(instance) * K * nx, U.data() + size_t(instance) * K * nu, t[instance]);
        }
    }
    // iterations done and flag (0 running, 1 converged, 2 failed, 4 iteration limit, 8 frozen) of every instance
    void getStatus(std::vector<int> &iterations, std::vector<int> &flags) const
    {
        require_solved(0);
        iterations.resize(N); flags.resize(N);
        check(scpp_b200_get_solution(engine, nullptr, nullptr, nullptr, iterations.data(), flags.data()));
    }
    // one step of the SC_sim closed loop (scpp/src/SC_sim.cpp:47-61) for every instance; returns how many instances are still flying
    int simulateStep(double time_step, std::vector<double> &x /* [N][nx] */, std::vector<double> &u0 /* [N][nu] */)
    {
        require_solved(0);
        x.resize(size_t(N) * nx); u0.resize(size_t(N) * nu);
        std::vector<int> reached(N);
        check(scpp_b200_sim_step(engine, time_step, x.data(), u0.data(), reached.data()));
        int flying = 0;
        for (int r : reached) flying += r == 0;
        return flying;
    }

    scpp_b200_sc_config cfg{};
    scpp_b200_model_params params{};
    int state_dim() const { return nx; }
    int input_dim() const { return nu; }

protected:
    void require_init() const { if (!initialized) throw std::runtime_error("SCAlgorithm: initialize() has not been called"); }
    void require_solved(int instance) const
    {
        require_init();
        if (!solved) throw std::runtime_error("SCAlgorithm: no solution yet");
        if (instance < 0 || instance >= N) throw std::runtime_error("SCAlgorithm: instance out of range");
    }
    void fill(trajectory_data_t &td, const double *X, const double *U, double t) const
    {
        const int K = cfg.K;
        const int KU = cfg.interpolate_input ? K : K - 1;      // trajectoryData.hpp:27-32: zero-order hold has K - 1 input columns (the engine's column K - 1 is a placeholder)
        td.X.assign(K, std::vector<double>(nx)); td.U.assign(KU, std::vector<double>(nu));
        for (int k = 0; k < K; k++) for (int i = 0; i < nx; i++) td.X[k][i] = X[size_t(k) * nx + i];
        for (int k = 0; k < KU; k++) for (int i = 0; i < nu; i++) td.U[k][i] = U[size_t(k) * nu + i];
        td.t = t;
    }
    int model, nx = 0, nu = 0, np = 0;
    std::string folder;
    int N, dev;
    bool initialized = false, have_states = false, solved = false;
    scpp_b200_engine *engine = nullptr;
    std::vector<double> x_init, x_final;
};

// scpp::SCvxAlgorithm (scpp_core/src/SCvxAlgorithm.cpp): same surface, parameters from SCvx.info
class SCvxAlgorithm : public SCAlgorithm {
public:
    using SCAlgorithm::SCAlgorithm;
    void loadParameters() override                      // SCvxAlgorithm::loadParameters, SCvxAlgorithm.cpp:23-44
    {
        scpp_b200_default_config(model, &cfg);
        check(scpp_b200_load_scvx_info((folder + "/SCvx.info").c_str(), &cfg));
    }
};

} // namespace scpp_b200
