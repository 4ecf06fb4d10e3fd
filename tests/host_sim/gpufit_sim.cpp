// tests/host_sim/gpufit_sim.cpp -- TEST SCAFFOLDING, NOT PRODUCT CODE.
// Compiles the per-spot arithmetic of the Gpufit-path kernel (picasso_b200/csrc/gpufit_core.cuh) with
// g++ so the CPU suite can compare it with the independent C restatement (oracle/gpufit_oracle.c).
#include "../../picasso_b200/csrc/gpufit_core.cuh"

namespace {
struct Roi {
    const float* p;
    float operator()(int k) const { return p[k]; }
};
template <int BOX>
void run(const float* spots, long long n, float tol, int max_it, float* params, int* states, float* chi2, int* nit) {
    for (long long s = 0; s < n; s++) {
        Roi roi{spots + s * BOX * BOX};
        float p[6];
        gpufit::initial_parameters<BOX>(roi, p);
        gpufit::fit<BOX>(roi, p, tol, max_it, &states[s], &chi2[s], &nit[s]);
        const float twopi = (float)(2.0 * 3.141592653589793);
        params[s * 6] = p[0] * ((twopi * p[3]) * p[4]);
        for (int k = 1; k < 6; k++) params[s * 6 + k] = p[k];
    }
}
}  // namespace

extern "C" int sim_gpufit(const float* spots, long long n, int box, float tol, int max_it, float* params,
                          int* states, float* chi2, int* nit) {
    switch (box) {
        case 5: run<5>(spots, n, tol, max_it, params, states, chi2, nit); return 0;
        case 7: run<7>(spots, n, tol, max_it, params, states, chi2, nit); return 0;
        case 9: run<9>(spots, n, tol, max_it, params, states, chi2, nit); return 0;
        case 13: run<13>(spots, n, tol, max_it, params, states, chi2, nit); return 0;
    }
    return 1;
}
