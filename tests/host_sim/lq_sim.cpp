// tests/host_sim/lq_sim.cpp -- TEST SCAFFOLDING, NOT PRODUCT CODE.
//
// Compiles the per-spot optimiser of the CUDA least-squares fit (picasso_b200/csrc/lq_core.cuh,
// the very functions the kernel calls) with g++ and runs it spot by spot on the CPU, so tests can
// compare its lmdif trajectory (nfev, info, theta) with the oracle / the reference's golden
// vectors without a GPU.  glibc's exp stands in for libdevice's; everything else is the same code.
#include "../../picasso_b200/csrc/lq_core.cuh"

namespace {
template <int BOX>
void run(const float* spots, long long n, int variant, float* thetas, int* infos, int* nfevs) {
    for (long long s = 0; s < n; s++) {
        double x[6];
        int info, nfev;
        if (variant == 1) lq::fit_spot_qr<BOX>(spots + s * BOX * BOX, x, &info, &nfev);
        else lq::fit_spot_ne<BOX>(spots + s * BOX * BOX, x, &info, &nfev);
        for (int k = 0; k < 6; k++) thetas[s * 6 + k] = (float)x[k];
        infos[s] = info;
        nfevs[s] = nfev;
    }
}
}  // namespace

extern "C" int sim_lq(const float* spots, long long n, int box, int variant, float* thetas,
                      int* infos, int* nfevs) {
    switch (box) {
        case 5: run<5>(spots, n, variant, thetas, infos, nfevs); return 0;
        case 7: run<7>(spots, n, variant, thetas, infos, nfevs); return 0;
        case 9: run<9>(spots, n, variant, thetas, infos, nfevs); return 0;
        case 11: run<11>(spots, n, variant, thetas, infos, nfevs); return 0;
        case 13: run<13>(spots, n, variant, thetas, infos, nfevs); return 0;
        case 15: run<15>(spots, n, variant, thetas, infos, nfevs); return 0;
    }
    return 1;
}
