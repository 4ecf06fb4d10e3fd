// tests/host_sim/mle_tps_sim.cpp -- TEST SCAFFOLDING, NOT PRODUCT CODE.
//
// Compiles the per-spot arithmetic of the thread-per-spot CUDA MLE path
// (picasso_b200/csrc/mle_tps_core.cuh, the very functions the kernels call) with g++ and
// runs it spot by spot on the CPU, so tests can compare it with the oracle without a GPU.
// glibc's exp/log stand in for libdevice's; everything else is the same code.
#include "../../picasso_b200/csrc/mle_tps_core.cuh"

namespace {

struct RoiPtr {
    const float* p;
    float operator()(int k) const { return p[k]; }
};
template <int BOX, typename T>
struct XfHost {
    T a[BOX][6];
    void put(int c, const double f[5]) {
        for (int k = 0; k < 5; k++) a[c][k] = (T)f[k];
        a[c][5] = (T)(f[0] - (double)a[c][0]);     // low word of PSFx (0 for T = double)
    }
    void put_f(int c, double psf, const float d[4]) {
        a[c][0] = (T)psf;
        for (int k = 0; k < 4; k++) a[c][k + 1] = (T)d[k];
        a[c][5] = (T)(psf - (double)a[c][0]);
    }
    void get(int c, T f[6]) const { for (int k = 0; k < 6; k++) f[k] = a[c][k]; }
};
template <int BOX>
struct Xf3Host {
    double a[BOX][3];
    void put(int c, const double f[5]) { a[c][0] = f[0]; a[c][1] = f[1]; a[c][2] = f[3]; }
    void get(int c, double f[3]) const { f[0] = a[c][0]; f[1] = a[c][1]; f[2] = a[c][2]; }
};

template <int BOX>
struct XfFHost {      // fast CRLB pass: PSF double, (d/dmu, d/dsigma) float
    double px[BOX];
    float c1[BOX], g1[BOX];
    void put_f(int c, double psf, const float d[4]) { px[c] = psf; c1[c] = d[0]; g1[c] = d[2]; }
    void get(int c, double& p, float& a, float& g) const { p = px[c]; a = c1[c]; g = g1[c]; }
};

template <int BOX, int METHOD, typename T, typename A>
void fit_range(const float* spots, long long n, double eps, int max_it, float* thetas, float* crlbs,
               float* logliks, int* iterations, int* status) {
    for (long long s = 0; s < n; s++) {
        RoiPtr roi{spots + s * BOX * BOX};
        float th[6], ms[6];
        int st = tps::initial_theta<BOX, METHOD>(roi, th);
        tps::max_steps(th, ms);
        int kk = 0;
        while (kk < max_it) {
            kk++;
            XfHost<BOX, T> xf;
            tps::column_stage<BOX, METHOD, T>(th, xf, tps::ErfTabDirect{});
            A num[6], den[6];
            tps::newton_sums<BOX, METHOD, T, A>(roi, th, xf, num, den, tps::ErfTabDirect{});
            if (tps::update_theta<BOX, METHOD, A>(th, ms, num, den, eps)) break;
        }
        Xf3Host<BOX> x3;
        float cr[6], ll;
        int cs = -1;
        if (sizeof(T) == 4) {      // the float32-pixel kernels use the fast CRLB pass with f64 fallback
            XfFHost<BOX> xff;
            cs = tps::crlb_loglik_fast<BOX, METHOD>(roi, th, xff, tps::LogTabDirect{}, tps::ErfTabDirect{}, cr, &ll);
        }
        if (cs < 0) cs = tps::crlb_loglik<BOX, METHOD>(roi, th, x3, cr, &ll);
        st |= cs;
        for (int l = 0; l < 6; l++) { thetas[s * 6 + l] = th[l]; crlbs[s * 6 + l] = cr[l]; }
        logliks[s] = ll;
        iterations[s] = kk;
        if (status) status[s] = st;
    }
}

template <int BOX>
int dispatch(const float* spots, long long n, double eps, int max_it, int method, int f32,
             float* th, float* cr, float* ll, int* it, int* st) {
    // f32: 0 = float64 pixel sums and row stage, 1 = float32 pixel sums / float64 row stage,
    //      2 = float32 pixel sums and row stage
    if (method == 1) {
        if (f32 == 2) fit_range<BOX, 1, float, float>(spots, n, eps, max_it, th, cr, ll, it, st);
        else if (f32) fit_range<BOX, 1, float, double>(spots, n, eps, max_it, th, cr, ll, it, st);
        else fit_range<BOX, 1, double, double>(spots, n, eps, max_it, th, cr, ll, it, st);
    } else {
        if (f32 == 2) fit_range<BOX, 0, float, float>(spots, n, eps, max_it, th, cr, ll, it, st);
        else if (f32) fit_range<BOX, 0, float, double>(spots, n, eps, max_it, th, cr, ll, it, st);
        else fit_range<BOX, 0, double, double>(spots, n, eps, max_it, th, cr, ll, it, st);
    }
    return 0;
}

}  // namespace

extern "C" int sim_mle_tps(const float* spots, long long n, int box, double eps, int max_it,
                           int method, int f32_pixels, float* thetas, float* crlbs, float* logliks,
                           int* iterations, int* status) {
    switch (box) {
        case 5: return dispatch<5>(spots, n, eps, max_it, method, f32_pixels, thetas, crlbs, logliks, iterations, status);
        case 7: return dispatch<7>(spots, n, eps, max_it, method, f32_pixels, thetas, crlbs, logliks, iterations, status);
        case 9: return dispatch<9>(spots, n, eps, max_it, method, f32_pixels, thetas, crlbs, logliks, iterations, status);
        case 11: return dispatch<11>(spots, n, eps, max_it, method, f32_pixels, thetas, crlbs, logliks, iterations, status);
        case 13: return dispatch<13>(spots, n, eps, max_it, method, f32_pixels, thetas, crlbs, logliks, iterations, status);
        default: return -1;
    }
}

// erf(z)/2 of the float32-pixel kernels (97-interval table), for the accuracy test
extern "C" void sim_half_erf(const double* z, long long n, double* out) {
    for (long long i = 0; i < n; i++) out[i] = tps::half_erf_tab(z[i], tps::ErfTabDirect{});
}

