// picasso_b200/csrc/mle_tps_core.cuh
//
// Per-spot arithmetic of the thread-per-spot MLE path (mle_tps.cu): start values,
// one Newton iteration, CRLB + log-likelihood.  Everything here is scalar code that
// one thread runs for one spot, written once for the device (nvcc, sm_100a) and for
// the host (g++): tests/host_sim compiles the same functions on the CPU so the
// arithmetic is checked against the oracle before a kernel ever runs.  The host
// build is test scaffolding only -- the product has no CPU path.
//
// Reference: picasso/gaussmle.py (jungmannlab/picasso @ 96e0da51):
//   _initial_parameters :28-168, _gaussian_integral :268-280,
//   _derivative_gaussian_integral :283-303, _G / _derivative_gaussian_integral_sigma
//   :306-336, _derivative_gaussian_integral_iso_sigma :339-383, _mlefit_sigma :533-644,
//   _update_theta_sigma :647-670, _mlefit_sigma_crlb :673-742, _mlefit_sigmaxy :745-857,
//   _update_theta_sigmaxy :860-884, _mlefit_sigmaxy_crlb :887-954.
//
// The model is separable: all transcendentals are evaluated per pixel EDGE
// (box+1 edges per axis: 1 exp + 1 erf polynomial each), every derivative is
// (column factor) x (row factor), and the Newton sums are accumulated as
// column-weighted row sums to which the row factors are applied once per row.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

// Under nvcc these are DEVICE functions (tables live in __constant__ memory, so the FP64
// instructions take their coefficients straight from the constant bank); under g++ they
// are plain inline functions with ordinary const tables.
#ifdef __CUDACC__
#define PB_HD __device__ __forceinline__
#define PB_HD_NOINLINE __device__ __noinline__
#define PB_TABLE static __constant__
#else
#define PB_HD inline
#define PB_HD_NOINLINE
#define PB_TABLE static const
#endif

#include "erf_table.cuh"

namespace tps {

constexpr double kInvSqrt2Pi = 0.3989422804014326779;   // 1/sqrt(2*pi)
constexpr double kInvSqrt2 = 0.70710678118654757;        // gaussmle.py:276
constexpr double kInvSqrtPi = 0.5641895835477562869;     // 1/sqrt(pi)

// ---- small math helpers ----------------------------------------------------
// 1/x for x in the normal range (MUFU seed + two Newton steps on the device).
PB_HD double rcp64(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    r = fma(r, fma(e, e, e), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
#else
    return 1.0 / x;
#endif
}
PB_HD float rcp32(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
PB_HD double rsqrt64(double x) {
#ifdef __CUDA_ARCH__
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
template <typename T> PB_HD T tfma(T a, T b, T c);
template <> PB_HD float tfma<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> PB_HD double tfma<double>(double a, double b, double c) { return fma(a, b, c); }
template <typename T> PB_HD T trcp(T x);
template <> PB_HD float trcp<float>(float x) { return rcp32(x); }
template <> PB_HD double trcp<double>(double x) { return rcp64(x); }

// exp(-q) for q >= 0 (tools/gen_exp_coeffs.py): n = rint(-q log2 e), r = -q - n ln2 (hi/lo),
// exp(r) by a degree-11 near-minimax polynomial (4e-18 relative before rounding) split into
// even and odd Horner chains, scaled by 2^n through the exponent field.  q is clamped to 700
// (exp(-700) ~ 1e-304 stands in for the underflowed tail), so the result stays normal and no
// special-case branch is needed.
PB_TABLE double kExpPoly[12] = {
    1.0, 1.0, 0.5000000000000019, 0.1666666666666668, 0.04166666666648738,
    0.008333333333319546, 0.0013888888952505406, 0.00019841269890193705,
    2.4801485278370062e-05, 2.7557240761739257e-06, 2.7632715071632255e-07,
    2.5110095611886237e-08,
};
PB_HD double exp_neg(double q) {
    q = q > 700.0 ? 700.0 : q;
    const double x = -q;
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);   // 1.5 * 2^52: n in the low word
    const double n = t - 6755399441055744.0;
    double r = fma(n, -0.6931471805598903, x);                          // ln2_hi (11 trailing zero bits)
    r = fma(n, -5.497923018708371e-14, r);                              // ln2_lo
    const double r2 = r * r;
    double pe = kExpPoly[10], po = kExpPoly[11];
#pragma unroll
    for (int k = 4; k >= 0; k--) {
        pe = fma(pe, r2, kExpPoly[2 * k]);
        po = fma(po, r2, kExpPoly[2 * k + 1]);
    }
    const double p = fma(po, r, pe);
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(p) + (__double2loint(t) << 20), __double2loint(p));
#else
    return ldexp(p, (int)n);
#endif
}

// erf(z) from the Gaussian term A = exp(-z^2) the fit needs anyway:
//   erf(z) = sign(z) (1 - A P(x)),  t = 1/(1 + |z|/2),  x = (8 t - 5)/3,
// P = degree-16 fit of erfcx on [0, 6] (tools/gen_erf_coeffs.py: max abs error 2.2e-15;
// beyond |z| = 6, A < 3e-16), evaluated as two Horner chains in x^2.
PB_TABLE double kErfcxPoly[17] = {
    0.3785374169292369,      0.4221875836134109,     0.16940759095488894,
    0.03299342947957943,     -0.0016702444148986467, -0.001665364388974453,
    0.00013358437327083263,  0.00010054202594586505, -2.2700690081718415e-05,
    -4.271484393949245e-06,  2.839667893118267e-06,  -2.8180849640529505e-07,
    -2.0096403079505821e-07, 8.619074557878024e-08,  -3.971810944694673e-09,
    -7.52808888456569e-09,   2.015803325273065e-09,
};
PB_HD double erf_from_gauss(double z, double A) {
    double a = fabs(z);
    a = a > 6.0 ? 6.0 : a;
    const double t = rcp64(fma(0.5, a, 1.0));
    const double x = fma(t, 2.6666666666666665, -1.6666666666666667);
    const double x2 = x * x;
    double pe = kErfcxPoly[16], po = kErfcxPoly[15];
    pe = fma(pe, x2, kErfcxPoly[14]);
#pragma unroll
    for (int k = 6; k >= 0; k--) {
        po = fma(po, x2, kErfcxPoly[2 * k + 1]);
        pe = fma(pe, x2, kErfcxPoly[2 * k]);
    }
    const double p = fma(po, x, pe);
    return copysign(fma(-A, p, 1.0), z);
}

// ln(x) for a positive, finite, normal x (|relative error| < 1e-15): x = m 2^e with
// m in [sqrt(1/2), sqrt(2)), ln m = 2 atanh(s), s = (m - 1)/(m + 1), nine odd terms.
// Coefficients come from the constant bank and there is no special-case branch (libdevice's
// log spends a third of its instructions on immediates and on denormal / inf / nan paths).
PB_TABLE double kLogPoly[9] = {
    2.0, 0.66666666666666663, 0.40000000000000002, 0.2857142857142857, 0.22222222222222221,
    0.18181818181818182, 0.15384615384615385, 0.13333333333333333, 0.11764705882352941,
};
PB_HD double log_pos(double x) {
#ifdef __CUDA_ARCH__
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                    // m in [1, 2)
    if (hi >= 0x3ff6a09f) { hi -= 0x00100000; e += 1; }     // m >= ~sqrt(2): halve
    const double m = __hiloint2double(hi, lo);
#else
    int e;
    double m = 2.0 * frexp(x, &e);
    e -= 1;
    if (m >= 1.4142141342163086) { m *= 0.5; e += 1; }     // same threshold: high word 0x3ff6a09f
#endif
    const double s = (m - 1.0) * rcp64(m + 1.0);
    const double s2 = s * s;
    double p = kLogPoly[8];
#pragma unroll
    for (int k = 7; k >= 0; k--) p = fma(p, s2, kLogPoly[k]);
    const double de = (double)e;
    return fma(de, 0.6931471805598903, fma(s, p, de * 5.497923018708371e-14));
}

// Table-driven ln(x) for a positive, finite, normal x (tools/gen_log_table.py; max abs error
// 1.8e-15 for x < 1e5):  x = m 2^e, m in [1, 2);  i = top 7 mantissa bits;  r = fma(m, rc_i, -1),
// |r| < 2^-8;  ln x = e ln2 + lc_i + log1p(r) with a degree-4 Taylor sum -- 7 FP64 instructions
// plus two table loads (log_pos above: 20 FP64 + a reciprocal).  `tab(i)` returns (rc_i, lc_i);
// on the device the table is staged in shared memory (the index diverges across lanes, which
// would serialise constant-bank reads).
PB_TABLE double kLogRc[128] = {
    0x1.fe01fe01fe020p-1, 0x1.fa11caa01fa12p-1, 0x1.f6310aca0dbb5p-1, 0x1.f25f644230ab5p-1,
    0x1.ee9c7f8458e02p-1, 0x1.eae807aba01ebp-1, 0x1.e741aa59750e4p-1, 0x1.e3a9179dc1a73p-1,
    0x1.e01e01e01e01ep-1, 0x1.dca01dca01dcap-1, 0x1.d92f2231e7f8ap-1, 0x1.d5cac807572b2p-1,
    0x1.d272ca3fc5b1ap-1, 0x1.cf26e5c44bfc6p-1, 0x1.cbe6d9601cbe7p-1, 0x1.c8b265afb8a42p-1,
    0x1.c5894d10d4986p-1, 0x1.c26b5392ea01cp-1, 0x1.bf583ee868d8bp-1, 0x1.bc4fd65883e7bp-1,
    0x1.b951e2b18ff23p-1, 0x1.b65e2e3beee05p-1, 0x1.b37484ad806cep-1, 0x1.b094b31d922a4p-1,
    0x1.adbe87f94905ep-1, 0x1.aaf1d2f87ebfdp-1, 0x1.a82e65130e159p-1, 0x1.a574107688a4ap-1,
    0x1.a2c2a87c51ca0p-1, 0x1.a01a01a01a01ap-1, 0x1.9d79f176b682dp-1, 0x1.9ae24ea5510dap-1,
    0x1.9852f0d8ec0ffp-1, 0x1.95cbb0be377aep-1, 0x1.934c67f9b2ce6p-1, 0x1.90d4f120190d5p-1,
    0x1.8e6527af1373fp-1, 0x1.8bfce8062ff3ap-1, 0x1.899c0f601899cp-1, 0x1.87427bcc092b9p-1,
    0x1.84f00c2780614p-1, 0x1.82a4a0182a4a0p-1, 0x1.8060180601806p-1, 0x1.7e225515a4f1dp-1,
    0x1.7beb3922e017cp-1, 0x1.79baa6bb6398bp-1, 0x1.77908119ac60dp-1, 0x1.756cac201756dp-1,
    0x1.734f0c541fe8dp-1, 0x1.713786d9c7c09p-1, 0x1.6f26016f26017p-1, 0x1.6d1a62681c861p-1,
    0x1.6b1490aa31a3dp-1, 0x1.691473a88d0c0p-1, 0x1.6719f3601671ap-1, 0x1.6524f853b4aa3p-1,
    0x1.63356b88ac0dep-1, 0x1.614b36831ae94p-1, 0x1.5f66434292dfcp-1, 0x1.5d867c3ece2a5p-1,
    0x1.5babcc647fa91p-1, 0x1.59d61f123ccaap-1, 0x1.5805601580560p-1, 0x1.56397ba7c52e2p-1,
    0x1.54725e6bb82fep-1, 0x1.52aff56a8054bp-1, 0x1.50f22e111c4c5p-1, 0x1.4f38f62dd4c9bp-1,
    0x1.4d843bedc2c4cp-1, 0x1.4bd3edda68fe1p-1, 0x1.4a27fad76014ap-1, 0x1.4880522014880p-1,
    0x1.46dce34596066p-1, 0x1.453d9e2c776cap-1, 0x1.43a2730abee4dp-1, 0x1.420b5265e5951p-1,
    0x1.40782d10e6566p-1, 0x1.3ee8f42a5af07p-1, 0x1.3d5d991aa75c6p-1, 0x1.3bd60d9232955p-1,
    0x1.3a524387ac822p-1, 0x1.38d22d366088ep-1, 0x1.3755bd1c945eep-1, 0x1.35dce5f9f2af8p-1,
    0x1.34679ace01346p-1, 0x1.32f5ced6a1dfap-1, 0x1.3187758e9ebb6p-1, 0x1.301c82ac40260p-1,
    0x1.2eb4ea1fed14bp-1, 0x1.2d50a012d50a0p-1, 0x1.2bef98e5a3711p-1, 0x1.2a91c92f3c105p-1,
    0x1.293725bb804a5p-1, 0x1.27dfa38a1ce4dp-1, 0x1.268b37cd60127p-1, 0x1.2539d7e9177b2p-1,
    0x1.23eb79717605bp-1, 0x1.22a0122a0122ap-1, 0x1.21579804855e6p-1, 0x1.2012012012012p-1,
    0x1.1ecf43c7fb84cp-1, 0x1.1d8f5672e4abdp-1, 0x1.1c522fc1ce059p-1, 0x1.1b17c67f2bae3p-1,
    0x1.19e0119e0119ep-1, 0x1.18ab083902bdbp-1, 0x1.1778a191bd684p-1, 0x1.1648d50fc3201p-1,
    0x1.151b9a3fdd5c9p-1, 0x1.13f0e8d344724p-1, 0x1.12c8b89edc0acp-1, 0x1.11a3019a74826p-1,
    0x1.107fbbe011080p-1, 0x1.0f5edfab325a2p-1, 0x1.0e40655826011p-1, 0x1.0d24456359e3ap-1,
    0x1.0c0a7868b4171p-1, 0x1.0af2f722eecb5p-1, 0x1.09ddba6af8360p-1, 0x1.08cabb37565e2p-1,
    0x1.07b9f29b8eae2p-1, 0x1.06ab59c7912fbp-1, 0x1.059eea0727586p-1, 0x1.04949cc1664c5p-1,
    0x1.038c6b78247fcp-1, 0x1.02864fc7729e9p-1, 0x1.0182436517a37p-1, 0x1.0080402010080p-1,
};
PB_TABLE double kLogLc[128] = {
    0x1.ff00aa2b10ba0p-9, 0x1.7dc475f810a69p-7, 0x1.3cea44346a584p-6, 0x1.b9fc027af919ap-6,
    0x1.1b0d98923d97fp-5, 0x1.58a5bafc8e4d3p-5, 0x1.95c830ec8e3f2p-5, 0x1.d276b8adb0b56p-5,
    0x1.075983598e471p-4, 0x1.253f62f0a1417p-4, 0x1.42edcbea646eep-4, 0x1.60658a93750c4p-4,
    0x1.7da766d7b12d0p-4, 0x1.9ab42462033aep-4, 0x1.b78c82bb0eda0p-4, 0x1.d4313d66cb35dp-4,
    0x1.f0a30c01162a4p-4, 0x1.0671512ca596fp-3, 0x1.14785846742acp-3, 0x1.2266f190a5acdp-3,
    0x1.303d718e47fd5p-3, 0x1.3dfc2b0ecc62ap-3, 0x1.4ba36f39a55e5p-3, 0x1.59338d9982085p-3,
    0x1.66acd4272ad51p-3, 0x1.740f8f54037a3p-3, 0x1.815c0a14357e9p-3, 0x1.8e928de886d41p-3,
    0x1.9bb362e7dfb85p-3, 0x1.a8becfc882f19p-3, 0x1.b5b519e8fb5a6p-3, 0x1.c2968558c18c2p-3,
    0x1.cf6354e09c5ddp-3, 0x1.dc1bca0abec7bp-3, 0x1.e8c0252aa5a60p-3, 0x1.f550a564b7b37p-3,
    0x1.00e6c45ad501dp-2, 0x1.071b85fcd590dp-2, 0x1.0d46b579ab74bp-2, 0x1.136870293a8b0p-2,
    0x1.1980d2dd4236fp-2, 0x1.1f8ff9e48a2f3p-2, 0x1.2596010df763ap-2, 0x1.2b9303ab89d25p-2,
    0x1.31871c9544185p-2, 0x1.3772662bfd85cp-2, 0x1.3d54fa5c1f710p-2, 0x1.432ef2a04e813p-2,
    0x1.49006804009d0p-2, 0x1.4ec9732600269p-2, 0x1.548a2c3add263p-2, 0x1.5a42ab0f4cfe2p-2,
    0x1.5ff3070a793d4p-2, 0x1.659b57303e1f2p-2, 0x1.6b3bb2235943dp-2, 0x1.70d42e2789236p-2,
    0x1.7664e1239dbcfp-2, 0x1.7bede0a37afbfp-2, 0x1.816f41da0d495p-2, 0x1.86e919a330ba1p-2,
    0x1.8c5b7c858b48bp-2, 0x1.91c67eb45a83ep-2, 0x1.972a341135159p-2, 0x1.9c86b02dc0862p-2,
    0x1.a1dc064d5b995p-2, 0x1.a72a4966bd9e9p-2, 0x1.ac718c258b0e5p-2, 0x1.b1b1e0ebdfc5ap-2,
    0x1.b6eb59d3cf35cp-2, 0x1.bc1e08b0dad0ap-2, 0x1.c149ff115f027p-2, 0x1.c66f4e3ff6ff9p-2,
    0x1.cb8e0744d7acap-2, 0x1.d0a63ae721e64p-2, 0x1.d5b7f9ae2c684p-2, 0x1.dac353e2c5955p-2,
    0x1.dfc859906d5b5p-2, 0x1.e4c71a8687704p-2, 0x1.e9bfa659861f5p-2, 0x1.eeb20c640ddf3p-2,
    0x1.f39e5bc811e5dp-2, 0x1.f884a36fe9ec1p-2, 0x1.fd64f20f61571p-2, 0x1.011fab125ff8ap-1,
    0x1.0389eefce633cp-1, 0x1.05f14bd26459cp-1, 0x1.0855c884b450ep-1, 0x1.0ab76bece14d2p-1,
    0x1.0d163ccb9d6b8p-1, 0x1.0f7241c9b497dp-1, 0x1.11cb81787ccf8p-1, 0x1.1422025243d45p-1,
    0x1.1675cababa60ep-1, 0x1.18c6e0ff5cf07p-1, 0x1.1b154b57da29ep-1, 0x1.1d610fe677003p-1,
    0x1.1faa34b87094cp-1, 0x1.21f0bfc65beecp-1, 0x1.2434b6f483934p-1, 0x1.26762013430e0p-1,
    0x1.28b500df60783p-1, 0x1.2af15f02640acp-1, 0x1.2d2b4012edc9dp-1, 0x1.2f62a99509546p-1,
    0x1.3197a0fa7fe6ap-1, 0x1.33ca2ba328994p-1, 0x1.35fa4edd36ea0p-1, 0x1.38280fe58797fp-1,
    0x1.3a5373e7ebdf9p-1, 0x1.3c7c7fff73206p-1, 0x1.3ea33936b2f5bp-1, 0x1.40c7a4880dceap-1,
    0x1.42e9c6ddf80bfp-1, 0x1.4509a5133bb0ap-1, 0x1.472743f33aaadp-1, 0x1.4942a83a2fc07p-1,
    0x1.4b5bd6956e273p-1, 0x1.4d72d3a39fd01p-1, 0x1.4f87a3f5026e9p-1, 0x1.519a4c0ba3446p-1,
    0x1.53aad05b99b7cp-1, 0x1.55b9354b40bcep-1, 0x1.57c57f336f191p-1, 0x1.59cfb25fae87fp-1,
    0x1.5bd7d30e71c73p-1, 0x1.5ddde57149923p-1, 0x1.5fe1edad18919p-1, 0x1.61e3efda46467p-1,
};
struct LogTabDirect {     // host build / staging source
    PB_HD void get(int i, double& rc, double& lc) const { rc = kLogRc[i]; lc = kLogLc[i]; }
};
template <class Tab>
PB_HD double log_tab(double x, const Tab& tab) {
#ifdef __CUDA_ARCH__
    const int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const int idx = (hi >> 13) & 127;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
#else
    uint64_t bits;
    memcpy(&bits, &x, 8);
    const int e = (int)((bits >> 52) & 0x7ff) - 1023;
    const int idx = (int)((bits >> 45) & 127);
    const uint64_t mb = (bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double m;
    memcpy(&m, &mb, 8);
#endif
    double rc, lc;
    tab.get(idx, rc, lc);
    const double r = fma(m, rc, -1.0);
    // log1p(r) = r - r^2/2 + r^3/3 - ...; |r| < 2^-8: the degree-4 truncation error is below 2e-13
    double p = fma(r, -0.25, 0.33333333333333331);
    p = fma(p, r, -0.5);
    p = fma(p, r, 1.0);
    return fma((double)e, 0.6931471805599453, fma(p, r, lc));
}

// ---- per-axis constants of one iteration ------------------------------------
// Reciprocals of sigma, f32(sigma^2), f32(sigma^3), f32(sigma^5): `float32 ** int`
// stays float32 in the reference (binary powering), so the powers are rounded first.
struct Axis {
    double rs, r2, r3, r5;   // 1/sigma, 1/f32(sigma^2), 1/f32(sigma^3), 1/f32(sigma^5)
    double rho;              // sigma^2 / f32(sigma^2) - 1   (for _G's exponent, :306-316)
    double rinvf;            // (double) f32(1/sigma)        (iso-sigma d2, :380-382)
};
PB_HD Axis make_axis(float sig) {
    Axis a;
    const float s2f = sig * sig;
    const float s3f = sig * s2f;
    const float s5f = sig * (s2f * s2f);
    const double sd = (double)sig;
    a.rs = rcp64(sd);
    a.r2 = rcp64((double)s2f);
    a.r3 = rcp64((double)s3f);
    a.r5 = rcp64((double)s5f);
    a.rho = fma(sd * sd, a.r2, -1.0);
    a.rinvf = (double)(1.0f / sig);
    return a;
}

// Everything the two pixels sharing an edge need from it.  Edge g of an axis sits at
// e = (g - mu) - 1/2: the minus edge of pixel g and (exactly, in f64) the plus edge of g-1.
template <int METHOD>
struct Edge {
    double E;    // erf(e / (sqrt2 sigma))
    double A;    // exp(-e^2 / (2 sigma^2))
    double eA;   // e * A
    double u;    // METHOD 1: e * AG  (AG = exp(-e^2 / (2 f32(sigma^2))));  METHOD 0: am * A
    double v;    // METHOD 1: e^3 * AG;                                     METHOD 0: am A (1 - 2 am^2)
};
template <int METHOD>
PB_HD Edge<METHOD> eval_edge(int g, float mu, const Axis& ax) {
    Edge<METHOD> r;
    const double e = ((double)g - (double)mu) - 0.5;
    const double t = e * ax.rs;
    const double q = 0.5 * t * t;
    const double A = exp_neg(q);
    r.A = A;
    r.E = erf_from_gauss(e * (kInvSqrt2 * ax.rs), A);
    r.eA = e * A;
    if (METHOD == 1) {
        const double z = -q * ax.rho;
        const double AG = fma(A, fma(0.5 * z, z, z), A);
        r.u = e * AG;
        r.v = e * e * r.u;
    } else {
        const double am = e * (ax.rs * kInvSqrt2);
        r.u = am * A;
        r.v = r.u * (1.0 - 2.0 * am * am);
    }
    return r;
}

// Factors of one pixel column (or row) from its two edges:
//   f[0] = PSF (integrated Gaussian), f[1], f[2] = d/dmu, d2/dmu2 (per photon, per other-axis PSF),
//   f[3], f[4] = d/dsigma, d2/dsigma2 parts.
template <int METHOD>
PB_HD void pixel_factors(const Edge<METHOD>& lo, const Edge<METHOD>& hi, const Axis& ax, double f[5]) {
    f[0] = 0.5 * (hi.E - lo.E);
    f[1] = (lo.A - hi.A) * ax.rs * kInvSqrt2Pi;
    f[2] = (lo.eA - hi.eA) * ax.r3 * kInvSqrt2Pi;
    if (METHOD == 1) {
        const double w1 = lo.u - hi.u, w3 = lo.v - hi.v;
        f[3] = w1 * ax.r2 * kInvSqrt2Pi;
        f[4] = (w3 * ax.r5 - 2.0 * w1 * ax.r3) * kInvSqrt2Pi;
    } else {
        const double F = lo.u - hi.u;
        const double dF = (hi.v - lo.v) * ax.rs;
        f[3] = F * ax.rs * kInvSqrtPi;
        f[4] = kInvSqrtPi * (-F * ax.r2 + ax.rinvf * dF);
    }
}

// ---- float32-pixel kernels: table-driven erf, float32 derivative factors -------------------------
// Only the PSF (a difference of two erf values that is multiplied into the model and cancels against
// the data) needs more than float32, and it needs ~1e-9 absolute (measured: rounding erf to 1e-9 leaves
// the iteration counts of 200 000 spots unchanged; 1e-7 still passes the parity bar).  So
//   * erf(z)/2 comes from a 97-interval table (erf_table.cuh, tools/gen_erf_table.py): k = rint(16 |z|),
//     d = |z| - k/16, degree 5 with float64 c0, c1 and a float32 tail -- 5 FP64 instructions and two
//     16-byte table loads instead of a reciprocal plus a degree-16 float64 polynomial; max abs error 1.2e-11;
//   * the Gaussian edge term and everything derived from it (d/dmu, d2/dmu2, d/dsigma, d2/dsigma2) run in
//     float32 with the hardware exp2: the reference stores dudt / d2udt2 in float32 arrays anyway
//     (gaussmle.py:776-779) and accumulates their products in float32.
// Per edge this leaves 7 FP64 instructions (FP64 issues at half rate on sm_100) of the 87 the all-float64
// evaluation needs.
struct ErfTabDirect {     // host build / staging source
    PB_HD void get(int k, double& c0, double& c1, float& c2, float& c3, float& c4, float& c5) const {
        c0 = kErfTabA[k][0]; c1 = kErfTabA[k][1];
        c2 = kErfTabB[k][0]; c3 = kErfTabB[k][1]; c4 = kErfTabB[k][2]; c5 = kErfTabB[k][3];
    }
};
// erf(z) / 2;  NaN propagates, |z| >= 6 (and inf) gives +-0.5
template <class Tab>
PB_HD double half_erf_tab(double z, const Tab& tab) {
#ifdef __CUDA_ARCH__
    // |z| >= 6 (and inf) is resolved by a select at the END, off the dependent chain
    // |z| -> interval -> table -> polynomial; NaN fails the test and propagates through d
    const unsigned ahi = (unsigned)__double2hiint(z) & 0x7fffffffu;
    const bool big = (ahi - 0x40180000u) <= (0x7ff00000u - 0x40180000u);
    const double a = fabs(z);
    const double t = fma(a, 16.0, 6755399441055744.0);      // 1.5 * 2^52: rint(16 a) in the low word
    unsigned k = (unsigned)__double2loint(t);
#else
    const double a = fabs(z);
    const bool big = a >= 6.0;
    const double t = fma(a, 16.0, 6755399441055744.0);
    uint64_t bits;
    memcpy(&bits, &t, 8);
    unsigned k = (unsigned)(bits & 0xffffffffu);
#endif
    k = k > 96u ? 96u : k;                                   // (huge / NaN: any in-range entry)
    const double n = t - 6755399441055744.0;
    const double d = fma(n, -0.0625, a);
    double c0, c1;
    float c2, c3, c4, c5;
    tab.get((int)k, c0, c1, c2, c3, c4, c5);
    const float df = (float)d;
    float pf = fmaf(c5, df, c4);
    pf = fmaf(pf, df, c3);
    pf = fmaf(pf, df, c2);
    double r = fma(fma((double)pf, d, c1), d, c0);
    r = big ? 0.5 : r;
    return copysign(r, z);
}
// exp(-q), q >= 0, float32 (device: MUFU.EX2, 2 ulp; underflows to 0)
PB_HD float exp_neg_f(float q) {
#ifdef __CUDA_ARCH__
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q * -1.4426950408889634f));
    return r;
#else
    return exp2f(q * -1.4426950408889634f);
#endif
}
struct AxisF {
    double c;                  // 1 / (sqrt2 sigma): the erf argument scale
    float rs;                  // 1 / sigma
    float rho;                 // sigma^2 / f32(sigma^2) - 1   (_G's exponent, :306-316)
    float k1, k2, k3, k5, k32; // METHOD 1: rs c, r3 c, r2 c, r5 c, 2 r3 c   (c = 1/sqrt(2 pi))
    float m0, m3, m4a, m4b;    // METHOD 0: rs/sqrt2, rs/sqrt(pi), r2/sqrt(pi), rs f32(1/sigma)/sqrt(pi)
};
template <int METHOD>
PB_HD AxisF make_axis_f(float sig) {
    AxisF x;
    const float s2f = sig * sig;
    const float s3f = sig * s2f;
    const float s5f = sig * (s2f * s2f);
    const double rsd = rcp64((double)sig);
    x.c = rsd * kInvSqrt2;
    x.rs = (float)rsd;
    // (IEEE divisions: per axis, not per edge -- the approximate reciprocal's ulp would be noise on
    // every derivative factor of the axis)
    const float r2 = 1.0f / s2f, r3 = 1.0f / s3f;
    x.rho = fmaf(sig, sig, -s2f) * r2;                       // the product's rounding error is exact in fmaf
    x.k1 = x.rs * (float)kInvSqrt2Pi;
    x.k2 = r3 * (float)kInvSqrt2Pi;
    if (METHOD == 1) {
        x.k3 = r2 * (float)kInvSqrt2Pi;
        x.k5 = (1.0f / s5f) * (float)kInvSqrt2Pi;
        x.k32 = 2.0f * x.k2;
        x.m0 = x.m3 = x.m4a = x.m4b = 0.f;
    } else {
        x.k3 = x.k5 = x.k32 = 0.f;
        x.m0 = x.rs * (float)kInvSqrt2;
        x.m3 = x.rs * (float)kInvSqrtPi;
        x.m4a = r2 * (float)kInvSqrtPi;
        x.m4b = x.m3 * (1.0f / sig);
    }
    return x;
}
template <int METHOD>
struct EdgeF {
    double E;          // erf(e / (sqrt2 sigma)) / 2, float64: the PSF is a difference of two of these
    float A, eA, u, v; // as in Edge, float32
};
// e = (g - mu) - 1/2 of edge g in float64 (for the erf argument) and, formed independently, in float32
// (for the Gaussian term: no conversion on the way)
template <int METHOD, class Tab>
PB_HD EdgeF<METHOD> eval_edge_f(double e, float ef, const AxisF& ax, const Tab& tab) {
    EdgeF<METHOD> r;
    r.E = half_erf_tab(e * ax.c, tab);
    const float t = ef * ax.rs;
    const float q = 0.5f * t * t;
    const float Af = exp_neg_f(q);
    r.A = Af;
    r.eA = ef * Af;
    if (METHOD == 1) {
        const float z = -q * ax.rho;
        const float AG = fmaf(Af, fmaf(0.5f * z, z, z), Af);
        r.u = ef * AG;
        r.v = ef * ef * r.u;
    } else {
        const float am = ef * ax.m0;
        r.u = am * Af;
        r.v = r.u * (1.0f - 2.0f * am * am);
    }
    return r;
}
// psf = PSF (float64); d[0..3] = d/dmu, d2/dmu2, d/dsigma, d2/dsigma2 parts (float32)
template <int METHOD>
PB_HD void pixel_factors_f(const EdgeF<METHOD>& lo, const EdgeF<METHOD>& hi, const AxisF& ax, double& psf,
                           float d[4]) {
    psf = hi.E - lo.E;
    d[0] = (lo.A - hi.A) * ax.k1;
    d[1] = (lo.eA - hi.eA) * ax.k2;
    if (METHOD == 1) {
        const float w1 = lo.u - hi.u, w3 = lo.v - hi.v;
        d[2] = w1 * ax.k3;
        d[3] = fmaf(w3, ax.k5, -(w1 * ax.k32));
    } else {
        const float F = lo.u - hi.u;
        d[2] = F * ax.m3;
        d[3] = fmaf(hi.v - lo.v, ax.m4b, -(F * ax.m4a));
    }
}

// ---- start values (gaussmle.py:28-168) ---------------------------------------
// roi(p) returns pixel p = row * BOX + col as float.  Returns status flags
// (bit 0: a centre-row/column sum was exactly 0 -- the reference raises there).
template <int BOX, int METHOD, class Roi>
PB_HD int initial_theta(const Roi& roi, float th[6]) {
    constexpr int H = BOX / 2;
    int st = 0;
    // sum and centre of mass, f64, row-major like the reference (:28-48)
    double s = 0.0, ys = 0.0, xs = 0.0;
    // 3x3 edge-truncated mean filter minimum (:61-91, :135).  f32(S / N) is monotone in S,
    // so the minimum is taken over the window sums per window size and divided once.
    double m4 = INFINITY, m6 = INFINITY, m9 = INFINITY;
    double h[3][BOX];   // horizontal 3-sums of the last three rows (rotating, fully unrolled)
#pragma unroll
    for (int m = 0; m <= BOX; m++) {
        if (m < BOX) {
            float v[BOX];
#pragma unroll
            for (int j = 0; j < BOX; j++) {
                v[j] = roi(m * BOX + j);
                const double d = (double)v[j];
                ys += d * (double)m;
                xs += d * (double)j;
                s += d;
            }
#pragma unroll
            for (int j = 0; j < BOX; j++) {
                double a = (double)v[j];
                if (j > 0) a = (double)v[j - 1] + a;
                if (j + 1 < BOX) a += (double)v[j + 1];
                h[m % 3][j] = a;
            }
        }
        // window sums of row k = m - 1 are complete once row m is in
        const int k = m - 1;
        if (k >= 0) {
#pragma unroll
            for (int j = 0; j < BOX; j++) {
                double S = h[k % 3][j];
                if (k > 0) S = h[(k - 1) % 3][j] + S;
                if (k + 1 < BOX) S += h[(k + 1) % 3][j];
                const bool er = (k == 0 || k == BOX - 1), ec = (j == 0 || j == BOX - 1);
                if (er && ec) m4 = S < m4 ? S : m4;
                else if (er || ec) m6 = S < m6 ? S : m6;
                else m9 = S < m9 ? S : m9;
            }
        }
    }
    double xc, yc;
    if (s <= 0.0) { s = 0.01; yc = (BOX - 1) / 2.0; xc = (BOX - 1) / 2.0; }
    else { yc = ys / s; xc = xs / s; }
    const float b4 = (float)(m4 / 4.0), b6 = (float)(m6 / 6.0), b9 = (float)(m9 / 9.0);
    float bg = b4 < b6 ? b4 : b6;
    bg = b9 < bg ? b9 : bg;
    double ph = s - (double)(BOX * BOX) * (double)bg;
    ph = ph > 1.0 ? ph : 1.0;
    // sigmas from the centre column / row of (spot - bg) (:94-124)
    double sdy = 0.0, sy_ = 0.0, sdx = 0.0, sx_ = 0.0;
#pragma unroll
    for (int i = 0; i < BOX; i++) {
        const float vy = roi(i * BOX + H) - bg;
        const float vx = roi(H * BOX + i) - bg;
        const double d2 = (double)((i - H) * (i - H));
        sdy += (double)vy * d2; sdx += (double)vx * d2;
        sy_ += (double)vy;      sx_ += (double)vx;
    }
    double sy0, sx0;
    if (sy_ == 0.0) { sy0 = 0.01; st |= 1; } else sy0 = sqrt(sdy / sy_);
    if (sx_ == 0.0) { sx0 = 0.01; st |= 1; } else sx0 = sqrt(sdx / sx_);
    if (!isfinite(sy0) || sy0 == 0.0) sy0 = 0.01;
    if (!isfinite(sx0) || sx0 == 0.0) sx0 = 0.01;
    th[0] = (float)xc; th[1] = (float)yc; th[2] = (float)ph; th[3] = bg;
    if (METHOD == 1) { th[4] = (float)sx0; th[5] = (float)sy0; }
    else { th[4] = (float)((sx0 + sy0) / 2.0); th[5] = th[4]; }
    return st;
}

// max_step from the start values (gaussmle.py:558-561, 770-773)
PB_HD void max_steps(const float th0[6], float ms[6]) {
    ms[0] = th0[4]; ms[1] = th0[4];
    ms[2] = (float)(0.1 * (double)th0[2]);
    ms[3] = (float)(0.1 * (double)th0[3]);
    ms[4] = (float)(0.2 * (double)th0[4]);
    ms[5] = (float)(0.2 * (double)th0[5]);
}

// ---- one Newton iteration ------------------------------------------------------
// Column factors of the x axis are written to `xf` (T = float or double; shared memory on
// the device) by stage 1 and read back per pixel by stage 2; the y factors are formed on
// the fly, one row at a time.  Xf concept: void put(int col, const double f[5]);
// void get(int col, T f[5]) const.
template <int BOX, int METHOD, typename T, class Xf, class Tab>
PB_HD void column_stage(const float th[6], Xf& xf, const Tab& tab) {
    if constexpr (sizeof(T) == 4) {
        const AxisF ax = make_axis_f<METHOD>(th[4]);
        EdgeF<METHOD> A, B = {};
        double e = -(double)th[0] - 0.5;      // edge 0; edges are 1 apart (exact in float64)
        float gf = 0.0f;                  // edge index as float32 (exact)
#pragma unroll 1
        for (int k = 0; k < (BOX + 1) / 2; k++) {
            double psf;
            float d[4];
            A = eval_edge_f<METHOD>(e, (gf - th[0]) - 0.5f, ax, tab);
            if (k > 0) {
                pixel_factors_f<METHOD>(B, A, ax, psf, d);
                xf.put_f(2 * k - 1, psf, d);
            }
            B = eval_edge_f<METHOD>(e + 1.0, (gf - th[0]) + 0.5f, ax, tab);
            pixel_factors_f<METHOD>(A, B, ax, psf, d);
            xf.put_f(2 * k, psf, d);
            e += 2.0;
            gf += 2.0f;
        }
        return;
    }
    const Axis ax = make_axis(th[4]);
    // edges are evaluated in pairs into two fixed register sets (A: even edges, B: odd edges):
    // no loop-carried copies, and the two evaluations of a trip are independent (ILP)
    Edge<METHOD> A, B = {};
#pragma unroll 1
    for (int k = 0; k < (BOX + 1) / 2; k++) {
        double f[5];
        A = eval_edge<METHOD>(2 * k, th[0], ax);
        if (k > 0) {
            pixel_factors<METHOD>(B, A, ax, f);
            xf.put(2 * k - 1, f);
        }
        B = eval_edge<METHOD>(2 * k + 1, th[0], ax);
        pixel_factors<METHOD>(A, B, ax, f);
        xf.put(2 * k, f);
    }
}

// One pixel row: column-weighted sums of cf = data/model - 1 and df = data/model^2 over the
// row in type T, then the row factors fy applied and accumulated over rows in type A
// (A = double: f64 row stage; A = float: everything after the edge terms runs on the FP32
// pipe -- the reference itself accumulates all b*b terms in float32, gaussmle.py:836-839).
template <int BOX, int METHOD, typename T, typename A, typename FY, class Roi, class Xf>
PB_HD void accumulate_row(int j, double psf_y, const FY fyd[4], const Roi& roi, const Xf& xf, float Nf, float bgf,
                          A num[6], A den[6]) {
    // fy[0] = PSF of the row (float64), fy[1..4] = its derivative factors (FY = double or float)
    const double fy0 = psf_y;
    const A N = (A)Nf;
    const A PSFy = (A)fy0;
    const A NPy = N * PSFy;
    const T NPy_t = (T)NPy, bg_t = (T)bgf;
    // float32 pixels: the residual data - model cancels to ~sqrt(model), so it is formed in
    // float-float arithmetic (N PSFy and PSFx carry a low word); everything downstream of the
    // residual is well conditioned and stays plain float32
    const float NPy_lo = (sizeof(T) == 4) ? (float)((double)Nf * fy0 - (double)(float)NPy_t) : 0.f;
    T c0 = 0, cpx = 0, cc1 = 0, cc2 = 0, cg1 = 0, cg2 = 0;
    T d0 = 0, dpx2 = 0, dc1 = 0, dg1 = 0, dgp = 0;
#pragma unroll
    for (int i = 0; i < BOX; i++) {
        T f[6];
        xf.get(i, f);
        const T data = (T)roi(j * BOX + i);
        T model, inv, cf, df;
        if (sizeof(T) == 4) {
            const float px = (float)f[0], NPh = (float)NPy_t;
            const float p = px * NPh;                                   // hi part of N PSFx PSFy
            // rounding error of p plus the low-word cross terms, three fused steps
            float pc = fmaf(px, NPy_lo, fmaf(px, NPh, -p));
            pc = fmaf((float)f[5], NPh, pc);
            model = (T)(p + (float)bg_t);
            const float a = (float)data - p;          // exact when data and p are within 2x
            const float r = (a - (float)bg_t) - pc;
            // guard (gaussmle.py:829-835): model > 0.01, else cf = df = 0 -- folded into 1/model
            const float iv = model > (T)10e-3 ? rcp32((float)model) : 0.0f;
            inv = (T)iv;
            cf = (T)(r * iv);
            df = data * inv * inv;
            cf = cf > (T)10e4 ? (T)10e4 : cf;          // cf, df <= 1e5
            df = df > (T)10e4 ? (T)10e4 : df;
        } else {
            model = tfma<T>(f[0], NPy_t, bg_t);
            // guards (gaussmle.py:829-835): model > 0.01, cf, df <= 1e5; NaN / inf from a
            // non-positive model are discarded by the select
            const bool okm = model > (T)10e-3;
            inv = trcp<T>(model);
            const T t = data * inv;
            cf = t - (T)1;
            df = t * inv;
            cf = cf > (T)10e4 ? (T)10e4 : cf;
            df = df > (T)10e4 ? (T)10e4 : df;
            cf = okm ? cf : (T)0;
            df = okm ? df : (T)0;
        }
        c0 += cf;
        cpx = tfma<T>(cf, f[0], cpx);
        cc1 = tfma<T>(cf, f[1], cc1);
        cc2 = tfma<T>(cf, f[2], cc2);
        cg1 = tfma<T>(cf, f[3], cg1);
        cg2 = tfma<T>(cf, f[4], cg2);
        d0 += df;
        dpx2 = tfma<T>(df, f[0] * f[0], dpx2);
        dc1 = tfma<T>(df, f[1] * f[1], dc1);
        dg1 = tfma<T>(df, f[3] * f[3], dg1);
        if (METHOD == 0) dgp = tfma<T>(df, f[3] * f[0], dgp);
    }
    const A Ncy1 = N * (A)fyd[0], Ncy2 = N * (A)fyd[1];
    const A cpx_a = (A)cpx, dpx2_a = (A)dpx2;
    num[0] = tfma<A>(NPy, (A)cc1, num[0]);
    den[0] += NPy * (A)cc2 - NPy * NPy * (A)dc1;
    num[1] = tfma<A>(Ncy1, cpx_a, num[1]);
    den[1] += Ncy2 * cpx_a - Ncy1 * Ncy1 * dpx2_a;
    num[2] = tfma<A>(PSFy, cpx_a, num[2]);
    den[2] -= PSFy * PSFy * dpx2_a;
    num[3] += (A)c0;
    den[3] -= (A)d0;
    if (METHOD == 1) {
        const A Ngy1 = N * (A)fyd[2], Ngy2 = N * (A)fyd[3];
        num[4] = tfma<A>(NPy, (A)cg1, num[4]);
        den[4] += NPy * (A)cg2 - NPy * NPy * (A)dg1;
        num[5] = tfma<A>(Ngy1, cpx_a, num[5]);
        den[5] += Ngy2 * cpx_a - Ngy1 * Ngy1 * dpx2_a;
    } else {
        // dudt = N (PSFy dPx + PSFx dPy); the reference's d2udt2 has photons on the
        // first term only (gaussmle.py:380-382)
        const A gy1 = (A)fyd[2], gy2 = (A)fyd[3];
        num[4] += N * (PSFy * (A)cg1 + gy1 * cpx_a);
        den[4] += (NPy * (A)cg2 + (A)2 * gy1 * (A)cg1 + gy2 * cpx_a) -
                  N * N * (PSFy * PSFy * (A)dg1 + (A)2 * PSFy * gy1 * (A)dgp +
                           gy1 * gy1 * dpx2_a);
    }
}

template <int BOX, int METHOD, typename T, typename A, class Roi, class Xf, class Tab>
PB_HD void newton_sums(const Roi& roi, const float th[6], const Xf& xf, A num[6], A den[6], const Tab& tab) {
#pragma unroll
    for (int l = 0; l < 6; l++) { num[l] = 0; den[l] = 0; }
    if constexpr (sizeof(T) == 4) {
        const AxisF ay = make_axis_f<METHOD>(METHOD == 1 ? th[5] : th[4]);
        EdgeF<METHOD> EA, EB = {};
        double e = -(double)th[1] - 0.5;
        float gf = 0.0f;
#pragma unroll 1
        for (int k = 0; k < (BOX + 1) / 2; k++) {
            double psf;
            float d[4];
            EA = eval_edge_f<METHOD>(e, (gf - th[1]) - 0.5f, ay, tab);
            if (k > 0) {
                pixel_factors_f<METHOD>(EB, EA, ay, psf, d);
                accumulate_row<BOX, METHOD, T, A>(2 * k - 1, psf, d, roi, xf, th[2], th[3], num, den);
            }
            EB = eval_edge_f<METHOD>(e + 1.0, (gf - th[1]) + 0.5f, ay, tab);
            pixel_factors_f<METHOD>(EA, EB, ay, psf, d);
            accumulate_row<BOX, METHOD, T, A>(2 * k, psf, d, roi, xf, th[2], th[3], num, den);
            e += 2.0;
            gf += 2.0f;
        }
        return;
    }
    const Axis ay = make_axis(METHOD == 1 ? th[5] : th[4]);
    Edge<METHOD> EA, EB = {};
#pragma unroll 1
    for (int k = 0; k < (BOX + 1) / 2; k++) {
        double fy[5];
        EA = eval_edge<METHOD>(2 * k, th[1], ay);
        if (k > 0) {
            pixel_factors<METHOD>(EB, EA, ay, fy);
            accumulate_row<BOX, METHOD, T, A>(2 * k - 1, fy[0], fy + 1, roi, xf, th[2], th[3], num, den);
        }
        EB = eval_edge<METHOD>(2 * k + 1, th[1], ay);
        pixel_factors<METHOD>(EA, EB, ay, fy);
        accumulate_row<BOX, METHOD, T, A>(2 * k, fy[0], fy + 1, roi, xf, th[2], th[3], num, den);
    }
}

// Clamped per-parameter Newton step in float32 (gaussmle.py:647-670, 860-884).
// Returns true when the stopping rule is met (:632-638, :844-852).
template <int BOX, int METHOD, typename A>
PB_HD bool update_theta(float th[6], const float ms[6], const A num[6], const A den[6],
                        double eps) {
    constexpr int NP = METHOD == 1 ? 6 : 5;
    float tn[6];
#pragma unroll
    for (int l = 0; l < NP; l++) {
        const float nu = (float)num[l], de = (float)den[l];
        float upd;
        if (de == 0.0f) {
            if (METHOD == 1) {
                const float sg = nu > 0.f ? 1.f : (nu < 0.f ? -1.f : nu);
                upd = sg * ms[l];
            } else {
                const float pr = nu * ms[l];
                upd = pr > 0.f ? 1.f : (pr < 0.f ? -1.f : pr);
            }
        } else {
            upd = fminf(fmaxf(nu / de, -ms[l]), ms[l]);
        }
        float v = th[l] - upd;
        if (l == 2) v = fmaxf(v, 1.0f);
        if (l >= 3) v = fmaxf(v, 0.01f);
        if (METHOD == 0 && l == 4) v = fminf(v, (float)BOX);
        tn[l] = v;
    }
    if (METHOD == 0) tn[5] = tn[4];
    bool conv = ((double)fabsf(th[0] - tn[0]) < eps) && ((double)fabsf(th[1] - tn[1]) < eps);
    if (METHOD == 1)
        conv = conv && ((double)fabsf(th[4] - tn[4]) < eps) && ((double)fabsf(th[5] - tn[5]) < eps);
#pragma unroll
    for (int l = 0; l < 6; l++) th[l] = tn[l];
    return conv;
}

// ---- CRLB: diagonal of the (pseudo-)inverse Fisher matrix --------------------------
// Jacobi eigen pseudo-inverse diagonal (rare fallback; np.linalg.pinv, rcond 1e-15)
template <int NP>
PB_HD_NOINLINE void pinv_diag_jacobi(const double* Msym /*NP*NP*/, double* diag) {
    double A[NP * NP], V[NP * NP];
    bool finite = true;
    for (int i = 0; i < NP * NP; i++) {
        A[i] = Msym[i];
        finite = finite && isfinite(A[i]);
    }
    if (!finite) {
        for (int i = 0; i < NP; i++) diag[i] = NAN;
        return;
    }
    for (int i = 0; i < NP; i++)
        for (int j = 0; j < NP; j++) V[i * NP + j] = (i == j) ? 1.0 : 0.0;
    // Cyclic Jacobi converges quadratically (off-diagonal weight 1e-4 -> 1e-9 -> 1e-20 -> 1e-36 of the diagonal's
    // in four sweeps); a matrix whose off-diagonal weight settles just above the 1e-34 target (rounding floor)
    // used to run all 60 sweeps -- 2 spots in 400 000, each ~0.7 ms of one lane, which was the tail of every
    // CRLB launch (SMs busy 43-55 % of the kernel's duration).  Stop as soon as a sweep no longer reduces it.
    double prev_off = INFINITY;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, dsum = 0.0;
        for (int i = 0; i < NP; i++)
            for (int j = 0; j < NP; j++) {
                double v = A[i * NP + j] * A[i * NP + j];
                if (i != j) off += v; else dsum += v;
            }
        if (off <= 1e-60 || off <= 1e-34 * dsum || off >= 0.25 * prev_off) break;
        prev_off = off;
        for (int p = 0; p < NP - 1; p++)
            for (int q = p + 1; q < NP; q++) {
                double apq = A[p * NP + q];
                if (apq == 0.0) continue;
                double th = (A[q * NP + q] - A[p * NP + p]) / (2.0 * apq);
                double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                if (!isfinite(th)) t = 0.0;
                double c = rsqrt64(t * t + 1.0), s = t * c;
                for (int k = 0; k < NP; k++) {
                    double akp = A[k * NP + p], akq = A[k * NP + q];
                    A[k * NP + p] = c * akp - s * akq;
                    A[k * NP + q] = s * akp + c * akq;
                }
                for (int k = 0; k < NP; k++) {
                    double apk = A[p * NP + k], aqk = A[q * NP + k];
                    A[p * NP + k] = c * apk - s * aqk;
                    A[q * NP + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < NP; k++) {
                    double vkp = V[k * NP + p], vkq = V[k * NP + q];
                    V[k * NP + p] = c * vkp - s * vkq;
                    V[k * NP + q] = s * vkp + c * vkq;
                }
            }
    }
    double smax = 0.0;
    for (int i = 0; i < NP; i++) smax = fmax(smax, fabs(A[i * NP + i]));
    double cutoff = 1e-15 * smax;
    for (int i = 0; i < NP; i++) {
        double acc = 0.0;
        for (int k = 0; k < NP; k++) {
            double lam = A[k * NP + k];
            if (fabs(lam) > cutoff) acc += V[i * NP + k] * V[i * NP + k] / lam;
        }
        diag[i] = acc;
    }
}

// Diagonal of inv(M), M symmetric positive definite, packed upper triangle m[idx(k,l)], k <= l.
// false when the diagonally scaled Cholesky meets a non-positive / tiny pivot.
template <int NP>
PB_HD bool inv_diag_cholesky(const double* m, double* diag, double* min_pivot = nullptr) {
#define PB_TRI(k, l) ((k) * NP - ((k) * ((k) - 1)) / 2 + ((l) - (k)))
    double d[NP];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        double mii = m[PB_TRI(i, i)];
        ok = ok && (mii > 0.0) && isfinite(mii);
        d[i] = rsqrt64(mii);
    }
    if (!ok) return false;
    double L[NP][NP];
    double rl[NP];
#pragma unroll
    for (int j = 0; j < NP; j++) {
        double s = 1.0;
#pragma unroll
        for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
        ok = ok && (s > 1e-12);
        if (min_pivot) *min_pivot = s < *min_pivot ? s : *min_pivot;
        rl[j] = rsqrt64(s);
        L[j][j] = s * rl[j];
#pragma unroll
        for (int i = j + 1; i < NP; i++) {
            double c = m[PB_TRI(j, i)] * d[i] * d[j];
#pragma unroll
            for (int k = 0; k < j; k++) c -= L[i][k] * L[j][k];
            L[i][j] = c * rl[j];
        }
    }
    if (!ok) return false;
#pragma unroll
    for (int j = 0; j < NP; j++) {
        double X[NP];
        double acc = 0.0;
#pragma unroll
        for (int i = j; i < NP; i++) {
            double s = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int k = j; k < i; k++) s -= L[i][k] * X[k];
            X[i] = s * rl[i];
            acc += X[i] * X[i];
        }
        diag[j] = acc * d[j] * d[j];
    }
#undef PB_TRI
    return true;
}

// One pixel row of the Fisher matrix and the log-likelihood.  dudt_k = (row factor) x (column
// factor of kind c1 / px / 1 / g1), so the row accumulates the ten 1/model-weighted products of
// column-factor pairs and the row factors are applied once per row.
template <int BOX, int METHOD, int NF, class Roi, class Xf3>
PB_HD void crlb_row(int j, const double fy[5], const Roi& roi, const Xf3& xf, double N, double bg,
                    double M[NF], double& ll) {
    constexpr int NP = METHOD == 1 ? 6 : 5;
    const double PSFy = fy[0], NPy = N * fy[0];
    double ac[10];
#pragma unroll
    for (int q = 0; q < 10; q++) ac[q] = 0.0;
#pragma unroll
    for (int i = 0; i < BOX; i++) {
        double f[3];   // px, c1, g1
        xf.get(i, f);
        const double model = fma(f[0], NPy, bg);
        const double w = rcp64(model);
        const double c1w = f[1] * w, pxw = f[0] * w, g1w = f[2] * w;
        ac[0] = fma(c1w, f[1], ac[0]);   // c1 c1
        ac[1] = fma(c1w, f[0], ac[1]);   // c1 px
        ac[2] += c1w;                    // c1 1
        ac[3] = fma(c1w, f[2], ac[3]);   // c1 g1
        ac[4] = fma(pxw, f[0], ac[4]);   // px px
        ac[5] += pxw;                    // px 1
        ac[6] = fma(pxw, f[2], ac[6]);   // px g1
        ac[7] += w;                      // 1 1
        ac[8] += g1w;                    // 1 g1
        ac[9] = fma(g1w, f[2], ac[9]);   // g1 g1
        const float dataf = roi(j * BOX + i);
        if (model > 0.0) {
            // d ln(model) - model - f32(d logf(d)) + d  (gaussmle.py:935-945)
            if (dataf > 0.0f)
                ll += (double)dataf * log_pos(model) - model - (double)(dataf * logf(dataf)) +
                      (double)dataf;
            else
                ll -= model;
        }
    }
    // pair index of column-factor kinds: 0 = c1, 1 = px, 2 = one, 3 = g1
#define PB_PR(p, q) ac[((p) < (q) ? (p) : (q)) * 4 - (((p) < (q) ? (p) : (q)) * (((p) < (q) ? (p) : (q)) - 1)) / 2 + \
                       (((p) < (q) ? (q) : (p)) - ((p) < (q) ? (p) : (q)))]
    if (METHOD == 1) {
        const double arow[6] = {NPy, N * fy[1], PSFy, 1.0, NPy, N * fy[3]};
        constexpr int kind[6] = {0, 1, 1, 2, 3, 1};
        int q = 0;
#pragma unroll
        for (int k = 0; k < 6; k++)
#pragma unroll
            for (int l = k; l < 6; l++) {
                M[q] = fma(arow[k] * arow[l], PB_PR(kind[k], kind[l]), M[q]);
                q++;
            }
    } else {
        // dudt_4 = N PSFy g1(i) + N gy1 px(i)
        const double arow[4] = {NPy, N * fy[1], PSFy, 1.0};
        constexpr int kind[4] = {0, 1, 1, 2};
        const double u = NPy, v = N * fy[3];
        int q = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int l = k; l < 4; l++) {
                M[q] = fma(arow[k] * arow[l], PB_PR(kind[k], kind[l]), M[q]);
                q++;
            }
            M[q] = fma(arow[k], u * PB_PR(kind[k], 3) + v * PB_PR(kind[k], 1), M[q]);
            q++;
        }
        M[q] += u * u * PB_PR(3, 3) + 2.0 * u * v * PB_PR(3, 1) + v * v * PB_PR(1, 1);
    }
#undef PB_PR
    (void)NP;
}

// CRLB + log-likelihood at theta (gaussmle.py:673-742, 887-954).  Xf3 concept: put(col, f[5])
// keeps (PSF, d/dmu, d/dsigma) of the column; get(col, double f[3]).
// Returns status flags (bit 1: pseudo-inverse fallback, bit 2: non-finite CRLB).
template <int BOX, int METHOD, class Roi, class Xf3>
PB_HD int crlb_loglik(const Roi& roi, const float th[6], Xf3& xf, float crlb[6], float* loglik) {
    constexpr int NP = METHOD == 1 ? 6 : 5;
    constexpr int NF = NP * (NP + 1) / 2;
    int st = 0;
    const double N = (double)th[2], bg = (double)th[3];
    {
        const Axis ax = make_axis(th[4]);
        Edge<METHOD> EA, EB = {};
#pragma unroll 1
        for (int k = 0; k < (BOX + 1) / 2; k++) {
            double f[5];
            EA = eval_edge<METHOD>(2 * k, th[0], ax);
            if (k > 0) {
                pixel_factors<METHOD>(EB, EA, ax, f);
                xf.put(2 * k - 1, f);
            }
            EB = eval_edge<METHOD>(2 * k + 1, th[0], ax);
            pixel_factors<METHOD>(EA, EB, ax, f);
            xf.put(2 * k, f);
        }
    }
    double M[NF];
#pragma unroll
    for (int q = 0; q < NF; q++) M[q] = 0.0;
    double ll = 0.0;
    {
        const Axis ay = make_axis(METHOD == 1 ? th[5] : th[4]);
        Edge<METHOD> EA, EB = {};
#pragma unroll 1
        for (int k = 0; k < (BOX + 1) / 2; k++) {
            double fy[5];
            EA = eval_edge<METHOD>(2 * k, th[1], ay);
            if (k > 0) {
                pixel_factors<METHOD>(EB, EA, ay, fy);
                crlb_row<BOX, METHOD, NF>(2 * k - 1, fy, roi, xf, N, bg, M, ll);
            }
            EB = eval_edge<METHOD>(2 * k + 1, th[1], ay);
            pixel_factors<METHOD>(EA, EB, ay, fy);
            crlb_row<BOX, METHOD, NF>(2 * k, fy, roi, xf, N, bg, M, ll);
        }
    }
    *loglik = (float)ll;
    double dg[NP];
    // np.linalg.pinv drops singular values below 1e-15 * max: a tiny diagonal entry means a
    // (numerically) null direction -> pseudo-inverse
    double dmin = INFINITY, dmax = 0.0;
    {
        int q = 0;
#pragma unroll
        for (int k = 0; k < NP; k++) {
            dmin = fmin(dmin, M[q]); dmax = fmax(dmax, M[q]);
            q += NP - k;
        }
    }
    bool ok = (dmin > 1e-13 * dmax) && inv_diag_cholesky<NP>(M, dg);
    if (!ok) {
        st |= 2;
        double Mfull[NP * NP];
        int q = 0;
        for (int k = 0; k < NP; k++)
            for (int l = k; l < NP; l++) {
                Mfull[k * NP + l] = M[q];
                Mfull[l * NP + k] = M[q];
                q++;
            }
        pinv_diag_jacobi<NP>(Mfull, dg);
    }
    bool bad = false;
#pragma unroll
    for (int l = 0; l < NP; l++) {
        crlb[l] = (float)dg[l];
        bad = bad || !isfinite(crlb[l]);
    }
    if (METHOD == 0) crlb[5] = crlb[4];
    if (bad) st |= 4;
    return st;
}

// ---- fast CRLB + log-likelihood pass ------------------------------------------------------
// Same quantities as crlb_row / crlb_loglik with the per-pixel work moved off the FP64 pipe:
//   * the ten 1/model-weighted pair sums of a row are accumulated in FLOAT32 (the reference forms
//     dudt[k] * dudt[l] in float32 itself, gaussmle.py:929-932: its Fisher entries carry the same
//     6e-8 relative rounding), the row factors and the accumulation over rows stay float64;
//   * ln(model) comes from the 128-entry table (log_tab).
// A (near-)singular Fisher matrix -- a sigma collapsed onto its floor, where np.linalg.pinv returns
// exact zeros -- amplifies float32 noise, so the result is only accepted when every pivot of the
// diagonally scaled Cholesky factorisation is above 1e-2 (error bound ~ 2e-7 / 1e-2); otherwise
// the caller repeats the pass with crlb_loglik (all float64).  Returns -1 in that case.
// XfF concept: put_f(col, psf, d[4]) keeps PSF as double and (d/dmu, d/dsigma) as floats;
// get(col, double& px, float& c1, float& g1).
template <int BOX, int METHOD, int NF, class Roi, class XfF, class Tab>
PB_HD void crlb_row_fast(int j, double PSFy, const float fyd[4], const Roi& roi, const XfF& xf, const Tab& tab,
                         double N, double bg, float M[NF], double& ll) {
    const double NPy = N * PSFy;
    const float Nf = (float)N;
    float ac[10];
#pragma unroll
    for (int q = 0; q < 10; q++) ac[q] = 0.0f;
#pragma unroll
    for (int i = 0; i < BOX; i++) {
        double px;
        float c1, g1;
        xf.get(i, px, c1, g1);
        const double model = fma(px, NPy, bg);
        const float pxf = (float)px;
        const float w = rcp32((float)model);
        const float c1w = c1 * w, pxw = pxf * w, g1w = g1 * w;
        ac[0] = fmaf(c1w, c1, ac[0]);    // c1 c1
        ac[1] = fmaf(c1w, pxf, ac[1]);   // c1 px
        ac[2] += c1w;                    // c1 1
        ac[3] = fmaf(c1w, g1, ac[3]);    // c1 g1
        ac[4] = fmaf(pxw, pxf, ac[4]);   // px px
        ac[5] += pxw;                    // px 1
        ac[6] = fmaf(pxw, g1, ac[6]);    // px g1
        ac[7] += w;                      // 1 1
        ac[8] += g1w;                    // 1 g1
        ac[9] = fmaf(g1w, g1, ac[9]);    // g1 g1
        const float dataf = roi(j * BOX + i);
        // d ln(model) - model - f32(d logf(d)) + d  (gaussmle.py:935-945); -model for d <= 0, nothing
        // for model <= 0.  Branch-free: the logs of non-positive arguments are discarded by the selects.
        const double d = (double)dataf;
        const double tpos = fma(d, log_tab(model, tab), d - model) - (double)(dataf * logf(dataf));
        const double t = dataf > 0.0f ? tpos : -model;
        ll += model > 0.0 ? t : 0.0;
    }
#define PB_PR(p, q) ac[((p) < (q) ? (p) : (q)) * 4 - (((p) < (q) ? (p) : (q)) * (((p) < (q) ? (p) : (q)) - 1)) / 2 + \
                               (((p) < (q) ? (q) : (p)) - ((p) < (q) ? (p) : (q)))]
    if (METHOD == 1) {
        // row factors and their products in float32 as well: the pair sums carry float32 rounding anyway
        const float arow[6] = {(float)NPy, Nf * fyd[0], (float)PSFy, 1.0f, (float)NPy, Nf * fyd[2]};
        constexpr int kind[6] = {0, 1, 1, 2, 3, 1};
        int q = 0;
#pragma unroll
        for (int k = 0; k < 6; k++)
#pragma unroll
            for (int l = k; l < 6; l++) {
                M[q] = fmaf(arow[k] * arow[l], PB_PR(kind[k], kind[l]), M[q]);
                q++;
            }
    } else {
        const float arow[4] = {(float)NPy, Nf * fyd[0], (float)PSFy, 1.0f};
        constexpr int kind[4] = {0, 1, 1, 2};
        const float u = (float)NPy, v = Nf * fyd[2];
        int q = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int l = k; l < 4; l++) {
                M[q] = fmaf(arow[k] * arow[l], PB_PR(kind[k], kind[l]), M[q]);
                q++;
            }
            M[q] = fmaf(arow[k], u * PB_PR(kind[k], 3) + v * PB_PR(kind[k], 1), M[q]);
            q++;
        }
        M[q] += u * u * PB_PR(3, 3) + 2.0f * u * v * PB_PR(3, 1) + v * v * PB_PR(1, 1);
    }
#undef PB_PR
}

template <int BOX, int METHOD, class Roi, class XfF, class Tab, class ETab>
PB_HD int crlb_loglik_fast(const Roi& roi, const float th[6], XfF& xf, const Tab& tab, const ETab& etab,
                           float crlb[6], float* loglik) {
    constexpr int NP = METHOD == 1 ? 6 : 5;
    constexpr int NF = NP * (NP + 1) / 2;
    const double N = (double)th[2], bg = (double)th[3];
    {
        const AxisF ax = make_axis_f<METHOD>(th[4]);
        EdgeF<METHOD> EA, EB = {};
        double e = -(double)th[0] - 0.5;
        float gf = 0.0f;                  // edge index as float32 (exact)
#pragma unroll 1
        for (int k = 0; k < (BOX + 1) / 2; k++) {
            double psf;
            float d[4];
            EA = eval_edge_f<METHOD>(e, (gf - th[0]) - 0.5f, ax, etab);
            if (k > 0) {
                pixel_factors_f<METHOD>(EB, EA, ax, psf, d);
                xf.put_f(2 * k - 1, psf, d);
            }
            EB = eval_edge_f<METHOD>(e + 1.0, (gf - th[0]) + 0.5f, ax, etab);
            pixel_factors_f<METHOD>(EA, EB, ax, psf, d);
            xf.put_f(2 * k, psf, d);
            e += 2.0;
            gf += 2.0f;
        }
    }
    float Mf[NF];
#pragma unroll
    for (int q = 0; q < NF; q++) Mf[q] = 0.0f;
    double ll = 0.0;
    {
        const AxisF ay = make_axis_f<METHOD>(METHOD == 1 ? th[5] : th[4]);
        EdgeF<METHOD> EA, EB = {};
        double e = -(double)th[1] - 0.5;
        float gf = 0.0f;
#pragma unroll 1
        for (int k = 0; k < (BOX + 1) / 2; k++) {
            double psf;
            float d[4];
            EA = eval_edge_f<METHOD>(e, (gf - th[1]) - 0.5f, ay, etab);
            if (k > 0) {
                pixel_factors_f<METHOD>(EB, EA, ay, psf, d);
                crlb_row_fast<BOX, METHOD, NF>(2 * k - 1, psf, d, roi, xf, tab, N, bg, Mf, ll);
            }
            EB = eval_edge_f<METHOD>(e + 1.0, (gf - th[1]) + 0.5f, ay, etab);
            pixel_factors_f<METHOD>(EA, EB, ay, psf, d);
            crlb_row_fast<BOX, METHOD, NF>(2 * k, psf, d, roi, xf, tab, N, bg, Mf, ll);
            e += 2.0;
            gf += 2.0f;
        }
    }
    double M[NF];
    bool finite = true;
#pragma unroll
    for (int q = 0; q < NF; q++) { M[q] = (double)Mf[q]; finite = finite && isfinite(Mf[q]); }
    if (!finite) return -1;           // float32 overflow of a row-factor product: repeat in float64
    double dg[NP];
    double dmin = INFINITY, dmax = 0.0;
    {
        int q = 0;
#pragma unroll
        for (int k = 0; k < NP; k++) {
            dmin = fmin(dmin, M[q]); dmax = fmax(dmax, M[q]);
            q += NP - k;
        }
    }
    double smin = 1.0;
    const bool ok = (dmin > 1e-13 * dmax) && inv_diag_cholesky<NP>(M, dg, &smin) && smin > 1e-2 &&
                    isfinite(ll);
    if (!ok) return -1;
    bool bad = false;
#pragma unroll
    for (int l = 0; l < NP; l++) {
        crlb[l] = (float)dg[l];
        bad = bad || !isfinite(crlb[l]);
    }
    if (bad) return -1;
    if (METHOD == 0) crlb[5] = crlb[4];
    *loglik = (float)ll;
    return 0;
}

}  // namespace tps
