// Per-node physics of the hot path, as register-level device functions.
//
// Each function restates one reference helper with the reference's exact operation order
// so that, compiled with -fmad=false, double results are bit-identical to the CPU build
// (which contains no FMA: CMakeLists.txt:71 has no -march).  Citations:
//   nodeRho        calcRho            LBmacroscopic.h:11-20  -> qSum   LBd3q19.h:173-180
//   firstMoment    qSumC              LBd3q19.h:182-190 (ascending q, zero terms skipped)
//   cDot           cDotAll            LBd3q19.h:129-153 (ascending dimension)
//   omegaBGK       calcOmegaBGK       LBcollision.h:27-46
//   omegaTRT       calcOmegaBGKTRT    LBcollision.h:50-75
//   deltaOmegaF    calcDeltaOmegaF    LBcollision.h:194-213
//   deltaOmegaFTRT calcDeltaOmegaFTRT LBcollision.h:215-236
//   deltaOmegaQ    calcDeltaOmegaQ    LBcollision.h:79-98
//   deltaOmegaQTRT calcDeltaOmegaQTRT LBcollision.h:100-121
//   deltaOmegaST   calcDeltaOmegaST   LBcollision2phase.h:7-20
//   deltaOmegaRC   calcDeltaOmegaRC   LBcollision2phase.h:70-86
//   latticeGrad    D3Q19::grad        LBd3q19.h:155-163 (grouped by weight class)
#pragma once
#include "lattices.cuh"

namespace chimp {

template <class L>
__device__ __forceinline__ double nodeRho(const double (&f)[L::nQ])
{
    double r = 0.0;
#pragma unroll
    for (int q = 0; q < L::nQ; ++q) r += f[q];
    return r;
}

template <class L, int d>
__device__ __forceinline__ double firstMoment(const double (&f)[L::nQ])
{
    double s = 0.0;
    bool first = true;
#pragma unroll
    for (int q = 0; q < L::nQ; ++q) {
        if (L::c(q, d) == 0) continue;
        if (first) { s = (L::c(q, d) > 0) ? f[q] : -f[q]; first = false; }
        else s = (L::c(q, d) > 0) ? s + f[q] : s - f[q];
    }
    return s;
}

// c_q . v with the reference's term order; 0.0 for the rest direction
template <class L, int q>
__device__ __forceinline__ double cDot(const double (&v)[3])
{
    double s = 0.0;
    bool first = true;
#pragma unroll
    for (int d = 0; d < L::nD; ++d) {
        if (L::c(q, d) == 0) continue;
        if (first) { s = (L::c(q, d) > 0) ? v[d] : -v[d]; first = false; }
        else s = (L::c(q, d) > 0) ? s + v[d] : s - v[d];
    }
    return s;
}

template <class L>
__device__ __forceinline__ double dotD(const double (&a)[3], const double (&b)[3])
{
    double s = a[0] * b[0] + a[1] * b[1];
    if (L::nD == 3) s = s + a[2] * b[2];
    return s;
}

// feq term  rho*w_q*(1 + 3 cu + 4.5 (cu^2 - u^2/3))   (LBcollision.h:43, LButilities.h:87)
template <class L, int q>
__device__ __forceinline__ double feq(double rho, double cu, double u2)
{
    return rho * L::w(q) * (1.0 + kC2Inv * cu + kC4Inv0_5 * (cu * cu - kC2 * u2));
}

template <class L, int q>
__device__ __forceinline__ double omegaBGK(double fq, double tauInv, double rho, double cu, double u2)
{
    return -tauInv * (fq - feq<L, q>(rho, cu, u2));
}

template <class L, int q>
__device__ __forceinline__ double omegaTRT(double fq, double fqRev, double tauSymInv, double tauAntiInv, double rho,
                                           double cu, double u2)
{
    const double fSym = 0.5 * (fq + fqRev);
    const double fAnti = 0.5 * (fq - fqRev);
    return -tauSymInv * (fSym - rho * L::w(q) * (1.0 + kC4Inv0_5 * (cu * cu - kC2 * u2))) -
           tauAntiInv * (fAnti - rho * L::w(q) * kC2Inv * cu);
}

template <class L, int q>
__device__ __forceinline__ double deltaOmegaF(double tauFactor, double cu, double uF, double cF)
{
    return L::w(q) * tauFactor * (kC2Inv * cF + kC4Inv * (cF * cu - kC2 * uF));
}

template <class L, int q>
__device__ __forceinline__ double deltaOmegaFTRT(double symFactor, double antiFactor, double phi, double cu, double uF,
                                                 double cF)
{
    return L::w(q) * phi * (antiFactor * kC2Inv * cF + symFactor * kC4Inv * (cF * cu - kC2 * uF));
}

template <class L, int q>
__device__ __forceinline__ double deltaOmegaQ(double tauFactor, double cu, double u2, double source)
{
    return tauFactor * source * L::w(q) * (1.0 + kC2Inv * cu + kC4Inv0_5 * (cu * cu - kC2 * u2));
}

template <class L, int q>
__device__ __forceinline__ double deltaOmegaQTRT(double symFactor, double antiFactor, double cu, double u2, double source)
{
    return source * L::w(q) * (antiFactor * kC2Inv * cu + symFactor * (1.0 + kC4Inv0_5 * (cu * cu - kC2 * u2)));
}

// cNorm[q]: 1 or SQRT2 (LBd3q19.h:34 with the literal of LBglobal.h:11)
template <class L, int q>
__device__ __forceinline__ constexpr double cNorm()
{
    return L::wclass(q) == 0 ? 0.0 : (L::wclass(q) == 1 ? 1.0 : 1.4142135623730950488);
}

// L::grad over the gathered neighbour scalars, grouped by weight class exactly as the
// generated header writes it:  w1c2Inv*( ... ) + w2c2Inv*( ... )
template <class L, int d>
__device__ __forceinline__ double latticeGrad(const double (&s)[L::nQ])
{
    double g1 = 0.0, g2 = 0.0;
    bool first1 = true, first2 = true;
#pragma unroll
    for (int q = 0; q < L::nQ; ++q) {
        if (L::c(q, d) == 0) continue;
        if (L::wclass(q) == 1) {
            if (first1) { g1 = (L::c(q, d) > 0) ? s[q] : -s[q]; first1 = false; }
            else g1 = (L::c(q, d) > 0) ? g1 + s[q] : g1 - s[q];
        } else {
            if (first2) { g2 = (L::c(q, d) > 0) ? s[q] : -s[q]; first2 = false; }
            else g2 = (L::c(q, d) > 0) ? g2 + s[q] : g2 - s[q];
        }
    }
    // w1c2Inv = w1*c2Inv, w2c2Inv = w2*c2Inv as compile-time constants (LBd3q19.h:27-30)
    constexpr double w1c2Inv = L::w(0) * kC2Inv;
    constexpr double w2c2Inv = (L::id == 0 ? L::w(1) : L::w(3)) * kC2Inv;
    return w1c2Inv * g1 + w2c2Inv * g2;
}

} // namespace chimp
