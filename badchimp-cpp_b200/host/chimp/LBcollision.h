// Per-node helper functions of the reference's API for HOST code, under the reference's names and argument lists:
// calcRho, calcVel (LBmacroscopic.h:11-45), calcOmegaBGK, calcOmegaBGKTRT, calcDeltaOmegaQ, calcDeltaOmegaQTRT,
// calcDeltaOmegaF, calcDeltaOmegaFTRT (LBcollision.h:27-121,194-236), calcDeltaOmegaST, calcDeltaOmegaRC
// (LBcollision2phase.h:7-20,70-86), calcfeq, grad, vecNorm (LButilities.h:12-22,62-91), initiateLbField
// (LBinitiatefield.h:33-57).  On the GPU these live inside the step kernels (csrc/collide.cuh, kernels.cuh); a main
// that switches its node loop to the engine keeps these for set-up, diagnostics and checks on the host.  Every
// expression keeps the reference's order of operations (no FMA contraction: build with -ffp-contract=off), so the
// host loop of host/apps/cpu_loop.cpp reproduces the reference's dumps bit for bit (tests/test_host_cpp.py).
#ifndef CHIMP_LBCOLLISION_H
#define CHIMP_LBCOLLISION_H

#include <cmath>

#include "LBfield.h"
#include "LBgrid.h"
#include "LBlattices.h"

// rho = sum_q f_q, ascending q from 0.0
template <typename DXQY, typename T>
inline lbBase_t calcRho(const T &f)
{
    return DXQY::qSum(f);
}

// u = sum_q c_q f_q / rho
template <typename DXQY, typename T1>
inline std::valarray<lbBase_t> calcVel(const T1 &f, const lbBase_t &rho)
{
    return DXQY::qSumC(f) / rho;
}

// Guo: u = (sum_q c_q f_q + F / 2) / rho
template <typename DXQY, typename T1, typename T2>
inline std::valarray<lbBase_t> calcVel(const T1 &f, const lbBase_t &rho, const T2 &force)
{
    std::valarray<lbBase_t> ret = DXQY::qSumC(f);
    for (int d = 0; d < DXQY::nD; ++d) ret[d] = (ret[d] + 0.5 * force[d]) / rho;
    return ret;
}

template <typename DXQY, typename T>
inline std::valarray<lbBase_t> calcfeq(const lbBase_t &rho, const lbBase_t &u_sq, const T &cu)
{
    std::valarray<lbBase_t> ret(DXQY::nQ);
    for (int q = 0; q < DXQY::nQ; ++q)
        ret[q] = rho * DXQY::w[q] * (1.0 + DXQY::c2Inv * cu[q] + DXQY::c4Inv0_5 * (cu[q] * cu[q] - DXQY::c2 * u_sq));
    return ret;
}

// Omega_q = -(1/tau) (f_q - f_q^eq)
template <typename DXQY, typename T>
inline std::valarray<lbBase_t> calcOmegaBGK(const T &f, const lbBase_t &tau, const lbBase_t &rho, const lbBase_t &u_sq,
                                            const std::valarray<lbBase_t> &cu)
{
    std::valarray<lbBase_t> ret(DXQY::nQ);
    const lbBase_t tau_inv = 1.0 / tau;
    for (int q = 0; q < DXQY::nQ; ++q)
        ret[q] = -tau_inv * (f[q] - rho * DXQY::w[q] * (1.0 + DXQY::c2Inv * cu[q] + DXQY::c4Inv0_5 * (cu[q] * cu[q] - DXQY::c2 * u_sq)));
    return ret;
}

// two relaxation times: even part with tauSym, odd part with tauAnti
template <typename DXQY, typename T>
inline std::valarray<lbBase_t> calcOmegaBGKTRT(const T &f, const lbBase_t &tauSym, const lbBase_t &tauAnti, const lbBase_t &rho,
                                               const lbBase_t &u_sq, const std::valarray<lbBase_t> &cu)
{
    std::valarray<lbBase_t> ret(DXQY::nQ);
    const lbBase_t tauSym_inv = 1.0 / tauSym, tauAnti_inv = 1.0 / tauAnti;
    for (int q = 0; q < DXQY::nQ; ++q) {
        const int r = DXQY::reverseDirection(q);
        const lbBase_t fSym = 0.5 * (f[q] + f[r]);
        const lbBase_t fAnti = 0.5 * (f[q] - f[r]);
        ret[q] = -tauSym_inv * (fSym - rho * DXQY::w[q] * (1.0 + DXQY::c4Inv0_5 * (cu[q] * cu[q] - DXQY::c2 * u_sq))) -
                 tauAnti_inv * (fAnti - rho * DXQY::w[q] * DXQY::c2Inv * cu[q]);
    }
    return ret;
}

// mass source correction
template <typename DXQY>
inline std::valarray<lbBase_t> calcDeltaOmegaQ(const lbBase_t &tau, const std::valarray<lbBase_t> &cu, const lbBase_t &u_sq,
                                               const lbBase_t &source)
{
    std::valarray<lbBase_t> ret(DXQY::nQ);
    const lbBase_t tau_factor = (1 - 0.5 / tau);
    for (int q = 0; q < DXQY::nQ; ++q)
        ret[q] = tau_factor * source * DXQY::w[q] * (1.0 + DXQY::c2Inv * cu[q] + DXQY::c4Inv0_5 * (cu[q] * cu[q] - DXQY::c2 * u_sq));
    return ret;
}

template <typename DXQY>
inline std::valarray<lbBase_t> calcDeltaOmegaQTRT(const lbBase_t &tauSym, const lbBase_t &tauAnti, const std::valarray<lbBase_t> &cu,
                                                  const lbBase_t &u_sq, const lbBase_t &source)
{
    std::valarray<lbBase_t> ret(DXQY::nQ);
    const lbBase_t tauSym_factor = (1 - 0.5 / tauSym), tauAnti_factor = (1 - 0.5 / tauAnti);
    for (int q = 0; q < DXQY::nQ; ++q)
        ret[q] = source * DXQY::w[q] *
                 (tauAnti_factor * DXQY::c2Inv * cu[q] + tauSym_factor * (1.0 + DXQY::c4Inv0_5 * (cu[q] * cu[q] - DXQY::c2 * u_sq)));
    return ret;
}

// Guo forcing term
template <typename DXQY>
inline std::valarray<lbBase_t> calcDeltaOmegaF(const lbBase_t &tau, const std::valarray<lbBase_t> &cu, const lbBase_t &uF,
                                               const std::valarray<lbBase_t> &cF)
{
    std::valarray<lbBase_t> ret(DXQY::nQ);
    const lbBase_t tau_factor = (1 - 0.5 / tau);
    for (int q = 0; q < DXQY::nQ; ++q)
        ret[q] = DXQY::w[q] * tau_factor * (DXQY::c2Inv * cF[q] + DXQY::c4Inv * (cF[q] * cu[q] - DXQY::c2 * uF));
    return ret;
}

template <typename DXQY>
inline std::valarray<lbBase_t> calcDeltaOmegaFTRT(const lbBase_t &tauSym, const lbBase_t &tauAnti, const lbBase_t &phi,
                                                  const std::valarray<lbBase_t> &cu, const lbBase_t &uF, const std::valarray<lbBase_t> &cF)
{
    std::valarray<lbBase_t> ret(DXQY::nQ);
    const lbBase_t tauSym_factor = (1 - 0.5 / tauSym), tauAnti_factor = (1 - 0.5 / tauAnti);
    for (int q = 0; q < DXQY::nQ; ++q)
        ret[q] = DXQY::w[q] * phi * (tauAnti_factor * DXQY::c2Inv * cF[q] + tauSym_factor * DXQY::c4Inv * (cF[q] * cu[q] - DXQY::c2 * uF));
    return ret;
}

// colour-gradient surface tension perturbation; cCGNorm = c_q . n with n the unit colour gradient
template <typename DXQY>
inline std::valarray<lbBase_t> calcDeltaOmegaST(const lbBase_t &tau, const lbBase_t &sigma, const lbBase_t &CGNorm,
                                                const std::valarray<lbBase_t> &cCGNorm)
{
    std::valarray<lbBase_t> ret(DXQY::nQ);
    const lbBase_t AF0_5 = 1.125 * CGNorm * sigma / tau;
    for (int q = 0; q < DXQY::nQNonZero_; ++q) ret[q] = AF0_5 * (DXQY::w[q] * cCGNorm[q] * cCGNorm[q] - DXQY::B[q]);
    ret[DXQY::nQNonZero_] = -AF0_5 * DXQY::B[DXQY::nQNonZero_];
    return ret;
}

// recolouring term; |c_q| is 1 on the axes and sqrt(2) on the diagonals (division, as in the reference)
template <typename DXQY>
inline std::valarray<lbBase_t> calcDeltaOmegaRC(const lbBase_t &beta, const lbBase_t &rho0, const lbBase_t &rho1, const lbBase_t &rho,
                                                const std::valarray<lbBase_t> &cCGNorm)
{
    std::valarray<lbBase_t> ret(DXQY::nQ);
    const lbBase_t rhoFacBeta = beta * rho0 * rho1 / rho;
    for (int q = 0; q < DXQY::nQNonZero_; ++q) {
        int c2 = 0;
        for (int d = 0; d < DXQY::nD; ++d) c2 += DXQY::c(q, d) * DXQY::c(q, d);
        const lbBase_t cNorm = c2 == 1 ? 1.0 : std::sqrt(lbBase_t(c2));
        ret[q] = rhoFacBeta * DXQY::w[q] * cCGNorm[q] / cNorm;
    }
    ret[DXQY::nQNonZero_] = 0.0;
    return ret;
}

// lattice gradient of a scalar field at a node: values of the Q neighbours through the neighbour table
template <typename DXQY>
inline std::valarray<lbBase_t> grad(const ScalarField &sField, const int fieldNum, const int nodeNo, const Grid<DXQY> &grid)
{
    std::valarray<lbBase_t> scalarTmp(DXQY::nQ);
    for (int q = 0; q < DXQY::nQ; ++q) scalarTmp[q] = sField(fieldNum, grid.neighbor(q, nodeNo));
    return DXQY::grad(scalarTmp);
}

template <typename DXQY, typename T>
inline lbBase_t vecNorm(const T &vec)
{
    return std::sqrt(DXQY::dot(vec, vec));
}

// f(lbFieldNo, q, n) = equilibrium of rho(rhoFieldNo, n), vel(velFieldNo, n) on the bulk nodes
template <typename DXQY>
void initiateLbField(const int lbFieldNo, const int rhoFieldNo, const int velFieldNo, const std::vector<int> &bulk, const ScalarField &rho,
                     const VectorField<DXQY> &vel, LbField<DXQY> &f)
{
    for (auto nodeNo : bulk) {
        const std::valarray<lbBase_t> u = vel(velFieldNo, nodeNo);
        const std::valarray<lbBase_t> cu = DXQY::cDotAll(u);
        const lbBase_t uu = DXQY::dot(u, u);
        for (int q = 0; q < DXQY::nQ; ++q)
            f(lbFieldNo, q, nodeNo) = DXQY::w[q] * rho(rhoFieldNo, nodeNo) * (1.0 + DXQY::c2Inv * cu[q] + DXQY::c4Inv0_5 * (cu[q] * cu[q] - DXQY::c2 * uu));
    }
}

#endif
