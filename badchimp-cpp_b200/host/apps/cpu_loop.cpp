// The reference's node loops written with the HOST mirror of its API (host/chimp/LBcollision.h: calcRho, calcVel,
// calcOmegaBGK[TRT], calcDeltaOmegaF[TRT], calcDeltaOmegaST/RC, grad, vecNorm, initiateLbField; LbField::propagateTo /
// swapData; HalfWayBounceBack::apply) on the CPU -- no engine, single rank.  It exists to pin those mirror
// functions: its output must equal the reference's dumps bit for bit (tests/test_host_cpp.py), so a main that
// mixes host-side calls of them with the GPU engine computes what the reference computes.
//
//   cpu_loop <std|trt|twophase> <D2Q9|D3Q19> <vtklb prefix> <out.bin> <steps> <params...>
//     std:      tau Fx Fy Fz            (std_case/main.cpp:109-146)
//     trt:      tauSym tauAnti Fx Fy Fz (same loop with calcOmegaBGKTRT / calcDeltaOmegaFTRT(phi = 1))
//     twophase: tau0 tau1 sigma beta momx Fy Fz      (twophase/main_TWOPHASE.cpp:236-392, Q = 0)
// out.bin: int size, f[size*nFields*nQ], rho[size*nFields], vel[size*nD]
#include <cstdio>

#include "../chimp/LBglobal.h"
#include "../chimp/LBlattices.h"
#include "../chimp/LBfield.h"
#include "../chimp/LBvtk.h"
#include "../chimp/LBgrid.h"
#include "../chimp/LBhalfwaybb.h"
#include "../chimp/LBcollision.h"

template <typename LT>
void writeOut(const std::string &file, int sz, LbField<LT> &f, ScalarField &rho, VectorField<LT> &vel)
{
    FILE *fp = std::fopen(file.c_str(), "wb");
    if (!fp) chimp_host::die("cannot open " + file);
    std::fwrite(&sz, sizeof(int), 1, fp);
    std::fwrite(f.data(), sizeof(double), std::size_t(sz) * f.num_fields() * LT::nQ, fp);
    std::fwrite(rho.data(), sizeof(double), std::size_t(sz) * rho.num_fields(), fp);
    std::fwrite(vel.data(), sizeof(double), std::size_t(sz) * LT::nD, fp);
    std::fclose(fp);
}

template <typename LT>
int runSingle(bool trt, const std::string &prefix, const std::string &out, int steps, char **a)
{
    LBvtk<LT> vtklb(prefix + "0.vtklb");
    Grid<LT> grid(vtklb);
    Nodes<LT> nodes(vtklb, grid);
    const std::vector<int> bulkNodes = findBulkNodes(nodes);
    const lbBase_t tau = std::atof(a[0]), tauAnti = trt ? std::atof(a[1]) : 0.0;
    std::valarray<lbBase_t> bodyForce(LT::nD);
    for (int d = 0; d < LT::nD; ++d) bodyForce[d] = std::atof(a[(trt ? 2 : 1) + d]);
    ScalarField rho(1, grid.size());
    vtklb.toAttribute("init_rho");
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) rho(0, n) = vtklb.template getScalarAttribute<lbBase_t>();
    VectorField<LT> vel(1, grid.size());
    HalfWayBounceBack<LT> bounceBackBnd(findFluidBndNodes(nodes), nodes, grid);
    LbField<LT> f(1, grid.size()), fTmp(1, grid.size());
    for (auto nodeNo : bulkNodes)
        for (int q = 0; q < LT::nQ; ++q) f(0, q, nodeNo) = LT::w[q] * rho(0, nodeNo);
    for (int i = 0; i < steps; ++i) {
        for (auto nodeNo : bulkNodes) {
            const std::valarray<lbBase_t> fNode = f(0, nodeNo);
            const lbBase_t rhoNode = calcRho<LT>(fNode);
            const auto velNode = calcVel<LT>(fNode, rhoNode, bodyForce);
            rho(0, nodeNo) = rhoNode;
            vel.set(0, nodeNo) = velNode;
            const lbBase_t u2 = LT::dot(velNode, velNode);
            const std::valarray<lbBase_t> cu = LT::cDotAll(velNode);
            const lbBase_t uF = LT::dot(velNode, bodyForce);
            const std::valarray<lbBase_t> cF = LT::cDotAll(bodyForce);
            if (trt) {
                const std::valarray<lbBase_t> omega = calcOmegaBGKTRT<LT>(fNode, tau, tauAnti, rhoNode, u2, cu);
                const std::valarray<lbBase_t> deltaOmegaF = calcDeltaOmegaFTRT<LT>(tau, tauAnti, 1.0, cu, uF, cF);
                fTmp.propagateTo(0, nodeNo, fNode + omega + deltaOmegaF, grid);
            } else {
                const std::valarray<lbBase_t> omegaBGK = calcOmegaBGK<LT>(fNode, tau, rhoNode, u2, cu);
                const std::valarray<lbBase_t> deltaOmegaF = calcDeltaOmegaF<LT>(tau, cu, uF, cF);
                fTmp.propagateTo(0, nodeNo, fNode + omegaBGK + deltaOmegaF, grid);
            }
        }
        f.swapData(fTmp);
        bounceBackBnd.apply(f, grid);
    }
    writeOut(out, grid.size(), f, rho, vel);
    return 0;
}

template <typename LT>
int runTwoPhase(const std::string &prefix, const std::string &out, int steps, char **a)
{
    LBvtk<LT> vtklb(prefix + "0.vtklb");
    Grid<LT> grid(vtklb);
    Nodes<LT> nodes(vtklb, grid);
    const std::vector<int> bulkNodes = findBulkNodes(nodes);
    const std::vector<int> solidBnd = findSolidBndNodes(nodes);
    HalfWayBounceBack<LT> bounceBackBnd(bulkNodes, nodes, grid);
    const lbBase_t tau0 = std::atof(a[0]), tau1 = std::atof(a[1]), sigma = std::atof(a[2]), beta = std::atof(a[3]), momx = std::atof(a[4]);
    std::valarray<lbBase_t> bodyForce(LT::nD);
    bodyForce[0] = 0.0;
    for (int d = 1; d < LT::nD; ++d) bodyForce[d] = std::atof(a[4 + d]);
    const lbBase_t nu0Inv = 1.0 / (LT::c2 * (tau0 - 0.5)), nu1Inv = 1.0 / (LT::c2 * (tau1 - 0.5));
    const int sz = grid.size();
    LbField<LT> f(2, sz), fTmp(2, sz);
    ScalarField rho(2, sz), cgField(1, sz);
    VectorField<LT> vel(1, sz);
    vtklb.toAttribute("rho0");
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) rho(0, n) = vtklb.template getScalarAttribute<float>();
    vtklb.toAttribute("rho1");
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) rho(1, n) = vtklb.template getScalarAttribute<float>();
    vtklb.toAttribute("wettability");
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) {
        const float val = vtklb.template getScalarAttribute<float>();
        if (nodes.isSolidBoundary(n)) { rho(0, n) = val; rho(1, n) = 1 - val; }
    }
    const int numNodesGlobal = int(bulkNodes.size());
    initiateLbField(0, 0, 0, bulkNodes, rho, vel, f);
    initiateLbField(1, 1, 0, bulkNodes, rho, vel, f);
    const lbBase_t lbBaseEps = 2.220446049250313e-16;
    for (int i = 0; i < steps; ++i) {
        for (auto nodeNo : bulkNodes) {
            const lbBase_t rho0Node = rho(0, nodeNo) = calcRho<LT>(f(0, nodeNo));
            const lbBase_t rho1Node = rho(1, nodeNo) = calcRho<LT>(f(1, nodeNo));
            cgField(0, nodeNo) = (rho0Node - rho1Node) / (rho0Node + rho1Node);
        }
        for (auto nodeNo : solidBnd) cgField(0, nodeNo) = (rho(0, nodeNo) - rho(1, nodeNo)) / (rho(0, nodeNo) + rho(1, nodeNo));
        lbBase_t meanfcX = 0.0;
        for (auto nodeNo : bulkNodes) {
            const std::valarray<lbBase_t> fTot = f(0, nodeNo) + f(1, nodeNo);
            meanfcX += LT::qSumC(fTot)[0];
        }
        meanfcX /= numNodesGlobal;
        bodyForce[0] = 2 * (momx - meanfcX);
        for (auto nodeNo : bulkNodes) {
            const std::valarray<lbBase_t> fTot = f(0, nodeNo) + f(1, nodeNo);
            const lbBase_t rho0Node = rho(0, nodeNo), rho1Node = rho(1, nodeNo);
            const lbBase_t rhoNode = rho0Node + rho1Node;
            const std::valarray<lbBase_t> velNode = calcVel<LT>(fTot, rhoNode, bodyForce);
            vel.set(0, nodeNo) = velNode;
            const lbBase_t tau = LT::c2Inv * rhoNode / (rho0Node * nu0Inv + rho1Node * nu1Inv) + 0.5;
            const lbBase_t uu = LT::dot(velNode, velNode);
            const std::valarray<lbBase_t> cu = LT::cDotAll(velNode);
            const std::valarray<lbBase_t> omegaBGK = calcOmegaBGK<LT>(fTot, tau, rhoNode, uu, cu);
            const lbBase_t uF = LT::dot(velNode, bodyForce);
            const std::valarray<lbBase_t> cF = LT::cDotAll(bodyForce);
            const std::valarray<lbBase_t> deltaOmegaF = calcDeltaOmegaF<LT>(tau, cu, uF, cF);
            std::valarray<lbBase_t> colorGradNode = grad(cgField, 0, nodeNo, grid);
            const lbBase_t CGNorm = vecNorm<LT>(colorGradNode);
            colorGradNode *= 1.0 / (CGNorm + (CGNorm < lbBaseEps));
            const std::valarray<lbBase_t> cCGNorm = LT::cDotAll(colorGradNode);
            const std::valarray<lbBase_t> deltaOmegaST = calcDeltaOmegaST<LT>(tau, sigma, CGNorm, cCGNorm);
            const std::valarray<lbBase_t> deltaOmegaRC = calcDeltaOmegaRC<LT>(beta, rho0Node, rho1Node, rhoNode, cCGNorm);
            const lbBase_t c0 = (rho0Node / rhoNode), c1 = (rho1Node / rhoNode);
            for (int q = 0; q < LT::nQ; ++q) {
                const int dst = grid.neighbor(q, nodeNo);
                fTmp(0, q, dst) = c0 * (fTot[q] + omegaBGK[q] + deltaOmegaF[q] + deltaOmegaST[q]) + deltaOmegaRC[q] + 0.0;
                fTmp(1, q, dst) = c1 * (fTot[q] + omegaBGK[q] + deltaOmegaF[q] + deltaOmegaST[q]) - deltaOmegaRC[q] + 0.0;
            }
        }
        f.swapData(fTmp);
        bounceBackBnd.apply(0, f, grid);
        bounceBackBnd.apply(1, f, grid);
    }
    writeOut(out, sz, f, rho, vel);
    return 0;
}

// prints calcDeltaOmegaQ, calcDeltaOmegaQTRT and calcfeq for one velocity (values the loops above do not reach)
template <typename LT>
int unitQ(char **a)
{
    const lbBase_t tau = std::atof(a[0]), tauSym = std::atof(a[1]), tauAnti = std::atof(a[2]), source = std::atof(a[3]), rho = std::atof(a[4]);
    std::valarray<lbBase_t> u(LT::nD);
    for (int d = 0; d < LT::nD; ++d) u[d] = std::atof(a[5 + d]);
    const std::valarray<lbBase_t> cu = LT::cDotAll(u);
    const lbBase_t u2 = LT::dot(u, u);
    const auto q = calcDeltaOmegaQ<LT>(tau, cu, u2, source);
    const auto qt = calcDeltaOmegaQTRT<LT>(tauSym, tauAnti, cu, u2, source);
    const auto feq = calcfeq<LT>(rho, u2, cu);
    for (int k = 0; k < LT::nQ; ++k) std::printf("%.17g %.17g %.17g\n", q[k], qt[k], feq[k]);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 2 && std::string(argv[1]) == "unitq") return std::string(argv[2]) == "D2Q9" ? unitQ<D2Q9>(argv + 3) : unitQ<D3Q19>(argv + 3);
    if (argc < 7) {
        std::cout << "usage: cpu_loop <std|trt|twophase> <D2Q9|D3Q19> <vtklb prefix> <out.bin> <steps> <params...>" << std::endl;
        return 2;
    }
    const std::string kind = argv[1], lattice = argv[2];
    const int steps = std::atoi(argv[5]);
    if (kind == "twophase") return lattice == "D2Q9" ? runTwoPhase<D2Q9>(argv[3], argv[4], steps, argv + 6) : runTwoPhase<D3Q19>(argv[3], argv[4], steps, argv + 6);
    const bool trt = kind == "trt";
    return lattice == "D2Q9" ? runSingle<D2Q9>(trt, argv[3], argv[4], steps, argv + 6) : runSingle<D3Q19>(trt, argv[3], argv[4], steps, argv + 6);
}
