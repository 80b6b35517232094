// oracle/integration_one_phase.cpp -- TEST INFRASTRUCTURE, compiled only where /root/reference exists.
//
// The reference's std_one_phase main (src/std_one_phase/main.cpp:226-652) with its mass-conservation sweep, node loop,
// swapData, ghost exchanges and the three link-boundary applies switched to the GPU engine through the binding stub of
// INTEGRATION.md (include/reference_binding/LBgpu.h).  Objects are the reference's OWN classes from the unmodified
// headers; the link finders and the attribute set-up restate what that main does in its own file (main.cpp:27-126,
// 253-345 -- they are not part of the reference's headers).  Parameters come from the command line; f, rho, vel, the
// mass change per interior domain and the mass flux through the pressure nodes are written as raw records for
// tests/test_integration_stub.py (golden: the reference's own CPU loop, tests/golden/onephase_d3q19_p1.npz).
//
//   integration_one_phase <dir with tmp0.vtklb> <output dir> <nIterations> <bgk|trt> <tau | tauSym> <0 | tauAnti> <Fx> <Fy> <Fz> <rhoW>
#include "LBSOLVER.h"
#include "IO.h"
#include "LBgpu.h"

#include <fstream>

#ifndef LT
#define LT D3Q19
#endif

namespace {
void record(std::ofstream &ofs, const std::string &name, const double *p, std::uint64_t n)
{
    ofs.write("REC1", 4);
    std::uint32_t l = (std::uint32_t)name.size();
    ofs.write((const char *)&l, 4);
    ofs.write(name.data(), l);
    const char t = 'd';
    ofs.write(&t, 1);
    ofs.write((const char *)&n, 8);
    ofs.write((const char *)p, 8 * n);
}

// links {nodeFluid, qUnknown, nodeWall, qKnown} from node tags (main.cpp:27-126): the node carries tag bit `bit`
// (3 solid interface, 4 pressure, 2 fluid-fluid for phase-1 nodes), its neighbour's two low tag bits equal `want`
template <typename DXQY>
std::vector<std::vector<int>> findLinks(const Nodes<DXQY> &nodes, const Grid<DXQY> &grid, int bit, int want, int needPhase)
{
    std::vector<std::vector<int>> ret;
    for (int n = 1; n < nodes.size(); n++) {
        const int isBoundary = (nodes.getTag(n) >> bit) & 1;
        if (!isBoundary || !nodes.isMyRank(n)) continue;
        if (needPhase >= 0 && (nodes.getTag(n) & 3) != needPhase) continue;
        for (int q = 0; q < DXQY::nQNonZero_; ++q) {
            const int nn = grid.neighbor(q, n);
            if ((nodes.getTag(nn) & 3) == want) ret.push_back({n, DXQY::reverseDirection(q), nn, q});
        }
    }
    return ret;
}
} // namespace

int main(int argc, char **argv)
{
    if (argc < 11) { std::cerr << "usage: " << argv[0] << " mpiDir outputDir nIterations bgk|trt tau tauAnti Fx Fy Fz rhoW" << std::endl; return 2; }
#ifdef CHIMP_MPI_SHIM
    mpishim::init(1);
#endif
    MPI_Init(NULL, NULL);
    int myRank;
    MPI_Comm_rank(MPI_COMM_WORLD, &myRank);
    const std::string mpiDir = std::string(argv[1]) + "/", outputDir = std::string(argv[2]) + "/";
    const int nIterations = std::atoi(argv[3]);
    const bool trt = std::string(argv[4]) == "trt";
    const lbBase_t tau = std::atof(argv[5]), tauAnti = std::atof(argv[6]), rhoW = std::atof(argv[10]);

    // GRID, NODES, TAGS                                                            (main.cpp:226-262)
    LBvtk<LT> vtklb(mpiDir + "tmp" + std::to_string(myRank) + ".vtklb");
    Grid<LT> grid(vtklb);
    Nodes<LT> nodes(vtklb, grid);
    BndMpi<LT> mpiBoundary(vtklb, nodes, grid);
    std::vector<int> bulkNodes = findBulkNodes(nodes);
    vtklb.toAttribute("nodetags");
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) nodes.setTag(vtklb.getScalarAttribute<int>(), n);
    // force indicator, interior domains, mass-source markers                       (main.cpp:279-345)
    ScalarField forceOn(1, grid.size());
    vtklb.toAttribute("force");
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) forceOn(0, n) = vtklb.getScalarAttribute<int>();
    std::vector<int> interiorDomainsLabel(grid.size(), 0);
    std::vector<lbBase_t> addMassSource(grid.size(), 0.0);
    vtklb.toAttribute("interior_domains");
    int localDomainLabelMax = 0;
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) {
        const int val = vtklb.getScalarAttribute<int>();
        interiorDomainsLabel[n] = val;
        if (nodes.isMyRank(n) && val > localDomainLabelMax) localDomainLabelMax = val;
    }
    int globalDomainLabelMax;
    MPI_Allreduce(&localDomainLabelMax, &globalDomainLabelMax, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
    std::vector<lbBase_t> massSourceScaleFactor(globalDomainLabelMax + 1, 0);
    {
        std::vector<lbBase_t> tmp(globalDomainLabelMax + 1, 0.0);
        for (int n = 1; n < grid.size(); ++n)
            if (nodes.isMyRank(n) && interiorDomainsLabel[n] > 0 && nodes.getTag(n) < 3) {
                tmp[interiorDomainsLabel[n]] += 1;
                addMassSource[n] = 1.0;
            }
        MPI_Allreduce(tmp.data(), massSourceScaleFactor.data(), globalDomainLabelMax + 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
        for (int i = 1; i < globalDomainLabelMax + 1; ++i) massSourceScaleFactor[i] = 1.0 / massSourceScaleFactor[i];
    }
    // boundary links and the pressure nodes whose mass flux is reported            (main.cpp:27-126, 347-360)
    const auto solidFluidLinks = findLinks(nodes, grid, 3, 0, -1);
    const auto pressureFluidLinks = findLinks(nodes, grid, 4, 3, -1);
    const auto fluidFluidLinks = findLinks(nodes, grid, 2, 2, 1);
    std::vector<int> pressureFluidNodes, fluidPhase;
    for (int n = 1; n < nodes.size(); n++)
        if (((nodes.getTag(n) >> 4) & 1) && nodes.isMyRank(n)) {
            pressureFluidNodes.push_back(n);
            fluidPhase.push_back((nodes.getTag(n) & 3) - 1);
        }
    VectorField<LT> bodyForce(1, 1);
    for (int d = 0; d < LT::nD; ++d) bodyForce(0, d, 0) = std::atof(argv[7 + d]);

    // FIELDS: rho = 1, u = 0, f = calcfeq                                         (main.cpp:441-474)
    ScalarField rho(1, grid.size());
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) rho(0, n) = 1.0;
    VectorField<LT> vel(1, grid.size());
    LbField<LT> f(1, grid.size());
    for (auto nodeNo : bulkNodes) {
        vel.set(0, nodeNo) = 0;
        const auto u2 = LT::dot(vel(0, nodeNo), vel(0, nodeNo));
        const auto cu = LT::cDotAll(vel(0, nodeNo));
        f.set(0, nodeNo) = calcfeq<LT>(rho(0, nodeNo), u2, cu);
    }

    // ---- new: the engine takes over the loop (replaces main.cpp:513-597)
    GpuLattice<LT> gpu(grid, bulkNodes, 1);
    gpu.addLinks(CHIMP_LINK_SOLID, solidFluidLinks);           // applySolidFluidBoundary       (:591)
    gpu.addLinks(CHIMP_LINK_PRESSURE, pressureFluidLinks);     // applyPressureFluidBoundary    (:593)
    gpu.addLinks(CHIMP_LINK_FLUID_SWAP, fluidFluidLinks);      // applyFluidFluidBoundary       (:597)
    gpu.finalize();
    gpu.setOnePhaseAttributes(forceOn, interiorDomainsLabel, addMassSource, massSourceScaleFactor, rhoW);
    gpu.upload(f);
    if (trt) gpu.stepTRT(tau, tauAnti, bodyForce(0, 0), nIterations);
    else gpu.stepBGK(tau, bodyForce(0, 0), nIterations);
    gpu.download(f);
    gpu.moments(rho, vel);
    std::vector<lbBase_t> massChange(globalDomainLabelMax + 1, 0.0);
    chimpCheck(chimp_download_mass_change(gpu.handle(), massChange.data()));
    const std::vector<lbBase_t> massFluxLocal = gpu.massFlux(pressureFluidNodes, fluidPhase);   // (:606-618)

    std::ofstream ofs(outputDir + "rank" + std::to_string(myRank) + ".rec", std::ios::binary);
    const std::string s = "step" + std::to_string(nIterations) + ".";
    record(ofs, s + "f", &f(0, 0, 0), (std::uint64_t)grid.size() * LT::nQ);
    record(ofs, s + "rho", &rho(0, 0), (std::uint64_t)grid.size());
    record(ofs, s + "vel", &vel(0, 0, 0), (std::uint64_t)grid.size() * LT::nD);
    record(ofs, s + "massChange", massChange.data(), massChange.size());
    record(ofs, s + "massFlux", massFluxLocal.data(), massFluxLocal.size());
    MPI_Finalize();
    return 0;
}
