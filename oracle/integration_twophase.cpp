// oracle/integration_twophase.cpp -- TEST INFRASTRUCTURE, compiled only where /root/reference exists.
//
// The reference's colour-gradient main (src/twophase/main_TWOPHASE.cpp:73-455) with its three node sweeps, the scalar
// and population ghost exchanges and the bounce-back apply switched to the GPU engine through the binding stub of
// INTEGRATION.md (include/reference_binding/LBgpu.h).  Every object is the reference's OWN class from the unmodified
// headers; set-up follows the reference main line by line (density attributes read through `float`, wettability on the
// solid boundary nodes, initiateLbField).  Parameters come from the command line instead of input.dat; f, rho, vel,
// the colour field and the flux-controlled force are written as raw records for tests/test_integration_stub.py
// (golden: the reference's own CPU loop, tests/golden/twophase_d3q19_p1.npz).
//
//   integration_twophase <dir with tmp0.vtklb> <output dir> <nIterations> <tau0> <tau1> <sigma> <beta> <momx> <Fy> <Fz>
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
} // namespace

int main(int argc, char **argv)
{
    if (argc < 11) { std::cerr << "usage: " << argv[0] << " mpiDir outputDir nIterations tau0 tau1 sigma beta momx Fy Fz" << std::endl; return 2; }
#ifdef CHIMP_MPI_SHIM
    mpishim::init(1);
#endif
    MPI_Init(NULL, NULL);
    int nProcs;
    MPI_Comm_size(MPI_COMM_WORLD, &nProcs);
    int myRank;
    MPI_Comm_rank(MPI_COMM_WORLD, &myRank);
    const std::string mpiDir = std::string(argv[1]) + "/", outputDir = std::string(argv[2]) + "/";
    const int nIterations = std::atoi(argv[3]);
    const lbBase_t tau0 = std::atof(argv[4]), tau1 = std::atof(argv[5]), sigma = std::atof(argv[6]), beta = std::atof(argv[7]);
    const lbBase_t momx = std::atof(argv[8]);

    // SETUP GRID AND GEOMETRY, BOUNDARIES, BULK NODES                              (unchanged, main_TWOPHASE.cpp:73-87)
    LBvtk<LT> vtklb(mpiDir + "tmp" + std::to_string(myRank) + ".vtklb");
    Grid<LT> grid(vtklb);
    Nodes<LT> nodes(vtklb, grid);
    BndMpi<LT> mpiBoundary(vtklb, nodes, grid);
    HalfWayBounceBack<LT> bbBnd(findBulkNodes(nodes), nodes, grid);
    std::vector<int> solidBnd = findSolidBndNodes(nodes);
    std::vector<int> bulkNodes = findBulkNodes(nodes);
    VectorField<LT> bodyForce(1, 1);
    bodyForce(0, 0, 0) = 0.0;
    for (int d = 1; d < LT::nD; ++d) bodyForce(0, d, 0) = std::atof(argv[8 + d]);

    // SETUP LB FIELDS, MACROSCOPIC FIELDS, MASS DENSITIES                          (unchanged, :136-181)
    LbField<LT> f(2, grid.size());
    ScalarField rho(2, grid.size());
    VectorField<LT> vel(1, grid.size());
    ScalarField cgField(1, grid.size());
    vtklb.toAttribute("rho0");
    for (int nodeNo = vtklb.beginNodeNo(); nodeNo < vtklb.endNodeNo(); ++nodeNo) {
        float val = vtklb.getScalar<float>();
        rho(0, nodeNo) = val;
        for (int d = 0; d < LT::nD; ++d) vel(0, d, nodeNo) = 0.0;
    }
    vtklb.toAttribute("rho1");
    for (int nodeNo = vtklb.beginNodeNo(); nodeNo < vtklb.endNodeNo(); ++nodeNo) {
        float val = vtklb.getScalar<float>();
        rho(1, nodeNo) = val;
    }
    vtklb.toAttribute("wettability");
    for (int nodeNo = vtklb.beginNodeNo(); nodeNo < vtklb.endNodeNo(); ++nodeNo) {
        float val = vtklb.getScalar<float>();
        if (nodes.isSolidBoundary(nodeNo)) {
            rho(0, nodeNo) = val;
            rho(1, nodeNo) = 1 - val;
        }
    }
    int numNodes = bulkNodes.size();
    int numNodesGlobal;
    MPI_Allreduce(&numNodes, &numNodesGlobal, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    initiateLbField(0, 0, 0, bulkNodes, rho, vel, f);                               // (:200-203)
    initiateLbField(1, 1, 0, bulkNodes, rho, vel, f);

    // ---- new: the engine takes over the loop body (replaces :236-392)
    GpuLattice<LT> gpu(grid, bulkNodes, 2);
    gpu.add(bbBnd);                        // replaces bbBnd.apply(0 / 1, f, grid) (:390-391)
    gpu.setSolidBoundary(solidBnd);        // the rows whose colour is constant (:280-284)
    gpu.finalize();
    gpu.setTwoPhaseDensity(rho);
    gpu.upload(f);
    gpu.stepTwoPhase(tau0, tau1, sigma, beta, momx, bodyForce(0, 0), numNodesGlobal, nIterations);
    gpu.download(f);
    gpu.moments(rho, vel);
    gpu.phaseField(cgField);
    const double forceX = gpu.lastFluxForce();

    std::ofstream ofs(outputDir + "rank" + std::to_string(myRank) + ".rec", std::ios::binary);
    const std::string s = "step" + std::to_string(nIterations) + ".";
    record(ofs, s + "f", &f(0, 0, 0), (std::uint64_t)grid.size() * 2 * LT::nQ);
    record(ofs, s + "rho", &rho(0, 0), (std::uint64_t)grid.size() * 2);
    record(ofs, s + "vel", &vel(0, 0, 0), (std::uint64_t)grid.size() * LT::nD);
    record(ofs, s + "cg", &cgField(0, 0), (std::uint64_t)grid.size());
    record(ofs, s + "forceX", &forceX, 1);
    MPI_Finalize();
    return 0;
}
