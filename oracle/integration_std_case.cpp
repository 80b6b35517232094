// oracle/integration_std_case.cpp -- TEST INFRASTRUCTURE, compiled only where /root/reference exists.
//
// The reference's std_case main (src/std_case/main.cpp:41-157) with its node loop, swapData, communicateLbField
// and bounce-back apply switched to the GPU engine exactly as INTEGRATION.md section 2 shows: every object is
// the reference's OWN class (LBvtk, Grid, Nodes, BndMpi, ScalarField, VectorField, LbField, HalfWayBounceBack,
// Output) from the unmodified headers under /root/reference/src; the only additions are the binding stub
// include/reference_binding/LBgpu.h and the GpuLattice calls.  Paths and parameters come from the command line
// instead of input.dat (the reference main hard-codes "./../input/"), and the final f / rho / vel are also written
// as raw records so that tests/test_integration_stub.py can compare them with the golden dump of the reference's
// own CPU loop (tests/golden/std_d3q19_p1.npz).
//
//   integration_std_case <dir with tmp0.vtklb> <output dir> <nIterations> <nItrWrite> <tau> <Fx> <Fy> <Fz> [pressure|inletoutlet]
//
// The optional last argument adds the reference's own PressureBnd / InletOutlet object (LBpressurebnd.h; no main of
// the reference uses them) on every third fluid boundary node, with the prescribed values of ref_driver --pressure-bnd.
#include "LBSOLVER.h"
#include "IO.h"
#include "LBgpu.h"

#include <cstdio>
#include <fstream>

#ifndef LT
#define LT D3Q19
#endif

namespace {
void record(std::ofstream &ofs, const std::string &name, const double *p, std::uint64_t n)
{   // record format of oracle/recfile.py
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
    if (argc < 9) { std::cerr << "usage: " << argv[0] << " mpiDir outputDir nIterations nItrWrite tau Fx Fy Fz" << std::endl; return 2; }
#ifdef CHIMP_MPI_SHIM
    mpishim::init(1); // oracle/mpi_shim: one in-process rank (with a real MPI this line does not exist)
#endif
    MPI_Init(NULL, NULL);
    int nProcs;
    MPI_Comm_size(MPI_COMM_WORLD, &nProcs);
    int myRank;
    MPI_Comm_rank(MPI_COMM_WORLD, &myRank);
    const std::string mpiDir = std::string(argv[1]) + "/", outputDir = std::string(argv[2]) + "/";
    const int nIterations = std::atoi(argv[3]), nItrWrite = std::atoi(argv[4]);
    const lbBase_t tau = std::atof(argv[5]);

    // SETUP GRID AND GEOMETRY                                                   (unchanged, main.cpp:41-49)
    LBvtk<LT> vtklb(mpiDir + "tmp" + std::to_string(myRank) + ".vtklb");
    Grid<LT> grid(vtklb);
    Nodes<LT> nodes(vtklb, grid);
    BndMpi<LT> mpiBoundary(vtklb, nodes, grid);
    std::vector<int> bulkNodes = findBulkNodes(nodes);
    VectorField<LT> bodyForce(1, 1);
    for (int d = 0; d < LT::nD; ++d) bodyForce(0, d, 0) = std::atof(argv[6 + d]);

    // MACROSCOPIC FIELDS                                                        (unchanged, main.cpp:62-79)
    ScalarField rho(1, grid.size());
    vtklb.toAttribute("init_rho");
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) rho(0, n) = vtklb.getScalarAttribute<lbBase_t>();
    VectorField<LT> vel(1, grid.size());
    for (auto nodeNo : bulkNodes)
        for (int d = 0; d < LT::nD; ++d) vel(0, d, nodeNo) = 0.0;

    // SETUP BOUNDARY, LB FIELDS                                                 (unchanged, main.cpp:84-96)
    HalfWayBounceBack<LT> bounceBackBnd(findFluidBndNodes(nodes), nodes, grid);
    LbField<LT> f(1, grid.size());
    for (auto nodeNo : bulkNodes)
        for (int q = 0; q < LT::nQ; ++q) f(0, q, nodeNo) = LT::w[q] * rho(0, nodeNo);

    // OUTPUT VTK                                                                (unchanged, main.cpp:101-104)
    Output<LT> output(grid, bulkNodes, outputDir, myRank, nProcs);
    output.add_file("lb_run");
    output.add_scalar_variables({"rho"}, {rho});
    output.add_vector_variables({"vel"}, {vel});

    // ---- new: the engine takes over the node loop (replaces main.cpp:110-145)
    GpuLattice<LT> gpu(grid, bulkNodes, 1);
    gpu.add(bounceBackBnd);                          // replaces bounceBackBnd.apply(f, grid)
    if (argc > 9) {
        const std::vector<int> fluidBnd = findFluidBndNodes(nodes);
        std::vector<int> bndNodes;
        for (std::size_t k = 0; k < fluidBnd.size(); k += 3) bndNodes.push_back(fluidBnd[k]);
        if (std::string(argv[9]) == "pressure") {
            ScalarField rhoBnd(1, grid.size());
            for (int n = 0; n < grid.size(); ++n) rhoBnd(0, n) = 1.0 + 0.01 * (n % 7);
            PressureBnd<LT> pressureBnd(bndNodes, nodes, grid);
            gpu.add(pressureBnd, 0, grid, rhoBnd);    // replaces pressureBnd.apply(0, f, grid, rhoBnd)
        } else {
            InletOutlet<LT> inletOutlet(bndNodes, nodes, grid);
            gpu.add(inletOutlet, grid, 1.02, std::vector<lbBase_t>{0.01, -0.005, 0.002});
        }
    }
    gpu.finalize();
    gpu.upload(f);

    // MAIN LOOP: nItrWrite iterations per engine call, then the fields the reference main holds at that point
    for (int done = 0; done < nIterations;) {
        const int n = std::min(nItrWrite, nIterations - done);
        gpu.stepBGK(tau, bodyForce(0, 0), n);
        done += n;
        gpu.moments(rho, vel);
        output.write(done);                                                      // unchanged (main.cpp:150-155)
        if (myRank == 0) std::cout << "PLOT AT ITERATION : " << done << std::endl;
    }
    gpu.download(f);

    std::ofstream ofs(outputDir + "rank" + std::to_string(myRank) + ".rec", std::ios::binary);
    const std::string s = "step" + std::to_string(nIterations) + ".";
    record(ofs, s + "f", &f(0, 0, 0), (std::uint64_t)grid.size() * LT::nQ);
    record(ofs, s + "rho", &rho(0, 0), (std::uint64_t)grid.size());
    record(ofs, s + "vel", &vel(0, 0, 0), (std::uint64_t)grid.size() * LT::nD);
    MPI_Finalize();
    return 0;
}
