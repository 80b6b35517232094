// twophase on the B200 engine: the structure of the reference's src/twophase/main_TWOPHASE.cpp:52-455
// (colour-gradient two-phase flow, two LbFields, flux-controlled body force) with the three node
// loops, swapData, both ghost exchanges and bbBnd.apply replaced by GpuLattice::stepTwoPhase.
//
//   twophase <D2Q9|D3Q19> <input.dat> <vtklb prefix> <first rank> <out.bin> [nRanksInProcess [vtk dir]]
//
// Reads the reference's files (input deck with <iterations> max/write and <fluid> tau a b / sigma /
// beta / momx / bodyforce, <prefix><rank>.vtklb with rho0 / rho1 / wettability attributes) and
// writes, per rank, raw f(2,size), rho(2,size), vel and cgField for the parity tests, plus the
// force.dat lines of main_TWOPHASE.cpp:427-437.
#include <cstdio>
#include <iomanip>
#include <memory>

#include "../chimp/LBSOLVER.h"

template <typename LT>
struct Rank {
    std::unique_ptr<LBvtk<LT>> vtklb;
    std::unique_ptr<Grid<LT>> grid;
    std::unique_ptr<Nodes<LT>> nodes;
    std::unique_ptr<BndMpi<LT>> mpiBoundary;
    std::vector<int> bulkNodes, solidBnd;
    std::unique_ptr<ScalarField> rho, cgField;
    std::unique_ptr<VectorField<LT>> vel;
    std::unique_ptr<LbField<LT>> f;
    std::unique_ptr<GpuLattice<LT>> gpu;
};

template <typename LT>
int run(const std::string &inputFile, const std::string &prefix, int firstRank, const std::string &outFile, int nRanks, const std::string &vtkDir)
{
    Input input(inputFile);
    auto &fluid = input["fluid"];
    const std::vector<double> bf = fluid["bodyforce"];
    std::valarray<lbBase_t> bodyForce(LT::nD);
    for (int d = 0; d < LT::nD; ++d) bodyForce[d] = bf[d];
    const int nIterations = input["iterations"]["max"];
    const int nWrite = std::max(1, int(input["iterations"]["write"]));
    const lbBase_t sigma = fluid["sigma"];
    const lbBase_t beta = fluid["beta"];
    const lbBase_t tau0 = fluid["tau"][0];
    const lbBase_t tau1 = fluid["tau"][1];
    const lbBase_t setMomX = fluid["momx"];

    std::vector<Rank<LT>> ranks(nRanks);
    int numNodesGlobal = 0;
    for (int r = 0; r < nRanks; ++r) {
        Rank<LT> &R = ranks[r];
        // SETUP GRID AND GEOMETRY (:73-87)
        R.vtklb.reset(new LBvtk<LT>(prefix + std::to_string(firstRank + r) + ".vtklb"));
        R.grid.reset(new Grid<LT>(*R.vtklb));
        R.nodes.reset(new Nodes<LT>(*R.vtklb, *R.grid));
        R.mpiBoundary.reset(new BndMpi<LT>(*R.vtklb, *R.nodes, *R.grid, prefix));
        HalfWayBounceBack<LT> bbBnd(findBulkNodes(*R.nodes), *R.nodes, *R.grid);
        R.solidBnd = findSolidBndNodes(*R.nodes);
        R.bulkNodes = findBulkNodes(*R.nodes);
        // SETUP MACROSCOPIC FIELDS (:140-181); the density attributes pass through float there
        const int sz = R.grid->size();
        R.rho.reset(new ScalarField(2, sz));
        R.vel.reset(new VectorField<LT>(1, sz));
        R.cgField.reset(new ScalarField(1, sz));
        R.vtklb->toAttribute("rho0");
        for (int n = R.vtklb->beginNodeNo(); n < R.vtklb->endNodeNo(); ++n) (*R.rho)(0, n) = R.vtklb->template getScalar<float>();
        R.vtklb->toAttribute("rho1");
        for (int n = R.vtklb->beginNodeNo(); n < R.vtklb->endNodeNo(); ++n) (*R.rho)(1, n) = R.vtklb->template getScalar<float>();
        R.vtklb->toAttribute("wettability");
        for (int n = R.vtklb->beginNodeNo(); n < R.vtklb->endNodeNo(); ++n) {
            const float val = R.vtklb->template getScalar<float>();
            if (R.nodes->isSolidBoundary(n)) {
                (*R.rho)(0, n) = val;
                (*R.rho)(1, n) = 1 - val;
            }
        }
        numNodesGlobal += int(R.bulkNodes.size()); // MPI_Allreduce at :196-198
        // INITIATE LB FIELDS (:202-205): equilibrium at u = 0
        R.f.reset(new LbField<LT>(2, sz));
        for (int fld = 0; fld < 2; ++fld)
            for (auto nodeNo : R.bulkNodes)
                for (int q = 0; q < LT::nQ; ++q)
                    (*R.f)(fld, q, nodeNo) = LT::w[q] * (*R.rho)(fld, nodeNo) * (1.0 + LT::c2Inv * 0.0 + LT::c4Inv0_5 * (0.0 * 0.0 - LT::c2 * 0.0));
        // hand the objects to the engine
        R.gpu.reset(new GpuLattice<LT>(*R.grid, R.bulkNodes, 2));
        R.gpu->add(*R.mpiBoundary);
        R.gpu->add(bbBnd);
        R.gpu->setSolidBoundary(R.solidBnd);
        R.gpu->finalize();
        R.gpu->setTwoPhaseDensity(*R.rho);
        R.gpu->upload(*R.f);
    }
    std::vector<chimp_lattice *> handles;
    std::vector<int> rankOf;
    for (int r = 0; r < nRanks; ++r) { handles.push_back(ranks[r].gpu->handle()); rankOf.push_back(firstRank + r); }
    InProcessRanks world(handles, rankOf);

    // MAIN LOOP (:236-452): the reference runs i = 0..nIterations and writes when i % nWrite == 0; the engine
    // advances to the next write iteration in one call and hands back the fields the reference holds there
    const std::string forceFile = outFile + ".force.dat";
    std::remove(forceFile.c_str());
    for (int i = 0; i <= nIterations;) {
        const int next = (i % nWrite == 0) ? i : std::min((i / nWrite + 1) * nWrite, nIterations);
        const int chunk = next - i + 1;
        world.run([&](int r) { ranks[r].gpu->stepTwoPhase(tau0, tau1, sigma, beta, setMomX, bodyForce, numNodesGlobal, chunk); });
        i = next + 1;
        if (next % nWrite == 0) {
            std::cout << "PLOT AT ITERATION : " << next << std::endl;
            std::ofstream ofs(forceFile, std::ios::app); // :427-437
            ofs << next << " " << std::setprecision(23) << ranks[0].gpu->lastFluxForce() << std::endl;
        }
    }

    FILE *fp = std::fopen(outFile.c_str(), "wb");
    if (!fp) chimp_host::die("cannot open " + outFile);
    for (auto &R : ranks) {
        R.gpu->download(*R.f);
        R.gpu->download(*R.rho, *R.vel);
        R.gpu->downloadPhaseField(*R.cgField);
        const int sz = R.grid->size();
        std::fwrite(&sz, sizeof(int), 1, fp);
        std::fwrite(R.f->data(), sizeof(double), std::size_t(sz) * 2 * LT::nQ, fp);
        std::fwrite(R.rho->data(), sizeof(double), std::size_t(sz) * 2, fp);
        std::fwrite(R.vel->data(), sizeof(double), std::size_t(sz) * LT::nD, fp);
        std::fwrite(R.cgField->data(), sizeof(double), sz, fp);
    }
    std::fclose(fp);
    if (!vtkDir.empty()) // OUTPUT VTK (main_TWOPHASE.cpp:214-223, 424)
        for (int r = 0; r < nRanks; ++r) {
            Rank<LT> &R = ranks[r];
            Output<LT> output(*R.grid, R.bulkNodes, vtkDir, firstRank + r, nRanks);
            output.add_file("fluid");
            output.add_scalar_variables({"rho"}, {*R.rho});
            output.add_vector_variables({"vel"}, {*R.vel});
            std::vector<int> geo(R.grid->size(), -1); // Nodes::geo (LBnodes.h:93-99)
            for (int n = R.vtklb->beginNodeNo(); n < R.vtklb->endNodeNo(); ++n) geo[n] = R.nodes->isSolid(n) ? 1 : 0;
            Output<LT, int> geoout(R.grid->pos(), vtkDir, firstRank + r, nRanks, "geo", geo);
            geoout.write();
            output.write(nIterations + 1);
        }
    std::cout << "twophase: " << nIterations + 1 << " iterations on " << nRanks << " rank(s) done" << std::endl;
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 6) {
        std::cout << "usage: twophase <D2Q9|D3Q19> <input.dat> <vtklb prefix> <first rank> <out.bin> [nRanksInProcess [vtk dir]]" << std::endl;
        return 2;
    }
    const std::string lattice = argv[1];
    const int rank = std::atoi(argv[4]);
    const int nRanks = argc > 6 ? std::atoi(argv[6]) : 1;
    const std::string vtkDir = argc > 7 ? argv[7] : "";
    if (lattice == "D2Q9") return run<D2Q9>(argv[2], argv[3], rank, argv[5], nRanks, vtkDir);
    if (lattice == "D3Q19") return run<D3Q19>(argv[2], argv[3], rank, argv[5], nRanks, vtkDir);
    chimp_host::die("twophase needs D2Q9 or D3Q19 (D3Q27 has no colour-gradient weights)");
}
