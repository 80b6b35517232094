// std_one_phase on the B200 engine: the structure of the reference's src/std_one_phase/main.cpp:205-652
// (single-phase flow in the pore space of one fluid of a two-fluid configuration: masked body force,
// mass-conservation source per interior domain, solid / pressure / fluid-fluid link boundaries) with
// the mass-change loop, the node loop, swapData, the ghost exchanges and the three apply*Boundary
// calls replaced by GpuLattice::stepBGK / stepTRT, and the mass-flux sum formed on the device.
//
//   std_one_phase <input.dat> <vtklb prefix> <first rank> <out.bin> [nRanksInProcess]
//
// Input deck: <iterations> max / write, <fluid> tau (or tausym + tauanti) / bodyforce / rhow (the
// reference hard-codes the pressure-boundary density 1.0 at main.cpp:593).  Geometry files carry the
// attributes nodetags, domains, force, interior_domains, normal_x/y/z (main.cpp:253-330).
// Writes per rank raw f, rho, vel and massChange, and <out.bin>.flux in the format of main.cpp:620-630.
#include <cstdio>
#include <memory>

#include "../chimp/LBSOLVER.h"

#define LT D3Q19

// findSolidFluidLinks / findPressureFluidLinks / findFluidFluidLinks (main.cpp:27-126): node carries tag
// bit `bit`, the neighbour's (tag & 3) equals wantTag; link = {nodeFluid, qUnknown, nodeWall, qKnown}
static std::vector<std::vector<int>> tagLinks(const Nodes<LT> &nodes, const Grid<LT> &grid, int bit, int wantTag, int needPhase)
{
    std::vector<std::vector<int>> ret;
    for (int n = 1; n < nodes.size(); n++) {
        const int flagged = (nodes.getTag(n) >> bit) & 1;
        if (!flagged || !nodes.isMyRank(n)) continue;
        if (needPhase >= 0 && (nodes.getTag(n) & 3) != needPhase) continue;
        for (int q = 0; q < LT::nQNonZero_; ++q) {
            const int neigNode = grid.neighbor(q, n);
            if ((nodes.getTag(neigNode) & 3) == wantTag) ret.push_back({n, LT::reverseDirection(q), neigNode, q});
        }
    }
    return ret;
}

struct Rank {
    std::unique_ptr<LBvtk<LT>> vtklb;
    std::unique_ptr<Grid<LT>> grid;
    std::unique_ptr<Nodes<LT>> nodes;
    std::unique_ptr<BndMpi<LT>> mpiBoundary;
    std::vector<int> bulkNodes, interiorDomainsLabel, pressureFluidNodes, fluidPhase;
    std::vector<lbBase_t> addMassSource;
    std::unique_ptr<ScalarField> rho, forceOn;
    std::unique_ptr<VectorField<LT>> vel;
    std::unique_ptr<LbField<LT>> f;
    std::unique_ptr<GpuLattice<LT>> gpu;
    std::vector<std::vector<int>> solidFluidLinks, pressureFluidLinks, fluidFluidLinks;
};

int main(int argc, char **argv)
{
    if (argc < 5) {
        std::cout << "usage: std_one_phase <input.dat> <vtklb prefix> <first rank> <out.bin> [nRanksInProcess]" << std::endl;
        return 2;
    }
    const std::string prefix = argv[2], outFile = argv[4];
    const int firstRank = std::atoi(argv[3]);
    const int nRanks = argc > 5 ? std::atoi(argv[5]) : 1;
    Input input(argv[1]);
    const int nIterations = input["iterations"]["max"];
    const int nItrWrite = std::max(1, int(input["iterations"]["write"]));
    const bool trt = input["fluid"].child_.count("tausym") != 0;
    const lbBase_t tau = trt ? 0.0 : lbBase_t(input["fluid"]["tau"]);
    const lbBase_t tauSym = trt ? lbBase_t(input["fluid"]["tausym"]) : 0.0;
    const lbBase_t tauAnti = trt ? lbBase_t(input["fluid"]["tauanti"]) : 0.0;
    const lbBase_t rhoW = input["fluid"].child_.count("rhow") ? lbBase_t(input["fluid"]["rhow"]) : 1.0;
    const std::vector<double> bf = input["fluid"]["bodyforce"];
    std::valarray<lbBase_t> bodyForce(LT::nD);
    for (int d = 0; d < LT::nD; ++d) bodyForce[d] = bf[d];

    std::vector<Rank> ranks(nRanks);
    int globalDomainLabelMax = 0;
    for (int r = 0; r < nRanks; ++r) {
        Rank &R = ranks[r];
        // SETUP GRID AND GEOMETRY (:244-249)
        R.vtklb.reset(new LBvtk<LT>(prefix + std::to_string(firstRank + r) + ".vtklb"));
        R.grid.reset(new Grid<LT>(*R.vtklb));
        R.nodes.reset(new Nodes<LT>(*R.vtklb, *R.grid));
        R.mpiBoundary.reset(new BndMpi<LT>(*R.vtklb, *R.nodes, *R.grid, prefix));
        R.bulkNodes = findBulkNodes(*R.nodes);
        const int sz = R.grid->size();
        // node tags, force indicator, interior domains (:260-303)
        R.vtklb->toAttribute("nodetags");
        for (int n = R.vtklb->beginNodeNo(); n < R.vtklb->endNodeNo(); ++n) R.nodes->setTag(R.vtklb->getScalarAttribute<int>(), n);
        R.forceOn.reset(new ScalarField(1, sz));
        R.vtklb->toAttribute("force");
        for (int n = R.vtklb->beginNodeNo(); n < R.vtklb->endNodeNo(); ++n) (*R.forceOn)(0, n) = R.vtklb->getScalarAttribute<int>();
        R.interiorDomainsLabel.assign(sz, 0);
        R.addMassSource.assign(sz, 0.0);
        R.vtklb->toAttribute("interior_domains");
        for (int n = R.vtklb->beginNodeNo(); n < R.vtklb->endNodeNo(); ++n) {
            const int val = R.vtklb->getScalarAttribute<int>();
            R.interiorDomainsLabel[n] = val;
            if (R.nodes->isMyRank(n) && val > globalDomainLabelMax) globalDomainLabelMax = val; // MPI_Allreduce(MAX) at :332
        }
    }
    // number of mass sources per interior domain (:337-361), summed over the ranks
    std::vector<lbBase_t> massSourceScaleFactor(globalDomainLabelMax + 1, 0);
    for (auto &R : ranks)
        for (int n = 1; n < R.grid->size(); ++n)
            if (R.nodes->isMyRank(n)) {
                const int label = R.interiorDomainsLabel[n];
                if (label > 0 && R.nodes->getTag(n) < 3) {
                    massSourceScaleFactor[label] += 1;
                    R.addMassSource[n] = 1.0;
                }
            }
    for (int i = 1; i < globalDomainLabelMax + 1; ++i) {
        if (massSourceScaleFactor[i] == 0.0) chimp_host::die("Interior domain " + std::to_string(i) + " has no interior nodes!");
        massSourceScaleFactor[i] = 1.0 / massSourceScaleFactor[i];
    }
    for (int r = 0; r < nRanks; ++r) {
        Rank &R = ranks[r];
        const int sz = R.grid->size();
        // pressure-boundary nodes for the mass flux (:80-93, :362) and the phase each belongs to (:610-615)
        for (int n = 1; n < R.nodes->size(); n++)
            if (((R.nodes->getTag(n) >> 4) & 1) && R.nodes->isMyRank(n)) {
                const int fluidPhase = (R.nodes->getTag(n) & 3) - 1;
                if (fluidPhase != 0 && fluidPhase != 1) chimp_host::die("Fluid phase = " + std::to_string(fluidPhase) + " in write mass flux");
                R.pressureFluidNodes.push_back(n);
                R.fluidPhase.push_back(fluidPhase);
            }
        // link boundaries (:418-433)
        R.solidFluidLinks = tagLinks(*R.nodes, *R.grid, 3, 0, -1);
        R.pressureFluidLinks = tagLinks(*R.nodes, *R.grid, 4, 3, -1);
        R.fluidFluidLinks = tagLinks(*R.nodes, *R.grid, 2, 2, 1);
        // macroscopic and lb fields (:441-474): rho = 1, u = 0, f = feq
        R.rho.reset(new ScalarField(1, sz));
        R.vel.reset(new VectorField<LT>(1, sz));
        R.f.reset(new LbField<LT>(1, sz));
        for (int n = R.vtklb->beginNodeNo(); n < R.vtklb->endNodeNo(); ++n) (*R.rho)(0, n) = 1.0;
        for (auto nodeNo : R.bulkNodes)
            for (int q = 0; q < LT::nQ; ++q)
                (*R.f)(0, q, nodeNo) = LT::w[q] * (*R.rho)(0, nodeNo) * (1 + LT::c2Inv * 0.0 + LT::c4Inv0_5 * (0.0 * 0.0 - LT::c2 * 0.0));
        // hand the objects to the engine
        R.gpu.reset(new GpuLattice<LT>(*R.grid, R.bulkNodes, 1));
        R.gpu->add(*R.mpiBoundary);
        R.gpu->addLinks(CHIMP_LINK_SOLID, R.solidFluidLinks);
        R.gpu->addLinks(CHIMP_LINK_PRESSURE, R.pressureFluidLinks);
        R.gpu->addLinks(CHIMP_LINK_FLUID_SWAP, R.fluidFluidLinks);
        R.gpu->finalize();
        R.gpu->setOnePhaseAttributes(*R.forceOn, R.interiorDomainsLabel, R.addMassSource, massSourceScaleFactor, rhoW);
        R.gpu->upload(*R.f);
    }
    std::vector<chimp_lattice *> handles;
    std::vector<int> rankOf;
    for (int r = 0; r < nRanks; ++r) { handles.push_back(ranks[r].gpu->handle()); rankOf.push_back(firstRank + r); }
    InProcessRanks world(handles, rankOf);

    // MAIN LOOP (:513-641): i = 0..nIterations, output when i % nItrWrite == 0
    std::vector<lbBase_t> oldMassFlux(2, 0.0);
    const std::string fluxFile = outFile + ".flux";
    std::remove(fluxFile.c_str());
    for (int i = 0; i <= nIterations;) {
        const int next = (i % nItrWrite == 0) ? i : std::min((i / nItrWrite + 1) * nItrWrite, nIterations);
        const int chunk = next - i + 1;
        world.run([&](int r) {
            if (trt) ranks[r].gpu->stepTRT(tauSym, tauAnti, bodyForce, chunk);
            else ranks[r].gpu->stepBGK(tau, bodyForce, chunk);
        });
        i = next + 1;
        if (next % nItrWrite != 0) continue;
        // WRITE TO FILE (:602-633): mass flux through the pressure-boundary nodes
        std::vector<lbBase_t> massFluxGlobal(2, 0.0);
        for (auto &R : ranks) { // MPI_Allreduce(SUM) at :619, rank order
            const std::vector<lbBase_t> local = R.gpu->massFlux(R.pressureFluidNodes, R.fluidPhase);
            massFluxGlobal[0] = massFluxGlobal[0] + local[0];
            massFluxGlobal[1] = massFluxGlobal[1] + local[1];
        }
        const lbBase_t q1 = 0.5 * massFluxGlobal[0];
        const lbBase_t q2 = 0.5 * massFluxGlobal[1];
        const lbBase_t q1_change = (q1 - oldMassFlux[0]) / (q1 + 1e-15);
        const lbBase_t q2_change = (q2 - oldMassFlux[1]) / (q2 + 1e-15);
        std::cout << "PLOT AT ITERATION: " << next << std::endl;
        std::cout << "q1 = " << q1 << " (" << q1_change << ")" << std::endl;
        std::cout << "q2 = " << q2 << " (" << q2_change << ")" << std::endl;
        std::ofstream myfile(fluxFile, std::ios::out | std::ios::app);
        myfile << "PLOT AT ITERATION: " << next << "\n";
        myfile << "q1 = " << q1 << " (" << q1_change << ")" << "\n";
        myfile << "q2 = " << q2 << " (" << q2_change << ")" << "\n";
        oldMassFlux[0] = q1;
        oldMassFlux[1] = q2;
    }

    FILE *fp = std::fopen(outFile.c_str(), "wb");
    if (!fp) chimp_host::die("cannot open " + outFile);
    for (auto &R : ranks) {
        R.gpu->download(*R.f);
        R.gpu->download(*R.rho, *R.vel);
        const std::vector<lbBase_t> mass = R.gpu->massChange(globalDomainLabelMax + 1);
        const int sz = R.grid->size(), nl = globalDomainLabelMax + 1;
        std::fwrite(&sz, sizeof(int), 1, fp);
        std::fwrite(R.f->data(), sizeof(double), std::size_t(sz) * LT::nQ, fp);
        std::fwrite(R.rho->data(), sizeof(double), sz, fp);
        std::fwrite(R.vel->data(), sizeof(double), std::size_t(sz) * LT::nD, fp);
        std::fwrite(&nl, sizeof(int), 1, fp);
        std::fwrite(mass.data(), sizeof(double), nl, fp);
    }
    std::fclose(fp);
    std::cout << "std_one_phase: " << nIterations + 1 << " iterations on " << nRanks << " rank(s) done" << std::endl;
    return 0;
}
