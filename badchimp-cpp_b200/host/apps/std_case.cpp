// std_case on the B200 engine: the structure of the reference's src/std_case/main.cpp:17-160 with
// the per-node loop, swapData, communicateLbField and bounceBackBnd.apply replaced by one call.
//
//   std_case <lattice D2Q9|D3Q19|D3Q27> <input.dat> <vtklb prefix> <rank> <out.bin> [nRanksInProcess [vtk dir|- [pressure|inletoutlet]]]
//
// The last argument adds the library's PressureBnd / InletOutlet (LBpressurebnd.h) on every third fluid boundary node,
// with the prescribed values of oracle/ref_driver --pressure-bnd (tests/test_library_bnd.py).
//
// Reads the same files as the reference main (input deck, <prefix><rank>.vtklb), runs
// iterations/max iterations and writes raw f (LbField layout), rho and vel for the parity tests.
// With nRanksInProcess > 1 all ranks of the decomposition run in this process on one GPU and the
// halos are moved with device copies between the engine's pack / unpack halves (no MPI needed).
#include <cuda_runtime_api.h>
#include <cstdio>
#include <memory>

#include "../chimp/LBSOLVER.h"

template <typename LT>
struct Rank {
    std::unique_ptr<LBvtk<LT>> vtklb;
    std::unique_ptr<Grid<LT>> grid;
    std::unique_ptr<Nodes<LT>> nodes;
    std::unique_ptr<BndMpi<LT>> mpiBoundary;
    std::vector<int> bulkNodes;
    std::unique_ptr<ScalarField> rho;
    std::unique_ptr<VectorField<LT>> vel;
    std::unique_ptr<LbField<LT>> f;
    std::unique_ptr<GpuLattice<LT>> gpu;
};

template <typename LT>
int run(const std::string &inputFile, const std::string &prefix, int firstRank, const std::string &outFile, int nRanks, const std::string &vtkDir,
        const std::string &libraryBnd)
{
    Input input(inputFile);
    const int nIterations = input["iterations"]["max"];
    const lbBase_t tau = input["fluid"]["tau"];
    const std::vector<double> F = input["fluid"]["bodyforce"];
    std::valarray<lbBase_t> bodyForce(LT::nD);
    for (int d = 0; d < LT::nD; ++d) bodyForce[d] = F[d];

    std::vector<Rank<LT>> ranks(nRanks);
    for (int r = 0; r < nRanks; ++r) {
        Rank<LT> &R = ranks[r];
        // SETUP GRID AND GEOMETRY (std_case/main.cpp:41-47)
        R.vtklb.reset(new LBvtk<LT>(prefix + std::to_string(firstRank + r) + ".vtklb"));
        R.grid.reset(new Grid<LT>(*R.vtklb));
        R.nodes.reset(new Nodes<LT>(*R.vtklb, *R.grid));
        R.mpiBoundary.reset(new BndMpi<LT>(*R.vtklb, *R.nodes, *R.grid, prefix));
        R.bulkNodes = findBulkNodes(*R.nodes);
        // MACROSCOPIC FIELDS (:62-80)
        R.rho.reset(new ScalarField(1, R.grid->size()));
        R.vtklb->toAttribute("init_rho");
        for (int n = R.vtklb->beginNodeNo(); n < R.vtklb->endNodeNo(); ++n) (*R.rho)(0, n) = R.vtklb->template getScalarAttribute<lbBase_t>();
        R.vel.reset(new VectorField<LT>(1, R.grid->size()));
        // BOUNDARY + LB FIELDS (:84-96)
        HalfWayBounceBack<LT> bounceBackBnd(findFluidBndNodes(*R.nodes), *R.nodes, *R.grid);
        R.f.reset(new LbField<LT>(1, R.grid->size()));
        for (auto nodeNo : R.bulkNodes)
            for (int q = 0; q < LT::nQ; ++q) (*R.f)(0, q, nodeNo) = LT::w[q] * (*R.rho)(0, nodeNo);
        // hand the objects to the engine
        R.gpu.reset(new GpuLattice<LT>(*R.grid, R.bulkNodes, 1));
        R.gpu->add(*R.mpiBoundary);
        R.gpu->add(bounceBackBnd);
        if (!libraryBnd.empty()) {
            const std::vector<int> fluidBnd = findFluidBndNodes(*R.nodes);
            std::vector<int> bndNodes;
            for (std::size_t k = 0; k < fluidBnd.size(); k += 3) bndNodes.push_back(fluidBnd[k]);
            if (libraryBnd == "pressure") {
                ScalarField rhoBnd(1, R.grid->size());
                for (int n = 0; n < R.grid->size(); ++n) rhoBnd(0, n) = 1.0 + 0.01 * (n % 7);
                PressureBnd<LT> pressureBnd(bndNodes, *R.nodes, *R.grid);
                R.gpu->add(pressureBnd, 0, *R.grid, rhoBnd);
            } else {
                const std::vector<lbBase_t> v{0.01, -0.005, 0.002};
                InletOutlet<LT> inletOutlet(bndNodes, *R.nodes, *R.grid);
                R.gpu->add(inletOutlet, *R.grid, 1.02, std::vector<lbBase_t>(v.begin(), v.begin() + LT::nD));
            }
        }
        R.gpu->finalize();
        R.gpu->upload(*R.f);
    }

    // MAIN LOOP (:109-146)
    if (nRanks == 1 && ranks[0].mpiBoundary->lists().empty()) {
        ranks[0].gpu->stepBGK(tau, bodyForce, nIterations);
    } else {
        chimp_single_params p{};
        p.collision = CHIMP_BGK;
        p.tau = tau;
        for (int d = 0; d < LT::nD; ++d) p.force[d] = bodyForce[d];
        for (int i = 0; i < nIterations; ++i) {
            for (auto &R : ranks) chimpCheck(chimp_step_begin(R.gpu->handle(), &p, i == nIterations - 1));
            for (auto &R : ranks) chimpCheck(chimp_synchronize(R.gpu->handle()));
            for (int r = 0; r < nRanks; ++r) {
                chimp_lattice *me = ranks[r].gpu->handle();
                for (int k = 0; k < chimp_num_neighbors(me); ++k) {
                    int nr;
                    long long ns, nrecv;
                    chimpCheck(chimp_neighbor_info(me, k, &nr, &ns, &nrecv));
                    chimp_lattice *peer = ranks[nr - firstRank].gpu->handle();
                    for (int j = 0; j < chimp_num_neighbors(peer); ++j) {
                        int pr;
                        chimpCheck(chimp_neighbor_info(peer, j, &pr, nullptr, nullptr));
                        if (pr == firstRank + r && nrecv)
                            cudaMemcpy(chimp_recv_buffer_dev(me, k), chimp_send_buffer_dev(peer, j), std::size_t(nrecv) * 8, cudaMemcpyDeviceToDevice);
                    }
                }
            }
            // the copies above ran on the legacy stream; the unpack kernels of step_end run on the engine's non-blocking streams
            cudaDeviceSynchronize();
            for (auto &R : ranks) chimpCheck(chimp_step_end(R.gpu->handle()));
        }
    }

    // results in reference layout and labels
    FILE *fp = std::fopen(outFile.c_str(), "wb");
    if (!fp) chimp_host::die("cannot open " + outFile);
    for (auto &R : ranks) {
        R.gpu->download(*R.f);
        R.gpu->download(*R.rho, *R.vel);
        const int sz = R.grid->size();
        std::fwrite(&sz, sizeof(int), 1, fp);
        std::fwrite(R.f->data(), sizeof(double), std::size_t(sz) * LT::nQ, fp);
        std::fwrite(R.rho->data(), sizeof(double), sz, fp);
        std::fwrite(R.vel->data(), sizeof(double), std::size_t(sz) * LT::nD, fp);
    }
    std::fclose(fp);
    if (!vtkDir.empty()) // OUTPUT VTK (std_case/main.cpp:101-104, 149-151)
        for (int r = 0; r < nRanks; ++r) {
            Rank<LT> &R = ranks[r];
            Output<LT> output(*R.grid, R.bulkNodes, vtkDir, firstRank + r, nRanks);
            output.add_file("lb_run");
            output.add_scalar_variables({"rho"}, {*R.rho});
            output.add_vector_variables({"vel"}, {*R.vel});
            output.write(nIterations);
        }
    std::cout << "std_case: " << nIterations << " iterations on " << nRanks << " rank(s) done" << std::endl;
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 6) {
        std::cout << "usage: std_case <D2Q9|D3Q19|D3Q27> <input.dat> <vtklb prefix> <rank> <out.bin> [nRanksInProcess [vtk output dir|- [pressure|inletoutlet]]]" << std::endl;
        return 2;
    }
    const std::string lattice = argv[1];
    const int rank = std::atoi(argv[4]);
    const int nRanks = argc > 6 ? std::atoi(argv[6]) : 1;
    const std::string vtkDir = argc > 7 && std::string(argv[7]) != "-" ? argv[7] : "";
    const std::string libraryBnd = argc > 8 ? argv[8] : "";
    if (lattice == "D2Q9") return run<D2Q9>(argv[2], argv[3], rank, argv[5], nRanks, vtkDir, libraryBnd);
    if (lattice == "D3Q19") return run<D3Q19>(argv[2], argv[3], rank, argv[5], nRanks, vtkDir, libraryBnd);
    if (lattice == "D3Q27") return run<D3Q27>(argv[2], argv[3], rank, argv[5], nRanks, vtkDir, libraryBnd);
    chimp_host::die("unknown lattice " + lattice);
}
