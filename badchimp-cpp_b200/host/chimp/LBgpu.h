// GpuLattice<DXQY>: the reference's objects handed to the C-ABI engine (include/chimp_b200.h).
// It replaces, in a main, the per-node loop + swapData + communicateLbField + boundary apply().
#ifndef CHIMP_LBGPU_H
#define CHIMP_LBGPU_H

#include <cstdint>

#include "../../../include/chimp_b200.h"
#include "LBbndmpi.h"
#include "LBfield.h"
#include "LBhalfwaybb.h"

inline void chimpCheck(int rc)
{
    if (rc) chimp_host::die(std::string("GPU engine: ") + chimp_last_error());
}

template <typename DXQY>
class GpuLattice
{
public:
    GpuLattice(const Grid<DXQY> &grid, const std::vector<int> &bulkNodes, int nFields, int device = -1) : nNodes_(grid.size())
    {
        chimpCheck(chimp_create(&h_, DXQY::chimpId, grid.size(), grid.neighborList().data(), int(bulkNodes.size()),
                                bulkNodes.data(), nFields, device));
    }
    // straight from a raw voxel array [nx][ny][nz] (0 = solid), reference numbering, half-way bounce back at every solid
    // link: the route to sizes the ASCII .vtklb loader cannot read (LBvtk.h:194-201).  Fields of such a lattice have
    // numFluidNodes() + 1 rows (row 0 is the reference's dummy node).
    GpuLattice(const std::vector<std::uint8_t> &voxels, int nx, int ny, int nz, int periodicMask, int nFields, int device = -1)
    {
        if (voxels.size() != std::size_t(nx) * ny * nz) chimp_host::die("voxel array does not hold nx*ny*nz cells");
        chimpCheck(chimp_create_from_voxels(&h_, DXQY::chimpId, nx, ny, nz, voxels.data(), periodicMask, nFields, CHIMP_INDEX_COMPACT, device));
        nNodes_ = chimp_num_own_nodes(h_) + 1;
    }
    int numFluidNodes() const { return chimp_num_own_nodes(h_); }
    void initUniform(lbBase_t rho) { chimpCheck(chimp_init_uniform(h_, rho)); }
    // two-field lattice from voxels: colour of the solid cells next to fluid (main_TWOPHASE.cpp:280-284)
    void setWallColour(const std::vector<std::uint8_t> &voxels, int nx, int ny, int nz, int periodicMask, const std::vector<double> &wallPhi)
    {
        chimpCheck(chimp_set_phi_table_from_voxels(h_, nx, ny, nz, voxels.data(), periodicMask, wallPhi.data()));
    }
    double stepBGKTimed(lbBase_t tau, const std::valarray<lbBase_t> &bodyForce, int nSteps)
    {
        chimp_single_params p{};
        p.collision = CHIMP_BGK;
        p.tau = tau;
        for (int d = 0; d < DXQY::nD; ++d) p.force[d] = bodyForce[d];
        double ms = 0.0;
        chimpCheck(chimp_step_timed(h_, &p, nSteps, &ms));
        return ms;
    }
    ~GpuLattice() { chimp_destroy(h_); }
    GpuLattice(const GpuLattice &) = delete;
    GpuLattice &operator=(const GpuLattice &) = delete;

    void add(const HalfWayBounceBack<DXQY> &bb)
    {
        chimpCheck(chimp_add_halfway_bb(h_, bb.size(), bb.nodeList().data(), bb.nBetaList().data(), bb.nGammaList().data(),
                                        bb.nDeltaList().data(), bb.linkList().data()));
    }
    // PressureBnd / InletOutlet (LBpressurebnd.h): add them after the bounce back, in the order the main applies them
    void add(const PressureBnd<DXQY> &bnd, int fieldNo, const Grid<DXQY> &grid, const ScalarField &rho)
    {
        std::vector<int> nodeQ;
        std::vector<lbBase_t> values;
        bnd.links(fieldNo, grid, rho, nodeQ, values);
        chimpCheck(chimp_add_constant_links(h_, int(values.size()), nodeQ.data(), values.data()));
    }
    void add(const InletOutlet<DXQY> &bnd, const Grid<DXQY> &grid, const lbBase_t &rho, const std::vector<lbBase_t> &vel)
    {
        std::vector<int> nodeQ;
        std::vector<lbBase_t> values;
        bnd.links(grid, rho, vel, nodeQ, values);
        chimpCheck(chimp_add_constant_links(h_, int(values.size()), nodeQ.data(), values.data()));
    }
    void add(const BndMpi<DXQY> &mpi)
    {
        for (const MonLatLists &m : mpi.lists())
            chimpCheck(chimp_add_neighbor(h_, m.neigRank, int(m.nodesToSend.size()), m.nodesToSend.data(), m.nDirPerNodeToSend.data(),
                                          m.dirListToSend.data(), int(m.nodesReceived.size()), m.nodesReceived.data(),
                                          m.nDirPerNodeReceived.data(), m.dirListReceived.data()));
    }
    // link lists {nodeFluid, qUnknown, nodeWall, qKnown} of std_one_phase (main.cpp:27-126)
    void addLinks(int kind, const std::vector<std::vector<int>> &links)
    {
        std::vector<int32_t> flat;
        for (const auto &l : links) flat.insert(flat.end(), l.begin(), l.end());
        chimpCheck(chimp_add_links(h_, kind, int(links.size()), flat.data()));
    }
    void setSolidBoundary(const std::vector<int> &nodes) { chimpCheck(chimp_set_solid_boundary(h_, int(nodes.size()), nodes.data())); }
    void finalize(int indexForm = CHIMP_INDEX_COMPACT, bool boundaryFirst = true) { chimpCheck(chimp_finalize(h_, indexForm, boundaryFirst)); }

    void upload(LbField<DXQY> &f) { chimpCheck(chimp_upload_lbfield(h_, f.data())); }
    void download(LbField<DXQY> &f) { chimpCheck(chimp_download_lbfield(h_, f.data())); }
    void download(ScalarField &rho, VectorField<DXQY> &vel)
    {
        chimpCheck(chimp_download_rho(h_, rho.data(), rho.num_fields()));
        chimpCheck(chimp_download_vel(h_, vel.data()));
    }
    // nSteps iterations of std_case/main.cpp:110-145 with calcOmegaBGK + calcDeltaOmegaF
    void stepBGK(lbBase_t tau, const std::valarray<lbBase_t> &bodyForce, int nSteps)
    {
        chimp_single_params p{};
        p.collision = CHIMP_BGK;
        p.tau = tau;
        for (int d = 0; d < DXQY::nD; ++d) p.force[d] = bodyForce[d];
        chimpCheck(chimp_step_single(h_, &p, nSteps));
    }
    // same with calcOmegaBGKTRT + calcDeltaOmegaFTRT(phi = 1)
    void stepTRT(lbBase_t tauSym, lbBase_t tauAnti, const std::valarray<lbBase_t> &bodyForce, int nSteps)
    {
        chimp_single_params p{};
        p.collision = CHIMP_TRT;
        p.tau_sym = tauSym;
        p.tau_anti = tauAnti;
        for (int d = 0; d < DXQY::nD; ++d) p.force[d] = bodyForce[d];
        chimpCheck(chimp_step_single(h_, &p, nSteps));
    }
    // std_one_phase: per-node attributes of main.cpp:253-345 (force switch, interior-domain label, mass-source
    // marker, 1/count per interior domain) and the pressure-boundary density of :593; afterwards stepBGK /
    // stepTRT run the loop of main.cpp:513-597
    void setOnePhaseAttributes(ScalarField &forceOn, const std::vector<int> &interiorDomainsLabel, const std::vector<lbBase_t> &addMassSource,
                               const std::vector<lbBase_t> &massSourceScaleFactor, lbBase_t rhoW)
    {
        chimpCheck(chimp_set_one_phase_attributes(h_, forceOn.data(), interiorDomainsLabel.data(), addMassSource.data(),
                                                  int(massSourceScaleFactor.size()), massSourceScaleFactor.data(), rhoW));
    }
    std::vector<lbBase_t> massChange(int nLabels)
    {
        std::vector<lbBase_t> m(nLabels, 0.0);
        chimpCheck(chimp_download_mass_change(h_, m.data()));
        return m;
    }
    // massFluxLocal of main.cpp:606-618: sum of vel(0,2,n)*rho(0,n) over the pressure-boundary nodes, per fluid phase
    std::vector<lbBase_t> massFlux(const std::vector<int> &pressureFluidNodes, const std::vector<int> &fluidPhase)
    {
        std::vector<lbBase_t> q(2, 0.0);
        chimpCheck(chimp_node_list_flux(h_, int(pressureFluidNodes.size()), pressureFluidNodes.data(), fluidPhase.data(), 2, 0,
                                        DXQY::nD - 1, q.data()));
        return q;
    }
    // calcFluxForceCartDir (LBglobalforcing.h:8-33)
    lbBase_t fluxForceCartDir(int fieldNo, int cartDir, lbBase_t fixedFlux, int numNodesGlobal)
    {
        lbBase_t F = 0.0;
        chimpCheck(chimp_flux_force(h_, fieldNo, cartDir, fixedFlux, numNodesGlobal, &F));
        return F;
    }
    // calcCapNumbForceCartDir (LBglobalforcing.h:35-98)
    lbBase_t capNumbForceCartDir(int cartDir, lbBase_t sigmaCapNumb, lbBase_t nu0, lbBase_t nu1, int numNodesGlobal)
    {
        lbBase_t F = 0.0;
        chimpCheck(chimp_capillary_force(h_, cartDir, sigmaCapNumb, nu0, nu1, numNodesGlobal, &F));
        return F;
    }
    // twophase: wall colour from rho(2,size) (main_TWOPHASE.cpp:173-181, 280-284), then nSteps iterations of :236-392
    void setTwoPhaseDensity(ScalarField &rho) { chimpCheck(chimp_set_twophase_density(h_, rho.data())); }
    void stepTwoPhase(lbBase_t tau0, lbBase_t tau1, lbBase_t sigma, lbBase_t beta, lbBase_t momx, const std::valarray<lbBase_t> &bodyForce,
                      long long numNodesGlobal, int nSteps)
    {
        chimp_twophase_params p{};
        p.tau0 = tau0;
        p.tau1 = tau1;
        p.sigma = sigma;
        p.beta = beta;
        p.momx = momx;
        for (int d = 0; d < DXQY::nD; ++d) p.force[d] = bodyForce[d];
        p.n_fluid_global = numNodesGlobal;
        chimpCheck(chimp_step_twophase(h_, &p, nSteps));
    }
    void downloadPhaseField(ScalarField &cgField) { chimpCheck(chimp_download_phase_field(h_, cgField.data())); }
    lbBase_t lastFluxForce() { return chimp_last_flux_force(h_); }
    chimp_lattice *handle() { return h_; }

private:
    chimp_lattice *h_ = nullptr;
    int nNodes_;
};

#endif
