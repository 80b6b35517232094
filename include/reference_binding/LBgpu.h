// LBgpu.h -- binding stub a BADChIMP maintainer adds next to src/lbsolver/LBgrid.h: a thin C++ wrapper over the
// C-ABI (chimp_b200.h) for the reference's OWN types.  This file is compiled against the unmodified reference
// headers by oracle/Makefile (target `integration`) and is the code block of INTEGRATION.md section 1.
#ifndef LBGPU_H
#define LBGPU_H

#include <cstdint>
#include <iostream>
#include <valarray>
#include <vector>

#include "chimp_b200.h"
#include "LBglobal.h"
#include "LBlatticetypes.h"
#include "LBgrid.h"
#include "LBnodes.h"
#include "LBfield.h"
#include "LBhalfwaybb.h"
#include "LBpressurebnd.h"

template <typename DXQY> struct ChimpLatticeId;
template <> struct ChimpLatticeId<D2Q9>  { static constexpr int value = CHIMP_D2Q9;  };
template <> struct ChimpLatticeId<D3Q19> { static constexpr int value = CHIMP_D3Q19; };

inline void chimpCheck(int rc) {          // reference convention: print + exit(1) (LBvtk.h:230-233)
    if (rc) { std::cout << "ERROR in GPU engine: " << chimp_last_error() << std::endl; exit(1); }
}

template <typename DXQY>
class GpuLattice {
public:
    GpuLattice(const Grid<DXQY> &grid, const std::vector<int> &bulkNodes, int nFields) {
        std::vector<int32_t> neigh(std::size_t(grid.size()) * DXQY::nQ);
        for (int n = 0; n < grid.size(); ++n)
            for (int q = 0; q < DXQY::nQ; ++q) neigh[std::size_t(n) * DXQY::nQ + q] = grid.neighbor(q, n);   // LBgrid.h:187
        chimpCheck(chimp_create(&h_, ChimpLatticeId<DXQY>::value, grid.size(), neigh.data(),
                                int(bulkNodes.size()), bulkNodes.data(), nFields, -1));
    }
    ~GpuLattice() { chimp_destroy(h_); }
    GpuLattice(const GpuLattice &) = delete;
    GpuLattice &operator=(const GpuLattice &) = delete;

    // HalfWayBounceBack<DXQY> bb(findFluidBndNodes(nodes), nodes, grid);  ->  gpu.add(bb);
    void add(const HalfWayBounceBack<DXQY> &bb) {
        std::vector<int32_t> node, nb, ng, nd, links;
        for (int n = 0; n < bb.size(); ++n) {
            node.push_back(bb.nodeNo(n)); nb.push_back(bb.nBeta(n)); ng.push_back(bb.nGamma(n)); nd.push_back(bb.nDelta(n));
            for (auto q : bb.beta(n)) links.push_back(q);
            for (auto q : bb.gamma(n)) links.push_back(q);
            for (auto q : bb.delta(n)) links.push_back(q);
        }
        chimpCheck(chimp_add_halfway_bb(h_, bb.size(), node.data(), nb.data(), ng.data(), nd.data(), links.data()));
    }
    // PressureBnd<DXQY> bnd(bndNodes, nodes, grid);  ->  gpu.add(bnd, fieldNo, grid, rho);   replaces bnd.apply(fieldNo, f, grid, rho)
    // (LBpressurebnd.h:19-41).  Same stores in the same order; the densities are taken now (a prescribed boundary density).
    void add(const PressureBnd<DXQY> &bnd, int fieldNo, const Grid<DXQY> &grid, const ScalarField &rho) {
        std::vector<int32_t> nodeQ;
        std::vector<lbBase_t> values;
        auto store = [&](int q, int nodeNo) {
            nodeQ.push_back(grid.neighbor(q, nodeNo)); nodeQ.push_back(q);
            values.push_back(DXQY::w[q] * rho(fieldNo, nodeNo));
        };
        for (int n = 0; n < bnd.size(); ++n) {
            for (auto beta : bnd.beta(n)) store(beta, bnd.nodeNo(n));
            for (auto delta : bnd.delta(n)) { store(delta, bnd.nodeNo(n)); store(bnd.dirRev(delta), bnd.nodeNo(n)); }
        }
        chimpCheck(chimp_add_constant_links(h_, int(values.size()), nodeQ.data(), values.data()));
    }
    // InletOutlet<DXQY> bnd(...);  ->  gpu.add(bnd, grid, rho, vel);   replaces bnd.apply(fieldNo, f, grid, rho, vel)  (:51-88)
    void add(const InletOutlet<DXQY> &bnd, const Grid<DXQY> &grid, const lbBase_t &rho, const std::vector<lbBase_t> vel) {
        std::vector<int32_t> nodeQ;
        std::vector<lbBase_t> values;
        lbBase_t u_sq = DXQY::dot(vel, vel);
        std::valarray<lbBase_t> cu = DXQY::cDotAll(vel);
        auto store = [&](int q, int nodeNo) {
            nodeQ.push_back(grid.neighbor(q, nodeNo)); nodeQ.push_back(q);
            values.push_back(rho * DXQY::w[q] * (1.0 + DXQY::c2Inv * cu[q] + DXQY::c4Inv0_5 * (cu[q] * cu[q] - DXQY::c2 * u_sq)));
        };
        for (int n = 0; n < bnd.size(); ++n) {
            for (auto beta : bnd.beta(n)) store(beta, bnd.nodeNo(n));
            for (auto delta : bnd.delta(n)) { store(delta, bnd.nodeNo(n)); store(bnd.dirRev(delta), bnd.nodeNo(n)); }
        }
        chimpCheck(chimp_add_constant_links(h_, int(values.size()), nodeQ.data(), values.data()));
    }
    // std_one_phase: auto solidFluidLinks = findSolidFluidLinks(nodes, grid);  ->  gpu.addLinks(CHIMP_LINK_SOLID, links)
    void addLinks(int kind, const std::vector<std::vector<int>> &l) {
        std::vector<int32_t> flat;
        for (auto &x : l) flat.insert(flat.end(), x.begin(), x.end());
        chimpCheck(chimp_add_links(h_, kind, int(l.size()), flat.data()));
    }
    // one call per MonLatMpi in BndMpi::mpiList_ (needs a friend accessor or the lists BndMpi::setup built)
    void addNeighbor(int rank, const std::vector<int> &sendNodes, const std::vector<int> &nDirSend, const std::vector<int> &dirSend,
                     const std::vector<int> &recvNodes, const std::vector<int> &nDirRecv, const std::vector<int> &dirRecv) {
        chimpCheck(chimp_add_neighbor(h_, rank, int(sendNodes.size()), sendNodes.data(), nDirSend.data(), dirSend.data(),
                                      int(recvNodes.size()), recvNodes.data(), nDirRecv.data(), dirRecv.data()));
    }
    void finalize() { chimpCheck(chimp_finalize(h_, CHIMP_INDEX_COMPACT, 1)); }

    // &f(0, 0, 0) is the address of LbField::data_[0] (LBfield.h:300): no change to the reference classes is needed
    void upload(LbField<DXQY> &f)   { chimpCheck(chimp_upload_lbfield(h_, &f(0, 0, 0))); }
    void download(LbField<DXQY> &f) { chimpCheck(chimp_download_lbfield(h_, &f(0, 0, 0))); }
    void moments(ScalarField &rho, VectorField<DXQY> &vel) {
        chimpCheck(chimp_download_rho(h_, &rho(0, 0), rho.num_fields()));
        chimpCheck(chimp_download_vel(h_, &vel(0, 0, 0)));
    }
    void stepBGK(lbBase_t tau, const std::valarray<lbBase_t> &F, int nSteps) {
        chimp_single_params p{CHIMP_BGK, tau, 0, 0, {F[0], F[1], DXQY::nD == 3 ? F[2] : 0.0}};
        chimpCheck(chimp_step_single(h_, &p, nSteps));
    }
    void stepTRT(lbBase_t tauSym, lbBase_t tauAnti, const std::valarray<lbBase_t> &F, int nSteps) {
        chimp_single_params p{CHIMP_TRT, 0, tauSym, tauAnti, {F[0], F[1], DXQY::nD == 3 ? F[2] : 0.0}};
        chimpCheck(chimp_step_single(h_, &p, nSteps));
    }
    // std_one_phase: per-node attributes of main.cpp:279-345 (forceOn, interiorDomainsLabel, addMassSource, the global
    // 1/count per interior domain) and the density of the pressure boundary (:593); afterwards stepBGK / stepTRT run
    // the loop of main.cpp:513-597, mass-conservation source included
    void setOnePhaseAttributes(ScalarField &forceOn, const std::vector<int> &interiorDomainsLabel, const std::vector<lbBase_t> &addMassSource,
                               const std::vector<lbBase_t> &massSourceScaleFactor, lbBase_t rhoW) {
        chimpCheck(chimp_set_one_phase_attributes(h_, &forceOn(0, 0), interiorDomainsLabel.data(), addMassSource.data(),
                                                  int(massSourceScaleFactor.size()), massSourceScaleFactor.data(), rhoW));
    }
    // massFluxLocal of main.cpp:606-618 from the moments of the last iteration, summed in list order
    std::vector<lbBase_t> massFlux(const std::vector<int> &pressureFluidNodes, const std::vector<int> &fluidPhase) {
        std::vector<lbBase_t> q(2, 0.0);
        chimpCheck(chimp_node_list_flux(h_, int(pressureFluidNodes.size()), pressureFluidNodes.data(), fluidPhase.data(), 2, 0,
                                        DXQY::nD - 1, q.data()));
        return q;
    }
    // twophase (main_TWOPHASE.cpp): solidBnd = findSolidBndNodes(nodes) before finalize(); rho(2, size) with the
    // wettability densities of the solid boundary nodes (:150-181) after it; then nSteps iterations of :236-392
    void setSolidBoundary(const std::vector<int> &solidBnd) { chimpCheck(chimp_set_solid_boundary(h_, int(solidBnd.size()), solidBnd.data())); }
    void setTwoPhaseDensity(ScalarField &rho) { chimpCheck(chimp_set_twophase_density(h_, &rho(0, 0))); }
    void stepTwoPhase(lbBase_t tau0, lbBase_t tau1, lbBase_t sigma, lbBase_t beta, lbBase_t momx, const std::valarray<lbBase_t> &F,
                      int numNodesGlobal, int nSteps) {
        chimp_twophase_params p{tau0, tau1, sigma, beta, momx, {F[0], F[1], DXQY::nD == 3 ? F[2] : 0.0}, numNodesGlobal};
        chimpCheck(chimp_step_twophase(h_, &p, nSteps));
    }
    void phaseField(ScalarField &cgField) { chimpCheck(chimp_download_phase_field(h_, &cgField(0, 0))); }
    lbBase_t lastFluxForce() { return chimp_last_flux_force(h_); }
    chimp_lattice *handle() { return h_; }
private:
    chimp_lattice *h_ = nullptr;
};

#endif
