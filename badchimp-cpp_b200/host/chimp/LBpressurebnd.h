// PressureBnd<DXQY> and InletOutlet<DXQY> (src/lbsolver/LBpressurebnd.h:10-88 on top of Boundary<DXQY>,
// LBboundary.h:77-160): after streaming, every beta link (and both directions of every delta link) of a boundary
// node stores a prescribed population at the neighbour it points to -- w[q] * rho(fieldNo, node) resp. the
// equilibrium of a given (rho, vel).  Boundary<DXQY> classes a pair by Nodes::isSolid of the two neighbours; isSolid
// is !isFluid (LBnodes.h:77-80), so the lists are those of the bounce-back helper.  apply() performs the stores on a
// host LbField like the reference; links() hands the same stores to the engine (GpuLattice::add), where they are
// folded into the pull table as constant sources: the values are taken when the boundary is added (a prescribed
// density / velocity -- the reference has no caller that varies them).
#ifndef CHIMP_LBPRESSUREBND_H
#define CHIMP_LBPRESSUREBND_H

#include "LBhalfwaybb.h"

template <typename DXQY>
class Boundary
{
public:
    Boundary(const std::vector<int> &bndNodes, const Nodes<DXQY> &nodes, const Grid<DXQY> &grid) : pairs_(bndNodes, nodes, grid) {}
    int size() const { return pairs_.size(); }
    int nodeNo(int b) const { return pairs_.nodeNo(b); }
    std::vector<int> beta(int b) const { return pairs_.beta(b); }
    std::vector<int> gamma(int b) const { return pairs_.gamma(b); }
    std::vector<int> delta(int b) const { return pairs_.delta(b); }
    int dirRev(int q) const { return DXQY::reverseDirection(q); }

protected:
    // calls store(q, node) for every store of the reference's apply() loops, in their order
    template <class Store>
    void forEachStore(Store store) const
    {
        for (int n = 0; n < size(); ++n) {
            const int node = nodeNo(n);
            for (int q : beta(n)) store(q, node);
            for (int q : delta(n)) {
                store(q, node);
                store(dirRev(q), node);
            }
        }
    }

private:
    HalfWayBounceBack<DXQY> pairs_;
};

template <typename DXQY>
class PressureBnd : public Boundary<DXQY>
{
public:
    using Boundary<DXQY>::Boundary;
    void apply(int fieldNo, LbField<DXQY> &f, const Grid<DXQY> &grid, const ScalarField &rho) const
    {
        this->forEachStore([&](int q, int node) { f(fieldNo, q, grid.neighbor(q, node)) = DXQY::w[q] * rho(fieldNo, node); });
    }
    // the stores as (destination node, direction) pairs and values
    void links(int fieldNo, const Grid<DXQY> &grid, const ScalarField &rho, std::vector<int> &nodeQ, std::vector<lbBase_t> &values) const
    {
        this->forEachStore([&](int q, int node) {
            nodeQ.push_back(grid.neighbor(q, node));
            nodeQ.push_back(q);
            values.push_back(DXQY::w[q] * rho(fieldNo, node));
        });
    }
};

template <typename DXQY>
class InletOutlet : public Boundary<DXQY>
{
public:
    using Boundary<DXQY>::Boundary;
    void apply(int fieldNo, LbField<DXQY> &f, const Grid<DXQY> &grid, const lbBase_t &rho, const std::vector<lbBase_t> vel) const
    {
        const lbBase_t u_sq = DXQY::dot(vel, vel);
        const std::valarray<lbBase_t> cu = DXQY::cDotAll(vel);
        this->forEachStore([&](int q, int node) {
            f(fieldNo, q, grid.neighbor(q, node)) = rho * DXQY::w[q] * (1.0 + DXQY::c2Inv * cu[q] + DXQY::c4Inv0_5 * (cu[q] * cu[q] - DXQY::c2 * u_sq));
        });
    }
    void links(const Grid<DXQY> &grid, const lbBase_t &rho, const std::vector<lbBase_t> vel, std::vector<int> &nodeQ, std::vector<lbBase_t> &values) const
    {
        const lbBase_t u_sq = DXQY::dot(vel, vel);
        const std::valarray<lbBase_t> cu = DXQY::cDotAll(vel);
        this->forEachStore([&](int q, int node) {
            nodeQ.push_back(grid.neighbor(q, node));
            nodeQ.push_back(q);
            values.push_back(rho * DXQY::w[q] * (1.0 + DXQY::c2Inv * cu[q] + DXQY::c4Inv0_5 * (cu[q] * cu[q] - DXQY::c2 * u_sq)));
        });
    }
};

#endif
