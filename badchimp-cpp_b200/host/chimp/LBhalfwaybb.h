// HalfWayBounceBack<DXQY> (src/lbsolver/LBhalfwaybb.h:25-63 on top of BoundaryHalwWayHelper,
// LBhalfwayhelperclass.h:110-253): per boundary node the direction pairs are classed beta (one side
// solid; the stored direction is the unknown one, pointing away from the wall), gamma (both fluid)
// or delta (both solid).  The copies the reference's apply() performs after streaming are folded
// into the engine's pull table: hand the object to GpuLattice::add().  apply() itself is kept for host
// loops (checks, small cases): it performs the same copies on a host LbField after swapData.
#ifndef CHIMP_LBHALFWAYBB_H
#define CHIMP_LBHALFWAYBB_H

#include "LBfield.h"
#include "LBgrid.h"

template <typename DXQY>
class HalfWayBounceBack
{
public:
    HalfWayBounceBack(const std::vector<int> &bndNodes, const Nodes<DXQY> &nodes, const Grid<DXQY> &grid)
        : nodes_(bndNodes), links_(bndNodes.size() * DXQY::nDirPairs_), nBeta_(bndNodes.size()), nGamma_(bndNodes.size()), nDelta_(bndNodes.size())
    {
        constexpr int P = DXQY::nDirPairs_;
        for (std::size_t b = 0; b < bndNodes.size(); ++b) {
            std::vector<int> beta, gamma, delta;
            for (int q = 0; q < P; ++q) {
                const bool fq = nodes.isFluid(grid.neighbor(q, bndNodes[b]));
                const bool fr = nodes.isFluid(grid.neighbor(q + P, bndNodes[b]));
                if (fq && fr) gamma.push_back(q);
                else if (fq) beta.push_back(q);
                else if (fr) beta.push_back(q + P);
                else delta.push_back(q);
            }
            nBeta_[b] = int(beta.size());
            nGamma_[b] = int(gamma.size());
            nDelta_[b] = int(delta.size());
            int *l = &links_[b * P];
            for (int q : beta) *l++ = q;
            for (int q : gamma) *l++ = q;
            for (int q : delta) *l++ = q;
        }
    }
    // half-way bounce back of one field after streaming: the unknown direction of a beta link takes the value that
    // left the node towards the wall; both directions of a delta link do
    void apply(int fieldNo, LbField<DXQY> &f, const Grid<DXQY> &grid) const
    {
        for (int b = 0; b < size(); ++b) {
            const int node = nodes_[b];
            for (int q : beta(b)) {
                const int r = dirRev(q);
                f(fieldNo, q, node) = f(fieldNo, r, grid.neighbor(r, node));
            }
            for (int q : delta(b)) {
                const int r = dirRev(q);
                f(fieldNo, q, node) = f(fieldNo, r, grid.neighbor(r, node));
                f(fieldNo, r, node) = f(fieldNo, q, grid.neighbor(q, node));
            }
        }
    }
    void apply(LbField<DXQY> &f, const Grid<DXQY> &grid) const
    {
        for (int n = 0; n < f.num_fields(); ++n) apply(n, f, grid);
    }
    int size() const { return int(nodes_.size()); }
    int nodeNo(int b) const { return nodes_[b]; }
    int nBeta(int b) const { return nBeta_[b]; }
    int nGamma(int b) const { return nGamma_[b]; }
    int nDelta(int b) const { return nDelta_[b]; }
    int dirRev(int q) const { return DXQY::reverseDirection(q); }
    std::vector<int> beta(int b) const { return slice(b, 0, nBeta_[b]); }
    std::vector<int> gamma(int b) const { return slice(b, nBeta_[b], nGamma_[b]); }
    std::vector<int> delta(int b) const { return slice(b, nBeta_[b] + nGamma_[b], nDelta_[b]); }
    const std::vector<int> &nodeList() const { return nodes_; }
    const std::vector<int> &linkList() const { return links_; }
    const std::vector<int> &nBetaList() const { return nBeta_; }
    const std::vector<int> &nGammaList() const { return nGamma_; }
    const std::vector<int> &nDeltaList() const { return nDelta_; }

private:
    std::vector<int> slice(int b, int off, int n) const
    {
        const int *p = &links_[std::size_t(b) * DXQY::nDirPairs_ + off];
        return std::vector<int>(p, p + n);
    }
    std::vector<int> nodes_, links_, nBeta_, nGamma_, nDelta_;
};

#endif
