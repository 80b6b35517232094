// Grid<DXQY> (src/lbsolver/LBgrid.h:91-198): node positions and the per-node neighbour list
// neigList_[node*nQ + q]; node 0 is the shared dummy node.  Nodes<DXQY> (LBnodes.h:55-209):
// rank, type (-1 default, 0 bulk solid, 1 solid boundary, 2 fluid boundary, 3 bulk fluid,
// 4 transient "fluid of another rank") and tag per node.
#ifndef CHIMP_LBGRID_H
#define CHIMP_LBGRID_H

#include "LBvtk.h"

template <typename DXQY>
class Grid
{
public:
    explicit Grid(LBvtk<DXQY> &vtk) : nNodes_(vtk.endNodeNo()), neigList_(std::size_t(nNodes_) * DXQY::nQ, 0), pos_(std::size_t(nNodes_) * DXQY::nD, -1)
    {
        vtk.toPos();
        for (int n = vtk.beginNodeNo(); n < vtk.endNodeNo(); ++n) {
            const std::vector<int> p = vtk.template getPos<int>();
            for (int d = 0; d < DXQY::nD; ++d) pos_[std::size_t(n) * DXQY::nD + d] = p[d];
        }
        vtk.toNeighbors();
        for (int n = vtk.beginNodeNo(); n < vtk.endNodeNo(); ++n) {
            const std::vector<int> nb = vtk.template getNeighbors<int>();
            for (int q = 0; q < DXQY::nQ; ++q) neigList_[std::size_t(n) * DXQY::nQ + q] = nb[q];
        }
    }
    int neighbor(int q, int nodeNo) const { return neigList_[std::size_t(nodeNo) * DXQY::nQ + q]; }
    std::vector<int> neighbor(int nodeNo) const
    {
        return std::vector<int>(neigList_.begin() + std::size_t(nodeNo) * DXQY::nQ, neigList_.begin() + std::size_t(nodeNo + 1) * DXQY::nQ);
    }
    int pos(int nodeNo, int d) const { return pos_[std::size_t(nodeNo) * DXQY::nD + d]; }
    std::vector<int> pos(int nodeNo) const
    {
        return std::vector<int>(pos_.begin() + std::size_t(nodeNo) * DXQY::nD, pos_.begin() + std::size_t(nodeNo + 1) * DXQY::nD);
    }
    int size() const { return nNodes_; }
    const std::vector<int> &neighborList() const { return neigList_; }
    // all positions (nD ints per node, dummy node 0 first) and the positions of a node list (LBgrid.h:112-114)
    const std::vector<int> &pos() const { return pos_; }
    std::vector<int> pos(const std::vector<int> &nodes) const
    {
        std::vector<int> v(nodes.size() * DXQY::nD);
        for (std::size_t n = 0; n < nodes.size(); ++n)
            for (int d = 0; d < DXQY::nD; ++d) v[n * DXQY::nD + d] = pos_[std::size_t(nodes[n]) * DXQY::nD + d];
        return v;
    }

private:
    int nNodes_;
    std::vector<int> neigList_, pos_;
};

template <typename DXQY>
class Nodes
{
public:
    Nodes(LBvtk<DXQY> &vtk, const Grid<DXQY> &grid)
        : nNodes_(grid.size()), myRank_(vtk.getRank()), nodeRank_(nNodes_, vtk.getRank()), nodeType_(nNodes_, -1), nodeTag_(nNodes_, -1)
    {
        nodeRank_[0] = -1;
        for (int k = 0; k < vtk.getNumNeigProc(); ++k)
            for (int node : vtk.getNeigNodesNo(k)) nodeRank_[node] = vtk.getNeigRank(k);
        vtk.toAttribute("nodetype");
        for (int n = vtk.beginNodeNo(); n < vtk.endNodeNo(); ++n) nodeType_[n] = vtk.template getScalarAttribute<int>();
        setupNodeType(grid);
    }
    // LBnodes.h:141-209
    void setupNodeType(const Grid<DXQY> &grid)
    {
        nodeType_[0] = -1;
        for (int n = 1; n < nNodes_; ++n) nodeType_[n] = nodeType_[n] == 0 ? 0 : 3;
        std::vector<short> out(nodeType_);
        for (int n = 1; n < nNodes_; ++n) {
            bool anyFluid = false, anySolid = false;
            for (int q = 0; q < DXQY::nQ; ++q) {
                const bool fl = nodeType_[grid.neighbor(q, n)] > 1;
                anyFluid |= fl;
                anySolid |= !fl;
            }
            if (nodeType_[n] < 2) out[n] = anyFluid ? 1 : 0;
            else out[n] = nodeRank_[n] != myRank_ ? 4 : (anySolid ? 2 : 3);
        }
        nodeType_.swap(out);
    }
    int size() const { return nNodes_; }
    int getType(int n) const { return nodeType_[n]; }
    int getRank(int n) const { return nodeRank_[n]; }
    int getTag(int n) const { return nodeTag_[n]; }
    bool isDefault(int n) const { return nodeType_[n] == -1; }
    bool isMyRank(int n) const { return nodeRank_[n] == myRank_; }
    bool isSolid(int n) const { return nodeType_[n] < 2; }
    bool isBulkSolid(int n) const { return nodeType_[n] == 0; }
    bool isSolidBoundary(int n) const { return nodeType_[n] == 1; }
    bool isFluid(int n) const { return nodeType_[n] > 1; }
    bool isBulkFluid(int n) const { return nodeType_[n] == 3; }
    bool isFluidBoundary(int n) const { return nodeType_[n] == 2; }
    bool isMpiBoundary(int n) const { return nodeRank_[n] != myRank_ && !isDefault(n); }
    void addNodeType(int type, int n) { nodeType_[n] = short(type); }
    void setTag(int tag, int n) { nodeTag_[n] = short(tag); }
    int myRank() const { return myRank_; }

private:
    int nNodes_, myRank_;
    std::vector<int> nodeRank_;
    std::vector<short> nodeType_, nodeTag_;
};

// node lists (src/lbsolver/LBgeometry.h:11-21, 37-56)
template <typename DXQY>
std::vector<int> findBulkNodes(const Nodes<DXQY> &nodes)
{
    std::vector<int> v;
    for (int n = 1; n < nodes.size(); ++n)
        if (nodes.isFluid(n) && nodes.isMyRank(n)) v.push_back(n);
    return v;
}
template <typename DXQY>
std::vector<int> findSolidBndNodes(const Nodes<DXQY> &nodes)
{
    std::vector<int> v;
    for (int n = 1; n < nodes.size(); ++n)
        if (nodes.isSolidBoundary(n)) v.push_back(n);
    return v;
}
template <typename DXQY>
std::vector<int> findFluidBndNodes(const Nodes<DXQY> &nodes)
{
    std::vector<int> v;
    for (int n = 1; n < nodes.size(); ++n)
        if (nodes.isFluidBoundary(n) && nodes.isMyRank(n)) v.push_back(n);
    return v;
}

#endif
