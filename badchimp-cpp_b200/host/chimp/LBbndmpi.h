// BndMpi<DXQY> (src/lbsolver/LBbndmpi.h:72-104, 182-331): the processor-boundary exchange lists.
// The reference negotiates them with blocking MPI_Send/Recv between neighbour ranks (tags 0-4,
// LBbndmpi.h:225-308).  Here MPI is only a launcher, so the same lists are derived without
// messages: what a neighbour rank would send in the handshake is a pure function of ITS geometry
// file, which this rank reads as well (file name pattern <prefix><rank>.vtklb).  The result -- send /
// receive node lists, directions per node, ghost node types -- is what MonLatMpi holds
// (LBmonlatmpi.h:72-98), and it is handed to the engine with GpuLattice::add().
#ifndef CHIMP_LBBNDMPI_H
#define CHIMP_LBBNDMPI_H

#include <memory>

#include "LBgrid.h"

struct MonLatLists {
    int neigRank;
    std::vector<int> nodesToSend, nDirPerNodeToSend, dirListToSend;
    std::vector<int> nodesReceived, nDirPerNodeReceived, dirListReceived;
};

// makeDirList (LBbndmpi.h:182-203): for every ghost node, the directions that point to a node of myRank
template <typename DXQY>
void makeDirList(int myRank, const Nodes<DXQY> &nodes, const Grid<DXQY> &grid, const std::vector<int> &ghostNodes,
                 std::vector<int> &nDirPerNode, std::vector<int> &dirList)
{
    nDirPerNode.assign(ghostNodes.size(), 0);
    for (std::size_t n = 0; n < ghostNodes.size(); ++n)
        for (int q = 0; q < DXQY::nQNonZero_; ++q)
            if (nodes.getRank(grid.neighbor(q, ghostNodes[n])) == myRank) {
                dirList.push_back(q);
                ++nDirPerNode[n];
            }
}

template <typename DXQY>
class BndMpi
{
public:
    // filePrefix: path such that filePrefix + std::to_string(rank) + ".vtklb" is rank's geometry file
    BndMpi(LBvtk<DXQY> &vtk, Nodes<DXQY> &nodes, const Grid<DXQY> &grid, const std::string &filePrefix) : myRank_(vtk.getRank())
    {
        for (int k = 0; k < vtk.getNumNeigProc(); ++k) {
            MonLatLists m;
            m.neigRank = vtk.getNeigRank(k);
            m.nodesReceived = vtk.getNeigNodesNo(k);
            makeDirList(myRank_, nodes, grid, m.nodesReceived, m.nDirPerNodeReceived, m.dirListReceived);
            // the neighbour's side of the handshake, from its own file
            LBvtk<DXQY> pv(filePrefix + std::to_string(m.neigRank) + ".vtklb");
            Grid<DXQY> pg(pv);
            Nodes<DXQY> pn(pv, pg);
            int mine = -1;
            for (int j = 0; j < pv.getNumNeigProc(); ++j)
                if (pv.getNeigRank(j) == myRank_) mine = j;
            if (mine < 0) chimp_host::die("rank " + std::to_string(m.neigRank) + " does not list rank " + std::to_string(myRank_) + " as neighbour");
            m.nodesToSend = pv.getNeigNodesNeigNo(mine); // my labels of the nodes it holds as ghosts
            makeDirList(m.neigRank, pn, pg, pv.getNeigNodesNo(mine), m.nDirPerNodeToSend, m.dirListToSend);
            // setupNodeType (LBbndmpi.h:317-331): ghost nodes take the type their owner gave them
            const std::vector<int> theirLabels = vtk.getNeigNodesNeigNo(k);
            for (std::size_t n = 0; n < m.nodesReceived.size(); ++n) nodes.addNodeType(pn.getType(theirLabels[n]), m.nodesReceived[n]);
            mpiList_.push_back(std::move(m));
        }
    }
    BndMpi(LBvtk<DXQY> &vtk, Nodes<DXQY> &, const Grid<DXQY> &) : myRank_(vtk.getRank())
    {
        if (vtk.getNumNeigProc() > 0) chimp_host::die("BndMpi: this rank has neighbours, pass the geometry file prefix");
    }
    const std::vector<MonLatLists> &lists() const { return mpiList_; }
    void printInfo() const { std::cout << "Number of neighbors = " << mpiList_.size() << std::endl; }

private:
    int myRank_;
    std::vector<MonLatLists> mpiList_;
};

#endif
