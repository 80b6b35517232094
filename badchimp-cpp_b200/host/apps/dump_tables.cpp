// Host-only check program (no GPU, no engine): reads <prefix><rank>.vtklb through the host mirror and
// prints the integer tables the parity tests compare with the reference's (tests/test_host_cpp.py).
//   dump_tables <lattice> <vtklb prefix> <rank>
#include "../chimp/LBglobal.h"
#include "../chimp/LBlattices.h"
#include "../chimp/LBvtk.h"
#include "../chimp/LBgrid.h"
#include "../chimp/LBhalfwaybb.h"
#include "../chimp/LBbndmpi.h"

static void put(const char *name, const std::vector<int> &v)
{
    std::cout << name << " " << v.size();
    for (int x : v) std::cout << " " << x;
    std::cout << "\n";
}

template <typename LT>
int run(const std::string &prefix, int rank)
{
    LBvtk<LT> vtklb(prefix + std::to_string(rank) + ".vtklb");
    Grid<LT> grid(vtklb);
    Nodes<LT> nodes(vtklb, grid);
    BndMpi<LT> mpi(vtklb, nodes, grid, prefix);
    std::vector<int> type(grid.size()), rk(grid.size());
    for (int n = 0; n < grid.size(); ++n) { type[n] = nodes.getType(n); rk[n] = nodes.getRank(n); }
    put("neigh", grid.neighborList());
    put("type", type);
    put("rank", rk);
    put("bulk", findBulkNodes(nodes));
    put("fluidBnd", findFluidBndNodes(nodes));
    put("solidBnd", findSolidBndNodes(nodes));
    HalfWayBounceBack<LT> bb(findFluidBndNodes(nodes), nodes, grid);
    put("bb.node", bb.nodeList());
    put("bb.nBeta", bb.nBetaList());
    put("bb.nGamma", bb.nGammaList());
    put("bb.nDelta", bb.nDeltaList());
    put("bb.links", bb.linkList());
    int k = 0;
    for (const MonLatLists &m : mpi.lists()) {
        const std::string p = "mpi" + std::to_string(k++) + ".";
        put((p + "neigRank").c_str(), {m.neigRank});
        put((p + "nodesToSend").c_str(), m.nodesToSend);
        put((p + "nDirPerNodeToSend").c_str(), m.nDirPerNodeToSend);
        put((p + "dirListToSend").c_str(), m.dirListToSend);
        put((p + "nodesReceived").c_str(), m.nodesReceived);
        put((p + "nDirPerNodeReceived").c_str(), m.nDirPerNodeReceived);
        put((p + "dirListReceived").c_str(), m.dirListReceived);
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    const std::string lattice = argv[1];
    if (lattice == "D2Q9") return run<D2Q9>(argv[2], std::atoi(argv[3]));
    if (lattice == "D3Q19") return run<D3Q19>(argv[2], std::atoi(argv[3]));
    if (lattice == "D3Q27") return run<D3Q27>(argv[2], std::atoi(argv[3]));
    return 2;
}
