// Writes the VTK files of one rank from fields held on the host -- Output<LT> (host/chimp/Output.h) used the
// way the reference mains use it (std_case/main.cpp:101-104, main_TWOPHASE.cpp:214-223).  No GPU involved:
// the CPU tests feed it the reference's own field dumps and compare the files byte for byte.
//
//   vtk_write <D2Q9|D3Q19|D3Q27> <vtklb prefix> <rank> <nRanks> <fields.bin> <nRhoFields> <out dir> <file name> <time> [geo|ascii]
//
// fields.bin: rho (ScalarField layout, nRhoFields) followed by vel (VectorField layout), raw doubles.
#include <cstdio>

#include "../chimp/LBfield.h"
#include "../chimp/LBgrid.h"
#include "../chimp/LBlattices.h"
#include "../chimp/LBvtk.h"
#include "../chimp/Output.h"

template <typename LT, int FMT>
int run(char **argv, bool withGeo)
{
    const int rank = std::atoi(argv[3]), nRanks = std::atoi(argv[4]), nRho = std::atoi(argv[6]);
    LBvtk<LT> vtklb(std::string(argv[2]) + std::to_string(rank) + ".vtklb");
    Grid<LT> grid(vtklb);
    Nodes<LT> nodes(vtklb, grid);
    std::vector<int> bulkNodes = findBulkNodes(nodes);
    ScalarField rho(nRho, grid.size());
    VectorField<LT> vel(1, grid.size());
    FILE *fp = std::fopen(argv[5], "rb");
    if (!fp) chimp_host::die(std::string("cannot open ") + argv[5]);
    const std::size_t nr = std::size_t(grid.size()) * nRho, nv = std::size_t(grid.size()) * LT::nD;
    if (std::fread(rho.data(), sizeof(double), nr, fp) != nr || std::fread(vel.data(), sizeof(double), nv, fp) != nv)
        chimp_host::die("fields file too short");
    std::fclose(fp);
    Output<LT, double, FMT> output(grid, bulkNodes, argv[7], rank, nRanks);
    output.add_file(argv[8]);
    output.add_scalar_variables({"rho"}, {rho});
    output.add_vector_variables({"vel"}, {vel});
    if (withGeo) {
        std::vector<int> geo(grid.size(), -1); // Nodes::geo (LBnodes.h:93-99)
        for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) geo[n] = nodes.isSolid(n) ? 1 : 0;
        Output<LT, int> geoout(grid.pos(), argv[7], rank, nRanks, "geo", geo);
        geoout.write();
    }
    output.write(std::atof(argv[9]));
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 10) {
        std::cout << "usage: vtk_write <lattice> <vtklb prefix> <rank> <nRanks> <fields.bin> <nRhoFields> <out dir> <file name> <time> [geo|ascii]" << std::endl;
        return 2;
    }
    const std::string lattice = argv[1];
    const bool withGeo = argc > 10 && std::string(argv[10]) == "geo";
    const bool ascii = argc > 10 && std::string(argv[10]) == "ascii";
    if (ascii) {
        if (lattice == "D2Q9") return run<D2Q9, VTK::ASCII>(argv, false);
        if (lattice == "D3Q19") return run<D3Q19, VTK::ASCII>(argv, false);
        if (lattice == "D3Q27") return run<D3Q27, VTK::ASCII>(argv, false);
    }
    if (lattice == "D2Q9") return run<D2Q9, VTK::BINARY>(argv, withGeo);
    if (lattice == "D3Q19") return run<D3Q19, VTK::BINARY>(argv, withGeo);
    if (lattice == "D3Q27") return run<D3Q27, VTK::BINARY>(argv, withGeo);
    chimp_host::die("unknown lattice " + lattice);
}
