// std_case physics on a raw voxel geometry: the C++ route to sizes the reference's loader cannot read.
//
//   voxel_case <D2Q9|D3Q19|D3Q27> <nx> <ny> <nz> <voxels.raw> <periodic axes, e.g. xyz or x or -> <iterations> <tau> <Fx> <Fy> <Fz> [out.bin]
//
// voxels.raw holds nx*ny*nz bytes in C-order (z fastest; 2-D: nz = 1), 0 = solid.  The reference reaches its Grid
// through vtklb.py -> ASCII .vtklb -> LBvtk (LBvtk.h:221-262), whose int offsets stop at 2 GiB of text
// (LBvtk.h:194-201); here the engine numbers the nodes itself (chimp_create_from_voxels: the reference's labels, 1..N in
// C-order) and the main keeps the reference's ScalarField / VectorField objects for the results.  Same loop body
// as std_case/main.cpp:109-146 (BGK + Guo force + half-way bounce back), rho = 1 and u = 0 initially (:62-96 with
// init_rho = 1).  Prints one JSON line; out.bin receives int N, rho[N+1], vel[(N+1)*nD] for the parity test.
#include <cstdio>
#include <fstream>

#include "../chimp/LBSOLVER.h"

template <typename LT>
int run(int nx, int ny, int nz, const std::string &file, const std::string &periodic, int nIterations, lbBase_t tau, const double *F,
        const std::string &outFile)
{
    std::vector<std::uint8_t> voxels(std::size_t(nx) * ny * nz);
    std::ifstream in(file, std::ios::binary);
    if (!in || !in.read(reinterpret_cast<char *>(voxels.data()), std::streamsize(voxels.size())))
        chimp_host::die("cannot read " + std::to_string(voxels.size()) + " voxels from " + file);
    int mask = 0;
    for (char ch : periodic) mask |= ch == 'x' ? 1 : ch == 'y' ? 2 : ch == 'z' ? 4 : 0;
    std::valarray<lbBase_t> bodyForce(LT::nD);
    for (int d = 0; d < LT::nD; ++d) bodyForce[d] = F[d];

    GpuLattice<LT> gpu(voxels, nx, ny, nz, mask, 1);
    const int n = gpu.numFluidNodes();
    gpu.initUniform(1.0);
    const double ms = gpu.stepBGKTimed(tau, bodyForce, nIterations);
    ScalarField rho(1, n + 1);
    VectorField<LT> vel(1, n + 1);
    gpu.download(rho, vel);
    double mean = 0.0;
    for (int i = 1; i <= n; ++i) mean += rho(0, i);
    mean /= n;
    std::printf("{\"fluid_nodes\": %d, \"iterations\": %d, \"ms_per_step\": %.6f, \"MLUPS\": %.3f, \"mean_rho\": %.15f}\n", n, nIterations,
                ms / nIterations, double(n) * nIterations / (ms * 1e-3) / 1e6, mean);
    if (!outFile.empty()) {
        FILE *fp = std::fopen(outFile.c_str(), "wb");
        if (!fp) chimp_host::die("cannot open " + outFile);
        std::fwrite(&n, sizeof(int), 1, fp);
        std::fwrite(rho.data(), sizeof(double), std::size_t(n) + 1, fp);
        std::fwrite(vel.data(), sizeof(double), (std::size_t(n) + 1) * LT::nD, fp);
        std::fclose(fp);
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 12) {
        std::cout << "usage: voxel_case <D2Q9|D3Q19|D3Q27> <nx> <ny> <nz> <voxels.raw> <periodic axes> <iterations> <tau> <Fx> <Fy> <Fz> [out.bin]" << std::endl;
        return 2;
    }
    const std::string lattice = argv[1], periodic = argv[6];
    const int nx = std::atoi(argv[2]), ny = std::atoi(argv[3]), nz = std::atoi(argv[4]), nIt = std::atoi(argv[7]);
    const lbBase_t tau = std::atof(argv[8]);
    const double F[3] = {std::atof(argv[9]), std::atof(argv[10]), std::atof(argv[11])};
    const std::string out = argc > 12 ? argv[12] : "";
    if (lattice == "D2Q9") return run<D2Q9>(nx, ny, nz, argv[5], periodic, nIt, tau, F, out);
    if (lattice == "D3Q19") return run<D3Q19>(nx, ny, nz, argv[5], periodic, nIt, tau, F, out);
    if (lattice == "D3Q27") return run<D3Q27>(nx, ny, nz, argv[5], periodic, nIt, tau, F, out);
    chimp_host::die("unknown lattice " + lattice);
}
