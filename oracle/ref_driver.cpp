// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Thin harness around the UNMODIFIED reference headers (compiled with
// -I /root/reference/src, see oracle/Makefile; output goes to oracle/_ref/).  It loads
// the per-rank .vtklb files through the reference's own LBvtk/Grid/Nodes/BndMpi, then
// drives the reference's own per-node functions (calcRho, calcVel, calcOmegaBGK[TRT],
// calcDeltaOmegaF[TRT], calcDeltaOmegaQ, calcDeltaOmegaST/RC, grad, LbField::propagateTo,
// BndMpi::communicate*, HalfWayBounceBack::apply) in the order the three target mains
// call them:
//   --case std_case   : src/std_case/main.cpp:89-145
//   --case one_phase  : src/std_one_phase/main.cpp:27-203 (link finders / BCs), 441-597 (loop)
//   --case twophase   : src/twophase/main_TWOPHASE.cpp:81-392
// and dumps raw integer tables and raw double fields as tagged records that
// oracle/recfile.py reads.  It replaces the mains only because they hard-code paths and
// never write f.  N "ranks" run as threads over oracle/mpi_shim.
//
// Record format: "REC1" | u32 namelen | name | char dtype ('i' int32, 'd' float64) |
//                u64 count | payload.
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <valarray>
#include <vector>

#include <mpi.h>
#include "io/Input.h" // Block is needed by LButilities.h:68
#define private public // read-only access to MonLatMpi / BndMpi lists for the table dump
#include "lbsolver/LBmonlatmpi.h"
#include "lbsolver/LBbndmpi.h"
#undef private
#include "lbsolver/LBlatticetypes.h"
#include "lbsolver/LBfield.h"
#include "lbsolver/LBgrid.h"
#include "lbsolver/LBnodes.h"
#include "lbsolver/LBgeometry.h"
#include "lbsolver/LBhalfwaybb.h"
#include "lbsolver/LBpressurebnd.h"
#include "lbsolver/LBmacroscopic.h"
#include "lbsolver/LBcollision.h"
#include "lbsolver/LBcollision2phase.h"
#include "lbsolver/LBinitiatefield.h"
#include "lbsolver/LButilities.h"
#include "lbsolver/LBglobalforcing.h" // calcFluxForceCartDir / calcCapNumbForceCartDir for the --global-forcing records
#include "lbsolver/LBvtk.h"
#include "io/Output.h" // the reference's VTK writer, for the --vtk goldens
#include "LBd3q27.h"

namespace {

struct Opts {
    std::string kase = "std_case", lattice = "D3Q19", dir = ".", prefix = "tmp", out = ".";
    int nranks = 1, steps = 1;
    std::set<int> dumpSteps;
    bool dumpTables = true, dumpF = true, timing = false;
    std::string vtk;        // directory for the reference's own Output<LT>::write() after the last step
    bool vtkAscii = false;  // ... through Output<LT, double, VTK::ASCII>
    std::string checkpoint; // prefix for the reference's own writeToFile() dumps after the last step
    int checkpointAt = -1;  // ... or after this step instead (the run continues)
    std::string restart;    // prefix of .lblbf files the reference's own LbField::readFromFile() loads before the first step
    bool globalForcing = false; // at every dump step: the reference's own calcFluxForceCartDir / calcCapNumbForceCartDir
    double fixedFlux = 1e-5, sigmaCapNumb = 1e-4, nu0 = 1.0 / 6.0, nu1 = 0.1;
    double tau = 0.8, tauSym = 0.0, tauAnti = 0.0; // TRT when tauSym > 0
    std::vector<double> force{0, 0, 0};
    double tau0 = 1, tau1 = 1, sigma = 0.01, beta = 1, momx = 1e-5; // twophase
    double rhoW = 1.0;                                             // one_phase pressure bnd
    // std_case: the reference's own PressureBnd / InletOutlet (LBpressurebnd.h:10-88) applied after the bounce back on
    // every `bndEvery`-th fluid boundary node: "pressure" (rho of node n prescribed as 1 + 0.01 (n % 7)) or "inletoutlet"
    std::string pressureBnd;
    int bndEvery = 3;
    double ioRho = 1.02;
    std::vector<double> ioVel{0.01, -0.005, 0.002};
};

class RecFile
{
public:
    explicit RecFile(const std::string &name) : ofs_(name, std::ios::binary) {
        if (!ofs_) { std::cerr << "cannot open " << name << std::endl; std::exit(1); }
    }
    void ints(const std::string &name, const int *p, std::size_t n) { head(name, 'i', n); ofs_.write((const char *)p, 4 * n); }
    void ints(const std::string &name, const std::vector<int> &v) { ints(name, v.data(), v.size()); }
    void reals(const std::string &name, const double *p, std::size_t n) { head(name, 'd', n); ofs_.write((const char *)p, 8 * n); }
    void reals(const std::string &name, const std::vector<double> &v) { reals(name, v.data(), v.size()); }
    void scalar(const std::string &name, int v) { ints(name, &v, 1); }
    void scalar(const std::string &name, double v) { reals(name, &v, 1); }
private:
    void head(const std::string &name, char t, std::size_t n) {
        ofs_.write("REC1", 4);
        std::uint32_t l = (std::uint32_t)name.size();
        ofs_.write((const char *)&l, 4);
        ofs_.write(name.data(), l);
        ofs_.write(&t, 1);
        std::uint64_t c = n;
        ofs_.write((const char *)&c, 8);
    }
    std::ofstream ofs_;
};

template <class LT>
std::vector<double> flat(const LbField<LT> &f, int nNodes)
{
    std::vector<double> v((std::size_t)f.num_fields() * LT::nQ * nNodes);
    std::size_t k = 0;
    for (int n = 0; n < nNodes; ++n)
        for (int fld = 0; fld < f.num_fields(); ++fld)
            for (int q = 0; q < LT::nQ; ++q) v[k++] = f(fld, q, n);
    return v;
}

std::vector<double> flat(const ScalarField &s, int nNodes)
{
    std::vector<double> v((std::size_t)s.num_fields() * nNodes);
    std::size_t k = 0;
    for (int n = 0; n < nNodes; ++n)
        for (int fld = 0; fld < s.num_fields(); ++fld) v[k++] = s(fld, n);
    return v;
}

template <class LT>
std::vector<double> flat(const VectorField<LT> &s, int nNodes)
{
    std::vector<double> v((std::size_t)s.num_fields() * LT::nD * nNodes);
    std::size_t k = 0;
    for (int n = 0; n < nNodes; ++n)
        for (int fld = 0; fld < s.num_fields(); ++fld)
            for (int d = 0; d < LT::nD; ++d) v[k++] = s(fld, d, n);
    return v;
}

template <class LT>
void dumpGridTables(RecFile &rec, const Grid<LT> &grid, const Nodes<LT> &nodes, BndMpi<LT> &mpi,
                    const std::vector<int> &bulk)
{
    const int sz = grid.size();
    rec.scalar("nD", LT::nD);
    rec.scalar("nQ", LT::nQ);
    rec.scalar("size", sz);
    std::vector<int> neigh((std::size_t)sz * LT::nQ), pos((std::size_t)sz * LT::nD), type(sz), rank(sz);
    for (int n = 0; n < sz; ++n) {
        for (int q = 0; q < LT::nQ; ++q) neigh[(std::size_t)n * LT::nQ + q] = grid.neighbor(q, n);
        for (int d = 0; d < LT::nD; ++d) pos[(std::size_t)n * LT::nD + d] = grid.pos(n, d);
        type[n] = nodes.getType(n);
        rank[n] = nodes.getRank(n);
    }
    rec.ints("neigh", neigh);
    rec.ints("pos", pos);
    rec.ints("type", type);
    rec.ints("rank", rank);
    rec.ints("bulk", bulk);
    rec.ints("fluidBnd", findFluidBndNodes(nodes));
    rec.ints("solidBnd", findSolidBndNodes(nodes));
    rec.scalar("nNeigRanks", (int)mpi.mpiList_.size());
    for (std::size_t k = 0; k < mpi.mpiList_.size(); ++k) {
        const MonLatMpi &m = mpi.mpiList_[k];
        const std::string p = "mpi" + std::to_string(k) + ".";
        rec.scalar(p + "neigRank", m.neigRank_);
        rec.ints(p + "nodesToSend", m.nodesToSend_);
        rec.ints(p + "nDirPerNodeToSend", m.nDirPerNodeToSend_);
        rec.ints(p + "dirListToSend", m.dirListToSend_);
        rec.ints(p + "nodesReceived", m.nodesReceived_);
        rec.ints(p + "nDirPerNodeReceived", m.nDirPerNodeReceived_);
        rec.ints(p + "dirListReceived", m.dirListReceived_);
    }
}

template <class LT>
void dumpBounceBack(RecFile &rec, const std::string &prefix, const HalfWayBounceBack<LT> &bb)
{
    std::vector<int> node, nb, ng, nd, links;
    for (int n = 0; n < bb.size(); ++n) {
        node.push_back(bb.nodeNo(n));
        nb.push_back(bb.nBeta(n));
        ng.push_back(bb.nGamma(n));
        nd.push_back(bb.nDelta(n));
        for (auto q : bb.beta(n)) links.push_back(q);
        for (auto q : bb.gamma(n)) links.push_back(q);
        for (auto q : bb.delta(n)) links.push_back(q);
    }
    rec.ints(prefix + "node", node);
    rec.ints(prefix + "nBeta", nb);
    rec.ints(prefix + "nGamma", ng);
    rec.ints(prefix + "nDelta", nd);
    rec.ints(prefix + "links", links);
}

template <class LT>
std::valarray<lbBase_t> forceArray(const Opts &o)
{
    std::valarray<lbBase_t> F(LT::nD);
    for (int d = 0; d < LT::nD; ++d) F[d] = o.force[d];
    return F;
}

// ---------------------------------------------------------------------------------------
// std_case: BGK (or TRT) + Guo force + half-way bounce back on the fluid boundary nodes
// ---------------------------------------------------------------------------------------
template <class LT>
double runStdCase(const Opts &o, RecFile &rec, LBvtk<LT> &vtklb, Grid<LT> &grid, Nodes<LT> &nodes,
                  BndMpi<LT> &mpi, const std::vector<int> &bulk)
{
    const int sz = grid.size();
    VectorField<LT> bodyForce(1, 1);
    bodyForce.set(0, 0) = forceArray<LT>(o);
    ScalarField rho(1, sz);
    vtklb.toAttribute("init_rho");
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) rho(0, n) = vtklb.template getScalarAttribute<lbBase_t>();
    VectorField<LT> vel(1, sz);
    for (auto n : bulk)
        for (int d = 0; d < LT::nD; ++d) vel(0, d, n) = 0.0;
    HalfWayBounceBack<LT> bb(findFluidBndNodes(nodes), nodes, grid);
    if (o.dumpTables) dumpBounceBack(rec, "bb.", bb);
    // library pressure boundaries (no caller in the reference's mains; exercised here through their own classes)
    std::vector<int> pbNodes;
    if (!o.pressureBnd.empty()) {
        const std::vector<int> fb = findFluidBndNodes(nodes);
        for (std::size_t k = 0; k < fb.size(); k += (std::size_t)o.bndEvery) pbNodes.push_back(fb[k]);
    }
    PressureBnd<LT> pressureBnd(pbNodes, nodes, grid);
    InletOutlet<LT> inletOutlet(pbNodes, nodes, grid);
    ScalarField rhoBnd(1, sz);
    for (int n = 0; n < sz; ++n) rhoBnd(0, n) = 1.0 + 0.01 * (n % 7);
    const std::vector<lbBase_t> ioVel(o.ioVel.begin(), o.ioVel.begin() + LT::nD);
    if (!o.pressureBnd.empty()) {
        rec.ints("pbnd.nodes", pbNodes);
        std::vector<int> nb, nd, links;
        for (int b = 0; b < pressureBnd.size(); ++b) {
            nb.push_back((int)pressureBnd.beta(b).size());
            nd.push_back((int)pressureBnd.delta(b).size());
            for (int q : pressureBnd.beta(b)) links.push_back(q);
            for (int q : pressureBnd.delta(b)) links.push_back(q);
        }
        rec.ints("pbnd.nBeta", nb);
        rec.ints("pbnd.nDelta", nd);
        rec.ints("pbnd.links", links);
    }
    LbField<LT> f(1, sz), fTmp(1, sz);
    for (auto n : bulk)
        for (int q = 0; q < LT::nQ; ++q) f(0, q, n) = LT::w[q] * rho(0, n);

    if (!o.restart.empty()) f.readFromFile(o.restart + std::to_string(vtklb.getRank())); // LBfield.h:394-421

    const bool trt = o.tauSym > 0.0;
    const lbBase_t tau = o.tau;
    int numNodes = (int)bulk.size(), numNodesGlobal = 0;
    MPI_Allreduce(&numNodes, &numNodesGlobal, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    auto dump = [&](int step) {
        if (o.globalForcing && o.dumpSteps.count(step)) { // LBglobalforcing.h:8-33, every rank takes part in its all-reduce
            std::vector<double> ff(LT::nD);
            for (int d = 0; d < LT::nD; ++d) ff[d] = calcFluxForceCartDir<LT>(0, f, bulk, d, o.fixedFlux, numNodesGlobal);
            rec.reals("step" + std::to_string(step) + ".fluxForce", ff);
        }
        if (step == o.checkpointAt && !o.checkpoint.empty()) {
            const std::string p = o.checkpoint + std::to_string(vtklb.getRank());
            f.writeToFile(p);
            rho.writeToFile(p);
            vel.writeToFile(p);
        }
        if (!o.dumpF || !o.dumpSteps.count(step)) return;
        const std::string s = "step" + std::to_string(step) + ".";
        rec.reals(s + "f", flat(f, sz));
        rec.reals(s + "rho", flat(rho, sz));
        rec.reals(s + "vel", flat(vel, sz));
    };
    dump(0);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 1; i <= o.steps; ++i) {
        for (auto nodeNo : bulk) {
            const std::valarray<lbBase_t> fNode = f(0, nodeNo);
            const lbBase_t rhoNode = calcRho<LT>(fNode);
            const auto velNode = calcVel<LT>(fNode, rhoNode, bodyForce(0, 0));
            rho(0, nodeNo) = rhoNode;
            vel.set(0, nodeNo) = velNode;
            const lbBase_t u2 = LT::dot(velNode, velNode);
            const std::valarray<lbBase_t> cu = LT::cDotAll(velNode);
            const lbBase_t uF = LT::dot(velNode, bodyForce(0, 0));
            const std::valarray<lbBase_t> cF = LT::cDotAll(bodyForce(0, 0));
            if (trt) {
                const std::valarray<lbBase_t> omega = calcOmegaBGKTRT<LT>(fNode, o.tauSym, o.tauAnti, rhoNode, u2, cu);
                const std::valarray<lbBase_t> dOmegaF = calcDeltaOmegaFTRT<LT>(o.tauSym, o.tauAnti, 1.0, cu, uF, cF);
                fTmp.propagateTo(0, nodeNo, fNode + omega + dOmegaF, grid);
            } else {
                const std::valarray<lbBase_t> omega = calcOmegaBGK<LT>(fNode, tau, rhoNode, u2, cu);
                const std::valarray<lbBase_t> dOmegaF = calcDeltaOmegaF<LT>(tau, cu, uF, cF);
                fTmp.propagateTo(0, nodeNo, fNode + omega + dOmegaF, grid);
            }
        }
        f.swapData(fTmp);
        mpi.communicateLbField(0, f, grid);
        bb.apply(f, grid);
        if (o.pressureBnd == "pressure") pressureBnd.apply(0, f, grid, rhoBnd);        // LBpressurebnd.h:19-41
        else if (o.pressureBnd == "inletoutlet") inletOutlet.apply(0, f, grid, o.ioRho, ioVel); // LBpressurebnd.h:51-88
        dump(i);
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!o.checkpoint.empty() && o.checkpointAt < 0) { // LBfield.h:102-114, 233-247, 378-392
        const std::string p = o.checkpoint + std::to_string(vtklb.getRank());
        f.writeToFile(p);
        rho.writeToFile(p);
        vel.writeToFile(p);
    }
    if (!o.vtk.empty()) { // std_case/main.cpp:101-104,149-151
        int nProcs = 1;
        MPI_Comm_size(MPI_COMM_WORLD, &nProcs);
        std::vector<int> bulkNodes = bulk;
        if (o.vtkAscii) {
            Output<LT, double, VTK::ASCII> output(grid, bulkNodes, o.vtk, vtklb.getRank(), nProcs);
            output.add_file("lb_run");
            output.add_scalar_variables({"rho"}, {rho});
            output.add_vector_variables({"vel"}, {vel});
            output.write(o.steps);
        } else {
            Output<LT> output(grid, bulkNodes, o.vtk, vtklb.getRank(), nProcs);
            output.add_file("lb_run");
            output.add_scalar_variables({"rho"}, {rho});
            output.add_vector_variables({"vel"}, {vel});
            output.write(o.steps);
        }
    }
    return secs;
}

// ---------------------------------------------------------------------------------------
// one_phase: std_one_phase/main.cpp -- BGK (or TRT) + masked Guo force + mass-conservation
// source, link boundaries (solid bounce back, anti bounce back pressure, fluid-fluid swap)
// ---------------------------------------------------------------------------------------
template <class LT>
std::vector<std::vector<int>> tagLinks(const Nodes<LT> &nodes, const Grid<LT> &grid, int bit, int wantTag, int needPhase)
{
    // std_one_phase/main.cpp:27-126: node carries tag bit `bit`; neighbor's (tag & 3) == wantTag
    std::vector<std::vector<int>> ret;
    for (int n = 1; n < nodes.size(); n++) {
        const int flagged = (nodes.getTag(n) >> bit) & 1;
        if (!flagged || !nodes.isMyRank(n)) continue;
        if (needPhase >= 0 && (nodes.getTag(n) & 3) != needPhase) continue;
        for (int q = 0; q < LT::nQNonZero_; ++q) {
            const int nn = grid.neighbor(q, n);
            if ((nodes.getTag(nn) & 3) == wantTag) ret.push_back({n, LT::reverseDirection(q), nn, q});
        }
    }
    return ret;
}

inline std::vector<int> flatLinks(const std::vector<std::vector<int>> &l)
{
    std::vector<int> v;
    for (auto &x : l) v.insert(v.end(), x.begin(), x.end());
    return v;
}

template <class LT>
double runOnePhase(const Opts &o, RecFile &rec, LBvtk<LT> &vtklb, Grid<LT> &grid, Nodes<LT> &nodes,
                   BndMpi<LT> &mpi, const std::vector<int> &bulk)
{
    const int sz = grid.size();
    auto readInt = [&](const char *name, std::vector<int> &dst) {
        dst.assign(sz, 0);
        vtklb.toAttribute(name);
        for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) dst[n] = vtklb.template getScalarAttribute<int>();
    };
    std::vector<int> tags, forceFlag, interior;
    readInt("nodetags", tags);
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) nodes.setTag(tags[n], n);
    readInt("force", forceFlag);
    readInt("interior_domains", interior);
    ScalarField forceOn(1, sz);
    for (int n = 0; n < sz; ++n) forceOn(0, n) = forceFlag[n];
    VectorField<LT> normals(1, sz);
    const char *nrm[3] = {"normal_x", "normal_y", "normal_z"};
    for (int d = 0; d < LT::nD; ++d) {
        vtklb.toAttribute(nrm[d]);
        for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) normals(0, d, n) = vtklb.template getScalarAttribute<lbBase_t>();
    }
    int localMax = 0, globalMax = 0;
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n)
        if (nodes.isMyRank(n) && interior[n] > localMax) localMax = interior[n];
    MPI_Allreduce(&localMax, &globalMax, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
    std::vector<lbBase_t> scale(globalMax + 1, 0), massLocal(globalMax + 1, 0), mass(globalMax + 1, 0), addSource(sz, 0.0);
    {
        std::vector<lbBase_t> cnt(globalMax + 1, 0.0);
        for (int n = 1; n < sz; ++n)
            if (nodes.isMyRank(n) && interior[n] > 0 && nodes.getTag(n) < 3) { cnt[interior[n]] += 1; addSource[n] = 1.0; }
        MPI_Allreduce(cnt.data(), scale.data(), globalMax + 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
        for (int i = 1; i < globalMax + 1; ++i) scale[i] = 1.0 / scale[i];
    }
    const auto solidLinks = tagLinks(nodes, grid, 3, 0, -1);
    const auto pressLinks = tagLinks(nodes, grid, 4, 3, -1);
    const auto fluidLinks = tagLinks(nodes, grid, 2, 2, 1);
    if (o.dumpTables) {
        rec.ints("tags", tags);
        rec.ints("interior", interior);
        rec.reals("addSource", addSource);
        rec.reals("scale", scale);
        rec.ints("solidLinks", flatLinks(solidLinks));
        rec.ints("pressLinks", flatLinks(pressLinks));
        rec.ints("fluidLinks", flatLinks(fluidLinks));
    }
    VectorField<LT> bodyForce(1, 1);
    bodyForce.set(0, 0) = forceArray<LT>(o);
    ScalarField rho(1, sz);
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) rho(0, n) = 1.0;
    VectorField<LT> vel(1, sz);
    LbField<LT> f(1, sz), fTmp(1, sz);
    for (auto n : bulk) {
        vel.set(0, n) = 0;
        auto u2 = LT::dot(vel(0, n), vel(0, n));
        auto cu = LT::cDotAll(vel(0, n));
        f.set(0, n) = calcfeq<LT>(rho(0, n), u2, cu);
        fTmp.set(0, n) = 0;
    }
    const bool trt = o.tauSym > 0.0;
    const lbBase_t tau = o.tau;
    auto dump = [&](int step) {
        if (!o.dumpF || !o.dumpSteps.count(step)) return;
        const std::string s = "step" + std::to_string(step) + ".";
        rec.reals(s + "f", flat(f, sz));
        rec.reals(s + "rho", flat(rho, sz));
        rec.reals(s + "vel", flat(vel, sz));
        rec.reals(s + "massChange", mass);
    };
    dump(0);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 1; i <= o.steps; ++i) {
        std::fill(massLocal.begin(), massLocal.end(), 0.0);
        for (auto nodeNo : bulk) {
            const std::valarray<lbBase_t> fNode = f(0, nodeNo);
            massLocal[interior[nodeNo]] += 1.0 - calcRho<LT>(fNode);
        }
        MPI_Allreduce(massLocal.data(), mass.data(), (int)mass.size(), MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
        for (auto nodeNo : bulk) {
            const std::valarray<lbBase_t> fNode = f(0, nodeNo);
            lbBase_t rhoNode = calcRho<LT>(fNode);
            const int label = interior[nodeNo];
            const lbBase_t qSrc = 0.9 * 2 * scale[label] * mass[label] * addSource[nodeNo];
            rhoNode += 0.5 * qSrc;
            const std::valarray<lbBase_t> forceNode = bodyForce(0, 0) * forceOn(0, nodeNo);
            const auto velNode = calcVel<LT>(fNode, rhoNode, forceNode);
            rho(0, nodeNo) = rhoNode;
            vel.set(0, nodeNo) = velNode;
            const lbBase_t u2 = LT::dot(velNode, velNode);
            const std::valarray<lbBase_t> cu = LT::cDotAll(velNode);
            const lbBase_t uF = LT::dot(velNode, forceNode);
            const std::valarray<lbBase_t> cF = LT::cDotAll(forceNode);
            if (trt) {
                const std::valarray<lbBase_t> omega = calcOmegaBGKTRT<LT>(fNode, o.tauSym, o.tauAnti, rhoNode, u2, cu);
                const std::valarray<lbBase_t> dF = calcDeltaOmegaFTRT<LT>(o.tauSym, o.tauAnti, 1.0, cu, uF, cF);
                const std::valarray<lbBase_t> dQ = calcDeltaOmegaQTRT<LT>(o.tauSym, o.tauAnti, cu, u2, qSrc);
                fTmp.propagateTo(0, nodeNo, fNode + omega + dF + dQ, grid);
            } else {
                const std::valarray<lbBase_t> omega = calcOmegaBGK<LT>(fNode, tau, rhoNode, u2, cu);
                const std::valarray<lbBase_t> dF = calcDeltaOmegaF<LT>(tau, cu, uF, cF);
                const std::valarray<lbBase_t> dQ = calcDeltaOmegaQ<LT>(tau, cu, u2, qSrc);
                fTmp.propagateTo(0, nodeNo, fNode + omega + dF + dQ, grid);
            }
        }
        f.swapData(fTmp);
        mpi.communicateLbField(f, grid);
        mpi.communciateVectorField_TEST(vel);
        // std_one_phase/main.cpp:138-151
        for (auto &l : solidLinks) f(0, l[1], l[0]) = f(0, l[3], l[2]);
        // std_one_phase/main.cpp:155-174 (anti bounce back, rho_w)
        for (auto &l : pressLinks) {
            const lbBase_t u2 = LT::dot(vel(0, l[0]), vel(0, l[0]));
            const lbBase_t cu = LT::cDotRef(l[3], vel(0, l[0]));
            const lbBase_t w = LT::w[l[3]];
            f(0, l[1], l[0]) = -f(0, l[3], l[2]) + 2 * w * o.rhoW * (1 + 0.5 * (LT::c4Inv * cu * cu - LT::c2Inv * u2));
        }
        // std_one_phase/main.cpp:178-203 (the velocity correction dfu is multiplied by 0 there)
        for (auto &l : fluidLinks) {
            const lbBase_t f1 = f(0, l[1], l[0]);
            f(0, l[1], l[0]) = f(0, l[3], l[2]);
            f(0, l[3], l[2]) = f1;
        }
        dump(i);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---------------------------------------------------------------------------------------
// twophase: colour gradient, two LbFields (main_TWOPHASE.cpp:236-392)
// ---------------------------------------------------------------------------------------
template <class LT>
double runTwoPhase(const Opts &o, RecFile &rec, LBvtk<LT> &vtklb, Grid<LT> &grid, Nodes<LT> &nodes,
                   BndMpi<LT> &mpi, const std::vector<int> &bulk)
{
    const int sz = grid.size();
    HalfWayBounceBack<LT> bb(findBulkNodes(nodes), nodes, grid);
    const std::vector<int> solidBnd = findSolidBndNodes(nodes);
    if (o.dumpTables) dumpBounceBack(rec, "bb.", bb);
    VectorField<LT> bodyForce(1, 1);
    bodyForce.set(0, 0) = forceArray<LT>(o);
    const lbBase_t sigma = o.sigma, beta = o.beta;
    const lbBase_t nu0Inv = 1.0 / (LT::c2 * (o.tau0 - 0.5));
    const lbBase_t nu1Inv = 1.0 / (LT::c2 * (o.tau1 - 0.5));
    ScalarField Q(2, sz);
    for (int n = 0; n < Q.size(); ++n) { Q(0, n) = 0.0; Q(1, n) = 0.0; }
    LbField<LT> f(2, sz), fTmp(2, sz);
    ScalarField rho(2, sz);
    VectorField<LT> vel(1, sz);
    ScalarField cgField(1, sz);
    vtklb.toAttribute("rho0");
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) {
        float val = vtklb.template getScalar<float>();
        rho(0, n) = val;
        for (int d = 0; d < LT::nD; ++d) vel(0, d, n) = 0.0;
    }
    vtklb.toAttribute("rho1");
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) {
        float val = vtklb.template getScalar<float>();
        rho(1, n) = val;
    }
    vtklb.toAttribute("wettability");
    for (int n = vtklb.beginNodeNo(); n < vtklb.endNodeNo(); ++n) {
        float val = vtklb.template getScalar<float>();
        if (nodes.isSolidBoundary(n)) { rho(0, n) = val; rho(1, n) = 1 - val; }
    }
    int numNodes = (int)bulk.size(), numNodesGlobal = 0;
    MPI_Allreduce(&numNodes, &numNodesGlobal, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    initiateLbField(0, 0, 0, bulk, rho, vel, f);
    initiateLbField(1, 1, 0, bulk, rho, vel, f);
    if (o.dumpTables) rec.reals("rhoInit", flat(rho, sz));

    auto dump = [&](int step) {
        if (o.globalForcing && o.dumpSteps.count(step)) { // LBglobalforcing.h:8-98 on the state a main holds at this point
            std::vector<double> ff0(LT::nD), ff1(LT::nD), cap(LT::nD);
            for (int d = 0; d < LT::nD; ++d) {
                ff0[d] = calcFluxForceCartDir<LT>(0, f, bulk, d, o.fixedFlux, numNodesGlobal);
                ff1[d] = calcFluxForceCartDir<LT>(1, f, bulk, d, o.fixedFlux, numNodesGlobal);
                cap[d] = calcCapNumbForceCartDir<LT>(0, f, rho, bulk, d, o.sigmaCapNumb, o.nu0, o.nu1, numNodesGlobal);
            }
            const std::string s = "step" + std::to_string(step) + ".";
            rec.reals(s + "fluxForce0", ff0);
            rec.reals(s + "fluxForce1", ff1);
            rec.reals(s + "capForce", cap);
        }
        if (!o.dumpF || !o.dumpSteps.count(step)) return;
        const std::string s = "step" + std::to_string(step) + ".";
        rec.reals(s + "f", flat(f, sz));
        rec.reals(s + "rho", flat(rho, sz));
        rec.reals(s + "vel", flat(vel, sz));
        rec.reals(s + "cg", flat(cgField, sz));
        rec.scalar(s + "forceX", (double)bodyForce(0, 0, 0));
    };
    dump(0);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 1; i <= o.steps; ++i) {
        for (auto nodeNo : bulk) {
            lbBase_t rho0Node = rho(0, nodeNo) = calcRho<LT>(f(0, nodeNo));
            lbBase_t rho1Node = rho(1, nodeNo) = calcRho<LT>(f(1, nodeNo));
            cgField(0, nodeNo) = (rho0Node - rho1Node) / (rho0Node + rho1Node);
        }
        for (auto nodeNo : solidBnd) {
            const lbBase_t rho0Node = rho(0, nodeNo);
            const lbBase_t rho1Node = rho(1, nodeNo);
            cgField(0, nodeNo) = (rho0Node - rho1Node) / (rho0Node + rho1Node);
        }
        mpi.communciateScalarField(cgField);
        lbBase_t meanfcX = 0.0, meanfcXGlobal;
        for (auto nodeNo : bulk) {
            std::valarray<lbBase_t> fTot = f(0, nodeNo) + f(1, nodeNo);
            std::valarray<lbBase_t> sumfc = LT::qSumC(fTot);
            meanfcX += sumfc[0];
        }
        MPI_Allreduce(&meanfcX, &meanfcXGlobal, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
        meanfcXGlobal /= numNodesGlobal;
        bodyForce(0, 0, 0) = 2 * (o.momx - meanfcXGlobal);
        for (auto nodeNo : bulk) {
            std::valarray<lbBase_t> fTot = f(0, nodeNo) + f(1, nodeNo);
            lbBase_t rho0Node = rho(0, nodeNo);
            lbBase_t rho1Node = rho(1, nodeNo);
            lbBase_t rhoNode = rho0Node + rho1Node;
            std::valarray<lbBase_t> forceNode = bodyForce(0, 0);
            std::valarray<lbBase_t> velNode = calcVel<LT>(fTot, rhoNode, forceNode);
            vel.set(0, nodeNo) = velNode;
            lbBase_t q0Node = Q(0, nodeNo);
            lbBase_t q1Node = Q(1, nodeNo);
            rho(0, nodeNo) = rho0Node += 0.5 * q0Node;
            rho(1, nodeNo) = rho1Node += 0.5 * q1Node;
            lbBase_t tau = LT::c2Inv * rhoNode / (rho0Node * nu0Inv + rho1Node * nu1Inv) + 0.5;
            lbBase_t uu = LT::dot(velNode, velNode);
            std::valarray<lbBase_t> cu = LT::cDotAll(velNode);
            std::valarray<lbBase_t> omegaBGK = calcOmegaBGK<LT>(fTot, tau, rhoNode, uu, cu);
            lbBase_t uF = LT::dot(velNode, forceNode);
            std::valarray<lbBase_t> cF = LT::cDotAll(forceNode);
            std::valarray<lbBase_t> deltaOmegaF = calcDeltaOmegaF<LT>(tau, cu, uF, cF);
            std::valarray<lbBase_t> deltaOmegaQ0 = calcDeltaOmegaQ<LT>(tau, cu, uu, q0Node);
            std::valarray<lbBase_t> deltaOmegaQ1 = calcDeltaOmegaQ<LT>(tau, cu, uu, q1Node);
            std::valarray<lbBase_t> colorGradNode = grad(cgField, 0, nodeNo, grid);
            lbBase_t CGNorm = vecNorm<LT>(colorGradNode);
            colorGradNode *= 1.0 / (CGNorm + (CGNorm < lbBaseEps));
            std::valarray<lbBase_t> cCGNorm = LT::cDotAll(colorGradNode);
            std::valarray<lbBase_t> deltaOmegaST = calcDeltaOmegaST<LT>(tau, sigma, CGNorm, cCGNorm);
            std::valarray<lbBase_t> deltaOmegaRC = calcDeltaOmegaRC<LT>(beta, rho0Node, rho1Node, rhoNode, cCGNorm);
            lbBase_t c0 = (rho0Node / rhoNode), c1 = (rho1Node / rhoNode);
            for (int q = 0; q < LT::nQ; ++q) {
                fTmp(0, q, grid.neighbor(q, nodeNo)) = c0 * (fTot[q] + omegaBGK[q] + deltaOmegaF[q] + deltaOmegaST[q]) + deltaOmegaRC[q] + deltaOmegaQ0[q];
                fTmp(1, q, grid.neighbor(q, nodeNo)) = c1 * (fTot[q] + omegaBGK[q] + deltaOmegaF[q] + deltaOmegaST[q]) - deltaOmegaRC[q] + deltaOmegaQ1[q];
            }
        }
        f.swapData(fTmp);
        mpi.communicateLbField(0, f, grid);
        mpi.communicateLbField(1, f, grid);
        bb.apply(0, f, grid);
        bb.apply(1, f, grid);
        dump(i);
    }
    if (!o.vtk.empty()) { // main_TWOPHASE.cpp:214-223,424
        int nProcs = 1;
        MPI_Comm_size(MPI_COMM_WORLD, &nProcs);
        std::vector<int> bulkNodes = bulk;
        Output<LT> output(grid, bulkNodes, o.vtk, vtklb.getRank(), nProcs);
        output.add_file("fluid");
        output.add_scalar_variables({"rho"}, {rho});
        output.add_vector_variables({"vel"}, {vel});
        auto geo = nodes.geo(grid, vtklb);
        Output<LT, int> geoout(grid.pos(), o.vtk, vtklb.getRank(), nProcs, "geo", geo);
        geoout.write();
        output.write(o.steps);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

template <class LT>
struct HasColourGradient { static constexpr bool value = true; };
template <>
struct HasColourGradient<D3Q27> { static constexpr bool value = false; };

std::mutex g_ioMutex;
std::vector<double> g_loopSeconds;
std::vector<long> g_bulkCount;

template <class LT>
void rankMain(const Opts &o, int rank)
{
    mpishim::my_rank() = rank;
    LBvtk<LT> vtklb(o.dir + "/" + o.prefix + std::to_string(rank) + ".vtklb");
    Grid<LT> grid(vtklb);
    Nodes<LT> nodes(vtklb, grid);
    BndMpi<LT> mpi(vtklb, nodes, grid);
    std::vector<int> bulk = findBulkNodes(nodes);
    RecFile rec(o.out + "/rank" + std::to_string(rank) + ".rec");
    if (o.dumpTables) dumpGridTables(rec, grid, nodes, mpi, bulk);
    double secs = 0;
    if (o.kase == "std_case") secs = runStdCase<LT>(o, rec, vtklb, grid, nodes, mpi, bulk);
    else if (o.kase == "one_phase") secs = runOnePhase<LT>(o, rec, vtklb, grid, nodes, mpi, bulk);
    else if (o.kase == "twophase") {
        if constexpr (HasColourGradient<LT>::value) secs = runTwoPhase<LT>(o, rec, vtklb, grid, nodes, mpi, bulk);
        else { std::cerr << "twophase needs a lattice with B[] weights" << std::endl; std::exit(1); }
    } else { std::cerr << "unknown case " << o.kase << std::endl; std::exit(1); }
    std::lock_guard<std::mutex> lk(g_ioMutex);
    g_loopSeconds[rank] = secs;
    g_bulkCount[rank] = (long)bulk.size();
}

std::vector<double> parseList(const std::string &s)
{
    std::vector<double> v;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) v.push_back(std::stod(tok));
    return v;
}

} // namespace

int main(int argc, char **argv)
{
    Opts o;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { std::cerr << "missing value for " << a << std::endl; std::exit(1); } return argv[++i]; };
        if (a == "--case") o.kase = next();
        else if (a == "--lattice") o.lattice = next();
        else if (a == "--dir") o.dir = next();
        else if (a == "--prefix") o.prefix = next();
        else if (a == "--out") o.out = next();
        else if (a == "--nranks") o.nranks = std::stoi(next());
        else if (a == "--steps") o.steps = std::stoi(next());
        else if (a == "--dump") { for (double d : parseList(next())) o.dumpSteps.insert((int)d); }
        else if (a == "--no-tables") o.dumpTables = false;
        else if (a == "--no-f") o.dumpF = false;
        else if (a == "--time") o.timing = true;
        else if (a == "--checkpoint") o.checkpoint = next();
        else if (a == "--checkpoint-at") o.checkpointAt = std::stoi(next());
        else if (a == "--restart") o.restart = next();
        else if (a == "--global-forcing") o.globalForcing = true;
        else if (a == "--fixed-flux") o.fixedFlux = std::stod(next());
        else if (a == "--cap-numb") { auto v = parseList(next()); o.sigmaCapNumb = v[0]; o.nu0 = v[1]; o.nu1 = v[2]; }
        else if (a == "--vtk") o.vtk = next();
        else if (a == "--vtk-ascii") o.vtkAscii = true;
        else if (a == "--tau") o.tau = std::stod(next());
        else if (a == "--trt") { auto v = parseList(next()); o.tauSym = v[0]; o.tauAnti = v[1]; }
        else if (a == "--force") { auto v = parseList(next()); v.resize(3, 0.0); o.force = v; }
        else if (a == "--tau2") { auto v = parseList(next()); o.tau0 = v[0]; o.tau1 = v[1]; }
        else if (a == "--sigma") o.sigma = std::stod(next());
        else if (a == "--beta") o.beta = std::stod(next());
        else if (a == "--momx") o.momx = std::stod(next());
        else if (a == "--rhow") o.rhoW = std::stod(next());
        else if (a == "--pressure-bnd") o.pressureBnd = next();
        else if (a == "--bnd-every") o.bndEvery = std::max(1, std::stoi(next()));
        else if (a == "--io-rho") o.ioRho = std::stod(next());
        else if (a == "--io-vel") { auto v = parseList(next()); v.resize(3, 0.0); o.ioVel = v; }
        else { std::cerr << "unknown option " << a << std::endl; return 1; }
    }
    mpishim::init(o.nranks);
    g_loopSeconds.assign(o.nranks, 0.0);
    g_bulkCount.assign(o.nranks, 0);
    std::vector<std::thread> th;
    for (int r = 0; r < o.nranks; ++r) {
        if (o.lattice == "D2Q9") th.emplace_back(rankMain<D2Q9>, std::cref(o), r);
        else if (o.lattice == "D3Q19") th.emplace_back(rankMain<D3Q19>, std::cref(o), r);
        else if (o.lattice == "D3Q27") th.emplace_back(rankMain<D3Q27>, std::cref(o), r);
        else { std::cerr << "unknown lattice " << o.lattice << std::endl; return 1; }
    }
    for (auto &t : th) t.join();
    if (o.timing) {
        double tmax = 0; long nodes = 0;
        for (int r = 0; r < o.nranks; ++r) { tmax = std::max(tmax, g_loopSeconds[r]); nodes += g_bulkCount[r]; }
        std::printf("{\"loop_seconds\": %.6f, \"fluid_nodes\": %ld, \"steps\": %d, \"ranks\": %d, \"mlups\": %.6f}\n",
                    tmax, nodes, o.steps, o.nranks, nodes * (double)o.steps / tmax / 1e6);
    }
    return 0;
}
