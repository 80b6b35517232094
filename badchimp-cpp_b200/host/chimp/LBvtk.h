// Reader of the reference's per-rank ASCII geometry file (.vtklb, doc/fileformat.md, written by
// PythonScripts/vtklb.py:222-325).  Same public interface as the reference's LBvtk<DXQY>
// (src/lbsolver/LBvtk.h:52-175) for the calls the three target mains make, but the file is parsed
// once into memory (the reference re-seeks into the stream and keeps `int` offsets, LBvtk.h:194-201,
// which limits it to files < 2 GiB).
#ifndef CHIMP_LBVTK_H
#define CHIMP_LBVTK_H

#include <fstream>
#include <map>
#include <sstream>

#include "LBglobal.h"

template <typename DXQY>
class LBvtk
{
public:
    explicit LBvtk(const std::string &filename) : filename_(filename)
    {
        std::ifstream in(filename);
        if (!in) chimp_host::die("reading file in LBvtk: could not open file " + filename + ".");
        std::string line, word;
        std::getline(in, line); // header
        std::getline(in, line); // title
        in >> word;             // ASCII | BINARY
        if (word != "ASCII") chimp_host::die("Files data type must be ASCII");
        in >> word;
        if (word != "DATASET") chimp_host::die("expected DATASET in " + filename);
        in >> word;
        if (word != "UNSTRUCTURED_LB_GRID") chimp_host::die("expected UNSTRUCTURED_LB_GRID in " + filename);
        nD_ = 3;
        zeroGhost_ = false;
        while (in >> word) {
            if (word == "NUM_DIMENSIONS") in >> nD_;
            else if (word == "GLOBAL_DIMENSIONS") { globalDim_.resize(nD_); for (auto &x : globalDim_) in >> x; }
            else if (word == "USE_ZERO_GHOST_NODE") zeroGhost_ = true;
            else if (word == "POINTS") break;
            else chimp_host::die("unexpected keyword " + word + " in " + filename);
        }
        if (nD_ != DXQY::nD) chimp_host::die("wrong number of dimensions in " + filename);
        in >> nPoints_ >> word;
        pos_.resize(std::size_t(nPoints_) * nD_);
        for (auto &x : pos_) in >> x;
        int nq;
        in >> word >> nq >> line;
        if (word != "LATTICE") chimp_host::die("Expected key word LATTICE got " + word);
        if (nq != DXQY::nQ) chimp_host::die("Wrong number of basis vectors in " + filename);
        std::vector<int> f2p(nq);
        std::vector<bool> seen(nq, false);
        for (int q = 0; q < nq; ++q) {
            std::vector<int> v(nD_);
            for (auto &x : v) in >> x;
            f2p[q] = DXQY::c2q(v); // file basis -> header basis (LBvtk.h:398-440)
            if (f2p[q] < 0 || seen[f2p[q]]) chimp_host::die("The LATTICE basis in file " + filename + " is not complete");
            seen[f2p[q]] = true;
        }
        in >> word >> line;
        if (word != "NEIGHBORS") chimp_host::die("Expected NEIGHBORS got " + word);
        neig_.assign(std::size_t(nPoints_) * nq, 0);
        for (int n = 0; n < nPoints_; ++n)
            for (int q = 0; q < nq; ++q) in >> neig_[std::size_t(n) * nq + f2p[q]];
        in >> word >> rank_;
        if (word != "PARALLEL_COMPUTING") chimp_host::die("Expected PARALLEL_COMPUTING, but got " + word);
        while (in >> word && word == "PROCESSOR") {
            int n, r;
            in >> n >> r;
            std::vector<int> cur(n), adj(n);
            for (int k = 0; k < n; ++k) in >> cur[k] >> adj[k];
            procs_.push_back({r, cur, adj});
        }
        std::stable_sort(procs_.begin(), procs_.end(), [](const Proc &a, const Proc &b) { return a.rank < b.rank; }); // LBvtk.h:534-544
        if (word != "POINT_DATA") chimp_host::die("Expected POINT_DATA got " + word);
        int nData;
        in >> nData;
        while (in >> word) {
            if (word == "POINT_DATA_SUBSET") break; // subsets are not used by the three target mains
            if (word != "SCALARS") chimp_host::die("Expected SCALARS got " + word);
            std::string name, type;
            in >> name >> type;
            std::vector<double> v(nPoints_);
            for (auto &x : v) in >> x;
            attr_[name] = std::move(v);
        }
    }

    int beginNodeNo() const { return zeroGhost_ ? 1 : 0; }
    int endNodeNo() const { return nPoints_ + beginNodeNo(); }
    int getRank() const { return rank_; }
    int getGlobaDimensions(int d) const { return globalDim_[d]; }
    int getNumNeigProc() const { return int(procs_.size()); }
    int getNeigRank(int n) const { return procs_[n].rank; }
    std::vector<int> getNeigNodesNo(int n) const { return procs_[n].cur; }
    std::vector<int> getNeigNodesNeigNo(int n) const { return procs_[n].adj; }

    void toPos() { cursor_ = 0; }
    template <typename T>
    std::vector<T> getPos()
    {
        std::vector<T> p(pos_.begin() + cursor_ * nD_, pos_.begin() + (cursor_ + 1) * nD_);
        ++cursor_;
        return p;
    }
    void toNeighbors() { cursor_ = 0; }
    template <typename T>
    std::vector<T> getNeighbors()
    {
        std::vector<T> p(neig_.begin() + cursor_ * DXQY::nQ, neig_.begin() + (cursor_ + 1) * DXQY::nQ);
        ++cursor_;
        return p;
    }
    void toAttribute(const std::string &name)
    {
        auto it = attr_.find(name);
        if (it == attr_.end()) chimp_host::die("Could not find attribute " + name + " in " + filename_);
        cur_ = &it->second;
        cursor_ = 0;
    }
    template <typename T>
    T getScalarAttribute() { return static_cast<T>((*cur_)[cursor_++]); }
    template <typename T>
    T getScalar() { return getScalarAttribute<T>(); }
    bool hasAttribute(const std::string &name) const { return attr_.count(name) != 0; }
    const std::string &fileName() const { return filename_; }

private:
    struct Proc { int rank; std::vector<int> cur, adj; };
    std::string filename_;
    int nD_, nPoints_ = 0, rank_ = 0;
    bool zeroGhost_;
    std::vector<int> globalDim_, pos_, neig_;
    std::vector<Proc> procs_;
    std::map<std::string, std::vector<double>> attr_;
    const std::vector<double> *cur_ = nullptr;
    std::size_t cursor_ = 0;
};

#endif
