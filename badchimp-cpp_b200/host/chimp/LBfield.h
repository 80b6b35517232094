// ScalarField / VectorField / LbField with the reference's storage layout and accessors
// (src/lbsolver/LBfield.h:48-100, 149-231, 285-376): AoS, node-major,
//   ScalarField  data[nFields*node + field]
//   VectorField  data[(nFields*nD)*node + nD*field + d]
//   LbField      data[(nFields*nQ)*node + nQ*field + q]
// The host arrays are what chimp_upload_lbfield / chimp_download_* read and write.
#ifndef CHIMP_LBFIELD_H
#define CHIMP_LBFIELD_H

#include <fstream>

#include "LBglobal.h"

namespace chimp_host {
// raw restart files of the reference (LBfield.h:102-138, 233-276, 378-421): int32 header + AoS doubles
inline void writeRaw(const std::string &file, std::initializer_list<int> header, const lbBase_t *data, std::size_t n)
{
    std::ofstream ofs(file, std::ios::out | std::ios::binary);
    if (!ofs) { std::cout << "Could not open file: " + file << std::endl; return; }
    for (int h : header) ofs.write((const char *)&h, sizeof(int));
    ofs.write((const char *)data, std::streamsize(n * sizeof(lbBase_t)));
}
inline bool readRaw(const std::string &file, std::initializer_list<int> expect, lbBase_t *data, std::size_t n)
{
    std::ifstream ifs(file, std::ios::in | std::ios::binary);
    if (!ifs) { std::cout << "Could not open file: " + file << std::endl; return false; }
    for (int e : expect) {
        int h = 0;
        ifs.read((char *)&h, sizeof(int));
        if (h != e) { std::cout << "WARNNING: Mismatch between field size and read field size in file: " << file << "\n          No data read!" << std::endl; return false; }
    }
    ifs.read((char *)data, std::streamsize(n * sizeof(lbBase_t)));
    return bool(ifs);
}
} // namespace chimp_host

class ScalarField
{
public:
    ScalarField(int nFields, int nNodes) : nFields_(nFields), nNodes_(nNodes), data_(0.0, std::size_t(nFields) * nNodes) {}
    lbBase_t &operator()(int fieldNo, int nodeNo) { return data_[std::size_t(nFields_) * nodeNo + fieldNo]; }
    const lbBase_t &operator()(int fieldNo, int nodeNo) const { return data_[std::size_t(nFields_) * nodeNo + fieldNo]; }
    int num_fields() const { return nFields_; }
    int getNumNodes() const { return nNodes_; }
    int size() const { return nNodes_; }
    lbBase_t *data() { return &data_[0]; }
    const lbBase_t *data() const { return &data_[0]; }
    int index(int fieldNo, int /*dimNo*/, int nodeNo) const { return nFields_ * nodeNo + fieldNo; } // LBfield.h:69-70
    constexpr int dim() const { return 1; }
    void writeToFile(const std::string &fileName) const { chimp_host::writeRaw(fileName + ".lbsca", {nFields_, nNodes_}, &data_[0], data_.size()); }
    void readFromFile(const std::string &fileName) { chimp_host::readRaw(fileName + ".lbsca", {nFields_, nNodes_}, &data_[0], data_.size()); }

private:
    int nFields_, nNodes_;
    std::valarray<lbBase_t> data_;
};

template <typename DXQY>
class VectorField
{
public:
    VectorField(int nFields, int nNodes) : nFields_(nFields), nNodes_(nNodes), data_(0.0, std::size_t(nFields) * DXQY::nD * nNodes) {}
    VectorField(int nFields, int nNodes, const std::vector<lbBase_t> &init) : VectorField(nFields, nNodes)
    {
        for (int n = 0; n < nNodes; ++n)
            for (int f = 0; f < nFields; ++f)
                for (int d = 0; d < DXQY::nD; ++d) (*this)(f, d, n) = init[d];
    }
    lbBase_t &operator()(int fieldNo, int dim, int nodeNo) { return data_[std::size_t(nFields_ * DXQY::nD) * nodeNo + DXQY::nD * fieldNo + dim]; }
    const lbBase_t &operator()(int fieldNo, int dim, int nodeNo) const { return data_[std::size_t(nFields_ * DXQY::nD) * nodeNo + DXQY::nD * fieldNo + dim]; }
    std::valarray<lbBase_t> operator()(int fieldNo, int nodeNo) const
    {
        return data_[std::slice(std::size_t(nFields_ * DXQY::nD) * nodeNo + DXQY::nD * fieldNo, DXQY::nD, 1)];
    }
    std::slice_array<lbBase_t> set(int fieldNo, int nodeNo)
    {
        return data_[std::slice(std::size_t(nFields_ * DXQY::nD) * nodeNo + DXQY::nD * fieldNo, DXQY::nD, 1)];
    }
    int num_fields() const { return nFields_; }
    int getNumNodes() const { return nNodes_; }
    lbBase_t *data() { return &data_[0]; }
    const lbBase_t *data() const { return &data_[0]; }
    int index(int fieldNo, int dimNo, int nodeNo) const { return nFields_ * DXQY::nD * nodeNo + DXQY::nD * fieldNo + dimNo; } // LBfield.h:219-220
    constexpr int dim() const { return DXQY::nD; }
    void writeToFile(const std::string &fileName) const { chimp_host::writeRaw(fileName + ".lbvec", {nFields_, DXQY::nD, nNodes_}, &data_[0], data_.size()); }
    void readFromFile(const std::string &fileName) { chimp_host::readRaw(fileName + ".lbvec", {nFields_, DXQY::nD, nNodes_}, &data_[0], data_.size()); }

private:
    int nFields_, nNodes_;
    std::valarray<lbBase_t> data_;
};

template <typename DXQY>
class LbField
{
public:
    LbField(int nFields, int nNodes) : nFields_(nFields), nNodes_(nNodes), data_(0.0, std::size_t(nFields) * DXQY::nQ * nNodes) {}
    lbBase_t &operator()(int fieldNo, int q, int nodeNo) { return data_[std::size_t(nFields_ * DXQY::nQ) * nodeNo + DXQY::nQ * fieldNo + q]; }
    const lbBase_t &operator()(int fieldNo, int q, int nodeNo) const { return data_[std::size_t(nFields_ * DXQY::nQ) * nodeNo + DXQY::nQ * fieldNo + q]; }
    std::valarray<lbBase_t> operator()(int fieldNo, int nodeNo) const
    {
        return data_[std::slice(std::size_t(nFields_ * DXQY::nQ) * nodeNo + DXQY::nQ * fieldNo, DXQY::nQ, 1)];
    }
    std::slice_array<lbBase_t> set(int fieldNo, int nodeNo)
    {
        return data_[std::slice(std::size_t(nFields_ * DXQY::nQ) * nodeNo + DXQY::nQ * fieldNo, DXQY::nQ, 1)];
    }
    void swapData(LbField &other) { data_.swap(other.data_); }
    // push streaming of one node: f(fieldNo, q, neighbor(q, nodeNo)) = fNew[q] for all q, rest direction included
    template <typename GRID, typename T>
    void propagateTo(int fieldNo, int nodeNo, const T &fNew, const GRID &grid)
    {
        for (int q = 0; q < DXQY::nQ; ++q) (*this)(fieldNo, q, grid.neighbor(q, nodeNo)) = fNew[q];
    }
    int num_fields() const { return nFields_; }
    int getNumNodes() const { return nNodes_; }
    lbBase_t *data() { return &data_[0]; }
    const lbBase_t *data() const { return &data_[0]; }
    void writeToFile(const std::string &fileName) const { chimp_host::writeRaw(fileName + ".lblbf", {nFields_, DXQY::nQ, nNodes_}, &data_[0], data_.size()); }
    void readFromFile(const std::string &fileName) { chimp_host::readRaw(fileName + ".lblbf", {nFields_, DXQY::nQ, nNodes_}, &data_[0], data_.size()); }

private:
    int nFields_, nNodes_;
    std::valarray<lbBase_t> data_;
};

#endif
