// ScalarField / VectorField / LbField with the reference's storage layout and accessors
// (src/lbsolver/LBfield.h:48-100, 149-231, 285-376): AoS, node-major,
//   ScalarField  data[nFields*node + field]
//   VectorField  data[(nFields*nD)*node + nD*field + d]
//   LbField      data[(nFields*nQ)*node + nQ*field + q]
// The host arrays are what chimp_upload_lbfield / chimp_download_* read and write.
#ifndef CHIMP_LBFIELD_H
#define CHIMP_LBFIELD_H

#include "LBglobal.h"

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
    int num_fields() const { return nFields_; }
    int getNumNodes() const { return nNodes_; }
    lbBase_t *data() { return &data_[0]; }
    const lbBase_t *data() const { return &data_[0]; }

private:
    int nFields_, nNodes_;
    std::valarray<lbBase_t> data_;
};

#endif
