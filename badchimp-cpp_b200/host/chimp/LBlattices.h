// Lattice descriptors D2Q9 / D3Q19 / D3Q27 with the member names of the reference structs
// (src/lbsolver/LBd2q9.h:12-107, LBd3q19.h:12-108; ordering contract LBlatticetypes.h:8-112):
// first nDirPairs_ directions, then their reverses in the same order, rest direction last.
// One generic implementation serves all three; sums follow the reference's conventions
// (ascending dimension in cDotAll, ascending q in qSumC, zero terms skipped; grad grouped by
// weight class), so host-side values equal the reference's bit for bit.  D3Q27 is new (the
// reference has none) and carries no colour-gradient weights.
#ifndef CHIMP_LBLATTICES_H
#define CHIMP_LBLATTICES_H

#include "LBglobal.h"

namespace chimp_host {

template <class Derived, int ND, int NQ>
struct LatticeBase {
    static constexpr int nD = ND;
    static constexpr int nQ = NQ;
    static constexpr int nDirPairs_ = (NQ - 1) / 2;
    static constexpr int nQNonZero_ = NQ - 1;
    static constexpr lbBase_t c2Inv = 3.0;
    static constexpr lbBase_t c4Inv = 9.0;
    static constexpr lbBase_t c2 = 1.0 / c2Inv;
    static constexpr lbBase_t c4 = 1.0 / c4Inv;
    static constexpr lbBase_t c4Inv0_5 = 0.5 * c4Inv;

    static int c(int q, int d) { return Derived::cDMajor_[nD * q + d]; }
    static std::vector<int> c(int q) { return std::vector<int>(Derived::cDMajor_ + nD * q, Derived::cDMajor_ + nD * q + nD); }
    static int reverseDirection(int q) { return q == nQ - 1 ? q : (q + nDirPairs_) % nQNonZero_; }
    static std::valarray<lbBase_t> cValarray(int q)
    {
        std::valarray<lbBase_t> v(nD);
        for (int d = 0; d < nD; ++d) v[d] = c(q, d);
        return v;
    }
    template <class T1, class T2>
    static lbBase_t dot(const T1 &a, const T2 &b)
    {
        lbBase_t s = a[0] * b[0] + a[1] * b[1];
        if (nD == 3) s = s + a[2] * b[2];
        return s;
    }
    template <class T>
    static lbBase_t cDotRef(int q, const T &v)
    {
        lbBase_t s = c(q, 0) * v[0] + c(q, 1) * v[1];
        if (nD == 3) s = s + c(q, 2) * v[2];
        return s;
    }
    template <class T>
    static std::valarray<lbBase_t> cDotAll(const T &v)
    {
        std::valarray<lbBase_t> ret(nQ);
        for (int q = 0; q < nQ; ++q) {
            lbBase_t s = 0.0;
            bool first = true;
            for (int d = 0; d < nD; ++d) {
                const int cq = c(q, d);
                if (!cq) continue;
                if (first) { s = cq > 0 ? +v[d] : -v[d]; first = false; }
                else s = cq > 0 ? s + v[d] : s - v[d];
            }
            ret[q] = s;
        }
        return ret;
    }
    template <class T>
    static lbBase_t qSum(const T &f)
    {
        lbBase_t r = 0.0;
        for (int q = 0; q < nQ; ++q) r += f[q];
        return r;
    }
    template <class T>
    static std::valarray<lbBase_t> qSumC(const T &f)
    {
        std::valarray<lbBase_t> ret(nD);
        for (int d = 0; d < nD; ++d) {
            lbBase_t s = 0.0;
            bool first = true;
            for (int q = 0; q < nQ; ++q) {
                const int cq = c(q, d);
                if (!cq) continue;
                if (first) { s = cq > 0 ? +f[q] : -f[q]; first = false; }
                else s = cq > 0 ? s + f[q] : s - f[q];
            }
            ret[d] = s;
        }
        return ret;
    }
    // lattice gradient, terms grouped by weight class (LBd3q19.h:155-163)
    template <class T>
    static std::valarray<lbBase_t> grad(const T &rho)
    {
        std::valarray<lbBase_t> ret(nD);
        for (int d = 0; d < nD; ++d) {
            lbBase_t g[2] = {0.0, 0.0};
            bool first[2] = {true, true};
            for (int q = 0; q < nQNonZero_; ++q) {
                const int cq = c(q, d);
                if (!cq) continue;
                const int k = Derived::w[q] == Derived::w1 ? 0 : 1;
                if (first[k]) { g[k] = cq > 0 ? +rho[q] : -rho[q]; first[k] = false; }
                else g[k] = cq > 0 ? g[k] + rho[q] : g[k] - rho[q];
            }
            ret[d] = Derived::w1 * c2Inv * g[0] + Derived::w2 * c2Inv * g[1];
        }
        return ret;
    }
    static int c2q(const std::vector<int> &v)
    {
        for (int q = 0; q < nQ; ++q)
            if (c(q) == v) return q;
        return -1;
    }
};

} // namespace chimp_host

struct D2Q9 : chimp_host::LatticeBase<D2Q9, 2, 9> {
    static constexpr lbBase_t w0 = 16.0 / 36.0, w1 = 4.0 / 36.0, w2 = 1.0 / 36.0;
    static constexpr lbBase_t w[9] = {w1, w2, w1, w2, w1, w2, w1, w2, w0};
    static constexpr int cDMajor_[18] = {1, 0, 1, 1, 0, 1, -1, 1, -1, 0, -1, -1, 0, -1, 1, -1, 0, 0};
    static constexpr lbBase_t B0 = -16.0 / 108.0, B1 = 8.0 / 108.0, B2 = 5.0 / 108.0;
    static constexpr lbBase_t B[9] = {B1, B2, B1, B2, B1, B2, B1, B2, B0};
    static constexpr int chimpId = 0;
};

struct D3Q19 : chimp_host::LatticeBase<D3Q19, 3, 19> {
    static constexpr lbBase_t w0 = 12.0 / 36.0, w1 = 2.0 / 36.0, w2 = 1.0 / 36.0;
    static constexpr lbBase_t w[19] = {w1, w1, w1, w2, w2, w2, w2, w2, w2, w1, w1, w1, w2, w2, w2, w2, w2, w2, w0};
    static constexpr int cDMajor_[57] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 0, 1, -1, 0, 1, 0, 1, 1, 0, -1, 0, 1, 1, 0, 1, -1,
                                         -1, 0, 0, 0, -1, 0, 0, 0, -1, -1, -1, 0, -1, 1, 0, -1, 0, -1, -1, 0, 1, 0, -1, -1, 0, -1, 1,
                                         0, 0, 0};
    static constexpr lbBase_t B0 = -12.0 / 54.0, B1 = 1.0 / 54.0, B2 = 2.0 / 54.0;
    static constexpr lbBase_t B[19] = {B1, B1, B1, B2, B2, B2, B2, B2, B2, B1, B1, B1, B2, B2, B2, B2, B2, B2, B0};
    static constexpr int chimpId = 1;
};

struct D3Q27 : chimp_host::LatticeBase<D3Q27, 3, 27> {
    static constexpr lbBase_t w0 = 64.0 / 216.0, w1 = 16.0 / 216.0, w2 = 4.0 / 216.0, w3 = 1.0 / 216.0;
    static constexpr lbBase_t w[27] = {w1, w1, w1, w2, w2, w2, w2, w2, w2, w3, w3, w3, w3,
                                       w1, w1, w1, w2, w2, w2, w2, w2, w2, w3, w3, w3, w3, w0};
    static constexpr int cDMajor_[81] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 0, 1, -1, 0, 1, 0, 1, 1, 0, -1, 0, 1, 1, 0, 1, -1,
                                         1, 1, 1, 1, 1, -1, 1, -1, 1, 1, -1, -1,
                                         -1, 0, 0, 0, -1, 0, 0, 0, -1, -1, -1, 0, -1, 1, 0, -1, 0, -1, -1, 0, 1, 0, -1, -1, 0, -1, 1,
                                         -1, -1, -1, -1, -1, 1, -1, 1, -1, -1, 1, 1,
                                         0, 0, 0};
    static constexpr int chimpId = 2;
};

#endif
