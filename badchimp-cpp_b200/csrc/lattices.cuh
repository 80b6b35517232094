// Lattice descriptors for the device path (D2Q9, D3Q19, D3Q27).
//
// Same constants and the same direction ordering as the reference's lattice structs
// (reference: src/lbsolver/LBd2q9.h:12-40, LBd3q19.h:12-42; contract LBlatticetypes.h:8-112):
// the first nPairs directions, then their reverses in the same order, rest direction last,
// reverse(q) = (q + nPairs) % (nQ - 1).  D3Q27 does not exist in the reference; it is written
// to the same contract (first 9 directions = D3Q19's, then the 4 cube corners).
//
// Everything is constexpr so that fully unrolled device loops fold the tables away.
// Summation conventions (ascending dimension in cDot, ascending q in the first moment, zero
// terms skipped) are those of the reference's generated code (LBd3q19.h:129-153,182-190) so
// double-precision results are bit-identical to the CPU build when FMA contraction is off.
#pragma once

#define CHIMP_HD __host__ __device__ __forceinline__

namespace chimp {

struct D2Q9 {
    static constexpr int nD = 2, nQ = 9, nPairs = 4;
    static constexpr int id = 0;
    CHIMP_HD static constexpr int c(int q, int d)
    {
        constexpr int t[18] = {1, 0, 1, 1, 0, 1, -1, 1, -1, 0, -1, -1, 0, -1, 1, -1, 0, 0};
        return t[q * 2 + d];
    }
    // weight class: 0 rest, 1 axis, 2 diagonal
    CHIMP_HD static constexpr int wclass(int q) { return q == 8 ? 0 : ((q & 1) ? 2 : 1); }
    CHIMP_HD static constexpr double w(int q)
    {
        return wclass(q) == 0 ? 16.0 / 36.0 : (wclass(q) == 1 ? 4.0 / 36.0 : 1.0 / 36.0);
    }
    // colour-gradient B weights (LBd2q9.h:36-39)
    CHIMP_HD static constexpr double B(int q)
    {
        return wclass(q) == 0 ? -16.0 / 108.0 : (wclass(q) == 1 ? 8.0 / 108.0 : 5.0 / 108.0);
    }
    static constexpr bool hasB = true;
};

struct D3Q19 {
    static constexpr int nD = 3, nQ = 19, nPairs = 9;
    static constexpr int id = 1;
    CHIMP_HD static constexpr int c(int q, int d)
    {
        constexpr int t[57] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 0, 1, -1, 0, 1, 0, 1, 1, 0, -1, 0, 1, 1, 0, 1, -1,
                               -1, 0, 0, 0, -1, 0, 0, 0, -1, -1, -1, 0, -1, 1, 0, -1, 0, -1, -1, 0, 1, 0, -1, -1, 0, -1, 1,
                               0, 0, 0};
        return t[q * 3 + d];
    }
    CHIMP_HD static constexpr int wclass(int q) { return q == 18 ? 0 : ((q % 9) < 3 ? 1 : 2); }
    CHIMP_HD static constexpr double w(int q)
    {
        return wclass(q) == 0 ? 12.0 / 36.0 : (wclass(q) == 1 ? 2.0 / 36.0 : 1.0 / 36.0);
    }
    // LBd3q19.h:37-40
    CHIMP_HD static constexpr double B(int q)
    {
        return wclass(q) == 0 ? -12.0 / 54.0 : (wclass(q) == 1 ? 1.0 / 54.0 : 2.0 / 54.0);
    }
    static constexpr bool hasB = true;
};

struct D3Q27 {
    static constexpr int nD = 3, nQ = 27, nPairs = 13;
    static constexpr int id = 2;
    CHIMP_HD static constexpr int c(int q, int d)
    {
        constexpr int t[81] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 0, 1, -1, 0, 1, 0, 1, 1, 0, -1, 0, 1, 1, 0, 1, -1,
                               1, 1, 1, 1, 1, -1, 1, -1, 1, 1, -1, -1,
                               -1, 0, 0, 0, -1, 0, 0, 0, -1, -1, -1, 0, -1, 1, 0, -1, 0, -1, -1, 0, 1, 0, -1, -1, 0, -1, 1,
                               -1, -1, -1, -1, -1, 1, -1, 1, -1, -1, 1, 1,
                               0, 0, 0};
        return t[q * 3 + d];
    }
    CHIMP_HD static constexpr int wclass(int q) { return q == 26 ? 0 : ((q % 13) < 3 ? 1 : ((q % 13) < 9 ? 2 : 3)); }
    CHIMP_HD static constexpr double w(int q)
    {
        return wclass(q) == 0 ? 64.0 / 216.0
                              : (wclass(q) == 1 ? 16.0 / 216.0 : (wclass(q) == 2 ? 4.0 / 216.0 : 1.0 / 216.0));
    }
    CHIMP_HD static constexpr double B(int) { return 0.0; } // no colour-gradient weights defined
    static constexpr bool hasB = false;
};

template <class L>
CHIMP_HD constexpr int reverseDir(int q)
{
    return q == L::nQ - 1 ? q : (q + L::nPairs) % (L::nQ - 1);
}

// Lattice constants shared by all three (LBd3q19.h:19-23)
constexpr double kC2Inv = 3.0;
constexpr double kC4Inv = 9.0;
constexpr double kC2 = 1.0 / 3.0;
constexpr double kC4Inv0_5 = 0.5 * 9.0;

} // namespace chimp
