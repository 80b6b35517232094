// Host-side mirror of the reference's API surface for the hot path (src/LBSOLVER.h), written
// from scratch around the C-ABI engine.  Names, argument meaning and storage layouts follow the
// reference so that its mains keep their structure; the per-node loop runs on the GPU.
// reference: src/lbsolver/LBglobal.h:10-14
#ifndef CHIMP_LBGLOBAL_H
#define CHIMP_LBGLOBAL_H

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <numeric>
#include <string>
#include <valarray>
#include <vector>

typedef double lbBase_t;
constexpr lbBase_t lbBaseEps = std::numeric_limits<lbBase_t>::epsilon();

namespace chimp_host {
[[noreturn]] inline void die(const std::string &msg)
{
    // reference convention: print and exit(1) (e.g. LBvtk.h:230-233)
    std::cout << "ERROR: " << msg << std::endl;
    std::exit(1);
}
} // namespace chimp_host

#endif
