// Colour-gradient two-phase step as ONE launch whose moment pass runs a short distance ahead of the
// collide pass, so that the collide pass finds the populations it gathers still in L2.
//
// The reference iteration (twophase/main_TWOPHASE.cpp:236-392) sweeps all nodes three times: rho0/rho1/phi
// (:238-246), the global x-momentum of the flux controller (:292-299) and the collision (:312-376), which
// needs phi of every neighbour.  The two-kernel form of kernels.cuh therefore reads both LbFields from HBM
// twice per step.  Here a launch consists of 2 * nTiles work items of 128 nodes,
//     A(t): gather both fields of tile t, phi = (rho0 - rho1) / (rho0 + rho1)            -> phi[]
//     D(t): gather both fields again, rho0/rho1 from the same sums, grad phi, collide, recolour, stream,
// dispensed in a fixed order (built by the host, chimp_lattice::tpSeq) in which A leads D by the reach of
// the stencil plus a margin of roughly one wave of blocks.  A block takes its item with an atomic ticket,
// so every item with a smaller sequence number is held by a block that is already running: a D item that
// waits for the A items it depends on (per-chunk completion counters) can never starve them.
//
// The global sum of the flux controller would force all A items before any D item.  It is taken out of the
// way by forming it one step earlier, over the populations as they are WRITTEN: node n contributes
// sum_q (+-)c_qx (f*_0q + f*_1q), minus for the links that bounce back into the reversed direction -- exactly
// the terms the destination nodes add up after streaming (:294-297), in another (fixed, deterministic)
// order.  outputMomentumKernel forms the same sum for a state that did not come out of a D pass
// (upload / initialisation), so results do not depend on how a run is split into calls.
#pragma once
#include "kernels.cuh"

namespace chimp {

#define CHIMP_FUSED_TILE 128        // nodes per work item = threads per block
#define CHIMP_FUSED_CHUNK_SHIFT 5   // completion counters cover 32 tiles (4096 nodes)
#ifndef CHIMP_FUSED_MIN_BLOCKS
#define CHIMP_FUSED_MIN_BLOCKS 4
#endif

struct FusedArgs {
    TwoPhaseArgs tp;
    const int32_t *seq;       // [2 * nTiles] item order: tile index, bit 31 set = D
    const int4 *deps;         // [nTiles] chunks a D item waits for: [x, y] and [z, w] (empty when z > w)
    unsigned *done;           // [nChunks] number of A items finished, accumulated over launches
    unsigned *ticket;         // item dispenser of this launch (zeroed before the launch)
    unsigned launchNo;        // 1, 2, ...: counters reach launchNo * tilesInChunk
    int nTiles;
    const double *force;      // device scalar: F_x of this step
    double *momPartial;       // [nTiles] per-tile x-momentum of the state written by this launch
};

__device__ __forceinline__ unsigned loadAcquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// sum over the 4 warps of a 128-thread block; valid in thread 0
__device__ __forceinline__ double blockSum128(double v, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    return (sh[0] + sh[1]) + (sh[2] + sh[3]);
}

// x-momentum carried by the populations node i writes: + c_qx for a link that streams on, - c_qx for one that
// comes back reversed (the pull of direction rev q at node i is a bounce, s[rev q] outside the plane)
template <class L>
__device__ __forceinline__ double writtenMomentumTerm(double acc, int q, double t, bool bounces, bool &first)
{
    const int cx = L::c(q, 0);
    if (cx == 0) return acc;
    const bool plus = (cx > 0) != bounces;
    if (first) { first = false; return plus ? t : -t; }
    return plus ? acc + t : acc - t;
}

template <class L, bool MOM, int IDX>
__global__ void __launch_bounds__(CHIMP_FUSED_TILE, CHIMP_FUSED_MIN_BLOCKS) twoPhaseFusedKernel(const FusedArgs fa)
{
    __shared__ int item;
    __shared__ double sh[4];
    const TwoPhaseArgs &a = fa.tp;
    if (threadIdx.x == 0) item = fa.seq[atomicAdd(fa.ticket, 1u)];
    __syncthreads();
    const int tile = item & 0x7fffffff;
    const bool isD = item < 0;
    const int i = tile * CHIMP_FUSED_TILE + threadIdx.x;
    const bool live = i < a.n;
    const long long field1 = (long long)L::nQ * a.stride;

    if (!isD) {
        // ---- A: rho0, rho1 -> phi (main_TWOPHASE.cpp:238-246)
        if ((i & ~31) < a.n) {
            int s[L::nQ];
            resolveSources<L, IDX>(a, i, live, s);
            double f0[L::nQ], f1[L::nQ];
#pragma unroll
            for (int q = 0; q < L::nQ; ++q) {
                f0[q] = live ? __ldg(a.pl.in[q] + s[q]) : 0.0;
                f1[q] = live ? __ldg(a.pl.in[q] + s[q] + field1) : 0.0;
            }
            if (live) {
                const double r0 = nodeRho<L>(f0), r1 = nodeRho<L>(f1);
                a.phi[i] = (r0 - r1) / (r0 + r1);
            }
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(fa.done + (tile >> CHIMP_FUSED_CHUNK_SHIFT), 1u);
        return;
    }

    // ---- D: wait until phi of every neighbour of this tile is final
    if (threadIdx.x < 32) {
        const int4 d = fa.deps[tile];
        const int lastChunk = (fa.nTiles - 1) >> CHIMP_FUSED_CHUNK_SHIFT;
        const int n1 = d.y - d.x + 1, n2 = d.w >= d.z ? d.w - d.z + 1 : 0;
        for (int k = threadIdx.x; k < n1 + n2; k += 32) {
            const int c = k < n1 ? d.x + k : d.z + (k - n1);
            const unsigned tilesInChunk = c == lastChunk ? (unsigned)(fa.nTiles - (lastChunk << CHIMP_FUSED_CHUNK_SHIFT)) : (1u << CHIMP_FUSED_CHUNK_SHIFT);
            const unsigned want = fa.launchNo * tilesInChunk;
            while (loadAcquire(fa.done + c) < want) __nanosleep(64);
        }
    }
    __syncthreads();
    const bool warpActive = (i & ~31) < a.n; // warps beyond the last node only take part in the tile sum
    double mom = 0.0;
    if (warpActive) {
        // colour gradient (LButilities.h:12-22 -> LBd3q19.h:155-163); phi was written during this launch by other
        // SMs: read it at L2 (ld.global.cg), never through the non-coherent path
        double g[3] = {0.0, 0.0, 0.0};
        double CGNorm = 0.0;
        if (live) {
            double ph[L::nQ];
#pragma unroll
            for (int q = 0; q < L::nQ; ++q) ph[q] = __ldcg(a.phi + __ldcs(a.ptable + ((unsigned)q * (unsigned)a.nPad + (unsigned)i)));
            g[0] = latticeGrad<L, 0>(ph);
            g[1] = latticeGrad<L, 1>(ph);
            if (L::nD == 3) g[2] = latticeGrad<L, 2>(ph);
            CGNorm = sqrt(dotD<L>(g, g));
            const double inv = 1.0 / (CGNorm + (CGNorm < 2.220446049250313e-16 ? 1.0 : 0.0)); // lbBaseEps (LBglobal.h:14)
#pragma unroll
            for (int d = 0; d < L::nD; ++d) g[d] *= inv;
        }
        int s[L::nQ];
        resolveSources<L, IDX>(a, i, live, s);
        // which of the populations written below come back reversed: the pull of the reversed direction bounces
        unsigned bounceMask = 0;
#pragma unroll
        for (int q = 0; q < L::nQ - 1; ++q) {
            constexpr int nz = L::nQ - 1;
            const int r = (q + L::nPairs) % nz;
            if (s[r] < 0 || (long long)s[r] >= a.stride) bounceMask |= 1u << q;
        }
        double fTot[L::nQ];
        double rho0 = 0.0, rho1 = 0.0;
        {
            double f0[L::nQ], f1[L::nQ];
#pragma unroll
            for (int q = 0; q < L::nQ; ++q) {
                f0[q] = live ? __ldcs(a.pl.in[q] + s[q]) : 0.0; // last use of these lines: evict first
                f1[q] = live ? __ldcs(a.pl.in[q] + s[q] + field1) : 0.0;
            }
            rho0 = nodeRho<L>(f0); // the sums pass A formed for phi (calcRho, LBmacroscopic.h:11-20)
            rho1 = nodeRho<L>(f1);
#pragma unroll
            for (int q = 0; q < L::nQ; ++q) fTot[q] = f0[q] + f1[q];
        }
        if (live) {
            const double rho = rho0 + rho1;
            double F[3] = {*fa.force, a.F[1], a.F[2]};
            double u[3] = {0.0, 0.0, 0.0};
            u[0] = (firstMoment<L, 0>(fTot) + 0.5 * F[0]) / rho;
            u[1] = (firstMoment<L, 1>(fTot) + 0.5 * F[1]) / rho;
            if (L::nD == 3) u[2] = (firstMoment<L, 2>(fTot) + 0.5 * F[2]) / rho;
            if (MOM) {
                a.rho[i] = rho0;
                a.rho[(long long)a.nPad + i] = rho1;
#pragma unroll
                for (int d = 0; d < L::nD; ++d) a.vel[(long long)d * a.nPad + i] = u[d];
            }
            // main_TWOPHASE.cpp:340
            const double tau = kC2Inv * rho / (rho0 * a.nu0Inv + rho1 * a.nu1Inv) + 0.5;
            const double tauInv = 1.0 / tau, tauFactor = (1 - 0.5 / tau);
            const double uu = dotD<L>(u, u);
            const double uF = dotD<L>(u, F);
            const double AF0_5 = 1.125 * CGNorm * a.sigma / tau;      // LBcollision2phase.h:12
            const double rhoFacBeta = a.beta * rho0 * rho1 / rho;     // LBcollision2phase.h:74
            const double c0 = (rho0 / rho), c1 = (rho1 / rho);
            const double c2uu = kC2 * uu, c2uF = kC2 * uF;
            const double negTauInv = -tauInv;
            bool firstTerm = true;
            auto pairBody = [&](auto pc) {
                constexpr int q = decltype(pc)::value, r = q + L::nPairs;
                const double cu = cDot<L, q>(u);
                const double cF = cDot<L, q>(F);
                const double cCG = cDot<L, q>(g);
                const double t = kC4Inv0_5 * (cu * cu - c2uu);
                const double m3 = kC2Inv * cu;
                const double rw = rho * L::w(q);
                const double eq = 1.0 + m3 + t, er = 1.0 - m3 + t;
                const double wtf = L::w(q) * tauFactor;
                const double gF = kC4Inv * (cF * cu - c2uF);
                const double h = kC2Inv * cF;
                const double st = AF0_5 * (L::w(q) * cCG * cCG - L::B(q));     // LBcollision2phase.h:16
                const double rc = rhoFacBeta * L::w(q) * cCG / cNorm<L, q>();  // LBcollision2phase.h:79
                const double commonQ = fTot[q] + negTauInv * (fTot[q] - rw * eq) + wtf * (h + gF) + st;
                const double commonR = fTot[r] + negTauInv * (fTot[r] - rw * er) + wtf * (gF - h) + st;
                const double q0 = c0 * commonQ + rc, q1 = c1 * commonQ - rc;
                const double r0 = c0 * commonR - rc, r1 = c1 * commonR + rc;
                __stcs(a.pl.out[q] + i, q0); // written once, read next step: must not displace the window in L2
                __stcs(a.pl.out[q] + field1 + i, q1);
                __stcs(a.pl.out[r] + i, r0);
                __stcs(a.pl.out[r] + field1 + i, r1);
                mom = writtenMomentumTerm<L>(mom, q, q0 + q1, (bounceMask >> q) & 1u, firstTerm);
                mom = writtenMomentumTerm<L>(mom, r, r0 + r1, (bounceMask >> r) & 1u, firstTerm);
            };
            staticFor<L::nPairs>(pairBody);
            {
                constexpr int q = L::nQ - 1; // rest direction (LBcollision2phase.h:18,84)
                const double om = omegaBGK<L, q>(fTot[q], tauInv, rho, 0.0, uu);
                const double dF = deltaOmegaF<L, q>(tauFactor, 0.0, uF, 0.0);
                const double st = -AF0_5 * L::B(q);
                const double common = fTot[q] + om + dF + st;
                __stcs(a.pl.out[q] + i, c0 * common + 0.0);
                __stcs(a.pl.out[q] + field1 + i, c1 * common - 0.0);
            }
        }
    }
    const double tileMom = blockSum128(mom, sh);
    if (threadIdx.x == 0) fa.momPartial[tile] = tileMom;
}

// the same per-tile sums for a state that was not written by a D pass: X holds f*_q of node i in its own slot
template <class L, int IDX>
__global__ void __launch_bounds__(CHIMP_FUSED_TILE) outputMomentumKernel(const TwoPhaseArgs a, double *momPartial)
{
    __shared__ double sh[4];
    const int tile = blockIdx.x;
    const int i = tile * CHIMP_FUSED_TILE + threadIdx.x;
    const bool live = i < a.n;
    const long long field1 = (long long)L::nQ * a.stride;
    double mom = 0.0;
    if ((i & ~31) < a.n) {
        int s[L::nQ];
        resolveSources<L, IDX>(a, i, live, s);
        if (live) {
            bool firstTerm = true;
            // same order of terms as the D pass: pairs (q, q + nPairs) for q = 0 .. nPairs-1
#pragma unroll
            for (int p = 0; p < L::nPairs; ++p) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int q = p + k * L::nPairs;
                    const int r = (q + L::nPairs) % (L::nQ - 1);
                    if (L::c(q, 0) == 0) continue;
                    const bool bounces = s[r] < 0 || (long long)s[r] >= a.stride;
                    const double t = a.pl.in[q][i] + a.pl.in[q][field1 + i];
                    mom = writtenMomentumTerm<L>(mom, q, t, bounces, firstTerm);
                }
            }
        }
    }
    const double tileMom = blockSum128(mom, sh);
    if (threadIdx.x == 0) momPartial[tile] = tileMom;
}

// Per tile: the range of phi slots of OWN nodes its nodes reference, split into the part near the tile
// (|slot - node| <= window) and the rest (periodic images at the other end of the numbering).
// out[tile] = {nearLo, nearHi, farLo, farHi}, far empty when farLo > farHi.
__global__ void tilePhiRangesKernel(const int32_t *__restrict__ ptable, int n, int nPad, int nQ, int window, int4 *out);

} // namespace chimp
