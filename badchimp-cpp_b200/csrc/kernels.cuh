// Device kernels of the hot path: fused collide-and-stream over the neighbour-list lattice.
//
// Storage ("pull" form).  Own fluid nodes (the reference's bulkNodes, LBgeometry.h:11-21)
// are renumbered 0..N-1 (device order).  Each LbField is stored SoA as planes
//      X[field][q][slot],  slot in [0, stride)
// holding the POST-COLLISION value f*_q produced by the node in that slot.  Slots
// [nPad, stride) of every plane are halo-in slots filled by the neighbour ranks.  The
// reference's post-stream / post-boundary state f_q(n) (LBfield.h:350-359 push + swap,
// LBmonlatmpi.h:236-297 ghost exchange, LBhalfwaybb.h:37-63 bounce back) is then
//      f_q(n) = X[q][ T[q][n] ]          if T[q][n] >= 0
//             = X[reverse(q)][n]         if T[q][n] == -1     (half-way bounce back, link swap)
// where the pull table T is produced by the host builder (engine.cu) by symbolically
// replaying the reference's push, ghost exchange and boundary copies.  One step reads the
// state through T (gather), collides, and writes the node's own slots of the other buffer
// (fully coalesced).  Two buffers (A/B) alternate.
//
// Index forms: IDX_TABLE reads the int32 table T[q][n]; IDX_RANK is the compressed form
// (per-node bounce-back bitmask + one base per (32-node tile, q); source of lane l is
// base + popc(non-bounce lanes below l); irregular (tile,q) pairs fall back to explicit rows).
#pragma once
#include <cstdint>
#include <utility>
#include "collide.cuh"

namespace chimp {

enum { COLL_BGK = 0, COLL_TRT = 1 };
enum { IDX_TABLE = 0, IDX_RANK = 1 };

struct IndexView {
    const int32_t *table;    // IDX_TABLE: [nQ][nPad]
    const uint32_t *bbmask;  // IDX_RANK : [nPad] bit q set -> f_q(n) = X[rev q][n]; bit 31: node is real
    const int32_t *base;     // IDX_RANK : [nQ][nTiles]; >= 0 regular base, < 0: -(row+1) into `rows`
    const int32_t *rows;     // IDX_RANK : explicit rows [nRows][32]
    int nTiles;
};

struct StepArgs {
    const double *fin;
    double *fout;
    long long stride;   // doubles per (field,q) plane
    int n;              // own nodes
    int nPad;           // padded own nodes (multiple of 32)
    int begin, end;     // node range of this launch (multiples of 32 except end == n)
    IndexView idx;
    // collision parameters
    double tauInv, tauFactor;             // BGK: 1/tau, 1 - 0.5/tau
    double tauSymInv, tauAntiInv, symFactor, antiFactor; // TRT
    double F[3];
    // optional per-node inputs of the one_phase variant (std_one_phase/main.cpp:534-575)
    const double *forceOn;     // [nPad] multiplies F
    const double *addSource;   // [nPad] 0/1
    const int32_t *label;      // [nPad] interior-domain label
    const double *srcPerLabel; // [nLabels] = 0.9*2*scale[label]*massChange[label]
    const uint32_t *pmask;     // [nPad] bit q: X[q][n] carries the anti-bounce-back value (main.cpp:155-174)
    double rhoW;
    // optional outputs
    double *rho; // [nPad]
    double *vel; // [nD][nPad]
};

template <class F, int... Q>
__device__ __forceinline__ void staticForImpl(F &f, std::integer_sequence<int, Q...>)
{
    (f(std::integral_constant<int, Q>{}), ...);
}
// calls f(integral_constant<int,0>) ... f(integral_constant<int,N-1>): compile-time q
template <int N, class F>
__device__ __forceinline__ void staticFor(F &f)
{
    staticForImpl(f, std::make_integer_sequence<int, N>{});
}

template <class L, int IDX>
struct Gather {
    // loads f_q(n) for all q of node i (lane = i & 31)
    __device__ __forceinline__ static void load(const StepArgs &a, const double *__restrict__ fin, int i, bool live,
                                                double (&f)[L::nQ])
    {
        if (IDX == IDX_TABLE) {
            if (!live) {
#pragma unroll
                for (int q = 0; q < L::nQ; ++q) f[q] = 0.0;
                return;
            }
            int src[L::nQ];
#pragma unroll
            for (int q = 0; q < L::nQ; ++q) src[q] = __ldg(a.idx.table + (long long)q * a.nPad + i);
#pragma unroll
            for (int q = 0; q < L::nQ; ++q) {
                const int s = src[q];
                const double *p = (s >= 0) ? fin + (long long)q * a.stride + s
                                           : fin + (long long)reverseDir<L>(q) * a.stride + i;
                f[q] = __ldg(p);
            }
        } else {
            const uint32_t m = live ? __ldg(a.idx.bbmask + i) : 0xffffffffu;
            const int tile = i >> 5;
            const unsigned lane = threadIdx.x & 31u;
            const unsigned below = (1u << lane) - 1u;
#pragma unroll
            for (int q = 0; q < L::nQ; ++q) {
                const int b = tile < a.idx.nTiles ? __ldg(a.idx.base + (long long)q * a.idx.nTiles + tile) : 0;
                const unsigned bb = __ballot_sync(0xffffffffu, (m >> q) & 1u);
                int s;
                if (b >= 0) s = b + __popc(~bb & below);
                else s = __ldg(a.idx.rows + ((long long)(-b - 1) << 5) + lane);
                const bool bounce = (m >> q) & 1u;
                const double *p = bounce ? fin + (long long)reverseDir<L>(q) * a.stride + i
                                         : fin + (long long)q * a.stride + s;
                f[q] = live ? __ldg(p) : 0.0;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
// Single-field collide-and-stream: BGK or TRT, Guo force, optional mass source / force mask
// / anti-bounce-back pressure links (one_phase variant), optional moment output.
// Reference loop bodies: std_case/main.cpp:110-135, std_one_phase/main.cpp:534-575.
// ---------------------------------------------------------------------------------------
template <class L, int COLL, bool ONEPHASE, bool MOM, int IDX>
__global__ void __launch_bounds__(256) collideStreamKernel(const StepArgs a)
{
    const int i = a.begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.end;
    if (IDX == IDX_TABLE && !live) return;
    if (IDX == IDX_RANK && (i & ~31) >= a.end) return; // whole warp out of range

    double f[L::nQ];
    Gather<L, IDX>::load(a, a.fin, i, live, f);
    if (!live) return;

    double rho = nodeRho<L>(f);
    double F[3] = {a.F[0], a.F[1], a.F[2]};
    double qSrc = 0.0;
    if (ONEPHASE) {
        const double on = a.forceOn[i];
        qSrc = a.srcPerLabel[a.label[i]] * a.addSource[i];
        rho += 0.5 * qSrc;
#pragma unroll
        for (int d = 0; d < L::nD; ++d) F[d] = F[d] * on;
    }
    double u[3] = {0.0, 0.0, 0.0};
    u[0] = (firstMoment<L, 0>(f) + 0.5 * F[0]) / rho;
    u[1] = (firstMoment<L, 1>(f) + 0.5 * F[1]) / rho;
    if (L::nD == 3) u[2] = (firstMoment<L, 2>(f) + 0.5 * F[2]) / rho;

    if (MOM) {
        a.rho[i] = rho;
#pragma unroll
        for (int d = 0; d < L::nD; ++d) a.vel[(long long)d * a.nPad + i] = u[d];
    }

    const double u2 = dotD<L>(u, u);
    const double uF = dotD<L>(u, F);
    uint32_t pm = 0;
    if (ONEPHASE && a.pmask) pm = a.pmask[i];

    double out[L::nQ];
    // compile-time q via recursive lambda-free unrolling
    auto body = [&](auto qc) {
        constexpr int q = decltype(qc)::value;
        const double cu = cDot<L, q>(u);
        const double cF = cDot<L, q>(F);
        double om, dF;
        if (COLL == COLL_BGK) {
            om = omegaBGK<L, q>(f[q], a.tauInv, rho, cu, u2);
            dF = deltaOmegaF<L, q>(a.tauFactor, cu, uF, cF);
        } else {
            om = omegaTRT<L, q>(f[q], f[reverseDir<L>(q)], a.tauSymInv, a.tauAntiInv, rho, cu, u2);
            dF = deltaOmegaFTRT<L, q>(a.symFactor, a.antiFactor, 1.0, cu, uF, cF);
        }
        double v = f[q] + om + dF;
        if (ONEPHASE) {
            const double dQ = (COLL == COLL_BGK) ? deltaOmegaQ<L, q>(a.tauFactor, cu, u2, qSrc)
                                                 : deltaOmegaQTRT<L, q>(a.symFactor, a.antiFactor, cu, u2, qSrc);
            v = v + dQ;
            if ((pm >> q) & 1u) {
                // anti bounce back: the slot is only ever pulled back by this node's reverse
                // direction (std_one_phase/main.cpp:168-172, w and cu of the known direction q)
                v = -v + 2 * L::w(q) * a.rhoW * (1 + 0.5 * (kC4Inv * cu * cu - kC2Inv * u2));
            }
        }
        out[q] = v;
    };
    staticFor<L::nQ>(body);

#pragma unroll
    for (int q = 0; q < L::nQ; ++q) a.fout[(long long)q * a.stride + i] = out[q];
}

// ---------------------------------------------------------------------------------------
// Per-label sum of (1 - rho) over the own nodes (std_one_phase/main.cpp:520-528) as a
// fixed-shape tree: warp shuffles, 8 warps per block, one partial per (label, block); the
// second kernel folds the partials in a fixed order, so the result is run-to-run
// deterministic (it differs from the reference's sequential sum only by rounding order).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double blockSum256(double v, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[w];
    }
    return t; // valid in thread 0
}

template <class L, int IDX>
__global__ void __launch_bounds__(256) massChangeKernel(const StepArgs a, int nLabels, double *partial)
{
    __shared__ double sh[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.n;
    double f[L::nQ];
    Gather<L, IDX>::load(a, a.fin, live ? i : 0, live, f);
    const double val = live ? 1.0 - nodeRho<L>(f) : 0.0;
    const int lab = live ? a.label[i] : -1;
    for (int l = 0; l < nLabels; ++l) {
        const double s = blockSum256(lab == l ? val : 0.0, sh);
        if (threadIdx.x == 0) partial[(long long)l * gridDim.x + blockIdx.x] = s;
    }
}

// one block per label: mass[l] = sum of partials; src[l] = 0.9*2*scale[l]*mass[l] (main.cpp:548)
__global__ void massFinalizeKernel(const double *__restrict__ partial, int nBlocks, const double *__restrict__ scale,
                                   double *mass, double *src);

// ---------------------------------------------------------------------------------------
// Layout conversion between the reference AoS field (LBfield.h:300:
// data[(nFields*nQ)*node + nQ*field + q], reference node labels) and the device planes.
// scatterState: X[T-resolved slot] = f_q(label(i)); halo-in slots are filled from the
// rank's own state so that the first step needs no exchange.  gatherState is the inverse.
// ---------------------------------------------------------------------------------------
template <class L>
__global__ void scatterStateKernel(const double *__restrict__ aos, double *X, const int32_t *__restrict__ table,
                                   const int32_t *__restrict__ label, int n, int nPad, long long stride, int nFields)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long row = (long long)label[i] * nFields * L::nQ;
    for (int fl = 0; fl < nFields; ++fl) {
        double *Xf = X + (long long)fl * L::nQ * stride;
#pragma unroll
        for (int q = 0; q < L::nQ; ++q) {
            const double v = aos[row + (long long)fl * L::nQ + q];
            const int s = table[(long long)q * nPad + i];
            if (s >= 0) Xf[(long long)q * stride + s] = v;
            else Xf[(long long)reverseDir<L>(q) * stride + i] = v;
        }
    }
}

template <class L>
__global__ void gatherStateKernel(double *aos, const double *__restrict__ X, const int32_t *__restrict__ table,
                                  const int32_t *__restrict__ label, int n, int nPad, long long stride, int nFields)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long row = (long long)label[i] * nFields * L::nQ;
    for (int fl = 0; fl < nFields; ++fl) {
        const double *Xf = X + (long long)fl * L::nQ * stride;
#pragma unroll
        for (int q = 0; q < L::nQ; ++q) {
            const int s = table[(long long)q * nPad + i];
            const double v = (s >= 0) ? Xf[(long long)q * stride + s] : Xf[(long long)reverseDir<L>(q) * stride + i];
            aos[row + (long long)fl * L::nQ + q] = v;
        }
    }
}

// out[label[i]*nComp + c] = planes[c][i]   (ScalarField / VectorField layout, LBfield.h:94,190)
__global__ void planesToAosKernel(double *aos, const double *__restrict__ planes, const int32_t *__restrict__ label,
                                  int n, int nPad, int nComp, int aosStride, int aosOffset);

// halo pack: buf[k] = X[src[k]] ; unpack: X[dst[k]] = buf[k]   (64-bit slot offsets)
__global__ void haloPackKernel(double *__restrict__ buf, const double *__restrict__ X,
                               const long long *__restrict__ src, int count);
__global__ void haloUnpackKernel(double *X, const double *__restrict__ buf, const long long *__restrict__ dst,
                                 int count);

} // namespace chimp
