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
// Index forms: IDX_TABLE reads the int32 table T[q][n]; IDX_COMPACT is the compressed form:
// one int32 base per (32-node tile, q) plus one byte per (node, q), packed four directions to
// a 32-bit word: source = base + byte, byte 255 = reversed own slot (bounce back).  vtklb
// numbers fluid nodes consecutively along z, so the sources of a tile span a short label
// range; the few (tile, q) pairs whose span exceeds 253 fall back to explicit 32-entry rows.
#pragma once
#include <cstdint>
#include <utility>
#include "collide.cuh"

namespace chimp {

// launch shape of the collide-stream kernels (tuned on B200)
#ifndef CHIMP_BLOCK
#define CHIMP_BLOCK 128
#endif
#ifndef CHIMP_MIN_BLOCKS
#define CHIMP_MIN_BLOCKS 6
#endif

// two-phase kernels: collide (pass D) and moments (passes A + C; its block reduction assumes 256 threads)
#ifndef CHIMP_TP_BLOCK
#define CHIMP_TP_BLOCK 128
#endif
#ifndef CHIMP_TP_MIN_BLOCKS
#define CHIMP_TP_MIN_BLOCKS 4
#endif
#ifndef CHIMP_PM_MIN_BLOCKS
#define CHIMP_PM_MIN_BLOCKS 3
#endif

enum { COLL_BGK = 0, COLL_TRT = 1 };
// one_phase variant of the single-field kernel: per-node attributes from four arrays (24 B/node) or from one packed word
enum { OP_NONE = 0, OP_ARRAYS = 1, OP_PACKED = 2 };
// IDX_COMPACT_MASK reads the tables of IDX_COMPACT; the spare base slot of a tile additionally says which of its delta
// words are plain runs (byte == lane for every live lane: the sources of those four directions are 32 consecutive slots),
// and such words are not read.  Only the single-field step kernel has this form (chimp_set_index_skip_mask).
enum { IDX_TABLE = 0, IDX_COMPACT = 1, IDX_COMPACT_MASK = 2 };

// The step kernels address one plane as  in[q] + s  with a signed 32-bit slot index s.  A bounce
// (reversed own slot X[rev q][i]) is expressed in the same form: s = i + (rev q - q) * stride,
// so the hot path has no select between two planes.  "Kernel form" below means T with its -1
// entries replaced that way.
struct IndexView {
    const int32_t *table;    // IDX_TABLE  : [nQ][nPad] kernel form
    const uint32_t *delta;   // IDX_COMPACT: [nWords][nPad], byte (q & 3) of word (q >> 2): source - base, 255 = bounce
    const int32_t *base;     // IDX_COMPACT: [nTiles][4*nWords] (16-byte aligned per tile); >= 0 smallest source of
                             //              the tile for direction q, < 0: -(row+1) into `rows`
    const int32_t *rows;     // IDX_COMPACT: explicit rows [nRows][32], kernel form
    int nTiles;
    int bounceOff[27];       // (rev q - q) * stride
};

struct Planes {
    const double *in[27];    // field 0, plane q of the buffer being read
    double *out[27];         // field 0, plane q of the buffer being written
};

// Peer halos fused into the step (the MPI ghost exchange of LBmonlatmpi.h:236-297 over NVLink): the first `blocks`
// thread blocks of a launch hold the halo-coupled nodes.  They wait in their prologue until every face's arrival
// counter says the neighbours' stores of the previous step have landed (and that the neighbours are done reading the
// slots written now), store each outgoing population straight into the neighbour GPU's halo-in slot next to the local
// store, and the last of them to finish publishes the step number in the neighbours' counters.
constexpr int kMaxFaces = 8;
struct PeerView {
    int blocks;                          // 0: plain launch
    int nFaces;
    int pad;                             // row length of `dst`
    const uint32_t *mask;                // [pad] bit q: X[q][i] also goes to a neighbour
    const int32_t *dst;                  // [nQ][pad] face << 28 | slot inside the neighbour's plane q
    const int32_t *extraStart;           // optional [pad + 1]: further destinations of node i are extra[extraStart[i] .. extraStart[i+1])
    const int2 *extra;                   //   {q, face << 28 | slot}: the reference's exchange lists name a node once per ghost
                                         //   image the receiver holds of it (periodic rims, LBbndmpi.h:207-315)
    double *out[kMaxFaces];              // the neighbours' buffers being written (field 0, plane 0)
    long long stride[kMaxFaces];         // their plane strides
    unsigned long long *flagOut[kMaxFaces]; // their arrival counters for my faces
    const unsigned long long *flagIn;    // my arrival counters, one per face
    unsigned long long expect, publish;  // step numbers: wait for >= expect, publish `publish`
    unsigned *counter;                   // blocks finished (reset by the last one)
    unsigned *error;                     // host-mapped error word (arrival timeout)
    unsigned long long timeoutNs;
    unsigned long long *trace;           // optional [4]: first block start, publish time, longest wait, last block end (ns)
};

__device__ __forceinline__ unsigned long long globalTimerNs()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// bounded spin on an arrival counter; a peer that never arrives raises the error word instead of hanging the job
__device__ __forceinline__ unsigned long long awaitCounter(const unsigned long long *flag, unsigned long long expect,
                                                           unsigned long long timeoutNs, unsigned *error, unsigned code)
{
    const volatile unsigned long long *f = flag;
    if (*f >= expect) return 0ull;
    // a wait of this context has already given up: the run is lost, do not spend another time-out per step
    if (error && *(volatile unsigned *)error != 0u) return 0ull;
    const unsigned long long t0 = globalTimerNs();
    unsigned long long waited = 0ull;
    while (*f < expect) {
        __nanosleep(100);
        waited = globalTimerNs() - t0;
        if (waited > timeoutNs) {
            if (error) { *(volatile unsigned *)error = code; __threadfence_system(); }
            break;
        }
    }
    return waited;
}

struct StepArgs {
    Planes pl;
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
    const uint32_t *attr;      // [nPad] ONEPHASE == OP_PACKED: the four arrays above in one word per node --
                               //   bit 0 forceOn, bit 1 addSource (both exactly 0.0 / 1.0), bits 2-5 label, bits 6.. pmask
    double rhoW;
    // optional outputs
    double *rho; // [nPad]
    double *vel; // [nD][nPad]
    PeerView peer;
};

// runtime-q weight lookup for the rare boundary branch (folds when q is a constant)
template <class L>
__device__ __forceinline__ double chimp_w(int q)
{
    return L::w(q);
}

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

// Resolves, for node i and every direction q, the slot s[q] with f_q(i) = in[q][s[q]] (kernel form: a bounce is
// the out-of-plane slot i + bounceOff[q]).  `a` is any argument block with nPad / idx members.  Dead lanes
// (live == false) get slot 0 and must not dereference it.
template <class L, int IDX, class Args>
__device__ __forceinline__ void resolveSources(const Args &a, int i, bool live, int (&s)[L::nQ])
{
    if (IDX == IDX_TABLE) {
#pragma unroll
        for (int q = 0; q < L::nQ; ++q) s[q] = live ? __ldg(a.idx.table + ((unsigned)q * (unsigned)a.nPad + (unsigned)i)) : 0;
    } else if (IDX == IDX_COMPACT_MASK) {
        constexpr int NW = (L::nQ + 3) / 4;
        const int tile = min(i >> 5, a.idx.nTiles - 1);
        // the tile's bases first: their spare slot (never a direction: 4 NW > nQ) holds the skip mask of the delta words
        const int4 *bp = reinterpret_cast<const int4 *>(a.idx.base) + (unsigned)tile * NW;
        int base[NW * 4];
#pragma unroll
        for (int g = 0; g < NW; ++g) {
            const int4 v = __ldg(bp + g);
            base[4 * g] = v.x; base[4 * g + 1] = v.y; base[4 * g + 2] = v.z; base[4 * g + 3] = v.w;
        }
        const uint32_t skip = (uint32_t)base[NW * 4 - 1];
        const uint32_t run = (threadIdx.x & 31u) * 0x01010101u; // byte == lane in all four directions of a word
        uint32_t wd[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w)
            wd[w] = ((skip >> w) & 1u) ? run : (live ? __ldg(a.idx.delta + ((unsigned)w * (unsigned)a.nPad + (unsigned)i)) : 0u);
        int anyRow = 0;
#pragma unroll
        for (int q = 0; q < L::nQ; ++q) {
            const int d = (int)((wd[q >> 2] >> (8 * (q & 3))) & 0xffu);
            s[q] = (d == 255) ? i + a.idx.bounceOff[q] : base[q] + d;
            anyRow |= base[q];
        }
        if (anyRow < 0) { // warp-uniform and rare: some (tile, q) of this tile keeps an explicit row
            const unsigned lane = threadIdx.x & 31u;
#pragma unroll
            for (int q = 0; q < L::nQ; ++q)
                if (base[q] < 0) s[q] = __ldg(a.idx.rows + (((unsigned)(-base[q] - 1) << 5) + lane));
        }
#pragma unroll
        for (int q = 0; q < L::nQ; ++q) s[q] = live ? s[q] : 0;
    } else {
        constexpr int NW = (L::nQ + 3) / 4;
        const int tile = min(i >> 5, a.idx.nTiles - 1);
        // every index word of the node and of its tile is requested before anything waits on one:
        // NW coalesced words of delta bytes, and the tile's bases as NW broadcast 16-byte loads
        uint32_t wd[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) wd[w] = live ? __ldg(a.idx.delta + ((unsigned)w * (unsigned)a.nPad + (unsigned)i)) : 0u;
        const int4 *bp = reinterpret_cast<const int4 *>(a.idx.base) + (unsigned)tile * NW;
        int base[NW * 4];
#pragma unroll
        for (int g = 0; g < NW; ++g) {
            const int4 v = __ldg(bp + g);
            base[4 * g] = v.x; base[4 * g + 1] = v.y; base[4 * g + 2] = v.z; base[4 * g + 3] = v.w;
        }
        int anyRow = 0;
#pragma unroll
        for (int q = 0; q < L::nQ; ++q) {
            const int d = (int)((wd[q >> 2] >> (8 * (q & 3))) & 0xffu);
            s[q] = (d == 255) ? i + a.idx.bounceOff[q] : base[q] + d;
            anyRow |= base[q];
        }
        if (anyRow < 0) { // warp-uniform and rare: some (tile, q) of this tile keeps an explicit row
            const unsigned lane = threadIdx.x & 31u;
#pragma unroll
            for (int q = 0; q < L::nQ; ++q)
                if (base[q] < 0) s[q] = __ldg(a.idx.rows + (((unsigned)(-base[q] - 1) << 5) + lane));
        }
#pragma unroll
        for (int q = 0; q < L::nQ; ++q) s[q] = live ? s[q] : 0;
    }
}

// hands the address of f_q(i) in field 0 of the buffer being read to fn(q, pointer)
template <class L, int IDX, class Args, class Fn>
__device__ __forceinline__ void forEachSource(const Args &a, int i, bool live, Fn &&fn)
{
    int s[L::nQ];
    resolveSources<L, IDX>(a, i, live, s);
#pragma unroll
    for (int q = 0; q < L::nQ; ++q) fn(q, a.pl.in[q] + s[q]);
}

template <class L, int IDX>
struct Gather {
    // loads f_q(n) for all q of node i (lane = i & 31); fieldOff = offset of the field in doubles
    template <class Args>
    __device__ __forceinline__ static void load(const Args &a, long long fieldOff, int i, bool live, double (&f)[L::nQ])
    {
        forEachSource<L, IDX>(a, i, live, [&](int q, const double *p) { f[q] = live ? __ldg(p + fieldOff) : 0.0; });
    }
};

// ---------------------------------------------------------------------------------------
// Single-field collide-and-stream: BGK or TRT, Guo force, optional mass source / force mask
// / anti-bounce-back pressure links (one_phase variant), optional moment output.
// Reference loop bodies: std_case/main.cpp:110-135, std_one_phase/main.cpp:534-575.
// ---------------------------------------------------------------------------------------
template <class L, int COLL, int ONEPHASE, bool MOM, int IDX, bool PEER = false>
__global__ void __launch_bounds__(CHIMP_BLOCK, CHIMP_MIN_BLOCKS) collideStreamKernel(const StepArgs a)
{
    const int i = a.begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.end;
    const bool peerBlock = PEER && (int)blockIdx.x < a.peer.blocks;
    if (!peerBlock) {
        if (IDX == IDX_TABLE && !live) return;
        if (IDX != IDX_TABLE && (i & ~31) >= a.end) return; // whole warp out of range
    }
    uint32_t sendMask = 0;
    if (PEER && peerBlock) {
        if (a.peer.trace && blockIdx.x == 0 && threadIdx.x == 0) a.peer.trace[0] = globalTimerNs();
        if ((int)threadIdx.x < a.peer.nFaces) {
            const unsigned long long w = awaitCounter(a.peer.flagIn + threadIdx.x, a.peer.expect, a.peer.timeoutNs, a.peer.error, 1u);
            if (a.peer.trace && w) atomicMax(a.peer.trace + 2, w);
            __threadfence_system(); // acquire: the neighbour's stores that preceded its counter are visible to what follows
        }
        __syncthreads();
        sendMask = live ? __ldg(a.peer.mask + i) : 0u;
    }

    double f[L::nQ];
    if (!PEER || (i & ~31) < a.end) Gather<L, IDX>::load(a, 0ll, i, live, f);
    if (!PEER && !live) return;
    if (!PEER || live) {

    double rho = nodeRho<L>(f);
    double F[3] = {a.F[0], a.F[1], a.F[2]};
    double qSrc = 0.0;
    uint32_t pm = 0;
    if (ONEPHASE) {
        double on;
        if (ONEPHASE == OP_PACKED) {
            // the same products as below with the 0.0 / 1.0 factors rebuilt from their bits
            const uint32_t w = __ldg(a.attr + i);
            on = (w & 1u) ? 1.0 : 0.0;
            qSrc = a.srcPerLabel[(w >> 2) & 15u] * ((w & 2u) ? 1.0 : 0.0);
            pm = w >> 6;
        } else {
            on = a.forceOn[i];
            qSrc = a.srcPerLabel[a.label[i]] * a.addSource[i];
            if (a.pmask) pm = a.pmask[i];
        }
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

    // Opposite directions are collided together.  With c_r = -c_q every intermediate of the
    // reversed direction is the exact IEEE negation (c.u, c.F, 3 c.u, 3 c.F, the TRT odd part) or
    // the identical value ((c.u)^2, (c.F)(c.u)) of the forward one, so sharing them reproduces the
    // reference's per-direction arithmetic (LBcollision.h:43,67-72,95,118,210,233) bit for bit
    // with a third fewer FP64 instructions.
    const double c2u2 = kC2 * u2, c2uF = kC2 * uF;
    const double negTauInv = -a.tauInv, negSymInv = -a.tauSymInv;
    auto finish = [&](int q, double v, double cu) {
        if (ONEPHASE && ((pm >> q) & 1u)) {
            // anti bounce back (std_one_phase/main.cpp:168-172): this slot is only pulled back by
            // the node's own reversed direction, so it can carry the boundary value directly
            v = -v + 2 * chimp_w<L>(q) * a.rhoW * (1 + 0.5 * (kC4Inv * cu * cu - kC2Inv * u2));
        }
        a.pl.out[q][i] = v;
        if (PEER && ((sendMask >> q) & 1u)) {
            const int d = __ldg(a.peer.dst + ((unsigned)q * (unsigned)a.peer.pad + (unsigned)i));
            const int k = d >> 28;
            a.peer.out[k][(long long)q * a.peer.stride[k] + (d & 0x0fffffff)] = v;
        }
    };
    auto pairBody = [&](auto pc) {
        constexpr int q = decltype(pc)::value, r = q + L::nPairs;
        const double cu = cDot<L, q>(u);
        const double cF = cDot<L, q>(F);
        const double t = kC4Inv0_5 * (cu * cu - c2u2);
        const double m3 = kC2Inv * cu;
        const double rw = rho * L::w(q);
        const double g = kC4Inv * (cF * cu - c2uF);
        const double h = kC2Inv * cF;
        double vq, vr;
        if (COLL == COLL_BGK) {
            const double eq = 1.0 + m3 + t, er = 1.0 - m3 + t;
            const double wtf = L::w(q) * a.tauFactor;
            vq = f[q] + negTauInv * (f[q] - rw * eq) + wtf * (h + g);
            vr = f[r] + negTauInv * (f[r] - rw * er) + wtf * (g - h);
            if (ONEPHASE) {
                const double tsw = a.tauFactor * qSrc * L::w(q);
                vq = vq + tsw * eq;
                vr = vr + tsw * er;
            }
        } else {
            const double fSym = 0.5 * (f[q] + f[r]);
            const double fAnti = 0.5 * (f[q] - f[r]);
            const double es = 1.0 + t;
            const double S = negSymInv * (fSym - rw * es);
            const double A = a.tauAntiInv * (fAnti - rw * kC2Inv * cu);
            const double w1 = L::w(q) * 1.0; // phi = 1.0 (call pattern of rans/main.cpp:530)
            const double ah = a.antiFactor * kC2Inv * cF;
            const double sg = a.symFactor * kC4Inv * (cF * cu - c2uF);
            vq = f[q] + (S - A) + w1 * (ah + sg);
            vr = f[r] + (S + A) + w1 * (sg - ah);
            if (ONEPHASE) {
                const double sw = qSrc * L::w(q);
                const double am = a.antiFactor * kC2Inv * cu;
                const double se = a.symFactor * es;
                vq = vq + sw * (am + se);
                vr = vr + sw * (se - am);
            }
        }
        finish(q, vq, cu);
        finish(r, vr, cu);
    };
    staticFor<L::nPairs>(pairBody);
    {
        // rest direction: c = 0
        constexpr int q = L::nQ - 1;
        double om, dF;
        if (COLL == COLL_BGK) {
            om = omegaBGK<L, q>(f[q], a.tauInv, rho, 0.0, u2);
            dF = deltaOmegaF<L, q>(a.tauFactor, 0.0, uF, 0.0);
        } else {
            om = omegaTRT<L, q>(f[q], f[q], a.tauSymInv, a.tauAntiInv, rho, 0.0, u2);
            dF = deltaOmegaFTRT<L, q>(a.symFactor, a.antiFactor, 1.0, 0.0, uF, 0.0);
        }
        double v = f[q] + om + dF;
        if (ONEPHASE) {
            const double dQ = (COLL == COLL_BGK) ? deltaOmegaQ<L, q>(a.tauFactor, 0.0, u2, qSrc)
                                                 : deltaOmegaQTRT<L, q>(a.symFactor, a.antiFactor, 0.0, u2, qSrc);
            v = v + dQ;
        }
        finish(q, v, 0.0);
    }

    if (PEER && peerBlock && a.peer.extraStart) {
        // further copies of populations this thread has just stored (its own stores are visible to it)
        for (int e = __ldg(a.peer.extraStart + i), e1 = __ldg(a.peer.extraStart + i + 1); e < e1; ++e) {
            const int2 x = __ldg(a.peer.extra + e);
            const int k = x.y >> 28;
            a.peer.out[k][(long long)x.x * a.peer.stride[k] + (x.y & 0x0fffffff)] = a.pl.out[x.x][i];
        }
    }
    } // live
    if (PEER && peerBlock) {
        // every thread's remote stores are ordered before the counter: fence, count blocks, the last one publishes
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned done = atomicAdd(a.peer.counter, 1u);
            if (done == (unsigned)a.peer.blocks - 1u) {
                *a.peer.counter = 0u;
                __threadfence_system();
                for (int k = 0; k < a.peer.nFaces; ++k) *(volatile unsigned long long *)a.peer.flagOut[k] = a.peer.publish;
                __threadfence_system();
                if (a.peer.trace) a.peer.trace[1] = globalTimerNs();
            }
        }
    }
    if (PEER && a.peer.trace && threadIdx.x == 0) atomicMax(a.peer.trace + 3, globalTimerNs());
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
    Gather<L, IDX>::load(a, 0ll, live ? i : 0, live, f);
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
// Colour-gradient two-phase path (twophase/main_TWOPHASE.cpp:236-392), two LbFields.
//   phaseMomentsKernel : pass A (:238-246) rho0, rho1, phi for own nodes, and the per-block
//                        partial sums of the x-momentum of pass C (:292-299)
//   fluxForceKernel    : folds the partials (fixed order) -> F_x = 2 (momx - mean) (:301-308)
//   twoPhaseCollideKernel : pass D (:312-376) collide + recolour + stream of both fields
// phi lives in an array of "phi slots": own nodes first, then solid-boundary nodes (constant
// wettability value, :280-284), ghost nodes (filled by the scalar halo exchange, :287) and one
// zero slot for every other neighbour (cgField is zero-initialised there).
// ---------------------------------------------------------------------------------------
struct TwoPhaseArgs {
    Planes pl;
    long long stride;
    int n, nPad, begin, end;
    IndexView idx;
    const int32_t *ptable; // [nQ][nPad] phi slot of neighbor(q, n)  (LButilities.h:12-22)
    // derived form of ptable: the phi slot of neighbor(q, n) is the pull source of direction rev(q) whenever that
    // neighbour is an own fluid node; only the other links (solid, ghost, zero slots) are stored.
    const int32_t *excInfo; // [nPad] 0: node has no stored link, else 1 + offset of its record in exc
    const int32_t *exc;     // record: bit mask of the stored directions, then their phi slots in direction order
    double *phi;
    double *rho; // [2][nPad]
    double *vel; // [nD][nPad]
    double nu0Inv, nu1Inv, sigma, beta;
    double F[3];
    const double *forceX; // device scalar written by fluxForceKernel
    double *partial;
};

template <class L, int IDX>
__global__ void __launch_bounds__(256, CHIMP_PM_MIN_BLOCKS) phaseMomentsKernel(const TwoPhaseArgs a)
{
    __shared__ double sh[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.n;
    double f0[L::nQ], f1[L::nQ];
    const long long field1 = (long long)L::nQ * a.stride;
    forEachSource<L, IDX>(a, live ? i : 0, live, [&](int q, const double *p) {
        f0[q] = live ? __ldg(p) : 0.0;
        f1[q] = live ? __ldg(p + field1) : 0.0;
    });
    double mom = 0.0;
    if (live) {
        const double r0 = nodeRho<L>(f0), r1 = nodeRho<L>(f1);
        a.rho[i] = r0;
        a.rho[(long long)a.nPad + i] = r1;
        a.phi[i] = (r0 - r1) / (r0 + r1);
#pragma unroll
        for (int q = 0; q < L::nQ; ++q) f0[q] = f0[q] + f1[q];
        mom = firstMoment<L, 0>(f0);
    }
    const double s = blockSum256(mom, sh);
    if (threadIdx.x == 0) a.partial[blockIdx.x] = s;
}

__global__ void fluxForceKernel(const double *partial, int nBlocks, double momx, double nGlobal,
                                double *sumOut, double *forceX, int finish);

// Library form of the flux controller (LBglobalforcing.h:8-33): per-block partial sums of
// qSumC(f(fieldNo, n))[cartDir] over the own nodes; fluxForceKernel folds them.
template <class L, int IDX>
__global__ void __launch_bounds__(256) momentumSumKernel(const StepArgs a, long long fieldOff, int cartDir, double *partial)
{
    __shared__ double sh[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.n;
    double f[L::nQ];
    Gather<L, IDX>::load(a, fieldOff, live ? i : 0, live, f);
    double m = 0.0;
    if (live) m = cartDir == 0 ? firstMoment<L, 0>(f) : (cartDir == 1 || L::nD == 2) ? firstMoment<L, 1>(f) : firstMoment<L, 2>(f);
    const double s = blockSum256(m, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Capillary-number controller (LBglobalforcing.h:35-98): per-block partial sums of
//   phi0 * m, phi1 * m, phi0, phi1   with m = qSumC(f(0, n))[cartDir], phi_s = rho_s / (rho_0 + rho_1)
// over the own nodes (the reference takes field 0 of f and the stored ScalarField rho); partial[k][block].
template <class L, int IDX>
__global__ void __launch_bounds__(256) capNumberSumsKernel(const StepArgs a, int cartDir, const double *__restrict__ rho2, double *partial)
{
    __shared__ double sh[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.n;
    double f[L::nQ];
    Gather<L, IDX>::load(a, 0ll, live ? i : 0, live, f);
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (live) {
        const double m = cartDir == 0 ? firstMoment<L, 0>(f) : (cartDir == 1 || L::nD == 2) ? firstMoment<L, 1>(f) : firstMoment<L, 2>(f);
        const double rho0 = rho2[i], rho1 = rho2[(long long)a.nPad + i];
        const double rhoTot = rho0 + rho1;
        const double phi0 = rho0 / rhoTot, phi1 = rho1 / rhoTot;
        v[0] = phi0 * m;
        v[1] = phi1 * m;
        v[2] = phi0;
        v[3] = phi1;
    }
    for (int k = 0; k < 4; ++k) {
        const double s = blockSum256(v[k], sh);
        if (threadIdx.x == 0) partial[(long long)k * gridDim.x + blockIdx.x] = s;
    }
}
// out[k] = sum of partial[k][0 .. nBlocks) in a fixed order; one 256-thread block per k
__global__ void foldRowsKernel(const double *__restrict__ partial, int nBlocks, double *out);

// Node-list products for callers (mass flux through the pressure nodes, std_one_phase/main.cpp:607-619):
// out[k] = vel(component, node_k) * rho(field, node_k); the host adds them in list order like the reference.
__global__ void invertLabelsKernel(const int32_t *__restrict__ label, int n, int32_t *__restrict__ slotOf);
__global__ void nodeFluxKernel(const int32_t *__restrict__ nodes, int count, const int32_t *__restrict__ slotOf, int nLabels,
                               const double *__restrict__ rho, const double *__restrict__ velComp, double *__restrict__ out);

// Set-up of the derived phi table: compares ptable with the slot the pull index yields and records the links where
// they differ.  counts != nullptr: counts[i] = words of node i's record (0: none); otherwise the records are written
// at the offsets in a.excInfo.
template <class L, int IDX>
__global__ void __launch_bounds__(128) phiExceptionKernel(const TwoPhaseArgs a, int32_t *counts, int32_t *exc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.n;
    if ((i & ~31) >= a.n) return;
    int s[L::nQ];
    resolveSources<L, IDX>(a, i, live, s);
    if (!live) return;
    uint32_t m = 0;
    int pt[L::nQ];
#pragma unroll
    for (int q = 0; q < L::nQ; ++q) {
        pt[q] = a.ptable[(long long)q * a.nPad + i];
        const int derived = (q == L::nQ - 1) ? i : s[reverseDir<L>(q)];
        if ((unsigned)derived >= (unsigned)a.n || derived != pt[q]) m |= 1u << q;
    }
    if (counts) {
        counts[i] = m ? 1 + __popc(m) : 0;
    } else if (m) {
        int32_t *rec = exc + (a.excInfo[i] - 1);
        *rec++ = (int32_t)m;
#pragma unroll
        for (int q = 0; q < L::nQ; ++q)
            if ((m >> q) & 1u) *rec++ = pt[q];
    }
}

template <class L, bool MOM, int IDX, bool DERIVED = false>
__global__ void __launch_bounds__(CHIMP_TP_BLOCK, CHIMP_TP_MIN_BLOCKS) twoPhaseCollideKernel(const TwoPhaseArgs a)
{
    const int i = a.begin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.end;
    if (IDX == IDX_TABLE && !live) return;
    if (IDX == IDX_COMPACT && (i & ~31) >= a.end) return;
    // DERIVED: the pull sources come first, the gradient's phi slots follow from them; otherwise the measured order of
    // the table form is kept (phi slots from ptable, gradient, then the pull sources)
    int s[L::nQ];
    if (DERIVED) resolveSources<L, IDX>(a, i, live, s);
    // colour gradient first (LButilities.h:12-22 -> LBd3q19.h:155-163): its Q gathered scalars are reduced
    // to nD numbers before the populations occupy the registers
    double g[3] = {0.0, 0.0, 0.0};
    double CGNorm = 0.0;
    if (live) {
        double ph[L::nQ];
        if (DERIVED) {
            const int info = __ldg(a.excInfo + i);
            const int32_t *rec = a.exc + (info ? info - 1 : 0);
            const uint32_t m = info ? (uint32_t)__ldg(rec) : 0u;
#pragma unroll
            for (int q = 0; q < L::nQ; ++q) {
                int slot = (q == L::nQ - 1) ? i : s[reverseDir<L>(q)];
                if ((m >> q) & 1u) slot = __ldg(rec + 1 + __popc(m & ((1u << q) - 1u)));
                ph[q] = a.phi[slot];
            }
        } else {
#pragma unroll
            for (int q = 0; q < L::nQ; ++q) ph[q] = a.phi[__ldg(a.ptable + ((unsigned)q * (unsigned)a.nPad + (unsigned)i))];
        }
        g[0] = latticeGrad<L, 0>(ph);
        g[1] = latticeGrad<L, 1>(ph);
        if (L::nD == 3) g[2] = latticeGrad<L, 2>(ph);
        CGNorm = sqrt(dotD<L>(g, g));
        const double inv = 1.0 / (CGNorm + (CGNorm < 2.220446049250313e-16 ? 1.0 : 0.0)); // lbBaseEps (LBglobal.h:14)
#pragma unroll
        for (int d = 0; d < L::nD; ++d) g[d] *= inv;
    }
    double fTot[L::nQ];
    const long long field1 = (long long)L::nQ * a.stride;
    if (!DERIVED) resolveSources<L, IDX>(a, i, live, s);
#pragma unroll
    for (int q = 0; q < L::nQ; ++q) {
        const double *p = a.pl.in[q] + s[q];
        fTot[q] = live ? __ldg(p) + __ldg(p + field1) : 0.0;
    }
    if (!live) return;
    const double rho0 = a.rho[i], rho1 = a.rho[(long long)a.nPad + i];
    const double rho = rho0 + rho1;
    double F[3] = {*a.forceX, a.F[1], a.F[2]};
    double u[3] = {0.0, 0.0, 0.0};
    u[0] = (firstMoment<L, 0>(fTot) + 0.5 * F[0]) / rho;
    u[1] = (firstMoment<L, 1>(fTot) + 0.5 * F[1]) / rho;
    if (L::nD == 3) u[2] = (firstMoment<L, 2>(fTot) + 0.5 * F[2]) / rho;
    if (MOM) {
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
    // opposite directions share every even intermediate and negate every odd one exactly (see the
    // single-field kernel); this also halves the divisions by |c_q| in the recolouring term
    const double c2uu = kC2 * uu, c2uF = kC2 * uF;
    const double negTauInv = -tauInv;
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
        a.pl.out[q][i] = c0 * commonQ + rc;
        a.pl.out[q][field1 + i] = c1 * commonQ - rc;
        a.pl.out[r][i] = c0 * commonR - rc;
        a.pl.out[r][field1 + i] = c1 * commonR + rc;
    };
    staticFor<L::nPairs>(pairBody);
    {
        constexpr int q = L::nQ - 1; // rest direction (LBcollision2phase.h:18,84)
        const double om = omegaBGK<L, q>(fTot[q], tauInv, rho, 0.0, uu);
        const double dF = deltaOmegaF<L, q>(tauFactor, 0.0, uF, 0.0);
        const double st = -AF0_5 * L::B(q);
        const double common = fTot[q] + om + dF + st;
        a.pl.out[q][i] = c0 * common + 0.0;
        a.pl.out[q][field1 + i] = c1 * common - 0.0;
    }
}

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

// state f_{s,q}(i) = w_q * rho_s(i) (initiateLbField with u = 0, LBinitiatefield.h:52-56), written through the
// pull table so that the first step pulls exactly these values; rho is [nFields][n] in device node order
template <class L>
__global__ void initEquilibriumKernel(const double *__restrict__ rho, double *X, const int32_t *__restrict__ table, int n,
                                      int nPad, long long stride, int nFields)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int fl = 0; fl < nFields; ++fl) {
        double *Xf = X + (long long)fl * L::nQ * stride;
        const double r = rho[(long long)fl * n + i];
#pragma unroll
        for (int q = 0; q < L::nQ; ++q) {
            const double v = L::w(q) * r * 1.0;
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

// Peer halos: the outgoing populations are stored straight into the neighbour GPU's halo-in slots over
// NVLink (peerX is the peer's buffer, mapped through CUDA IPC), field by field; the last block to finish
// publishes the step number in the peer's arrival flag.  waitFlagsKernel is the consumer side.
__global__ void haloPushKernel(double *peerX, const double *__restrict__ X, const long long *__restrict__ src,
                               const long long *__restrict__ dst, int count, int nFields, long long fieldStride,
                               long long peerFieldStride, unsigned *blockCounter, unsigned long long *peerFlag,
                               unsigned long long value);
// one double per rank summed over all ranks through peer memory (kernels.cu)
__global__ void sumWaitFoldKernel(const void *mail, int world, int parity, unsigned long long seq, double momx, double nGlobal,
                                  double *sumOut, double *forceX, const unsigned long long *flags, unsigned flagMask,
                                  unsigned long long flagExpect, unsigned long long timeoutNs, unsigned *error);
__global__ void foldAndPushKernel(const double *__restrict__ partial, int nBlocks, double *sumOut, void *const *peerMail, int rank,
                                  int world, int parity, unsigned long long seq);
__global__ void waitFlagsKernel(const unsigned long long *flags, unsigned mask, unsigned long long expect, unsigned long long timeoutNs,
                                unsigned *error);

// halo pack: buf[k] = X[src[k]] ; unpack: X[dst[k]] = buf[k]   (64-bit slot offsets)
__global__ void haloPackKernel(double *__restrict__ buf, const double *__restrict__ X,
                               const long long *__restrict__ src, int count);
__global__ void haloUnpackKernel(double *X, const double *__restrict__ buf, const long long *__restrict__ dst,
                                 int count);

} // namespace chimp
