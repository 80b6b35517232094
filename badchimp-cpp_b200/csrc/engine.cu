// Engine: host-side table builder, device memory, launch logic and the C-ABI (include/chimp_b200.h).
//
// The builder turns the reference's tables (Grid::neigList_, bulk list, bounce-back / link
// lists, MonLatMpi exchange lists) into the pull table T described in kernels.cuh by
// symbolically replaying one reference iteration:
//   1. push      fTmp(q, neighbor(q,n)) = f*_q(n) for bulk n in order     (LBfield.h:350-357)
//   2. exchange  real(q) <- ghost(q) of the neighbour rank                 (LBmonlatmpi.h:236-297)
//   3. boundary copies in the order the main applies them                  (LBhalfwaybb.h:37-63,
//                                                                           std_one_phase/main.cpp:138-203)
// Every slot of the reference's f array carries a symbolic source (which own node's
// post-collision value, in which orientation, or which halo element); sequential replay
// reproduces last-writer-wins and chained copies exactly.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <climits>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/chimp_b200.h"
#include "kernels.cuh"

namespace chimp {
__global__ void fillKernel(double *p, double v, long long count);
__global__ void classifyTilesKernel(const int32_t *, int, int, int, int, int32_t *, uint8_t *, int *);
__global__ void skipMaskKernel(const uint32_t *, int, int, int, int, int32_t *, unsigned long long *);
__global__ void fillRowsKernel(const int32_t *, int, int, int, int, const int32_t *, int32_t *, int, long long);
__global__ void kernelTableKernel(const int32_t *, int32_t *, int, int, int, int, long long);
} // namespace chimp

using namespace chimp;

namespace {

thread_local std::string g_err;
std::map<std::string, void *> g_ipcOpen; // CUDA IPC mappings of this process (kept until exit)
int openIpc(const unsigned char *handle64, void **out);
int labelRangeOf(chimp_lattice *c);
void preloadForPeerStepping(const chimp_lattice *c);
int buildPeerTables(chimp_lattice *c);
int buildPhiExceptions(chimp_lattice *c);
int checkDeviceError(chimp_lattice *c);
void awaitPeers(chimp_lattice *c);
void restoreConstants(chimp_lattice *c, double *X);
std::atomic<long long> g_launches{0};

int fail(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CUDA_OK(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct LatInfo { int nQ, nD, nPairs; };
LatInfo latInfo(int id)
{
    switch (id) {
    case CHIMP_D2Q9: return {D2Q9::nQ, D2Q9::nD, D2Q9::nPairs};
    case CHIMP_D3Q19: return {D3Q19::nQ, D3Q19::nD, D3Q19::nPairs};
    case CHIMP_D3Q27: return {D3Q27::nQ, D3Q27::nD, D3Q27::nPairs};
    }
    return {0, 0, 0};
}
int revDir(const LatInfo &li, int q) { return q == li.nQ - 1 ? q : (q + li.nPairs) % (li.nQ - 1); }

// one rank's contribution to a global sum, delivered into every rank's mailbox: value first, then its sequence number
struct MailSlot { double value; unsigned long long seq; };
constexpr int kMaxWorld = 64;

struct Op { int kind, dn, dq, sn, sq; }; // kind 0 copy, 1 anti bounce back, 2 swap, 3 constant (sn = index of the value)

struct Neighbor {
    int rank = -1;
    std::vector<int32_t> sendNodes, nDirSend, dirSend, recvNodes, nDirRecv, dirRecv;
    long long sendCount = 0, recvCount = 0; // per field
    long long *d_sendSrc = nullptr, *d_recvDst = nullptr;
    std::vector<long long> hSrc, hPeerDst; // host copies: send list and the peer slots it lands in (fused peer tables)
    std::vector<long long> hRecv;          // host copy of the receive list of a structured-ingest face (handed to the peer)
    long long *d_phiSendSrc = nullptr, *d_phiRecvDst = nullptr; // scalar (phi) halo: slots of the phi array
    double *d_phiSendBuf = nullptr, *d_phiRecvBuf = nullptr;
    long long phiSendCount = 0, phiRecvCount = 0;
    bool ownPhiBufs = true; // false: caller-owned scalar halo buffers (chimp_add_scalar_halo_face)
    // peer halos (CUDA IPC): the neighbour's two population buffers and its arrival flag for my face
    double *peerX[2] = {nullptr, nullptr};
    unsigned long long *peerFlags = nullptr;
    long long *d_peerDst = nullptr;
    long long peerFieldStride = 0;
    int peerFace = -1;
    unsigned *d_blockCounter = nullptr;
    // scalar (phi) peer halo: the neighbour's phi array, the ghost slots my values go to, its arrival flags
    double *peerPhi = nullptr;
    long long *d_peerPhiDst = nullptr;
    unsigned *d_phiBlockCounter = nullptr;
    double *d_sendBuf = nullptr, *d_recvBuf = nullptr;   // owned
    double *x_sendBuf = nullptr, *x_recvBuf = nullptr;   // caller-owned overrides (chimp_set_halo_buffers)
    double *sendBuf() const { return x_sendBuf ? x_sendBuf : d_sendBuf; }
    double *recvBuf() const { return x_recvBuf ? x_recvBuf : d_recvBuf; }
};

template <class T>
void freeDev(T *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

} // namespace

struct chimp_lattice {
    int lattice = 0;
    LatInfo li{};
    int device = 0;
    int nFields = 1;
    // host-side inputs (released by finalize)
    int nNodes = 0;
    std::vector<int32_t> neigh, bulk;
    std::vector<Op> ops;
    // constant links (PressureBnd / InletOutlet, LBpressurebnd.h:10-88): values by registration order, the halo-in slot each
    // one occupies (offset inside one field's planes), and their device copies
    std::vector<double> constValues;
    std::vector<long long> hConstDst;
    double *d_constVal = nullptr;
    long long *d_constDst = nullptr;
    bool constDirty = false; // the buffer read next holds uploaded / initial values in the constant links' slots
    std::vector<Neighbor> nbrs;
    std::vector<int32_t> solidBnd;
    bool finalized = false, hostBuilt = false;
    // host-built tables (chimp_build_host)
    std::vector<int32_t> hTable, hLabel;
    std::vector<uint32_t> hPmask;
    std::vector<std::vector<long long>> hSendSrc, hRecvDst;
    // device layout
    int n = 0, nPad = 0, nHalo = 0, nBoundary = 0;
    long long stride = 0;
    int indexForm = CHIMP_INDEX_TABLE;
    int32_t *d_table = nullptr, *d_ktable = nullptr, *d_label = nullptr;
    int labelMin = 0, labelMax = -1;
    bool labelsContiguous = false;
    uint32_t *d_delta = nullptr, *d_pmask = nullptr;
    bool skipMask = false;              // single-field step kernel in IDX_COMPACT_MASK form (chimp_set_index_skip_mask)
    unsigned long long skippedWords = 0; // (tile, delta word) pairs that form skips
    uint32_t *d_attr = nullptr; // one_phase attributes packed into one word per node (kernels.cuh, StepArgs::attr), if they pack
    bool attrPackEnv = true;
    int nWords = 0;
    int32_t *d_base = nullptr, *d_rows = nullptr;
    int nTiles = 0, nRows = 0;
    double *d_f[2] = {nullptr, nullptr};
    int cur = 0;
    double *d_rho = nullptr, *d_vel = nullptr;
    bool hasPressure = false;
    // two-phase (colour gradient)
    std::vector<int32_t> hPtable;
    int nPhi = 0, nSolid = 0, nGhost = 0;
    int32_t *d_ptable = nullptr;
    int32_t *d_excInfo = nullptr, *d_exc = nullptr; // derived form of ptable (kernels.cuh, TwoPhaseArgs)
    long long excWords = 0;
    bool phiDerived = false, phiDerivedEnv = true;
    double *d_phi = nullptr, *d_fluxPartial = nullptr, *d_fluxSum = nullptr, *d_forceX = nullptr;
    bool densitySet = false;
    // one-phase attributes
    bool onePhase = false;
    double *d_forceOn = nullptr, *d_addSource = nullptr, *d_srcPerLabel = nullptr, *d_massPartial = nullptr;
    int32_t *d_labelAttr = nullptr;
    double *d_scale = nullptr, *d_mass = nullptr;
    int nLabels = 0;
    std::vector<double> scalePerLabel;
    double rhoW = 1.0;
    // streams
    cudaStream_t stream = nullptr, haloStream = nullptr;
    bool ownStream = false;
    cudaEvent_t evBoundary = nullptr, evHalo = nullptr, evStep = nullptr;
    chimp_exchange_fn exchange = nullptr, scalarExchange = nullptr;
    void *exchangeUser = nullptr, *scalarExchangeUser = nullptr;
    unsigned long long *d_flags = nullptr; // arrival counters written by the neighbours, one per face (max 8)
    bool peerHalos = false;
    // peer exchange fused into the step kernel (PeerView): per halo-coupled node the directions that leave the rank
    // and where they land; built once every face is connected
    bool peerFused = false;
    uint32_t *d_sendMask = nullptr;
    int32_t *d_sendDst = nullptr, *d_extraStart = nullptr;
    int2 *d_extra = nullptr;
    int peerPad = 0, peerBlocks = 0;
    unsigned *d_peerCounter = nullptr;
    unsigned *h_error = nullptr, *d_error = nullptr; // host-mapped error word raised by a device-side arrival timeout
    unsigned long long timeoutNs = 20000000000ull;
    unsigned long long *d_trace = nullptr;           // CHIMP_TRACE: per-step device timestamps [traceCap][4]
    unsigned long long *traceCursor = nullptr;
    int traceCap = 0;
    bool trace = false, peerFusedEnv = true;
    std::string peerWhy;
    // two-phase over peer memory: sum mailbox of every rank (world pointers on the device), my rank / world size
    MailSlot *d_mail = nullptr;
    MailSlot **d_peerMail = nullptr;
    int worldRank = -1, worldSize = 0;
    bool peerTwoPhase = false;
    chimp_allreduce_fn allreduce = nullptr;
    void *allreduceUser = nullptr;
    std::vector<std::vector<long long>> hPhiSendSrc, hPhiRecvDst;
    int32_t *d_slotOf = nullptr; // reference label -> device slot (-1: not an own node), built on first use
    long long steps = 0;
};

namespace {

// the constant links' slots of buffer X take their values again (after anything that wrote whole planes or
// scattered a state through the pull table); stream-ordered on the engine's stream
void restoreConstants(chimp_lattice *c, double *X)
{
    const long long n = (long long)c->hConstDst.size();
    if (!n || !c->d_constDst) return;
    haloUnpackKernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(X, c->d_constVal, c->d_constDst, (int)n);
    ++g_launches;
}

int check(chimp_lattice *c, bool needFinal)
{
    if (!c) return fail("null lattice handle");
    if (needFinal && !c->finalized) return fail("lattice not finalized (call chimp_finalize first)");
    if (!needFinal && c->finalized) return fail("lattice already finalized");
    return 0;
}

int allocateState(chimp_lattice *c)
{
    const size_t planeBytes = (size_t)c->stride * c->li.nQ * c->nFields * sizeof(double);
    for (int b = 0; b < 2; ++b) {
        CUDA_OK(cudaMalloc(&c->d_f[b], planeBytes));
        CUDA_OK(cudaMemsetAsync(c->d_f[b], 0, planeBytes, c->stream));
    }
    CUDA_OK(cudaMalloc(&c->d_rho, (size_t)c->nPad * c->nFields * sizeof(double)));
    CUDA_OK(cudaMalloc(&c->d_vel, (size_t)c->nPad * c->li.nD * sizeof(double)));
    CUDA_OK(cudaMemsetAsync(c->d_rho, 0, (size_t)c->nPad * c->nFields * sizeof(double), c->stream));
    CUDA_OK(cudaMemsetAsync(c->d_vel, 0, (size_t)c->nPad * c->li.nD * sizeof(double), c->stream));
    // arrival counters written by the neighbours: [0, 8) population faces, [8, 16) scalar (phi) faces
    CUDA_OK(cudaMalloc(&c->d_flags, 16 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemsetAsync(c->d_flags, 0, 16 * sizeof(unsigned long long), c->stream));
    // error word raised by device-side waits that time out (a peer that never arrives): host-mapped, so it can be
    // read even while a kernel is still running
    CUDA_OK(cudaHostAlloc((void **)&c->h_error, sizeof(unsigned), cudaHostAllocMapped));
    *c->h_error = 0u;
    CUDA_OK(cudaHostGetDevicePointer((void **)&c->d_error, c->h_error, 0));
    // environment switches are read once, here
    auto envInt = [](const char *name, long long dflt) {
        const char *v = getenv(name);
        return v ? atoll(v) : dflt;
    };
    c->trace = envInt("CHIMP_TRACE", 0) == 1;
    c->peerFusedEnv = envInt("CHIMP_PEER_FUSED", 1) != 0;
    c->phiDerivedEnv = envInt("CHIMP_PHI_DERIVED", 1) != 0;
    c->attrPackEnv = envInt("CHIMP_ATTR_PACKED", 1) != 0;
    c->timeoutNs = (unsigned long long)std::max(1ll, envInt("CHIMP_PEER_TIMEOUT_MS", 20000)) * 1000000ull;
    return 0;
}

// a device-side wait gave up (CHIMP_PEER_TIMEOUT_MS): the state is no longer meaningful
int checkDeviceError(chimp_lattice *c)
{
    if (!c->h_error || *c->h_error == 0u) return 0;
    const unsigned code = *c->h_error;
    return fail("rank %d: a neighbour rank did not arrive within %.1f s (%s); all ranks must step the same number of times",
                c->worldRank, c->timeoutNs * 1e-9,
                code == 1u ? "population halo counter" : code == 2u ? "global-sum mailbox" : "scalar halo counter");
}

// peer halos: everything that reads or overwrites halo-in slots outside the step kernels (transfers, reductions, the
// mass-change pass) first waits, on the engine's stream, until every neighbour's stores of the last step have landed
void awaitPeers(chimp_lattice *c)
{
    if (!c->peerHalos || c->nbrs.empty() || c->steps == 0) return;
    waitFlagsKernel<<<1, 32, 0, c->stream>>>(c->d_flags, (1u << c->nbrs.size()) - 1u, (unsigned long long)c->steps, c->timeoutNs, c->d_error);
    ++g_launches;
}

int buildKernelTable(chimp_lattice *c)
{
    const long long total = (long long)c->li.nQ * c->nPad;
    CUDA_OK(cudaMalloc(&c->d_ktable, (size_t)total * sizeof(int32_t)));
    kernelTableKernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(c->d_table, c->d_ktable, c->n, c->nPad, c->li.nQ, c->li.nPairs, c->stride);
    ++g_launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int checkIndexRange(chimp_lattice *c)
{
    // slot indices are signed 32-bit relative to a plane, bounces reach nPairs planes away
    if ((long long)c->li.nPairs * c->stride + c->nPad >= (1ll << 31))
        return fail("lattice too large for 32-bit plane-relative indices (%lld slots per plane)", c->stride);
    return 0;
}

int buildRankIndex(chimp_lattice *c)
{
    const int nQ = c->li.nQ;
    c->nTiles = c->nPad / 32;
    c->nWords = (nQ + 3) / 4;
    CUDA_OK(cudaMalloc(&c->d_delta, (size_t)c->nWords * c->nPad * sizeof(uint32_t)));
    CUDA_OK(cudaMemsetAsync(c->d_delta, 0xff, (size_t)c->nWords * c->nPad * sizeof(uint32_t), c->stream));
    CUDA_OK(cudaMalloc(&c->d_base, (size_t)c->nTiles * c->nWords * 4 * sizeof(int32_t)));
    CUDA_OK(cudaMemsetAsync(c->d_base, 0, (size_t)c->nTiles * c->nWords * 4 * sizeof(int32_t), c->stream));
    int *d_cnt = nullptr;
    CUDA_OK(cudaMalloc(&d_cnt, sizeof(int)));
    CUDA_OK(cudaMemsetAsync(d_cnt, 0, sizeof(int), c->stream));
    const long long threads = (long long)c->nTiles * nQ * 32;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    classifyTilesKernel<<<grid, 256, 0, c->stream>>>(c->d_table, c->n, c->nPad, nQ, c->nTiles, c->d_base, (uint8_t *)c->d_delta, d_cnt);
    ++g_launches;
    CUDA_OK(cudaGetLastError());
    int cnt = 0;
    CUDA_OK(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    cudaFree(d_cnt);
    c->nRows = cnt;
    {   // skip mask of the delta words (IDX_COMPACT_MASK): always built, it lives in a slot the plain compact form never reads
        unsigned long long *d_skipped = nullptr;
        CUDA_OK(cudaMalloc(&d_skipped, sizeof(unsigned long long)));
        CUDA_OK(cudaMemsetAsync(d_skipped, 0, sizeof(unsigned long long), c->stream));
        skipMaskKernel<<<(unsigned)((c->nTiles + 127) / 128), 128, 0, c->stream>>>(c->d_delta, c->n, c->nPad, nQ, c->nTiles, c->d_base, d_skipped);
        ++g_launches;
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpyAsync(&c->skippedWords, d_skipped, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        cudaFree(d_skipped);
    }
    CUDA_OK(cudaMalloc(&c->d_rows, (size_t)std::max(cnt, 1) * 32 * sizeof(int32_t)));
    if (cnt > 0) {
        fillRowsKernel<<<grid, 256, 0, c->stream>>>(c->d_table, c->n, c->nPad, nQ, c->nTiles, c->d_base, c->d_rows, c->li.nPairs, c->stride);
        ++g_launches;
        CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int setupStreams(chimp_lattice *c)
{
    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->ownStream = true;
    // halo work (unpack kernels, the host's transport) must not queue behind the interior blocks
    int prioLow = 0, prioHigh = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh));
    CUDA_OK(cudaStreamCreateWithPriority(&c->haloStream, cudaStreamNonBlocking, prioHigh));
    CUDA_OK(cudaEventCreateWithFlags(&c->evBoundary, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->evHalo, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->evStep, cudaEventDisableTiming));
    return 0;
}

template <class L, int COLL, int ONEPHASE, int IDX>
void launchSingle(const StepArgs &a, bool mom, cudaStream_t s)
{
    const int count = a.end - a.begin;
    if (count <= 0) return;
    const int block = CHIMP_BLOCK;
    const unsigned grid = (unsigned)((count + block - 1) / block);
    if (IDX == IDX_COMPACT && a.peer.blocks > 0) {
        // halo-coupled blocks first, their remote stores and arrival counters inside the launch
        if (mom) collideStreamKernel<L, COLL, ONEPHASE, true, IDX_COMPACT, true><<<grid, block, 0, s>>>(a);
        else collideStreamKernel<L, COLL, ONEPHASE, false, IDX_COMPACT, true><<<grid, block, 0, s>>>(a);
    } else if (mom) collideStreamKernel<L, COLL, ONEPHASE, true, IDX><<<grid, block, 0, s>>>(a);
    else collideStreamKernel<L, COLL, ONEPHASE, false, IDX><<<grid, block, 0, s>>>(a);
    ++g_launches;
}

template <class L>
void dispatchSingle(const chimp_lattice *c, const StepArgs &a, int coll, bool mom, cudaStream_t s)
{
    const int op = !c->onePhase ? OP_NONE : (a.attr ? OP_PACKED : OP_ARRAYS);
    const bool rk = c->indexForm == CHIMP_INDEX_COMPACT;
#define CH_LAUNCH(COLL, OP, IDX) launchSingle<L, COLL, OP, IDX>(a, mom, s)
#define CH_INDEX(COLL, OP) do { if (rk) CH_LAUNCH(COLL, OP, IDX_COMPACT); else CH_LAUNCH(COLL, OP, IDX_TABLE); } while (0)
    // the skip-mask form exists for the plain single-field kernel outside the fused peer launch
    const bool masked = rk && c->skipMask && op == OP_NONE && a.peer.blocks == 0;
#define CH_ATTR(COLL) do { if (masked) CH_LAUNCH(COLL, OP_NONE, IDX_COMPACT_MASK); else if (op == OP_PACKED) CH_INDEX(COLL, OP_PACKED); \
                           else if (op == OP_ARRAYS) CH_INDEX(COLL, OP_ARRAYS); else CH_INDEX(COLL, OP_NONE); } while (0)
    if (coll == CHIMP_BGK) CH_ATTR(COLL_BGK);
    else CH_ATTR(COLL_TRT);
#undef CH_ATTR
#undef CH_INDEX
#undef CH_LAUNCH
}

int massChangePass(chimp_lattice *c, const StepArgs &a);

template <class L>
void launchMassChange(const chimp_lattice *c, const StepArgs &a, unsigned grid)
{
    if (c->indexForm == CHIMP_INDEX_COMPACT) massChangeKernel<L, IDX_COMPACT><<<grid, 256, 0, c->stream>>>(a, c->nLabels, c->d_massPartial);
    else massChangeKernel<L, IDX_TABLE><<<grid, 256, 0, c->stream>>>(a, c->nLabels, c->d_massPartial);
}

int massChangePass(chimp_lattice *c, const StepArgs &a)
{
    if (!c->nbrs.empty() && !c->allreduce)
        return fail("mass-conservation source across ranks needs chimp_set_allreduce_callback (std_one_phase/main.cpp:528)");
    const unsigned grid = (unsigned)((c->n + 255) / 256);
    if (!c->d_massPartial) {
        CUDA_OK(cudaMalloc(&c->d_massPartial, (size_t)grid * c->nLabels * sizeof(double)));
    }
    switch (c->lattice) {
    case CHIMP_D2Q9: launchMassChange<D2Q9>(c, a, grid); break;
    case CHIMP_D3Q19: launchMassChange<D3Q19>(c, a, grid); break;
    case CHIMP_D3Q27: launchMassChange<D3Q27>(c, a, grid); break;
    }
    massFinalizeKernel<<<c->nLabels, 256, 0, c->stream>>>(c->d_massPartial, (int)grid, c->d_scale, c->d_mass, c->d_srcPerLabel);
    g_launches += 2;
    CUDA_OK(cudaGetLastError());
    if (!c->nbrs.empty()) {
        // MPI_Allreduce of massChangeLocal (main.cpp:528), then the source factors from the global sums
        if (c->allreduce(c->allreduceUser, c->d_mass, c->nLabels, (void *)c->stream)) return fail("allreduce callback failed");
        massFinalizeKernel<<<c->nLabels, 256, 0, c->stream>>>(c->d_mass, -1, c->d_scale, c->d_mass, c->d_srcPerLabel);
        ++g_launches;
        CUDA_OK(cudaGetLastError());
    }
    return 0;
}

void dispatchSingleLattice(const chimp_lattice *c, const StepArgs &a, int coll, bool mom, cudaStream_t s)
{
    switch (c->lattice) {
    case CHIMP_D2Q9: dispatchSingle<D2Q9>(c, a, coll, mom, s); break;
    case CHIMP_D3Q19: dispatchSingle<D3Q19>(c, a, coll, mom, s); break;
    case CHIMP_D3Q27: dispatchSingle<D3Q27>(c, a, coll, mom, s); break;
    }
}

} // namespace

// =========================================================================================
// C-ABI
// =========================================================================================
extern "C" {

const char *chimp_last_error(void) { return g_err.c_str(); }
int chimp_version(void) { return 100; }
long long chimp_launch_count(void) { return g_launches.load(); }
int chimp_lattice_nq(int l) { return latInfo(l).nQ; }
int chimp_lattice_nd(int l) { return latInfo(l).nD; }
int chimp_lattice_c(int l, int q, int d)
{
    switch (l) {
    case CHIMP_D2Q9: return D2Q9::c(q, d);
    case CHIMP_D3Q19: return D3Q19::c(q, d);
    case CHIMP_D3Q27: return D3Q27::c(q, d);
    }
    return 0;
}
double chimp_lattice_w(int l, int q)
{
    switch (l) {
    case CHIMP_D2Q9: return D2Q9::w(q);
    case CHIMP_D3Q19: return D3Q19::w(q);
    case CHIMP_D3Q27: return D3Q27::w(q);
    }
    return 0.0;
}
int chimp_lattice_reverse(int l, int q) { return revDir(latInfo(l), q); }

int chimp_create(chimp_lattice **out, int lattice, int n_nodes, const int32_t *neigh, int n_bulk,
                 const int32_t *bulk, int n_fields, int device)
{
    if (!out) return fail("out is null");
    *out = nullptr;
    const LatInfo li = latInfo(lattice);
    if (li.nQ == 0) return fail("unknown lattice id %d", lattice);
    if (n_fields < 1 || n_fields > 2) return fail("n_fields must be 1 or 2, got %d", n_fields);
    if (n_nodes < 1 || n_bulk < 0 || !neigh || (!bulk && n_bulk > 0)) return fail("bad table arguments");
    for (int b = 0; b < n_bulk; ++b) {
        if (bulk[b] <= 0 || bulk[b] >= n_nodes) return fail("bulk[%d]=%d outside (0,%d)", b, bulk[b], n_nodes);
        if (b && bulk[b] <= bulk[b - 1]) return fail("bulk list must be strictly ascending (LBgeometry.h:11-21)");
    }
    for (long long k = 0; k < (long long)n_nodes * li.nQ; ++k)
        if (neigh[k] < 0 || neigh[k] >= n_nodes) return fail("neighbor entry %lld = %d outside [0,%d)", k, neigh[k], n_nodes);
    chimp_lattice *c = new chimp_lattice;
    c->lattice = lattice;
    c->li = li;
    c->device = device;
    c->nFields = n_fields;
    c->nNodes = n_nodes;
    c->neigh.assign(neigh, neigh + (size_t)n_nodes * li.nQ);
    c->bulk.assign(bulk, bulk + n_bulk);
    *out = c;
    return 0;
}

int chimp_add_halfway_bb(chimp_lattice *c, int n_bnd, const int32_t *nodes, const int32_t *n_beta,
                         const int32_t *n_gamma, const int32_t *n_delta, const int32_t *links)
{
    if (check(c, false)) return 1;
    const int nQ = c->li.nQ, P = c->li.nPairs;
    if (n_bnd < 0 || (n_bnd > 0 && (!nodes || !n_beta || !n_gamma || !n_delta || !links))) return fail("bad bounce-back arguments");
    for (int b = 0; b < n_bnd; ++b) {
        const int node = nodes[b];
        if (node <= 0 || node >= c->nNodes) return fail("bounce-back node %d out of range", node);
        if (n_beta[b] < 0 || n_gamma[b] < 0 || n_delta[b] < 0) return fail("bounce-back node %d: negative link count", node);
        if (n_beta[b] + n_gamma[b] + n_delta[b] != P) return fail("bounce-back node %d: link counts do not sum to nDirPairs", node);
        const int32_t *l = links + (size_t)b * P;
        for (int k = 0; k < P; ++k)
            if (l[k] < 0 || l[k] >= nQ - 1) return fail("bounce-back node %d: link direction %d outside [0,%d)", node, l[k], nQ - 1);
        // LBhalfwaybb.h:52-61
        for (int k = 0; k < n_beta[b]; ++k) {
            const int beta = l[k], br = revDir(c->li, beta);
            c->ops.push_back({0, node, beta, c->neigh[(size_t)node * nQ + br], br});
        }
        for (int k = 0; k < n_delta[b]; ++k) {
            const int d = l[n_beta[b] + n_gamma[b] + k], dr = revDir(c->li, d);
            c->ops.push_back({0, node, d, c->neigh[(size_t)node * nQ + dr], dr});
            c->ops.push_back({0, node, dr, c->neigh[(size_t)node * nQ + d], d});
        }
    }
    return 0;
}

int chimp_add_links(chimp_lattice *c, int kind, int n_links, const int32_t *l4)
{
    if (check(c, false)) return 1;
    if (kind < 0 || kind > 2) return fail("unknown link kind %d", kind);
    for (int k = 0; k < n_links; ++k) {
        const int nf = l4[4 * k], qu = l4[4 * k + 1], nw = l4[4 * k + 2], qk = l4[4 * k + 3];
        if (nf <= 0 || nf >= c->nNodes || nw < 0 || nw >= c->nNodes || qu < 0 || qu >= c->li.nQ || qk < 0 || qk >= c->li.nQ)
            return fail("link %d out of range", k);
        c->ops.push_back({kind, nf, qu, nw, qk});
    }
    if (kind == CHIMP_LINK_PRESSURE && n_links > 0) c->hasPressure = true;
    return 0;
}

int chimp_add_constant_links(chimp_lattice *c, int n_links, const int32_t *node_q, const double *values)
{
    if (check(c, false)) return 1;
    if (c->nFields != 1) return fail("constant links need a one-field lattice (both fields of a two-field lattice share the pull table)");
    if (n_links < 0 || (n_links > 0 && (!node_q || !values))) return fail("bad constant-link arguments");
    for (int k = 0; k < n_links; ++k) {
        const int node = node_q[2 * k], q = node_q[2 * k + 1];
        if (node < 0 || node >= c->nNodes || q < 0 || q >= c->li.nQ) return fail("constant link %d: node %d / direction %d out of range", k, node, q);
        c->ops.push_back({3, node, q, (int)c->constValues.size(), 0});
        c->constValues.push_back(values[k]);
    }
    return 0;
}

int chimp_add_neighbor(chimp_lattice *c, int neig_rank, int n_send, const int32_t *nodes_to_send,
                       const int32_t *n_dir_send, const int32_t *dir_list_send, int n_recv,
                       const int32_t *nodes_received, const int32_t *n_dir_recv, const int32_t *dir_list_recv)
{
    if (check(c, false)) return 1;
    if (!c->nbrs.empty() && c->nbrs.back().rank >= neig_rank) return fail("neighbour ranks must be added in ascending order");
    if (n_send < 0 || n_recv < 0 || (n_send > 0 && (!nodes_to_send || !n_dir_send || !dir_list_send)) ||
        (n_recv > 0 && (!nodes_received || !n_dir_recv || !dir_list_recv)))
        return fail("bad exchange lists for neighbour %d", neig_rank);
    for (int k = 0; k < n_send; ++k)
        if (n_dir_send[k] < 0) return fail("neighbour %d: negative direction count", neig_rank);
    for (int k = 0; k < n_recv; ++k)
        if (n_dir_recv[k] < 0) return fail("neighbour %d: negative direction count", neig_rank);
    Neighbor nb;
    nb.rank = neig_rank;
    nb.sendNodes.assign(nodes_to_send, nodes_to_send + n_send);
    nb.nDirSend.assign(n_dir_send, n_dir_send + n_send);
    long long ns = 0;
    for (int k = 0; k < n_send; ++k) ns += n_dir_send[k];
    nb.dirSend.assign(dir_list_send, dir_list_send + ns);
    nb.recvNodes.assign(nodes_received, nodes_received + n_recv);
    nb.nDirRecv.assign(n_dir_recv, n_dir_recv + n_recv);
    long long nr = 0;
    for (int k = 0; k < n_recv; ++k) nr += n_dir_recv[k];
    nb.dirRecv.assign(dir_list_recv, dir_list_recv + nr);
    nb.sendCount = ns;
    nb.recvCount = nr;
    c->nbrs.push_back(std::move(nb));
    return 0;
}

int chimp_set_solid_boundary(chimp_lattice *c, int n_solid, const int32_t *solid_nodes)
{
    if (check(c, false)) return 1;
    if (n_solid < 0 || (n_solid > 0 && !solid_nodes)) return fail("bad solid boundary list");
    c->solidBnd.assign(solid_nodes, solid_nodes + n_solid);
    return 0;
}

int chimp_build_host(chimp_lattice *c, int boundary_first)
{
    if (check(c, false)) return 1;
    if (c->hostBuilt) return fail("host tables already built");
    const LatInfo li = c->li;
    const int nQ = li.nQ;
    const int nBulk = (int)c->bulk.size();
    const int32_t UNSET = -1;
    const int32_t POISON = INT32_MIN; // a value whose orientation the kernels cannot express; an error only if an own node pulls it
    // symbolic slot contents: >= 0: (bulkIndex << 1) | reversed ; -1 unset ; <= -2: halo element -2-h
    std::vector<int32_t> sym((size_t)c->nNodes * nQ, UNSET);
    std::vector<int32_t> bulkIndex(c->nNodes, -1);
    for (int b = 0; b < nBulk; ++b) bulkIndex[c->bulk[b]] = b;
    if (nBulk >= (1 << 30)) return fail("too many nodes for the 32-bit symbolic builder");
    // 1. push (LBfield.h:350-357), bulk order => last writer wins
    for (int b = 0; b < nBulk; ++b) {
        const size_t row = (size_t)c->bulk[b] * nQ;
        for (int q = 0; q < nQ; ++q) sym[(size_t)c->neigh[row + q] * nQ + q] = b << 1;
    }
    // 2. ghost exchange (LBmonlatmpi.h:236-297)
    long long haloTotal = 0;
    std::vector<std::vector<long long>> sendSrc(c->nbrs.size()); // q * nBulk + b (bulk-index space)
    std::vector<std::vector<int>> haloDirCount(1, std::vector<int>(nQ, 0));
    std::vector<std::pair<int, int>> haloQH; // halo element -> (q, index within direction q)
    for (size_t k = 0; k < c->nbrs.size(); ++k) {
        Neighbor &nb = c->nbrs[k];
        size_t cnt = 0;
        for (size_t s = 0; s < nb.sendNodes.size(); ++s)
            for (int j = 0; j < nb.nDirSend[s]; ++j, ++cnt) {
                const int q = nb.dirSend[cnt];
                const int node = nb.sendNodes[s];
                if (q < 0 || q >= nQ || node < 0 || node >= c->nNodes) return fail("neighbour %d: bad send entry", nb.rank);
                const int ghost = c->neigh[(size_t)node * nQ + q];
                const int32_t code = sym[(size_t)ghost * nQ + q];
                if (code < 0 || (code & 1)) return fail("neighbour %d: send slot (q=%d, ghost node %d) was not pushed by an own node", nb.rank, q, ghost);
                sendSrc[k].push_back((long long)q * nBulk + (code >> 1));
            }
    }
    for (size_t k = 0; k < c->nbrs.size(); ++k) {
        Neighbor &nb = c->nbrs[k];
        size_t cnt = 0;
        for (size_t s = 0; s < nb.recvNodes.size(); ++s)
            for (int j = 0; j < nb.nDirRecv[s]; ++j, ++cnt) {
                const int q = nb.dirRecv[cnt];
                const int node = nb.recvNodes[s];
                if (q < 0 || q >= nQ || node < 0 || node >= c->nNodes) return fail("neighbour %d: bad recv entry", nb.rank);
                const int real = c->neigh[(size_t)node * nQ + q];
                sym[(size_t)real * nQ + q] = (int32_t)(-2 - haloTotal);
                haloQH.push_back({q, haloDirCount[0][q]++});
                ++haloTotal;
                if (haloTotal > (1ll << 30)) return fail("halo too large");
            }
    }
    // 3. boundary copies in application order
    std::vector<uint32_t> pmaskB(nBulk, 0);
    std::vector<std::pair<long long, int>> constElem; // (halo element, index of its value)
    auto orient = [&](int32_t code, int sq, int dq, int32_t &outCode) -> bool {
        if (code == UNSET || code == POISON) { outCode = code; return true; }
        if (code <= -2) { outCode = sq == dq ? code : POISON; return true; }
        const int actual = (code & 1) ? revDir(li, sq) : sq;
        if (actual == dq) outCode = code & ~1;
        else if (actual == revDir(li, dq)) outCode = code | 1;
        else outCode = POISON;
        return true;
    };
    for (const Op &op : c->ops) {
        const size_t d = (size_t)op.dn * nQ + op.dq, s = (size_t)op.sn * nQ + op.sq;
        int32_t v;
        if (op.kind == 3) {
            // the destination takes a constant: a halo-in element nobody sends to, filled once by the engine
            sym[d] = (int32_t)(-2 - haloTotal);
            haloQH.push_back({op.dq, haloDirCount[0][op.dq]++});
            constElem.push_back({haloTotal, op.sn});
            ++haloTotal;
            if (haloTotal > (1ll << 30)) return fail("halo too large");
        } else if (op.kind == 0) {
            if (!orient(sym[s], op.sq, op.dq, v)) return fail("boundary copy (%d,%d)<-(%d,%d): unsupported direction change", op.dn, op.dq, op.sn, op.sq);
            sym[d] = v;
        } else if (op.kind == 1) {
            const int32_t code = sym[s];
            const int b = bulkIndex[op.dn];
            if (b < 0 || code != (b << 1) || op.dq != revDir(li, op.sq))
                return fail("pressure link at node %d: the known population must be the node's own pushed value", op.dn);
            pmaskB[b] |= 1u << op.sq;
            sym[d] = code | 1;
        } else {
            int32_t a, bcode;
            if (!orient(sym[s], op.sq, op.dq, a) || !orient(sym[d], op.dq, op.sq, bcode))
                return fail("fluid-fluid link (%d,%d)<->(%d,%d): unsupported direction change", op.dn, op.dq, op.sn, op.sq);
            sym[d] = a;
            sym[s] = bcode;
        }
    }
    // 4. device order: halo-coupled nodes first when requested
    std::vector<int32_t> devOf(nBulk, -1), order;
    order.reserve(nBulk);
    int nBoundary = 0;
    if (boundary_first && !c->nbrs.empty()) {
        std::vector<char> coupled(nBulk, 0);
        for (auto &v : sendSrc)
            for (long long e : v) coupled[e % nBulk] = 1;
        for (int b = 0; b < nBulk; ++b) {
            const size_t row = (size_t)c->bulk[b] * nQ;
            for (int q = 0; q < nQ && !coupled[b]; ++q)
                if (sym[row + q] <= -2) coupled[b] = 1;
        }
        for (int b = 0; b < nBulk; ++b) if (coupled[b]) order.push_back(b);
        nBoundary = (int)order.size();
        for (int b = 0; b < nBulk; ++b) if (!coupled[b]) order.push_back(b);
    } else {
        for (int b = 0; b < nBulk; ++b) order.push_back(b);
    }
    const int nSlots = (int)order.size();
    if (nSlots == 0) return fail("the lattice has no own fluid nodes (empty bulk list)");
    for (int i = 0; i < nSlots; ++i) if (order[i] >= 0) devOf[order[i]] = i;
    c->n = nSlots;
    // the first launch covers whole warp tiles: it may include a few uncoupled nodes
    c->nBoundary = nBoundary ? std::min(((nBoundary + 31) / 32) * 32, nSlots) : 0;
    c->nPad = ((nSlots + 31) / 32) * 32;
    int maxHaloDir = 0;
    for (int q = 0; q < nQ; ++q) maxHaloDir = std::max(maxHaloDir, haloDirCount[0][q]);
    c->nHalo = ((maxHaloDir + 15) / 16) * 16;
    c->stride = (long long)c->nPad + c->nHalo;
    // 5. pull table in device order
    std::vector<int32_t> table((size_t)nQ * c->nPad, -1), label(c->nPad, 0);
    std::vector<uint32_t> pmask(c->nPad, 0);
    std::vector<uint8_t> pulled((size_t)nBulk * nQ, 0);
    for (int i = 0; i < nSlots; ++i) {
        const int b = order[i];
        label[i] = c->bulk[b];
        pmask[i] = pmaskB[b];
        const size_t row = (size_t)c->bulk[b] * nQ;
        for (int q = 0; q < nQ; ++q) {
            const int32_t code = sym[row + q];
            int32_t t;
            if (code == UNSET)
                return fail("node %d direction %d has no upstream writer: geometry must be closed (walls with bounce back) or periodic", c->bulk[b], q);
            if (code == POISON)
                return fail("node %d direction %d receives a population through a boundary copy that changes its direction (unsupported link structure)", c->bulk[b], q);
            if (code <= -2) {
                const auto &qh = haloQH[(size_t)(-2 - code)];
                t = c->nPad + qh.second;
            } else {
                const int sb = code >> 1;
                const int sq = (code & 1) ? revDir(li, q) : q;
                if (pulled[(size_t)sb * nQ + sq]++) return fail("population (q=%d, node %d) is pulled twice: open boundary through the shared dummy node 0", sq, c->bulk[sb]);
                if (code & 1) {
                    if (sb != b) return fail("node %d direction %d bounces a foreign node's population (unsupported)", c->bulk[b], q);
                    t = -1;
                } else t = devOf[sb];
            }
            table[(size_t)q * c->nPad + i] = t;
        }
    }
    // 6. halo lists (slot offsets inside one field's planes)
    long long haloOff = 0;
    c->hSendSrc.resize(c->nbrs.size());
    c->hRecvDst.resize(c->nbrs.size());
    for (size_t k = 0; k < c->nbrs.size(); ++k) {
        Neighbor &nb = c->nbrs[k];
        std::vector<long long> &src = c->hSendSrc[k], &dst = c->hRecvDst[k];
        src.resize(sendSrc[k].size());
        dst.resize((size_t)nb.recvCount);
        for (size_t e = 0; e < src.size(); ++e) {
            const long long q = sendSrc[k][e] / nBulk, b = sendSrc[k][e] % nBulk;
            src[e] = q * c->stride + devOf[b];
        }
        for (long long e = 0; e < nb.recvCount; ++e) {
            const auto &qh = haloQH[(size_t)(haloOff + e)];
            dst[e] = (long long)qh.first * c->stride + c->nPad + qh.second;
        }
        haloOff += nb.recvCount;
    }
    c->hConstDst.assign(c->constValues.size(), -1);
    for (const auto &ce : constElem) {
        const auto &qh = haloQH[(size_t)ce.first];
        c->hConstDst[(size_t)ce.second] = (long long)qh.first * c->stride + c->nPad + qh.second;
    }
    if (c->nFields == 2) {
        // phi slots: own nodes, solid-boundary nodes, ghost nodes (in neighbour / list order), one zero slot
        std::vector<int32_t> slotOf(c->nNodes, -1);
        for (int b = 0; b < nBulk; ++b) slotOf[c->bulk[b]] = devOf[b];
        c->nSolid = (int)c->solidBnd.size();
        for (int k = 0; k < c->nSolid; ++k) {
            const int node = c->solidBnd[k];
            if (node <= 0 || node >= c->nNodes) return fail("solid boundary node %d out of range", node);
            if (slotOf[node] < 0) slotOf[node] = c->nPad + k;
        }
        int g = 0;
        for (auto &nb : c->nbrs)
            for (int node : nb.recvNodes) {
                if (slotOf[node] < 0) slotOf[node] = c->nPad + c->nSolid + g;
                ++g;
            }
        c->nGhost = g;
        const int zeroSlot = c->nPad + c->nSolid + c->nGhost;
        c->nPhi = zeroSlot + 1;
        // MonLatMpi::communicateScalarField (LBmonlatmpi.h:181-205): field(nodesToSend) -> field(nodesReceived)
        c->hPhiSendSrc.resize(c->nbrs.size());
        c->hPhiRecvDst.resize(c->nbrs.size());
        for (size_t k = 0; k < c->nbrs.size(); ++k) {
            for (int node : c->nbrs[k].sendNodes) {
                if (node < 0 || node >= c->nNodes || bulkIndex[node] < 0) return fail("scalar halo: node %d to send is not an own fluid node", node);
                c->hPhiSendSrc[k].push_back(devOf[bulkIndex[node]]);
            }
            for (int node : c->nbrs[k].recvNodes) c->hPhiRecvDst[k].push_back(slotOf[node]);
        }
        c->hPtable.assign((size_t)nQ * c->nPad, zeroSlot);
        for (int i = 0; i < nSlots; ++i) {
            const size_t row = (size_t)c->bulk[order[i]] * nQ;
            for (int q = 0; q < nQ; ++q) {
                const int s = slotOf[c->neigh[row + q]];
                c->hPtable[(size_t)q * c->nPad + i] = s >= 0 ? s : zeroSlot;
            }
        }
    }
    c->hTable.swap(table);
    c->hLabel.swap(label);
    c->hPmask.swap(pmask);
    // release host inputs
    std::vector<int32_t>().swap(c->neigh);
    std::vector<Op>().swap(c->ops);
    c->hostBuilt = true;
    return 0;
}

int chimp_finalize(chimp_lattice *c, int index_form, int boundary_first)
{
    if (check(c, false)) return 1;
    if (index_form != CHIMP_INDEX_TABLE && index_form != CHIMP_INDEX_COMPACT) return fail("unknown index form %d", index_form);
    if (!c->hostBuilt && chimp_build_host(c, boundary_first)) return 1;
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
        return fail("no CUDA device available: this engine has no CPU fallback");
    if (c->device < 0) CUDA_OK(cudaGetDevice(&c->device));
    CUDA_OK(cudaSetDevice(c->device));
    if (setupStreams(c)) return 1;
    const std::vector<int32_t> &table = c->hTable, &label = c->hLabel;
    const std::vector<uint32_t> &pmask = c->hPmask;
    CUDA_OK(cudaMalloc(&c->d_table, table.size() * sizeof(int32_t)));
    CUDA_OK(cudaMemcpy(c->d_table, table.data(), table.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&c->d_label, label.size() * sizeof(int32_t)));
    CUDA_OK(cudaMemcpy(c->d_label, label.data(), label.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&c->d_pmask, pmask.size() * sizeof(uint32_t)));
    CUDA_OK(cudaMemcpy(c->d_pmask, pmask.data(), pmask.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    for (size_t k = 0; k < c->nbrs.size(); ++k) {
        Neighbor &nb = c->nbrs[k];
        const std::vector<long long> &src = c->hSendSrc[k], &dst = c->hRecvDst[k];
        nb.hSrc = src;
        if (!src.empty()) {
            CUDA_OK(cudaMalloc(&nb.d_sendSrc, src.size() * sizeof(long long)));
            CUDA_OK(cudaMemcpy(nb.d_sendSrc, src.data(), src.size() * sizeof(long long), cudaMemcpyHostToDevice));
            CUDA_OK(cudaMalloc(&nb.d_sendBuf, src.size() * c->nFields * sizeof(double)));
        }
        if (!dst.empty()) {
            CUDA_OK(cudaMalloc(&nb.d_recvDst, dst.size() * sizeof(long long)));
            CUDA_OK(cudaMemcpy(nb.d_recvDst, dst.data(), dst.size() * sizeof(long long), cudaMemcpyHostToDevice));
            CUDA_OK(cudaMalloc(&nb.d_recvBuf, dst.size() * c->nFields * sizeof(double)));
        }
    }
    if (c->nFields == 2) {
        for (size_t k = 0; k < c->nbrs.size(); ++k) {
            Neighbor &nb = c->nbrs[k];
            const auto &src = c->hPhiSendSrc[k], &dst = c->hPhiRecvDst[k];
            nb.phiSendCount = (long long)src.size();
            nb.phiRecvCount = (long long)dst.size();
            if (!src.empty()) {
                CUDA_OK(cudaMalloc(&nb.d_phiSendSrc, src.size() * sizeof(long long)));
                CUDA_OK(cudaMemcpy(nb.d_phiSendSrc, src.data(), src.size() * sizeof(long long), cudaMemcpyHostToDevice));
                CUDA_OK(cudaMalloc(&nb.d_phiSendBuf, src.size() * sizeof(double)));
            }
            if (!dst.empty()) {
                CUDA_OK(cudaMalloc(&nb.d_phiRecvDst, dst.size() * sizeof(long long)));
                CUDA_OK(cudaMemcpy(nb.d_phiRecvDst, dst.data(), dst.size() * sizeof(long long), cudaMemcpyHostToDevice));
                CUDA_OK(cudaMalloc(&nb.d_phiRecvBuf, dst.size() * sizeof(double)));
            }
        }
        CUDA_OK(cudaMalloc(&c->d_ptable, c->hPtable.size() * sizeof(int32_t)));
        CUDA_OK(cudaMemcpy(c->d_ptable, c->hPtable.data(), c->hPtable.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&c->d_phi, (size_t)c->nPhi * sizeof(double)));
        CUDA_OK(cudaMemset(c->d_phi, 0, (size_t)c->nPhi * sizeof(double)));
        const int nBlocks = (c->n + 255) / 256;
        CUDA_OK(cudaMalloc(&c->d_fluxPartial, (size_t)std::max(nBlocks, 1) * sizeof(double)));
        CUDA_OK(cudaMalloc(&c->d_fluxSum, sizeof(double)));
        CUDA_OK(cudaMalloc(&c->d_forceX, 4 * sizeof(double)));
        CUDA_OK(cudaMemset(c->d_forceX, 0, 4 * sizeof(double)));
    }
    c->indexForm = index_form;
    if (checkIndexRange(c)) return 1;
    if (index_form == CHIMP_INDEX_COMPACT && buildRankIndex(c)) return 1;
    if (index_form == CHIMP_INDEX_TABLE && buildKernelTable(c)) return 1;
    if (allocateState(c)) return 1;
    if (!c->hConstDst.empty()) {
        for (long long d : c->hConstDst)
            if (d < 0) return fail("internal error: a constant link has no slot");
        CUDA_OK(cudaMalloc(&c->d_constVal, c->constValues.size() * sizeof(double)));
        CUDA_OK(cudaMemcpy(c->d_constVal, c->constValues.data(), c->constValues.size() * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&c->d_constDst, c->hConstDst.size() * sizeof(long long)));
        CUDA_OK(cudaMemcpy(c->d_constDst, c->hConstDst.data(), c->hConstDst.size() * sizeof(long long), cudaMemcpyHostToDevice));
        restoreConstants(c, c->d_f[0]);
        restoreConstants(c, c->d_f[1]);
        CUDA_OK(cudaGetLastError());
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (labelRangeOf(c)) return 1; // row range of the host arrays, so that the first transfer does not pay for it
    if (c->d_ptable && buildPhiExceptions(c)) return 1;
    c->finalized = true;
    return 0;
}

int chimp_create_from_device_table(chimp_lattice **out, int lattice, int n_bulk, int n_pad, int n_halo,
                                   const int32_t *table_dev, const int32_t *label_dev, int n_fields,
                                   int index_form, int device)
{
    if (!out) return fail("out is null");
    *out = nullptr;
    const LatInfo li = latInfo(lattice);
    if (li.nQ == 0) return fail("unknown lattice id %d", lattice);
    if (n_fields < 1 || n_fields > 2) return fail("n_fields must be 1 or 2");
    if (n_pad % 32 || n_pad < n_bulk || n_halo % 16) return fail("n_pad must be a multiple of 32 >= n_bulk, n_halo a multiple of 16");
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
        return fail("no CUDA device available: this engine has no CPU fallback");
    if (device < 0) CUDA_OK(cudaGetDevice(&device));
    CUDA_OK(cudaSetDevice(device));
    chimp_lattice *c = new chimp_lattice;
    c->lattice = lattice;
    c->li = li;
    c->device = device;
    c->nFields = n_fields;
    if (setupStreams(c)) { delete c; return 1; }
    c->n = n_bulk;
    c->nNodes = n_bulk + 1; // labels 1..N are the leading rows of the reference's LbField (fluid nodes come first, vtklb.py:92-94)
    c->nPad = n_pad;
    c->nHalo = n_halo;
    c->stride = (long long)n_pad + n_halo;
    // the tables may still be in flight on the producer's stream (e.g. torch's): wait for the device once
    CUDA_OK(cudaDeviceSynchronize());
    const size_t tb = (size_t)li.nQ * n_pad * sizeof(int32_t);
    CUDA_OK(cudaMalloc(&c->d_table, tb));
    CUDA_OK(cudaMemcpyAsync(c->d_table, table_dev, tb, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(cudaMalloc(&c->d_label, (size_t)n_pad * sizeof(int32_t)));
    CUDA_OK(cudaMemcpyAsync(c->d_label, label_dev, (size_t)n_pad * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    c->hostBuilt = true;
    c->indexForm = index_form;
    if (checkIndexRange(c)) { chimp_destroy(c); return 1; }
    if (index_form == CHIMP_INDEX_COMPACT && buildRankIndex(c)) { chimp_destroy(c); return 1; }
    if (index_form == CHIMP_INDEX_TABLE && buildKernelTable(c)) { chimp_destroy(c); return 1; }
    if (allocateState(c)) { chimp_destroy(c); return 1; }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (labelRangeOf(c)) { chimp_destroy(c); return 1; }
    c->finalized = true;
    *out = c;
    return 0;
}

void chimp_destroy(chimp_lattice *c)
{
    if (!c) return;
    if (c->device >= 0) cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    freeDev(c->d_table); freeDev(c->d_ktable); freeDev(c->d_label); freeDev(c->d_delta); freeDev(c->d_pmask); freeDev(c->d_attr);
    freeDev(c->d_base); freeDev(c->d_rows); freeDev(c->d_f[0]); freeDev(c->d_f[1]);
    freeDev(c->d_rho); freeDev(c->d_vel); freeDev(c->d_flags); freeDev(c->d_slotOf);
    freeDev(c->d_mail); freeDev(c->d_peerMail);
    freeDev(c->d_sendMask); freeDev(c->d_sendDst); freeDev(c->d_extraStart); freeDev(c->d_extra);
    freeDev(c->d_peerCounter); freeDev(c->d_trace);
    if (c->h_error) cudaFreeHost(c->h_error);
    freeDev(c->d_ptable); freeDev(c->d_phi); freeDev(c->d_fluxPartial); freeDev(c->d_fluxSum); freeDev(c->d_forceX);
    freeDev(c->d_excInfo); freeDev(c->d_exc);
    freeDev(c->d_constVal); freeDev(c->d_constDst);
    freeDev(c->d_forceOn); freeDev(c->d_addSource); freeDev(c->d_srcPerLabel); freeDev(c->d_massPartial);
    freeDev(c->d_labelAttr); freeDev(c->d_scale); freeDev(c->d_mass);
    for (auto &nb : c->nbrs) {
        freeDev(nb.d_peerDst); freeDev(nb.d_blockCounter); freeDev(nb.d_peerPhiDst); freeDev(nb.d_phiBlockCounter);
        freeDev(nb.d_sendSrc); freeDev(nb.d_recvDst); freeDev(nb.d_sendBuf); freeDev(nb.d_recvBuf);
        freeDev(nb.d_phiSendSrc); freeDev(nb.d_phiRecvDst);
        if (nb.ownPhiBufs) { freeDev(nb.d_phiSendBuf); freeDev(nb.d_phiRecvBuf); }
    }
    if (c->evBoundary) cudaEventDestroy(c->evBoundary);
    if (c->evHalo) cudaEventDestroy(c->evHalo);
    if (c->evStep) cudaEventDestroy(c->evStep);
    if (c->haloStream) cudaStreamDestroy(c->haloStream);
    if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// ---- state transfer ---------------------------------------------------------------------
// Own bulk rows of a host array span the label range [labelMin, labelMax] (vtklb numbers a rank's fluid
// nodes 1..N first, so the range is normally exactly the bulk rows).  Transfers touch only that row
// range; when it holds nothing but bulk rows (labelsContiguous) a download is a pure device->host copy,
// otherwise the caller's rows are staged first so that non-bulk rows keep their content.
static int labelRange(chimp_lattice *c) { return labelRangeOf(c); }

extern "C++" {
namespace {
int labelRangeOf(chimp_lattice *c)
{
    if (c->labelMax >= 0) return 0;
    std::vector<int32_t> lab(c->n);
    CUDA_OK(cudaMemcpy(lab.data(), c->d_label, (size_t)c->n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    int lo = INT32_MAX, hi = -1;
    for (int v : lab) { lo = std::min(lo, v); hi = std::max(hi, v); }
    if (c->n == 0) { lo = 0; hi = 0; }
    c->labelMin = lo;
    c->labelMax = hi;
    c->labelsContiguous = (long long)hi - lo + 1 == c->n;
    return 0;
}
} // namespace
} // extern "C++"

// Staging area for layout conversion: the population buffer that is NOT current holds the state of the previous
// step, which nothing reads any more (the next step overwrites it), so transfers need no allocation.  Only when the
// caller's row range is larger than that buffer (labels far from contiguous) a temporary allocation is made.
struct Staging {
    double *ptr = nullptr;
    bool owned = false;
    ~Staging() { if (owned) cudaFree(ptr); }
};
static int acquireStaging(chimp_lattice *c, size_t bytes, Staging &st)
{
    const size_t have = (size_t)c->nFields * c->li.nQ * (size_t)c->stride * sizeof(double);
    // not with peer halos: the neighbours may already be storing the next step's populations into that buffer;
    // not with constant links: their slots of that buffer keep their values from finalize on
    if (bytes <= have && !c->peerHalos && c->hConstDst.empty()) {
        st.ptr = c->d_f[c->cur ^ 1];
        return 0;
    }
    CUDA_OK(cudaMalloc(&st.ptr, bytes));
    st.owned = true;
    return 0;
}

int chimp_upload_lbfield(chimp_lattice *c, const double *f_aos)
{
    if (check(c, true)) return 1;
    if (c->nNodes <= 0) return fail("upload in reference layout needs a lattice created from reference tables");
    CUDA_OK(cudaSetDevice(c->device));
    if (labelRange(c)) return 1;
    const size_t rowDoubles = (size_t)c->nFields * c->li.nQ;
    const size_t rows = (size_t)(c->labelMax - c->labelMin + 1);
    const size_t bytes = rows * rowDoubles * sizeof(double);
    Staging st;
    if (acquireStaging(c, bytes, st)) return 1;
    double *d_aos = st.ptr;
    awaitPeers(c); // a neighbour's stores of the last step must not land on top of the uploaded halo-in slots
    CUDA_OK(cudaMemcpyAsync(d_aos, f_aos + (size_t)c->labelMin * rowDoubles, bytes, cudaMemcpyHostToDevice, c->stream));
    const unsigned grid = (unsigned)((c->n + 255) / 256);
    double *X = c->d_f[c->cur];
    const double *shifted = d_aos - (size_t)c->labelMin * rowDoubles; // kernels index rows by label
    switch (c->lattice) {
    case CHIMP_D2Q9: scatterStateKernel<D2Q9><<<grid, 256, 0, c->stream>>>(shifted, X, c->d_table, c->d_label, c->n, c->nPad, c->stride, c->nFields); break;
    case CHIMP_D3Q19: scatterStateKernel<D3Q19><<<grid, 256, 0, c->stream>>>(shifted, X, c->d_table, c->d_label, c->n, c->nPad, c->stride, c->nFields); break;
    case CHIMP_D3Q27: scatterStateKernel<D3Q27><<<grid, 256, 0, c->stream>>>(shifted, X, c->d_table, c->d_label, c->n, c->nPad, c->stride, c->nFields); break;
    }
    ++g_launches;
    c->constDirty = !c->hConstDst.empty(); // the constant links' slots hold uploaded values until the first step has read them
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return checkDeviceError(c);
}

int chimp_download_lbfield(chimp_lattice *c, double *f_aos)
{
    if (check(c, true)) return 1;
    if (c->nNodes <= 0) return fail("download in reference layout needs a lattice created from reference tables");
    CUDA_OK(cudaSetDevice(c->device));
    if (labelRange(c)) return 1;
    const size_t rowDoubles = (size_t)c->nFields * c->li.nQ;
    const size_t rows = (size_t)(c->labelMax - c->labelMin + 1);
    const size_t bytes = rows * rowDoubles * sizeof(double);
    double *host = f_aos + (size_t)c->labelMin * rowDoubles;
    Staging st;
    if (acquireStaging(c, bytes, st)) return 1;
    double *d_aos = st.ptr;
    if (!c->labelsContiguous) CUDA_OK(cudaMemcpyAsync(d_aos, host, bytes, cudaMemcpyHostToDevice, c->stream));
    awaitPeers(c); // the gather reads the halo-in slots the neighbours filled during the last step
    const unsigned grid = (unsigned)((c->n + 255) / 256);
    const double *X = c->d_f[c->cur];
    double *shifted = d_aos - (size_t)c->labelMin * rowDoubles;
    switch (c->lattice) {
    case CHIMP_D2Q9: gatherStateKernel<D2Q9><<<grid, 256, 0, c->stream>>>(shifted, X, c->d_table, c->d_label, c->n, c->nPad, c->stride, c->nFields); break;
    case CHIMP_D3Q19: gatherStateKernel<D3Q19><<<grid, 256, 0, c->stream>>>(shifted, X, c->d_table, c->d_label, c->n, c->nPad, c->stride, c->nFields); break;
    case CHIMP_D3Q27: gatherStateKernel<D3Q27><<<grid, 256, 0, c->stream>>>(shifted, X, c->d_table, c->d_label, c->n, c->nPad, c->stride, c->nFields); break;
    }
    ++g_launches;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(host, d_aos, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return checkDeviceError(c);
}

static int downloadPlanes(chimp_lattice *c, double *host, const double *planes, int nComp, int aosStride, int aosOffset)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (labelRange(c)) return 1;
    const size_t rows = (size_t)(c->labelMax - c->labelMin + 1);
    const size_t bytes = rows * aosStride * sizeof(double);
    double *hostRows = host + (size_t)c->labelMin * aosStride;
    Staging st;
    if (acquireStaging(c, bytes, st)) return 1;
    double *d_aos = st.ptr;
    // a full overwrite of the row range needs no staging; partial rows (aosStride > nComp) or
    // non-bulk rows inside the range keep the caller's content
    if (!c->labelsContiguous || nComp != aosStride) CUDA_OK(cudaMemcpyAsync(d_aos, hostRows, bytes, cudaMemcpyHostToDevice, c->stream));
    const unsigned grid = (unsigned)((c->n + 255) / 256);
    planesToAosKernel<<<grid, 256, 0, c->stream>>>(d_aos - (size_t)c->labelMin * aosStride, planes, c->d_label, c->n, c->nPad, nComp, aosStride, aosOffset);
    ++g_launches;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(hostRows, d_aos, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return checkDeviceError(c);
}

int chimp_download_rho(chimp_lattice *c, double *rho_sca, int n_fields_host)
{
    if (check(c, true)) return 1;
    if (c->nNodes <= 0) return fail("needs a lattice created from reference tables");
    if (n_fields_host < c->nFields) return fail("host ScalarField has fewer fields than the lattice");
    for (int f = 0; f < c->nFields; ++f)
        if (downloadPlanes(c, rho_sca, c->d_rho + (size_t)f * c->nPad, 1, n_fields_host, f)) return 1;
    return 0;
}

int chimp_download_vel(chimp_lattice *c, double *vel_vec)
{
    if (check(c, true)) return 1;
    if (c->nNodes <= 0) return fail("needs a lattice created from reference tables");
    return downloadPlanes(c, vel_vec, c->d_vel, c->li.nD, c->li.nD, 0);
}

int chimp_set_one_phase_attributes(chimp_lattice *c, const double *force_on, const int32_t *interior_label,
                                   const double *add_mass_source, int n_labels, const double *scale_per_label,
                                   double rho_w)
{
    if (check(c, true)) return 1;
    if (c->nNodes <= 0) return fail("needs a lattice created from reference tables");
    if (n_labels < 1) return fail("n_labels must be >= 1");
    CUDA_OK(cudaSetDevice(c->device));
    std::vector<int32_t> label(c->nPad);
    CUDA_OK(cudaMemcpy(label.data(), c->d_label, (size_t)c->nPad * sizeof(int32_t), cudaMemcpyDeviceToHost));
    std::vector<double> on(c->nPad, 0.0), add(c->nPad, 0.0);
    std::vector<int32_t> lab(c->nPad, 0);
    for (int i = 0; i < c->n; ++i) {
        const int l = label[i];
        if (l <= 0) continue;
        on[i] = force_on[l];
        add[i] = add_mass_source[l];
        lab[i] = interior_label[l];
        if (lab[i] < 0 || lab[i] >= n_labels) return fail("interior label %d of node %d outside [0,%d)", lab[i], l, n_labels);
    }
    freeDev(c->d_forceOn); freeDev(c->d_addSource); freeDev(c->d_labelAttr); freeDev(c->d_srcPerLabel);
    CUDA_OK(cudaMalloc(&c->d_forceOn, on.size() * sizeof(double)));
    CUDA_OK(cudaMemcpy(c->d_forceOn, on.data(), on.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&c->d_addSource, add.size() * sizeof(double)));
    CUDA_OK(cudaMemcpy(c->d_addSource, add.data(), add.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&c->d_labelAttr, lab.size() * sizeof(int32_t)));
    CUDA_OK(cudaMemcpy(c->d_labelAttr, lab.data(), lab.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&c->d_srcPerLabel, (size_t)n_labels * sizeof(double)));
    CUDA_OK(cudaMemset(c->d_srcPerLabel, 0, (size_t)n_labels * sizeof(double)));
    freeDev(c->d_scale); freeDev(c->d_mass); freeDev(c->d_massPartial);
    CUDA_OK(cudaMalloc(&c->d_scale, (size_t)n_labels * sizeof(double)));
    CUDA_OK(cudaMemcpy(c->d_scale, scale_per_label, (size_t)n_labels * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&c->d_mass, (size_t)n_labels * sizeof(double)));
    CUDA_OK(cudaMemset(c->d_mass, 0, (size_t)n_labels * sizeof(double)));
    c->nLabels = n_labels;
    c->scalePerLabel.assign(scale_per_label, scale_per_label + n_labels);
    c->rhoW = rho_w;
    c->onePhase = true;
    // One word per node instead of four arrays (24 -> 4 bytes per node and step) when the switches are exactly 0.0 / 1.0
    // (they are integers in the reference's input, std_one_phase/main.cpp:337-346), at most 16 labels, and the link
    // mask fits: bit 0 forceOn, bit 1 addSource, bits 2-5 label, bits 6.. pmask (nQ - 1 <= 26 directions).
    freeDev(c->d_attr);
    bool packs = c->attrPackEnv && n_labels <= 16 && c->li.nQ - 1 <= 26;
    for (int i = 0; i < c->n && packs; ++i) packs = (on[i] == 0.0 || on[i] == 1.0) && (add[i] == 0.0 || add[i] == 1.0);
    if (packs) {
        std::vector<uint32_t> pmask(c->nPad, 0u);
        if (c->hasPressure && c->d_pmask) CUDA_OK(cudaMemcpy(pmask.data(), c->d_pmask, (size_t)c->nPad * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        std::vector<uint32_t> attr(c->nPad, 0u);
        for (int i = 0; i < c->n && packs; ++i) {
            if (pmask[i] >> 26) packs = false; // a link bit beyond direction 25 (cannot happen: the rest direction has no link)
            attr[i] = (on[i] == 1.0 ? 1u : 0u) | (add[i] == 1.0 ? 2u : 0u) | ((uint32_t)lab[i] << 2) | (pmask[i] << 6);
        }
        if (packs) {
            CUDA_OK(cudaMalloc(&c->d_attr, attr.size() * sizeof(uint32_t)));
            CUDA_OK(cudaMemcpy(c->d_attr, attr.data(), attr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
    }
    return 0;
}

// ---- stepping -----------------------------------------------------------------------------
namespace {

void fillIndexView(const chimp_lattice *c, IndexView &v)
{
    v.table = c->d_ktable;
    v.delta = c->d_delta;
    v.base = c->d_base;
    v.rows = c->d_rows;
    v.nTiles = c->nTiles;
    for (int q = 0; q < c->li.nQ; ++q) v.bounceOff[q] = (int)((long long)(revDir(c->li, q) - q) * c->stride);
}

void fillPlanes(const chimp_lattice *c, Planes &pl)
{
    const double *in = c->d_f[c->cur];
    double *out = c->d_f[c->cur ^ 1];
    for (int q = 0; q < c->li.nQ; ++q) {
        pl.in[q] = in + (long long)q * c->stride;
        pl.out[q] = out + (long long)q * c->stride;
    }
}

int fillStepArgs(chimp_lattice *c, const chimp_single_params *p, StepArgs &a)
{
    if (!p) return fail("params is null");
    if (c->nFields != 1) return fail("chimp_step_single needs a one-field lattice");
    if (p->collision != CHIMP_BGK && p->collision != CHIMP_TRT) return fail("unknown collision %d", p->collision);
    if (c->hasPressure && !c->onePhase) return fail("pressure links need chimp_set_one_phase_attributes");
    a = StepArgs{};
    a.stride = c->stride;
    a.n = c->n;
    a.nPad = c->nPad;
    fillIndexView(c, a.idx);
    // LBcollision.h:39,65-66,91,113-114,206,228-229: per-call constants of the reference helpers
    a.tauInv = 1.0 / p->tau;
    a.tauFactor = (1 - 0.5 / p->tau);
    if (p->collision == CHIMP_TRT) {
        a.tauSymInv = 1.0 / p->tau_sym;
        a.tauAntiInv = 1.0 / p->tau_anti;
        a.symFactor = (1 - 0.5 / p->tau_sym);
        a.antiFactor = (1 - 0.5 / p->tau_anti);
    }
    for (int d = 0; d < 3; ++d) a.F[d] = d < c->li.nD ? p->force[d] : 0.0;
    a.forceOn = c->d_forceOn;
    a.addSource = c->d_addSource;
    a.label = c->d_labelAttr;
    a.srcPerLabel = c->d_srcPerLabel;
    a.pmask = c->hasPressure ? c->d_pmask : nullptr;
    a.attr = c->d_attr;
    a.rhoW = c->rhoW;
    a.rho = c->d_rho;
    a.vel = c->d_vel;
    return 0;
}

void packHalos(chimp_lattice *c, const double *X)
{
    const long long fieldStride = (long long)c->li.nQ * c->stride;
    for (auto &nb : c->nbrs)
        for (int f = 0; f < c->nFields && nb.sendCount; ++f) {
            haloPackKernel<<<(unsigned)((nb.sendCount + 255) / 256), 256, 0, c->haloStream>>>(nb.sendBuf() + f * nb.sendCount, X + f * fieldStride, nb.d_sendSrc, (int)nb.sendCount);
            ++g_launches;
        }
}

void unpackHalos(chimp_lattice *c, double *X)
{
    const long long fieldStride = (long long)c->li.nQ * c->stride;
    for (auto &nb : c->nbrs)
        for (int f = 0; f < c->nFields && nb.recvCount; ++f) {
            haloUnpackKernel<<<(unsigned)((nb.recvCount + 255) / 256), 256, 0, c->haloStream>>>(X + f * fieldStride, nb.recvBuf() + f * nb.recvCount, nb.d_recvDst, (int)nb.recvCount);
            ++g_launches;
        }
}

// first half of an iteration: collide + stream of all own nodes, halo-coupled nodes first, and
// packing of the populations the neighbour ranks will pull
int stepBegin(chimp_lattice *c, const chimp_single_params *p, bool mom, bool callExchange)
{
    StepArgs a;
    if (fillStepArgs(c, p, a)) return 1;
    fillPlanes(c, a.pl);
    double *const foutBuf = c->d_f[c->cur ^ 1];
    if (c->onePhase && c->nLabels > 1) {
        awaitPeers(c); // the pass gathers through the halo-in slots
        if (massChangePass(c, a)) return 1;
    }
    if (c->nbrs.empty()) {
        a.begin = 0;
        a.end = c->n;
        dispatchSingleLattice(c, a, p->collision, mom, c->stream);
        return 0;
    }
    if (c->peerHalos && c->peerFused) {
        // one launch per step: the halo-coupled nodes occupy the first blocks; those blocks wait for the neighbours'
        // arrival counters, store the outgoing populations into the neighbours' halo-in slots and publish this step
        PeerView &pv = a.peer;
        pv.blocks = c->peerBlocks;
        pv.nFaces = (int)c->nbrs.size();
        pv.pad = c->peerPad;
        pv.mask = c->d_sendMask;
        pv.dst = c->d_sendDst;
        pv.extraStart = c->d_extraStart;
        pv.extra = c->d_extra;
        const int outIdx = c->cur ^ 1;
        for (int k = 0; k < pv.nFaces; ++k) {
            const Neighbor &nb = c->nbrs[k];
            pv.out[k] = nb.peerX[outIdx];
            pv.stride[k] = nb.peerFieldStride / c->li.nQ;
            pv.flagOut[k] = nb.peerFlags + nb.peerFace;
        }
        pv.flagIn = c->d_flags;
        pv.expect = (unsigned long long)c->steps;
        pv.publish = (unsigned long long)c->steps + 1;
        pv.counter = c->d_peerCounter;
        pv.error = c->d_error;
        pv.timeoutNs = c->timeoutNs;
        pv.trace = c->traceCursor;
        a.begin = 0;
        a.end = c->n;
        dispatchSingleLattice(c, a, p->collision, mom, c->stream);
        return 0;
    }
    // The halo-coupled nodes (first slots) are stepped and packed on the high-priority halo stream while
    // the interior nodes run on the main stream: both read the old buffer and write disjoint slots of the
    // new one.  evStep orders the halo stream behind everything the main stream did up to here.
    CUDA_OK(cudaEventRecord(c->evStep, c->stream));
    CUDA_OK(cudaStreamWaitEvent(c->haloStream, c->evStep, 0));
    if (c->peerHalos) {
        // the halo-in slots of the buffer read now were stored by the neighbours during their previous
        // step: wait until every face has reported that step (the flag counts completed pushes)
        waitFlagsKernel<<<1, 32, 0, c->haloStream>>>(c->d_flags, (1u << c->nbrs.size()) - 1u, (unsigned long long)c->steps, c->timeoutNs, c->d_error);
        ++g_launches;
    }
    a.begin = 0;
    a.end = c->nBoundary ? c->nBoundary : c->n;
    dispatchSingleLattice(c, a, p->collision, mom, c->haloStream);
    if (c->peerHalos) {
        const long long fieldStride = (long long)c->li.nQ * c->stride;
        const int outIdx = c->cur ^ 1;
        for (auto &nb : c->nbrs)
            if (nb.sendCount) {
                haloPushKernel<<<(unsigned)((nb.sendCount + 255) / 256), 256, 0, c->haloStream>>>(
                    nb.peerX[outIdx], foutBuf, nb.d_sendSrc, nb.d_peerDst, (int)nb.sendCount, c->nFields, fieldStride,
                    nb.peerFieldStride, nb.d_blockCounter, nb.peerFlags + nb.peerFace, (unsigned long long)(c->steps + 1));
                ++g_launches;
            }
        if (c->nBoundary && c->nBoundary < c->n) {
            a.begin = c->nBoundary;
            a.end = c->n;
            dispatchSingleLattice(c, a, p->collision, mom, c->stream);
        }
        return 0;
    }
    packHalos(c, foutBuf);
    // the transport is enqueued before the interior launch so that it is ahead of it in issue order
    if (callExchange && c->exchange && c->exchange(c->exchangeUser, (void *)c->haloStream)) return fail("exchange callback failed");
    if (c->nBoundary && c->nBoundary < c->n) {
        a.begin = c->nBoundary;
        a.end = c->n;
        dispatchSingleLattice(c, a, p->collision, mom, c->stream);
    }
    return 0;
}

// second half: incoming populations go to the halo-in slots of the new buffer; buffers swap
int stepEnd(chimp_lattice *c)
{
    double *fout = c->d_f[c->cur ^ 1];
    if (!c->nbrs.empty() && !(c->peerHalos && c->peerFused)) {
        if (!c->peerHalos) unpackHalos(c, fout);
        CUDA_OK(cudaEventRecord(c->evHalo, c->haloStream));
        CUDA_OK(cudaStreamWaitEvent(c->stream, c->evHalo, 0));
    }
    if (c->constDirty) {
        // An upload or initialisation put ordinary values into the constant links' slots of the buffer this step has
        // read -- the reference's field before the first apply() (LBpressurebnd.h:19-41).  From now on they hold the
        // constants again (the other buffer's were never touched).
        restoreConstants(c, c->d_f[c->cur]);
        c->constDirty = false;
    }
    c->cur ^= 1;
    ++c->steps;
    return 0;
}

} // namespace

int chimp_step_begin(chimp_lattice *c, const chimp_single_params *p, int store_moments)
{
    if (check(c, true)) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    if (stepBegin(c, p, store_moments != 0, false)) return 1;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int chimp_step_end(chimp_lattice *c)
{
    if (check(c, true)) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    if (stepEnd(c)) return 1;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int chimp_step_single(chimp_lattice *c, const chimp_single_params *p, int n_steps)
{
    if (check(c, true)) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    // CHIMP_TRACE=1 with the fused peer step: device timestamps per step (first block start, arrival counters
    // published, longest counter wait, last block end), printed per rank as averages over the call
    const bool trace = c->trace && c->peerHalos && c->peerFused && n_steps > 0;
    if (trace) {
        if (c->traceCap < n_steps) {
            freeDev(c->d_trace);
            CUDA_OK(cudaMalloc(&c->d_trace, (size_t)n_steps * 4 * sizeof(unsigned long long)));
            c->traceCap = n_steps;
        }
        CUDA_OK(cudaMemsetAsync(c->d_trace, 0, (size_t)n_steps * 4 * sizeof(unsigned long long), c->stream));
    }
    for (int s = 0; s < n_steps; ++s) {
        c->traceCursor = trace ? c->d_trace + 4 * (size_t)s : nullptr;
        if (stepBegin(c, p, s == n_steps - 1, true)) return 1;
        if (stepEnd(c)) return 1;
    }
    c->traceCursor = nullptr;
    CUDA_OK(cudaGetLastError());
    if (trace) {
        std::vector<unsigned long long> t((size_t)n_steps * 4);
        CUDA_OK(cudaMemcpyAsync(t.data(), c->d_trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        double pub = 0, run = 0, wait = 0, gap = 0, waitMax = 0;
        for (int s = 0; s < n_steps; ++s) {
            const unsigned long long *e = &t[4 * (size_t)s];
            pub += (double)(e[1] - e[0]);
            run += (double)(e[3] - e[0]);
            wait += (double)e[2];
            waitMax = std::max(waitMax, (double)e[2]);
            if (s) gap += (double)((long long)e[0] - (long long)t[4 * (size_t)(s - 1) + 3]);
        }
        const double k = 1e-6 / n_steps;
        fprintf(stderr,
                "[chimp trace] rank %d: %d steps, %d own nodes (%d halo-coupled blocks of %d); ms/step: kernel %.4f | start -> halo published %.4f | "
                "longest counter wait %.4f (max %.4f) | gap between launches %.4f\n",
                c->worldRank, n_steps, c->n, c->peerBlocks, (c->n + CHIMP_BLOCK - 1) / CHIMP_BLOCK, run * k, pub * k, wait * k, waitMax * 1e-6,
                n_steps > 1 ? gap * 1e-6 / (n_steps - 1) : 0.0);
    }
    return 0;
}

int chimp_step_timed(chimp_lattice *c, const chimp_single_params *p, int n_steps, double *ms)
{
    if (check(c, true)) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    CUDA_OK(cudaEventRecord(e0, c->stream));
    const int rc = chimp_step_single(c, p, n_steps);
    CUDA_OK(cudaEventRecord(e1, c->stream));
    CUDA_OK(cudaEventSynchronize(e1));
    float t = 0.f;
    CUDA_OK(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = t;
    return rc;
}

int chimp_init_uniform(chimp_lattice *c, double rho)
{
    if (check(c, true)) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    std::vector<double> wq(c->li.nQ);
    for (int q = 0; q < c->li.nQ; ++q) wq[q] = chimp_lattice_w(c->lattice, q) * rho;
    for (int f = 0; f < c->nFields; ++f)
        for (int q = 0; q < c->li.nQ; ++q) {
            const long long count = c->stride;
            fillKernel<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(c->d_f[c->cur] + ((long long)f * c->li.nQ + q) * c->stride, wq[q], count);
            ++g_launches;
        }
    c->constDirty = !c->hConstDst.empty();
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int chimp_init_equilibrium_dev(chimp_lattice *c, const double *rho_dev)
{
    if (check(c, true)) return 1;
    if (!rho_dev) return fail("rho_dev is null");
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaDeviceSynchronize()); // rho may come from another stream
    const unsigned grid = (unsigned)((c->n + 255) / 256);
    double *X = c->d_f[c->cur];
    switch (c->lattice) {
    case CHIMP_D2Q9: initEquilibriumKernel<D2Q9><<<grid, 256, 0, c->stream>>>(rho_dev, X, c->d_table, c->n, c->nPad, c->stride, c->nFields); break;
    case CHIMP_D3Q19: initEquilibriumKernel<D3Q19><<<grid, 256, 0, c->stream>>>(rho_dev, X, c->d_table, c->n, c->nPad, c->stride, c->nFields); break;
    case CHIMP_D3Q27: initEquilibriumKernel<D3Q27><<<grid, 256, 0, c->stream>>>(rho_dev, X, c->d_table, c->n, c->nPad, c->stride, c->nFields); break;
    }
    ++g_launches;
    c->constDirty = !c->hConstDst.empty();
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C++" {
namespace {
template <class L>
void launchPhiExceptions(chimp_lattice *c, const TwoPhaseArgs &a, int32_t *counts, int32_t *exc)
{
    const unsigned grid = (unsigned)((c->n + 127) / 128);
    if (c->indexForm == CHIMP_INDEX_COMPACT) phiExceptionKernel<L, IDX_COMPACT><<<grid, 128, 0, c->stream>>>(a, counts, exc);
    else phiExceptionKernel<L, IDX_TABLE><<<grid, 128, 0, c->stream>>>(a, counts, exc);
    ++g_launches;
}
// Derived form of the phi table (TwoPhaseArgs::excInfo / exc): wherever ptable names the slot the pull index already
// yields -- neighbor(q, n) is an own fluid node, the pull source of direction rev(q) -- nothing is stored; the other
// links (solid boundary nodes, ghost nodes, the zero slot) are kept per node.  Built by comparing the two on the
// device, so whatever the caller's ptable says is what the collide pass sees; ptable itself stays for CHIMP_PHI_DERIVED=0.
int buildPhiExceptions(chimp_lattice *c)
{
    freeDev(c->d_excInfo);
    freeDev(c->d_exc);
    c->phiDerived = false;
    c->excWords = 0;
    if (!c->phiDerivedEnv || !c->d_ptable || c->lattice == CHIMP_D3Q27 || c->n <= 0) return 0;
    if (c->indexForm == CHIMP_INDEX_COMPACT ? !c->d_delta : !c->d_ktable) return 0; // index not built yet (finalize calls again)
    TwoPhaseArgs a{};
    a.stride = c->stride;
    a.n = c->n;
    a.nPad = c->nPad;
    fillIndexView(c, a.idx);
    a.ptable = c->d_ptable;
    int32_t *d_counts = nullptr;
    CUDA_OK(cudaMalloc(&d_counts, (size_t)c->nPad * sizeof(int32_t)));
    CUDA_OK(cudaMemsetAsync(d_counts, 0, (size_t)c->nPad * sizeof(int32_t), c->stream));
    if (c->lattice == CHIMP_D2Q9) launchPhiExceptions<D2Q9>(c, a, d_counts, nullptr);
    else launchPhiExceptions<D3Q19>(c, a, d_counts, nullptr);
    std::vector<int32_t> info((size_t)c->nPad, 0);
    cudaError_t e = cudaMemcpyAsync(info.data(), d_counts, (size_t)c->nPad * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { cudaFree(d_counts); return fail("phi exception count failed: %s", cudaGetErrorString(e)); }
    long long total = 0;
    for (int i = 0; i < c->n; ++i) {
        const int32_t words = info[(size_t)i];
        info[(size_t)i] = words ? (int32_t)(total + 1) : 0;
        total += words;
        if (total >= INT_MAX) { cudaFree(d_counts); return 0; } // offsets would not fit: keep the table form
    }
    // d_counts becomes excInfo
    e = cudaMemcpyAsync(d_counts, info.data(), (size_t)c->nPad * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) { cudaFree(d_counts); return fail("phi exception offsets failed: %s", cudaGetErrorString(e)); }
    c->d_excInfo = d_counts;
    CUDA_OK(cudaMalloc(&c->d_exc, (size_t)std::max(total, 1ll) * sizeof(int32_t)));
    CUDA_OK(cudaMemsetAsync(c->d_exc, 0, (size_t)std::max(total, 1ll) * sizeof(int32_t), c->stream));
    a.excInfo = c->d_excInfo;
    if (c->lattice == CHIMP_D2Q9) launchPhiExceptions<D2Q9>(c, a, nullptr, c->d_exc);
    else launchPhiExceptions<D3Q19>(c, a, nullptr, c->d_exc);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(c->stream)); // info (host vector) is read by the copy above
    c->excWords = total;
    c->phiDerived = true;
    return 0;
}
} // namespace
} // extern "C++"

int chimp_set_phi_table_dev(chimp_lattice *c, const int32_t *ptable_dev, int n_extra, const double *phi_extra_dev)
{
    if (check(c, true)) return 1;
    if (c->nFields != 2) return fail("needs a two-field lattice");
    if (!ptable_dev || n_extra < 0 || (n_extra > 0 && !phi_extra_dev)) return fail("bad phi table arguments");
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaDeviceSynchronize());
    freeDev(c->d_ptable); freeDev(c->d_phi); freeDev(c->d_fluxPartial); freeDev(c->d_fluxSum); freeDev(c->d_forceX);
    const size_t tb = (size_t)c->li.nQ * c->nPad * sizeof(int32_t);
    CUDA_OK(cudaMalloc(&c->d_ptable, tb));
    CUDA_OK(cudaMemcpy(c->d_ptable, ptable_dev, tb, cudaMemcpyDeviceToDevice));
    c->nSolid = n_extra;
    c->nGhost = 0;
    c->nPhi = c->nPad + n_extra + 1;
    CUDA_OK(cudaMalloc(&c->d_phi, (size_t)c->nPhi * sizeof(double)));
    CUDA_OK(cudaMemset(c->d_phi, 0, (size_t)c->nPhi * sizeof(double)));
    if (n_extra) CUDA_OK(cudaMemcpy(c->d_phi + c->nPad, phi_extra_dev, (size_t)n_extra * sizeof(double), cudaMemcpyDeviceToDevice));
    const int nBlocks = (c->n + 255) / 256;
    CUDA_OK(cudaMalloc(&c->d_fluxPartial, (size_t)std::max(nBlocks, 1) * sizeof(double)));
    CUDA_OK(cudaMalloc(&c->d_fluxSum, sizeof(double)));
    CUDA_OK(cudaMalloc(&c->d_forceX, 4 * sizeof(double)));
    CUDA_OK(cudaMemset(c->d_forceX, 0, 4 * sizeof(double)));
    if (buildPhiExceptions(c)) return 1;
    c->densitySet = true;
    return 0;
}

int chimp_download_moments_device_order(chimp_lattice *c, double *rho, double *vel)
{
    if (check(c, true)) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (checkDeviceError(c)) return 1;
    if (rho) CUDA_OK(cudaMemcpy(rho, c->d_rho, (size_t)c->n * sizeof(double), cudaMemcpyDeviceToHost));
    if (vel)
        for (int d = 0; d < c->li.nD; ++d)
            CUDA_OK(cudaMemcpy(vel + (size_t)d * c->n, c->d_vel + (size_t)d * c->nPad, (size_t)c->n * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int chimp_download_mass_change(chimp_lattice *c, double *mass_per_label)
{
    if (check(c, true)) return 1;
    if (!c->onePhase) return fail("one-phase attributes not set");
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaMemcpy(mass_per_label, c->d_mass, (size_t)c->nLabels * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int chimp_set_twophase_density(chimp_lattice *c, const double *rho2)
{
    if (check(c, true)) return 1;
    if (c->nFields != 2) return fail("needs a two-field lattice");
    if (!rho2) return fail("rho is null");
    CUDA_OK(cudaSetDevice(c->device));
    // main_TWOPHASE.cpp:280-284: colour of the solid boundary nodes from their (constant) densities
    std::vector<double> phiWall(std::max(c->nSolid, 1), 0.0);
    for (int k = 0; k < c->nSolid; ++k) {
        const double r0 = rho2[2 * (size_t)c->solidBnd[k]], r1 = rho2[2 * (size_t)c->solidBnd[k] + 1];
        phiWall[k] = (r0 - r1) / (r0 + r1);
    }
    if (c->nSolid) CUDA_OK(cudaMemcpy(c->d_phi + c->nPad, phiWall.data(), (size_t)c->nSolid * sizeof(double), cudaMemcpyHostToDevice));
    c->densitySet = true;
    return 0;
}

extern "C++" {
namespace {
template <class L>
void launchPhaseMoments(chimp_lattice *c, const TwoPhaseArgs &a, unsigned gridAll)
{
    if (c->indexForm == CHIMP_INDEX_COMPACT) phaseMomentsKernel<L, IDX_COMPACT><<<gridAll, 256, 0, c->stream>>>(a);
    else phaseMomentsKernel<L, IDX_TABLE><<<gridAll, 256, 0, c->stream>>>(a);
    ++g_launches;
}
template <class L>
void launchTwoPhaseCollide(chimp_lattice *c, const TwoPhaseArgs &a, bool mom, cudaStream_t st)
{
    if (a.end <= a.begin) return;
    const unsigned grid = (unsigned)((a.end - a.begin + CHIMP_TP_BLOCK - 1) / CHIMP_TP_BLOCK);
    if (c->phiDerived) {
        if (c->indexForm == CHIMP_INDEX_COMPACT) {
            if (mom) twoPhaseCollideKernel<L, true, IDX_COMPACT, true><<<grid, CHIMP_TP_BLOCK, 0, st>>>(a);
            else twoPhaseCollideKernel<L, false, IDX_COMPACT, true><<<grid, CHIMP_TP_BLOCK, 0, st>>>(a);
        } else {
            if (mom) twoPhaseCollideKernel<L, true, IDX_TABLE, true><<<grid, CHIMP_TP_BLOCK, 0, st>>>(a);
            else twoPhaseCollideKernel<L, false, IDX_TABLE, true><<<grid, CHIMP_TP_BLOCK, 0, st>>>(a);
        }
    } else if (c->indexForm == CHIMP_INDEX_COMPACT) {
        if (mom) twoPhaseCollideKernel<L, true, IDX_COMPACT><<<grid, CHIMP_TP_BLOCK, 0, st>>>(a);
        else twoPhaseCollideKernel<L, false, IDX_COMPACT><<<grid, CHIMP_TP_BLOCK, 0, st>>>(a);
    } else {
        if (mom) twoPhaseCollideKernel<L, true, IDX_TABLE><<<grid, CHIMP_TP_BLOCK, 0, st>>>(a);
        else twoPhaseCollideKernel<L, false, IDX_TABLE><<<grid, CHIMP_TP_BLOCK, 0, st>>>(a);
    }
    ++g_launches;
}
void phaseMoments(chimp_lattice *c, const TwoPhaseArgs &a, unsigned gridAll)
{
    if (c->lattice == CHIMP_D2Q9) launchPhaseMoments<D2Q9>(c, a, gridAll);
    else launchPhaseMoments<D3Q19>(c, a, gridAll);
}
void twoPhaseCollide(chimp_lattice *c, const TwoPhaseArgs &a, bool mom, cudaStream_t st = nullptr)
{
    if (!st) st = c->stream;
    if (c->lattice == CHIMP_D2Q9) launchTwoPhaseCollide<D2Q9>(c, a, mom, st);
    else launchTwoPhaseCollide<D3Q19>(c, a, mom, st);
}
} // namespace
} // extern "C++"

int chimp_step_twophase(chimp_lattice *c, const chimp_twophase_params *p, int n_steps)
{
    if (check(c, true)) return 1;
    if (!p) return fail("params is null");
    if (c->nFields != 2) return fail("chimp_step_twophase needs a two-field lattice");
    if (c->lattice == CHIMP_D3Q27) return fail("D3Q27 has no colour-gradient weights B[] (not defined by the reference)");
    if (!c->densitySet) return fail("call chimp_set_twophase_density first (wall colour, main_TWOPHASE.cpp:173-181)");
    const bool multi = !c->nbrs.empty();
    if (multi && !c->peerTwoPhase && (!c->exchange || !c->scalarExchange || !c->allreduce))
        return fail("N-rank twophase needs the exchange, scalar-exchange and allreduce callbacks (or peer connections)");
    if (p->n_fluid_global <= 0) return fail("n_fluid_global must be positive");
    CUDA_OK(cudaSetDevice(c->device));
    TwoPhaseArgs a{};
    a.stride = c->stride;
    a.n = c->n;
    a.nPad = c->nPad;
    fillIndexView(c, a.idx);
    a.ptable = c->d_ptable;
    a.excInfo = c->d_excInfo;
    a.exc = c->d_exc;
    a.phi = c->d_phi;
    a.rho = c->d_rho;
    a.vel = c->d_vel;
    // main_TWOPHASE.cpp:126-129
    a.nu0Inv = 1.0 / (kC2 * (p->tau0 - 0.5));
    a.nu1Inv = 1.0 / (kC2 * (p->tau1 - 0.5));
    a.sigma = p->sigma;
    a.beta = p->beta;
    for (int d = 0; d < 3; ++d) a.F[d] = d < c->li.nD ? p->force[d] : 0.0;
    a.forceX = c->d_forceX;
    a.partial = c->d_fluxPartial;
    const unsigned gridAll = (unsigned)((c->n + 255) / 256);
    // CHIMP_TRACE=1: phase durations on the main stream (moment pass / exchange + global sum / boundary collide /
    // interior collide + halo completion), averaged over the call and printed per rank
    const bool trace = c->trace;
    std::vector<cudaEvent_t> ev;
    auto mark = [&]() {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, c->stream);
        ev.push_back(e);
    };
    for (int s = 0; s < n_steps; ++s) {
        fillPlanes(c, a.pl);
        double *const foutBuf = c->d_f[c->cur ^ 1];
        const bool mom = (s == n_steps - 1);
        mark();
        if (multi && c->peerTwoPhase) {
            // the halo-in slots read by the moment pass were stored by the neighbours during their previous step
            waitFlagsKernel<<<1, 32, 0, c->stream>>>(c->d_flags, (1u << c->nbrs.size()) - 1u, (unsigned long long)c->steps, c->timeoutNs, c->d_error);
            ++g_launches;
        }
        mark();
        // passes A + C (:238-246, :292-299): rho0, rho1, phi and the local x-momentum sum
        phaseMoments(c, a, gridAll);
        if (multi && c->peerTwoPhase) {
            // fold of the partials and delivery of the local sum into every rank's mailbox in one launch
            foldAndPushKernel<<<1, 256, 0, c->stream>>>(c->d_fluxPartial, (int)gridAll, c->d_fluxSum, (void *const *)c->d_peerMail, c->worldRank,
                                                         c->worldSize, (int)(c->steps & 1), (unsigned long long)c->steps + 1);
        } else {
            fluxForceKernel<<<1, 256, 0, c->stream>>>(c->d_fluxPartial, (int)gridAll, p->momx, (double)p->n_fluid_global, c->d_fluxSum, c->d_forceX, multi ? 0 : 1);
        }
        ++g_launches;
        mark();
        if (multi && c->peerTwoPhase) {
            // everything between the passes goes over peer memory: the momentum sums through the mailboxes (above), my
            // boundary colours into the neighbours' ghost slots; then wait for theirs and fold the sums in rank order
            const unsigned long long seq = (unsigned long long)c->steps + 1;
            const int parity = (int)(c->steps & 1);
            for (size_t k = 0; k < c->nbrs.size(); ++k) {
                Neighbor &nb = c->nbrs[k];
                if (!nb.phiSendCount) continue;
                haloPushKernel<<<(unsigned)((nb.phiSendCount + 255) / 256), 256, 0, c->stream>>>(
                    nb.peerPhi, c->d_phi, nb.d_phiSendSrc, nb.d_peerPhiDst, (int)nb.phiSendCount, 1, 0, 0, nb.d_phiBlockCounter,
                    nb.peerFlags + 8 + nb.peerFace, seq);
                ++g_launches;
            }
            unsigned phiMask = 0;
            for (size_t k = 0; k < c->nbrs.size(); ++k)
                if (c->nbrs[k].phiRecvCount) phiMask |= 1u << k;
            sumWaitFoldKernel<<<1, kMaxWorld + 32, 0, c->stream>>>(c->d_mail, c->worldSize, parity, seq, p->momx, (double)p->n_fluid_global,
                                                                  c->d_fluxSum, c->d_forceX, c->d_flags + 8, phiMask, seq, c->timeoutNs, c->d_error);
            ++g_launches;
        } else if (multi) {
            // communciateScalarField(cgField) (:287) and MPI_Allreduce of the momentum sum (:299)
            for (auto &nb : c->nbrs)
                if (nb.phiSendCount) {
                    haloPackKernel<<<(unsigned)((nb.phiSendCount + 255) / 256), 256, 0, c->stream>>>(nb.d_phiSendBuf, c->d_phi, nb.d_phiSendSrc, (int)nb.phiSendCount);
                    ++g_launches;
                }
            if (c->scalarExchange(c->scalarExchangeUser, (void *)c->stream)) return fail("scalar exchange callback failed");
            if (c->allreduce(c->allreduceUser, c->d_fluxSum, 1, (void *)c->stream)) return fail("allreduce callback failed");
            for (auto &nb : c->nbrs)
                if (nb.phiRecvCount) {
                    haloUnpackKernel<<<(unsigned)((nb.phiRecvCount + 255) / 256), 256, 0, c->stream>>>(c->d_phi, nb.d_phiRecvBuf, nb.d_phiRecvDst, (int)nb.phiRecvCount);
                    ++g_launches;
                }
            fluxForceKernel<<<1, 256, 0, c->stream>>>(c->d_fluxSum, 1, p->momx, (double)p->n_fluid_global, c->d_fluxSum, c->d_forceX, 1);
            ++g_launches;
        }
        mark();
        // pass D (:312-376) + ghost exchange of both fields (:387-388)
        if (!multi) {
            a.begin = 0;
            a.end = c->n;
            twoPhaseCollide(c, a, mom);
            mark();
        } else {
            a.begin = 0;
            a.end = c->nBoundary ? c->nBoundary : c->n;
            // peer mode: the halo-coupled nodes run on the high-priority halo stream, concurrently with the interior nodes
            // (both read the old buffer and phi, and write disjoint slots); callback mode keeps them on the main stream
            const bool split = c->peerTwoPhase && c->nBoundary && c->nBoundary < c->n;
            if (split) {
                CUDA_OK(cudaEventRecord(c->evBoundary, c->stream));
                CUDA_OK(cudaStreamWaitEvent(c->haloStream, c->evBoundary, 0));
                twoPhaseCollide(c, a, mom, c->haloStream);
            } else {
                twoPhaseCollide(c, a, mom);
                CUDA_OK(cudaEventRecord(c->evBoundary, c->stream));
                CUDA_OK(cudaStreamWaitEvent(c->haloStream, c->evBoundary, 0));
            }
            mark();
            if (c->peerTwoPhase) {
                // both fields of the outgoing faces straight into the neighbours' halo-in slots (LBmonlatmpi.h:253-257)
                const long long fieldStride = (long long)c->li.nQ * c->stride;
                const int outIdx = c->cur ^ 1;
                for (auto &nb : c->nbrs)
                    if (nb.sendCount) {
                        haloPushKernel<<<(unsigned)((nb.sendCount + 255) / 256), 256, 0, c->haloStream>>>(
                            nb.peerX[outIdx], foutBuf, nb.d_sendSrc, nb.d_peerDst, (int)nb.sendCount, c->nFields, fieldStride,
                            nb.peerFieldStride, nb.d_blockCounter, nb.peerFlags + nb.peerFace, (unsigned long long)(c->steps + 1));
                        ++g_launches;
                    }
            } else {
                packHalos(c, foutBuf);
                if (c->exchange(c->exchangeUser, (void *)c->haloStream)) return fail("exchange callback failed");
            }
            if (c->nBoundary && c->nBoundary < c->n) {
                a.begin = c->nBoundary;
                a.end = c->n;
                twoPhaseCollide(c, a, mom);
            }
            if (!c->peerTwoPhase) unpackHalos(c, foutBuf);
            CUDA_OK(cudaEventRecord(c->evHalo, c->haloStream));
            CUDA_OK(cudaStreamWaitEvent(c->stream, c->evHalo, 0));
        }
        mark();
        c->cur ^= 1;
        ++c->steps;
    }
    CUDA_OK(cudaGetLastError());
    if (trace && !ev.empty()) {
        CUDA_OK(cudaStreamSynchronize(c->stream));
        const int per = (int)ev.size() / n_steps; // marks per step: start, wait done, A done, exchange done, D(b) done, end
        std::vector<double> sum(per, 0.0);
        for (int s = 0; s < n_steps; ++s)
            for (int k = 0; k + 1 < per; ++k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[(size_t)s * per + k], ev[(size_t)s * per + k + 1]);
                sum[k] += ms;
            }
        float total = 0.f;
        cudaEventElapsedTime(&total, ev.front(), ev.back());
        fprintf(stderr, "[chimp trace] rank %d: %d steps, %.4f ms/step; phases (ms/step):", c->worldRank, n_steps, total / n_steps);
        for (int k = 0; k + 1 < per; ++k) fprintf(stderr, " %.4f", sum[k] / n_steps);
        fprintf(stderr, "  [flag wait | moment pass + fold | phi + sum exchange | boundary collide | interior collide + halo]\n");
        for (auto e : ev) cudaEventDestroy(e);
    }
    return 0;
}

int chimp_step_twophase_timed(chimp_lattice *c, const chimp_twophase_params *p, int n_steps, double *ms)
{
    if (check(c, true)) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    CUDA_OK(cudaEventRecord(e0, c->stream));
    const int rc = chimp_step_twophase(c, p, n_steps);
    CUDA_OK(cudaEventRecord(e1, c->stream));
    CUDA_OK(cudaEventSynchronize(e1));
    float t = 0.f;
    CUDA_OK(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = t;
    return rc;
}

int chimp_download_phase_field(chimp_lattice *c, double *cg)
{
    if (check(c, true)) return 1;
    if (c->nFields != 2) return fail("needs a two-field lattice");
    if (downloadPlanes(c, cg, c->d_phi, 1, 1, 0)) return 1;
    // solid-boundary rows carry the constant wall colour (main_TWOPHASE.cpp:280-284)
    std::vector<double> phiWall(std::max(c->nSolid, 1));
    if (c->nSolid) CUDA_OK(cudaMemcpy(phiWall.data(), c->d_phi + c->nPad, (size_t)c->nSolid * sizeof(double), cudaMemcpyDeviceToHost));
    for (int k = 0; k < c->nSolid; ++k) cg[c->solidBnd[k]] = phiWall[k];
    return 0;
}

double chimp_last_flux_force(chimp_lattice *c)
{
    if (!c || !c->finalized || !c->d_forceX) return 0.0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    double v = 0.0;
    cudaMemcpy(&v, c->d_forceX, sizeof(double), cudaMemcpyDeviceToHost);
    return v;
}

// ---- halo plumbing ----------------------------------------------------------------------
extern "C++" {
namespace {
template <class L>
void launchMomentumSum(chimp_lattice *c, const StepArgs &a, long long fieldOff, int cartDir, unsigned grid)
{
    if (c->indexForm == CHIMP_INDEX_COMPACT) momentumSumKernel<L, IDX_COMPACT><<<grid, 256, 0, c->stream>>>(a, fieldOff, cartDir, c->d_fluxPartial);
    else momentumSumKernel<L, IDX_TABLE><<<grid, 256, 0, c->stream>>>(a, fieldOff, cartDir, c->d_fluxPartial);
    ++g_launches;
}
} // namespace
} // extern "C++"

int chimp_flux_force(chimp_lattice *c, int field_no, int cart_dir, double fixed_flux, long long n_nodes_global, double *force_out)
{
    if (check(c, true)) return 1;
    if (field_no < 0 || field_no >= c->nFields) return fail("field %d out of range", field_no);
    if (cart_dir < 0 || cart_dir >= c->li.nD) return fail("cartesian direction %d out of range", cart_dir);
    if (n_nodes_global <= 0 || !force_out) return fail("bad arguments");
    if (!c->nbrs.empty() && !c->allreduce) return fail("flux force across ranks needs chimp_set_allreduce_callback (LBglobalforcing.h:28)");
    CUDA_OK(cudaSetDevice(c->device));
    const unsigned grid = (unsigned)((c->n + 255) / 256);
    if (!c->d_fluxPartial) CUDA_OK(cudaMalloc(&c->d_fluxPartial, (size_t)std::max(grid, 1u) * sizeof(double)));
    if (!c->d_fluxSum) CUDA_OK(cudaMalloc(&c->d_fluxSum, sizeof(double)));
    if (!c->d_forceX) CUDA_OK(cudaMalloc(&c->d_forceX, 4 * sizeof(double)));
    StepArgs a{};
    a.stride = c->stride;
    a.n = c->n;
    a.nPad = c->nPad;
    fillIndexView(c, a.idx);
    fillPlanes(c, a.pl);
    const long long fieldOff = (long long)field_no * c->li.nQ * c->stride;
    awaitPeers(c);
    switch (c->lattice) {
    case CHIMP_D2Q9: launchMomentumSum<D2Q9>(c, a, fieldOff, cart_dir, grid); break;
    case CHIMP_D3Q19: launchMomentumSum<D3Q19>(c, a, fieldOff, cart_dir, grid); break;
    case CHIMP_D3Q27: launchMomentumSum<D3Q27>(c, a, fieldOff, cart_dir, grid); break;
    }
    const bool multi = !c->nbrs.empty();
    fluxForceKernel<<<1, 256, 0, c->stream>>>(c->d_fluxPartial, (int)grid, fixed_flux, (double)n_nodes_global, c->d_fluxSum, c->d_forceX + 2, multi ? 0 : 1);
    ++g_launches;
    if (multi) {
        if (c->allreduce(c->allreduceUser, c->d_fluxSum, 1, (void *)c->stream)) return fail("allreduce callback failed");
        fluxForceKernel<<<1, 256, 0, c->stream>>>(c->d_fluxSum, 1, fixed_flux, (double)n_nodes_global, c->d_fluxSum, c->d_forceX + 2, 1);
        ++g_launches;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(force_out, c->d_forceX + 2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return checkDeviceError(c);
}

extern "C++" {
namespace {
template <class L>
void launchCapNumberSums(chimp_lattice *c, const StepArgs &a, int cartDir, unsigned grid, double *partial)
{
    if (c->indexForm == CHIMP_INDEX_COMPACT) capNumberSumsKernel<L, IDX_COMPACT><<<grid, 256, 0, c->stream>>>(a, cartDir, c->d_rho, partial);
    else capNumberSumsKernel<L, IDX_TABLE><<<grid, 256, 0, c->stream>>>(a, cartDir, c->d_rho, partial);
    ++g_launches;
}
} // namespace
} // extern "C++"

int chimp_capillary_force(chimp_lattice *c, int cart_dir, double sigma_cap_numb, double nu0, double nu1, long long n_nodes_global,
                          double *force_out)
{
    if (check(c, true)) return 1;
    if (c->nFields != 2) return fail("the capillary-number force needs a two-field lattice (rho of both fluids)");
    if (cart_dir < 0 || cart_dir >= c->li.nD) return fail("cartesian direction %d out of range", cart_dir);
    if (n_nodes_global <= 0 || !force_out) return fail("bad arguments");
    if (!c->nbrs.empty() && !c->allreduce) return fail("capillary-number force across ranks needs chimp_set_allreduce_callback (LBglobalforcing.h:78-81)");
    CUDA_OK(cudaSetDevice(c->device));
    const unsigned grid = (unsigned)((c->n + 255) / 256);
    double *d_partial = nullptr, *d_sums = nullptr;
    CUDA_OK(cudaMalloc(&d_partial, (size_t)4 * std::max(grid, 1u) * sizeof(double)));
    CUDA_OK(cudaMalloc(&d_sums, 4 * sizeof(double)));
    StepArgs a{};
    a.stride = c->stride;
    a.n = c->n;
    a.nPad = c->nPad;
    fillIndexView(c, a.idx);
    fillPlanes(c, a.pl);
    awaitPeers(c);
    switch (c->lattice) {
    case CHIMP_D2Q9: launchCapNumberSums<D2Q9>(c, a, cart_dir, grid, d_partial); break;
    case CHIMP_D3Q19: launchCapNumberSums<D3Q19>(c, a, cart_dir, grid, d_partial); break;
    case CHIMP_D3Q27: launchCapNumberSums<D3Q27>(c, a, cart_dir, grid, d_partial); break;
    }
    foldRowsKernel<<<4, 256, 0, c->stream>>>(d_partial, (int)grid, d_sums);
    ++g_launches;
    int rc = 0;
    if (!c->nbrs.empty() && c->allreduce(c->allreduceUser, d_sums, 4, (void *)c->stream)) rc = fail("allreduce callback failed");
    double h[4] = {0, 0, 0, 0};
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_sums, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_partial);
    cudaFree(d_sums);
    if (rc) return rc;
    if (e != cudaSuccess) return fail("capillary-number sums failed: %s", cudaGetErrorString(e));
    // LBglobalforcing.h:85-95
    for (double &x : h) x /= (double)n_nodes_global;
    *force_out = 2 * (sigma_cap_numb - (h[0] * nu0 + h[1] * nu1)) / (h[2] * nu0 + h[3] * nu1);
    return 0;
}

int chimp_node_list_flux(chimp_lattice *c, int n_list, const int32_t *nodes, const int32_t *bin, int n_bins, int field_no,
                         int component, double *out)
{
    if (check(c, true)) return 1;
    if (n_list < 0 || (n_list && (!nodes || !bin)) || n_bins <= 0 || !out) return fail("bad arguments");
    if (field_no < 0 || field_no >= c->nFields) return fail("field %d out of range", field_no);
    if (component < 0 || component >= c->li.nD) return fail("velocity component %d out of range", component);
    for (int k = 0; k < n_list; ++k)
        if (bin[k] < 0 || bin[k] >= n_bins) return fail("bin %d of list entry %d out of range", bin[k], k);
    CUDA_OK(cudaSetDevice(c->device));
    for (int b = 0; b < n_bins; ++b) out[b] = 0.0;
    if (n_list == 0) return 0;
    if (labelRange(c)) return 1;
    const int nLabels = c->labelMax + 1;
    if (!c->d_slotOf) {
        CUDA_OK(cudaMalloc(&c->d_slotOf, (size_t)nLabels * sizeof(int32_t)));
        CUDA_OK(cudaMemsetAsync(c->d_slotOf, 0xff, (size_t)nLabels * sizeof(int32_t), c->stream));
        invertLabelsKernel<<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(c->d_label, c->n, c->d_slotOf);
        ++g_launches;
    }
    int32_t *d_nodes = nullptr;
    double *d_out = nullptr;
    CUDA_OK(cudaMalloc(&d_nodes, (size_t)n_list * sizeof(int32_t)));
    CUDA_OK(cudaMalloc(&d_out, (size_t)n_list * sizeof(double)));
    CUDA_OK(cudaMemcpyAsync(d_nodes, nodes, (size_t)n_list * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    nodeFluxKernel<<<(unsigned)((n_list + 255) / 256), 256, 0, c->stream>>>(d_nodes, n_list, c->d_slotOf, nLabels,
                                                                           c->d_rho + (size_t)field_no * c->nPad,
                                                                           c->d_vel + (size_t)component * c->nPad, d_out);
    ++g_launches;
    std::vector<double> prod(n_list);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(prod.data(), d_out, (size_t)n_list * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_nodes);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail("node list flux failed: %s", cudaGetErrorString(e));
    // the reference adds the products in list order (std_one_phase/main.cpp:609-618)
    for (int k = 0; k < n_list; ++k) out[bin[k]] += prod[k];
    return 0;
}

int chimp_num_neighbors(chimp_lattice *c) { return c ? (int)c->nbrs.size() : 0; }
int chimp_neighbor_info(chimp_lattice *c, int k, int *neig_rank, long long *send_count, long long *recv_count)
{
    if (!c || k < 0 || k >= (int)c->nbrs.size()) return fail("bad neighbour index");
    if (neig_rank) *neig_rank = c->nbrs[k].rank;
    if (send_count) *send_count = c->nbrs[k].sendCount * c->nFields;
    if (recv_count) *recv_count = c->nbrs[k].recvCount * c->nFields;
    return 0;
}
void *chimp_send_buffer_dev(chimp_lattice *c, int k) { return (c && k >= 0 && k < (int)c->nbrs.size()) ? c->nbrs[k].sendBuf() : nullptr; }
void *chimp_recv_buffer_dev(chimp_lattice *c, int k) { return (c && k >= 0 && k < (int)c->nbrs.size()) ? c->nbrs[k].recvBuf() : nullptr; }
int chimp_set_halo_buffers(chimp_lattice *c, int k, void *send_dev, void *recv_dev)
{
    if (!c || k < 0 || k >= (int)c->nbrs.size()) return fail("bad neighbour index");
    c->nbrs[k].x_sendBuf = (double *)send_dev;
    c->nbrs[k].x_recvBuf = (double *)recv_dev;
    return 0;
}
void *chimp_halo_stream(chimp_lattice *c) { return c ? (void *)c->haloStream : nullptr; }

int chimp_add_halo_face(chimp_lattice *c, int neig_rank, long long n_send, const long long *send_src, long long n_recv,
                        const long long *recv_dst)
{
    if (check(c, true)) return 1;
    if (n_send < 0 || n_recv < 0 || n_send >= (1ll << 31) || n_recv >= (1ll << 31)) return fail("bad halo face size");
    CUDA_OK(cudaSetDevice(c->device));
    const long long planeSlots = c->stride * c->li.nQ;
    for (long long k = 0; k < n_send; ++k)
        if (send_src[k] < 0 || send_src[k] >= planeSlots) return fail("send offset %lld out of range", send_src[k]);
    for (long long k = 0; k < n_recv; ++k)
        if (recv_dst[k] < 0 || recv_dst[k] >= planeSlots) return fail("recv offset %lld out of range", recv_dst[k]);
    Neighbor nb;
    nb.rank = neig_rank;
    nb.sendCount = n_send;
    nb.recvCount = n_recv;
    nb.hSrc.assign(send_src, send_src + n_send);
    nb.hRecv.assign(recv_dst, recv_dst + n_recv);
    if (n_send) {
        CUDA_OK(cudaMalloc(&nb.d_sendSrc, (size_t)n_send * sizeof(long long)));
        CUDA_OK(cudaMemcpy(nb.d_sendSrc, send_src, (size_t)n_send * sizeof(long long), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&nb.d_sendBuf, (size_t)n_send * c->nFields * sizeof(double)));
    }
    if (n_recv) {
        CUDA_OK(cudaMalloc(&nb.d_recvDst, (size_t)n_recv * sizeof(long long)));
        CUDA_OK(cudaMemcpy(nb.d_recvDst, recv_dst, (size_t)n_recv * sizeof(long long), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&nb.d_recvBuf, (size_t)n_recv * c->nFields * sizeof(double)));
    }
    c->nbrs.push_back(std::move(nb));
    return 0;
}

int chimp_add_scalar_halo_face(chimp_lattice *c, int k, long long n_send, const long long *send_src, long long n_recv,
                               const long long *recv_dst, void *send_dev, void *recv_dev)
{
    if (check(c, true)) return 1;
    if (c->nFields != 2 || !c->d_phi) return fail("scalar halos need a two-field lattice with its phi table set");
    if (k < 0 || k >= (int)c->nbrs.size()) return fail("bad neighbour index");
    if (n_send < 0 || n_recv < 0 || (n_send && !send_src) || (n_recv && !recv_dst)) return fail("bad scalar halo lists");
    for (long long e = 0; e < n_send; ++e)
        if (send_src[e] < 0 || send_src[e] >= c->n) return fail("scalar send slot %lld is not an own node", send_src[e]);
    for (long long e = 0; e < n_recv; ++e)
        if (recv_dst[e] < c->nPad || recv_dst[e] >= c->nPhi - 1) return fail("scalar receive slot %lld is not a ghost slot", recv_dst[e]);
    CUDA_OK(cudaSetDevice(c->device));
    Neighbor &nb = c->nbrs[k];
    freeDev(nb.d_phiSendSrc); freeDev(nb.d_phiRecvDst);
    if (nb.ownPhiBufs) { freeDev(nb.d_phiSendBuf); freeDev(nb.d_phiRecvBuf); }
    nb.phiSendCount = n_send;
    nb.phiRecvCount = n_recv;
    nb.ownPhiBufs = !(send_dev || recv_dev);
    if ((send_dev == nullptr) != (recv_dev == nullptr)) return fail("pass both scalar halo buffers or neither");
    if (n_send) {
        CUDA_OK(cudaMalloc(&nb.d_phiSendSrc, (size_t)n_send * sizeof(long long)));
        CUDA_OK(cudaMemcpy(nb.d_phiSendSrc, send_src, (size_t)n_send * sizeof(long long), cudaMemcpyHostToDevice));
        if (nb.ownPhiBufs) CUDA_OK(cudaMalloc(&nb.d_phiSendBuf, (size_t)n_send * sizeof(double)));
    }
    if (n_recv) {
        CUDA_OK(cudaMalloc(&nb.d_phiRecvDst, (size_t)n_recv * sizeof(long long)));
        CUDA_OK(cudaMemcpy(nb.d_phiRecvDst, recv_dst, (size_t)n_recv * sizeof(long long), cudaMemcpyHostToDevice));
        if (nb.ownPhiBufs) CUDA_OK(cudaMalloc(&nb.d_phiRecvBuf, (size_t)n_recv * sizeof(double)));
    }
    if (!nb.ownPhiBufs) {
        nb.d_phiSendBuf = (double *)send_dev;
        nb.d_phiRecvBuf = (double *)recv_dev;
    }
    return 0;
}

int chimp_ipc_handles(chimp_lattice *c, unsigned char *out192)
{
    if (check(c, true)) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    void *ptrs[3] = {c->d_f[0], c->d_f[1], c->d_flags};
    for (int k = 0; k < 3; ++k) {
        cudaIpcMemHandle_t h;
        CUDA_OK(cudaIpcGetMemHandle(&h, ptrs[k]));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(out192 + 64 * k, &h, 64);
    }
    return 0;
}

int chimp_connect_peer(chimp_lattice *c, int k, const unsigned char *peer_handles192, int same_process,
                       void *const *peer_ptrs3, long long peer_field_stride, int peer_face, long long n_dst,
                       const long long *peer_dst)
{
    if (check(c, true)) return 1;
    if (k < 0 || k >= (int)c->nbrs.size() || k >= 8) return fail("bad neighbour index");
    if (peer_face < 0 || peer_face >= 8) return fail("bad peer face index");
    Neighbor &nb = c->nbrs[k];
    if (n_dst != nb.sendCount) return fail("peer destination list has %lld entries, my send list %lld", n_dst, nb.sendCount);
    CUDA_OK(cudaSetDevice(c->device));
    void *ptrs[3] = {nullptr, nullptr, nullptr};
    if (same_process) {
        // both contexts live in this process (tests): the peer's device pointers are directly usable
        for (int j = 0; j < 3; ++j) ptrs[j] = peer_ptrs3[j];
    } else {
        for (int j = 0; j < 3; ++j)
            if (openIpc(peer_handles192 + 64 * j, &ptrs[j])) return 1;
    }
    nb.peerX[0] = (double *)ptrs[0];
    nb.peerX[1] = (double *)ptrs[1];
    nb.peerFlags = (unsigned long long *)ptrs[2];
    nb.peerFieldStride = peer_field_stride;
    nb.peerFace = peer_face;
    freeDev(nb.d_peerDst);
    nb.hPeerDst.assign(peer_dst, peer_dst + n_dst);
    if (n_dst) {
        CUDA_OK(cudaMalloc(&nb.d_peerDst, (size_t)n_dst * sizeof(long long)));
        CUDA_OK(cudaMemcpy(nb.d_peerDst, peer_dst, (size_t)n_dst * sizeof(long long), cudaMemcpyHostToDevice));
    }
    if (!nb.d_blockCounter) {
        CUDA_OK(cudaMalloc(&nb.d_blockCounter, sizeof(unsigned)));
        CUDA_OK(cudaMemset(nb.d_blockCounter, 0, sizeof(unsigned)));
    }
    bool all = true;
    for (auto &x : c->nbrs) all = all && x.peerFlags != nullptr;
    c->peerHalos = all;
    if (all) {
        if (buildPeerTables(c)) return 1;
        preloadForPeerStepping(c);
    }
    return 0;
}

extern "C++" {
namespace {
// Tables of the fused peer exchange (PeerView): for every halo-coupled node the directions whose written population
// also goes to a neighbour rank, and the slot of the neighbour's plane it lands in.  The fused form needs the
// halo-coupled nodes in the leading slots (boundary_first), every outgoing population in the plane of its own
// direction on the other side (true for MonLatMpi lists and the structured ingest) and the compact index; otherwise
// the step keeps the separate haloPushKernel launches.
int buildPeerTables(chimp_lattice *c)
{
    c->peerFused = false;
    freeDev(c->d_sendMask); freeDev(c->d_sendDst); freeDev(c->d_extraStart); freeDev(c->d_extra);
    c->peerWhy = "";
    auto notFused = [&](const char *why) { c->peerWhy = why; return 0; };
    if (!c->peerFusedEnv) return notFused("switched off (CHIMP_PEER_FUSED=0)");
    if (c->nFields != 1) return notFused("two-field lattice");
    if (c->indexForm != CHIMP_INDEX_COMPACT) return notFused("table index form");
    if (c->nBoundary <= 0) return notFused("halo-coupled nodes are not in the leading slots (boundary_first)");
    if (c->nbrs.size() > (size_t)kMaxFaces) return notFused("more than 8 faces");
    const int nQ = c->li.nQ;
    const int blocks = (c->nBoundary + CHIMP_BLOCK - 1) / CHIMP_BLOCK;
    const int pad = blocks * CHIMP_BLOCK;
    std::vector<uint32_t> mask(pad, 0u);
    std::vector<int32_t> dst((size_t)nQ * pad, -1);
    std::vector<std::pair<int, int2>> more; // (slot, {q, word})
    for (size_t k = 0; k < c->nbrs.size(); ++k) {
        const Neighbor &nb = c->nbrs[k];
        if ((long long)nb.hSrc.size() != nb.sendCount || (long long)nb.hPeerDst.size() != nb.sendCount) return notFused("send list not available on the host");
        if (nb.peerFieldStride % nQ) return notFused("peer field stride is not a multiple of nQ");
        const long long peerPlane = nb.peerFieldStride / nQ;
        for (long long e = 0; e < nb.sendCount; ++e) {
            const long long q = nb.hSrc[e] / c->stride, slot = nb.hSrc[e] % c->stride;
            const long long dq = nb.hPeerDst[e] / peerPlane, dslot = nb.hPeerDst[e] % peerPlane;
            if (slot >= c->nBoundary) return notFused("a sent population belongs to a node outside the leading halo-coupled slots");
            if (dq != q) return notFused("a sent population changes its direction plane on the other side");
            if (dslot >= (1ll << 28)) return notFused("peer plane too long for the packed destination word");
            const int32_t word = (int32_t)((k << 28) | dslot);
            if ((mask[slot] >> q) & 1u) {
                // the reference's lists name a node once per ghost image the receiver holds of it: further destinations
                more.push_back({(int)slot, make_int2((int)q, word)});
                continue;
            }
            mask[slot] |= 1u << q;
            dst[(size_t)q * pad + slot] = word;
        }
    }
    CUDA_OK(cudaMalloc(&c->d_sendMask, mask.size() * sizeof(uint32_t)));
    CUDA_OK(cudaMemcpy(c->d_sendMask, mask.data(), mask.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&c->d_sendDst, dst.size() * sizeof(int32_t)));
    CUDA_OK(cudaMemcpy(c->d_sendDst, dst.data(), dst.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (!more.empty()) {
        std::stable_sort(more.begin(), more.end(), [](const std::pair<int, int2> &x, const std::pair<int, int2> &y) { return x.first < y.first; });
        std::vector<int32_t> start(pad + 1, 0);
        std::vector<int2> ent(more.size());
        for (auto &m : more) ++start[m.first + 1];
        for (int i = 0; i < pad; ++i) start[i + 1] += start[i];
        for (size_t e = 0; e < more.size(); ++e) ent[e] = more[e].second;
        CUDA_OK(cudaMalloc(&c->d_extraStart, start.size() * sizeof(int32_t)));
        CUDA_OK(cudaMemcpy(c->d_extraStart, start.data(), start.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&c->d_extra, ent.size() * sizeof(int2)));
        CUDA_OK(cudaMemcpy(c->d_extra, ent.data(), ent.size() * sizeof(int2), cudaMemcpyHostToDevice));
    }
    if (!c->d_peerCounter) {
        CUDA_OK(cudaMalloc(&c->d_peerCounter, sizeof(unsigned)));
        CUDA_OK(cudaMemset(c->d_peerCounter, 0, sizeof(unsigned)));
    }
    c->peerPad = pad;
    c->peerBlocks = blocks;
    c->peerFused = true;
    return 0;
}
// one mapping per handle and process (a 2-rank ring reaches the same peer through both faces)
int openIpc(const unsigned char *handle64, void **out)
{
    const std::string key((const char *)handle64, 64);
    auto it = g_ipcOpen.find(key);
    if (it == g_ipcOpen.end()) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, 64);
        void *p = nullptr;
        CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        it = g_ipcOpen.emplace(key, p).first;
    }
    *out = it->second;
    return 0;
}
int ensureMailbox(chimp_lattice *c)
{
    if (c->d_mail) return 0;
    CUDA_OK(cudaMalloc(&c->d_mail, 2 * kMaxWorld * sizeof(MailSlot)));
    CUDA_OK(cudaMemset(c->d_mail, 0, 2 * kMaxWorld * sizeof(MailSlot)));
    return 0;
}
} // namespace
} // extern "C++"

extern "C++" {
namespace {
// With lazy module loading (the CUDA 12 default) the first launch of a kernel loads it, and loading waits for
// running kernels -- including an arrival-flag wait that only ends when the peer's push, issued after the load,
// has run: a deadlock.  Every kernel a step may launch is therefore loaded before the first wait kernel exists.
template <class K>
void preloadKernel(K kernel)
{
    cudaFuncAttributes attr;
    cudaFuncGetAttributes(&attr, kernel);
}
template <class L, int IDX, int OP>
void preloadOnePhaseForms()
{
    preloadKernel(collideStreamKernel<L, COLL_BGK, OP, false, IDX>);
    preloadKernel(collideStreamKernel<L, COLL_BGK, OP, true, IDX>);
    preloadKernel(collideStreamKernel<L, COLL_TRT, OP, false, IDX>);
    preloadKernel(collideStreamKernel<L, COLL_TRT, OP, true, IDX>);
    if constexpr (IDX == IDX_COMPACT) {
        preloadKernel(collideStreamKernel<L, COLL_BGK, OP, false, IDX_COMPACT, true>);
        preloadKernel(collideStreamKernel<L, COLL_BGK, OP, true, IDX_COMPACT, true>);
        preloadKernel(collideStreamKernel<L, COLL_TRT, OP, false, IDX_COMPACT, true>);
        preloadKernel(collideStreamKernel<L, COLL_TRT, OP, true, IDX_COMPACT, true>);
    }
}
template <class L, int IDX>
void preloadStepKernels(bool twoField)
{
    preloadOnePhaseForms<L, IDX, OP_NONE>();
    preloadOnePhaseForms<L, IDX, OP_ARRAYS>();
    preloadOnePhaseForms<L, IDX, OP_PACKED>();
    if constexpr (IDX == IDX_COMPACT) {
        preloadKernel(collideStreamKernel<L, COLL_BGK, OP_NONE, false, IDX_COMPACT_MASK>);
        preloadKernel(collideStreamKernel<L, COLL_BGK, OP_NONE, true, IDX_COMPACT_MASK>);
        preloadKernel(collideStreamKernel<L, COLL_TRT, OP_NONE, false, IDX_COMPACT_MASK>);
        preloadKernel(collideStreamKernel<L, COLL_TRT, OP_NONE, true, IDX_COMPACT_MASK>);
    }
    preloadKernel(massChangeKernel<L, IDX>);
    if constexpr (L::id != D3Q27::id) {
        if (twoField) {
            preloadKernel(phaseMomentsKernel<L, IDX>);
            preloadKernel(twoPhaseCollideKernel<L, false, IDX>);
            preloadKernel(twoPhaseCollideKernel<L, true, IDX>);
            preloadKernel(twoPhaseCollideKernel<L, false, IDX, true>);
            preloadKernel(twoPhaseCollideKernel<L, true, IDX, true>);
        }
    }
}
void preloadForPeerStepping(const chimp_lattice *c)
{
    const bool two = c->nFields == 2;
    const bool compact = c->indexForm == CHIMP_INDEX_COMPACT;
    switch (c->lattice) {
    case CHIMP_D2Q9: compact ? preloadStepKernels<D2Q9, IDX_COMPACT>(two) : preloadStepKernels<D2Q9, IDX_TABLE>(two); break;
    case CHIMP_D3Q19: compact ? preloadStepKernels<D3Q19, IDX_COMPACT>(two) : preloadStepKernels<D3Q19, IDX_TABLE>(two); break;
    case CHIMP_D3Q27: compact ? preloadStepKernels<D3Q27, IDX_COMPACT>(two) : preloadStepKernels<D3Q27, IDX_TABLE>(two); break;
    }
    preloadKernel(haloPushKernel);
    preloadKernel(sumWaitFoldKernel);
    preloadKernel(foldAndPushKernel);
    preloadKernel(waitFlagsKernel);
    preloadKernel(fluxForceKernel);
    preloadKernel(massFinalizeKernel);
    preloadKernel(haloPackKernel);
    preloadKernel(haloUnpackKernel);
    preloadKernel(fillKernel);
}
} // namespace
} // extern "C++"

int chimp_ipc_handles_twophase(chimp_lattice *c, unsigned char *out128)
{
    if (check(c, true)) return 1;
    if (c->nFields != 2 || !c->d_phi) return fail("needs a two-field lattice with its phi table set");
    CUDA_OK(cudaSetDevice(c->device));
    if (ensureMailbox(c)) return 1;
    CUDA_OK(cudaStreamSynchronize(c->stream));
    void *ptrs[2] = {c->d_phi, c->d_mail};
    for (int k = 0; k < 2; ++k) {
        cudaIpcMemHandle_t h;
        CUDA_OK(cudaIpcGetMemHandle(&h, ptrs[k]));
        memcpy(out128 + 64 * k, &h, 64);
    }
    return 0;
}

int chimp_local_pointers_twophase(chimp_lattice *c, void **out2)
{
    if (check(c, true)) return 1;
    if (c->nFields != 2 || !c->d_phi) return fail("needs a two-field lattice with its phi table set");
    CUDA_OK(cudaSetDevice(c->device));
    if (ensureMailbox(c)) return 1;
    out2[0] = c->d_phi;
    out2[1] = c->d_mail;
    return 0;
}

int chimp_connect_peer_scalar(chimp_lattice *c, int k, const unsigned char *peer_phi_handle64, int same_process, void *peer_phi_ptr,
                              long long n_dst, const long long *peer_phi_dst)
{
    if (check(c, true)) return 1;
    if (k < 0 || k >= (int)c->nbrs.size() || k >= 8) return fail("bad neighbour index");
    Neighbor &nb = c->nbrs[k];
    if (!nb.peerFlags) return fail("connect the population halo of face %d first (chimp_connect_peer)", k);
    if (n_dst != nb.phiSendCount) return fail("peer phi list has %lld entries, my scalar send list %lld", n_dst, nb.phiSendCount);
    CUDA_OK(cudaSetDevice(c->device));
    void *p = peer_phi_ptr;
    if (!same_process && openIpc(peer_phi_handle64, &p)) return 1;
    nb.peerPhi = (double *)p;
    freeDev(nb.d_peerPhiDst);
    if (n_dst) {
        CUDA_OK(cudaMalloc(&nb.d_peerPhiDst, (size_t)n_dst * sizeof(long long)));
        CUDA_OK(cudaMemcpy(nb.d_peerPhiDst, peer_phi_dst, (size_t)n_dst * sizeof(long long), cudaMemcpyHostToDevice));
    }
    if (!nb.d_phiBlockCounter) {
        CUDA_OK(cudaMalloc(&nb.d_phiBlockCounter, sizeof(unsigned)));
        CUDA_OK(cudaMemset(nb.d_phiBlockCounter, 0, sizeof(unsigned)));
    }
    return 0;
}

int chimp_connect_world(chimp_lattice *c, int rank, int world, const unsigned char *mail_handles, int same_process, void *const *mail_ptrs)
{
    if (check(c, true)) return 1;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail("bad rank %d / world size %d (at most %d ranks)", rank, world, kMaxWorld);
    CUDA_OK(cudaSetDevice(c->device));
    if (ensureMailbox(c)) return 1;
    std::vector<MailSlot *> ptrs(world, nullptr);
    for (int w = 0; w < world; ++w) {
        void *p = nullptr;
        if (w == rank) p = c->d_mail;
        else if (same_process) p = mail_ptrs[w];
        else if (openIpc(mail_handles + 64 * (size_t)w, &p)) return 1;
        ptrs[w] = (MailSlot *)p;
    }
    freeDev(c->d_peerMail);
    CUDA_OK(cudaMalloc(&c->d_peerMail, (size_t)world * sizeof(MailSlot *)));
    CUDA_OK(cudaMemcpy(c->d_peerMail, ptrs.data(), (size_t)world * sizeof(MailSlot *), cudaMemcpyHostToDevice));
    c->worldRank = rank;
    c->worldSize = world;
    bool all = c->peerHalos;
    for (auto &x : c->nbrs) all = all && x.peerPhi != nullptr;
    c->peerTwoPhase = all;
    if (!all) return fail("connect every face (chimp_connect_peer, chimp_connect_peer_scalar) before chimp_connect_world");
    preloadForPeerStepping(c);
    return 0;
}

int chimp_local_pointers(chimp_lattice *c, void **out3)
{
    if (check(c, true)) return 1;
    out3[0] = c->d_f[0];
    out3[1] = c->d_f[1];
    out3[2] = c->d_flags;
    return 0;
}

int chimp_peer_mode(chimp_lattice *c, char *why, int why_len)
{
    if (!c) return 0;
    if (why && why_len > 0) snprintf(why, (size_t)why_len, "%s", c->peerWhy.c_str());
    return !c->peerHalos ? 0 : c->peerFused ? 2 : 1;
}

int chimp_set_boundary_count(chimp_lattice *c, int n_boundary)
{
    if (check(c, true)) return 1;
    if (n_boundary < 0 || n_boundary > c->n) return fail("boundary count out of range");
    c->nBoundary = std::min(((n_boundary + 31) / 32) * 32, c->n);
    if (c->peerHalos && buildPeerTables(c)) return 1;
    return 0;
}
int chimp_set_exchange_callback(chimp_lattice *c, chimp_exchange_fn fn, void *user)
{
    if (!c) return fail("null lattice handle");
    c->exchange = fn;
    c->exchangeUser = user;
    return 0;
}
int chimp_set_scalar_exchange_callback(chimp_lattice *c, chimp_exchange_fn fn, void *user)
{
    if (!c) return fail("null lattice handle");
    c->scalarExchange = fn;
    c->scalarExchangeUser = user;
    return 0;
}
int chimp_set_allreduce_callback(chimp_lattice *c, chimp_allreduce_fn fn, void *user)
{
    if (!c) return fail("null lattice handle");
    c->allreduce = fn;
    c->allreduceUser = user;
    return 0;
}
int chimp_scalar_neighbor_info(chimp_lattice *c, int k, long long *send_count, long long *recv_count)
{
    if (!c || k < 0 || k >= (int)c->nbrs.size()) return fail("bad neighbour index");
    if (send_count) *send_count = c->nbrs[k].phiSendCount;
    if (recv_count) *recv_count = c->nbrs[k].phiRecvCount;
    return 0;
}
void *chimp_scalar_send_buffer_dev(chimp_lattice *c, int k) { return (c && k >= 0 && k < (int)c->nbrs.size()) ? c->nbrs[k].d_phiSendBuf : nullptr; }
void *chimp_scalar_recv_buffer_dev(chimp_lattice *c, int k) { return (c && k >= 0 && k < (int)c->nbrs.size()) ? c->nbrs[k].d_phiRecvBuf : nullptr; }
int chimp_set_stream(chimp_lattice *c, void *s)
{
    if (!c) return fail("null lattice handle");
    if (c->ownStream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = (cudaStream_t)s;
    c->ownStream = false;
    return 0;
}
int chimp_synchronize(chimp_lattice *c)
{
    if (!c) return fail("null lattice handle");
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaStreamSynchronize(c->haloStream));
    return checkDeviceError(c);
}

// ---- introspection ------------------------------------------------------------------------
// ---- structured geometry ingest for host code --------------------------------------------------------
// A raw voxel array becomes the pull table with the reference's numbering (own fluid nodes labelled 1..N in C-order
// of geo[x][y][z], vtklb.py:92-94; device slot = label - 1) and std_case semantics (half-way bounce back on every
// link whose upstream cell is solid or lies outside a non-periodic axis, LBhalfwaybb.h:37-63):
//     T[q][i] = label(pos_i - c_q) - 1   if that cell is fluid,   -1 otherwise.
// Pure host code, threaded over x; the reference's route (vtklb.py -> ASCII file -> LBvtk, LBvtk.h:221-262) cannot
// load cases beyond 2 GiB of text (LBvtk.h:194-201).
extern "C++" {
namespace {
struct VoxelGrid {
    int nd, nx, ny, nz, per[3];
    const uint8_t *v;
    long long cells() const { return (long long)nx * ny * nz; }
    long long at(int x, int y, int z) const { return ((long long)x * ny + y) * nz + z; }
    // cell index of pos + (dx, dy, dz) or -1 when it leaves a non-periodic axis
    long long shifted(int x, int y, int z, int dx, int dy, int dz) const
    {
        int p[3] = {x + dx, y + dy, z + dz};
        const int n[3] = {nx, ny, nz};
        for (int a = 0; a < 3; ++a) {
            if (p[a] < 0 || p[a] >= n[a]) {
                if (!per[a]) return -1;
                p[a] = (p[a] + n[a]) % n[a];
            }
        }
        return at(p[0], p[1], p[2]);
    }
};

int voxelGrid(VoxelGrid &g, int lattice, int nx, int ny, int nz, const uint8_t *voxels, int periodicMask)
{
    const LatInfo li = latInfo(lattice);
    if (li.nQ == 0) return fail("unknown lattice id %d", lattice);
    if (!voxels || nx < 1 || ny < 1 || nz < 1) return fail("bad voxel array");
    if (li.nD == 2 && nz != 1) return fail("a 2-D lattice takes a voxel array [nx][ny] (nz = 1)");
    if ((long long)nx * ny * nz >= (1ll << 31)) return fail("voxel array too large for 32-bit cell indices");
    g = VoxelGrid{li.nD, nx, ny, nz, {periodicMask & 1, (periodicMask >> 1) & 1, li.nD == 3 ? (periodicMask >> 2) & 1 : 1}, voxels};
    return 0;
}

void lookupC(int lattice, int q, int c[3])
{
    const LatInfo li = latInfo(lattice);
    for (int d = 0; d < 3; ++d) c[d] = d < li.nD ? chimp_lattice_c(lattice, q, d) : 0;
}

template <class Fn>
void parallelOverX(int nx, Fn fn)
{
    const int nThreads = std::max(1, std::min<int>(nx, (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()))));
    std::vector<std::thread> th;
    for (int t = 0; t < nThreads; ++t)
        th.emplace_back([=] { fn((int)((long long)nx * t / nThreads), (int)((long long)nx * (t + 1) / nThreads)); });
    for (auto &x : th) x.join();
}

// label[cell] = 1..N for fluid cells in C-order, 0 for solid
long long labelCells(const VoxelGrid &g, std::vector<int32_t> &label)
{
    label.assign((size_t)g.cells(), 0);
    long long n = 0;
    for (long long k = 0; k < g.cells(); ++k)
        if (g.v[k]) label[(size_t)k] = (int32_t)++n;
    return n;
}

// one direction of the pull table for the x-range [x0, x1)
void fillTableRow(const VoxelGrid &g, const std::vector<int32_t> &label, const int c[3], int32_t *row, int x0, int x1)
{
    for (int x = x0; x < x1; ++x)
        for (int y = 0; y < g.ny; ++y)
            for (int z = 0; z < g.nz; ++z) {
                const int32_t me = label[(size_t)g.at(x, y, z)];
                if (!me) continue;
                const long long up = g.shifted(x, y, z, -c[0], -c[1], -c[2]);
                const int32_t l = up >= 0 ? label[(size_t)up] : 0;
                row[me - 1] = l ? l - 1 : -1;
            }
}
} // namespace
} // extern "C++"

int chimp_voxel_table_host(int lattice, int nx, int ny, int nz, const uint8_t *voxels, int periodic_mask, int *n_own, int *n_pad,
                           int32_t *table, int32_t *labels)
{
    VoxelGrid g;
    if (voxelGrid(g, lattice, nx, ny, nz, voxels, periodic_mask)) return 1;
    std::vector<int32_t> label;
    const long long n = labelCells(g, label);
    if (n == 0) return fail("the voxel array has no fluid cell");
    const int nPad = (int)(((n + 31) / 32) * 32);
    if (n_own) *n_own = (int)n;
    if (n_pad) *n_pad = nPad;
    if (labels) {
        for (int i = 0; i < nPad; ++i) labels[i] = i < n ? i + 1 : 0;
    }
    if (table) {
        const int nQ = latInfo(lattice).nQ;
        for (int q = 0; q < nQ; ++q) {
            int c[3];
            lookupC(lattice, q, c);
            int32_t *row = table + (size_t)q * nPad;
            std::fill(row, row + nPad, -1);
            parallelOverX(nx, [&](int x0, int x1) { fillTableRow(g, label, c, row, x0, x1); });
        }
    }
    return 0;
}

int chimp_create_from_voxels(chimp_lattice **out, int lattice, int nx, int ny, int nz, const uint8_t *voxels, int periodic_mask,
                             int n_fields, int index_form, int device)
{
    if (!out) return fail("out is null");
    *out = nullptr;
    VoxelGrid g;
    if (voxelGrid(g, lattice, nx, ny, nz, voxels, periodic_mask)) return 1;
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) return fail("no CUDA device available: this engine has no CPU fallback");
    if (device < 0) CUDA_OK(cudaGetDevice(&device));
    CUDA_OK(cudaSetDevice(device));
    std::vector<int32_t> label;
    const long long n = labelCells(g, label);
    if (n == 0) return fail("the voxel array has no fluid cell");
    const int nPad = (int)(((n + 31) / 32) * 32);
    const int nQ = latInfo(lattice).nQ;
    // one direction at a time: the host holds a single row of the table
    int32_t *d_table = nullptr, *d_label = nullptr;
    CUDA_OK(cudaMalloc(&d_table, (size_t)nQ * nPad * sizeof(int32_t)));
    std::vector<int32_t> row((size_t)nPad);
    int rc = 0;
    for (int q = 0; q < nQ && !rc; ++q) {
        int c[3];
        lookupC(lattice, q, c);
        std::fill(row.begin(), row.end(), -1);
        parallelOverX(nx, [&](int x0, int x1) { fillTableRow(g, label, c, row.data(), x0, x1); });
        if (cudaMemcpy(d_table + (size_t)q * nPad, row.data(), (size_t)nPad * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess)
            rc = fail("copy of the pull table to the device failed");
    }
    if (!rc) {
        for (int i = 0; i < nPad; ++i) row[(size_t)i] = i < n ? i + 1 : 0;
        if (cudaMalloc(&d_label, (size_t)nPad * sizeof(int32_t)) != cudaSuccess ||
            cudaMemcpy(d_label, row.data(), (size_t)nPad * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess)
            rc = fail("copy of the label array to the device failed");
    }
    if (!rc) rc = chimp_create_from_device_table(out, lattice, (int)n, nPad, 0, d_table, d_label, n_fields, index_form, device);
    cudaFree(d_table);
    cudaFree(d_label);
    return rc;
}

// Colour-gradient support tables of a two-field lattice from the same voxel array (twophase/main_TWOPHASE.cpp:280-284,
// LButilities.h:12-22): phi slot of neighbor(q, n) -- own fluid node -> its slot, solid cell next to a fluid cell (the
// reference's solid boundary nodes, LBgeometry.h:37-45; numbered in C-order) -> n_pad + k, anything else -> the zero
// slot n_pad + n_extra -- and the constant colour wall_phi[cell] of those solid cells.
int chimp_voxel_phi_table_host(int lattice, int nx, int ny, int nz, const uint8_t *voxels, int periodic_mask, const double *wall_phi,
                               int *n_extra, int32_t *ptable, double *phi_extra)
{
    VoxelGrid g;
    if (voxelGrid(g, lattice, nx, ny, nz, voxels, periodic_mask)) return 1;
    const int nQ = latInfo(lattice).nQ;
    std::vector<int32_t> label;
    const long long n = labelCells(g, label);
    if (n == 0) return fail("the voxel array has no fluid cell");
    const int nPad = (int)(((n + 31) / 32) * 32);
    // wall cells: solid with a fluid neighbour in any non-rest direction
    std::vector<int32_t> wallNo((size_t)g.cells(), -1);
    std::vector<uint8_t> isWall((size_t)g.cells(), 0);
    parallelOverX(nx, [&](int x0, int x1) {
        for (int x = x0; x < x1; ++x)
            for (int y = 0; y < g.ny; ++y)
                for (int z = 0; z < g.nz; ++z) {
                    const long long k = g.at(x, y, z);
                    if (g.v[k]) continue;
                    for (int q = 0; q < nQ - 1; ++q) {
                        int c[3];
                        lookupC(lattice, q, c);
                        const long long nb = g.shifted(x, y, z, c[0], c[1], c[2]);
                        if (nb >= 0 && g.v[nb]) { isWall[(size_t)k] = 1; break; }
                    }
                }
    });
    int nWall = 0;
    for (long long k = 0; k < g.cells(); ++k)
        if (isWall[(size_t)k]) wallNo[(size_t)k] = nWall++;
    if (n_extra) *n_extra = nWall;
    if (phi_extra) {
        if (!wall_phi) return fail("wall_phi is null");
        for (long long k = 0; k < g.cells(); ++k)
            if (isWall[(size_t)k]) phi_extra[wallNo[(size_t)k]] = wall_phi[k];
    }
    if (ptable) {
        const int zeroSlot = nPad + nWall;
        for (int q = 0; q < nQ; ++q) {
            int c[3];
            lookupC(lattice, q, c);
            int32_t *row = ptable + (size_t)q * nPad;
            std::fill(row, row + nPad, zeroSlot);
            parallelOverX(nx, [&](int x0, int x1) {
                for (int x = x0; x < x1; ++x)
                    for (int y = 0; y < g.ny; ++y)
                        for (int z = 0; z < g.nz; ++z) {
                            const int32_t me = label[(size_t)g.at(x, y, z)];
                            if (!me) continue;
                            const long long nb = g.shifted(x, y, z, c[0], c[1], c[2]);
                            if (nb < 0) continue;
                            if (label[(size_t)nb]) row[me - 1] = label[(size_t)nb] - 1;
                            else if (isWall[(size_t)nb]) row[me - 1] = nPad + wallNo[(size_t)nb];
                        }
            });
        }
    }
    return 0;
}

int chimp_set_phi_table_from_voxels(chimp_lattice *c, int nx, int ny, int nz, const uint8_t *voxels, int periodic_mask, const double *wall_phi)
{
    if (check(c, true)) return 1;
    if (c->nFields != 2) return fail("needs a two-field lattice");
    int nExtra = 0;
    if (chimp_voxel_phi_table_host(c->lattice, nx, ny, nz, voxels, periodic_mask, nullptr, &nExtra, nullptr, nullptr)) return 1;
    std::vector<int32_t> ptable((size_t)c->li.nQ * c->nPad);
    std::vector<double> extra((size_t)std::max(nExtra, 1), 0.0);
    if (chimp_voxel_phi_table_host(c->lattice, nx, ny, nz, voxels, periodic_mask, wall_phi, &nExtra, ptable.data(), extra.data())) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    int32_t *d_pt = nullptr;
    double *d_ex = nullptr;
    CUDA_OK(cudaMalloc(&d_pt, ptable.size() * sizeof(int32_t)));
    CUDA_OK(cudaMalloc(&d_ex, extra.size() * sizeof(double)));
    int rc = 0;
    if (cudaMemcpy(d_pt, ptable.data(), ptable.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_ex, extra.data(), extra.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
        rc = fail("copy of the phi table to the device failed");
    if (!rc) rc = chimp_set_phi_table_dev(c, d_pt, nExtra, d_ex);
    cudaFree(d_pt);
    cudaFree(d_ex);
    return rc;
}

// ---- z-slab of a decomposition from a voxel array (host code on N GPUs) --------------------------------------
// voxels_ext is the rank's own slab plus one layer of the neighbour below (z index 0) and above (z index nz + 1);
// x and y are periodic, the z neighbours are other ranks (a ring; with one rank they are the slab itself).  Same
// tables as the torch ingest (ingest.build_slab_tables, checked cell by cell in tests/test_voxel_ingest.py): own fluid
// cells carry the reference's per-rank labels 1..N (C-order of the own slab, vtklb.py:92-94); device slots put the
// two halo-coupled layers first and then go layer by layer; a population entering from a neighbour rank is pulled
// from a halo-in slot numbered per direction in C-order of the receiving cells.
extern "C++" {
namespace {
struct SlabTables {
    int n = 0, nPad = 0, nHalo = 0, nBoundary = 0;
    long long stride = 0;
    std::vector<int32_t> table, labels, slotOfCell; // slotOfCell: [nx*ny*nz] slot of an own cell, -1 solid
    std::vector<long long> send[2], recv[2];        // face 0 = down (towards z - 1), 1 = up
};

int buildSlabTables(int lattice, int nx, int ny, int nz, const uint8_t *ext, SlabTables &t)
{
    const LatInfo li = latInfo(lattice);
    if (li.nQ == 0) return fail("unknown lattice id %d", lattice);
    if (li.nD != 3) return fail("z-slabs need a 3-D lattice");
    if (!ext || nx < 1 || ny < 1 || nz < 2) return fail("bad slab voxel array (at least two own layers)");
    if ((long long)nx * ny * (nz + 2) >= (1ll << 31)) return fail("slab voxel array too large for 32-bit cell indices");
    const int nzE = nz + 2, nQ = li.nQ;
    auto E = [&](int x, int y, int ze) -> uint8_t { return ext[((size_t)x * ny + y) * nzE + ze]; }; // ze in [0, nz + 2)
    auto wrapX = [&](int x) { return (x % nx + nx) % nx; };
    auto wrapY = [&](int y) { return (y % ny + ny) % ny; };
    // labels in C-order of the own slab
    std::vector<int32_t> label((size_t)nx * ny * nz, 0);
    int n = 0;
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z)
                if (E(x, y, z + 1)) label[((size_t)x * ny + y) * nz + z] = ++n;
    if (n == 0) return fail("the slab has no fluid cell");
    t.n = n;
    t.nPad = ((n + 31) / 32) * 32;
    // slot order: layer 0, layer nz - 1, layers 1 .. nz - 2; inside a layer the C-order of (x, y)
    t.slotOfCell.assign((size_t)nx * ny * nz, -1);
    t.labels.assign((size_t)t.nPad, 0);
    int slot = 0;
    auto placeLayer = [&](int z) {
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y) {
                const size_t cell = ((size_t)x * ny + y) * nz + z;
                if (label[cell]) { t.slotOfCell[cell] = slot; t.labels[(size_t)slot] = label[cell]; ++slot; }
            }
    };
    placeLayer(0);
    placeLayer(nz - 1);
    t.nBoundary = slot;
    for (int z = 1; z < nz - 1; ++z) placeLayer(z);
    // halo elements per direction: receiving cells of my bottom layer (c_z = +1) / top layer (c_z = -1)
    std::vector<int> counts(nQ, 0);
    std::vector<std::vector<int32_t>> haloNo(nQ); // [q][x*ny + y] -> k or -1
    for (int q = 0; q < nQ; ++q) {
        int c[3];
        lookupC(lattice, q, c);
        if (c[2] == 0) continue;
        haloNo[q].assign((size_t)nx * ny, -1);
        const int zOwn = c[2] == 1 ? 0 : nz - 1, zHalo = c[2] == 1 ? 0 : nz + 1;
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y)
                if (E(x, y, zOwn + 1) && E(wrapX(x - c[0]), wrapY(y - c[1]), zHalo)) haloNo[q][(size_t)x * ny + y] = counts[q]++;
    }
    int maxCount = 0;
    for (int q = 0; q < nQ; ++q) maxCount = std::max(maxCount, counts[q]);
    t.nHalo = ((maxCount + 15) / 16) * 16;
    t.stride = (long long)t.nPad + t.nHalo;
    // pull table
    t.table.assign((size_t)nQ * t.nPad, -1);
    for (int q = 0; q < nQ; ++q) {
        int c[3];
        lookupC(lattice, q, c);
        int32_t *row = t.table.data() + (size_t)q * t.nPad;
        parallelOverX(nx, [&](int x0, int x1) {
            for (int x = x0; x < x1; ++x)
                for (int y = 0; y < ny; ++y)
                    for (int z = 0; z < nz; ++z) {
                        const int s = t.slotOfCell[((size_t)x * ny + y) * nz + z];
                        if (s < 0) continue;
                        const int zs = z - c[2];
                        if (zs < 0 || zs >= nz) { // from the neighbour rank: halo-in slot, or a wall there
                            const int k = haloNo[q][(size_t)x * ny + y];
                            row[s] = k >= 0 ? t.nPad + k : -1;
                        } else {
                            row[s] = t.slotOfCell[((size_t)wrapX(x - c[0]) * ny + wrapY(y - c[1])) * nz + zs];
                        }
                    }
        });
    }
    // faces: directions ascending, receiving cells in C-order of (x, y)
    for (int q = 0; q < nQ; ++q) {
        int c[3];
        lookupC(lattice, q, c);
        if (c[2] == 0) continue;
        const int recvFace = c[2] == 1 ? 0 : 1, sendFace = 1 - recvFace;
        for (int k = 0; k < counts[q]; ++k) t.recv[recvFace].push_back((long long)q * t.stride + t.nPad + k);
        // what the neighbour on the other side receives from me: its cell (x, y) of the adjacent layer is fluid and
        // my cell (x - cx, y - cy) of the layer facing it is fluid
        const int zMine = c[2] == 1 ? nz - 1 : 0, zTheirs = c[2] == 1 ? nz + 1 : 0;
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y) {
                if (!E(x, y, zTheirs)) continue;
                const int xs = wrapX(x - c[0]), ys = wrapY(y - c[1]);
                const int s = t.slotOfCell[((size_t)xs * ny + ys) * nz + zMine];
                if (s >= 0) t.send[sendFace].push_back((long long)q * t.stride + s);
            }
    }
    return 0;
}
} // namespace
} // extern "C++"

int chimp_slab_tables_host(int lattice, int nx, int ny, int nz_own, const uint8_t *voxels_ext, long long *info8, int32_t *table,
                           int32_t *labels, long long *send_down, long long *recv_down, long long *send_up, long long *recv_up)
{
    SlabTables t;
    if (buildSlabTables(lattice, nx, ny, nz_own, voxels_ext, t)) return 1;
    if (info8) {
        info8[0] = t.n; info8[1] = t.nPad; info8[2] = t.nHalo; info8[3] = t.nBoundary;
        info8[4] = (long long)t.send[0].size(); info8[5] = (long long)t.recv[0].size();
        info8[6] = (long long)t.send[1].size(); info8[7] = (long long)t.recv[1].size();
    }
    if (table) memcpy(table, t.table.data(), t.table.size() * sizeof(int32_t));
    if (labels) memcpy(labels, t.labels.data(), t.labels.size() * sizeof(int32_t));
    long long *dst[4] = {send_down, recv_down, send_up, recv_up};
    const std::vector<long long> *src[4] = {&t.send[0], &t.recv[0], &t.send[1], &t.recv[1]};
    for (int k = 0; k < 4; ++k)
        if (dst[k] && !src[k]->empty()) memcpy(dst[k], src[k]->data(), src[k]->size() * sizeof(long long));
    return 0;
}

int chimp_create_slab_from_voxels(chimp_lattice **out, int lattice, int nx, int ny, int nz_own, const uint8_t *voxels_ext, int n_fields,
                                  int index_form, int device, int rank_down, int rank_up)
{
    if (!out) return fail("out is null");
    *out = nullptr;
    if (n_fields != 1) return fail("z-slabs from voxels are offered for one-field lattices (two-field slabs: the device-table route)");
    SlabTables t;
    if (buildSlabTables(lattice, nx, ny, nz_own, voxels_ext, t)) return 1;
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) return fail("no CUDA device available: this engine has no CPU fallback");
    if (device < 0) CUDA_OK(cudaGetDevice(&device));
    CUDA_OK(cudaSetDevice(device));
    int32_t *d_table = nullptr, *d_label = nullptr;
    int rc = 0;
    if (cudaMalloc(&d_table, t.table.size() * sizeof(int32_t)) != cudaSuccess || cudaMalloc(&d_label, t.labels.size() * sizeof(int32_t)) != cudaSuccess ||
        cudaMemcpy(d_table, t.table.data(), t.table.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_label, t.labels.data(), t.labels.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess)
        rc = fail("copy of the slab tables to the device failed");
    if (!rc) rc = chimp_create_from_device_table(out, lattice, t.n, t.nPad, t.nHalo, d_table, d_label, n_fields, index_form, device);
    cudaFree(d_table);
    cudaFree(d_label);
    if (rc) return rc;
    chimp_lattice *c = *out;
    if (chimp_add_halo_face(c, rank_down, (long long)t.send[0].size(), t.send[0].data(), (long long)t.recv[0].size(), t.recv[0].data()) ||
        chimp_add_halo_face(c, rank_up, (long long)t.send[1].size(), t.send[1].data(), (long long)t.recv[1].size(), t.recv[1].data()) ||
        chimp_set_boundary_count(c, t.nBoundary)) {
        chimp_destroy(c);
        *out = nullptr;
        return 1;
    }
    return 0;
}

long long chimp_halo_face_recv_count(chimp_lattice *c, int k)
{
    if (!c || k < 0 || k >= (int)c->nbrs.size()) return -1;
    return (long long)c->nbrs[k].hRecv.size();
}
int chimp_halo_face_recv_list(chimp_lattice *c, int k, long long *recv_dst)
{
    if (!c || k < 0 || k >= (int)c->nbrs.size() || !recv_dst) return fail("bad halo face index");
    const auto &v = c->nbrs[k].hRecv;
    if (!v.empty()) memcpy(recv_dst, v.data(), v.size() * sizeof(long long));
    return 0;
}

int chimp_num_own_nodes(chimp_lattice *c) { return c ? c->n : 0; }
int chimp_host_table_info(chimp_lattice *c, long long *info6)
{
    if (!c || !c->hostBuilt) return fail("host tables not built");
    info6[0] = c->n; info6[1] = c->nPad; info6[2] = c->nHalo; info6[3] = c->stride; info6[4] = c->nBoundary; info6[5] = 0;
    return 0;
}
int chimp_host_table(chimp_lattice *c, int32_t *table, int32_t *labels, uint32_t *pmask)
{
    if (!c || !c->hostBuilt) return fail("host tables not built");
    for (int q = 0; q < c->li.nQ; ++q)
        memcpy(table + (size_t)q * c->n, c->hTable.data() + (size_t)q * c->nPad, (size_t)c->n * sizeof(int32_t));
    memcpy(labels, c->hLabel.data(), (size_t)c->n * sizeof(int32_t));
    memcpy(pmask, c->hPmask.data(), (size_t)c->n * sizeof(uint32_t));
    return 0;
}
int chimp_host_halo_lists(chimp_lattice *c, int k, long long *send_src, long long *recv_dst)
{
    if (!c || !c->hostBuilt) return fail("host tables not built");
    if (k < 0 || k >= (int)c->nbrs.size()) return fail("bad neighbour index");
    if (!c->hSendSrc[k].empty()) memcpy(send_src, c->hSendSrc[k].data(), c->hSendSrc[k].size() * sizeof(long long));
    if (!c->hRecvDst[k].empty()) memcpy(recv_dst, c->hRecvDst[k].data(), c->hRecvDst[k].size() * sizeof(long long));
    return 0;
}
int chimp_host_constant_links(chimp_lattice *c, long long *dst, double *values)
{
    if (!c || !c->hostBuilt) { fail("host tables not built"); return -1; }
    const size_t n = c->hConstDst.size();
    if (dst && n) memcpy(dst, c->hConstDst.data(), n * sizeof(long long));
    if (values && n) memcpy(values, c->constValues.data(), n * sizeof(double));
    return (int)n;
}
int chimp_host_scalar_halo_lists(chimp_lattice *c, int k, long long *send_src, long long *recv_dst)
{
    if (!c || !c->hostBuilt) return fail("host tables not built");
    if (k < 0 || k >= (int)c->hPhiSendSrc.size()) return fail("bad neighbour index");
    if (send_src && !c->hPhiSendSrc[k].empty()) memcpy(send_src, c->hPhiSendSrc[k].data(), c->hPhiSendSrc[k].size() * sizeof(long long));
    if (recv_dst && !c->hPhiRecvDst[k].empty()) memcpy(recv_dst, c->hPhiRecvDst[k].data(), c->hPhiRecvDst[k].size() * sizeof(long long));
    return 0;
}
double chimp_irregular_fraction(chimp_lattice *c)
{
    if (!c || c->indexForm != CHIMP_INDEX_COMPACT || c->nTiles == 0) return 0.0;
    return (double)c->nRows / ((double)c->nTiles * c->li.nQ);
}
double chimp_index_bytes_per_node(chimp_lattice *c)
{
    if (!c || c->n == 0) return 0.0;
    if (c->indexForm == CHIMP_INDEX_TABLE) return 4.0 * c->li.nQ;
    // delta words (minus the skipped ones when the skip-mask form is on) + bases + explicit rows
    const double skipped = c->skipMask && !c->onePhase ? 128.0 * (double)c->skippedWords : 0.0;
    return (4.0 * c->nWords * c->nPad - skipped + 16.0 * c->nTiles * c->nWords + 128.0 * c->nRows) / c->n;
}
int chimp_set_index_skip_mask(chimp_lattice *c, int on)
{
    if (check(c, true)) return 1;
    if (on && c->indexForm != CHIMP_INDEX_COMPACT) return fail("the skip mask belongs to the compact index form");
    c->skipMask = on != 0;
    return 0;
}
double chimp_index_skipped_word_fraction(chimp_lattice *c)
{
    if (!c || c->indexForm != CHIMP_INDEX_COMPACT || c->nTiles == 0) return 0.0;
    return (double)c->skippedWords / ((double)c->nTiles * c->nWords);
}
double chimp_one_phase_attribute_bytes_per_node(chimp_lattice *c)
{
    if (!c || !c->onePhase) return 0.0;
    return c->d_attr ? 4.0 : 24.0;
}
double chimp_phi_index_bytes_per_node(chimp_lattice *c)
{
    if (!c || c->n == 0 || !c->d_ptable) return 0.0;
    if (!c->phiDerived) return 4.0 * c->li.nQ;
    return 4.0 + 4.0 * (double)c->excWords / c->n;
}
long long chimp_plane_stride(chimp_lattice *c) { return c ? c->stride : 0; }

} // extern "C"
