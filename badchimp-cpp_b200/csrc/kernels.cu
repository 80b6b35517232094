// Non-templated helper kernels (layout conversion, halo pack/unpack, index compression).
#include "kernels.cuh"

namespace chimp {

__global__ void planesToAosKernel(double *aos, const double *__restrict__ planes, const int32_t *__restrict__ label,
                                  int n, int nPad, int nComp, int aosStride, int aosOffset)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long row = (long long)label[i] * aosStride + aosOffset;
    for (int c = 0; c < nComp; ++c) aos[row + c] = planes[(long long)c * nPad + i];
}

__global__ void haloPackKernel(double *__restrict__ buf, const double *__restrict__ X,
                               const long long *__restrict__ src, int count)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) buf[k] = X[src[k]];
}

__global__ void haloUnpackKernel(double *X, const double *__restrict__ buf, const long long *__restrict__ dst,
                                 int count)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) X[dst[k]] = buf[k];
}

__global__ void haloPushKernel(double *peerX, const double *__restrict__ X, const long long *__restrict__ src,
                               const long long *__restrict__ dst, int count, int nFields, long long fieldStride,
                               long long peerFieldStride, unsigned *blockCounter, unsigned long long *peerFlag,
                               unsigned long long value)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) {
        const long long s = src[k], d = dst[k];
        for (int f = 0; f < nFields; ++f) peerX[d + f * peerFieldStride] = X[s + f * fieldStride];
    }
    // every thread's remote stores are ordered before the flag: fence, count blocks, last one signals
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(blockCounter, 1u);
        if (done == gridDim.x - 1) {
            *blockCounter = 0u;
            __threadfence_system();
            *(volatile unsigned long long *)peerFlag = value;
            __threadfence_system();
        }
    }
}

__global__ void invertLabelsKernel(const int32_t *__restrict__ label, int n, int32_t *__restrict__ slotOf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slotOf[label[i]] = i;
}

__global__ void nodeFluxKernel(const int32_t *__restrict__ nodes, int count, const int32_t *__restrict__ slotOf, int nLabels,
                               const double *__restrict__ rho, const double *__restrict__ velComp, double *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int node = nodes[k];
    const int s = (node >= 0 && node < nLabels) ? slotOf[node] : -1;
    // rows of the reference fields that no own node writes stay zero-initialised (LBfield.h:75)
    out[k] = s >= 0 ? velComp[s] * rho[s] : 0.0;
}

// Global sums over peer memory (replaces MPI_Allreduce of one double, main_TWOPHASE.cpp:299): every rank stores its
// local sum into slot `rank` of every rank's mailbox (value, fence, sequence number) ...
struct MailSlotDev { double value; unsigned long long seq; };
// fold of the per-block partials (fixed order, as fluxForceKernel) and delivery of the local sum in one launch
__global__ void foldAndPushKernel(const double *__restrict__ partial, int nBlocks, double *sumOut, void *const *peerMail, int rank,
                                  int world, int parity, unsigned long long seq)
{
    __shared__ double sh[8];
    __shared__ double total;
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.0;
    int b = threadIdx.x;
    for (; b + 15 * (int)blockDim.x < nBlocks; b += 16 * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] += partial[b + k * blockDim.x];
    }
    for (; b < nBlocks; b += blockDim.x) v[0] += partial[b];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += v[k + 8];
    const double v0 = v[0] + v[4], v1 = v[1] + v[5], v2 = v[2] + v[6], v3 = v[3] + v[7];
    const double s = blockSum256((v0 + v1) + (v2 + v3), sh);
    if (threadIdx.x == 0) {
        *sumOut = s;
        total = s;
    }
    __syncthreads();
    const int w = threadIdx.x;
    if (w < world) {
        MailSlotDev *slot = (MailSlotDev *)peerMail[w] + parity * 64 + rank;
        *(volatile double *)&slot->value = total;
        __threadfence_system();
        *(volatile unsigned long long *)&slot->seq = seq;
    }
}

// ... and, once all `world` contributions of this step have arrived, adds them in rank order -- every rank
// obtains the same bits -- and forms F_x = 2 (momx - sum / nGlobal) (main_TWOPHASE.cpp:301-308)
__global__ void sumWaitFoldKernel(const void *mail, int world, int parity, unsigned long long seq, double momx, double nGlobal,
                                  double *sumOut, double *forceX, const unsigned long long *flags, unsigned flagMask,
                                  unsigned long long flagExpect, unsigned long long timeoutNs, unsigned *error)
{
    const MailSlotDev *slots = (const MailSlotDev *)mail + parity * 64;
    const int w = threadIdx.x;
    if (w < world) {
        awaitCounter(&slots[w].seq, seq, timeoutNs, error, 2u);
    } else if (w >= 64 && ((flagMask >> (w - 64)) & 1u)) {
        // threads 64.. also wait for the arrival counters of the scalar halo (one launch instead of three)
        awaitCounter(flags + (w - 64), flagExpect, timeoutNs, error, 3u);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < world; ++k) s += *(const volatile double *)&slots[k].value;
        *sumOut = s;
        double mean = s;
        mean /= nGlobal;
        *forceX = 2 * (momx - mean);
    }
}

// waits until every arrival counter selected by `mask` (bit k -> flags[k]) has reached `expect`
__global__ void waitFlagsKernel(const unsigned long long *flags, unsigned mask, unsigned long long expect, unsigned long long timeoutNs,
                                unsigned *error)
{
    const unsigned k = threadIdx.x;
    if ((mask >> k) & 1u) awaitCounter(flags + k, expect, timeoutNs, error, 1u);
    __threadfence_system();
}

__global__ void fillKernel(double *p, double v, long long count)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) p[k] = v;
}

__global__ void massFinalizeKernel(const double *__restrict__ partial, int nBlocks, const double *__restrict__ scale,
                                   double *mass, double *src)
{
    __shared__ double sh[8];
    const int l = blockIdx.x;
    if (nBlocks < 0) { // partial[] already holds the (all-reduced) sums
        if (threadIdx.x == 0) src[l] = 0.9 * 2 * scale[l] * partial[l];
        return;
    }
    double v = 0.0;
    for (int b = threadIdx.x; b < nBlocks; b += blockDim.x) v += partial[(long long)l * nBlocks + b];
    const double s = blockSum256(v, sh);
    if (threadIdx.x == 0) {
        mass[l] = s;
        src[l] = 0.9 * 2 * scale[l] * s;
    }
}

// Folds the per-block x-momentum partials in a fixed order.  finish == 0: only the local sum is
// written (an all-reduce over ranks follows); finish == 1: partial[] already holds the global
// sum(s) and F_x = 2 (momx - sum / nGlobal) is produced (main_TWOPHASE.cpp:299-308).
__global__ void fluxForceKernel(const double *partial, int nBlocks, double momx, double nGlobal,
                                double *sumOut, double *forceX, int finish)
{
    __shared__ double sh[8];
    // sixteen independent accumulators per thread keep sixteen loads in flight (the kernel is one block and purely
    // latency-bound); the summation order is fixed by (thread, slot), so the result is deterministic
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.0;
    int b = threadIdx.x;
    for (; b + 15 * (int)blockDim.x < nBlocks; b += 16 * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] += partial[b + k * blockDim.x];
    }
    for (; b < nBlocks; b += blockDim.x) v[0] += partial[b];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += v[k + 8];
    const double v0 = v[0] + v[4], v1 = v[1] + v[5], v2 = v[2] + v[6], v3 = v[3] + v[7];
    const double s = blockSum256((v0 + v1) + (v2 + v3), sh);
    if (threadIdx.x == 0) {
        *sumOut = s;
        if (finish) {
            double mean = s;
            mean /= nGlobal;
            *forceX = 2 * (momx - mean);
        }
    }
}

__global__ void foldRowsKernel(const double *__restrict__ partial, int nBlocks, double *out)
{
    __shared__ double sh[8];
    const double *row = partial + (long long)blockIdx.x * nBlocks;
    double v = 0.0;
    for (int b = threadIdx.x; b < nBlocks; b += blockDim.x) v += row[b];
    const double s = blockSum256(v, sh);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// ---- index compression (IDX_COMPACT) ---------------------------------------------------
// One warp per (tile, q): base = smallest non-bounce source of the tile; a pair whose
// sources span more than 253 keeps an explicit row.
__global__ void classifyTilesKernel(const int32_t *__restrict__ table, int n, int nPad, int nQ, int nTiles,
                                    int32_t *__restrict__ base, uint8_t *__restrict__ deltaBytes, int *irregularCount)
{
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (warp >= (long long)nTiles * nQ) return;
    const int q = (int)(warp / nTiles), tile = (int)(warp % nTiles);
    const int i = tile * 32 + lane;
    const bool live = i < n;
    const int t = live ? table[(long long)q * nPad + i] : -1;
    const bool bounce = (t < 0);
    int mn = bounce ? 0x7fffffff : t, mx = bounce ? -1 : t;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const bool any = mx >= 0;
    const bool regular = !any || (mx - mn) <= 253;
    const int b0 = any ? mn : 0;
    // byte (q & 3) of word (q >> 2) of node i
    deltaBytes[((long long)(q >> 2) * nPad + i) * 4 + (q & 3)] = (uint8_t)((bounce || !regular) ? 255 : (t - b0));
    if (lane == 0) {
        const int nSlots = ((nQ + 3) / 4) * 4; // tile-major, padded to whole int4 groups
        if (regular) base[(long long)tile * nSlots + q] = b0;
        else base[(long long)tile * nSlots + q] = -(atomicAdd(irregularCount, 1) + 1);
    }
}

// Skip mask of IDX_COMPACT_MASK, one thread per tile: bit w of the tile's spare base slot (4 NW - 1, never a direction)
// is set when every live lane of the tile has byte == lane in all directions of delta word w -- the sources of those
// directions are 32 consecutive slots starting at the base, and the step kernel need not read the word.
__global__ void skipMaskKernel(const uint32_t *__restrict__ delta, int n, int nPad, int nQ, int nTiles, int32_t *__restrict__ base,
                               unsigned long long *skippedWords)
{
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= nTiles) return;
    const int nW = (nQ + 3) / 4, nSlots = nW * 4;
    unsigned mask = 0;
    for (int w = 0; w < nW; ++w) {
        const int dirs = min(4, nQ - 4 * w);                       // the last word may hold fewer than four directions
        const uint32_t cmp = dirs == 4 ? 0xffffffffu : ((1u << (8 * dirs)) - 1u);
        bool run = true;
        for (int lane = 0; lane < 32 && run; ++lane) {
            const int i = tile * 32 + lane;
            if (i >= n) break;
            run = ((delta[(long long)w * nPad + i] ^ ((uint32_t)lane * 0x01010101u)) & cmp) == 0u;
        }
        if (run) mask |= 1u << w;
    }
    base[(long long)tile * nSlots + nSlots - 1] = (int32_t)mask;
    if (mask) atomicAdd(skippedWords, (unsigned long long)__popc(mask));
}

// kernel form of the pull table: -1 (reversed own slot) becomes i + (rev q - q) * stride
__global__ void kernelTableKernel(const int32_t *__restrict__ table, int32_t *__restrict__ ktable, int n, int nPad,
                                  int nQ, int nPairs, long long stride)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (long long)nQ * nPad) return;
    const int q = (int)(k / nPad), i = (int)(k % nPad);
    const int rq = q == nQ - 1 ? q : (q + nPairs) % (nQ - 1);
    const int t = i < n ? table[k] : i;
    ktable[k] = t >= 0 ? t : (int)(i + (long long)(rq - q) * stride);
}

__global__ void fillRowsKernel(const int32_t *__restrict__ table, int n, int nPad, int nQ, int nTiles,
                               const int32_t *__restrict__ base, int32_t *__restrict__ rows, int nPairs, long long stride)
{
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (warp >= (long long)nTiles * nQ) return;
    const int q = (int)(warp / nTiles), tile = (int)(warp % nTiles);
    const int b = base[(long long)tile * (((nQ + 3) / 4) * 4) + q];
    if (b >= 0) return;
    const int i = tile * 32 + lane;
    const int rq = q == nQ - 1 ? q : (q + nPairs) % (nQ - 1);
    const int t = (i < n) ? table[(long long)q * nPad + i] : i;
    rows[((long long)(-b - 1) << 5) + lane] = t >= 0 ? t : (int)(i + (long long)(rq - q) * stride);
}

} // namespace chimp
