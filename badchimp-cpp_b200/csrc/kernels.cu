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
    double v = 0.0;
    for (int b = threadIdx.x; b < nBlocks; b += blockDim.x) v += partial[(long long)l * nBlocks + b];
    const double s = blockSum256(v, sh);
    if (threadIdx.x == 0) {
        mass[l] = s;
        src[l] = 0.9 * 2 * scale[l] * s;
    }
}

// ---- index compression (IDX_RANK) ------------------------------------------------------
// One warp per (tile, q).  A pair is regular when every non-bounce lane l satisfies
// T[q][32*tile + l] == base + (number of non-bounce lanes below l).
__global__ void classifyTilesKernel(const int32_t *__restrict__ table, int n, int nPad, int nQ, int nTiles,
                                    int32_t *__restrict__ base, uint32_t *__restrict__ bbmask, int *irregularCount)
{
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (warp >= (long long)nTiles * nQ) return;
    const int q = (int)(warp / nTiles), tile = (int)(warp % nTiles);
    const int i = tile * 32 + lane;
    const bool live = i < n;
    const int t = live ? table[(long long)q * nPad + i] : -1;
    const bool bounce = (t == -1);
    const unsigned bb = __ballot_sync(0xffffffffu, bounce);
    const unsigned nb = ~bb;
    int b0 = 0;
    bool regular = true;
    if (nb) {
        const int first = __ffs(nb) - 1;
        b0 = __shfl_sync(0xffffffffu, t, first);
        const int expect = b0 + __popc(nb & ((1u << lane) - 1u));
        regular = __all_sync(0xffffffffu, bounce || t == expect);
    }
    if (live && bounce) atomicOr(bbmask + i, 1u << q);
    if (lane == 0) {
        if (regular) base[(long long)q * nTiles + tile] = b0;
        else base[(long long)q * nTiles + tile] = -(atomicAdd(irregularCount, 1) + 1);
    }
}

__global__ void fillRowsKernel(const int32_t *__restrict__ table, int n, int nPad, int nQ, int nTiles,
                               const int32_t *__restrict__ base, int32_t *__restrict__ rows)
{
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (warp >= (long long)nTiles * nQ) return;
    const int q = (int)(warp / nTiles), tile = (int)(warp % nTiles);
    const int b = base[(long long)q * nTiles + tile];
    if (b >= 0) return;
    const int i = tile * 32 + lane;
    const int t = (i < n) ? table[(long long)q * nPad + i] : 0;
    rows[((long long)(-b - 1) << 5) + lane] = t < 0 ? 0 : t;
}

} // namespace chimp
