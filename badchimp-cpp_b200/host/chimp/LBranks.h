// InProcessRanks: all ranks of a decomposition as engine contexts in ONE process, one host thread per
// rank -- the arrangement used when MPI is only the launcher (or absent).  It supplies the three
// transport callbacks of include/chimp_b200.h (population halos, scalar halos, all-reduce) by
// rendezvous on a barrier and device-to-device copies, i.e. what MPI_Send/MPI_Recv
// (LBmonlatmpi.h:253-257) and MPI_Allreduce do between the reference's ranks.  With one GPU per rank
// in separate processes the same callbacks are bound to NCCL (badchimp-cpp_b200/multi.py) or replaced
// by the engine's peer stores (chimp_connect_peer).
#ifndef CHIMP_LBRANKS_H
#define CHIMP_LBRANKS_H

#include <cuda_runtime_api.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "../../../include/chimp_b200.h"
#include "LBglobal.h"

class InProcessRanks
{
public:
    // rankOf[k]: global rank number of handles[k] (the numbers chimp_neighbor_info reports)
    InProcessRanks(const std::vector<chimp_lattice *> &handles, const std::vector<int> &rankOf)
        : h_(handles), rankOf_(rankOf), shared_(handles.size()), ctx_(handles.size())
    {
        for (std::size_t r = 0; r < h_.size(); ++r) {
            ctx_[r] = {this, int(r)};
            if (chimp_set_exchange_callback(h_[r], &InProcessRanks::exchangeCb, &ctx_[r]) ||
                chimp_set_scalar_exchange_callback(h_[r], &InProcessRanks::scalarCb, &ctx_[r]) ||
                chimp_set_allreduce_callback(h_[r], &InProcessRanks::allreduceCb, &ctx_[r]))
                chimp_host::die(std::string("GPU engine: ") + chimp_last_error());
        }
    }

    // runs fn(r) for every rank on its own thread and joins them
    void run(const std::function<void(int)> &fn)
    {
        if (h_.size() == 1) { fn(0); return; }
        std::vector<std::thread> threads;
        for (std::size_t r = 0; r < h_.size(); ++r) threads.emplace_back([&fn, r] { fn(int(r)); });
        for (auto &t : threads) t.join();
    }

    // MPI_Allreduce(MPI_SUM) of host values between the rank threads, added in rank order
    void allreduceHost(int r, double *vals, int count)
    {
        shared_[r].assign(vals, vals + count);
        barrier();
        std::vector<double> total(shared_[0]);
        for (std::size_t k = 1; k < h_.size(); ++k)
            for (int i = 0; i < count; ++i) total[i] = total[i] + shared_[k][i];
        barrier();
        std::copy(total.begin(), total.end(), vals);
    }

private:
    struct Ctx { InProcessRanks *self; int r; };

    void barrier()
    {
        std::unique_lock<std::mutex> lk(m_);
        const long gen = gen_;
        if (++waiting_ == int(h_.size())) {
            waiting_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen_ != gen; });
        }
    }
    int localIndex(int globalRank) const
    {
        for (std::size_t k = 0; k < rankOf_.size(); ++k)
            if (rankOf_[k] == globalRank) return int(k);
        chimp_host::die("neighbour rank " + std::to_string(globalRank) + " is not in this process");
    }
    int exchange(int r, void *stream, bool scalar)
    {
        if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return 1;
        barrier(); // every rank's send buffers are complete
        chimp_lattice *me = h_[r];
        int rc = 0;
        for (int k = 0; k < chimp_num_neighbors(me); ++k) {
            int nr = -1;
            long long ns = 0, nrecv = 0;
            if (chimp_neighbor_info(me, k, &nr, &ns, &nrecv)) { rc = 1; break; }
            if (scalar && chimp_scalar_neighbor_info(me, k, &ns, &nrecv)) { rc = 1; break; }
            chimp_lattice *peer = h_[localIndex(nr)];
            for (int j = 0; j < chimp_num_neighbors(peer); ++j) {
                int pr = -1;
                chimp_neighbor_info(peer, j, &pr, nullptr, nullptr);
                if (pr != rankOf_[r] || !nrecv) continue;
                const void *src = scalar ? chimp_scalar_send_buffer_dev(peer, j) : chimp_send_buffer_dev(peer, j);
                void *dst = scalar ? chimp_scalar_recv_buffer_dev(me, k) : chimp_recv_buffer_dev(me, k);
                // population counts reported by chimp_neighbor_info already cover all LbFields.  The copy goes onto the
                // stream the engine handed over: the unpack kernels it enqueues there afterwards are ordered behind it (the
                // engine's streams are non-blocking, so a copy on the legacy default stream would not be)
                if (cudaMemcpyAsync(dst, src, std::size_t(nrecv) * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess) rc = 1;
            }
        }
        // my copies have read the peers' send buffers completely before anybody may overwrite them
        if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = 1;
        barrier();
        return rc;
    }
    int allreduce(int r, void *dev, int count, void *stream)
    {
        // everything on the engine's stream: the kernels that consume the sum are enqueued there after this call returns
        cudaStream_t s = (cudaStream_t)stream;
        std::vector<double> mine(count);
        if (cudaMemcpyAsync(mine.data(), dev, std::size_t(count) * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess) return 1;
        if (cudaStreamSynchronize(s) != cudaSuccess) return 1;
        allreduceHost(r, mine.data(), count);
        if (cudaMemcpyAsync(dev, mine.data(), std::size_t(count) * sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess) return 1;
        return cudaStreamSynchronize(s) != cudaSuccess; // `mine` leaves scope: the copy must have read it
    }
    static int exchangeCb(void *u, void *s) { auto *c = (Ctx *)u; return c->self->exchange(c->r, s, false); }
    static int scalarCb(void *u, void *s) { auto *c = (Ctx *)u; return c->self->exchange(c->r, s, true); }
    static int allreduceCb(void *u, double *dev, int count, void *s) { auto *c = (Ctx *)u; return c->self->allreduce(c->r, dev, count, s); }

    std::vector<chimp_lattice *> h_;
    std::vector<int> rankOf_;
    std::vector<std::vector<double>> shared_;
    std::vector<Ctx> ctx_;
    std::mutex m_;
    std::condition_variable cv_;
    int waiting_ = 0;
    long gen_ = 0;
};

#endif
