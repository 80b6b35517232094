// oracle/mpi_shim/mpi.h -- TEST INFRASTRUCTURE, not product code.
//
// In-process stand-in for the handful of MPI calls the reference headers use
// (LBbndmpi.h:225-308, LBmonlatmpi.h:181-297, the Allreduce calls in the mains).
// No MPI is installed in the build container, so "ranks" are std::threads of one
// process: every rank thread has a thread_local rank id, point-to-point messages go
// through unbounded mailboxes keyed by (source, dest, tag) -- a send never blocks,
// which is a superset of the blocking semantics the reference relies on -- and
// Allreduce is a two-phase barrier with a fixed rank-order summation.
//
// Only what the oracle driver needs is implemented.  Datatype handles carry their
// byte size, reductions support int/double SUM and MAX.
#ifndef ORACLE_MPI_SHIM_H
#define ORACLE_MPI_SHIM_H

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
// io/VTK.h:100,1220,1288-1303 (parallel .pvtu piece records): MPI-IO on one node is positional POSIX I/O
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct { int fd; long long pos; } MPI_File;
#define MPI_LONG 8
#define MPI_CHAR 1
#define MPI_MODE_APPEND 1
#define MPI_MODE_WRONLY 2
#define MPI_INFO_NULL 0
#define MPI_SEEK_SET 0

namespace mpishim {

struct World {
    int nranks = 1;
    std::mutex mtx;
    std::condition_variable cv;
    std::map<std::tuple<int, int, int>, std::deque<std::vector<char>>> box;
    // allreduce state
    int arrived = 0;
    long generation = 0;
    std::vector<std::vector<char>> contrib;
    std::vector<char> result;
};

inline World &world()
{
    static World w;
    return w;
}

inline int &my_rank()
{
    static thread_local int r = 0;
    return r;
}

inline void init(int nranks)
{
    World &w = world();
    w.nranks = nranks;
    w.contrib.assign(nranks, {});
}

} // namespace mpishim

inline int MPI_Init(int *, char ***) { return 0; }
inline int MPI_Finalize() { return 0; }
inline int MPI_Comm_size(MPI_Comm, int *n) { *n = mpishim::world().nranks; return 0; }
inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = mpishim::my_rank(); return 0; }

inline int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm)
{
    mpishim::World &w = mpishim::world();
    std::vector<char> msg((size_t)count * dt);
    if (!msg.empty()) std::memcpy(msg.data(), buf, msg.size());
    {
        std::lock_guard<std::mutex> lk(w.mtx);
        w.box[std::make_tuple(mpishim::my_rank(), dest, tag)].push_back(std::move(msg));
    }
    w.cv.notify_all();
    return 0;
}

inline int MPI_Recv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm, MPI_Status *)
{
    mpishim::World &w = mpishim::world();
    std::unique_lock<std::mutex> lk(w.mtx);
    auto key = std::make_tuple(source, mpishim::my_rank(), tag);
    w.cv.wait(lk, [&] { auto it = w.box.find(key); return it != w.box.end() && !it->second.empty(); });
    std::vector<char> msg = std::move(w.box[key].front());
    w.box[key].pop_front();
    size_t n = (size_t)count * dt;
    if (msg.size() < n) n = msg.size();
    if (n) std::memcpy(buf, msg.data(), n);
    return 0;
}

inline int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm)
{
    mpishim::World &w = mpishim::world();
    const size_t nbytes = (size_t)count * dt;
    std::unique_lock<std::mutex> lk(w.mtx);
    const long gen = w.generation;
    w.contrib[mpishim::my_rank()].assign((const char *)sendbuf, (const char *)sendbuf + nbytes);
    if (++w.arrived == w.nranks) {
        // last rank in: reduce in rank order 0..P-1 (deterministic)
        w.result = w.contrib[0];
        for (int r = 1; r < w.nranks; ++r) {
            for (int i = 0; i < count; ++i) {
                if (dt == MPI_DOUBLE) {
                    double a, b;
                    std::memcpy(&a, w.result.data() + 8 * (size_t)i, 8);
                    std::memcpy(&b, w.contrib[r].data() + 8 * (size_t)i, 8);
                    a = (op == MPI_SUM) ? a + b : (a > b ? a : b);
                    std::memcpy(w.result.data() + 8 * (size_t)i, &a, 8);
                } else {
                    int a, b;
                    std::memcpy(&a, w.result.data() + 4 * (size_t)i, 4);
                    std::memcpy(&b, w.contrib[r].data() + 4 * (size_t)i, 4);
                    a = (op == MPI_SUM) ? a + b : (a > b ? a : b);
                    std::memcpy(w.result.data() + 4 * (size_t)i, &a, 4);
                }
            }
        }
        w.arrived = 0;
        ++w.generation;
        w.cv.notify_all();
    } else {
        w.cv.wait(lk, [&] { return w.generation != gen; });
    }
    // every rank copies the result out before anyone can start the next reduction's
    // "last rank in" phase: result is only rewritten when all nranks have arrived again,
    // which needs this rank to have left this call.
    std::memcpy(recvbuf, w.result.data(), nbytes);
    return 0;
}

inline int MPI_Barrier(MPI_Comm c)
{
    int a = 0, b = 0;
    return MPI_Allreduce(&a, &b, 1, MPI_INT, MPI_SUM, c);
}

inline int MPI_Initialized(int *flag) { *flag = 1; return 0; }

// root's buffer to every rank: an all-reduce in which only the root contributes (bytes are OR-ed as
// ints/doubles would not be exact for arbitrary payloads, so the payload travels through the mailbox)
inline int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm c)
{
    const int me = mpishim::my_rank(), n = mpishim::world().nranks;
    if (me == root) {
        for (int r = 0; r < n; ++r)
            if (r != root) MPI_Send(buf, count, dt, r, 9001, c);
    } else {
        MPI_Recv(buf, count, dt, root, 9001, c, MPI_STATUS_IGNORE);
    }
    return 0;
}

#include <fcntl.h>
#include <unistd.h>
inline int MPI_File_open(MPI_Comm, const char *name, int, MPI_Info, MPI_File *f)
{
    f->fd = ::open(name, O_WRONLY);
    f->pos = 0;
    return f->fd < 0;
}
inline int MPI_File_seek(MPI_File &f, MPI_Offset off, int) { f.pos = off; return 0; }
inline int MPI_File_write(MPI_File &f, const void *buf, int count, MPI_Datatype dt, MPI_Status *)
{
    const ssize_t n = ::pwrite(f.fd, buf, (size_t)count * dt, f.pos);
    f.pos += n > 0 ? n : 0;
    return n < 0;
}
inline int MPI_File_close(MPI_File *f) { return ::close(f->fd); }

#endif // ORACLE_MPI_SHIM_H
