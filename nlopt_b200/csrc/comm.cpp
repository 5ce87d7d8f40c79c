// comm.cpp -- see comm.hpp.  NCCL entry points are resolved with dlsym; the handful of enum
// values used (ncclFloat64 = 8, ncclSum = 0, ncclSuccess = 0) are ABI constants of nccl.h.
#include "comm.hpp"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <chrono>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/nlopt_b200.h"

namespace nb200 {

namespace {
struct NcclId { char internal[128]; };
typedef int (*fn_get_id)(NcclId *);
typedef int (*fn_init_rank)(void **, int, NcclId, int);
typedef int (*fn_destroy)(void *);
typedef int (*fn_all_gather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*fn_errstr)(int);

struct Api {
    void *lib = nullptr;
    fn_get_id get_id = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_errstr errstr = nullptr;
};

Api &api() { static Api a; return a; }

bool load(std::string *err)
{
    Api &a = api();
    if (a.lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names)
        if ((a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!a.lib) { if (err) *err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
    a.get_id = (fn_get_id) dlsym(a.lib, "ncclGetUniqueId");
    a.init_rank = (fn_init_rank) dlsym(a.lib, "ncclCommInitRank");
    a.destroy = (fn_destroy) dlsym(a.lib, "ncclCommDestroy");
    a.all_gather = (fn_all_gather) dlsym(a.lib, "ncclAllGather");
    a.all_reduce = (fn_all_reduce) dlsym(a.lib, "ncclAllReduce");
    a.errstr = (fn_errstr) dlsym(a.lib, "ncclGetErrorString");
    if (!a.get_id || !a.init_rank || !a.destroy || !a.all_gather || !a.all_reduce) {
        if (err) *err = "libnccl is missing expected symbols";
        return false;
    }
    return true;
}

std::string nccl_msg(const char *what, int code)
{
    std::string s = what;
    s += ": ";
    s += api().errstr ? api().errstr(code) : "NCCL error";
    return s;
}
}  // namespace

Comm &Comm::instance() { static Comm c; return c; }

int Comm::unique_id(unsigned char id[128], std::string *err)
{
    if (!load(err)) return -1;
    NcclId nid;
    int rc = api().get_id(&nid);
    if (rc) { if (err) *err = nccl_msg("ncclGetUniqueId", rc); return -1; }
    std::memcpy(id, nid.internal, 128);
    return 0;
}

int Comm::init(const unsigned char id[128], int r, int w, int dev, std::string *err)
{
    if (w < 1 || r < 0 || r >= w || (8 % w) != 0) {
        if (err) *err = "world size must be 1, 2, 4 or 8 (it has to divide the 8 virtual shards)";
        return -1;
    }
    if (comm_) finalize();
    rank = r; world = w; device = dev;
    seq_counter_ = 0;
    if (w == 1) return 0;
    if (!load(err)) return -1;
    if (cudaSetDevice(dev) != cudaSuccess) { if (err) *err = "cudaSetDevice failed"; return -1; }
    NcclId nid;
    std::memcpy(nid.internal, id, 128);
    int rc = api().init_rank(&comm_, w, nid, r);
    if (rc) { comm_ = nullptr; world = 1; rank = 0; if (err) *err = nccl_msg("ncclCommInitRank", rc); return -1; }
    const char *ex = std::getenv("NLOPT_B200_EXCHANGE");
    force_nccl_ = ex && std::strcmp(ex, "nccl") == 0;
    std::string perr;
    const int prc = setup_p2p(&perr);
    if (prc == -2) {             // a bootstrap collective itself failed: the communicator is unusable
        teardown_p2p();
        if (err) *err = "mailbox bootstrap: " + perr;
        return -1;
    }
    if (prc != 0) {              // agreed by ALL ranks (see setup_p2p): everyone takes the NCCL path
        std::fprintf(stderr, "nlopt_b200: peer mailboxes unavailable (%s); using NCCL all-gather per dual evaluation\n",
                     perr.c_str());
        teardown_p2p();
    }
    return 0;
}

// Allocate this rank's mailbox, all-gather the CUDA IPC handles over NCCL, map every peer's mailbox.
// The outcome is COLLECTIVE: a local failure (allocation, IPC export, IPC import of some peer) does not
// return early -- every rank still takes part in both collectives below, and the mailbox path is enabled
// only if the closing all-reduce shows that all ranks succeeded.  (A per-rank decision would leave some ranks
// launching the mailbox kernel while others wait in ncclAllGather: a deadlock.)
int Comm::setup_p2p(std::string *err)
{
    p2p_ready = false;
    bool ok = true;
    std::string why;
    auto note = [&](const char *what, cudaError_t ce) {
        if (ok) why = std::string(what) + ": " + cudaGetErrorString(ce);
        ok = false;
        cudaGetLastError();
    };
    cudaError_t ce;
    if ((ce = cudaMalloc(&box_local_, kBoxDoubles * sizeof(double))) != cudaSuccess) { box_local_ = nullptr; note("cudaMalloc mailbox", ce); }
    if (box_local_) cudaMemset(box_local_, 0, kBoxDoubles * sizeof(double));
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof mine);
    if (ok && (ce = cudaIpcGetMemHandle(&mine, box_local_)) != cudaSuccess) note("cudaIpcGetMemHandle", ce);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    const size_t per = sizeof(cudaIpcMemHandle_t) / sizeof(double) + 1;      // handle + "this rank is fine so far"
    double *stage = nullptr;
    if (cudaMalloc(&stage, (size_t) world * per * sizeof(double)) != cudaSuccess) {
        // without even this buffer the rank cannot join a collective at all: nothing sensible is left to do
        *err = "cudaMalloc of the bootstrap buffer failed";
        return -2;
    }
    double rec[sizeof(cudaIpcMemHandle_t) / sizeof(double) + 1];
    std::memcpy(rec, &mine, sizeof mine);
    rec[per - 1] = ok ? 1.0 : 0.0;
    cudaMemcpy(stage + (size_t) rank * per, rec, per * sizeof(double), cudaMemcpyHostToDevice);
    std::string e2;
    if (all_gather_inplace(stage, per, 0, &e2) != 0) { *err = e2; cudaFree(stage); return -2; }
    cudaStreamSynchronize(0);
    std::vector<double> all((size_t) world * per);
    cudaMemcpy(all.data(), stage, all.size() * sizeof(double), cudaMemcpyDeviceToHost);
    bool everyone = ok;
    for (int r = 0; r < world; ++r) everyone = everyone && all[(size_t) r * per + per - 1] == 1.0;
    if (everyone)
        for (int r = 0; r < world; ++r) {
            if (r == rank) { box_peer[r] = box_local_; continue; }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, &all[(size_t) r * per], sizeof h);
            void *p = nullptr;
            if ((ce = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) { note("cudaIpcOpenMemHandle", ce); break; }
            box_peer[r] = (double *) p;
        }
    // closing collective: (a) nobody may start writing before everyone has mapped and zeroed, (b) the verdict
    double bad = ok ? 0.0 : 1.0;
    cudaMemcpy(stage, &bad, sizeof bad, cudaMemcpyHostToDevice);
    if (all_reduce_sum(stage, 1, 0, &e2) != 0) { *err = e2; cudaFree(stage); return -2; }
    cudaStreamSynchronize(0);
    cudaMemcpy(&bad, stage, sizeof bad, cudaMemcpyDeviceToHost);
    cudaFree(stage);
    if (bad != 0.0) {
        *err = ok ? "a peer rank could not set up its mailbox" : why;
        return -1;
    }
    p2p_ready = true;
    return 0;
}

void Comm::teardown_p2p()
{
    for (int r = 0; r < 8; ++r) {
        if (box_peer[r] && box_peer[r] != box_local_) cudaIpcCloseMemHandle(box_peer[r]);
        box_peer[r] = nullptr;
    }
    if (box_local_) cudaFree(box_local_);
    box_local_ = nullptr;
    p2p_ready = false;
}

// ---- host memory shared by the ranks (see comm.hpp) ------------------------------------------------------------------
void Comm::release_shared_host()
{
    if (shm_base_) {
        cudaHostUnregister(shm_base_);
        cudaGetLastError();
        munmap(shm_base_, shm_bytes_);
    }
    shm_base_ = nullptr;
    shm_bytes_ = 0;
}

double *Comm::shared_host(size_t doubles, std::string *err)
{
    constexpr size_t kHeader = 4096;
    const size_t want = kHeader + doubles * sizeof(double);
    if (!active() || shm_failed_) return nullptr;
    if (const char *e = std::getenv("NLOPT_B200_SHARED_HOST_X"))
        if (e[0] == '0') return nullptr;                 // every rank gathers x on the device and copies all of it down
    if (shm_base_ && shm_bytes_ >= want) return reinterpret_cast<double *>(static_cast<char *>(shm_base_) + kHeader);
    // (re)create: every rank takes every step, whatever its local outcome; the verdict is collective
    cudaDeviceSynchronize();
    release_shared_host();
    double *tmp = nullptr;
    if (cudaMalloc(&tmp, 2 * sizeof(double)) != cudaSuccess) { shm_failed_ = true; return nullptr; }
    auto exchange = [&](double v) -> double {          // sum over ranks of one double (also a barrier)
        cudaMemcpy(tmp, &v, sizeof v, cudaMemcpyHostToDevice);
        std::string e;
        all_reduce_sum(tmp, 1, 0, &e);
        cudaStreamSynchronize(0);
        cudaMemcpy(&v, tmp, sizeof v, cudaMemcpyDeviceToHost);
        return v;
    };
    const double pid0 = exchange(rank == 0 ? (double) getpid() : 0.0);
    char name[96];
    std::snprintf(name, sizeof name, "/nlopt_b200_%ld_%u", (long) pid0, ++shm_gen_);
    const size_t bytes = (want + ((size_t) 2 << 20) - 1) / ((size_t) 2 << 20) * ((size_t) 2 << 20);
    bool ok = true;
    int fd = -1;
    if (rank == 0) {
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t) bytes) != 0) ok = false;
    }
    const double created = exchange(rank == 0 && ok ? 1.0 : 0.0);          // barrier: the segment exists (or not)
    if (created != 1.0) ok = false;
    if (ok && rank != 0) {
        fd = shm_open(name, O_RDWR, 0600);
        if (fd < 0) ok = false;
    }
    void *base = nullptr;
    if (ok) {
        base = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (base == MAP_FAILED) { base = nullptr; ok = false; }
    }
    if (fd >= 0) close(fd);
    if (ok && rank == 0) std::memset(base, 0, kHeader);
    if (ok && cudaHostRegister(base, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); munmap(base, bytes); base = nullptr; ok = false; }
    const double bad = exchange(ok ? 0.0 : 1.0);                             // everyone has mapped (or someone failed)
    if (rank == 0) shm_unlink(name);                                          // the mapping outlives the name
    cudaFree(tmp);
    if (bad != 0.0) {
        if (base) { cudaHostUnregister(base); munmap(base, bytes); }
        shm_failed_ = true;
        if (err) *err = "shared host segment unavailable on some rank (ranks on different nodes?)";
        return nullptr;
    }
    shm_base_ = base;
    shm_bytes_ = bytes;
    shm_sense_ = 0;
    return reinterpret_cast<double *>(static_cast<char *>(shm_base_) + kHeader);
}

bool Comm::host_barrier()
{
    if (!shm_base_) return false;
    volatile unsigned *count = static_cast<volatile unsigned *>(shm_base_);
    volatile unsigned *sense = count + 16;                                   // its own cache line
    shm_sense_ ^= 1;
    const unsigned mine = (unsigned) shm_sense_;
    if (__atomic_add_fetch(const_cast<unsigned *>(count), 1u, __ATOMIC_ACQ_REL) == (unsigned) world) {
        __atomic_store_n(const_cast<unsigned *>(count), 0u, __ATOMIC_RELAXED);
        __atomic_store_n(const_cast<unsigned *>(sense), mine, __ATOMIC_RELEASE);
        return true;
    }
    const auto t0 = std::chrono::steady_clock::now();
    unsigned long spins = 0;
    while (__atomic_load_n(const_cast<unsigned *>(sense), __ATOMIC_ACQUIRE) != mine)
        if ((++spins & 0xfffff) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0) return false;
    return true;
}

int Comm::finalize()
{
    if (comm_) {
        cudaDeviceSynchronize();
        teardown_p2p();
        release_shared_host();
    }
    shm_failed_ = false;
    if (comm_) api().destroy(comm_);
    comm_ = nullptr;
    rank = 0; world = 1;
    return 0;
}

int Comm::all_gather_inplace(double *buf, size_t count_per_rank, cudaStream_t s, std::string *err)
{
    if (world == 1) return 0;
    int rc = api().all_gather(buf + (size_t) rank * count_per_rank, buf, count_per_rank, /*ncclFloat64*/ 8, comm_, s);
    if (rc) { if (err) *err = nccl_msg("ncclAllGather", rc); return -1; }
    return 0;
}

int Comm::all_reduce_sum(double *buf, size_t count, cudaStream_t s, std::string *err)
{
    if (world == 1) return 0;
    int rc = api().all_reduce(buf, buf, count, 8, /*ncclSum*/ 0, comm_, s);
    if (rc) { if (err) *err = nccl_msg("ncclAllReduce", rc); return -1; }
    return 0;
}

}  // namespace nb200

extern "C" {

int nlopt_b200_comm_unique_id(unsigned char id128[128])
{
    std::string e;
    return nb200::Comm::unique_id(id128, &e);
}

int nlopt_b200_comm_init(const unsigned char id128[128], int rank, int world, int device)
{
    std::string e;
    int rc = nb200::Comm::instance().init(id128, rank, world, device, &e);
    if (rc) std::fprintf(stderr, "nlopt_b200_comm_init: %s\n", e.c_str());
    return rc;
}

int nlopt_b200_comm_finalize(void) { return nb200::Comm::instance().finalize(); }
int nlopt_b200_comm_rank(void) { return nb200::Comm::instance().rank; }
int nlopt_b200_comm_world(void) { return nb200::Comm::instance().world; }

}  // extern "C"
