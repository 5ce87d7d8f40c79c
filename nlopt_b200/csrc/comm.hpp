// comm.hpp -- process-wide communicator for the sharded (one process per GPU) mode.
//
// The path shards by contiguous blocks of variables; its only exchange step is the m+3 partial
// sums of a dual evaluation (and 3 sums at the end of an outer iteration).  NCCL is loaded with
// dlopen at nlopt_b200_comm_init time (the library has no link-time NCCL dependency, so a
// single-GPU or CPU-only host can load it); the unique id travels through the host program's
// own bootstrap (torch.distributed in bench.py / tests).
#pragma once

#include <cuda_runtime.h>

#include <string>

namespace nb200 {

struct Comm {
    int rank = 0, world = 1, device = 0;
    bool active() const { return world > 1; }

    static Comm &instance();
    static int unique_id(unsigned char id[128], std::string *err);
    int init(const unsigned char id[128], int rank, int world, int device, std::string *err);
    int finalize();

    // Fused exchange: every rank owns a small mailbox in its HBM that all peers map through CUDA IPC
    // (NVLink peer stores).  The dual kernel's last warp writes this rank's shard sums straight into
    // every peer's mailbox, raises a per-rank flag and spins until all peers' flags for this sequence
    // number arrive -- the all-gather happens inside the kernel, no NCCL call, no second launch.
    // Layout: [2 buffers][8 virtual shards][kBoxStride] slots of 16 bytes {double value; u64 tag}.  Value and
    // tag (= the exchange's sequence number) travel in ONE 128-bit store, so a reader that sees the tag
    // also sees the value: no fences, no separate flag (the "LL128" idea of NCCL's low-latency protocol).
    static constexpr int kBoxStride = 24;
    static constexpr int kBoxDoubles = 2 * 8 * kBoxStride * 2;
    double *box_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool p2p_ready = false;
    bool use_p2p() const { return p2p_ready && !force_nccl_; }
    // Mailbox flags carry a launch sequence number.  The mailbox is shared by every solver object of the
    // process, so the number is process-wide; all ranks issue the same launches in the same order, which
    // keeps the counters of all ranks in lockstep.
    unsigned long long next_seq() { return ++seq_counter_; }
    void advance_seq(unsigned long long by) { seq_counter_ += by; }   // a dual-solve kernel used `by` numbers

    // Host memory shared by the ranks of a node (POSIX shm, page-locked in every process): with plain nlopt_func host
    // callbacks every rank needs the full x on the host; each rank copies only ITS shard down, into this segment, and a
    // host-side barrier makes the whole vector visible to all -- n/world instead of n doubles over each PCIe link.
    // Collective: every rank calls with the same size; returns nullptr on all ranks if any rank failed (the caller
    // falls back to gathering on the device).  host_barrier() returns false after 120 s (a peer died).
    double *shared_host(size_t doubles, std::string *err);
    bool host_barrier();

    // collectives on a stream; return 0 on success
    int all_gather_inplace(double *buf, size_t count_per_rank, cudaStream_t s, std::string *err);
    int all_reduce_sum(double *buf, size_t count, cudaStream_t s, std::string *err);

private:
    int setup_p2p(std::string *err);
    void teardown_p2p();
    double *box_local_ = nullptr;
    unsigned long long seq_counter_ = 0;
    bool force_nccl_ = false;
    void *shm_base_ = nullptr;          // mapping: 4 KB header (barrier words) + payload
    size_t shm_bytes_ = 0;
    unsigned shm_gen_ = 0;
    int shm_sense_ = 0;
    bool shm_failed_ = false;
    void release_shared_host();
    void *handle_ = nullptr;   // dlopen handle
    void *comm_ = nullptr;     // ncclComm_t
};

}  // namespace nb200
