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

    // collectives on a stream; return 0 on success
    int all_gather_inplace(double *buf, size_t count_per_rank, cudaStream_t s, std::string *err);
    int all_reduce_sum(double *buf, size_t count, cudaStream_t s, std::string *err);

private:
    void *handle_ = nullptr;   // dlopen handle
    void *comm_ = nullptr;     // ncclComm_t
};

}  // namespace nb200
