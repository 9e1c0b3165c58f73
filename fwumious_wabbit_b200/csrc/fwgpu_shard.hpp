// fwgpu_shard.hpp -- one weight table, physically sharded over the HBM of the GPUs of one NVSwitch box, mapped into ONE
// contiguous virtual address range on every rank (BASELINE config 4: "hashed weight table shards by high hash bits").
//
// The reference has a single table in host memory that every Hogwild thread updates (hogwild.rs:24-103).  Here every
// rank (one process per GPU) creates the physical memory of its own hash range with the CUDA virtual-memory API, exports
// it as a POSIX file descriptor, imports the other ranks' ranges and maps all of them back to back.  Index h of the
// table is then the same address expression on every GPU, `table + h`, whoever owns it: the fused learn kernels gather
// remote rows with ordinary 128-bit loads and scatter their AdaGrad updates with ordinary L2 atomics, which the hardware
// carries over NVLink to the owner's L2 -- the exchange of "gathered vectors and returned gradients" happens inside the
// kernel, row by row, overlapped with the arithmetic, and windows that cross a shard boundary need no halo.
//
// Rendezvous (no MPI, no torch): rank r listens on the unix socket "<prefix>.<r>"; peers connect to fetch a shard's file
// descriptor (SCM_RIGHTS) or to wait on a barrier phase.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace fwgpu {

struct ShardedArray {
    CUdeviceptr va = 0;
    size_t total_bytes = 0;            // reserved address range = sum of sizes
    std::vector<size_t> sizes;         // [world] bytes owned by each rank (0 = none), multiples of the granularity
    std::vector<size_t> offsets;       // [world] byte offset of each rank's range
    std::vector<CUmemGenericAllocationHandle> handles; // [world], own + imported
    int own_fd = -1;
};

// How `bytes` of table (+ `tail_bytes` that must follow the last element) are split over `world` ranks: equal hash ranges
// when every range is a whole number of allocation granules, otherwise the whole array lives on rank 0 (small tables).
// Pure function, the same on every rank.
void shard_plan(size_t bytes, size_t tail_bytes, uint32_t world, size_t granularity, std::vector<size_t> &sizes);

struct ShardGroup {
    uint32_t rank = 0, world = 1;
    int device = 0;
    std::string prefix;
    uint32_t timeout_ms = 60000;
    size_t granularity = 0;
    std::vector<ShardedArray *> arrays; // served by index
    // server
    int listen_fd = -1;
    std::thread server;
    std::atomic<bool> stop{false};
    std::mutex mu;
    std::condition_variable cv;
    uint64_t phase = 0;                // barrier phases this rank has reached
    int active_workers = 0;            // request threads in flight (guarded by mu)
    std::map<uint32_t, std::string> blobs; // small byte strings this rank publishes to its peers (the NCCL unique id)
    void *nccl_comm = nullptr;         // ncclComm_t over the group's ranks (libnccl is loaded at run time)
    std::string error;

    ~ShardGroup();
    bool start(uint32_t rank_, uint32_t world_, int device_, const char *prefix_, uint32_t timeout_ms_);
    bool create_array(ShardedArray &a, const std::vector<size_t> &sizes);
    bool barrier();
    void destroy_array(ShardedArray &a);
    // NCCL communicator over the ranks of the group (collective; the unique id travels over the rendezvous sockets).  Used for
    // the one exchange step of the sharded learn path: an all-gather of the per-owner push counts after every chunk, which is
    // also the barrier between "every rank has pushed its gradient rows" and "owners apply them".
    bool comm_init();
    bool all_gather_u32(const uint32_t *send_dev, uint32_t *recv_dev, uint32_t count_per_rank, cudaStream_t stream);
    void publish_blob(uint32_t id, const std::string &data);
    bool fetch_blob(uint32_t from_rank, uint32_t id, std::string &out);
    size_t round_up(size_t bytes) const { return (bytes + granularity - 1) / granularity * granularity; }
};

} // namespace fwgpu
