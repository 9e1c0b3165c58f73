// fwgpu_shard.cu -- sharded-table plumbing: CUDA virtual-memory API + unix-socket rendezvous (see fwgpu_shard.hpp).
#include "fwgpu_shard.hpp"

#include <chrono>
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

namespace fwgpu {

// ---- driver entry points, resolved at run time (the library links against the runtime only) ----
struct Drv {
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    bool ok = false;
    std::string err;
    Drv()
    {
        auto get = [&](const char *name, void **fn) {
            cudaDriverEntryPointQueryResult q;
            cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
            if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) { err = std::string("driver entry point ") + name + " unavailable"; return false; }
            return true;
        };
        ok = get("cuMemCreate", (void **)&MemCreate) && get("cuMemRelease", (void **)&MemRelease) &&
             get("cuMemAddressReserve", (void **)&MemAddressReserve) && get("cuMemAddressFree", (void **)&MemAddressFree) &&
             get("cuMemMap", (void **)&MemMap) && get("cuMemUnmap", (void **)&MemUnmap) && get("cuMemSetAccess", (void **)&MemSetAccess) &&
             get("cuMemExportToShareableHandle", (void **)&MemExportToShareableHandle) &&
             get("cuMemImportFromShareableHandle", (void **)&MemImportFromShareableHandle) &&
             get("cuMemGetAllocationGranularity", (void **)&MemGetAllocationGranularity) && get("cuGetErrorString", (void **)&GetErrorString);
    }
};
static Drv &drv()
{
    static Drv d;
    return d;
}
static std::string cu_err(const char *what, CUresult r)
{
    const char *s = nullptr;
    if (drv().GetErrorString) drv().GetErrorString(r, &s);
    return std::string(what) + ": " + (s ? s : "CUDA driver error ") + " (" + std::to_string((int)r) + ")";
}
static CUmemAllocationProp alloc_prop(int device)
{
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

// ---- unix-socket helpers ----
static bool send_all(int fd, const void *buf, size_t n)
{
    const char *p = (const char *)buf;
    while (n) { ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL); if (w <= 0) return false; p += w; n -= (size_t)w; }
    return true;
}
static bool recv_all(int fd, void *buf, size_t n, int timeout_ms)
{
    char *p = (char *)buf;
    while (n) {
        pollfd pf{fd, POLLIN, 0};
        if (::poll(&pf, 1, timeout_ms) <= 0) return false;
        ssize_t r = ::recv(fd, p, n, 0);
        if (r <= 0) return false;
        p += r; n -= (size_t)r;
    }
    return true;
}
static bool send_fd(int sock, int fd)
{
    char dummy = 'F', ctrl[CMSG_SPACE(sizeof(int))];
    memset(ctrl, 0, sizeof(ctrl));
    iovec io{&dummy, 1};
    msghdr msg{};
    msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
    cmsghdr *cm = CMSG_FIRSTHDR(&msg);
    cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(cm), &fd, sizeof(int));
    return ::sendmsg(sock, &msg, MSG_NOSIGNAL) == 1;
}
static int recv_fd(int sock, int timeout_ms)
{
    pollfd pf{sock, POLLIN, 0};
    if (::poll(&pf, 1, timeout_ms) <= 0) return -1;
    char dummy = 0, ctrl[CMSG_SPACE(sizeof(int))];
    iovec io{&dummy, 1};
    msghdr msg{};
    msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
    if (::recvmsg(sock, &msg, 0) != 1) return -1;
    cmsghdr *cm = CMSG_FIRSTHDR(&msg);
    if (!cm || cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS) return -1;
    int fd = -1;
    memcpy(&fd, CMSG_DATA(cm), sizeof(int));
    return fd;
}
static bool sock_addr(const std::string &path, sockaddr_un &a)
{
    memset(&a, 0, sizeof(a));
    a.sun_family = AF_UNIX;
    if (path.size() + 1 > sizeof(a.sun_path)) return false;
    memcpy(a.sun_path, path.c_str(), path.size() + 1);
    return true;
}
static int connect_retry(const std::string &path, uint32_t timeout_ms)
{
    sockaddr_un a;
    if (!sock_addr(path, a)) return -1;
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        int s = ::socket(AF_UNIX, SOCK_STREAM, 0);
        if (s < 0) return -1;
        if (::connect(s, (sockaddr *)&a, sizeof(a)) == 0) return s;
        ::close(s);
        if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > (long)timeout_ms) return -1;
        ::usleep(20000);
    }
}

enum { REQ_FD = 1, REQ_BARRIER = 2, REQ_BLOB = 3 };

// ---- NCCL, resolved at run time: libfwgpu.so loads (and every single-GPU entry point works) on machines without libnccl ----
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string err;
    NcclApi()
    {
        void *h = nullptr;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) if ((h = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        auto get = [&](const char *n, void **fn) { *fn = dlsym(h, n); if (!*fn) err = std::string("libnccl lacks ") + n; return *fn != nullptr; };
        ok = get("ncclGetUniqueId", (void **)&GetUniqueId) && get("ncclCommInitRank", (void **)&CommInitRank) && get("ncclAllGather", (void **)&AllGather) &&
             get("ncclCommDestroy", (void **)&CommDestroy) && get("ncclGetErrorString", (void **)&GetErrorString);
    }
};
static NcclApi &nccl()
{
    static NcclApi a;
    return a;
}

static void serve_connection(ShardGroup *g, int s)
{
    uint32_t req[2];
    if (recv_all(s, req, sizeof(req), (int)g->timeout_ms)) {
        std::unique_lock<std::mutex> lk(g->mu);
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(g->timeout_ms);
        if (req[0] == REQ_FD) {
            const bool ready = g->cv.wait_until(lk, deadline, [&] { return g->stop.load() || g->arrays.size() > req[1]; });
            // a destroyed array leaves a null slot (indices are the protocol); duplicate the fd under the lock so that a
            // concurrent destroy_array cannot close it between the look-up and the send
            const ShardedArray *arr = (ready && !g->stop.load()) ? g->arrays[req[1]] : nullptr;
            const int fd = (arr && arr->own_fd >= 0) ? ::dup(arr->own_fd) : -1;
            lk.unlock();
            if (fd >= 0) { send_fd(s, fd); ::close(fd); }
        } else if (req[0] == REQ_BLOB) {
            const bool ready = g->cv.wait_until(lk, deadline, [&] { return g->stop.load() || g->blobs.count(req[1]); });
            const std::string data = (ready && !g->stop.load()) ? g->blobs[req[1]] : std::string();
            lk.unlock();
            const uint32_t len = (uint32_t)data.size();
            if (send_all(s, &len, 4) && len) send_all(s, data.data(), len);
        } else if (req[0] == REQ_BARRIER) {
            const bool ready = g->cv.wait_until(lk, deadline, [&] { return g->stop.load() || g->phase >= req[1]; });
            lk.unlock();
            const char ok = (ready && !g->stop.load()) ? 1 : 0;
            send_all(s, &ok, 1);
        }
    }
    ::close(s);
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->active_workers--;
    }
    g->cv.notify_all();
}

bool ShardGroup::start(uint32_t rank_, uint32_t world_, int device_, const char *prefix_, uint32_t timeout_ms_)
{
    rank = rank_; world = world_; device = device_; prefix = prefix_ ? prefix_ : ""; if (timeout_ms_) timeout_ms = timeout_ms_;
    if (world == 0 || rank >= world || prefix.empty()) { error = "bad shard group arguments"; return false; }
    if (!drv().ok) { error = drv().err; return false; }
    CUmemAllocationProp prop = alloc_prop(device);
    CUresult r = drv().MemGetAllocationGranularity(&granularity, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS || granularity == 0) { error = cu_err("cuMemGetAllocationGranularity", r); return false; }
    const std::string path = prefix + "." + std::to_string(rank);
    sockaddr_un a;
    if (!sock_addr(path, a)) { error = "rendezvous path too long"; return false; }
    ::unlink(path.c_str());
    listen_fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
    // whoever can connect can obtain the exported GPU-memory handles: the socket is created owner-only
    const mode_t old_mask = ::umask(0177);
    const bool bound = listen_fd >= 0 && ::bind(listen_fd, (sockaddr *)&a, sizeof(a)) == 0;
    ::umask(old_mask);
    if (!bound || ::chmod(path.c_str(), 0600) != 0 || ::listen(listen_fd, 64) != 0) {
        error = "cannot listen on " + path + ": " + strerror(errno);
        return false;
    }
    server = std::thread([this] {
        // one short-lived detached thread per request (a barrier request blocks until this rank arrives); the destructor
        // waits for active_workers to drain, so nothing accumulates however many barriers a run performs
        while (!stop.load()) {
            pollfd pf{listen_fd, POLLIN, 0};
            if (::poll(&pf, 1, 100) <= 0) continue;
            int s = ::accept(listen_fd, nullptr, nullptr);
            if (s < 0) continue;
            {
                std::lock_guard<std::mutex> lk(mu);
                active_workers++;
            }
            std::thread(serve_connection, this, s).detach();
        }
    });
    return true;
}

ShardGroup::~ShardGroup()
{
    if (nccl_comm && nccl().ok) { nccl().CommDestroy((ncclComm_t)nccl_comm); nccl_comm = nullptr; }
    {
        std::lock_guard<std::mutex> lk(mu);
        stop.store(true);
    }
    cv.notify_all();
    if (server.joinable()) server.join();
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait_for(lk, std::chrono::milliseconds(timeout_ms + 1000), [&] { return active_workers == 0; });
    }
    if (listen_fd >= 0) { ::close(listen_fd); ::unlink((prefix + "." + std::to_string(rank)).c_str()); }
}

void shard_plan(size_t bytes, size_t tail_bytes, uint32_t world, size_t granularity, std::vector<size_t> &sizes)
{
    auto up = [&](size_t x) { return (x + granularity - 1) / granularity * granularity; };
    sizes.assign(world, 0);
    if (world > 1 && bytes % world == 0 && (bytes / world) % granularity == 0) {
        for (uint32_t s = 0; s < world; s++) sizes[s] = bytes / world;
        sizes[world - 1] += up(tail_bytes);
    } else {
        sizes[0] = up(bytes + tail_bytes);
    }
}

bool ShardGroup::create_array(ShardedArray &a, const std::vector<size_t> &sizes)
{
    Drv &d = drv();
    if (sizes.size() != world) { error = "shard plan does not match the group size"; return false; }
    a.sizes = sizes;
    a.offsets.assign(world, 0);
    a.total_bytes = 0;
    for (uint32_t s = 0; s < world; s++) {
        if (sizes[s] % granularity) { error = "shard sizes must be multiples of the allocation granularity"; return false; }
        a.offsets[s] = a.total_bytes;
        a.total_bytes += sizes[s];
    }
    a.handles.assign(world, 0);
    CUresult r;
    if (sizes[rank]) {
        CUmemAllocationProp prop = alloc_prop(device);
        r = d.MemCreate(&a.handles[rank], sizes[rank], &prop, 0);
        if (r != CUDA_SUCCESS) { error = cu_err("cuMemCreate", r); return false; }
        r = d.MemExportToShareableHandle(&a.own_fd, a.handles[rank], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
        if (r != CUDA_SUCCESS) { error = cu_err("cuMemExportToShareableHandle", r); return false; }
    }
    uint32_t index;
    {
        std::lock_guard<std::mutex> lk(mu);
        index = (uint32_t)arrays.size();
        arrays.push_back(&a);
    }
    cv.notify_all();
    r = d.MemAddressReserve(&a.va, a.total_bytes, granularity, 0, 0);
    if (r != CUDA_SUCCESS) { error = cu_err("cuMemAddressReserve", r); return false; }
    for (uint32_t s = 0; s < world; s++) {
        if (!sizes[s]) continue;
        if (s != rank) {
            int sock = connect_retry(prefix + "." + std::to_string(s), timeout_ms);
            if (sock < 0) { error = "cannot reach rank " + std::to_string(s) + " at " + prefix; return false; }
            const uint32_t req[2] = {REQ_FD, index};
            int fd = send_all(sock, req, sizeof(req)) ? recv_fd(sock, (int)timeout_ms) : -1;
            ::close(sock);
            if (fd < 0) { error = "rank " + std::to_string(s) + " did not send its shard handle"; return false; }
            r = d.MemImportFromShareableHandle(&a.handles[s], (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
            ::close(fd);
            if (r != CUDA_SUCCESS) { error = cu_err("cuMemImportFromShareableHandle", r); return false; }
        }
        r = d.MemMap(a.va + a.offsets[s], sizes[s], 0, a.handles[s], 0);
        if (r != CUDA_SUCCESS) { error = cu_err("cuMemMap", r); return false; }
    }
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = d.MemSetAccess(a.va, a.total_bytes, &acc, 1);
    if (r != CUDA_SUCCESS) { error = cu_err("cuMemSetAccess (is peer access between the GPUs available?)", r); return false; }
    return true;
}

void ShardGroup::destroy_array(ShardedArray &a)
{
    Drv &d = drv();
    {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &p : arrays) if (p == &a) p = nullptr; // the slot stays (indices are the protocol), the pointer goes
    }
    if (!d.ok) return;
    if (a.va) {
        d.MemUnmap(a.va, a.total_bytes);
        d.MemAddressFree(a.va, a.total_bytes);
        a.va = 0;
    }
    for (auto h : a.handles) if (h) d.MemRelease(h);
    a.handles.clear();
    if (a.own_fd >= 0) { ::close(a.own_fd); a.own_fd = -1; }
}

void ShardGroup::publish_blob(uint32_t id, const std::string &data)
{
    {
        std::lock_guard<std::mutex> lk(mu);
        blobs[id] = data;
    }
    cv.notify_all();
}

bool ShardGroup::fetch_blob(uint32_t from_rank, uint32_t id, std::string &out)
{
    int sock = connect_retry(prefix + "." + std::to_string(from_rank), timeout_ms);
    if (sock < 0) { error = "cannot reach rank " + std::to_string(from_rank); return false; }
    const uint32_t req[2] = {REQ_BLOB, id};
    uint32_t len = 0;
    bool ok = send_all(sock, req, sizeof(req)) && recv_all(sock, &len, 4, (int)timeout_ms) && len > 0 && len < (1u << 20);
    if (ok) { out.resize(len); ok = recv_all(sock, &out[0], len, (int)timeout_ms); }
    ::close(sock);
    if (!ok) error = "rank " + std::to_string(from_rank) + " did not publish blob " + std::to_string(id);
    return ok;
}

bool ShardGroup::comm_init()
{
    if (nccl_comm) return true;
    NcclApi &n = nccl();
    if (!n.ok) { error = n.err; return false; }
    ncclUniqueId id;
    if (rank == 0) {
        ncclResult_t r = n.GetUniqueId(&id);
        if (r != ncclSuccess) { error = std::string("ncclGetUniqueId: ") + n.GetErrorString(r); return false; }
        publish_blob(1, std::string((const char *)&id, sizeof(id)));
    } else {
        std::string blob;
        if (!fetch_blob(0, 1, blob) || blob.size() != sizeof(id)) { if (error.empty()) error = "bad NCCL id blob"; return false; }
        memcpy(&id, blob.data(), sizeof(id));
    }
    ncclComm_t comm = nullptr;
    ncclResult_t r = n.CommInitRank(&comm, (int)world, id, (int)rank);
    if (r != ncclSuccess) { error = std::string("ncclCommInitRank: ") + n.GetErrorString(r); return false; }
    nccl_comm = comm;
    return true;
}

bool ShardGroup::all_gather_u32(const uint32_t *send_dev, uint32_t *recv_dev, uint32_t count_per_rank, cudaStream_t stream)
{
    NcclApi &n = nccl();
    if (!nccl_comm) { error = "NCCL communicator not initialised"; return false; }
    ncclResult_t r = n.AllGather(send_dev, recv_dev, count_per_rank, ncclUint32, (ncclComm_t)nccl_comm, stream);
    if (r != ncclSuccess) { error = std::string("ncclAllGather: ") + n.GetErrorString(r); return false; }
    return true;
}

bool ShardGroup::barrier()
{
    uint64_t my;
    {
        std::lock_guard<std::mutex> lk(mu);
        my = ++phase;
    }
    cv.notify_all();
    for (uint32_t s = 0; s < world; s++) {
        if (s == rank) continue;
        int sock = connect_retry(prefix + "." + std::to_string(s), timeout_ms);
        if (sock < 0) { error = "barrier: cannot reach rank " + std::to_string(s); return false; }
        const uint32_t req[2] = {REQ_BARRIER, (uint32_t)my};
        char ok = 0;
        const bool got = send_all(sock, req, sizeof(req)) && recv_all(sock, &ok, 1, (int)timeout_ms) && ok == 1;
        ::close(sock);
        if (!got) { error = "barrier: rank " + std::to_string(s) + " did not arrive"; return false; }
    }
    return true;
}

} // namespace fwgpu
