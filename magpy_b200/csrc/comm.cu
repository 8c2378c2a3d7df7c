// comm.cu — the one collective of the multi-GPU path behind the C ABI: ncclAllReduce(sum, fp64) of the [S][4]
// ensemble sums over NVLink, one process per GPU.  Replaces the result gathering of the
// reference's joblib process pool (magpy/model.py:204-207, pickled Results sent back to the parent).
//
// NCCL is loaded with dlopen at the first communicator request: the library has no link-time dependency on it (a
// one-GPU user never needs it) and the process may already hold a copy (e.g. the one bundled with PyTorch) — whichever
// libnccl.so.2 is resident is the one used.  The unique id travels over a tiny TCP rendezvous that follows the
// launcher's contract (RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT as set by torchrun / mpirun wrappers): rank 0 serves
// the 128-byte id on MASTER_PORT + 1 (or MAGPY_B200_COMM_PORT), the other ranks fetch it.
#include <arpa/inet.h>
#include <dlfcn.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>

#include <nccl.h>

#include "../../include/magpy_b200.h"
#include "host_util.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mutex;

int load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.handle) return MAGPY_B200_OK;
    const char* names[] = {std::getenv("MAGPY_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (h) break;
    }
    if (!h) return mbh::fail(MAGPY_B200_ERR_COMM, "cannot load libnccl.so.2 (%s); set MAGPY_B200_NCCL_LIB", dlerror());
    NcclApi api;
    api.handle = h;
#define MB_SYM(field, name)                                                                   \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));                        \
    if (!api.field) return mbh::fail(MAGPY_B200_ERR_COMM, "libnccl lacks the symbol %s", name)
    MB_SYM(GetVersion, "ncclGetVersion");
    MB_SYM(GetUniqueId, "ncclGetUniqueId");
    MB_SYM(CommInitRank, "ncclCommInitRank");
    MB_SYM(CommDestroy, "ncclCommDestroy");
    MB_SYM(AllReduce, "ncclAllReduce");
    MB_SYM(GetErrorString, "ncclGetErrorString");
#undef MB_SYM
    g_nccl = api;
    return MAGPY_B200_OK;
}

#define NCCL_TRY(expr)                                                                                         \
    do {                                                                                                       \
        ncclResult_t r__ = (expr);                                                                             \
        if (r__ != ncclSuccess)                                                                                \
            return mbh::fail(MAGPY_B200_ERR_COMM, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r__), \
                             __FILE__, __LINE__);                                                              \
    } while (0)

#define CU_TRY(expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return mbh::fail(MAGPY_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                                        \
    } while (0)

bool send_all(int fd, const void* buf, size_t n) {
    const char* p = static_cast<const char*>(buf);
    while (n) {
        const ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
        if (k <= 0) return false;
        p += k;
        n -= (size_t)k;
    }
    return true;
}

bool recv_all(int fd, void* buf, size_t n) {
    char* p = static_cast<char*>(buf);
    while (n) {
        const ssize_t k = ::recv(fd, p, n, 0);
        if (k <= 0) return false;
        p += k;
        n -= (size_t)k;
    }
    return true;
}

// rank 0: hand `id` to the world - 1 peers that connect; other ranks: fetch it
int exchange_id(uint8_t id[MAGPY_B200_COMM_ID_BYTES], int rank, int world, const char* addr, int port, double timeout_s) {
    const uint32_t magic = 0xB2000C01u;
    if (rank == 0) {
        const int srv = ::socket(AF_INET, SOCK_STREAM, 0);
        if (srv < 0) return mbh::fail(MAGPY_B200_ERR_COMM, "socket(): %s", std::strerror(errno));
        int one = 1;
        ::setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        sockaddr_in sa{};
        sa.sin_family = AF_INET;
        sa.sin_addr.s_addr = htonl(INADDR_ANY);
        sa.sin_port = htons((uint16_t)port);
        if (::bind(srv, reinterpret_cast<sockaddr*>(&sa), sizeof sa) != 0 || ::listen(srv, world) != 0) {
            const int e = errno;
            ::close(srv);
            return mbh::fail(MAGPY_B200_ERR_COMM, "rank 0 cannot listen on port %d for the id rendezvous: %s "
                             "(set MAGPY_B200_COMM_PORT)", port, std::strerror(e));
        }
        timeval tv{(time_t)timeout_s, 0};
        ::setsockopt(srv, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
        for (int served = 0; served < world - 1;) {
            const int fd = ::accept(srv, nullptr, nullptr);
            if (fd < 0) {
                ::close(srv);
                return mbh::fail(MAGPY_B200_ERR_COMM, "id rendezvous: %d of %d peers connected within %.0f s", served,
                                 world - 1, timeout_s);
            }
            ::setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
            uint32_t hello[2] = {0, 0};
            if (recv_all(fd, hello, sizeof hello) && hello[0] == magic && send_all(fd, id, MAGPY_B200_COMM_ID_BYTES)) ++served;
            ::close(fd);
        }
        ::close(srv);
        return MAGPY_B200_OK;
    }
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    char port_s[16];
    std::snprintf(port_s, sizeof port_s, "%d", port);
    if (getaddrinfo(addr, port_s, &hints, &res) != 0 || !res)
        return mbh::fail(MAGPY_B200_ERR_COMM, "id rendezvous: cannot resolve MASTER_ADDR '%s'", addr);
    const auto t0 = std::chrono::steady_clock::now();
    int rc = MAGPY_B200_ERR_COMM;
    for (;;) {
        const int fd = ::socket(AF_INET, SOCK_STREAM, 0);
        if (fd >= 0 && ::connect(fd, res->ai_addr, res->ai_addrlen) == 0) {
            const uint32_t hello[2] = {magic, (uint32_t)rank};
            if (send_all(fd, hello, sizeof hello) && recv_all(fd, id, MAGPY_B200_COMM_ID_BYTES)) {
                ::close(fd);
                rc = MAGPY_B200_OK;
                break;
            }
        }
        if (fd >= 0) ::close(fd);
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) {
            mbh::fail(MAGPY_B200_ERR_COMM, "id rendezvous: rank %d could not reach rank 0 at %s:%d within %.0f s", rank, addr,
                      port, timeout_s);
            break;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
    freeaddrinfo(res);
    return rc;
}

}  // namespace

struct magpy_b200_comm {
    ncclComm_t nccl = nullptr;
    int rank = 0, world = 1, device = 0;
    cudaStream_t stream = nullptr;   // for the host-buffer helpers
    double* d_buf = nullptr;         // staging for host-buffer reductions
    size_t d_cap = 0;
};

namespace mbh {
// used by plan_run (magpy_b200.cu): in-place sum of a device buffer on the plan's stream
int comm_allreduce_device(magpy_b200_comm* c, double* dptr, size_t n, int op, cudaStream_t stream) {
    if (!c) return fail(MAGPY_B200_ERR_BAD_ARG, "comm is NULL");
    if (c->world == 1) return MAGPY_B200_OK;
    NCCL_TRY(g_nccl.AllReduce(dptr, dptr, n, ncclDouble, op == MAGPY_B200_COMM_MAX ? ncclMax : ncclSum, c->nccl, stream));
    return MAGPY_B200_OK;
}
int comm_device(const magpy_b200_comm* c) { return c->device; }
}  // namespace mbh

extern "C" {

int magpy_b200_comm_unique_id(uint8_t id[MAGPY_B200_COMM_ID_BYTES]) {
    if (!id) return mbh::fail(MAGPY_B200_ERR_BAD_ARG, "id is NULL");
    int rc = load_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == MAGPY_B200_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    NCCL_TRY(g_nccl.GetUniqueId(&u));
    std::memcpy(id, &u, sizeof u);
    return MAGPY_B200_OK;
}

int magpy_b200_comm_create(const uint8_t id[MAGPY_B200_COMM_ID_BYTES], int rank, int world_size, int device,
                           magpy_b200_comm** comm) {
    if (!comm) return mbh::fail(MAGPY_B200_ERR_BAD_ARG, "comm is NULL");
    *comm = nullptr;
    if (world_size < 1 || rank < 0 || rank >= world_size)
        return mbh::fail(MAGPY_B200_ERR_BAD_ARG, "rank %d outside world of size %d", rank, world_size);
    if (world_size > 1 && !id) return mbh::fail(MAGPY_B200_ERR_BAD_ARG, "id is NULL");
    int rc = mbh::select_device(device);
    if (rc) return rc;
    magpy_b200_comm* c = new (std::nothrow) magpy_b200_comm();
    if (!c) return mbh::fail(MAGPY_B200_ERR_NOMEM, "out of host memory");
    c->rank = rank; c->world = world_size; c->device = device;
    if (world_size > 1) {
        rc = load_nccl();
        if (rc) { delete c; return rc; }
        ncclUniqueId u;
        std::memcpy(&u, id, sizeof u);
        ncclResult_t r = g_nccl.CommInitRank(&c->nccl, world_size, u, rank);
        if (r != ncclSuccess) {
            delete c;
            return mbh::fail(MAGPY_B200_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        }
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        magpy_b200_comm_destroy(c);
        return mbh::fail(MAGPY_B200_ERR_CUDA, "cudaStreamCreate failed");
    }
    *comm = c;
    return MAGPY_B200_OK;
}

int magpy_b200_comm_create_from_env(int device, magpy_b200_comm** comm) {
    if (!comm) return mbh::fail(MAGPY_B200_ERR_BAD_ARG, "comm is NULL");
    *comm = nullptr;
    const char* e_rank = std::getenv("RANK");
    const char* e_world = std::getenv("WORLD_SIZE");
    if (!e_rank || !e_world)
        return mbh::fail(MAGPY_B200_ERR_COMM, "RANK / WORLD_SIZE are not set: launch one process per GPU (e.g. torchrun) or "
                         "pass an id to magpy_b200_comm_create");
    const int rank = std::atoi(e_rank), world = std::atoi(e_world);
    if (device < 0) {
        const char* e_local = std::getenv("LOCAL_RANK");
        device = e_local ? std::atoi(e_local) : rank;
    }
    uint8_t id[MAGPY_B200_COMM_ID_BYTES] = {0};
    if (world > 1) {
        const char* addr = std::getenv("MASTER_ADDR");
        const char* e_port = std::getenv("MAGPY_B200_COMM_PORT");
        int port = e_port ? std::atoi(e_port) : 0;
        if (port <= 0) {
            const char* mp = std::getenv("MASTER_PORT");
            if (!mp) return mbh::fail(MAGPY_B200_ERR_COMM, "MASTER_PORT (or MAGPY_B200_COMM_PORT) is not set");
            port = std::atoi(mp) + 1;
        }
        if (!addr || !*addr) addr = "127.0.0.1";
        int rc = mbh::select_device(device);
        if (rc) return rc;
        if (rank == 0) {
            rc = magpy_b200_comm_unique_id(id);
            if (rc) return rc;
        }
        rc = exchange_id(id, rank, world, addr, port, 120.0);
        if (rc) return rc;
    }
    return magpy_b200_comm_create(id, rank, world, device, comm);
}

int magpy_b200_comm_rank(const magpy_b200_comm* comm, int* rank, int* world_size) {
    if (!comm) return mbh::fail(MAGPY_B200_ERR_BAD_ARG, "comm is NULL");
    if (rank) *rank = comm->rank;
    if (world_size) *world_size = comm->world;
    return MAGPY_B200_OK;
}

int magpy_b200_comm_allreduce(magpy_b200_comm* c, double* host_values, size_t n, int op) {
    if (!c || (!host_values && n)) return mbh::fail(MAGPY_B200_ERR_BAD_ARG, "bad arguments");
    if (op != MAGPY_B200_COMM_SUM && op != MAGPY_B200_COMM_MAX) return mbh::fail(MAGPY_B200_ERR_BAD_ARG, "op must be SUM or MAX");
    if (c->world == 1 || n == 0) return MAGPY_B200_OK;
    CU_TRY(cudaSetDevice(c->device));
    if (c->d_cap < n) {
        if (c->d_buf) cudaFree(c->d_buf);
        c->d_buf = nullptr;
        c->d_cap = 0;
        CU_TRY(cudaMalloc(&c->d_buf, n * sizeof(double)));
        c->d_cap = n;
    }
    CU_TRY(cudaMemcpyAsync(c->d_buf, host_values, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    int rc = mbh::comm_allreduce_device(c, c->d_buf, n, op, c->stream);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(host_values, c->d_buf, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return MAGPY_B200_OK;
}

int magpy_b200_comm_barrier(magpy_b200_comm* c) {
    double token = 0.0;
    return magpy_b200_comm_allreduce(c, &token, 1, MAGPY_B200_COMM_SUM);
}

int magpy_b200_comm_destroy(magpy_b200_comm* c) {
    if (!c) return MAGPY_B200_OK;
    cudaSetDevice(c->device);
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    if (c->d_buf) cudaFree(c->d_buf);
    if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
    delete c;
    return MAGPY_B200_OK;
}

}  // extern "C"
