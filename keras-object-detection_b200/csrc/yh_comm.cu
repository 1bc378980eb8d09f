// yh_comm.cu - single-process multi-GPU exchange step of the mAP: a
// communicator over the devices of one box and the all-gather of per-detection records + all-reduce
// of the per-class ground-truth counts between yh_map_match (per shard) and yh_map_reduce.
//
// This is the C-ABI for hosts that drive all GPUs from ONE process (ncclCommInitAll).  The Python
// mirror uses one process per GPU with torch.distributed instead (yolohot/dist.py); both produce the
// same rank-ordered concatenation.  NCCL is resolved at run time (dlopen of the libnccl.so.2 that is
// already in the process, e.g. the one torch loaded), so libyolohot.so has no link-time dependency
// on it and two NCCL copies never meet in one process.
#include <dlfcn.h>

#include <mutex>
#include <vector>

#include <string.h>

#include "yh_common.cuh"

namespace yh {

// the slice of the NCCL API used here (ABI of nccl.h 2.x)
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
enum { kNcclUint8 = 1, kNcclInt32 = 2, kNcclUint64 = 5, kNcclSum = 0 };
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int load_nccl()
{
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.handle) return YH_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy already in the process, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("comm: cannot load libnccl.so.2 (%s)", dlerror());
        return YH_ERR_NCCL;
    }
#define YH_SYM(field, name)                                                       \
    do {                                                                          \
        *reinterpret_cast<void **>(&g_nccl.field) = dlsym(h, name);               \
        if (!g_nccl.field) {                                                      \
            set_error("comm: libnccl lacks %s", name);                            \
            return YH_ERR_NCCL;                                                   \
        }                                                                         \
    } while (0)
    YH_SYM(CommInitAll, "ncclCommInitAll");
    YH_SYM(CommDestroy, "ncclCommDestroy");
    YH_SYM(GroupStart, "ncclGroupStart");
    YH_SYM(GroupEnd, "ncclGroupEnd");
    YH_SYM(Broadcast, "ncclBroadcast");
    YH_SYM(AllReduce, "ncclAllReduce");
    YH_SYM(GetErrorString, "ncclGetErrorString");
#undef YH_SYM
    g_nccl.handle = h;
    return YH_OK;
}

#define YH_NCCL(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != 0) {                                                                         \
            set_error("NCCL error %d (%s) at %s", r_, g_nccl.GetErrorString(r_), #call);       \
            return YH_ERR_NCCL;                                                                \
        }                                                                                      \
    } while (0)

}  // namespace yh

using namespace yh;

extern "C" int yh_comm_init_all(int ndev, const int *devs, void **comm)
{
    YH_REQUIRE(ndev >= 1 && ndev <= kMaxPeers && comm != nullptr, "comm_init_all: bad arguments (at most %d devices)", kMaxPeers);
    int rc = load_nccl();
    if (rc != YH_OK) return rc;
    Comm *c = new Comm;
    c->ndev = ndev;
    for (int i = 0; i < ndev; ++i) c->devs[i] = devs ? devs[i] : i;
    ncclResult_t r = g_nccl.CommInitAll(reinterpret_cast<ncclComm_t *>(c->comms), ndev, c->devs);
    if (r != 0) {
        set_error("NCCL error %d (%s) at ncclCommInitAll", r, g_nccl.GetErrorString(r));
        delete c;
        return YH_ERR_NCCL;
    }
    // peer access for the fused match + scatter kernel (yh_map_match_p2p) and one event per device for yh_comm_barrier
    int prev = 0;
    cudaGetDevice(&prev);
    c->p2p = true;
    for (int i = 0; i < ndev; ++i) {
        cudaSetDevice(c->devs[i]);
        cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming);
        for (int j = 0; j < ndev; ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, c->devs[i], c->devs[j]);
            if (!can) { c->p2p = false; continue; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(c->devs[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) c->p2p = false;
            cudaGetLastError();
        }
    }
    cudaSetDevice(prev);
    *comm = c;
    return YH_OK;
}

extern "C" int yh_comm_destroy(void *comm)
{
    if (!comm) return YH_OK;
    Comm *c = static_cast<Comm *>(comm);
    for (int i = 0; i < c->ndev; ++i) {
        if (c->comms[i]) g_nccl.CommDestroy(static_cast<ncclComm_t>(c->comms[i]));
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    }
    delete c;
    return YH_OK;
}

extern "C" int yh_comm_p2p(void *comm) { return (comm && static_cast<Comm *>(comm)->p2p) ? 1 : 0; }

// Cross-device barrier on the devices' streams (no host wait): everything enqueued so far on every stream
// happens before anything enqueued afterwards on any of them.
extern "C" int yh_comm_barrier(void *comm, void *const *streams)
{
    YH_REQUIRE(comm != nullptr, "comm_barrier: null communicator");
    Comm *c = static_cast<Comm *>(comm);
    int prev = 0;
    YH_CUDA(cudaGetDevice(&prev));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
    for (int d = 0; d < c->ndev; ++d) {
        YH_CUDA(cudaSetDevice(c->devs[d]));
        YH_CUDA(cudaEventRecord(c->ev[d], streams ? static_cast<cudaStream_t>(streams[d]) : nullptr));
    }
    for (int d = 0; d < c->ndev; ++d) {
        YH_CUDA(cudaSetDevice(c->devs[d]));
        for (int e = 0; e < c->ndev; ++e)
            if (e != d) YH_CUDA(cudaStreamWaitEvent(streams ? static_cast<cudaStream_t>(streams[d]) : nullptr, c->ev[e], 0));
    }
    return YH_OK;
}

// Device d holds nrec[d] records (keys[d], tp[d]) of its image shard and gt[d] (C int32).  Afterwards
// every device holds, in out_keys[d] / out_tp[d], the records of all devices concatenated in device
// order (= image order, which the stable sort of yh_map_reduce needs for equal confidences) and
// gt[d] holds the sum over devices.  nrec is a HOST array (the caller knows its row counts); the
// variable-sized all-gather is one NCCL group of broadcasts, one per source device.
extern "C" int yh_map_allgather(void *comm, const uint64_t *const *keys, const uint8_t *const *tp, const int64_t *nrec,
                                int32_t *const *gt_per_class, int C, uint64_t *const *out_keys, uint8_t *const *out_tp,
                                int64_t out_capacity, void *const *streams)
{
    YH_REQUIRE(comm && nrec && gt_per_class && C >= 1, "map_allgather: bad arguments");
    Comm *c = static_cast<Comm *>(comm);
    int64_t total = 0;
    for (int d = 0; d < c->ndev; ++d) {
        YH_REQUIRE(nrec[d] >= 0, "map_allgather: negative record count");
        total += nrec[d];
    }
    YH_REQUIRE(total <= out_capacity, "map_allgather: %lld records do not fit out_capacity %lld",
               static_cast<long long>(total), static_cast<long long>(out_capacity));
    YH_REQUIRE(total == 0 || (keys && tp && out_keys && out_tp), "map_allgather: null pointer");
    int prev = 0;
    YH_CUDA(cudaGetDevice(&prev));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
    YH_NCCL(g_nccl.GroupStart());
    int64_t off = 0;
    for (int r = 0; r < c->ndev; ++r) {                                   // source device r
        if (nrec[r] > 0) {
            for (int d = 0; d < c->ndev; ++d) {
                cudaStream_t st = streams ? static_cast<cudaStream_t>(streams[d]) : nullptr;
                YH_NCCL(g_nccl.Broadcast(keys[r], out_keys[d] + off, static_cast<size_t>(nrec[r]), kNcclUint64, r, static_cast<ncclComm_t>(c->comms[d]), st));
                YH_NCCL(g_nccl.Broadcast(tp[r], out_tp[d] + off, static_cast<size_t>(nrec[r]), kNcclUint8, r, static_cast<ncclComm_t>(c->comms[d]), st));
            }
        }
        off += nrec[r];
    }
    for (int d = 0; d < c->ndev; ++d) {
        cudaStream_t st = streams ? static_cast<cudaStream_t>(streams[d]) : nullptr;
        YH_NCCL(g_nccl.AllReduce(gt_per_class[d], gt_per_class[d], static_cast<size_t>(C), kNcclInt32, kNcclSum, static_cast<ncclComm_t>(c->comms[d]), st));
    }
    YH_NCCL(g_nccl.GroupEnd());
    return YH_OK;
}

// Device scratch an operation allocates internally (stream-ordered pool / cached per device), in
// bytes, for a problem of n images (decode/NMS/loss) or n rows (mAP stages); 0 = none.  Informational:
// callers never pass workspaces, but can budget memory with it.
extern "C" size_t yh_workspace_bytes(int op, int64_t n, int S, int B, int C)
{
    (void)B;
    const int64_t M = static_cast<int64_t>(S) * S;
    switch (op) {
        case YH_OP_DECODE_NMS: return 0;                                   // shared memory only
        case YH_OP_DECODE_NMS_HOST: {                                      // 3 slots x (input + padded rows + counts [+ idx]) chunks
            const int64_t per = 4 * M * (C + 5 * B) + 24 * M + 4 + 4 * M;
            return static_cast<size_t>(3 * std::min<int64_t>(n * per, (96ll << 20) / (4 * M * (C + 5 * B)) * per + per));
        }
        case YH_OP_LOSS: return static_cast<size_t>(148 * 16) * 5 * sizeof(double) + sizeof(unsigned);
        case YH_OP_MAP_MATCH: return static_cast<size_t>(n) * 48 + (1u << 20);      // sort keys/values, double buffers, claims
        case YH_OP_MAP_REDUCE: return static_cast<size_t>(n) * 40 + (1u << 20);
        default: return 0;
    }
}

// ------------------------------------------------------------------------------------------
// One process per GPU: record buffers other processes of the box can map (CUDA IPC), so that a rank's match
// kernel stores straight into every peer's buffer (yh_map_match_peers) instead of all-gathering afterwards.
// ------------------------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == YH_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int yh_ipc_alloc(size_t bytes, void **ptr, unsigned char *handle)
{
    YH_REQUIRE(ptr && handle && bytes > 0, "ipc_alloc: bad arguments");
    *ptr = nullptr;
    YH_CUDA(cudaMalloc(ptr, bytes));               // a whole cudaMalloc allocation: the handle maps exactly this range
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
    if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
    if (e != cudaSuccess) {
        cudaFree(*ptr);
        *ptr = nullptr;
        return cuda_fail(e, "cudaIpcGetMemHandle");
    }
    memcpy(handle, &h, sizeof(h));
    return YH_OK;
}

extern "C" int yh_ipc_open(const unsigned char *handle, void **ptr)
{
    YH_REQUIRE(ptr && handle, "ipc_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    YH_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return YH_OK;
}

extern "C" int yh_ipc_close(void *ptr)
{
    if (ptr) YH_CUDA(cudaIpcCloseMemHandle(ptr));
    return YH_OK;
}

extern "C" int yh_ipc_free(void *ptr)
{
    if (ptr) YH_CUDA(cudaFree(ptr));
    return YH_OK;
}
