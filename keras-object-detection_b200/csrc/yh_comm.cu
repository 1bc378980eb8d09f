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

#include <algorithm>

#include "yh_common.cuh"
#include "yh_map_internal.cuh"

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
    // peer access for the kernel-level exchange (yh_map_exchange) and one event per device for yh_comm_barrier
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

// Device d holds nrec[d] packed records (rec[d]) of its image shard and gt[d] (C int32).  Afterwards every
// device holds, in out_rec[d], the records of all devices concatenated in device order (= image order, which the
// stable sort of yh_map_reduce needs for equal confidences) and gt[d] holds the sum over devices.  nrec is a HOST
// array; the variable-sized all-gather is one NCCL group of broadcasts, one per source device.  This is the
// collective form of the exchange step; yh_map_exchange below is the one fused with kernels over NVLink.
extern "C" int yh_map_allgather(void *comm, const uint64_t *const *rec, const int64_t *nrec, int32_t *const *gt_per_class, int C,
                                uint64_t *const *out_rec, int64_t out_capacity, void *const *streams)
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
    YH_REQUIRE(total == 0 || (rec && out_rec), "map_allgather: null pointer");
    int prev = 0;
    YH_CUDA(cudaGetDevice(&prev));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
    YH_NCCL(g_nccl.GroupStart());
    int64_t off = 0;
    for (int r = 0; r < c->ndev; ++r) {                                   // source device r
        if (nrec[r] > 0) {
            for (int d = 0; d < c->ndev; ++d) {
                cudaStream_t st = streams ? static_cast<cudaStream_t>(streams[d]) : nullptr;
                YH_NCCL(g_nccl.Broadcast(rec[r], out_rec[d] + off, static_cast<size_t>(nrec[r]), kNcclUint64, r, static_cast<ncclComm_t>(c->comms[d]), st));
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

// ------------------------------------------------------------------------------------------
// The exchange step as kernels over peer-mapped memory (NVLink / NVSwitch), no collective call and no host
// synchronisation.  Every rank owns one exchange buffer that all its peers can write (cudaDeviceEnablePeerAccess in
// one process, CUDA IPC between processes):
//     flags [2][16] u64 | count [2][16] i64 | ticket | gt [2][n][Cpad] i32 | rec [2][n][capacity] u64
// (2 = parity of the epoch: double buffered, so a fast rank may deliver epoch e+1 while a slow one still reduces
// epoch e; epoch e+2 cannot start anywhere before every rank delivered e+1, i.e. finished reducing e.)
// yh_map_exchange on rank r stores its records, its record count and its per-class ground-truth counts into slot r
// of EVERY rank's buffer; each thread fences its stores system-wide, and the last CTA to finish then releases
// flags[parity][r] = epoch on every rank.  The reduce kernel of a rank spins (bounded) on its own flags row, then
// reads the n segments in rank order - rank order = image order, which the stable sort relies on.
// ------------------------------------------------------------------------------------------
namespace yh {

struct ExLayout {
    size_t flags, count, ticket, gt, rec, total;
    int cpad;
};

static ExLayout ex_layout(int n_peers, int C, int64_t capacity)
{
    ExLayout l;
    l.cpad = (C + 63) / 64 * 64;
    l.flags = 0;
    l.count = l.flags + 2 * kMaxPeers * 8;
    l.ticket = l.count + 2 * kMaxPeers * 8;
    l.gt = l.ticket + 256;
    l.rec = align_up(l.gt + static_cast<size_t>(2) * n_peers * l.cpad * 4, 256);
    l.total = l.rec + static_cast<size_t>(2) * n_peers * static_cast<size_t>(capacity) * 8;
    return l;
}

struct ExArgs {
    unsigned char *buf[kMaxPeers];
    ExLayout l;
    int n, self, C, parity;
    long long capacity;
    const unsigned long long *rec;
    long long nrec_max;
    const long long *nrec_dev;
    const int32_t *gt;
    unsigned long long epoch;
};

__global__ void __launch_bounds__(256) map_exchange_kernel(ExArgs a)
{
    __shared__ int last_s;
    long long n = a.nrec_dev ? *a.nrec_dev : a.nrec_max;
    if (n > a.nrec_max) n = a.nrec_max;
    if (n < 0) n = 0;
    const bool fits = n <= a.capacity;
    const size_t slot = static_cast<size_t>(a.parity) * a.n + a.self;
    if (fits) {
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const unsigned long long v = a.rec[i];
            for (int d = 0; d < a.n; ++d)
                reinterpret_cast<unsigned long long *>(a.buf[d] + a.l.rec)[slot * a.capacity + i] = v;
        }
    }
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
            const int32_t v = a.gt[c];
            for (int d = 0; d < a.n; ++d) reinterpret_cast<int32_t *>(a.buf[d] + a.l.gt)[slot * a.l.cpad + c] = v;
        }
        if (threadIdx.x < a.n)      // -1: the shard does not fit its region (the reduce stage reports YH_MAP_ERR_OVERFLOW)
            reinterpret_cast<long long *>(a.buf[threadIdx.x] + a.l.count)[a.parity * kMaxPeers + a.self] = fits ? n : -1;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned *ticket = reinterpret_cast<unsigned *>(a.buf[a.self] + a.l.ticket);
        last_s = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        if (last_s) *ticket = 0;
    }
    __syncthreads();
    if (last_s && threadIdx.x < a.n) {
        __threadfence_system();
        unsigned long long *f = reinterpret_cast<unsigned long long *>(a.buf[threadIdx.x] + a.l.flags) + a.parity * kMaxPeers + a.self;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(a.epoch) : "memory");
    }
}

}  // namespace yh

extern "C" size_t yh_map_exchange_bytes(int n_peers, int C, int64_t capacity)
{
    if (n_peers < 1 || n_peers > kMaxPeers || C < 1 || capacity < 0) return 0;
    return ex_layout(n_peers, C, capacity).total;
}

extern "C" int yh_map_exchange(int n_peers, int self, void *const *bufs, int C, int64_t capacity, const uint64_t *rec,
                               int64_t nrec_max, const int64_t *nrec_dev, const int32_t *gt_per_class, uint64_t epoch, void *stream)
{
    YH_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers && self >= 0 && self < n_peers,
               "map_exchange: bad peer count / index (at most %d peers)", kMaxPeers);
    YH_REQUIRE(C >= 1 && C <= kMaxMapClasses && capacity >= 0 && nrec_max >= 0 && epoch >= 1, "map_exchange: bad sizes / epoch");
    YH_REQUIRE(bufs && gt_per_class && (nrec_max == 0 || rec), "map_exchange: null pointer");
    ExArgs a{};
    for (int d = 0; d < n_peers; ++d) {
        YH_REQUIRE(bufs[d] != nullptr, "map_exchange: null buffer for peer %d", d);
        a.buf[d] = static_cast<unsigned char *>(bufs[d]);
    }
    a.l = ex_layout(n_peers, C, capacity);
    a.n = n_peers; a.self = self; a.C = C; a.parity = static_cast<int>(epoch & 1);
    a.capacity = capacity;
    a.rec = reinterpret_cast<const unsigned long long *>(rec);
    a.nrec_max = nrec_max;
    a.nrec_dev = reinterpret_cast<const long long *>(nrec_dev);
    a.gt = gt_per_class;
    a.epoch = epoch;
    const long long want = (nrec_max + 255) / 256;
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(want, 2 * sm_count())));
    map_exchange_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    YH_LAUNCH_CHECK("map_exchange_kernel");
    return YH_OK;
}

// reduce stage over the records delivered for `epoch` into this rank's own exchange buffer
extern "C" int yh_map_reduce_exchanged(int n_peers, void *own_buf, int C, int64_t capacity, uint64_t epoch, int64_t n_hint,
                                       float *out_ap, float *out_map, int32_t *err, void *workspace, size_t workspace_bytes,
                                       void *stream)
{
    YH_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers && own_buf && C >= 1 && capacity >= 0 && epoch >= 1,
               "map_reduce_exchanged: bad arguments");
    const ExLayout l = ex_layout(n_peers, C, capacity);
    unsigned char *b = static_cast<unsigned char *>(own_buf);
    const int parity = static_cast<int>(epoch & 1);
    const uint64_t *rec[kMaxPeers];
    const int64_t *cnt[kMaxPeers];
    const int32_t *gt[kMaxPeers];
    int64_t cmax[kMaxPeers];
    for (int s = 0; s < n_peers; ++s) {
        const size_t slot = static_cast<size_t>(parity) * n_peers + s;
        rec[s] = reinterpret_cast<const uint64_t *>(b + l.rec) + slot * capacity;
        cnt[s] = reinterpret_cast<const int64_t *>(b + l.count) + parity * kMaxPeers + s;
        gt[s] = reinterpret_cast<const int32_t *>(b + l.gt) + slot * l.cpad;
        cmax[s] = capacity;
    }
    const uint64_t *flags = reinterpret_cast<const uint64_t *>(b + l.flags) + parity * kMaxPeers;
    return reduce_impl(n_peers, rec, cmax, cnt, n_peers, gt, C, out_ap, out_map, flags, n_peers, epoch, err, n_hint, workspace,
                       workspace_bytes, stream);
}

// Device scratch of an operation, in bytes, for a problem of n images (decode/NMS/loss) or n rows (mAP stages);
// 0 = none.  The mAP stages take it as a caller workspace (or allocate it from the stream-ordered pool when the
// caller passes NULL); the others allocate internally and the figure is informational.
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
        case YH_OP_MAP_MATCH: return match_ws_bytes(n, n, false);                  // n ground-truth rows and n detections, any order
        case YH_OP_MAP_REDUCE: return radix_ws_bytes(n, C);                        // two record buffers + tables
        default: return 0;
    }
}

// ------------------------------------------------------------------------------------------
// One process per GPU: record buffers other processes of the box can map (CUDA IPC), so that a rank's match
// exchange kernel stores straight into every peer's buffer (yh_map_exchange) instead of all-gathering afterwards.
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
