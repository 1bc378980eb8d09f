// yh_map.cu - K6/K7: IoU matching behind mAP and the per-class AP reduction, plus the
// evaluator's row compaction.  sm_100a.
//
// Replaces utils.py:303-456 (mean_average_precision, change_tensor :280-299) and the append
// of MeanAveragePrecision.update_state utils.py:476-489.
//
// The reference walks, per class, the detections in stable descending-confidence order and
// lets each one claim the ground truth of its image with the highest IoU (strict >, first
// wins) if that IoU is > thr and the GT is still free (utils.py:373-422).  Which GT a
// detection points at does not depend on the claims, so:
//   K6a  ground truths are grouped by (class, image) with a stable radix sort (row order kept
//        inside a group = the reference's `ground_truth_img` order, utils.py:378)
//   K6b  detections are put in (class asc, conf desc, row asc) order - the reference's
//        processing order, utils.py:367 - by one stable radix sort
//   K6c  one thread per detection: binary search of its (class, image) GT group, best IoU;
//        hits do atomicMin(claim[gt], sorted position): the earliest detection owns the GT
//   K6d  TP <=> the detection owns the GT it points at
//   K7   (after shards are concatenated) stable sort by (class, ~conf), inclusive scan of
//        TP, float32 precision/recall points exactly as utils.py:430-439, trapezoid terms as
//        np.trapz on float32, summed per class in float64 by one CTA in a fixed order.
// Radix sort / scan are CUB device primitives (library calls, like cuBLAS would be).
#include <cub/cub.cuh>

#include <algorithm>

#include "yh_common.cuh"

namespace yh {

__device__ __forceinline__ uint32_t orderable(float f)
{
    f = __fadd_rn(f, 0.0f);                       // -0 -> +0 so that equal floats get equal keys
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// class of a row as an integer in [0, C), or C when the reference would never select the row
// (utils.py:329-330 compare the float class with float(c), c = 0..C-1)
__device__ __forceinline__ uint32_t class_of(float cf, int C)
{
    const int ci = static_cast<int>(cf);
    return (cf >= 0.0f && cf < static_cast<float>(C) && static_cast<float>(ci) == cf) ? static_cast<uint32_t>(ci)
                                                                                    : static_cast<uint32_t>(C);
}

__global__ void gt_keys_kernel(const float *__restrict__ rows, int64_t n, int C, uint64_t *__restrict__ keys,
                               uint32_t *__restrict__ vals, int *__restrict__ gt_per_class)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = class_of(rows[7 * i + 1], C);
    keys[i] = (static_cast<uint64_t>(c) << 32) | orderable(rows[7 * i]);
    vals[i] = static_cast<uint32_t>(i);
    if (c < static_cast<uint32_t>(C)) atomicAdd(gt_per_class + c, 1);
}

__global__ void det_keys_kernel(const float *__restrict__ rows, int64_t n, int C, uint64_t *__restrict__ keys,
                                uint32_t *__restrict__ vals)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = class_of(rows[7 * i + 1], C);
    const uint32_t k = (c < static_cast<uint32_t>(C)) ? ~orderable(rows[7 * i + 2]) : 0xffffffffu;
    keys[i] = (static_cast<uint64_t>(c) << 32) | k;
    vals[i] = static_cast<uint32_t>(i);
}

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t *a, int64_t n, uint64_t key)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// K6c: s = sorted detection position
__global__ void match_kernel(const float *__restrict__ pred_rows, const uint32_t *__restrict__ det_vals,
                             const uint64_t *__restrict__ det_keys, int64_t np, const float *__restrict__ true_rows,
                             const uint64_t *__restrict__ gt_keys, const uint32_t *__restrict__ gt_vals, int64_t nt, int C,
                             float iou_thr, int32_t *__restrict__ hit, uint32_t *__restrict__ claim)
{
    const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= np) return;
    int32_t h = -1;
    const uint32_t c = static_cast<uint32_t>(det_keys[s] >> 32);
    if (c < static_cast<uint32_t>(C)) {
        const float *d = pred_rows + 7ll * det_vals[s];
        const uint64_t gk = (static_cast<uint64_t>(c) << 32) | orderable(d[0]);
        const int64_t lo = lower_bound_u64(gt_keys, nt, gk);
        const float dx = d[3], dy = d[4], dw = d[5], dh = d[6];
        float best = 0.0f;                               // utils.py:382 (unwritten slot reads 0)
        int64_t bj = lo;                                 // utils.py:383 (index defaults to 0)
        int64_t g = lo;
        for (; g < nt && gt_keys[g] == gk; ++g) {        // utils.py:386
            const float *t = true_rows + 7ll * gt_vals[g];
            const float v = iou_ref(dx, dy, dw, dh, t[3], t[4], t[5], t[6]);   // utils.py:387 (det, gt)
            if (v > best) { best = v; bj = g; }          // utils.py:389
        }
        if (g > lo && best > iou_thr) {                  // utils.py:395
            h = static_cast<int32_t>(bj);
            atomicMin(claim + bj, static_cast<uint32_t>(s));
        }
    }
    hit[s] = h;
}

// K6d
__global__ void tp_kernel(const int32_t *__restrict__ hit, const uint32_t *__restrict__ claim, int64_t np,
                          uint8_t *__restrict__ tp)
{
    const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= np) return;
    const int32_t h = hit[s];
    tp[s] = (h >= 0 && claim[h] == static_cast<uint32_t>(s)) ? 1 : 0;   // utils.py:408-418
}

// K6d', fused with the exchange step: the TP decision of a detection and its sort key are stored straight into the
// record buffers of EVERY device of the box (peer-mapped memory over NVLink / NVSwitch) at this shard's offset, so
// no separate all-gather runs afterwards.  The sorted keys already sit in the local buffer (the radix sort wrote them).
struct PeerSet {
    int n, self;
    uint64_t *keys[kMaxPeers];
    uint8_t *tp[kMaxPeers];
    int32_t *gt[kMaxPeers];
    int64_t offset;
    int has_gt;
};

__global__ void tp_scatter_kernel(const int32_t *__restrict__ hit, const uint32_t *__restrict__ claim, int64_t np, PeerSet ps)
{
    const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= np) return;
    const int32_t h = hit[s];
    const uint8_t tp = (h >= 0 && claim[h] == static_cast<uint32_t>(s)) ? 1 : 0;   // utils.py:408-418
    const uint64_t key = ps.keys[ps.self][ps.offset + s];
    for (int d = 0; d < ps.n; ++d) {
        ps.tp[d][ps.offset + s] = tp;
        if (d != ps.self) ps.keys[d][ps.offset + s] = key;
    }
}

// per-class ground-truth counts of this shard added into every device's accumulator (system-scope atomics)
__global__ void gt_scatter_kernel(const int32_t *__restrict__ gt_local, int C, PeerSet ps)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int v = gt_local[c];
    if (v == 0) return;
    for (int d = 0; d < ps.n; ++d) atomicAdd_system(ps.gt[d] + c, v);
}

// K7: class segment starts in the sorted record keys: start[c] = lower_bound(c << 32), c = 0..C
__global__ void class_starts_kernel(const uint64_t *__restrict__ keys, int64_t n, int C, int64_t *__restrict__ start)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > C) return;
    start[c] = lower_bound_u64(keys, n, static_cast<uint64_t>(c) << 32);
}

struct U8ToI32 {
    __host__ __device__ int operator()(uint8_t v) const { return static_cast<int>(v); }
};

// precision / recall point of sorted record i of a class starting at s0 (utils.py:430-435)
__device__ __forceinline__ void pr_point(const int *__restrict__ cum, int64_t i, int64_t s0, int base, float total,
                                         float &rec, float &prec)
{
    const int tpi = cum[i] - base;
    const float tpc = static_cast<float>(tpi);
    const float fpc = static_cast<float>(static_cast<int>(i - s0 + 1) - tpi);
    rec = __fdiv_rn(tpc, __fadd_rn(total, 1e-6f));
    prec = __fdiv_rn(tpc, __fadd_rn(__fadd_rn(tpc, fpc), 1e-6f));
}

// one CTA per class: AP = sum of np.trapz terms (float32 each), accumulated in float64
__global__ void __launch_bounds__(256) ap_kernel(const int *__restrict__ cum, const int64_t *__restrict__ start,
                                                 const int *__restrict__ gt_per_class, float *__restrict__ ap_out)
{
    const int c = blockIdx.x;
    __shared__ double red[256];
    const int64_t s0 = start[c], s1 = start[c + 1];
    const int ngt = gt_per_class[c];
    double acc = 0.0;
    if (ngt > 0) {                                                        // utils.py:334-336
        const float total = static_cast<float>(ngt);
        const int base = (s0 > 0) ? cum[s0 - 1] : 0;
        for (int64_t i = s0 + threadIdx.x; i < s1; i += blockDim.x) {
            float r1, p1, r0 = 0.0f, p0 = 1.0f;                           // utils.py:438-439
            pr_point(cum, i, s0, base, total, r1, p1);
            if (i > s0) pr_point(cum, i - 1, s0, base, total, r0, p0);
            const float term = __fmul_rn(__fmul_rn(__fsub_rn(r1, r0), __fadd_rn(p1, p0)), 0.5f);   // np.trapz
            acc += static_cast<double>(term);
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (static_cast<int>(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) ap_out[c] = static_cast<float>(red[0]);
}

__global__ void map_mean_kernel(const float *__restrict__ ap, int C, float *__restrict__ out_ap, float *__restrict__ out_map)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int c = 0; c < C; ++c) {
            s += static_cast<double>(ap[c]);
            if (out_ap) out_ap[c] = ap[c];
        }
        *out_map = static_cast<float>(s / static_cast<double>(C));         // utils.py:456
    }
}

// ---- evaluator rows ----------------------------------------------------------------------
__global__ void rows_append_kernel(const float *__restrict__ boxes, const int *__restrict__ count,
                                   const int64_t *__restrict__ offs, int64_t n, int M, int64_t img_base,
                                   float *__restrict__ out_rows, int64_t capacity, const int64_t *__restrict__ cursor)
{
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= n * M) return;
    const int64_t img = idx / M;
    const int slot = static_cast<int>(idx % M);
    if (slot >= count[img]) return;
    const int64_t row = *cursor + offs[img] + slot;
    if (row >= capacity) return;
    const float *b = boxes + idx * 6;
    float *o = out_rows + row * 7;
    o[0] = static_cast<float>(img_base + img);                            // utils.py:476
    o[1] = b[0]; o[2] = b[1]; o[3] = b[2]; o[4] = b[3]; o[5] = b[4]; o[6] = b[5];
}

__global__ void cursor_advance_kernel(const int *__restrict__ count, const int64_t *__restrict__ offs, int64_t n,
                                      int64_t *__restrict__ cursor)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *cursor += offs[n - 1] + count[n - 1];
}

struct I32ToI64 {
    __host__ __device__ int64_t operator()(int v) const { return static_cast<int64_t>(v); }
};

static inline int blocks_for(int64_t n, int t) { return static_cast<int>((n + t - 1) / t); }

struct AsyncBuf {   // stream-ordered scratch, freed on scope exit
    cudaStream_t st;
    void *p = nullptr;
    explicit AsyncBuf(cudaStream_t s) : st(s) {}
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 16, st); }
    ~AsyncBuf() { if (p) cudaFreeAsync(p, st); }
    template <class T> T *as() { return static_cast<T *>(p); }
};

// The stream-ordered allocator trims its pool at every synchronisation unless a release threshold
// is set; the evaluator synchronises often (result() returns a host scalar), so keep the scratch.
static int keep_pool()
{
    static bool done[64] = {false};
    int dev = 0;
    YH_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !done[dev]) {
        cudaMemPool_t pool;
        YH_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t thr = 1ull << 30;      // retain up to 1 GiB of freed scratch
        YH_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        done[dev] = true;
    }
    return YH_OK;
}

static int bits_for(int C)
{
    int b = 1;
    while ((1ll << b) <= C) ++b;
    return 32 + b;
}

}  // namespace yh

using namespace yh;

static int map_match_impl(const float *true_rows, int64_t nt, const float *pred_rows, int64_t np, int C, float iou_thr,
                          uint64_t *out_keys, uint8_t *out_tp, int32_t *out_gt_per_class, void *stream, const PeerSet *peers)
{
    YH_REQUIRE(C >= 1 && nt >= 0 && np >= 0, "map_match: bad sizes");
    YH_REQUIRE(nt < (1ll << 31) && np < (1ll << 31), "map_match: more than 2^31 rows");
    YH_REQUIRE(out_gt_per_class != nullptr, "map_match: out_gt_per_class is null");
    YH_REQUIRE((nt == 0 || true_rows) && (np == 0 || (pred_rows && out_keys && out_tp)), "map_match: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    { int rc_ = keep_pool(); if (rc_ != YH_OK) return rc_; }
    YH_CUDA(cudaMemsetAsync(out_gt_per_class, 0, sizeof(int32_t) * C, st));
    const int end_bit = bits_for(C);

    AsyncBuf gk_in(st), gv_in(st), gk(st), gv(st), dk_in(st), dv_in(st), dv(st), claim(st), hit(st), tmp(st);
    YH_CUDA(gk_in.alloc(8 * nt)); YH_CUDA(gv_in.alloc(4 * nt)); YH_CUDA(gk.alloc(8 * nt)); YH_CUDA(gv.alloc(4 * nt));
    YH_CUDA(dk_in.alloc(8 * np)); YH_CUDA(dv_in.alloc(4 * np)); YH_CUDA(dv.alloc(4 * np));
    YH_CUDA(claim.alloc(4 * nt)); YH_CUDA(hit.alloc(4 * np));

    if (nt > 0) {
        gt_keys_kernel<<<blocks_for(nt, 256), 256, 0, st>>>(true_rows, nt, C, gk_in.as<uint64_t>(), gv_in.as<uint32_t>(),
                                                            out_gt_per_class);
        YH_LAUNCH_CHECK("gt_keys_kernel");
    }
    if (peers && peers->has_gt) {
        gt_scatter_kernel<<<blocks_for(C, 128), 128, 0, st>>>(out_gt_per_class, C, *peers);
        YH_LAUNCH_CHECK("gt_scatter_kernel");
    }
    if (np == 0) return YH_OK;
    det_keys_kernel<<<blocks_for(np, 256), 256, 0, st>>>(pred_rows, np, C, dk_in.as<uint64_t>(), dv_in.as<uint32_t>());
    YH_LAUNCH_CHECK("det_keys_kernel");

    size_t tb1 = 0, tb2 = 0;
    YH_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb1, gk_in.as<uint64_t>(), gk.as<uint64_t>(), gv_in.as<uint32_t>(),
                                            gv.as<uint32_t>(), static_cast<int>(nt), 0, end_bit, st));
    YH_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, dk_in.as<uint64_t>(), out_keys, dv_in.as<uint32_t>(),
                                            dv.as<uint32_t>(), static_cast<int>(np), 0, end_bit, st));
    YH_CUDA(tmp.alloc(std::max(tb1, tb2)));
    size_t tb = std::max(tb1, tb2);
    if (nt > 0) {
        YH_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, gk_in.as<uint64_t>(), gk.as<uint64_t>(), gv_in.as<uint32_t>(),
                                                gv.as<uint32_t>(), static_cast<int>(nt), 0, end_bit, st));
        count_launch(4);
        YH_CUDA(cudaMemsetAsync(claim.p, 0xff, 4 * nt, st));
    }
    tb = std::max(tb1, tb2);
    YH_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk_in.as<uint64_t>(), out_keys, dv_in.as<uint32_t>(),
                                            dv.as<uint32_t>(), static_cast<int>(np), 0, end_bit, st));
    count_launch(4);
    match_kernel<<<blocks_for(np, 128), 128, 0, st>>>(pred_rows, dv.as<uint32_t>(), out_keys, np, true_rows,
                                                      gk.as<uint64_t>(), gv.as<uint32_t>(), nt, C, iou_thr,
                                                      hit.as<int32_t>(), claim.as<uint32_t>());
    YH_LAUNCH_CHECK("match_kernel");
    if (peers) {
        tp_scatter_kernel<<<blocks_for(np, 256), 256, 0, st>>>(hit.as<int32_t>(), claim.as<uint32_t>(), np, *peers);
        YH_LAUNCH_CHECK("tp_scatter_kernel");
    } else {
        tp_kernel<<<blocks_for(np, 256), 256, 0, st>>>(hit.as<int32_t>(), claim.as<uint32_t>(), np, out_tp);
        YH_LAUNCH_CHECK("tp_kernel");
    }
    return YH_OK;
}

extern "C" int yh_map_match(const float *true_rows, int64_t nt, const float *pred_rows, int64_t np, int C, float iou_thr,
                            uint64_t *out_keys, uint8_t *out_tp, int32_t *out_gt_per_class, void *stream)
{
    return map_match_impl(true_rows, nt, pred_rows, np, C, iou_thr, out_keys, out_tp, out_gt_per_class, stream, nullptr);
}

// Stage 1 fused with the exchange step over peer-mapped memory: this device (index `self` of `n_peers`) matches its
// shard and stores its records at [offset, offset + np) of EVERY peer's record buffers.  The pointers may come from
// cudaDeviceEnablePeerAccess in one process (yh_map_match_p2p) or from CUDA IPC handles of other processes (yh_ipc_open).
// gt_sum_all (nullable): per-device accumulators that receive this shard's per-class GT counts by system-scope atomics;
// out_gt_local (nullable): the shard's own counts (for a caller that sums them itself).
extern "C" int yh_map_match_peers(int n_peers, int self, const float *true_rows, int64_t nt, const float *pred_rows, int64_t np,
                                  int C, float iou_thr, uint64_t *const *out_keys_all, uint8_t *const *out_tp_all, int64_t offset,
                                  int32_t *const *gt_sum_all, int32_t *out_gt_local, void *stream)
{
    YH_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers && self >= 0 && self < n_peers && offset >= 0,
               "map_match_peers: bad peer count / index / offset (at most %d peers)", kMaxPeers);
    YH_REQUIRE(C >= 1 && out_keys_all && out_tp_all, "map_match_peers: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PeerSet ps;
    ps.n = n_peers; ps.self = self; ps.offset = offset;
    for (int d = 0; d < n_peers; ++d) {
        ps.keys[d] = out_keys_all[d]; ps.tp[d] = out_tp_all[d]; ps.gt[d] = gt_sum_all ? gt_sum_all[d] : nullptr;
        YH_REQUIRE((!gt_sum_all || ps.gt[d]) && (np == 0 || (ps.keys[d] && ps.tp[d])), "map_match_peers: null buffer for peer %d", d);
    }
    ps.has_gt = gt_sum_all ? 1 : 0;
    AsyncBuf gt_tmp(st);
    int32_t *gt_local = out_gt_local;
    if (!gt_local) {
        YH_CUDA(gt_tmp.alloc(sizeof(int32_t) * C));
        gt_local = gt_tmp.as<int32_t>();
    }
    return map_match_impl(true_rows, nt, pred_rows, np, C, iou_thr, ps.keys[self] + offset, ps.tp[self] + offset, gt_local,
                          stream, &ps);
}

// Single-process flavour: the peers are the devices of a communicator (peer access enabled by yh_comm_init_all).
extern "C" int yh_map_match_p2p(void *comm, int dev_index, const float *true_rows, int64_t nt, const float *pred_rows, int64_t np,
                                int C, float iou_thr, uint64_t *const *out_keys_all, uint8_t *const *out_tp_all, int64_t offset,
                                int32_t *const *gt_sum_all, void *stream)
{
    YH_REQUIRE(comm != nullptr, "map_match_p2p: null communicator");
    const Comm *c = static_cast<const Comm *>(comm);
    YH_REQUIRE(c->p2p, "map_match_p2p: the devices of this communicator cannot access each other's memory; use yh_map_allgather");
    YH_REQUIRE(dev_index >= 0 && dev_index < c->ndev && gt_sum_all, "map_match_p2p: bad device index / null pointer");
    return yh_map_match_peers(c->ndev, dev_index, true_rows, nt, pred_rows, np, C, iou_thr, out_keys_all, out_tp_all, offset,
                              gt_sum_all, nullptr, stream);
}

extern "C" int yh_map_reduce(const uint64_t *keys, const uint8_t *tp, int64_t nrec, const int32_t *gt_per_class, int C,
                             float *out_ap, float *out_map, void *stream)
{
    YH_REQUIRE(C >= 1 && nrec >= 0 && nrec < (1ll << 31), "map_reduce: bad sizes");
    YH_REQUIRE(gt_per_class && out_map && (nrec == 0 || (keys && tp)), "map_reduce: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    { int rc_ = keep_pool(); if (rc_ != YH_OK) return rc_; }
    const int end_bit = bits_for(C);
    AsyncBuf sk(st), stp(st), cum(st), start(st), ap(st), tmp(st);
    YH_CUDA(sk.alloc(8 * nrec)); YH_CUDA(stp.alloc(nrec)); YH_CUDA(cum.alloc(4 * nrec));
    YH_CUDA(start.alloc(8 * (C + 1))); YH_CUDA(ap.alloc(4 * C));
    if (nrec > 0) {
        size_t tb1 = 0, tb2 = 0;
        YH_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb1, keys, sk.as<uint64_t>(), tp, stp.as<uint8_t>(),
                                                static_cast<int>(nrec), 0, end_bit, st));
        cub::TransformInputIterator<int, U8ToI32, const uint8_t *> it(stp.as<uint8_t>(), U8ToI32());
        YH_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb2, it, cum.as<int>(), static_cast<int>(nrec), st));
        YH_CUDA(tmp.alloc(std::max(tb1, tb2)));
        size_t tb = std::max(tb1, tb2);
        YH_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys, sk.as<uint64_t>(), tp, stp.as<uint8_t>(),
                                                static_cast<int>(nrec), 0, end_bit, st));
        tb = std::max(tb1, tb2);
        YH_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, it, cum.as<int>(), static_cast<int>(nrec), st));
        count_launch(6);
    }
    class_starts_kernel<<<blocks_for(C + 1, 128), 128, 0, st>>>(sk.as<uint64_t>(), nrec, C, start.as<int64_t>());
    YH_LAUNCH_CHECK("class_starts_kernel");
    ap_kernel<<<C, 256, 0, st>>>(cum.as<int>(), start.as<int64_t>(), gt_per_class, ap.as<float>());
    YH_LAUNCH_CHECK("ap_kernel");
    map_mean_kernel<<<1, 32, 0, st>>>(ap.as<float>(), C, out_ap, out_map);
    YH_LAUNCH_CHECK("map_mean_kernel");
    return YH_OK;
}

extern "C" int yh_rows_append(const float *boxes, const int32_t *count, int64_t n, int M, int64_t img_base,
                              float *out_rows, int64_t out_capacity, int64_t *row_cursor, void *stream)
{
    YH_REQUIRE(n >= 0 && M >= 1 && out_capacity >= 0, "rows_append: bad sizes");
    if (n == 0) return YH_OK;
    YH_REQUIRE(boxes && count && row_cursor && (out_capacity == 0 || out_rows), "rows_append: null pointer");
    YH_REQUIRE(n < (1ll << 31), "rows_append: too many images in one call");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    { int rc_ = keep_pool(); if (rc_ != YH_OK) return rc_; }
    AsyncBuf offs(st), tmp(st);
    YH_CUDA(offs.alloc(8 * n));
    size_t tb = 0;
    cub::TransformInputIterator<int64_t, I32ToI64, const int *> it(count, I32ToI64());
    YH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, offs.as<int64_t>(), static_cast<int>(n), st));
    YH_CUDA(tmp.alloc(tb));
    YH_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, it, offs.as<int64_t>(), static_cast<int>(n), st));
    count_launch(2);
    rows_append_kernel<<<blocks_for(n * M, 256), 256, 0, st>>>(boxes, count, offs.as<int64_t>(), n, M, img_base, out_rows,
                                                               out_capacity, row_cursor);
    YH_LAUNCH_CHECK("rows_append_kernel");
    cursor_advance_kernel<<<1, 32, 0, st>>>(count, offs.as<int64_t>(), n, row_cursor);
    YH_LAUNCH_CHECK("cursor_advance_kernel");
    return YH_OK;
}
