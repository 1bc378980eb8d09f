// yh_map.cu - K6: the evaluator's accumulation step and the IoU matching behind mAP.  sm_100a, hand-written
// (no CUB): one kernel per update, no host synchronisation anywhere.
//
// Replaces MeanAveragePrecision.update_state (utils.py:470-491: prefix img_idx, append) and the matching loop
// of mean_average_precision (utils.py:373-422, change_tensor :280-299).
//
//   eval_update_kernel   one launch per update_state: a single-pass chained scan (decoupled look-back over the
//                        CTAs' tile totals) turns the kept-row counts of the batch into global row offsets; one
//                        warp per image appends its rows [img, cls, conf, cx, cy, w, h] to the prediction and
//                        ground-truth row buffers in image order and - because a detection can only ever claim a
//                        ground truth of its own image, and NMS already emits an image's detections in the
//                        reference's processing order (confidence descending, stable) - matches the image right
//                        there: best same-class IoU (strict >, first wins, utils.py:386-393), TP iff that IoU is
//                        > thr and the ground truth is still free (utils.py:395-418).  Out comes one packed record
//                        per detection (yh_common.cuh) and the per-class ground-truth counts; result() is then
//                        only the reduce stage (yh_map_reduce.cu).
//   general path         mean_average_precision() on arbitrary rows: prep (claim table, per-class GT counts,
//                        optionally image keys for a radix sort when the ground-truth rows are not grouped by
//                        image) -> one thread per detection: binary search of its image's ground truths, best IoU,
//                        atomicMin of (confidence desc, row asc) into the claim of that ground truth -> TP iff the
//                        detection owns the ground truth it points at: the sequential claim loop without sequencing.
#include <cstdlib>
#include <map>
#include <mutex>

#include "yh_common.cuh"
#include "yh_map_internal.cuh"

namespace yh {

int keep_pool()
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

struct EvalArgs {
    const float *boxes[2];            // padded NMS output (n, M, 6): [0] predictions, [1] ground truth (nullable)
    const int32_t *count[2];          // (n)
    float *rows[2];                   // (cap, 7) append buffers
    long long cap[2];
    long long *cursor[2];             // device row counters
    unsigned long long *rec;          // (cap[0]) packed records, parallel to the prediction rows (MATCH)
    int32_t *gt_per_class;            // (C) accumulates (MATCH)
    long long n, img_base;
    int M, C, tile;
    float iou_thr;
};

template <bool MATCH>
__global__ void __launch_bounds__(256) eval_update_kernel(EvalArgs a, ScanWs *ws)
{
    extern __shared__ int dyn[];                                  // [2][tile] row offsets inside the tile, [C] GT histogram
    __shared__ int tile_id_s, last_s;
    __shared__ long long base_s[2];
    __shared__ int warp_tot[8];
    __shared__ uint32_t claimed_s[8][YH_MAX_CELLS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsets = a.boxes[1] ? 2 : 1;
    const int ntiles = gridDim.x;
    if (tid == 0) tile_id_s = static_cast<int>(atomicAdd(&ws->ticket, 1u));
    __syncthreads();
    const int b = tile_id_s;
    const long long img0 = static_cast<long long>(b) * a.tile;
    const int nimg = static_cast<int>(min(static_cast<long long>(a.tile), a.n - img0));
    int *pre[2] = {dyn, dyn + a.tile};
    int *hist = dyn + 2 * a.tile;
    long long agg[2] = {0, 0};

    // ---- 1. kept-row counts of the tile -> offsets inside the tile
    for (int s = 0; s < nsets; ++s) {
        int run = 0;
        for (int i0 = 0; i0 < nimg; i0 += 256) {
            const int i = i0 + tid;
            const int c = (i < nimg) ? min(max(a.count[s][img0 + i], 0), a.M) : 0;
            int v = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            if (lane == 31) warp_tot[warp] = v;
            __syncthreads();
            int woff = 0, all = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const int t = warp_tot[w];
                if (w < warp) woff += t;
                all += t;
            }
            if (i < nimg) pre[s][i] = run + woff + v - c;
            run += all;
            __syncthreads();
        }
        agg[s] = run;
    }
    if (MATCH)
        for (int c = tid; c < a.C; c += 256) hist[c] = 0;

    // ---- 2. publish the tile total, look back for the row offset of the tile (warp s <-> set s)
    if (warp < nsets) {
        const int s = warp;
        volatile unsigned long long *st = ws->st[s];
        long long excl;
        if (b == 0) {
            excl = *a.cursor[s];                                  // only tile 0 reads the cursor, only the last tile writes it
        } else {
            if (lane == 0) st[b] = kFlagAgg | static_cast<unsigned long long>(agg[s]);
            excl = lookback(st, b, lane);
        }
        if (lane == 0) {
            st[b] = kFlagIncl | static_cast<unsigned long long>(excl + agg[s]);
            base_s[s] = excl;
            if (b == ntiles - 1) *a.cursor[s] = excl + agg[s];
        }
    }
    __syncthreads();

    // ---- 3. one warp per image: append the rows (utils.py:476-489), match the image (utils.py:373-422)
    for (int i = warp; i < nimg; i += 8) {
        const long long img = img0 + i;
        const float imgf = static_cast<float>(a.img_base + img);         // utils.py:476: the index travels as float32
        int cnt[2] = {0, 0};
        long long row0[2] = {0, 0};
        for (int s = 0; s < nsets; ++s) {
            cnt[s] = min(max(a.count[s][img], 0), a.M);
            row0[s] = base_s[s] + pre[s][i];
            const float *src = a.boxes[s] + img * a.M * 6;
            for (int slot = lane; slot < cnt[s]; slot += 32) {
                const long long r = row0[s] + slot;
                if (r >= a.cap[s]) continue;
                const float2 v0 = *reinterpret_cast<const float2 *>(src + slot * 6);
                const float2 v1 = *reinterpret_cast<const float2 *>(src + slot * 6 + 2);
                const float2 v2 = *reinterpret_cast<const float2 *>(src + slot * 6 + 4);
                float *o = a.rows[s] + r * 7;
                o[0] = imgf; o[1] = v0.x; o[2] = v0.y; o[3] = v1.x; o[4] = v1.y; o[5] = v2.x; o[6] = v2.y;
            }
        }
        if (!MATCH) continue;
        const long long r0 = row0[0];
        match_image(a.boxes[0] + img * a.M * 6, a.boxes[1] + img * a.M * 6, cnt[0], cnt[1], a.C, a.iou_thr, claimed_s[warp], hist, lane,
                    [&](int d, unsigned long long rec) {
                        const long long r = r0 + d;
                        if (r < a.cap[0]) a.rec[r] = rec;
                    });
        __syncwarp();
    }
    __syncthreads();
    if (MATCH)
        for (int c = tid; c < a.C; c += 256)
            if (hist[c]) atomicAdd(a.gt_per_class + c, hist[c]);

    // ---- 4. the last CTA to finish returns the scan state to all-zero
    if (tid == 0) {
        __threadfence();
        last_s = (atomicAdd(&ws->done, 1u) == static_cast<unsigned>(ntiles - 1));
    }
    __syncthreads();
    if (last_s) {
        for (int j = tid; j < ntiles; j += 256) { ws->st[0][j] = 0; ws->st[1][j] = 0; }
        if (tid == 0) { ws->ticket = 0; ws->done = 0; }
    }
}

// scan state per (device, stream): launches on one stream are ordered, so they can share it
static std::mutex g_scan_mu;
static std::map<std::pair<int, cudaStream_t>, ScanWs *> g_scan_ws;

int scan_ws_for(cudaStream_t st, ScanWs **out)
{
    int dev = 0;
    YH_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_scan_mu);
    auto key = std::make_pair(dev, st);
    auto it = g_scan_ws.find(key);
    if (it == g_scan_ws.end()) {
        ScanWs *p = nullptr;
        YH_CUDA(cudaMalloc(&p, sizeof(ScanWs)));
        cudaError_t e = cudaMemsetAsync(p, 0, sizeof(ScanWs), st);
        if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaMemsetAsync"); }
        it = g_scan_ws.emplace(key, p).first;
    }
    *out = it->second;
    return YH_OK;
}

static int eval_update_impl(EvalArgs a, bool match, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScanWs *ws = nullptr;
    int rc = scan_ws_for(st, &ws);
    if (rc != YH_OK) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        YH_CUDA(cudaFuncSetAttribute(eval_update_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        YH_CUDA(cudaFuncSetAttribute(eval_update_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_done = true;
    }
    const long long n_all = a.n, base_all = a.img_base;
    const float *boxes0[2] = {a.boxes[0], a.boxes[1]};
    const int32_t *count0[2] = {a.count[0], a.count[1]};
    const long long chunk = static_cast<long long>(kScanMaxTiles) * 1024;       // images per launch
    for (long long lo = 0; lo < n_all; lo += chunk) {
        a.n = std::min(chunk, n_all - lo);
        a.img_base = base_all + lo;
        for (int s = 0; s < 2; ++s) {
            a.boxes[s] = boxes0[s] ? boxes0[s] + lo * a.M * 6 : nullptr;
            a.count[s] = count0[s] ? count0[s] + lo : nullptr;
        }
        // tiles: enough CTAs to fill the machine four times over, at most kScanMaxTiles, a multiple of 8 images each
        long long tile = (a.n + 4 * sm_count() - 1) / (4 * sm_count());
        tile = std::max<long long>(8, std::min<long long>(1024, (tile + 7) / 8 * 8));
        if (const char *v = getenv("YH_EVAL_TILE")) {                       // experiments: any tile size is correct
            const long long t = atoll(v);
            if (t >= 8 && t <= 1024 && (a.n + t - 1) / t <= kScanMaxTiles) tile = t / 8 * 8;
        }
        a.tile = static_cast<int>(tile);
        const int ntiles = static_cast<int>((a.n + tile - 1) / tile);
        const size_t smem = sizeof(int) * (2 * static_cast<size_t>(a.tile) + (match ? a.C : 0));
        if (match) eval_update_kernel<true><<<ntiles, 256, smem, st>>>(a, ws);
        else eval_update_kernel<false><<<ntiles, 256, smem, st>>>(a, ws);
        YH_LAUNCH_CHECK("eval_update_kernel");
    }
    return YH_OK;
}

// ---- general path: arbitrary rows ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) map_prep_kernel(const float *__restrict__ true_rows, long long nt_max,
                                                       const long long *__restrict__ nt_dev, int C,
                                                       unsigned long long *__restrict__ claim, unsigned long long *__restrict__ gk,
                                                       int32_t *__restrict__ gt_per_class)
{
    extern __shared__ int hist[];
    for (int c = threadIdx.x; c < C; c += blockDim.x) hist[c] = 0;
    __syncthreads();
    const long long nt = nt_dev ? min(max(*nt_dev, 0ll), nt_max) : nt_max;
    for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < nt;
         g += static_cast<long long>(gridDim.x) * blockDim.x) {
        claim[g] = ~0ull;
        const uint32_t c = class_of(true_rows[7 * g + 1], C);
        if (c < static_cast<uint32_t>(C)) atomicAdd(&hist[c], 1);
        if (gk) gk[g] = (static_cast<unsigned long long>(orderable(true_rows[7 * g])) << 32) | static_cast<unsigned long long>(g);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        if (hist[c]) atomicAdd(gt_per_class + c, hist[c]);
}

// one thread per detection.  Ground truths of an image: rows [lo, ...) with the image's key, either in the rows
// themselves (grouped by image, nondecreasing) or through the sorted (image key << 32 | row) list gk.
__global__ void __launch_bounds__(128) map_match_kernel(const float *__restrict__ true_rows, long long nt_max,
                                                        const long long *__restrict__ nt_dev, const float *__restrict__ pred_rows,
                                                        long long np_max, const long long *__restrict__ np_dev, int C, float iou_thr,
                                                        const unsigned long long *__restrict__ gk, const long long *__restrict__ gk_n,
                                                        unsigned long long *__restrict__ claim, int32_t *__restrict__ hit)
{
    const long long np = np_dev ? min(max(*np_dev, 0ll), np_max) : np_max;
    long long nt = nt_dev ? min(max(*nt_dev, 0ll), nt_max) : nt_max;
    if (gk && gk_n) nt = *gk_n;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < np;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float *d = pred_rows + 7 * i;
        const uint32_t dc = class_of(d[1], C);
        int32_t h = -1;
        if (dc < static_cast<uint32_t>(C)) {
            const uint32_t ik = orderable(d[0]);
            long long lo = 0, hi = nt;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                const uint32_t mk = gk ? static_cast<uint32_t>(gk[mid] >> 32) : orderable(true_rows[7 * mid]);
                if (mk < ik) lo = mid + 1; else hi = mid;
            }
            const float dx = d[3], dy = d[4], dw = d[5], dh = d[6];
            float best = 0.0f;                               // utils.py:382
            long long bj = 0;                                // utils.py:383
            for (long long m = lo; m < nt; ++m) {            // utils.py:386: row order inside the image
                long long g = m;
                if (gk) {
                    const unsigned long long e = gk[m];
                    if (static_cast<uint32_t>(e >> 32) != ik) break;
                    g = static_cast<long long>(e & 0xffffffffull);
                } else if (orderable(true_rows[7 * m]) != ik) {
                    break;
                }
                const float *t = true_rows + 7 * g;
                if (class_of(t[1], C) != dc) continue;       // utils.py:330, :378
                const float v = iou_ref(dx, dy, dw, dh, t[3], t[4], t[5], t[6]);    // utils.py:387 (det, gt)
                if (v > best) { best = v; bj = g; }          // utils.py:389
            }
            if (best > iou_thr) {                            // utils.py:395
                h = static_cast<int32_t>(bj);
                atomicMin(claim + bj, (static_cast<unsigned long long>(~orderable(d[2])) << 32) | static_cast<unsigned long long>(i));
            }
        }
        hit[i] = h;
    }
}

__global__ void __launch_bounds__(256) map_tp_kernel(const float *__restrict__ pred_rows, long long np_max,
                                                     const long long *__restrict__ np_dev, int C,
                                                     const int32_t *__restrict__ hit, const unsigned long long *__restrict__ claim,
                                                     unsigned long long *__restrict__ rec)
{
    const long long np = np_dev ? min(max(*np_dev, 0ll), np_max) : np_max;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < np;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float conf = pred_rows[7 * i + 2];
        const int32_t h = hit[i];
        const unsigned long long mine = (static_cast<unsigned long long>(~orderable(conf)) << 32) | static_cast<unsigned long long>(i);
        const uint32_t tp = (h >= 0 && claim[h] == mine) ? 1u : 0u;               // utils.py:408-418
        rec[i] = make_rec(class_of(pred_rows[7 * i + 1], C), conf, tp);
    }
}

static inline int blocks_for(long long n, int t, int cap)
{
    return static_cast<int>(std::max<long long>(1, std::min<long long>((n + t - 1) / t, cap)));
}

size_t match_ws_bytes(int64_t nt, int64_t np, bool sorted)
{
    size_t b = align_up(static_cast<size_t>(nt) * 8, 256) + align_up(static_cast<size_t>(np) * 4, 256) + 256;
    if (!sorted) b += align_up(static_cast<size_t>(nt) * 8, 256) + radix_ws_bytes(nt, 1);
    return b;
}

}  // namespace yh

using namespace yh;

extern "C" int yh_eval_update(const float *pred_boxes, const int32_t *pred_count, const float *true_boxes,
                              const int32_t *true_count, int64_t n, int M, int64_t img_base, int C, float iou_thr,
                              float *pred_rows, int64_t pred_capacity, float *true_rows, int64_t true_capacity,
                              uint64_t *rec, int64_t *cursors, int32_t *gt_per_class, void *stream)
{
    YH_REQUIRE(n >= 0 && M >= 1 && M <= YH_MAX_CELLS, "eval_update: bad sizes (M = %d, limit %d)", M, YH_MAX_CELLS);
    YH_REQUIRE(C >= 1 && C <= kMaxMapClasses, "eval_update: C = %d outside [1, %d]", C, kMaxMapClasses);
    if (n == 0) return YH_OK;
    YH_REQUIRE(pred_boxes && pred_count && true_boxes && true_count && cursors && rec && gt_per_class,
               "eval_update: null pointer");
    YH_REQUIRE(pred_capacity >= 0 && true_capacity >= 0 && (pred_capacity == 0 || pred_rows) && (true_capacity == 0 || true_rows),
               "eval_update: bad row buffers");
    EvalArgs a{};
    a.boxes[0] = pred_boxes; a.boxes[1] = true_boxes;
    a.count[0] = pred_count; a.count[1] = true_count;
    a.rows[0] = pred_rows; a.rows[1] = true_rows;
    a.cap[0] = pred_capacity; a.cap[1] = true_capacity;
    a.cursor[0] = reinterpret_cast<long long *>(cursors); a.cursor[1] = reinterpret_cast<long long *>(cursors) + 1;
    a.rec = reinterpret_cast<unsigned long long *>(rec);
    a.gt_per_class = gt_per_class;
    a.n = n; a.img_base = img_base; a.M = M; a.C = C; a.iou_thr = iou_thr;
    return eval_update_impl(a, true, stream);
}

extern "C" int yh_rows_append(const float *boxes, const int32_t *count, int64_t n, int M, int64_t img_base,
                              float *out_rows, int64_t out_capacity, int64_t *row_cursor, void *stream)
{
    YH_REQUIRE(n >= 0 && M >= 1 && out_capacity >= 0, "rows_append: bad sizes");
    if (n == 0) return YH_OK;
    YH_REQUIRE(boxes && count && row_cursor && (out_capacity == 0 || out_rows), "rows_append: null pointer");
    EvalArgs a{};
    a.boxes[0] = boxes; a.count[0] = count; a.rows[0] = out_rows; a.cap[0] = out_capacity;
    a.cursor[0] = reinterpret_cast<long long *>(row_cursor);
    a.n = n; a.img_base = img_base; a.M = M; a.C = 1; a.iou_thr = 0.0f;
    return eval_update_impl(a, false, stream);
}

extern "C" int yh_map_match(const float *true_rows, int64_t nt, const int64_t *nt_dev, const float *pred_rows, int64_t np,
                            const int64_t *np_dev, int C, float iou_thr, int flags, uint64_t *out_rec,
                            int32_t *out_gt_per_class, void *workspace, size_t workspace_bytes, void *stream)
{
    YH_REQUIRE(C >= 1 && C <= kMaxMapClasses, "map_match: C = %d outside [1, %d]", C, kMaxMapClasses);
    YH_REQUIRE(nt >= 0 && np >= 0 && nt < (1ll << 31) && np < (1ll << 31), "map_match: bad row counts");
    YH_REQUIRE(out_gt_per_class != nullptr, "map_match: out_gt_per_class is null");
    YH_REQUIRE((nt == 0 || true_rows) && (np == 0 || (pred_rows && out_rec)), "map_match: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool sorted = (flags & YH_MAP_TRUE_ROWS_BY_IMAGE) != 0;
    AsyncBuf own(st);
    const size_t need = match_ws_bytes(nt, np, sorted);
    if (!workspace) {
        int rc = keep_pool();
        if (rc != YH_OK) return rc;
        YH_CUDA(own.alloc(need));
        workspace = own.p;
        workspace_bytes = need;
    }
    YH_REQUIRE(workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "map_match: a 256-byte aligned workspace of %zu bytes is needed", need);
    unsigned char *w = static_cast<unsigned char *>(workspace);
    unsigned long long *claim = reinterpret_cast<unsigned long long *>(w); w += align_up(static_cast<size_t>(nt) * 8, 256);
    int32_t *hit = reinterpret_cast<int32_t *>(w); w += align_up(static_cast<size_t>(np) * 4, 256);
    long long *gk_n = reinterpret_cast<long long *>(w); w += 256;
    unsigned long long *gk = nullptr;
    if (!sorted) { gk = reinterpret_cast<unsigned long long *>(w); w += align_up(static_cast<size_t>(nt) * 8, 256); }

    YH_CUDA(cudaMemsetAsync(out_gt_per_class, 0, sizeof(int32_t) * C, st));
    const int cap = 8 * sm_count();
    map_prep_kernel<<<blocks_for(nt, 256, cap), 256, sizeof(int) * C, st>>>(true_rows, nt, reinterpret_cast<const long long *>(nt_dev), C,
                                                                             claim, gk, out_gt_per_class);
    YH_LAUNCH_CHECK("map_prep_kernel");
    if (np == 0) return YH_OK;
    const unsigned long long *gk_sorted = nullptr;
    if (!sorted) {                                   // ground-truth rows in any order: stable radix sort by image
        ReduceArgs ra{};
        ra.in.nseg = 1;
        ra.in.ptr[0] = gk;
        ra.in.cnt_dev[0] = reinterpret_cast<const long long *>(nt_dev);
        ra.in.cnt_max[0] = nt;
        ra.bit_lo = 32;
        ra.npass = 4;
        ra.mode = YH_RADIX_SORT_ONLY;
        ra.C = 1;
        ra.out_n = gk_n;
        int rc = radix_launch(ra, nt, w, radix_ws_bytes(nt, 1), st);
        if (rc != YH_OK) return rc;
        gk_sorted = ra.buf[(ra.npass - 1) & 1];
    }
    map_match_kernel<<<blocks_for(np, 128, 16 * sm_count()), 128, 0, st>>>(
        true_rows, nt, reinterpret_cast<const long long *>(nt_dev), pred_rows, np, reinterpret_cast<const long long *>(np_dev), C,
        iou_thr, gk_sorted, gk_sorted ? gk_n : nullptr, claim, hit);
    YH_LAUNCH_CHECK("map_match_kernel");
    map_tp_kernel<<<blocks_for(np, 256, cap), 256, 0, st>>>(pred_rows, np, reinterpret_cast<const long long *>(np_dev), C, hit, claim,
                                                             reinterpret_cast<unsigned long long *>(out_rec));
    YH_LAUNCH_CHECK("map_tp_kernel");
    return YH_OK;
}
