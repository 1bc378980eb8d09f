// yh_eval_fused.cu - K6f: MeanAveragePrecision.update_state (utils.py:470-491) in ONE launch.  sm_100a.
//
// The evaluator's accumulation step used to be three launches - fused decode + NMS of the predictions, of the ground
// truth (utils.py:475 / :480), and yh_eval_update (append + per-image matching) - with two padded (n, S*S, 6) row
// tensors written and re-read in between.  At the batch sizes an evaluator sees (BASELINE cfg4: 5,000 images) every one
// of those launches is latency, not bandwidth.  Here one warp does everything for an image while its rows are still in
// shared memory:
//   1  decode + NMS of the image's prediction cells and of its label cells (decode_cell / nms_warp of
//      yh_decode_nms_impl.cuh: the very code of the fused decode + NMS kernels, kept rows written to SHARED memory)
//   2  matching of the image (match_image: utils.py:373-422), packed records to shared memory
//   3  the CTA's kept-row counts -> global row offsets by the single-pass chained scan of yh_map.cu (ticket, publish,
//      look back), then the rows [img, cls, conf, cx, cy, w, h] of both sets and the records go out coalesced
// Same outputs, bit for bit, as the three-launch path (tests/test_gpu_map.py::test_fused_update_state_equals_three_launches).
// Grids of more than 64 cells keep the three-launch path (their NMS wants the team kernel).
#include <mutex>

#include "yh_decode_nms_impl.cuh"
#include "yh_map_internal.cuh"

namespace yh {

struct EvalStateArgs {
    const float *y[2];                // (n, M, D): [0] y_pred, [1] y_true
    float *rows[2];                   // (cap, 7) append buffers: [0] predictions, [1] ground truth
    long long cap[2];
    long long *cursor[2];
    unsigned long long *rec;          // (cap[0]) packed records, parallel to the prediction rows
    int32_t *gt_per_class;            // (C) accumulates
    long long n, img_base;
    int tile;                         // images per CTA
    float match_thr;                  // utils.py:496 -> mean_average_precision's iou_threshold
};

constexpr int kEvalWarps = 8;

template <int NS, int CT, int BT>
__global__ void __launch_bounds__(32 * kEvalWarps) eval_state_kernel(EvalStateArgs a, NmsCfg cfg, ScanWs *sw)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int tile_id_s, last_s;
    __shared__ long long base_s[2];
    __shared__ uint32_t claimed_s[kEvalWarps][YH_MAX_CELLS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = cfg.M, D = cfg.D, C = cfg.C;
    // shared memory: per-warp NMS workspace | kept rows [2][tile][M * 6] | records [tile][M] | counts / offsets [2][2][tile] | hist [C]
    unsigned char *p = smem;
    unsigned char *ws_base = p;                                   p += static_cast<size_t>(kEvalWarps) * cfg.ws_bytes;
    float *rows_s = reinterpret_cast<float *>(p);                 p += static_cast<size_t>(2) * a.tile * M * 6 * 4;
    unsigned long long *rec_s = reinterpret_cast<unsigned long long *>(p);   p += static_cast<size_t>(a.tile) * M * 8;
    int *cnt_s = reinterpret_cast<int *>(p);                      p += static_cast<size_t>(2) * a.tile * 4;
    int *pre_s = reinterpret_cast<int *>(p);                      p += static_cast<size_t>(2) * a.tile * 4;
    int *hist = reinterpret_cast<int *>(p);

    const int ntiles = gridDim.x;
    if (tid == 0) tile_id_s = static_cast<int>(atomicAdd(&sw->ticket, 1u));
    for (int c = tid; c < C; c += 32 * kEvalWarps) hist[c] = 0;
    WarpWs<NS, false> ws(ws_base + static_cast<size_t>(warp) * cfg.ws_bytes);
    for (int i = lane; i < cfg.tbl_rows * NS; i += 32) ws.tbl[i] = 0u;
    __syncthreads();
    const int b = tile_id_s;
    const long long img0 = static_cast<long long>(b) * a.tile;
    const int nimg = static_cast<int>(min(static_cast<long long>(a.tile), a.n - img0));

    float colf[NS], rowf[NS];
    bool valid[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        const int cell = lane + 32 * t;
        valid[t] = cell < M;
        rowf[t] = static_cast<float>(cell / cfg.S);
        colf[t] = static_cast<float>(cell % cfg.S);
    }

    // ---- 1 + 2. one warp per image: decode + NMS of both tensors into shared memory, then the matching
    for (int i = warp; i < nimg; i += kEvalWarps) {
        const long long img = img0 + i;
        int K[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const float *base = a.y[s] + img * M * D;
            float conf[NS];
            float4 box[NS];
            int cls[NS];
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                conf[t] = -INFINITY; cls[t] = 0; box[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid[t]) decode_cell<CT, BT>(base + (lane + 32 * t) * D, cfg, colf[t], rowf[t], cls[t], conf[t], box[t]);
            }
            K[s] = nms_warp<NS, false>(conf, box, cls, valid, cfg, ws, rows_s + (static_cast<size_t>(s) * a.tile + i) * M * 6, nullptr);
            if (lane == 0) cnt_s[s * a.tile + i] = K[s];
        }
        __syncwarp();
        unsigned long long *rec_i = rec_s + static_cast<size_t>(i) * M;
        match_image(rows_s + static_cast<size_t>(i) * M * 6, rows_s + (static_cast<size_t>(a.tile) + i) * M * 6, K[0], K[1], C, a.match_thr,
                    claimed_s[warp], hist, lane, [&](int d, unsigned long long rec) { rec_i[d] = rec; });
    }
    __syncthreads();

    // ---- 3. kept-row counts of the tile -> offsets inside the tile (warp s scans set s), tile totals -> chained scan
    if (warp < 2) {
        const int s = warp;
        int run = 0;
        for (int i0 = 0; i0 < nimg; i0 += 32) {
            const int i = i0 + lane;
            const int c = i < nimg ? cnt_s[s * a.tile + i] : 0;
            int v = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            if (i < nimg) pre_s[s * a.tile + i] = run + v - c;
            run += __shfl_sync(0xffffffffu, v, 31);
        }
        volatile unsigned long long *st = sw->st[s];
        long long excl;
        if (b == 0) {
            excl = *a.cursor[s];                                  // only tile 0 reads the cursor, only the last tile writes it
        } else {
            if (lane == 0) st[b] = kFlagAgg | static_cast<unsigned long long>(run);
            excl = lookback(st, b, lane);
        }
        if (lane == 0) {
            st[b] = kFlagIncl | static_cast<unsigned long long>(excl + run);
            base_s[s] = excl;
            if (b == ntiles - 1) *a.cursor[s] = excl + run;
        }
    }
    __syncthreads();

    // ---- rows [img, cls, conf, cx, cy, w, h] (utils.py:476-489) and records, coalesced over (row, field)
    for (int i = warp; i < nimg; i += kEvalWarps) {
        const float imgf = static_cast<float>(a.img_base + img0 + i);    // utils.py:476: the index travels as float32
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int cnt = cnt_s[s * a.tile + i];
            const long long row0 = base_s[s] + pre_s[s * a.tile + i];
            const long long room = a.cap[s] - row0;                      // rows past the capacity are dropped (the host sizes the buffers)
            const int take = static_cast<int>(max(0ll, min(static_cast<long long>(cnt), room)));
            const float *src = rows_s + (static_cast<size_t>(s) * a.tile + i) * M * 6;
            float *dst = a.rows[s] + row0 * 7;
            for (int e = lane; e < take * 7; e += 32) {
                const int r = e / 7, f = e - 7 * r;
                dst[e] = f == 0 ? imgf : src[r * 6 + f - 1];
            }
            if (s == 0) {
                const unsigned long long *rec_i = rec_s + static_cast<size_t>(i) * M;
                for (int d = lane; d < take; d += 32) a.rec[row0 + d] = rec_i[d];
            }
        }
    }
    for (int c = tid; c < C; c += 32 * kEvalWarps)
        if (hist[c]) atomicAdd(a.gt_per_class + c, hist[c]);

    // ---- the last CTA to finish returns the scan state to all-zero
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        last_s = (atomicAdd(&sw->done, 1u) == static_cast<unsigned>(ntiles - 1));
    }
    __syncthreads();
    if (last_s) {
        for (int j = tid; j < ntiles; j += 32 * kEvalWarps) { sw->st[0][j] = 0; sw->st[1][j] = 0; }
        if (tid == 0) { sw->ticket = 0; sw->done = 0; }
    }
}

template <int NS, int CT, int BT>
static int launch_eval_state(EvalStateArgs a, NmsCfg cfg, cudaStream_t st)
{
    cfg.ws_bytes = WarpWs<NS, false>::bytes(cfg.tbl_rows);
    ScanWs *sw = nullptr;
    int rc = scan_ws_for(st, &sw);
    if (rc != YH_OK) return rc;
    auto kern = eval_state_kernel<NS, CT, BT>;
    auto smem_of = [&](int tile) {
        return static_cast<size_t>(kEvalWarps) * cfg.ws_bytes + static_cast<size_t>(tile) * cfg.M * (2 * 6 * 4 + 8) + static_cast<size_t>(tile) * 16 +
               static_cast<size_t>(cfg.C) * 4 + 64;
    };
    if (smem_of(kEvalWarps) > 200 * 1024) {
        set_error("eval_update_state: C = %d needs more shared memory than an SM has", cfg.C);
        return YH_ERR_UNSUPPORTED;
    }
    static std::mutex mu;
    static size_t smem_set[3][64] = {};
    const int which = CT > 0 ? 0 : NS;
    int dev = 0;
    YH_CUDA(cudaGetDevice(&dev));
    const long long n_all = a.n, base_all = a.img_base;
    const float *y0[2] = {a.y[0], a.y[1]};
    const long long per_img = static_cast<long long>(cfg.M) * cfg.D;
    for (long long lo = 0; lo < n_all;) {
        // tile: images per CTA, a multiple of the warps of a CTA; enough CTAs to fill the machine about three times over,
        // few enough images that their rows fit shared memory, at most kScanMaxTiles tiles per launch
        long long tile = (n_all - lo + 3 * sm_count() - 1) / (3 * sm_count());
        tile = std::max<long long>(kEvalWarps, std::min<long long>(4 * kEvalWarps, (tile + kEvalWarps - 1) / kEvalWarps * kEvalWarps));
        while (tile > kEvalWarps && smem_of(static_cast<int>(tile)) > 72 * 1024) tile -= kEvalWarps;
        if (const char *v = getenv("YH_EVAL_TILE")) {                       // experiments: any tile size that fits is correct
            const long long t = atoll(v) / kEvalWarps * kEvalWarps;
            if (t >= kEvalWarps && smem_of(static_cast<int>(t)) <= 200 * 1024) tile = t;
        }
        a.n = std::min(n_all - lo, tile * kScanMaxTiles);
        a.img_base = base_all + lo;
        a.y[0] = y0[0] + lo * per_img;
        a.y[1] = y0[1] + lo * per_img;
        a.tile = static_cast<int>(tile);
        const size_t smem = smem_of(a.tile);
        {
            std::lock_guard<std::mutex> lock(mu);
            if (dev >= 0 && dev < 64 && smem > smem_set[which][dev]) {
                YH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                smem_set[which][dev] = smem;
            }
        }
        const int ntiles = static_cast<int>((a.n + tile - 1) / tile);
        kern<<<ntiles, 32 * kEvalWarps, smem, st>>>(a, cfg, sw);
        YH_LAUNCH_CHECK("eval_state_kernel");
        lo += a.n;
    }
    return YH_OK;
}

}  // namespace yh

using namespace yh;

extern "C" int yh_eval_update_state(const float *y_true, const float *y_pred, int64_t n, int S, int B, int C, float nms_iou_thr,
                                    float nms_conf_thr, int64_t img_base, float match_iou_thr, float *pred_rows,
                                    int64_t pred_capacity, float *true_rows, int64_t true_capacity, uint64_t *rec,
                                    int64_t *cursors, int32_t *gt_per_class, int restart, void *stream)
{
    NmsCfg cfg;
    int rc = fill_cfg(cfg, S, B, C, nms_iou_thr, nms_conf_thr);
    if (rc != YH_OK) return rc;
    YH_REQUIRE(n >= 0, "eval_update_state: n < 0");
    YH_REQUIRE(C <= kMaxMapClasses, "eval_update_state: C = %d outside [1, %d]", C, kMaxMapClasses);
    if (cfg.M > 64) {
        set_error("eval_update_state: grids of more than 64 cells take the three-launch path (yh_decode_nms x 2 + yh_eval_update)");
        return YH_ERR_UNSUPPORTED;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (restart) {                    // utils.py:484-486: the first image of an epoch overwrites - cursors and counts start at zero
        YH_REQUIRE(cursors && gt_per_class, "eval_update_state: null pointer");
        YH_CUDA(cudaMemsetAsync(cursors, 0, 2 * sizeof(int64_t), st));
        YH_CUDA(cudaMemsetAsync(gt_per_class, 0, static_cast<size_t>(C) * sizeof(int32_t), st));
    }
    if (n == 0) return YH_OK;
    YH_REQUIRE(y_true && y_pred && cursors && rec && gt_per_class, "eval_update_state: null pointer");
    YH_REQUIRE(pred_capacity >= 0 && true_capacity >= 0 && (pred_capacity == 0 || pred_rows) && (true_capacity == 0 || true_rows),
               "eval_update_state: bad row buffers");
    YH_REQUIRE((reinterpret_cast<uintptr_t>(y_true) | reinterpret_cast<uintptr_t>(y_pred)) % 8 == 0,
               "eval_update_state: y_true / y_pred must be 8-byte aligned");
    EvalStateArgs a{};
    a.y[0] = y_pred; a.y[1] = y_true;
    a.rows[0] = pred_rows; a.rows[1] = true_rows;
    a.cap[0] = pred_capacity; a.cap[1] = true_capacity;
    a.cursor[0] = reinterpret_cast<long long *>(cursors); a.cursor[1] = reinterpret_cast<long long *>(cursors) + 1;
    a.rec = reinterpret_cast<unsigned long long *>(rec);
    a.gt_per_class = gt_per_class;
    a.n = n; a.img_base = img_base; a.match_thr = match_iou_thr;
    const int ns = pick_ns(cfg.M);
    if (ns == 2 && C == 20 && B == 2) return launch_eval_state<2, 20, 2>(a, cfg, st);
    if (ns == 1) return launch_eval_state<1, 0, 0>(a, cfg, st);
    return launch_eval_state<2, 0, 0>(a, cfg, st);
}
