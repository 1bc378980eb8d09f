// yh_host.cu - host-buffer front end of the fused decode+NMS path: the call a reference user
// makes with NumPy arrays (utils.py:588-620 `MeanAveragePrecisionNumpy.update_state`,
// evaluate.py:40) and the call bench.py times end to end.
//
// The batch is cut into chunks; chunk c uses slot c % kSlots = {stream, device in/out
// buffers}.  H2D copy, kernel and D2H copies of one chunk are ordered on its stream; the
// slots' streams overlap, so the copy engines (one per direction) and the SMs all stay busy.
// That overlap needs PINNED host buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory):
// with pageable memory the call is still correct, but the driver stages every copy through its
// own pinned bounce buffer and the copies run one after the other.
// State is per device (slots, streams, one lock per device): calls for different devices run
// concurrently, calls for one device are serialised.  Rows past count[i] of an image come back
// as zeros (the device slot is cleared before each kernel), like the np.zeros the caller made.
#include <algorithm>
#include <mutex>

#include "yh_common.cuh"

namespace yh {

int decode_nms_device(const float *pred, int64_t n, int S, int B, int C, float iou_thr, float conf_thr,
                      float *out_boxes, int32_t *out_count, int32_t *out_keep_idx, cudaStream_t st, int score_mode);

constexpr int kSlots = 3;

struct HostCtx {
    bool init = false;
    cudaStream_t st[kSlots] = {};
    float *d_in[kSlots] = {};
    float *d_boxes[kSlots] = {};
    int32_t *d_count[kSlots] = {};
    int32_t *d_idx[kSlots] = {};
    float *d_rows[kSlots] = {};          // compact rows [img, cls, conf, cx, cy, w, h] of a chunk (yh_decode_nms_host_rows)
    int64_t *d_cursor[kSlots] = {};      // their count, on the device ...
    int64_t *h_cursor = nullptr;         // ... and in pinned host memory, one per slot
    size_t cap_in = 0, cap_boxes = 0, cap_count = 0, cap_idx = 0, cap_rows = 0;   // bytes per slot
};

static HostCtx g_ctx[64];
static std::mutex g_mu[64];               // one per device: a single-process multi-GPU host does not serialise its devices

static int grow(void **p, size_t &cap, size_t need, bool &changed)
{
    (void)changed;
    if (need <= cap && *p) return YH_OK;
    if (*p) YH_CUDA(cudaFree(*p));
    *p = nullptr;
    YH_CUDA(cudaMalloc(p, need));
    return YH_OK;
}

}  // namespace yh

using namespace yh;

static int host_impl(const void *pred_host_v, int dtype, int64_t n, int S, int B, int C, float iou_thr, float conf_thr,
                     float *out_boxes_host, int32_t *out_count_host, int32_t *out_keep_idx_host, int device)
{
    YH_REQUIRE(dtype == YH_DTYPE_F32 || dtype == YH_DTYPE_F16 || dtype == YH_DTYPE_BF16, "decode_nms_host: unknown dtype %d", dtype);
    const int64_t esize = dtype == YH_DTYPE_F32 ? 4 : 2;
    const unsigned char *pred_host = static_cast<const unsigned char *>(pred_host_v);
    YH_REQUIRE(n >= 0 && S >= 1 && B >= 1 && C >= 1, "decode_nms_host: bad sizes");
    YH_REQUIRE(device >= 0 && device < 64, "decode_nms_host: bad device %d", device);
    if (n == 0) return YH_OK;
    YH_REQUIRE(pred_host && out_boxes_host && out_count_host, "decode_nms_host: null pointer");
    std::lock_guard<std::mutex> lock(g_mu[device]);
    int prev = 0;
    YH_CUDA(cudaGetDevice(&prev));
    YH_CUDA(cudaSetDevice(device));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};

    const int64_t M = static_cast<int64_t>(S) * S, D = C + 5 * B;
    const int64_t img_in = esize * M * D, img_boxes = 4 * M * 6;
    // chunk: ~96 MiB of input, a multiple of 16 images so every chunk keeps the TMA alignment
    int64_t chunk = std::max<int64_t>(16, ((96ll << 20) / img_in) & ~15ll);
    chunk = std::min<int64_t>(chunk, (n + 15) & ~15ll);

    HostCtx &cx = g_ctx[device];
    if (!cx.init) {
        for (int s = 0; s < kSlots; ++s) YH_CUDA(cudaStreamCreateWithFlags(&cx.st[s], cudaStreamNonBlocking));
        cx.init = true;
    }
    bool ch = false;
    for (int s = 0; s < kSlots; ++s) {
        int rc;
        if ((rc = grow(reinterpret_cast<void **>(&cx.d_in[s]), cx.cap_in, chunk * img_in, ch)) != YH_OK) return rc;
        if ((rc = grow(reinterpret_cast<void **>(&cx.d_boxes[s]), cx.cap_boxes, chunk * img_boxes, ch)) != YH_OK) return rc;
        if ((rc = grow(reinterpret_cast<void **>(&cx.d_count[s]), cx.cap_count, chunk * 4, ch)) != YH_OK) return rc;
        if (out_keep_idx_host)
            if ((rc = grow(reinterpret_cast<void **>(&cx.d_idx[s]), cx.cap_idx, chunk * M * 4, ch)) != YH_OK) return rc;
    }
    cx.cap_in = std::max(cx.cap_in, static_cast<size_t>(chunk * img_in));
    cx.cap_boxes = std::max(cx.cap_boxes, static_cast<size_t>(chunk * img_boxes));
    cx.cap_count = std::max(cx.cap_count, static_cast<size_t>(chunk * 4));
    if (out_keep_idx_host) cx.cap_idx = std::max(cx.cap_idx, static_cast<size_t>(chunk * M * 4));

    int rc = YH_OK;
    int64_t c = 0;
    // any failure leaves the loop through `rc`: the slots' streams are ALWAYS synchronised below, so no copy into
    // the caller's buffers and no kernel on the cached slots is still in flight when the call returns
    auto ck = [&](cudaError_t e, const char *what) { if (e != cudaSuccess && rc == YH_OK) rc = cuda_fail(e, what); return e == cudaSuccess; };
    for (int64_t lo = 0; lo < n && rc == YH_OK; lo += chunk, ++c) {
        const int s = static_cast<int>(c % kSlots);
        const int64_t cnt = std::min(chunk, n - lo);
        cudaStream_t st = cx.st[s];
        if (!ck(cudaMemcpyAsync(cx.d_in[s], pred_host + lo * img_in, cnt * img_in, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync H2D")) break;
        if (!ck(cudaMemsetAsync(cx.d_boxes[s], 0, cnt * img_boxes, st), "cudaMemsetAsync")) break;      // rows past count[i] read as zeros
        if (out_keep_idx_host && !ck(cudaMemsetAsync(cx.d_idx[s], 0xff, cnt * M * 4, st), "cudaMemsetAsync")) break;   // and indices as -1
        if (dtype == YH_DTYPE_F32)
            rc = decode_nms_device(cx.d_in[s], cnt, S, B, C, iou_thr, conf_thr, cx.d_boxes[s], cx.d_count[s],
                                   out_keep_idx_host ? cx.d_idx[s] : nullptr, st, YH_SCORE_CONF);
        else        // float16 / bfloat16 head: half the bytes over PCIe and HBM, widened exactly inside the kernel
            rc = yh_decode_nms_typed(cx.d_in[s], dtype, cnt, S, B, C, iou_thr, conf_thr, YH_SCORE_CONF, cx.d_boxes[s],
                                     cx.d_count[s], out_keep_idx_host ? cx.d_idx[s] : nullptr, st);
        if (rc != YH_OK) break;
        if (!ck(cudaMemcpyAsync(out_boxes_host + lo * M * 6, cx.d_boxes[s], cnt * img_boxes, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync D2H")) break;
        if (!ck(cudaMemcpyAsync(out_count_host + lo, cx.d_count[s], cnt * 4, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync D2H")) break;
        if (out_keep_idx_host &&
            !ck(cudaMemcpyAsync(out_keep_idx_host + lo * M, cx.d_idx[s], cnt * M * 4, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync D2H")) break;
    }
    for (int s = 0; s < kSlots; ++s) {
        cudaError_t e = cudaStreamSynchronize(cx.st[s]);
        if (e != cudaSuccess && rc == YH_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    return rc;
}

extern "C" int yh_decode_nms_host(const float *pred_host, int64_t n, int S, int B, int C, float iou_thr, float conf_thr,
                                  float *out_boxes_host, int32_t *out_count_host, int32_t *out_keep_idx_host, int device)
{
    return host_impl(pred_host, YH_DTYPE_F32, n, S, B, C, iou_thr, conf_thr, out_boxes_host, out_count_host, out_keep_idx_host, device);
}

extern "C" int yh_decode_nms_host_typed(const void *pred_host, int dtype, int64_t n, int S, int B, int C, float iou_thr,
                                        float conf_thr, float *out_boxes_host, int32_t *out_count_host,
                                        int32_t *out_keep_idx_host, int device)
{
    return host_impl(pred_host, dtype, n, S, B, C, iou_thr, conf_thr, out_boxes_host, out_count_host, out_keep_idx_host, device);
}

// Compact form: only the kept rows cross PCIe on the way back.  Per chunk the padded NMS output stays on the device and
// yh_rows_append compacts it into rows [img, cls, conf, cx, cy, w, h] (img = index in the batch, as float32 like the
// reference's accumulator, utils.py:476); the chunk's row count comes back first (8 bytes), and once it is known - two
// chunks later, so the pipeline never drains - exactly rows * 28 bytes are copied behind the rows of the chunks before.
// D2H bytes: 28 * kept rows + 4 * n instead of 24 * S^2 * n (VOC-like outputs: ~3 kept of 49 slots per image).
extern "C" int yh_decode_nms_host_rows(const void *pred_host_v, int dtype, int64_t n, int S, int B, int C, float iou_thr,
                                       float conf_thr, float *out_rows_host, int64_t row_capacity, int32_t *out_count_host,
                                       int64_t *out_total_rows, int device)
{
    YH_REQUIRE(dtype == YH_DTYPE_F32 || dtype == YH_DTYPE_F16 || dtype == YH_DTYPE_BF16, "decode_nms_host_rows: unknown dtype %d", dtype);
    const int64_t esize = dtype == YH_DTYPE_F32 ? 4 : 2;
    const unsigned char *pred_host = static_cast<const unsigned char *>(pred_host_v);
    YH_REQUIRE(n >= 0 && S >= 1 && B >= 1 && C >= 1 && row_capacity >= 0, "decode_nms_host_rows: bad sizes");
    YH_REQUIRE(device >= 0 && device < 64, "decode_nms_host_rows: bad device %d", device);
    YH_REQUIRE(out_total_rows != nullptr, "decode_nms_host_rows: out_total_rows is null");
    *out_total_rows = 0;
    if (n == 0) return YH_OK;
    YH_REQUIRE(pred_host && out_count_host && (row_capacity == 0 || out_rows_host), "decode_nms_host_rows: null pointer");
    std::lock_guard<std::mutex> lock(g_mu[device]);
    int prev = 0;
    YH_CUDA(cudaGetDevice(&prev));
    YH_CUDA(cudaSetDevice(device));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};

    const int64_t M = static_cast<int64_t>(S) * S, D = C + 5 * B;
    const int64_t img_in = esize * M * D, img_boxes = 4 * M * 6, img_rows = 4 * M * 7;
    int64_t chunk = std::max<int64_t>(16, ((96ll << 20) / img_in) & ~15ll);
    chunk = std::min<int64_t>(chunk, (n + 15) & ~15ll);
    HostCtx &cx = g_ctx[device];
    if (!cx.init) {
        for (int s = 0; s < kSlots; ++s) YH_CUDA(cudaStreamCreateWithFlags(&cx.st[s], cudaStreamNonBlocking));
        cx.init = true;
    }
    if (!cx.h_cursor) YH_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&cx.h_cursor), sizeof(int64_t) * kSlots, cudaHostAllocPortable));
    bool ch = false;
    for (int s = 0; s < kSlots; ++s) {
        int rc;
        if ((rc = grow(reinterpret_cast<void **>(&cx.d_in[s]), cx.cap_in, chunk * img_in, ch)) != YH_OK) return rc;
        if ((rc = grow(reinterpret_cast<void **>(&cx.d_boxes[s]), cx.cap_boxes, chunk * img_boxes, ch)) != YH_OK) return rc;
        if ((rc = grow(reinterpret_cast<void **>(&cx.d_count[s]), cx.cap_count, chunk * 4, ch)) != YH_OK) return rc;
        if ((rc = grow(reinterpret_cast<void **>(&cx.d_rows[s]), cx.cap_rows, chunk * img_rows, ch)) != YH_OK) return rc;
        if (!cx.d_cursor[s]) YH_CUDA(cudaMalloc(reinterpret_cast<void **>(&cx.d_cursor[s]), sizeof(int64_t)));
    }
    cx.cap_in = std::max(cx.cap_in, static_cast<size_t>(chunk * img_in));
    cx.cap_boxes = std::max(cx.cap_boxes, static_cast<size_t>(chunk * img_boxes));
    cx.cap_count = std::max(cx.cap_count, static_cast<size_t>(chunk * 4));
    cx.cap_rows = std::max(cx.cap_rows, static_cast<size_t>(chunk * img_rows));

    int rc = YH_OK;
    int64_t total = 0, need = 0;
    auto ck = [&](cudaError_t e, const char *what) { if (e != cudaSuccess && rc == YH_OK) rc = cuda_fail(e, what); return e == cudaSuccess; };
    // chunk k is complete on its stream: its row count is known, copy its rows behind those of the chunks before it
    auto finish = [&](int64_t k) {
        const int s = static_cast<int>(k % kSlots);
        if (!ck(cudaStreamSynchronize(cx.st[s]), "cudaStreamSynchronize")) return;
        const int64_t rows = cx.h_cursor[s];
        need += rows;
        const int64_t fit = std::max<int64_t>(0, std::min(rows, row_capacity - total));
        if (fit > 0) ck(cudaMemcpyAsync(out_rows_host + total * 7, cx.d_rows[s], fit * 28, cudaMemcpyDeviceToHost, cx.st[s]), "cudaMemcpyAsync D2H rows");
        total += fit;
    };
    int64_t c = 0;
    const int64_t n_chunks = (n + chunk - 1) / chunk;
    for (int64_t lo = 0; lo < n && rc == YH_OK; lo += chunk, ++c) {
        const int s = static_cast<int>(c % kSlots);
        const int64_t cnt = std::min(chunk, n - lo);
        cudaStream_t st = cx.st[s];
        if (!ck(cudaMemcpyAsync(cx.d_in[s], pred_host + lo * img_in, cnt * img_in, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync H2D")) break;
        if (!ck(cudaMemsetAsync(cx.d_cursor[s], 0, sizeof(int64_t), st), "cudaMemsetAsync")) break;
        if (dtype == YH_DTYPE_F32)
            rc = decode_nms_device(cx.d_in[s], cnt, S, B, C, iou_thr, conf_thr, cx.d_boxes[s], cx.d_count[s], nullptr, st, YH_SCORE_CONF);
        else
            rc = yh_decode_nms_typed(cx.d_in[s], dtype, cnt, S, B, C, iou_thr, conf_thr, YH_SCORE_CONF, cx.d_boxes[s], cx.d_count[s], nullptr, st);
        if (rc != YH_OK) break;
        rc = yh_rows_append(cx.d_boxes[s], cx.d_count[s], cnt, static_cast<int>(M), lo, cx.d_rows[s], cnt * M, cx.d_cursor[s], st);
        if (rc != YH_OK) break;
        if (!ck(cudaMemcpyAsync(out_count_host + lo, cx.d_count[s], cnt * 4, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync D2H")) break;
        if (!ck(cudaMemcpyAsync(cx.h_cursor + s, cx.d_cursor[s], sizeof(int64_t), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync D2H")) break;
        if (c >= kSlots - 1) finish(c - (kSlots - 1));
    }
    if (rc == YH_OK)
        for (int64_t k = std::max<int64_t>(0, n_chunks - (kSlots - 1)); k < n_chunks && rc == YH_OK; ++k) finish(k);
    for (int s = 0; s < kSlots; ++s) {
        cudaError_t e = cudaStreamSynchronize(cx.st[s]);
        if (e != cudaSuccess && rc == YH_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    *out_total_rows = need;
    if (rc == YH_OK && need > row_capacity) {
        set_error("decode_nms_host_rows: %lld kept rows do not fit row_capacity %lld (the first %lld were copied)",
                  static_cast<long long>(need), static_cast<long long>(row_capacity), static_cast<long long>(total));
        return YH_ERR_ARG;
    }
    return rc;
}

// Pinned host buffers for the callers of the *_host entry points (what the chunk pipeline above needs to overlap its
// copies).  write_combined = 1 asks for cudaHostAllocWriteCombined: not snooped on its way over PCIe (faster H2D on
// some hosts) but very slow to READ from the CPU - for input staging buffers that the CPU only writes.
extern "C" int yh_host_alloc(size_t bytes, int write_combined, void **ptr)
{
    YH_REQUIRE(ptr != nullptr && bytes > 0, "host_alloc: bad arguments");
    *ptr = nullptr;
    YH_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
    return YH_OK;
}

extern "C" int yh_host_free(void *ptr)
{
    if (ptr) YH_CUDA(cudaFreeHost(ptr));
    return YH_OK;
}
