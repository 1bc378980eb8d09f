// yh_abi.cu - library state (error string, launch counter) and the DLPack front ends.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "yh_common.cuh"
#include "yh_dlpack.h"

namespace yh {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
    return YH_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- DLPack validation ----------------------------------------------------------------
struct View {
    void *ptr;
    int ndim;
    int64_t shape[8];
    int64_t numel;
};

// code/bits: expected dtype.  Checks device, dtype, lanes, compact row-major strides.
static int view_of(const DLManagedTensor *m, const char *name, uint8_t code, uint8_t bits, View &v)
{
    YH_REQUIRE(m != nullptr, "%s: null DLManagedTensor", name);
    const DLTensor &t = m->dl_tensor;
    YH_REQUIRE(t.device.device_type == kDLCUDA,
               "%s: tensor is on device type %d, need kDLCUDA (2); there is no CPU fallback", name,
               static_cast<int>(t.device.device_type));
    int cur = -1;
    YH_CUDA(cudaGetDevice(&cur));
    YH_REQUIRE(t.device.device_id == cur, "%s: tensor on cuda:%d but the current device is cuda:%d", name,
               t.device.device_id, cur);
    YH_REQUIRE(t.dtype.code == code && t.dtype.bits == bits && t.dtype.lanes == 1,
               "%s: dtype (code %d, bits %d, lanes %d), expected (code %d, bits %d)", name, t.dtype.code, t.dtype.bits,
               t.dtype.lanes, code, bits);
    YH_REQUIRE(t.ndim >= 0 && t.ndim <= 8, "%s: ndim %d unsupported", name, t.ndim);
    v.ndim = t.ndim;
    v.numel = 1;
    for (int i = 0; i < t.ndim; ++i) {
        YH_REQUIRE(t.shape[i] >= 0, "%s: negative extent", name);
        v.shape[i] = t.shape[i];
        v.numel *= t.shape[i];
    }
    if (t.strides != nullptr && v.numel > 0) {
        int64_t expect = 1;
        for (int i = t.ndim - 1; i >= 0; --i) {
            YH_REQUIRE(t.shape[i] == 1 || t.strides[i] == expect, "%s: tensor is not C-contiguous", name);
            expect *= t.shape[i];
        }
    }
    v.ptr = static_cast<char *>(t.data) + t.byte_offset;
    YH_REQUIRE(v.numel == 0 || v.ptr != nullptr, "%s: null data pointer", name);
    YH_REQUIRE(reinterpret_cast<uintptr_t>(v.ptr) % (bits / 8) == 0, "%s: misaligned data pointer", name);
    return YH_OK;
}

#define YH_VIEW(var, m, name, code, bits)                  \
    View var;                                              \
    do {                                                   \
        int rc_ = view_of(m, name, code, bits, var);       \
        if (rc_ != YH_OK) return rc_;                      \
    } while (0)

}  // namespace yh

using namespace yh;

extern "C" int yh_version(void) { return 1000 * 0 + 2; }

extern "C" const char *yh_last_error(void) { return g_err; }

extern "C" int64_t yh_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int yh_device_info(int *sm, int *cc_major, int *cc_minor)
{
    int dev = 0;
    YH_CUDA(cudaGetDevice(&dev));
    int a = 0, b = 0, c = 0;
    YH_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
    YH_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
    YH_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm) *sm = a;
    if (cc_major) *cc_major = b;
    if (cc_minor) *cc_minor = c;
    return YH_OK;
}

extern "C" int yh_iou_dl(const DLManagedTensor *b1, const DLManagedTensor *b2, DLManagedTensor *out, void *stream)
{
    YH_VIEW(a, b1, "iou.boxes1", kDLFloat, 32);
    YH_VIEW(b, b2, "iou.boxes2", kDLFloat, 32);
    YH_VIEW(o, out, "iou.out", kDLFloat, 32);
    YH_REQUIRE(a.ndim >= 1 && a.shape[a.ndim - 1] == 4, "iou.boxes1: last dimension must be 4");
    YH_REQUIRE(b.ndim == a.ndim && b.numel == a.numel, "iou: boxes1 and boxes2 must have the same shape");
    for (int i = 0; i < a.ndim; ++i) YH_REQUIRE(a.shape[i] == b.shape[i], "iou: boxes1 and boxes2 must have the same shape");
    YH_REQUIRE(o.numel == a.numel / 4, "iou.out: expected %lld elements, got %lld", static_cast<long long>(a.numel / 4),
               static_cast<long long>(o.numel));
    return yh_iou(static_cast<const float *>(a.ptr), static_cast<const float *>(b.ptr), a.numel / 4,
                  static_cast<float *>(o.ptr), stream);
}

static int grid_of(const View &p, const char *name, int B, int C, int64_t &n, int &S)
{
    YH_REQUIRE(p.ndim == 4, "%s: expected a (N, S, S, C+5B) tensor, got ndim %d", name, p.ndim);
    YH_REQUIRE(p.shape[1] == p.shape[2], "%s: grid is not square (%lld x %lld)", name,
               static_cast<long long>(p.shape[1]), static_cast<long long>(p.shape[2]));
    YH_REQUIRE(p.shape[3] == C + 5 * B, "%s: last dimension %lld != C + 5B = %d", name,
               static_cast<long long>(p.shape[3]), C + 5 * B);
    n = p.shape[0];
    S = static_cast<int>(p.shape[1]);
    return YH_OK;
}

extern "C" int yh_decode_dl(const DLManagedTensor *pred, int B, int C, DLManagedTensor *out_boxes, void *stream)
{
    YH_VIEW(p, pred, "decode.pred", kDLFloat, 32);
    YH_VIEW(o, out_boxes, "decode.out_boxes", kDLFloat, 32);
    int64_t n;
    int S;
    int rc = grid_of(p, "decode.pred", B, C, n, S);
    if (rc != YH_OK) return rc;
    YH_REQUIRE(o.numel == n * S * S * 6, "decode.out_boxes: expected (N, S*S, 6)");
    return yh_decode(static_cast<const float *>(p.ptr), n, S, B, C, static_cast<float *>(o.ptr), stream);
}

extern "C" int yh_nms_dl(const DLManagedTensor *boxes, float iou_thr, float conf_thr, DLManagedTensor *out_boxes,
                         DLManagedTensor *out_count, DLManagedTensor *out_keep_idx, void *stream)
{
    YH_VIEW(b, boxes, "nms.boxes", kDLFloat, 32);
    YH_VIEW(o, out_boxes, "nms.out_boxes", kDLFloat, 32);
    YH_VIEW(c, out_count, "nms.out_count", kDLInt, 32);
    YH_REQUIRE((b.ndim == 2 || b.ndim == 3) && b.shape[b.ndim - 1] == 6, "nms.boxes: expected (M, 6) or (N, M, 6)");
    const int64_t n = b.ndim == 3 ? b.shape[0] : 1;
    const int64_t M = b.shape[b.ndim - 2];
    YH_REQUIRE(M >= 1 && M <= YH_MAX_CELLS, "nms.boxes: M = %lld outside [1, %d]", static_cast<long long>(M), YH_MAX_CELLS);
    YH_REQUIRE(o.numel == b.numel, "nms.out_boxes: must have the shape of boxes");
    YH_REQUIRE(c.numel == n, "nms.out_count: expected %lld elements", static_cast<long long>(n));
    int32_t *kidx = nullptr;
    if (out_keep_idx) {
        YH_VIEW(k, out_keep_idx, "nms.out_keep_idx", kDLInt, 32);
        YH_REQUIRE(k.numel == n * M, "nms.out_keep_idx: expected (N, M)");
        kidx = static_cast<int32_t *>(k.ptr);
    }
    return yh_nms(static_cast<const float *>(b.ptr), n, static_cast<int>(M), iou_thr, conf_thr,
                  static_cast<float *>(o.ptr), static_cast<int32_t *>(c.ptr), kidx, stream);
}

extern "C" int yh_decode_nms_dl(const DLManagedTensor *pred, int B, int C, float iou_thr, float conf_thr,
                                DLManagedTensor *out_boxes, DLManagedTensor *out_count, DLManagedTensor *out_keep_idx,
                                void *stream)
{
    YH_VIEW(p, pred, "decode_nms.pred", kDLFloat, 32);
    YH_VIEW(o, out_boxes, "decode_nms.out_boxes", kDLFloat, 32);
    YH_VIEW(c, out_count, "decode_nms.out_count", kDLInt, 32);
    int64_t n;
    int S;
    int rc = grid_of(p, "decode_nms.pred", B, C, n, S);
    if (rc != YH_OK) return rc;
    YH_REQUIRE(o.numel == n * S * S * 6, "decode_nms.out_boxes: expected (N, S*S, 6)");
    YH_REQUIRE(c.numel == n, "decode_nms.out_count: expected (N)");
    int32_t *kidx = nullptr;
    if (out_keep_idx) {
        YH_VIEW(k, out_keep_idx, "decode_nms.out_keep_idx", kDLInt, 32);
        YH_REQUIRE(k.numel == n * S * S, "decode_nms.out_keep_idx: expected (N, S*S)");
        kidx = static_cast<int32_t *>(k.ptr);
    }
    return yh_decode_nms(static_cast<const float *>(p.ptr), n, S, B, C, iou_thr, conf_thr, static_cast<float *>(o.ptr),
                         static_cast<int32_t *>(c.ptr), kidx, stream);
}

extern "C" int yh_loss_dl(const DLManagedTensor *y_true, const DLManagedTensor *y_pred, int B, int C, float lambda_coord,
                          float lambda_noobj, DLManagedTensor *out_terms, DLManagedTensor *out_grad, void *stream)
{
    YH_VIEW(t, y_true, "loss.y_true", kDLFloat, 32);
    YH_VIEW(p, y_pred, "loss.y_pred", kDLFloat, 32);
    YH_VIEW(o, out_terms, "loss.out_terms", kDLFloat, 32);
    const int D = C + 5 * B;
    YH_REQUIRE(B >= 1 && C >= 1, "loss: B and C must be >= 1");
    YH_REQUIRE(t.ndim >= 1 && t.shape[t.ndim - 1] == D, "loss.y_true: last dimension must be C + 5B = %d", D);
    YH_REQUIRE(p.ndim == t.ndim && p.numel == t.numel, "loss: y_true and y_pred must have the same shape");
    for (int i = 0; i < t.ndim; ++i) YH_REQUIRE(t.shape[i] == p.shape[i], "loss: y_true and y_pred must have the same shape");
    YH_REQUIRE(o.numel == 6, "loss.out_terms: expected 6 floats");
    float *grad = nullptr;
    if (out_grad) {
        YH_VIEW(g, out_grad, "loss.out_grad", kDLFloat, 32);
        YH_REQUIRE(g.numel == p.numel, "loss.out_grad: must have the shape of y_pred");
        grad = static_cast<float *>(g.ptr);
    }
    return yh_loss(static_cast<const float *>(t.ptr), static_cast<const float *>(p.ptr), t.numel / D, B, C, lambda_coord,
                   lambda_noobj, static_cast<float *>(o.ptr), grad, stream);
}
