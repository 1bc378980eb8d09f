// yh_decode_nms.cu - float32 instantiation of the fused decode + NMS kernels (yh_decode_nms_impl.cuh), and K4:
// NMS over decoded rows, decode only, element-wise IoU - the reference's separate calls
// (utils.py:79-114, 152-218, 9-43).  sm_100a.
#include "yh_decode_nms_impl.cuh"

namespace yh {

// ------------------------------------------------------------------------------------------
// NMS over already decoded rows (n, M, 6): utils.py:79-114 as a batched call.
// ------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(256) nms_rows_kernel(const float *__restrict__ rows, int64_t n, NmsCfg cfg,
                                                       float *__restrict__ out_boxes, int *__restrict__ out_count,
                                                       int *__restrict__ out_idx)
{
    extern __shared__ uint4 smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    WarpWs<NS, true> ws(reinterpret_cast<unsigned char *>(smem_raw) + warp * cfg.ws_bytes);
    for (int i = lane; i < cfg.tbl_rows * NS; i += 32) ws.tbl[i] = 0u;
    __syncwarp();
    for (int64_t img = static_cast<int64_t>(blockIdx.x) * wpb + warp; img < n;
         img += static_cast<int64_t>(gridDim.x) * wpb) {
        const float2 *base = reinterpret_cast<const float2 *>(rows + img * cfg.M * 6);
        float conf[NS];
        float4 box[NS];
        int cls[NS];
        bool valid[NS];
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            const int i = lane + 32 * t;
            valid[t] = i < cfg.M;
            conf[t] = -INFINITY; cls[t] = 0; box[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid[t]) {
                const float2 a = base[3 * i], b = base[3 * i + 1], c = base[3 * i + 2];
                cls[t] = __float_as_int(a.x);
                conf[t] = a.y;
                box[t] = make_float4(b.x, b.y, c.x, c.y);
            }
        }
        const int K = nms_warp<NS, true>(conf, box, cls, valid, cfg, ws, out_boxes + img * cfg.M * 6,
                                         out_idx ? out_idx + img * cfg.M : nullptr);
        if (lane == 0) out_count[img] = K;
    }
}

// ------------------------------------------------------------------------------------------
// decode only (utils.py:152-218).  A block takes 256 consecutive cells: their 256 * D input floats
// are one contiguous chunk, staged in shared memory with coalesced 128-bit loads; thread = cell
// decodes from shared memory; the 256 rows (6 floats each, again one contiguous chunk) go back
// through shared memory as coalesced 128-bit stores.
// ------------------------------------------------------------------------------------------
constexpr int kDecodeCells = 256;

template <int CT, int BT>
__global__ void __launch_bounds__(kDecodeCells) decode_kernel(const float *__restrict__ pred, int64_t n_cells, NmsCfg cfg,
                                                              float *__restrict__ out)
{
    extern __shared__ float4 dsm4[];
    float *sin = reinterpret_cast<float *>(dsm4);                       // [256][D]
    float *sout = sin + kDecodeCells * cfg.D;                           // [256][6]
    const int64_t n_blk = (n_cells + kDecodeCells - 1) / kDecodeCells;
    for (int64_t blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
        const int64_t c0 = blk * kDecodeCells;
        const int cells = static_cast<int>(min(static_cast<int64_t>(kDecodeCells), n_cells - c0));
        const float *src = pred + c0 * cfg.D;
        const int nfl = cells * cfg.D;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
            for (int i = threadIdx.x; i < (nfl >> 2); i += kDecodeCells) dsm4[i] = __ldcs(s4 + i);
            for (int i = (nfl & ~3) + threadIdx.x; i < nfl; i += kDecodeCells) sin[i] = src[i];
        } else {
            for (int i = threadIdx.x; i < nfl; i += kDecodeCells) sin[i] = src[i];
        }
        __syncthreads();
        if (static_cast<int>(threadIdx.x) < cells) {
            const int cell = static_cast<int>((c0 + threadIdx.x) % cfg.M);
            int cls;
            float conf;
            float4 box;
            decode_cell<CT, BT>(sin + threadIdx.x * cfg.D, cfg, static_cast<float>(cell % cfg.S), static_cast<float>(cell / cfg.S),
                                cls, conf, box);
            float2 *o = reinterpret_cast<float2 *>(sout + threadIdx.x * 6);
            o[0] = make_float2(static_cast<float>(cls), conf);                    // utils.py:175, 213
            o[1] = make_float2(box.x, box.y);
            o[2] = make_float2(box.z, box.w);
        }
        __syncthreads();
        float *dst = out + c0 * 6;
        const int nout = cells * 6;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            const float4 *so4 = reinterpret_cast<const float4 *>(sout);
            float4 *d4 = reinterpret_cast<float4 *>(dst);
            for (int i = threadIdx.x; i < (nout >> 2); i += kDecodeCells) __stcs(d4 + i, so4[i]);
            for (int i = (nout & ~3) + threadIdx.x; i < nout; i += kDecodeCells) dst[i] = sout[i];
        } else {
            for (int i = threadIdx.x; i < nout; i += kDecodeCells) dst[i] = sout[i];
        }
        __syncthreads();                                                          // shared memory is reused
    }
}

// element-wise IoU (utils.py:9-43), (n,4) x (n,4) -> (n)
__global__ void __launch_bounds__(256) iou_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b, int64_t n,
                                                  float *__restrict__ out)
{
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = iou_ref(a[i], b[i]);
}
__global__ void __launch_bounds__(256) iou_kernel_unaligned(const float *__restrict__ a, const float *__restrict__ b,
                                                            int64_t n, float *__restrict__ out)
{
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = iou_ref(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3], b[4 * i], b[4 * i + 1], b[4 * i + 2],
                         b[4 * i + 3]);
}

int decode_nms_device(const float *pred, int64_t n, int S, int B, int C, float iou_thr, float conf_thr,
                      float *out_boxes, int32_t *out_count, int32_t *out_keep_idx, cudaStream_t st, int score_mode)
{
    return decode_nms_typed<float>(pred, n, S, B, C, iou_thr, conf_thr, out_boxes, out_count, out_keep_idx, st, score_mode);
}

template <int NS>
static int launch_nms_rows(const float *rows, int64_t n, NmsCfg cfg, float *out_boxes, int *out_count, int *out_idx,
                           cudaStream_t st)
{
    cfg.tbl_rows = 32 * NS;
    cfg.ws_bytes = WarpWs<NS, true>::bytes(cfg.tbl_rows);
    int wpb = 8;
    while (wpb > 1 && static_cast<size_t>(wpb) * cfg.ws_bytes > 200 * 1024) wpb >>= 1;
    const size_t smem = static_cast<size_t>(wpb) * cfg.ws_bytes;
    auto kern = nms_rows_kernel<NS>;
    YH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 1;
    YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t want = (n + wpb - 1) / wpb;
    const int grid = static_cast<int>(std::min<int64_t>(want, static_cast<int64_t>(sm_count()) * per_sm));
    kern<<<grid, wpb * 32, smem, st>>>(rows, n, cfg, out_boxes, out_count, out_idx);
    YH_LAUNCH_CHECK("nms_rows_kernel");
    return YH_OK;
}

}  // namespace yh

using namespace yh;

extern "C" int yh_decode_nms(const float *pred, int64_t n, int S, int B, int C, float iou_thr, float conf_thr,
                             float *out_boxes, int32_t *out_count, int32_t *out_keep_idx, void *stream)
{
    return decode_nms_device(pred, n, S, B, C, iou_thr, conf_thr, out_boxes, out_count, out_keep_idx,
                             static_cast<cudaStream_t>(stream), YH_SCORE_CONF);
}

extern "C" int yh_decode_nms_ex(const float *pred, int64_t n, int S, int B, int C, float iou_thr, float conf_thr,
                                int score_mode, float *out_boxes, int32_t *out_count, int32_t *out_keep_idx, void *stream)
{
    return decode_nms_device(pred, n, S, B, C, iou_thr, conf_thr, out_boxes, out_count, out_keep_idx,
                             static_cast<cudaStream_t>(stream), score_mode);
}

extern "C" int yh_nms(const float *boxes, int64_t n, int M, float iou_thr, float conf_thr, float *out_boxes,
                      int32_t *out_count, int32_t *out_keep_idx, void *stream)
{
    YH_REQUIRE(n >= 0 && M >= 1, "nms: bad sizes n=%lld M=%d", static_cast<long long>(n), M);
    if (M > YH_MAX_CELLS) {
        set_error("nms: M = %d exceeds the compiled limit %d", M, YH_MAX_CELLS);
        return YH_ERR_UNSUPPORTED;
    }
    if (n == 0) return YH_OK;
    YH_REQUIRE(boxes && out_boxes && out_count, "nms: null pointer");
    YH_REQUIRE(reinterpret_cast<uintptr_t>(boxes) % 8 == 0 && reinterpret_cast<uintptr_t>(out_boxes) % 8 == 0,
               "nms: boxes and out_boxes must be 8-byte aligned");
    NmsCfg cfg;
    cfg.S = 0; cfg.B = 0; cfg.C = 0; cfg.M = M; cfg.D = 6; cfg.inv_s = 0.f;
    cfg.iou_thr = iou_thr; cfg.conf_thr = conf_thr; cfg.ws_bytes = 0; cfg.tbl_rows = 0; cfg.score_mode = 0;
    set_thr(cfg);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (pick_ns(M)) {
        case 1: return launch_nms_rows<1>(boxes, n, cfg, out_boxes, out_count, out_keep_idx, st);
        case 2: return launch_nms_rows<2>(boxes, n, cfg, out_boxes, out_count, out_keep_idx, st);
        case 4: return launch_nms_rows<4>(boxes, n, cfg, out_boxes, out_count, out_keep_idx, st);
        case 7: return launch_nms_rows<7>(boxes, n, cfg, out_boxes, out_count, out_keep_idx, st);
        case 8: return launch_nms_rows<8>(boxes, n, cfg, out_boxes, out_count, out_keep_idx, st);
    }
    return YH_ERR_UNSUPPORTED;
}

extern "C" int yh_decode(const float *pred, int64_t n, int S, int B, int C, float *out_boxes, void *stream)
{
    NmsCfg cfg;
    YH_REQUIRE(S >= 1 && B >= 1 && C >= 1 && n >= 0, "decode: bad sizes");
    cfg.S = S; cfg.B = B; cfg.C = C; cfg.M = S * S; cfg.D = C + 5 * B;
    cfg.inv_s = static_cast<float>(1.0 / static_cast<double>(S));
    cfg.iou_thr = cfg.conf_thr = 0.f; cfg.ws_bytes = 0; cfg.tbl_rows = 0;
    cfg.thr_fast = 0; cfg.band = INFINITY; cfg.score_mode = 0;
    if (n == 0) return YH_OK;
    YH_REQUIRE(pred && out_boxes, "decode: null pointer");
    YH_REQUIRE(reinterpret_cast<uintptr_t>(out_boxes) % 8 == 0, "decode: out_boxes must be 8-byte aligned");
    const int64_t cells = n * cfg.M;
    size_t smem = static_cast<size_t>(kDecodeCells) * (cfg.D + 6) * 4;
    smem = (smem + 15) & ~static_cast<size_t>(15);
    if (smem > 227 * 1024) {
        set_error("decode: C + 5B = %d too large for the shared-memory tile", cfg.D);
        return YH_ERR_UNSUPPORTED;
    }
    const bool voc = (C == 20 && B == 2);
    auto kern = voc ? decode_kernel<20, 2> : decode_kernel<0, 0>;
    YH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 1;
    YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kDecodeCells, smem));
    if (per_sm < 1) per_sm = 1;
    const int grid = static_cast<int>(std::min<int64_t>((cells + kDecodeCells - 1) / kDecodeCells, static_cast<int64_t>(sm_count()) * per_sm));
    kern<<<grid, kDecodeCells, smem, static_cast<cudaStream_t>(stream)>>>(pred, cells, cfg, out_boxes);
    YH_LAUNCH_CHECK("decode_kernel");
    return YH_OK;
}

extern "C" int yh_iou(const float *boxes1, const float *boxes2, int64_t n, float *out, void *stream)
{
    YH_REQUIRE(n >= 0, "iou: n < 0");
    if (n == 0) return YH_OK;
    YH_REQUIRE(boxes1 && boxes2 && out, "iou: null pointer");
    const int grid = static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(sm_count()) * 16));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (reinterpret_cast<uintptr_t>(boxes1) % 16 == 0 && reinterpret_cast<uintptr_t>(boxes2) % 16 == 0)
        iou_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(boxes1),
                                         reinterpret_cast<const float4 *>(boxes2), n, out);
    else
        iou_kernel_unaligned<<<grid, 256, 0, st>>>(boxes1, boxes2, n, out);
    YH_LAUNCH_CHECK("iou_kernel");
    return YH_OK;
}
