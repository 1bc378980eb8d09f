// yh_adapters.cu - the callers either side of the hot path.  sm_100a.
//
//   N2  yh_encode_labels   dataset.py:88-112  YoloV1Generator._get_labels, batched over images:
//                          ragged [cx, cy, w, h, class] box lists -> (N, S, S, C+5B) label grids
//   N3  yh_head_to_f32     train.py:208 / model.py:107: the flat Dense(S*S*D) head output is the
//                          same memory as (N, S, S, D) (no kernel needed); a half-precision head
//                          (fp16 / bf16, mixed-precision Keras) is widened exactly to float32 here
//   N4  yh_pixel_boxes     utils.py:645-655 (get_tagged_img): kept rows -> int pixel corners
#include <algorithm>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "yh_common.cuh"

namespace yh {

constexpr unsigned FULLM = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// N2.  One CTA builds a tile of T consecutive label grids in shared memory (zero fill, then one
// warp per image scatters that image's boxes) and hands the finished tile to the TMA engine
// (cp.async.bulk shared -> global); two tile buffers, so the store of tile i overlaps the
// zero fill + scatter of tile i+1.  The kernel is write-bound: 4*S*S*D bytes per image.
//
// Reference semantics kept (float64 like the NumPy code, then the float32 cast of
// dataset.py:85 `batch_labels[i] = ...`):
//   loc = [grid*cy, grid*cx]; loc_i = int(loc[0]); y = loc[0] - loc_i   (x uses the column)   :101-105
//   first writer wins per cell: later boxes of an already claimed cell are dropped              :107
//   write order class -> [x, y, w, h] -> confidence, class index taken as a channel index       :108-110
//   Python indexing: negative cell / class indices wrap once; anything further out raises
//   IndexError in the reference - here the box is skipped and counted in *out_bad.
// ------------------------------------------------------------------------------------------
struct EncCfg {
    int S, C, D, M;
    int T;                  // images per tile
    int img_floats;         // M * D
    uint32_t tile_bytes;    // T * img_floats * 4
    int64_t n, n_tiles;
    int bulk;               // tiles may leave through cp.async.bulk (out 16-byte aligned, tile_bytes % 16 == 0)
};

__global__ void __launch_bounds__(256) encode_labels_kernel(const double *__restrict__ boxes,
                                                            const int64_t *__restrict__ offsets, EncCfg cfg,
                                                            float *__restrict__ out, int *__restrict__ bad)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int tile_floats = cfg.T * cfg.img_floats;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < cfg.n_tiles; tile += gridDim.x, ++it) {
        float *buf = reinterpret_cast<float *>(smem + static_cast<size_t>(it & 1) * cfg.tile_bytes);
        // the bulk store that last read this buffer (two tiles ago) must be done reading it
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        {
            float4 *b4 = reinterpret_cast<float4 *>(buf);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = threadIdx.x; i < (tile_floats >> 2); i += blockDim.x) b4[i] = z;
            for (int i = (tile_floats & ~3) + threadIdx.x; i < tile_floats; i += blockDim.x) buf[i] = 0.f;
        }
        __syncthreads();
        const int imgs = static_cast<int>(min(static_cast<int64_t>(cfg.T), cfg.n - tile * cfg.T));
        for (int il = warp; il < imgs; il += nwarp) {
            const int64_t img = tile * cfg.T + il;
            const int64_t b0 = offsets[img], b1 = offsets[img + 1];
            float *L = buf + il * cfg.img_floats;
            for (int64_t c0 = b0; c0 < b1; c0 += 32) {
                const int64_t i = c0 + lane;
                const bool valid = i < b1;
                bool ok = false;
                int cell = 0, ch = 0;
                float x = 0.f, y = 0.f, w = 0.f, h = 0.f;
                if (valid) {
                    const double *bx = boxes + 5 * i;
                    const double cx = bx[0], cy = bx[1];
                    const double li_d = __dmul_rn(static_cast<double>(cfg.S), cy);      // dataset.py:101
                    const double lj_d = __dmul_rn(static_cast<double>(cfg.S), cx);
                    const double ti = trunc(li_d), tj = trunc(lj_d), tc = trunc(bx[4]); // int(): toward zero
                    // Python index rules: [-S, S) for the cell, [-D, D) for the channel
                    if (ti >= -cfg.S && ti < cfg.S && tj >= -cfg.S && tj < cfg.S && tc >= -cfg.D && tc < cfg.D) {
                        ok = true;
                        int li = static_cast<int>(ti), lj = static_cast<int>(tj);
                        ch = static_cast<int>(tc);
                        y = static_cast<float>(__dsub_rn(li_d, ti));                   // :104
                        x = static_cast<float>(__dsub_rn(lj_d, tj));                   // :105
                        w = static_cast<float>(bx[2]);
                        h = static_cast<float>(bx[3]);
                        if (li < 0) li += cfg.S;
                        if (lj < 0) lj += cfg.S;
                        if (ch < 0) ch += cfg.D;
                        cell = li * cfg.S + lj;
                    }
                }
                // first box of a cell within this chunk = lowest lane of its match group
                const unsigned m = __match_any_sync(FULLM, ok ? cell : (0x40000000 + lane));
                float *cp = L + cell * cfg.D;
                if (ok && (__ffs(m) - 1) == lane && cp[cfg.C] == 0.0f) {                 // :107 (earlier chunks)
                    cp[ch] = 1.0f;                                                      // :108
                    cp[cfg.C + 1] = x; cp[cfg.C + 2] = y; cp[cfg.C + 3] = w; cp[cfg.C + 4] = h;   // :109
                    cp[cfg.C] = 1.0f;                                                   // :110
                }
                if (valid && !ok && bad) atomicAdd(bad, 1);
                __syncwarp();
            }
        }
        const bool bulk = cfg.bulk && imgs == cfg.T;
        if (bulk) fence_proxy_async_smem();
        __syncthreads();
        float *dst = out + tile * static_cast<int64_t>(tile_floats);
        if (bulk) {
            if (threadIdx.x == 0) bulk_s2g(dst, buf, cfg.tile_bytes);
        } else {
            const int nfl = imgs * cfg.img_floats;
            for (int i = threadIdx.x; i < nfl; i += blockDim.x) dst[i] = buf[i];
        }
    }
    if (threadIdx.x == 0) bulk_store_wait_all();
}

// ------------------------------------------------------------------------------------------
// N3.  half / bfloat16 -> float32, exact.  8 elements per thread: one 128-bit load, two 128-bit stores.
// ------------------------------------------------------------------------------------------
template <typename H>
__device__ __forceinline__ float widen(H v);
template <>
__device__ __forceinline__ float widen<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float widen<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename H>
__global__ void __launch_bounds__(256) head_to_f32_kernel(const H *__restrict__ src, int64_t n, float *__restrict__ dst,
                                                          int vec_ok)
{
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (vec_ok) {
        const int64_t n8 = n >> 3;
        const uint4 *s8 = reinterpret_cast<const uint4 *>(src);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (int64_t i = tid; i < n8; i += stride) {
            const uint4 raw = __ldcs(s8 + i);
            const H *h = reinterpret_cast<const H *>(&raw);
            __stcs(d4 + 2 * i, make_float4(widen(h[0]), widen(h[1]), widen(h[2]), widen(h[3])));
            __stcs(d4 + 2 * i + 1, make_float4(widen(h[4]), widen(h[5]), widen(h[6]), widen(h[7])));
        }
        for (int64_t i = (n8 << 3) + tid; i < n; i += stride) dst[i] = widen(src[i]);
    } else {
        for (int64_t i = tid; i < n; i += stride) dst[i] = widen(src[i]);
    }
}

// ------------------------------------------------------------------------------------------
// N4.  utils.py:652-655 on float32 tensors: xmin = int((x - (w / 2)) * width) ... ; int() truncates.
// rows (n, M, 6) [cls, conf, cx, cy, w, h]; rows at or beyond count[i] (when given) become -1.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pixel_boxes_kernel(const float *__restrict__ rows, const int *__restrict__ count,
                                                          int64_t n_rows, int M, float width, float height,
                                                          int4 *__restrict__ out)
{
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_rows;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        int4 o = make_int4(-1, -1, -1, -1);
        if (!count || static_cast<int>(i % M) < count[i / M]) {
            const float2 *r = reinterpret_cast<const float2 *>(rows + 6 * i);
            const float2 xy = r[1], wh = r[2];
            const float hw = __fmul_rn(wh.x, 0.5f), hh = __fmul_rn(wh.y, 0.5f);          // w / 2 (exact either way)
            o.x = __float2int_rz(__fmul_rn(__fsub_rn(xy.x, hw), width));
            o.y = __float2int_rz(__fmul_rn(__fsub_rn(xy.y, hh), height));
            o.z = __float2int_rz(__fmul_rn(__fadd_rn(xy.x, hw), width));
            o.w = __float2int_rz(__fmul_rn(__fadd_rn(xy.y, hh), height));
        }
        out[i] = o;
    }
}

// ------------------------------------------------------------------------------------------
// Rows with conf > thr, cell order kept (the stale evaluator of metric.py:35-37, 81 thresholds the ground
// truth but does not run NMS on it).  One warp per image, ballot/popc compaction slot by slot.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) filter_rows_kernel(const float *__restrict__ rows, int64_t n, int M, float conf_thr,
                                                          float *__restrict__ out, int *__restrict__ count)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int64_t img = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); img < n;
         img += static_cast<int64_t>(gridDim.x) * wpb) {
        const float2 *src = reinterpret_cast<const float2 *>(rows + img * M * 6);
        float2 *dst = reinterpret_cast<float2 *>(out + img * M * 6);
        int base = 0;
        for (int i0 = 0; i0 < M; i0 += 32) {
            const int i = i0 + lane;
            float2 a = make_float2(0.f, 0.f), b = a, c = a;
            if (i < M) { a = src[3 * i]; b = src[3 * i + 1]; c = src[3 * i + 2]; }
            const bool pass = i < M && a.y > conf_thr;
            const unsigned bal = __ballot_sync(FULLM, pass);
            if (pass) {
                const int pos = base + __popc(bal & ((1u << lane) - 1u));
                dst[3 * pos] = a; dst[3 * pos + 1] = b; dst[3 * pos + 2] = c;
            }
            base += __popc(bal);
        }
        if (lane == 0) count[img] = base;
    }
}

}  // namespace yh

using namespace yh;

extern "C" int yh_filter_rows(const float *rows, int64_t n, int M, float conf_thr, float *out_rows, int32_t *out_count,
                              void *stream)
{
    YH_REQUIRE(n >= 0 && M >= 1, "filter_rows: bad sizes");
    if (n == 0) return YH_OK;
    YH_REQUIRE(rows && out_rows && out_count, "filter_rows: null pointer");
    YH_REQUIRE(rows != out_rows, "filter_rows: in-place filtering is not supported");
    YH_REQUIRE(reinterpret_cast<uintptr_t>(rows) % 8 == 0 && reinterpret_cast<uintptr_t>(out_rows) % 8 == 0,
               "filter_rows: rows and out_rows must be 8-byte aligned");
    const int grid = static_cast<int>(std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(sm_count()) * 8));
    filter_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, n, M, conf_thr, out_rows, out_count);
    YH_LAUNCH_CHECK("filter_rows_kernel");
    return YH_OK;
}

extern "C" int yh_encode_labels(const double *boxes, const int64_t *offsets, int64_t n, int S, int B, int C, float *out,
                                int32_t *out_bad, void *stream)
{
    YH_REQUIRE(n >= 0 && S >= 1 && B >= 1 && C >= 1, "encode_labels: bad sizes");
    if (n == 0) return YH_OK;
    YH_REQUIRE(offsets && out, "encode_labels: null pointer");
    YH_REQUIRE(reinterpret_cast<uintptr_t>(out) % 4 == 0 && reinterpret_cast<uintptr_t>(offsets) % 8 == 0 &&
                   reinterpret_cast<uintptr_t>(boxes) % 8 == 0,
               "encode_labels: misaligned pointer");
    EncCfg cfg;
    cfg.S = S; cfg.C = C; cfg.D = C + 5 * B; cfg.M = S * S;
    cfg.img_floats = cfg.M * cfg.D;
    const int64_t img_bytes = 4ll * cfg.img_floats;
    if (2 * img_bytes + 256 > 227 * 1024) {
        set_error("encode_labels: one label grid (%lld B) does not fit the shared-memory tile", static_cast<long long>(img_bytes));
        return YH_ERR_UNSUPPORTED;
    }
    // tile: as many images as fit ~48 KB, and a multiple of 4 images when possible so tile_bytes % 16 == 0
    int T = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(48 * 1024 / img_bytes, 64)));
    if (T >= 4) T &= ~3;
    cfg.T = T;
    cfg.tile_bytes = static_cast<uint32_t>(T * img_bytes);
    cfg.n = n;
    cfg.n_tiles = (n + T - 1) / T;
    cfg.bulk = (reinterpret_cast<uintptr_t>(out) % 16 == 0 && cfg.tile_bytes % 16 == 0) ? 1 : 0;
    const size_t smem = 2 * static_cast<size_t>((cfg.tile_bytes + 127) & ~127u);
    // buffers are addressed as smem + (it & 1) * tile_bytes: keep the second one 16-byte aligned
    YH_REQUIRE(!cfg.bulk || cfg.tile_bytes % 16 == 0, "encode_labels: internal tile alignment");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (out_bad) YH_CUDA(cudaMemsetAsync(out_bad, 0, sizeof(int32_t), st));
    YH_CUDA(cudaFuncSetAttribute(encode_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 1;
    YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_labels_kernel, 256, smem));
    if (per_sm < 1) per_sm = 1;
    const int grid = static_cast<int>(std::min<int64_t>(cfg.n_tiles, static_cast<int64_t>(sm_count()) * per_sm));
    encode_labels_kernel<<<grid, 256, smem, st>>>(boxes, offsets, cfg, out, out_bad);
    YH_LAUNCH_CHECK("encode_labels_kernel");
    return YH_OK;
}

extern "C" int yh_head_to_f32(const void *src, int src_dtype, int64_t n, float *dst, void *stream)
{
    YH_REQUIRE(n >= 0, "head_to_f32: n < 0");
    YH_REQUIRE(src_dtype == YH_DTYPE_F16 || src_dtype == YH_DTYPE_BF16, "head_to_f32: src_dtype %d is neither YH_DTYPE_F16 nor YH_DTYPE_BF16", src_dtype);
    if (n == 0) return YH_OK;
    YH_REQUIRE(src && dst, "head_to_f32: null pointer");
    YH_REQUIRE(reinterpret_cast<uintptr_t>(src) % 2 == 0 && reinterpret_cast<uintptr_t>(dst) % 4 == 0, "head_to_f32: misaligned pointer");
    const int vec_ok = (reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0) ? 1 : 0;
    const int64_t work = vec_ok ? (n + 7) / 8 : n;
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, static_cast<int64_t>(sm_count()) * 16)));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (src_dtype == YH_DTYPE_F16)
        head_to_f32_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half *>(src), n, dst, vec_ok);
    else
        head_to_f32_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16 *>(src), n, dst, vec_ok);
    YH_LAUNCH_CHECK("head_to_f32_kernel");
    return YH_OK;
}

extern "C" int yh_pixel_boxes(const float *rows, const int32_t *count, int64_t n, int M, int width, int height,
                              int32_t *out, void *stream)
{
    YH_REQUIRE(n >= 0 && M >= 1, "pixel_boxes: bad sizes");
    if (n == 0) return YH_OK;
    YH_REQUIRE(rows && out, "pixel_boxes: null pointer");
    YH_REQUIRE(reinterpret_cast<uintptr_t>(rows) % 8 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
               "pixel_boxes: rows must be 8-byte and out 16-byte aligned");
    const int64_t n_rows = n * M;
    const int grid = static_cast<int>(std::min<int64_t>((n_rows + 255) / 256, static_cast<int64_t>(sm_count()) * 16));
    pixel_boxes_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, count, n_rows, M, static_cast<float>(width),
                                                                          static_cast<float>(height),
                                                                          reinterpret_cast<int4 *>(out));
    YH_LAUNCH_CHECK("pixel_boxes_kernel");
    return YH_OK;
}
