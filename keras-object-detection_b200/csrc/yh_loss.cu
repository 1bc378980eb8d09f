// yh_loss.cu - K5: YoloV1Loss forward + hand-written backward in one pass.  sm_100a.
//
// Replaces loss.py:120-215 (YoloV1Loss.call) and the TF autodiff backward of it:
//   IoU(true box, each pred box) loss.py:126-133, first-max responsible box :136, the five
//   batch-sum terms :171-207, total = 5*box + obj + 0.5*noobj + cls :210-213.
// Element-wise terms are float32 with the reference's op order; the batch sums are carried
// in float64 through a fixed two-stage tree (no float atomics), so the result is
// run-to-run reproducible and within 1e-5 relative of any float32 summation order.
// Backward: gradient flows through the IoU target (no stop_gradient,
// loss.py:189); TF sub-gradient conventions: clip passes on [0,1] inclusive, max/min send
// ties to their first argument (the true box), sign(0) = 0.
#include <algorithm>

#include "yh_common.cuh"

namespace yh {

struct LossCfg {
    int B, C, D;
    float lc, ln;           // lambda_coord, lambda_noobj
    int64_t n_cells;
    int tile_cells;         // == blockDim.x
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

// one tile = tile_cells consecutive cells; y_true / y_pred tiles staged in shared memory with
// coalesced 128-bit loads, one thread per cell, gradient written back through the same tile.
template <bool kGrad>
__global__ void __launch_bounds__(256) loss_kernel(const float *__restrict__ yt, const float *__restrict__ yp, LossCfg cfg,
                                                   float *__restrict__ grad, double *__restrict__ partials)
{
    extern __shared__ float4 smem4[];
    float *st = reinterpret_cast<float *>(smem4);
    float *sp = st + static_cast<size_t>(cfg.tile_cells) * cfg.D;
    __shared__ double red[8][5];

    const int C = cfg.C, B = cfg.B, D = cfg.D;
    const int64_t n_tiles = (cfg.n_cells + cfg.tile_cells - 1) / cfg.tile_cells;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(yt) | reinterpret_cast<uintptr_t>(yp) |
                          (kGrad ? reinterpret_cast<uintptr_t>(grad) : 0)) % 16 == 0) &&
                        ((static_cast<int64_t>(cfg.tile_cells) * D) % 4 == 0);
    double sxy = 0, swh = 0, sob = 0, snb = 0, scl = 0;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t cell0 = tile * cfg.tile_cells;
        const int cells = static_cast<int>(min(static_cast<int64_t>(cfg.tile_cells), cfg.n_cells - cell0));
        const int nfl = cells * D;
        const float *gt = yt + cell0 * D;
        const float *gp = yp + cell0 * D;
        if (vec_ok) {
            const int n4 = nfl >> 2;
            const float4 *gt4 = reinterpret_cast<const float4 *>(gt);
            const float4 *gp4 = reinterpret_cast<const float4 *>(gp);
            float4 *st4 = reinterpret_cast<float4 *>(st);
            float4 *sp4 = reinterpret_cast<float4 *>(sp);
            for (int i = threadIdx.x; i < n4; i += blockDim.x) {
                st4[i] = __ldcs(gt4 + i);
                sp4[i] = __ldcs(gp4 + i);
            }
            for (int i = (n4 << 2) + threadIdx.x; i < nfl; i += blockDim.x) {
                st[i] = gt[i];
                sp[i] = gp[i];
            }
        } else {
            for (int i = threadIdx.x; i < nfl; i += blockDim.x) {
                st[i] = gt[i];
                sp[i] = gp[i];
            }
        }
        __syncthreads();

        if (static_cast<int>(threadIdx.x) < cells) {
            const float *t = st + threadIdx.x * D;
            float *p = sp + threadIdx.x * D;
            const float obj = t[C];                                           // loss.py:162
            const float tx = t[C + 1], ty = t[C + 2], tw = t[C + 3], th = t[C + 4];
            // responsible box: first max of IoU(true, pred_b)               // loss.py:126-137
            int k = 0;
            float u = iou_ref(tx, ty, tw, th, p[C + 1], p[C + 2], p[C + 3], p[C + 4]);
            for (int b = 1; b < B; ++b) {
                const float *q = p + C + 5 * b;
                const float v = iou_ref(tx, ty, tw, th, q[1], q[2], q[3], q[4]);
                if (v > u) { u = v; k = b; }
            }
            float *q = p + C + 5 * k;
            const float c = q[0], px = q[1], py = q[2], pw = q[3], ph = q[4];
            const float noobj = __fsub_rn(1.0f, obj);                         // loss.py:163
            // ---- forward terms (float32 element-wise, as the reference) ----
            const float dx = __fsub_rn(tx, px), dy = __fsub_rn(ty, py);
            sxy += static_cast<double>(__fmul_rn(obj, __fmul_rn(dx, dx)));    // loss.py:171
            sxy += static_cast<double>(__fmul_rn(obj, __fmul_rn(dy, dy)));
            const float sw = sgn(pw), sh = sgn(ph);
            const float rw = __fsqrt_rn(__fadd_rn(fabsf(pw), 1e-6f)), rh = __fsqrt_rn(__fadd_rn(fabsf(ph), 1e-6f));
            const float dw = __fsub_rn(__fsqrt_rn(tw), __fmul_rn(sw, rw));    // loss.py:176-178
            const float dh = __fsub_rn(__fsqrt_rn(th), __fmul_rn(sh, rh));
            swh += static_cast<double>(__fmul_rn(obj, __fmul_rn(dw, dw)));
            swh += static_cast<double>(__fmul_rn(obj, __fmul_rn(dh, dh)));
            const float e = __fsub_rn(u, c);
            sob += static_cast<double>(__fmul_rn(obj, __fmul_rn(e, e)));      // loss.py:189
            const float z = __fsub_rn(0.0f, c);
            snb += static_cast<double>(__fmul_rn(noobj, __fmul_rn(z, z)));    // loss.py:197
            if (obj != 0.0f) {
                for (int j = 0; j < C; ++j) {                                  // loss.py:206
                    const float d = __fsub_rn(t[j], p[j]);
                    scl += static_cast<double>(__fmul_rn(obj, __fmul_rn(d, d)));
                }
            }
            // ---- backward w.r.t. y_pred ----
            if (kGrad) {
                float g_c = cfg.ln * 2.0f * noobj * c;
                float g_x = 0.f, g_y = 0.f, g_w = 0.f, g_h = 0.f;
                if (obj != 0.0f) {
                    for (int j = 0; j < C; ++j) p[j] = -2.0f * obj * (t[j] - p[j]);
                    // IoU pieces again, with their partial derivatives
                    const float x1n = (tx - tw) * 0.5f, x1x = (tx + tw) * 0.5f;
                    const float y1n = (ty - th) * 0.5f, y1x = (ty + th) * 0.5f;
                    const float x2n = (px - pw) * 0.5f, x2x = (px + pw) * 0.5f;
                    const float y2n = (py - ph) * 0.5f, y2x = (py + ph) * 0.5f;
                    const float ddx = fminf(x1x, x2x) - fmaxf(x1n, x2n);
                    const float ddy = fminf(y1x, y2x) - fmaxf(y1n, y2n);
                    const float cw = clip01(ddx), ch = clip01(ddy);
                    const float inter = cw * ch;
                    const float a1 = fabsf((x1x - x1n) * (y1x - y1n));
                    const float w2 = x2x - x2n, h2 = y2x - y2n;
                    const float a2 = fabsf(w2 * h2);
                    const float dn = ((a1 + a2) - inter) + 1e-6f;
                    const float inv = 1.0f / dn;
                    const float du_dI = inv + inter * inv * inv;
                    const float du_da2 = -inter * inv * inv;
                    const float in_x = (ddx >= 0.f && ddx <= 1.f) ? 1.f : 0.f;
                    const float in_y = (ddy >= 0.f && ddy <= 1.f) ? 1.f : 0.f;
                    const float mx = (x2x < x1x) ? 1.f : 0.f, nx = (x2n > x1n) ? 1.f : 0.f;
                    const float my = (y2x < y1x) ? 1.f : 0.f, ny = (y2n > y1n) ? 1.f : 0.f;
                    const float sa = sgn(w2 * h2);
                    const float du_dpx = du_dI * in_x * ch * 0.5f * (mx - nx);
                    const float du_dpy = du_dI * in_y * cw * 0.5f * (my - ny);
                    const float du_dpw = du_dI * in_x * ch * 0.5f * (mx + nx) + du_da2 * sa * h2;
                    const float du_dph = du_dI * in_y * cw * 0.5f * (my + ny) + du_da2 * sa * w2;
                    const float e2 = 2.0f * obj * e;
                    g_c -= e2;
                    g_x = -2.0f * cfg.lc * obj * dx + e2 * du_dpx;
                    g_y = -2.0f * cfg.lc * obj * dy + e2 * du_dpy;
                    g_w = -2.0f * cfg.lc * obj * dw * (sw * sw) / (2.0f * rw) + e2 * du_dpw;
                    g_h = -2.0f * cfg.lc * obj * dh * (sh * sh) / (2.0f * rh) + e2 * du_dph;
                } else {
                    for (int j = 0; j < C; ++j) p[j] = 0.f;
                }
                for (int j = C; j < D; ++j) p[j] = 0.f;
                q[0] = g_c; q[1] = g_x; q[2] = g_y; q[3] = g_w; q[4] = g_h;
            }
        }
        __syncthreads();
        if (kGrad) {
            float *gg = grad + cell0 * D;
            if (vec_ok) {
                const int n4 = nfl >> 2;
                float4 *gg4 = reinterpret_cast<float4 *>(gg);
                const float4 *sp4 = reinterpret_cast<const float4 *>(sp);
                for (int i = threadIdx.x; i < n4; i += blockDim.x) __stcs(gg4 + i, sp4[i]);
                for (int i = (n4 << 2) + threadIdx.x; i < nfl; i += blockDim.x) gg[i] = sp[i];
            } else {
                for (int i = threadIdx.x; i < nfl; i += blockDim.x) gg[i] = sp[i];
            }
            __syncthreads();
        }
    }

    // stage 1 of the deterministic reduction: lanes -> warp -> block, fixed order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    sxy = warp_sum(sxy); swh = warp_sum(swh); sob = warp_sum(sob); snb = warp_sum(snb); scl = warp_sum(scl);
    if (lane == 0) {
        red[warp][0] = sxy; red[warp][1] = swh; red[warp][2] = sob; red[warp][3] = snb; red[warp][4] = scl;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0;
        const int nw = blockDim.x >> 5;
        for (int w = 0; w < nw; ++w) s += red[w][threadIdx.x];
        partials[static_cast<size_t>(blockIdx.x) * 5 + threadIdx.x] = s;
    }
}

// stage 2: one warp, term `lane` (< 5) summed over blocks in index order
__global__ void loss_finalize_kernel(const double *__restrict__ partials, int n_blocks, float lc, float ln,
                                     float *__restrict__ out_terms)
{
    __shared__ double tot[5];
    if (threadIdx.x < 5) {
        double s = 0;
        for (int b = 0; b < n_blocks; ++b) s += partials[static_cast<size_t>(b) * 5 + threadIdx.x];
        tot[threadIdx.x] = s;
        out_terms[threadIdx.x] = static_cast<float>(s);
    }
    __syncwarp();
    if (threadIdx.x == 0)                                                       // loss.py:210-213
        out_terms[5] = static_cast<float>(static_cast<double>(lc) * (tot[0] + tot[1]) + tot[2] +
                                          static_cast<double>(ln) * tot[3] + tot[4]);
}

}  // namespace yh

using namespace yh;

extern "C" int yh_loss(const float *y_true, const float *y_pred, int64_t n_cells, int B, int C, float lambda_coord,
                       float lambda_noobj, float *out_terms, float *out_grad, void *stream)
{
    YH_REQUIRE(B >= 1 && C >= 1 && n_cells >= 0, "loss: bad sizes");
    YH_REQUIRE(out_terms != nullptr, "loss: out_terms is null");
    YH_REQUIRE(n_cells == 0 || (y_true && y_pred), "loss: null input");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LossCfg cfg;
    cfg.B = B; cfg.C = C; cfg.D = C + 5 * B; cfg.lc = lambda_coord; cfg.ln = lambda_noobj; cfg.n_cells = n_cells;
    int tile = 256;
    while (tile > 32 && static_cast<size_t>(tile) * cfg.D * 8 > 96 * 1024) tile >>= 1;
    const size_t smem = static_cast<size_t>(tile) * cfg.D * 8;
    if (smem > 227 * 1024) {
        set_error("loss: C + 5B = %d too large for the shared-memory tile", cfg.D);
        return YH_ERR_UNSUPPORTED;
    }
    cfg.tile_cells = tile;
    const int64_t n_tiles = (n_cells + tile - 1) / tile;
    auto kern = out_grad ? loss_kernel<true> : loss_kernel<false>;
    YH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 1;
    YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, tile, smem));
    if (per_sm < 1) per_sm = 1;
    const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_tiles, static_cast<int64_t>(sm_count()) * per_sm)));
    double *partials = nullptr;
    YH_CUDA(cudaMallocAsync(&partials, sizeof(double) * 5 * grid, st));
    kern<<<grid, tile, smem, st>>>(y_true, y_pred, cfg, out_grad, partials);
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { cudaFreeAsync(partials, st); return cuda_fail(e, "loss_kernel"); }
        count_launch();
    }
    loss_finalize_kernel<<<1, 32, 0, st>>>(partials, grid, lambda_coord, lambda_noobj, out_terms);
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { cudaFreeAsync(partials, st); return cuda_fail(e, "loss_finalize_kernel"); }
        count_launch();
    }
    YH_CUDA(cudaFreeAsync(partials, st));
    return YH_OK;
}
