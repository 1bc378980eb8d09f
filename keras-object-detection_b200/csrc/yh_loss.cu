// yh_loss.cu - K5: YoloV1Loss forward + hand-written backward in one launch.  sm_100a.
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
#include <cstdlib>
#include <mutex>
#include <vector>

#include "yh_common.cuh"

namespace yh {

struct LossCfg {
    int B, C, D;
    float lc, ln;           // lambda_coord, lambda_noobj
    int64_t n_cells;
    int tile_cells;         // cells per shared-memory tile
    int stages;             // depth of the TMA ring (<= kLossMaxStages)
    int defer_cap;          // capacity of the deferred heavy-cell list
};

constexpr int kLossThreads = 128;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Programmatic dependent launch: launched with cudaLaunchAttributeProgrammaticStreamSerialization, a kernel's CTAs may be
// scheduled while the kernel before it in the stream is still draining; griddepcontrol.wait then blocks until that kernel
// has completed and its writes are visible, and is the first thing each thread does - before any global access - so the
// stream's ordering is unchanged and only launch latency / CTA start-up overlap the predecessor's tail.  Both instructions
// are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_wait_then_release()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

#ifdef YH_LOSS_TIMELINE
__device__ unsigned long long g_loss_tl[8][2048];
#define LOSS_STAMP(k)                                                                         \
    do {                                                                                      \
        if (threadIdx.x == 0 && blockIdx.x < 2048) {                                          \
            unsigned long long t_;                                                            \
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)::"memory");                   \
            g_loss_tl[k][blockIdx.x] = t_;                                                    \
        }                                                                                     \
    } while (0)
#else
#define LOSS_STAMP(k)
#endif

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

// Box / confidence terms of one "heavy" cell (object, or a non-zero true box) and their gradients.  t, p: the cell's
// D values of y_true / y_pred (shared or global memory); gc: the cell's gradient row (kGrad only).
template <bool kGrad>
__device__ __forceinline__ void heavy_box_terms(const float *t, const float *p, float *gc, const LossCfg &cfg, double &sxy,
                                                double &swh, double &sob, double &snb)
{
    const int C = cfg.C, B = cfg.B;
    const float obj = t[C];                                           // loss.py:162
    const float tx = t[C + 1], ty = t[C + 2], tw = t[C + 3], th = t[C + 4];
    int k = 0;                                                        // loss.py:126-137
    float u = iou_ref(tx, ty, tw, th, p[C + 1], p[C + 2], p[C + 3], p[C + 4]);
    for (int b = 1; b < B; ++b) {
        const float *q = p + C + 5 * b;
        const float v = iou_ref(tx, ty, tw, th, q[1], q[2], q[3], q[4]);
        if (v > u) { u = v; k = b; }
    }
    const float *q = p + C + 5 * k;
    const float c = q[0], px = q[1], py = q[2], pw = q[3], ph_ = q[4];
    const float noobj = __fsub_rn(1.0f, obj);
    const float z = __fsub_rn(0.0f, c);
    snb += static_cast<double>(__fmul_rn(noobj, __fmul_rn(z, z)));
    const float dx = __fsub_rn(tx, px), dy = __fsub_rn(ty, py);
    sxy += static_cast<double>(__fmul_rn(obj, __fmul_rn(dx, dx)));    // loss.py:171
    sxy += static_cast<double>(__fmul_rn(obj, __fmul_rn(dy, dy)));
    const float sw = sgn(pw), sh = sgn(ph_);
    const float rw = __fsqrt_rn(__fadd_rn(fabsf(pw), 1e-6f)), rh = __fsqrt_rn(__fadd_rn(fabsf(ph_), 1e-6f));
    const float dw = __fsub_rn(__fsqrt_rn(tw), __fmul_rn(sw, rw));    // loss.py:176-178
    const float dh = __fsub_rn(__fsqrt_rn(th), __fmul_rn(sh, rh));
    swh += static_cast<double>(__fmul_rn(obj, __fmul_rn(dw, dw)));
    swh += static_cast<double>(__fmul_rn(obj, __fmul_rn(dh, dh)));
    const float e = __fsub_rn(u, c);
    sob += static_cast<double>(__fmul_rn(obj, __fmul_rn(e, e)));      // loss.py:189
    if (kGrad) {
        float g_c = cfg.ln * 2.0f * noobj * c;
        float g_x = 0.f, g_y = 0.f, g_w = 0.f, g_h = 0.f;
        if (obj != 0.0f) {
            // IoU pieces again, with their partial derivatives
            const float x1n = (tx - tw) * 0.5f, x1x = (tx + tw) * 0.5f;
            const float y1n = (ty - th) * 0.5f, y1x = (ty + th) * 0.5f;
            const float x2n = (px - pw) * 0.5f, x2x = (px + pw) * 0.5f;
            const float y2n = (py - ph_) * 0.5f, y2x = (py + ph_) * 0.5f;
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
        }
        float *gq = gc + C + 5 * k;                                   // every other box of the cell stays 0
        gq[0] = g_c; gq[1] = g_x; gq[2] = g_y; gq[3] = g_w; gq[4] = g_h;
    }
}

// Class term of one heavy cell with an object (loss.py:206) by one warp, lanes over classes.
template <bool kGrad>
__device__ __forceinline__ void heavy_class_term(const float *t, const float *p, float *gc, int C, int lane, double &scl)
{
    const float obj = t[C];
    if (obj != 0.0f) {
        for (int j = lane; j < C; j += 32) {
            const float d = __fsub_rn(t[j], p[j]);
            scl += static_cast<double>(__fmul_rn(obj, __fmul_rn(d, d)));
            if (kGrad) gc[j] = -2.0f * obj * d;
        }
    }
}

// Deterministic batch sums: lanes -> warp -> block partials (fixed order); the last block to finish (ticket counter)
// adds the partials of all blocks in index order and writes the six outputs, so the result does not depend on which
// block came last.  Called by every thread of the block (whole warps, at most kLossMaxWarps), once.
constexpr int kLossMaxWarps = 8;
__device__ __forceinline__ void block_finish(double sxy, double swh, double sob, double snb, double scl, const LossCfg &cfg,
                                             double *__restrict__ partials, unsigned *__restrict__ ticket,
                                             float *__restrict__ out_terms)
{
    __shared__ double red[kLossMaxWarps][5];
    __shared__ double fin[5][26];
    __shared__ bool is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    // stage 1 of the deterministic reduction: lanes -> warp -> block, fixed order
    sxy = warp_sum(sxy); swh = warp_sum(swh); sob = warp_sum(sob); snb = warp_sum(snb); scl = warp_sum(scl);
    if (lane == 0) {
        red[warp][0] = sxy; red[warp][1] = swh; red[warp][2] = sob; red[warp][3] = snb; red[warp][4] = scl;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0;
        for (int w = 0; w < nwarp; ++w) s += red[w][threadIdx.x];
        partials[static_cast<size_t>(blockIdx.x) * 5 + threadIdx.x] = s;
    }
    // stage 2: the last block (ticket) sums all partials in block-index order.  Barrier, then ONE thread fences and takes
    // the ticket: its fence is cumulative over the partials the barrier ordered before it, and the other 100+ threads
    // no longer wait for their own gradient stores to drain before the block may leave
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(ticket, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // thread i: term i % 5, blocks (i / 5), (i / 5) + 25, ...; then 25 sub-sums per term in index order.
    // The loads of a thread are issued together (one L2 round trip, not one per block) and added in
    // block order afterwards.
    const int term = threadIdx.x % 5, slot = threadIdx.x / 5;
    const int nslot = min(25, static_cast<int>(blockDim.x) / 5);              // 25 for blocks of >= 125 threads
    if (slot < nslot) {
        double s = 0;
        const int nb = static_cast<int>(gridDim.x);
        for (int b0 = slot; b0 < nb; b0 += nslot * 16) {
            double v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int b = b0 + nslot * j;
                v[j] = (b < nb) ? __ldcg(partials + static_cast<size_t>(b) * 5 + term) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) s += v[j];
        }
        fin[term][slot] = s;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double tot = 0;
        for (int i = 0; i < nslot; ++i) tot += fin[threadIdx.x][i];
        fin[threadIdx.x][25] = tot;
        out_terms[threadIdx.x] = static_cast<float>(tot);
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                                   // loss.py:210-213
        out_terms[5] = static_cast<float>(static_cast<double>(cfg.lc) * (fin[0][25] + fin[1][25]) + fin[2][25] +
                                          static_cast<double>(cfg.ln) * fin[3][25] + fin[4][25]);
        *ticket = 0u;                                                         // ready for the next launch
    }
}

// Persistent CTAs; each walks its tiles (tile_cells consecutive cells of y_true and y_pred) through a
// kLossStages-deep shared-memory ring filled by cp.async.bulk (TMA, one elected thread, mbarrier
// complete_tx), so the loads of tile i+1.. overlap the arithmetic of tile i.  Tiles that are not
// 16-byte aligned/sized (tail, odd bases) are loaded with plain coalesced loads by the whole CTA.
//
// The gradient is almost all zeros (about 95 % of the cells of a VOC batch hold no object, and of a
// no-object cell only the confidence of box 0 gets a gradient), so it is not assembled element by
// element: the CTA zero-fills its gradient tile in global memory with 128-bit stores, and the few
// non-zero entries are written on top afterwards (same CTA, ordered by __syncthreads; both land in
// L2 before the line is evicted, so DRAM sees full lines once).
//   pass A  (per tile) thread per cell: cells without object and with an all-zero true box ("light")
//           only owe the no-object term on box 0 (every IoU is exactly +0 there, so the first-max
//           responsible box is box 0, loss.py:136,197).  The others ("heavy", ~5 %) are NOT worked on
//           here - their long dependent chains (divisions, square roots) would stall the whole
//           tile pipeline for a handful of busy threads: their 2*D values are copied to a deferred
//           list in shared memory
//   pass B  (when the list is full, and once at the end) one thread per deferred cell: IoUs,
//           responsible box, the four box/confidence terms and their gradients
//   pass C  class term and gradient of the deferred cells with an object: one warp per cell, lanes
//           over classes (every other cell contributes obj * (...) = 0 exactly, loss.py:206)
// Sums: per-thread float64 -> warp -> block partials; the last block to finish (ticket counter) adds
// the partials of all blocks in index order and writes the six outputs: one launch, and the result
// does not depend on which block came last.
constexpr int kLossMaxStages = 8;

template <bool kGrad>
__global__ void __launch_bounds__(kLossThreads) loss_kernel(const float *__restrict__ yt, const float *__restrict__ yp,
                                                            LossCfg cfg, float *__restrict__ grad,
                                                            double *__restrict__ partials, unsigned *__restrict__ ticket,
                                                            float *__restrict__ out_terms)
{
    extern __shared__ __align__(128) unsigned char smem[];
    pdl_wait_then_release();
    const int C = cfg.C, D = cfg.D;
    const int tile_fl = cfg.tile_cells * D;
    const uint32_t tile_bytes = static_cast<uint32_t>(tile_fl) * 4u;
    float *ring = reinterpret_cast<float *>(smem);                     // [stage][0: y_true tile | 1: y_pred tile]
    const int kLossStages = cfg.stages;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + static_cast<size_t>(kLossStages) * 2 * tile_bytes);
    int *heavy = reinterpret_cast<int *>(full + kLossStages);           // [tile_cells] heavy cells of the current tile
    int64_t *hcell = reinterpret_cast<int64_t *>(heavy + ((cfg.tile_cells + 1) & ~1));   // [defer_cap] global cell index
    float *hdat = reinterpret_cast<float *>(hcell + cfg.defer_cap);     // [defer_cap][2 D]: y_true row | y_pred row
    __shared__ int wcount[kLossThreads / 32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int64_t n_tiles = (cfg.n_cells + cfg.tile_cells - 1) / cfg.tile_cells;
    const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool base_ok = ((reinterpret_cast<uintptr_t>(yt) | reinterpret_cast<uintptr_t>(yp)) % 16 == 0) && (tile_bytes % 16 == 0);
    const bool gvec_ok = kGrad && (reinterpret_cast<uintptr_t>(grad) % 16 == 0) && (tile_bytes % 16 == 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kLossStages; ++s) mbar_init(full + s, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // tile `it` of this CTA -> global tile index / cells; bulk path only for full, aligned tiles
    auto tile_cells_of = [&](int64_t it) {
        const int64_t cell0 = (blockIdx.x + it * gridDim.x) * cfg.tile_cells;
        return static_cast<int>(min(static_cast<int64_t>(cfg.tile_cells), cfg.n_cells - cell0));
    };
    auto is_bulk = [&](int64_t it) { return base_ok && tile_cells_of(it) == cfg.tile_cells; };
    auto issue = [&](int64_t it) {       // thread 0 only
        const int s = static_cast<int>(it % kLossStages);
        const int64_t off = (blockIdx.x + it * gridDim.x) * static_cast<int64_t>(tile_fl);
        mbar_arrive_expect_tx(full + s, 2 * tile_bytes);
        const uint64_t pol = l2_evict_first_policy();
        bulk_g2s(ring + static_cast<size_t>(s) * 2 * tile_fl, yt + off, tile_bytes, full + s, pol);
        bulk_g2s(ring + static_cast<size_t>(s) * 2 * tile_fl + tile_fl, yp + off, tile_bytes, full + s, pol);
    };
    if (threadIdx.x == 0)
        for (int64_t it = 0; it < my_tiles && it < kLossStages - 1; ++it)
            if (is_bulk(it)) issue(it);

    double sxy = 0, swh = 0, sob = 0, snb = 0, scl = 0;
    int n_def = 0;                                  // deferred heavy cells (CTA-uniform)
    // pass B + C over the deferred list; called by all threads
    auto flush = [&]() {
        __syncthreads();                                                      // list complete
        for (int h = threadIdx.x; h < n_def; h += blockDim.x) {
            const float *t = hdat + h * 2 * D;
            heavy_box_terms<kGrad>(t, t + D, kGrad ? grad + hcell[h] * D : nullptr, cfg, sxy, swh, sob, snb);
        }
        for (int h = warp; h < n_def; h += nwarp) {
            const float *t = hdat + h * 2 * D;
            heavy_class_term<kGrad>(t, t + D, kGrad ? grad + hcell[h] * D : nullptr, C, lane, scl);
        }
        __syncthreads();                                                      // list may be refilled
    };
    for (int64_t it = 0; it < my_tiles; ++it) {
        const int s = static_cast<int>(it % kLossStages);
        const uint32_t ph = static_cast<uint32_t>((it / kLossStages) & 1);
        const int64_t cell0 = (blockIdx.x + it * gridDim.x) * cfg.tile_cells;
        const int cells = tile_cells_of(it);
        const int nfl = cells * D;
        const float *st = ring + static_cast<size_t>(s) * 2 * tile_fl;
        const float *sp = st + tile_fl;
        float *gg = kGrad ? grad + cell0 * D : nullptr;
        const bool bulk = is_bulk(it);
        // refill the stage that tile it-1 used (every thread left it at the barrier that ends an iteration)
        if (threadIdx.x == 0) {
            const int64_t nx = it + kLossStages - 1;
            if (nx < my_tiles && is_bulk(nx)) issue(nx);
        }
        // ---- gradient tile := 0 (overwritten below where it is not) ----
        if (kGrad) {
            if (gvec_ok && cells == cfg.tile_cells) {
                float4 *g4 = reinterpret_cast<float4 *>(gg);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int i = threadIdx.x; i < (nfl >> 2); i += blockDim.x) g4[i] = z;
            } else {
                for (int i = threadIdx.x; i < nfl; i += blockDim.x) gg[i] = 0.f;
            }
        }
        if (bulk) {
            mbar_wait(full + s, ph);
        } else {
            const float *gt = yt + cell0 * D, *gp = yp + cell0 * D;
            float *wt = ring + static_cast<size_t>(s) * 2 * tile_fl;
            for (int i = threadIdx.x; i < nfl; i += blockDim.x) {
                wt[i] = gt[i];
                wt[tile_fl + i] = gp[i];
            }
            __syncthreads();
        }

        // ---- pass A: light cells + compaction of the heavy ones ----
        int n_heavy = 0;
        for (int c0 = 0; c0 < cells; c0 += blockDim.x) {
            const int cell = c0 + threadIdx.x;
            bool hv = false;
            float g_light = 0.f;
            if (cell < cells) {
                const float *t = st + cell * D;
                const float obj = t[C];
                hv = (obj != 0.0f) || (t[C + 1] != 0.0f) || (t[C + 2] != 0.0f) || (t[C + 3] != 0.0f) || (t[C + 4] != 0.0f);
                if (!hv) {
                    const float c = sp[cell * D + C];                         // responsible = box 0
                    const float noobj = __fsub_rn(1.0f, obj);                 // loss.py:163
                    const float z = __fsub_rn(0.0f, c);
                    snb += static_cast<double>(__fmul_rn(noobj, __fmul_rn(z, z)));   // loss.py:197
                    g_light = cfg.ln * 2.0f * noobj * c;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hv);
            if (lane == 0) wcount[warp] = __popc(bal);
            __syncthreads();                                                  // also: zero fill before the stores below
            int base = n_heavy;
            for (int w = 0; w < warp; ++w) base += wcount[w];
            if (hv) heavy[base + __popc(bal & ((1u << lane) - 1u))] = cell;
            if (kGrad && cell < cells && !hv) gg[cell * D + C] = g_light;
            for (int w = 0; w < nwarp; ++w) n_heavy += wcount[w];
            __syncthreads();
        }
        // ---- defer the heavy cells of this tile (flush first if they would not fit) ----
        for (int h0 = 0; h0 < n_heavy;) {
            if (n_def == cfg.defer_cap) {
                flush();
                n_def = 0;
            }
            const int take = min(n_heavy - h0, cfg.defer_cap - n_def);
            for (int h = warp; h < take; h += nwarp) {                     // one warp per cell, lanes over its 2 D values
                const int cell = heavy[h0 + h];
                float *dst = hdat + (n_def + h) * 2 * D;
                for (int j = lane; j < D; j += 32) {
                    dst[j] = st[cell * D + j];
                    dst[D + j] = sp[cell * D + j];
                }
            }
            for (int h = threadIdx.x; h < take; h += blockDim.x) hcell[n_def + h] = cell0 + heavy[h0 + h];
            n_def += take;
            h0 += take;
        }
        __syncthreads();              // everybody is done with this stage and with heavy[]
    }
    if (n_def > 0) flush();

    block_finish(sxy, swh, sob, snb, scl, cfg, partials, ticket, out_terms);
}

// ------------------------------------------------------------------------------------------
// Gather variant (large batches): no staging.
// A "light" cell (no object, all-zero true box: ~95 % of a VOC batch) is decided by y_true[C .. C+4] and owes only the
// no-object term on the confidence of box 0, y_pred[C].  So the tile is not staged through shared memory at all:
// thread = cell, the six values come straight from global memory (read-only path, all loads of a thread in flight
// together), and only the heavy cells (~5 %) read their full rows, through L1/L2, in the same passes B and C as above.
// The gradient tile is zero-filled with 128-bit stores while the loads are in flight and the few non-zero entries are
// written on top after the barrier.  DRAM traffic is NOT reduced - ncu shows the same bytes read as the ring kernel
// (1.54 GB at batch 131,072: the memory system fetches the sectors between the 24 needed bytes of each 120-byte cell
// anyway, whatever cudaLimitMaxL2FetchGranularity says) - but the kernel issues half the instructions (no ring, no
// deferral list, one barrier pair per 256 cells) at twice the occupancy: 350 us against 376 us for the ring kernel at
// batch 131,072 (DRAM at 6.6 TB/s).  On the 4,096-image cfg3 batch the ring kernel keeps a small edge (22.0 vs 22.9 us).
// ------------------------------------------------------------------------------------------
#ifndef YH_GATHER_THREADS
#define YH_GATHER_THREADS 256
#endif
#ifndef YH_GATHER_CPT
#define YH_GATHER_CPT 2
#endif
#ifndef YH_GATHER_QUAD
#define YH_GATHER_QUAD 1
#endif
constexpr int kGatherThreads = YH_GATHER_THREADS;
constexpr int kGatherCPT = YH_GATHER_CPT;                       // cells per thread: fewer, fatter threads keep a cfg3-sized batch in ONE wave
constexpr int kGatherTile = kGatherThreads * kGatherCPT;
constexpr int kZeroFloats = 4096;                               // 16 KB of zeros for the bulk zero fill of the gradient tiles

template <bool kGrad>
__global__ void __launch_bounds__(kGatherThreads, 3) loss_gather_kernel(const float *__restrict__ yt, const float *__restrict__ yp,
                                                                     LossCfg cfg, float *__restrict__ grad,
                                                                     double *__restrict__ partials, unsigned *__restrict__ ticket,
                                                                     float *__restrict__ out_terms)
{
    __shared__ int heavy[kGatherTile];
    __shared__ int wcount[kGatherCPT][kGatherThreads / 32];
    // the gradient tile is zero-filled by the TMA engine from this block of zeros (cp.async.bulk shared -> global): threads
    // that issue 24 MB of plain stores sit in the store queue for 4 - 9 us (timeline in profiles/README.md) before they can
    // even look at the values they loaded; a bulk copy costs one thread a few instructions
    __shared__ __align__(128) float zeros[kZeroFloats];
    LOSS_STAMP(0);
    pdl_wait_then_release();
    LOSS_STAMP(1);
    const int C = cfg.C, D = cfg.D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int nwarp = kGatherThreads / 32;
    const int64_t n_tiles = (cfg.n_cells + kGatherTile - 1) / kGatherTile;
    const bool gvec_ok = kGrad && (reinterpret_cast<uintptr_t>(grad) % 16 == 0) && ((static_cast<int64_t>(kGatherTile) * D) % 4 == 0);
    const bool pair_ok = (((C | D) & 1) == 0) && (reinterpret_cast<uintptr_t>(yt) % 8 == 0);   // y_true[C..C+3] as two aligned pairs
    // Thread = two CONSECUTIVE cells (an even / odd pair starts on a 16-byte boundary when C % 4 == 0 and D is even): the five
    // deciding values of y_true come as one 16-byte and one 8-byte load per cell - LDG.128 [C..C+3] + LDG.64 [C+4..C+5] for the
    // even cell, LDG.64 [C..C+1] + LDG.128 [C+2..C+5] for the odd one - three requests per cell with y_pred[C] instead of four.
    // The gathers are bound by the SM's miss tracking (timeline in profiles/README.md), so requests are what counts.
    static_assert(kGatherCPT == 2, "the consecutive-pair mapping below is written for two cells per thread");
    const bool quad_ok = (C % 4 == 0) && (D % 2 == 0) && (D >= C + 6) /* [C+5] is read: B >= 2 */ && (reinterpret_cast<uintptr_t>(yt) % 16 == 0) && YH_GATHER_QUAD;
    double sxy = 0, swh = 0, sob = 0, snb = 0, scl = 0;
    if (kGrad && gvec_ok) {
        float4 *z4 = reinterpret_cast<float4 *>(zeros);
        for (int i = threadIdx.x; i < kZeroFloats / 4; i += kGatherThreads) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        fence_proxy_async_smem();                                             // generic-proxy writes -> visible to the TMA engine
        __syncthreads();
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t cell0 = tile * kGatherTile;
        const int cells = static_cast<int>(min(static_cast<int64_t>(kGatherTile), cfg.n_cells - cell0));
        float obj[kGatherCPT], b1[kGatherCPT], b2[kGatherCPT], b3[kGatherCPT], b4[kGatherCPT], c0[kGatherCPT];
        bool in[kGatherCPT];
#pragma unroll
        for (int j = 0; j < kGatherCPT; ++j) {                               // every load of the thread in flight together
            const int cell = quad_ok ? 2 * static_cast<int>(threadIdx.x) + j : j * kGatherThreads + static_cast<int>(threadIdx.x);
            in[j] = cell < cells;
            obj[j] = b1[j] = b2[j] = b3[j] = b4[j] = c0[j] = 0.f;
            if (in[j]) {
                const float *t = yt + (cell0 + cell) * D + C;
                if (quad_ok) {                                               // (cell0 is even: the parity of the cell is j)
                    if (j == 0) {
                        const float4 u = __ldg(reinterpret_cast<const float4 *>(t));
                        const float2 v = __ldg(reinterpret_cast<const float2 *>(t + 4));
                        obj[j] = u.x; b1[j] = u.y; b2[j] = u.z; b3[j] = u.w; b4[j] = v.x;
                    } else {
                        const float2 u = __ldg(reinterpret_cast<const float2 *>(t));
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(t + 2));
                        obj[j] = u.x; b1[j] = u.y; b2[j] = v.x; b3[j] = v.y; b4[j] = v.z;
                    }
                    c0[j] = __ldg(yp + (cell0 + cell) * D + C);
                    continue;
                }
                if (pair_ok) {
                    const float2 u = __ldg(reinterpret_cast<const float2 *>(t));
                    const float2 v = __ldg(reinterpret_cast<const float2 *>(t + 2));
                    obj[j] = u.x; b1[j] = u.y; b2[j] = v.x; b3[j] = v.y;
                } else {
                    obj[j] = __ldg(t); b1[j] = __ldg(t + 1); b2[j] = __ldg(t + 2); b3[j] = __ldg(t + 3);
                }
                b4[j] = __ldg(t + 4);
                c0[j] = __ldg(yp + (cell0 + cell) * D + C);
            }
        }
        // ---- gradient tile := 0 while the loads are in flight (overwritten below where it is not) ----
        if (kGrad) {
            float *gg = grad + cell0 * D;
            const int nfl = cells * D;
            if (gvec_ok && cells == kGatherTile) {
                if (threadIdx.x == 0) {                                       // nfl * 4 bytes in pieces of the zero block
                    const uint32_t total = static_cast<uint32_t>(nfl) * 4u;
                    for (uint32_t o = 0; o < total; o += kZeroFloats * 4u) {
                        const uint32_t nbytes = min(static_cast<uint32_t>(kZeroFloats) * 4u, total - o);
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<char *>(gg) + o),
                                     "r"(smem_u32(zeros)), "r"(nbytes)
                                     : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else {
                for (int i = threadIdx.x; i < nfl; i += kGatherThreads) gg[i] = 0.f;
            }
        }
        LOSS_STAMP(2);
        // ---- pass A: light cells, compaction of the heavy ones ----
        bool hv[kGatherCPT];
        float g_light[kGatherCPT];
        unsigned bal[kGatherCPT];
#pragma unroll
        for (int j = 0; j < kGatherCPT; ++j) {
            hv[j] = in[j] && ((obj[j] != 0.0f) || (b1[j] != 0.0f) || (b2[j] != 0.0f) || (b3[j] != 0.0f) || (b4[j] != 0.0f));
            g_light[j] = 0.f;
            if (in[j] && !hv[j]) {
                const float noobj = __fsub_rn(1.0f, obj[j]);                      // loss.py:163
                const float z = __fsub_rn(0.0f, c0[j]);                           // responsible = box 0
                snb += static_cast<double>(__fmul_rn(noobj, __fmul_rn(z, z)));    // loss.py:197
                g_light[j] = cfg.ln * 2.0f * noobj * c0[j];
            }
            bal[j] = __ballot_sync(0xffffffffu, hv[j]);
            if (lane == 0) wcount[j][warp] = __popc(bal[j]);
        }
        if (kGrad && threadIdx.x == 0) bulk_store_wait_all();                 // the zero fill has been written (long ago: it was issued before the loads came back)
        __syncthreads();                                                      // also: zero fill before the stores below
        LOSS_STAMP(3);
        int n_heavy = 0;
#pragma unroll
        for (int j = 0; j < kGatherCPT; ++j) {
            int base = n_heavy;
#pragma unroll
            for (int w = 0; w < nwarp; ++w) {
                const int k = wcount[j][w];
                base += (w < warp) ? k : 0;
                n_heavy += k;
            }
            const int cell = quad_ok ? 2 * static_cast<int>(threadIdx.x) + j : j * kGatherThreads + static_cast<int>(threadIdx.x);
            if (hv[j]) heavy[base + __popc(bal[j] & ((1u << lane) - 1u))] = cell;
            if (kGrad && in[j] && !hv[j]) grad[(cell0 + cell) * D + C] = g_light[j];
        }
        __syncthreads();                                                      // heavy[] complete
        // ---- pass C first: class term of the heavy cells with an object (loss.py:206), flattened over (cell, class) so that
        //      every thread has ALL its loads in flight at once - one L2 round trip for the CTA instead of one per cell and
        //      warp (a warp-per-cell loop cost three to four dependent round trips: 3.5 of the kernel's 18 us) - and they fly
        //      while pass B (box / confidence terms, thread per heavy cell: IoUs, square roots, divisions) computes ----
        constexpr int CI = 4;                                                 // (cell, class) items per thread and batch
        const int n_items = n_heavy * C;
        const int n_batches = max((n_items + CI * kGatherThreads - 1) / (CI * kGatherThreads), (n_heavy + kGatherThreads - 1) / kGatherThreads);
        const int sh = kGatherThreads / C, sj = kGatherThreads % C;
        int ih = threadIdx.x / C, ij = threadIdx.x % C;                       // item tid + k * 256 = (heavy cell ih, class ij)
        for (int b = 0; b < n_batches; ++b) {
            float to[CI], tj[CI], pj[CI];
            int64_t go[CI];
#pragma unroll
            for (int q = 0; q < CI; ++q) {
                go[q] = -1;
                to[q] = tj[q] = pj[q] = 0.f;
                if ((b * CI + q) * kGatherThreads + static_cast<int>(threadIdx.x) < n_items) {
                    const int64_t off = (cell0 + heavy[ih]) * D;
                    to[q] = __ldg(yt + off + C);
                    tj[q] = __ldg(yt + off + ij);
                    pj[q] = __ldg(yp + off + ij);
                    go[q] = off + ij;
                }
                ih += sh;
                ij += sj;
                if (ij >= C) { ij -= C; ++ih; }
            }
            const int h = b * kGatherThreads + static_cast<int>(threadIdx.x);   // pass B: thread per heavy cell
            if (h < n_heavy) {
                const int64_t off = (cell0 + heavy[h]) * D;
                heavy_box_terms<kGrad>(yt + off, yp + off, kGrad ? grad + off : nullptr, cfg, sxy, swh, sob, snb);
            }
#pragma unroll
            for (int q = 0; q < CI; ++q) {
                if (go[q] >= 0 && to[q] != 0.0f) {
                    const float d = __fsub_rn(tj[q], pj[q]);
                    scl += static_cast<double>(__fmul_rn(to[q], __fmul_rn(d, d)));
                    if (kGrad) grad[go[q]] = -2.0f * to[q] * d;
                }
            }
        }
        if (tile + gridDim.x < n_tiles) __syncthreads();                      // heavy[] / wcount[] are reused by the next tile
    }
    LOSS_STAMP(4);
    block_finish(sxy, swh, sob, snb, scl, cfg, partials, ticket, out_terms);
    LOSS_STAMP(5);
}

// ------------------------------------------------------------------------------------------
// Stream variant (mid-size batches, cfg3): TMA in, TMA out, nothing else touches global memory.
// What the per-CTA timeline of the gather kernel showed at cfg3 (profiles/README.md): its 4-byte gathers are bound by the
// SM's miss tracking (threads spend 4 - 9 us just ISSUING their loads), and everything that needs a second look at
// memory - the full rows of the heavy cells - queues behind the whole backlog (3 - 4 us).  Here a tile of 64 cells (both
// tensors, 2 x 7,680 B) arrives by cp.async.bulk into a 2-stage ring (as in loss_kernel), thread = cell works on it in
// shared memory - the heavy cells' rows are already there, so there is no second round trip and no deferral list - and
// the gradient tile is COMPOSED in shared memory (zeros + the few non-zero entries) and leaves as one bulk store
// (cp.async.bulk shared -> global): no zero-fill stores through the LSU, no ordering between a fill and the sparse
// writes.  Small CTAs (2 warps, 4 per SM) keep the three barriers of a tile cheap and 60 - 120 KB per SM in flight.
// Bit-identical gradients and per-cell terms to the other two kernels (same device functions).
// ------------------------------------------------------------------------------------------
constexpr int kStreamThreads = 64;                              // = cells per tile
#ifndef YH_STREAM_STAGES
#define YH_STREAM_STAGES 2
#endif
constexpr int kStreamStages = YH_STREAM_STAGES;
#ifndef YH_STREAM_GBUFS
#define YH_STREAM_GBUFS 2
#endif
constexpr int kStreamGbufs = YH_STREAM_GBUFS;              // gradient tiles in shared memory (2: a store may still be read while the next is composed)

template <bool kGrad>
__global__ void __launch_bounds__(kStreamThreads) loss_stream_kernel(const float *__restrict__ yt, const float *__restrict__ yp,
                                                                     LossCfg cfg, float *__restrict__ grad,
                                                                     double *__restrict__ partials, unsigned *__restrict__ ticket,
                                                                     float *__restrict__ out_terms)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int wcount[kStreamThreads / 32];
    pdl_wait_then_release();
    constexpr int TT = kStreamThreads;
    const int C = cfg.C, D = cfg.D;
    const int tile_fl = TT * D;
    const uint32_t tile_bytes = static_cast<uint32_t>(tile_fl) * 4u;                 // 256 * D: a multiple of 16
    float *ring = reinterpret_cast<float *>(smem);                                    // [stage][y_true tile | y_pred tile]
    float *gbuf = ring + static_cast<size_t>(kStreamStages) * 2 * tile_fl;            // [2][tile] gradient tiles (double buffered)
    uint64_t *full = reinterpret_cast<uint64_t *>(gbuf + kStreamGbufs * static_cast<size_t>(tile_fl));
    int *heavy = reinterpret_cast<int *>(full + kStreamStages);                       // [TT]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (cfg.n_cells + TT - 1) / TT;
    const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool in_ok = ((reinterpret_cast<uintptr_t>(yt) | reinterpret_cast<uintptr_t>(yp)) % 16 == 0);
    const bool out_ok = kGrad && (reinterpret_cast<uintptr_t>(grad) % 16 == 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStreamStages; ++s) mbar_init(full + s, 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto cells_of = [&](int64_t it) {
        const int64_t cell0 = (blockIdx.x + it * gridDim.x) * TT;
        return static_cast<int>(min(static_cast<int64_t>(TT), cfg.n_cells - cell0));
    };
    auto issue = [&](int64_t it) {       // thread 0 only; full, aligned tiles
        const int s = static_cast<int>(it % kStreamStages);
        const int64_t off = (blockIdx.x + it * gridDim.x) * static_cast<int64_t>(tile_fl);
        mbar_arrive_expect_tx(full + s, 2 * tile_bytes);
        const uint64_t pol = l2_evict_first_policy();
        bulk_g2s(ring + static_cast<size_t>(s) * 2 * tile_fl, yt + off, tile_bytes, full + s, pol);
        bulk_g2s(ring + static_cast<size_t>(s) * 2 * tile_fl + tile_fl, yp + off, tile_bytes, full + s, pol);
    };
    if (threadIdx.x == 0)
        for (int64_t it = 0; it < my_tiles && it < kStreamStages - 1; ++it)
            if (in_ok && cells_of(it) == TT) issue(it);
    if (kGrad) {                                                                      // both gradient tiles start as zeros
        float4 *g4 = reinterpret_cast<float4 *>(gbuf);
        for (int i = threadIdx.x; i < (kStreamGbufs * tile_fl) >> 2; i += TT) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    double sxy = 0, swh = 0, sob = 0, snb = 0, scl = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
        const int s = static_cast<int>(it % kStreamStages);
        const uint32_t ph = static_cast<uint32_t>((it / kStreamStages) & 1);
        const int64_t cell0 = (blockIdx.x + it * gridDim.x) * TT;
        const int cells = cells_of(it);
        float *st = ring + static_cast<size_t>(s) * 2 * tile_fl;
        float *sp = st + tile_fl;
        float *gb = gbuf + static_cast<size_t>(it % kStreamGbufs) * tile_fl;
        const bool bulk_in = in_ok && cells == TT;
        // the stage of tile it + stages - 1 was read by tile it-1, which every thread left at the barrier that ends an iteration
        if (threadIdx.x == 0 && it + kStreamStages - 1 < my_tiles && in_ok && cells_of(it + kStreamStages - 1) == TT) issue(it + kStreamStages - 1);
        if (kGrad && it >= kStreamGbufs) {
            // this gradient buffer left with tile it - kStreamGbufs: once the TMA engine has read it, make it zeros again
            if (threadIdx.x == 0) {
                if (kStreamGbufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncthreads();
            float4 *g4 = reinterpret_cast<float4 *>(gb);
            for (int i = threadIdx.x; i < tile_fl >> 2; i += TT) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (bulk_in) {
            mbar_wait(full + s, ph);
        } else {
            const float *gt = yt + cell0 * D, *gp = yp + cell0 * D;
            for (int i = threadIdx.x; i < cells * D; i += TT) {
                st[i] = gt[i];
                sp[i] = gp[i];
            }
            __syncthreads();
        }
        // ---- pass A: thread = cell; light cells owe the no-object term on box 0 (loss.py:136,197) ----
        const int cell = threadIdx.x;
        const bool in = cell < cells;
        bool hv = false;
        float g_light = 0.f;
        if (in) {
            const float *t = st + cell * D;
            const float obj = t[C];
            hv = (obj != 0.0f) || (t[C + 1] != 0.0f) || (t[C + 2] != 0.0f) || (t[C + 3] != 0.0f) || (t[C + 4] != 0.0f);
            if (!hv) {
                const float c = sp[cell * D + C];
                const float noobj = __fsub_rn(1.0f, obj);                             // loss.py:163
                const float z = __fsub_rn(0.0f, c);
                snb += static_cast<double>(__fmul_rn(noobj, __fmul_rn(z, z)));        // loss.py:197
                g_light = cfg.ln * 2.0f * noobj * c;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hv);
        if (lane == 0) wcount[warp] = __popc(bal);
        __syncthreads();                                                              // also: the zeros of gb are in place
        int base = 0, n_heavy = 0;
#pragma unroll
        for (int w = 0; w < TT / 32; ++w) {
            const int k = wcount[w];
            base += (w < warp) ? k : 0;
            n_heavy += k;
        }
        if (hv) heavy[base + __popc(bal & ((1u << lane) - 1u))] = cell;
        if (kGrad && in && !hv) gb[cell * D + C] = g_light;
        __syncthreads();                                                              // heavy[] complete
        // ---- heavy cells (~5 %): class term flattened over (cell, class), box / confidence terms thread per cell - all
        //      from the stage in shared memory, gradients into the tile in shared memory ----
        for (int itx = threadIdx.x; itx < n_heavy * C; itx += TT) {
            const int h = itx / C, j = itx - h * C;
            const int hc = heavy[h];
            const float obj = st[hc * D + C];
            if (obj != 0.0f) {                                                        // loss.py:206
                const float d = __fsub_rn(st[hc * D + j], sp[hc * D + j]);
                scl += static_cast<double>(__fmul_rn(obj, __fmul_rn(d, d)));
                if (kGrad) gb[hc * D + j] = -2.0f * obj * d;
            }
        }
        if (static_cast<int>(threadIdx.x) < n_heavy) {
            const int hc = heavy[threadIdx.x];
            heavy_box_terms<kGrad>(st + hc * D, sp + hc * D, kGrad ? gb + hc * D : nullptr, cfg, sxy, swh, sob, snb);
        }
        // ---- the gradient tile leaves as one bulk store ----
        const bool bulk_out = out_ok && cells == TT;
        if (kGrad && bulk_out) fence_proxy_async_smem();                              // generic-proxy writes -> visible to the TMA engine
        __syncthreads();                                                              // everybody is done with this stage, heavy[] and gb
        if (kGrad) {
            if (bulk_out) {
                if (threadIdx.x == 0) bulk_s2g(grad + cell0 * D, gb, tile_bytes);
            } else {
                float *gg = grad + cell0 * D;                                         // ragged last tile, unaligned gradient
                for (int i = threadIdx.x; i < cells * D; i += TT) gg[i] = gb[i];
            }
        }
    }
    if (kGrad && threadIdx.x == 0) bulk_store_wait_read_all();                        // shared memory must outlive the TMA reads
    block_finish(sxy, swh, sob, snb, scl, cfg, partials, ticket, out_terms);
}

// Per-device scratch of the loss: block partials + ticket.  Calls on different streams are
// ordered through an event so the scratch is never shared by two running launches.
struct LossScratch {
    double *partials = nullptr;
    unsigned *ticket = nullptr;
    int cap_blocks = 0;
    cudaEvent_t ev = nullptr;         // recorded only when the scratch changes hands between streams
    cudaStream_t last = nullptr;
    bool used = false;
};
constexpr int kLossStreams = 8;       // streams per device with a scratch block of their own (no cross-stream ordering needed)
struct LossStreamSlot {
    cudaStream_t st = nullptr;
    bool taken = false;
    LossScratch sc;
};
struct LossGeo {
    size_t smem = 0;
    int per_sm = 0;
};
static LossScratch g_loss[64];                     // shared fall-back once a device has used more than kLossStreams streams
static LossStreamSlot g_loss_slot[64][kLossStreams];
static LossGeo g_geo[6][64];
static std::mutex g_loss_mu;

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
    // tile / ring geometry: small tiles keep the static tile split even across the persistent CTAs (the
    // cfg3 batch is only ~10 tiles per CTA) and a deep ring keeps >= 44 KB per SM in flight
    static const int env_tile = [] { const char *v = getenv("YH_LOSS_TILE"); return (v && *v) ? atoi(v) : 64; }();
    static const int env_stages = [] { const char *v = getenv("YH_LOSS_STAGES"); return (v && *v) ? atoi(v) : 2; }();
    static const int env_ctas = [] { const char *v = getenv("YH_LOSS_CTAS"); return (v && *v) ? atoi(v) : 4; }();
    int tile = std::max(16, std::min(1024, env_tile));
    cfg.stages = std::max(2, std::min(kLossMaxStages, env_stages));
    static const int env_defer = [] { const char *v = getenv("YH_LOSS_DEFER"); return (v && *v) ? atoi(v) : 64; }();
    cfg.defer_cap = std::max(8, std::min(std::min(128, env_defer), (24 * 1024) / (8 * cfg.D)));
    auto smem_of = [&](int tl) {
        return static_cast<size_t>(cfg.stages) * 2 * tl * cfg.D * 4 + cfg.stages * 8 + static_cast<size_t>(tl + 2) * 4 +
               static_cast<size_t>(cfg.defer_cap) * (8 + 8 * cfg.D) + 64;
    };
    while (tile > 16 && smem_of(tile) > static_cast<size_t>(100) * 1024) tile >>= 1;
    const size_t smem_tma = smem_of(tile);
    if (smem_tma > 227 * 1024) {
        set_error("loss: C + 5B = %d too large for the shared-memory tile", cfg.D);
        return YH_ERR_UNSUPPORTED;
    }
    cfg.tile_cells = tile;
    // two kernels: the TMA ring (small batches) and the gather variant (half the instructions: 352 vs 389 us on a batch of
    // 131,072; with two cells per thread the cfg3 batch is a single wave of CTAs: 21.0 vs 21.8 us); YH_LOSS_GATHER = 0 / 1
    // forces one of them
    const char *gv = getenv("YH_LOSS_GATHER");                     // read per call: the tests switch it
    const int env_gather = (gv && *gv) ? atoi(gv) : -1;            // 0 = ring, 1 = gather, 2 = stream
    const size_t smem_stream = static_cast<size_t>(kStreamStages * 2 + kStreamGbufs) * kStreamThreads * cfg.D * 4 + kStreamStages * 8 + kStreamThreads * 4 + 64;
    // measured (B200, fwd+bwd us at batch 1,024 / 4,096 / 16,384 / 65,536): ring 11.3 / 20.5 / 57.4 / 196, gather 10.6 / 19.7 /
    // 57.7 / 188, stream 11.4 / 20.2 / 52.2 / 179 (6.45 TB/s); forward only: gather 8.2 / 14.3 / 38.7 / 119 beats both
    int which = env_gather >= 0 ? env_gather : (n_cells < (1 << 15) ? 0 : (out_grad && n_cells >= 393216) ? 2 : 1);
    if (which == 2 && smem_stream > static_cast<size_t>(100) * 1024) which = 1;
    const bool gather = which == 1, stream_k = which == 2;
    const int64_t n_tiles = gather ? (n_cells + kGatherTile - 1) / kGatherTile
                                   : stream_k ? (n_cells + kStreamThreads - 1) / kStreamThreads : (n_cells + tile - 1) / tile;
    auto kern = gather ? (out_grad ? loss_gather_kernel<true> : loss_gather_kernel<false>)
                       : stream_k ? (out_grad ? loss_stream_kernel<true> : loss_stream_kernel<false>)
                                  : (out_grad ? loss_kernel<true> : loss_kernel<false>);
    const size_t smem = gather ? 0 : stream_k ? smem_stream : smem_tma;
    const int threads = gather ? kGatherThreads : stream_k ? kStreamThreads : kLossThreads;

    int dev = 0;
    YH_CUDA(cudaGetDevice(&dev));
    YH_REQUIRE(dev >= 0 && dev < 64, "loss: device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_loss_mu);
    // launch geometry is cached: the attribute / occupancy queries cost more than the kernel
    LossGeo &g = g_geo[(out_grad ? 1 : 0) + 2 * which][dev];
    if (g.smem != smem || g.per_sm == 0) {
        if (smem) YH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        int per_sm = 1;
        YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        g.smem = smem;
        g.per_sm = per_sm < 1 ? 1 : per_sm;
    }
    int grid = static_cast<int>(
        std::max<int64_t>(1, std::min<int64_t>(n_tiles, static_cast<int64_t>(sm_count()) * ((gather || stream_k) ? g.per_sm : std::min(g.per_sm, std::max(1, env_ctas))))));
    if (stream_k) {
        // same number of tiles for every CTA (the CTAs walk their tiles in lock step: a last round that only some of them
        // take leaves the memory system half idle for a whole tile time)
        static const int env_bal = [] { const char *v = getenv("YH_STREAM_BALANCE"); return (v && *v) ? atoi(v) : 1; }();
        if (env_bal) {
            const int64_t rounds = (n_tiles + grid - 1) / grid;
            grid = static_cast<int>((n_tiles + rounds - 1) / rounds);
        }
    }
    // every stream gets its own partials / ticket (launches of one stream are ordered anyway, so nothing has to be
    // recorded or waited for between them); only past kLossStreams streams per device is one block shared through an event
    LossScratch *scp = nullptr;
    bool shared = false;
    for (int i = 0; i < kLossStreams && !scp; ++i) {
        LossStreamSlot &sl = g_loss_slot[dev][i];
        if (sl.taken && sl.st == st) scp = &sl.sc;
    }
    for (int i = 0; i < kLossStreams && !scp; ++i) {
        LossStreamSlot &sl = g_loss_slot[dev][i];
        if (!sl.taken) {
            sl.taken = true;
            sl.st = st;
            scp = &sl.sc;
        }
    }
    if (!scp) {
        scp = &g_loss[dev];
        shared = true;
    }
    LossScratch &sc = *scp;
    if (sc.cap_blocks < grid) {
        if (sc.partials) {
            YH_CUDA(cudaDeviceSynchronize());
            YH_CUDA(cudaFree(sc.partials));
            YH_CUDA(cudaFree(sc.ticket));
        }
        const int cap = std::max(grid, sm_count() * 16);
        YH_CUDA(cudaMalloc(&sc.partials, sizeof(double) * 5 * cap));
        YH_CUDA(cudaMalloc(&sc.ticket, sizeof(unsigned)));
        YH_CUDA(cudaMemset(sc.ticket, 0, sizeof(unsigned)));
        YH_CUDA(cudaDeviceSynchronize());
        sc.cap_blocks = cap;
        if (shared && !sc.ev) YH_CUDA(cudaEventCreateWithFlags(&sc.ev, cudaEventDisableTiming));
    }
    if (shared && sc.used && sc.last != st) YH_CUDA(cudaStreamWaitEvent(st, sc.ev, 0));
    static const bool pdl = [] { const char *v = getenv("YH_PDL"); return !(v && *v && atoi(v) == 0); }();
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(threads);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = pdl ? 1 : 0;
    YH_CUDA(cudaLaunchKernelEx(&lc, kern, y_true, y_pred, cfg, out_grad, sc.partials, sc.ticket, out_terms));
    YH_LAUNCH_CHECK("loss_kernel");
#ifdef YH_LOSS_TIMELINE
    if (getenv("YH_LOSS_DBG") && gather) {
        static unsigned long long h[8][2048];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_loss_tl, sizeof(h));
        const int nb = std::min(grid, 2048);
        unsigned long long t0 = ~0ull;
        for (int b = 0; b < nb; ++b) t0 = std::min(t0, h[0][b]);
        const char *nm[] = {"start", "pdl-wait", "loads+zero issued", "light done", "heavy done", "finish"};
        fprintf(stderr, "loss timeline (us after the first CTA started; min / median / max over %d CTAs):", nb);
        for (int k = 0; k < 6; ++k) {
            std::vector<double> v;
            for (int b = 0; b < nb; ++b) v.push_back((double)(h[k][b] - t0) / 1000.0);
            std::sort(v.begin(), v.end());
            fprintf(stderr, "  %s %.2f/%.2f/%.2f", nm[k], v.front(), v[v.size() / 2], v.back());
        }
        fprintf(stderr, "\n");
    }
#endif
    if (shared) YH_CUDA(cudaEventRecord(sc.ev, st));
    sc.used = true;
    sc.last = st;
    return YH_OK;
}
