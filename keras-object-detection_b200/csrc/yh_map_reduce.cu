// yh_map_reduce.cu - K7: the reduce stage of the mAP (utils.py:364-456) as ONE persistent cooperative
// kernel: stable LSD radix sort of the packed detection records by (class asc, confidence desc), cumulative
// TP / FP, float32 precision / recall points, np.trapz terms, per-class AP and their mean.  sm_100a.
//
// Everything is hand-written (no CUB): the kernel owns the whole grid (cooperative launch, <= 2 CTAs per
// SM), every CTA keeps a contiguous range of the array, and the phases are separated by grid barriers:
//   per 8-bit digit pass   histogram of the CTA's range -> hist[cta][digit] | barrier | every CTA derives its 256
//                          output cursors from that matrix (digit-major, cta-minor = stable) | tiles of 4,096
//                          records: warp-level multi-split ranks (match.any), tile-local order staged in shared
//                          memory, coalesced runs written out | barrier
//   epilogue               class starts by binary search, TP counts per class and per CTA | barrier | every CTA
//                          rescans its range: running TP count, the two precision / recall points of each TRUE
//                          positive (a false positive contributes an exact 0 to np.trapz: its recall step is 0),
//                          terms accumulated per class as 2^-56 fixed point integers (order independent, so the
//                          result is reproducible bit for bit) | barrier | AP[c], mean over all C classes
// The record count is read from device memory (the evaluator's cursors), the input may be split in up to
// kMaxSegs segments (one per rank of the exchange step, concatenated in rank order = image order), and the
// kernel can first wait for the peers' "records delivered" flags - so the stage needs no host synchronisation
// and can be captured in a CUDA graph.  The same kernel in sort-only mode orders ground-truth rows by image
// for yh_map_match on unsorted rows.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "yh_common.cuh"
#include "yh_map_internal.cuh"

namespace cg = cooperative_groups;

namespace yh {

constexpr int RS_T = 512;                 // threads per CTA
constexpr int RS_W = RS_T / 32;           // warps
constexpr int RS_IPT = 8;                 // records per thread per tile
constexpr int RS_TILE = RS_T * RS_IPT;    // 4,096 records
constexpr int RS_BINS = 256;

struct RadixSmem {
    unsigned long long stage[RS_TILE];    // tile in digit order
    uint32_t wc[RS_W][RS_BINS];           // per-warp digit counters -> exclusive warp offsets
    uint32_t tile_cnt[RS_BINS];
    uint32_t tile_start[RS_BINS];
    uint32_t dig_off[RS_BINS];            // next output position of this CTA per digit
    uint32_t hist[RS_BINS];
    uint32_t wsum[RS_W];
    long long seg_off[kMaxSegs + 1];
    unsigned cta_tp;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long load_in(const ReduceArgs &a, const RadixSmem &sm, const unsigned long long *src,
                                                      long long i)
{
    if (src) return src[i];
    int s = 0;
    while (i >= sm.seg_off[s + 1]) ++s;
    return a.in.ptr[s][i - sm.seg_off[s]];
}

// exclusive scan of v over the first 256 threads (8 warps); `total` = sum.  All RS_T threads call it.
__device__ __forceinline__ uint32_t scan256(uint32_t v, RadixSmem &sm, uint32_t &total)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
    }
    if (lane == 31 && warp < 8) sm.wsum[warp] = x;
    __syncthreads();
    uint32_t off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t s = sm.wsum[w];
        if (w < warp) off += s;
        tot += s;
    }
    total = tot;
    __syncthreads();
    return off + x - v;
}

__global__ void __launch_bounds__(RS_T, 2) map_radix_kernel(ReduceArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RadixSmem &sm = *reinterpret_cast<RadixSmem *>(smem_raw);
    unsigned long long *ap_s = reinterpret_cast<unsigned long long *>(smem_raw + sizeof(RadixSmem));   // [C+1] fixed-point AP sums
    uint32_t *tp_s = reinterpret_cast<uint32_t *>(ap_s + (a.C + 1));                                   // [C+1] TP counts
    int32_t *gt_s = reinterpret_cast<int32_t *>(tp_s + (a.C + 1));                                     // [C+1] ground truths per class
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;

    // exchange step: the records of every rank must have landed in this rank's buffers (yh_map_exchange)
    if (a.wait_flags && tid < a.wait_n) {
        const long long t0 = clock64();
        while (ld_acquire_sys(a.wait_flags + tid) != a.wait_epoch) {
            if (clock64() - t0 > a.wait_cycles) {
                if (a.err) atomicExch(a.err, YH_MAP_ERR_TIMEOUT);
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (tid == 0) {
        long long off = 0;
        for (int s = 0; s < a.in.nseg; ++s) {
            sm.seg_off[s] = off;
            long long c = a.in.cnt_dev[s] ? *a.in.cnt_dev[s] : a.in.cnt_max[s];
            if (c < 0 || c > a.in.cnt_max[s]) {            // a peer's shard outgrew the exchange region
                if (a.err) atomicExch(a.err, YH_MAP_ERR_OVERFLOW);
                c = c < 0 ? 0 : a.in.cnt_max[s];
            }
            off += c;
        }
        sm.seg_off[a.in.nseg] = off;
    }
    __syncthreads();
    const long long n = sm.seg_off[a.in.nseg];
    const long long lo = n * cta / G, hi = n * (cta + 1) / G;
    const uint32_t lt_mask = (1u << lane) - 1u;

    if (a.mode == YH_RADIX_AP && cta == 0) {              // epilogue accumulators (read after several barriers)
        for (int c = tid; c <= a.C; c += RS_T) {
            a.class_tp[c] = 0;
            if (c < a.C) a.apfix[c] = 0;
        }
    }

    const unsigned long long *src = nullptr;              // pass 0 reads the segments
    unsigned long long *dst = a.buf[0];
    for (int p = 0; p < a.npass; ++p) {
        const int shift = a.bit_lo + 8 * p;
        // ---- histogram of this CTA's range
        if (tid < RS_BINS) sm.hist[tid] = 0;
        __syncthreads();
        for (long long i0 = lo; i0 < hi; i0 += RS_T) {
            const long long i = i0 + tid;
            const bool valid = i < hi;
            const uint32_t d = valid ? static_cast<uint32_t>(load_in(a, sm, src, i) >> shift) & 0xffu : 256u + lane;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (valid && lane == __ffs(peers) - 1) atomicAdd(&sm.hist[d], __popc(peers));
        }
        __syncthreads();
        if (tid < RS_BINS) a.hist[static_cast<size_t>(cta) * RS_BINS + tid] = sm.hist[tid];
        grid.sync();
        // ---- output cursors of this CTA: digits ascending, within a digit CTAs ascending (stable)
        // (all 512 threads: digit = tid & 255, rows of parity tid >> 8, eight loads in flight; one L2 round trip per
        // eight rows instead of one per row - at a few thousand records per CTA this loop IS the pass)
        uint32_t tot = 0, mine = 0;
        {
            const int dgt = tid & (RS_BINS - 1), par = tid >> 8;
            for (int c0 = par; c0 < G; c0 += 16) {
                uint32_t v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + 2 * j;
                    v[j] = (c < G) ? __ldcg(a.hist + static_cast<size_t>(c) * RS_BINS + dgt) : 0u;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (c0 + 2 * j < cta) mine += v[j];
                    tot += v[j];
                }
            }
            if (par == 1) { sm.tile_cnt[dgt] = tot; sm.tile_start[dgt] = mine; }      // scratch until the tile loop
            __syncthreads();
            if (par == 0) { tot += sm.tile_cnt[dgt]; mine += sm.tile_start[dgt]; }
            __syncthreads();
        }
        uint32_t all;
        const uint32_t base = scan256(tid < RS_BINS ? tot : 0u, sm, all);
        if (tid < RS_BINS) sm.dig_off[tid] = base + mine;
        __syncthreads();
        // ---- tiles
        for (long long t0 = lo; t0 < hi; t0 += RS_TILE) {
            unsigned long long k[RS_IPT];
            uint32_t d[RS_IPT], r[RS_IPT];
            const long long wbase = t0 + static_cast<long long>(warp) * (32 * RS_IPT);
#pragma unroll
            for (int j = 0; j < RS_IPT; ++j) {                   // all loads of the thread in flight together
                const long long pos = wbase + j * 32 + lane;
                const bool valid = pos < hi;
                k[j] = valid ? load_in(a, sm, src, pos) : ~0ull;
                d[j] = valid ? static_cast<uint32_t>(k[j] >> shift) & 0xffu : 256u;
            }
#pragma unroll
            for (int j = 0; j < RS_BINS / 32; ++j) sm.wc[warp][j * 32 + lane] = 0;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < RS_IPT; ++j) {
                if (wbase + j * 32 < hi) {                       // warp-uniform: a short last tile costs what it holds
                    const uint32_t peers = __match_any_sync(0xffffffffu, d[j]);
                    const bool valid = d[j] < 256u;
                    const uint32_t before = valid ? sm.wc[warp][d[j]] : 0u;
                    r[j] = before + __popc(peers & lt_mask);
                    __syncwarp();
                    if (valid && lane == __ffs(peers) - 1) sm.wc[warp][d[j]] = before + __popc(peers);
                    __syncwarp();
                }
            }
            __syncthreads();
            uint32_t cnt = 0;
            if (tid < RS_BINS) {
#pragma unroll
                for (int w = 0; w < RS_W; ++w) {
                    const uint32_t t = sm.wc[w][tid];
                    sm.wc[w][tid] = cnt;
                    cnt += t;
                }
                sm.tile_cnt[tid] = cnt;
            }
            uint32_t tile_n;
            const uint32_t ts = scan256(tid < RS_BINS ? cnt : 0u, sm, tile_n);
            if (tid < RS_BINS) sm.tile_start[tid] = ts;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < RS_IPT; ++j)
                if (d[j] < 256u) sm.stage[sm.tile_start[d[j]] + sm.wc[warp][d[j]] + r[j]] = k[j];
            __syncthreads();
            for (uint32_t q = tid; q < tile_n; q += RS_T) {
                const unsigned long long key = sm.stage[q];
                const uint32_t dd = static_cast<uint32_t>(key >> shift) & 0xffu;
                dst[sm.dig_off[dd] + (q - sm.tile_start[dd])] = key;
            }
            __syncthreads();
            if (tid < RS_BINS) sm.dig_off[tid] += sm.tile_cnt[tid];
            __syncthreads();
        }
        grid.sync();
        src = dst;
        dst = (dst == a.buf[0]) ? a.buf[1] : a.buf[0];
    }
    if (a.mode != YH_RADIX_AP) {
        if (a.out_n && cta == 0 && tid == 0) *a.out_n = n;
        return;
    }

    // ---- epilogue 1: class starts, TP counts per class (all CTAs) and per CTA
    const int C = a.C;
    for (int c = tid; c <= C; c += RS_T) {
        ap_s[c] = 0;
        tp_s[c] = 0;
        int g = 0;
        if (c < C)
            for (int s = 0; s < a.n_gt; ++s) g += a.gt_part[s][c];            // the shards' counts, summed here
        gt_s[c] = g;
    }
    if (tid == 0) sm.cta_tp = 0;
    __syncthreads();
    if (cta == 0) {
        for (int c = tid; c <= C + 1; c += RS_T) {
            long long l = 0, h = n;
            if (c > C) l = n;
            const unsigned long long key = static_cast<unsigned long long>(c) << kRecClassShift;
            while (l < h) {
                const long long mid = (l + h) >> 1;
                if (src[mid] < key) l = mid + 1; else h = mid;
            }
            a.class_start[c] = l;
        }
    }
    for (long long i0 = lo; i0 < hi; i0 += RS_T) {
        const long long i = i0 + tid;
        const bool valid = i < hi;
        const unsigned long long key = valid ? src[i] : 0ull;
        const bool tp = valid && (key & 1ull);
        const uint32_t c = tp ? static_cast<uint32_t>(key >> kRecClassShift) : 0x10000u + lane;
        const uint32_t peers = __match_any_sync(0xffffffffu, c);
        if (tp && lane == __ffs(peers) - 1) {
            atomicAdd(&tp_s[c], __popc(peers));
            atomicAdd(&sm.cta_tp, __popc(peers));
        }
    }
    __syncthreads();
    for (int c = tid; c <= C; c += RS_T)
        if (tp_s[c]) atomicAdd(&a.class_tp[c], tp_s[c]);
    if (tid == 0) a.cta_tp[cta] = sm.cta_tp;
    grid.sync();

    // ---- epilogue 2: class_base[c] = TPs of the classes before c (tp_s reused); TPs of the CTAs before this one
    {
        uint32_t run = 0;
        for (int c0 = 0; c0 <= C; c0 += RS_BINS) {                      // chunks of 256 classes
            const int c = c0 + tid;
            const uint32_t v = (tid < RS_BINS && c <= C) ? a.class_tp[c] : 0u;
            uint32_t chunk;
            const uint32_t ex = scan256(v, sm, chunk);
            if (tid < RS_BINS && c <= C) tp_s[c] = run + ex;
            run += chunk;
        }
    }
    uint32_t cta_base = 0;
    for (int c = 0; c < cta; ++c) cta_base += a.cta_tp[c];
    __syncthreads();
    uint32_t running = cta_base;
    for (long long i0 = lo; i0 < hi; i0 += RS_T) {
        const long long i = i0 + tid;
        const bool valid = i < hi;
        const unsigned long long key = valid ? src[i] : 0ull;
        const bool tp = valid && (key & 1ull);
        const uint32_t ball = __ballot_sync(0xffffffffu, tp);
        if (lane == 0) sm.wsum[warp] = __popc(ball);
        __syncthreads();
        uint32_t woff = 0, blk = 0;
#pragma unroll
        for (int w = 0; w < RS_W; ++w) {
            const uint32_t s = sm.wsum[w];
            if (w < warp) woff += s;
            blk += s;
        }
        unsigned long long fx = 0;
        uint32_t c = 0x10000u + lane;
        if (tp) {
            const uint32_t cls = static_cast<uint32_t>(key >> kRecClassShift);
            const int ngt = cls < static_cast<uint32_t>(C) ? gt_s[cls] : 0;
            if (ngt > 0) {                                                 // utils.py:334-336
                const uint32_t cum = running + woff + __popc(ball & lt_mask) + 1u;   // TPs up to and including i
                const int tpi = static_cast<int>(cum - tp_s[cls]);
                const int pos = static_cast<int>(i - a.class_start[cls]) + 1;        // detections of the class so far
                const float total = static_cast<float>(ngt);
                const float tpc = static_cast<float>(tpi), fpc = static_cast<float>(pos - tpi);
                const float r1 = __fdiv_rn(tpc, __fadd_rn(total, 1e-6f));            // utils.py:434
                const float p1 = __fdiv_rn(tpc, __fadd_rn(__fadd_rn(tpc, fpc), 1e-6f));   // utils.py:435
                float r0 = 0.0f, p0 = 1.0f;                                         // utils.py:438-439
                if (pos > 1) {
                    const float tq = static_cast<float>(tpi - 1);                   // the point before: one TP fewer, same FPs
                    r0 = __fdiv_rn(tq, __fadd_rn(total, 1e-6f));
                    p0 = __fdiv_rn(tq, __fadd_rn(__fadd_rn(tq, fpc), 1e-6f));
                }
                const float term = __fmul_rn(__fmul_rn(__fsub_rn(r1, r0), __fadd_rn(p1, p0)), 0.5f);   // np.trapz, utils.py:444
                fx = static_cast<unsigned long long>(static_cast<double>(term) * 72057594037927936.0);  // 2^56
                c = cls;
            }
        }
        // warp-aggregated add of the fixed-point terms of equal classes
        const uint32_t peers = __match_any_sync(0xffffffffu, c);
        if (__any_sync(0xffffffffu, c < 0x10000u)) {
            unsigned long long sum = 0;
            uint32_t rest = peers;
            // peers differ per class group: every lane sums its own group by stepping through the group's lanes
            while (rest) {
                const int src_lane = __ffs(rest) - 1;
                rest &= rest - 1;
                sum += __shfl_sync(peers, fx, src_lane);
            }
            if (c < 0x10000u && lane == __ffs(peers) - 1) atomicAdd(&ap_s[c], sum);
        }
        running += blk;
        __syncthreads();
    }
    __syncthreads();
    for (int c = tid; c < C; c += RS_T)
        if (ap_s[c]) atomicAdd(&a.apfix[c], ap_s[c]);
    grid.sync();

    // ---- epilogue 3: AP per class and the mean over ALL classes (utils.py:456)
    if (cta == 0) {
        double *red = reinterpret_cast<double *>(sm.stage);
        double acc = 0.0;
        for (int c = tid; c < C; c += RS_T) {
            const float ap = gt_s[c] > 0 ? static_cast<float>(static_cast<double>(a.apfix[c]) * (1.0 / 72057594037927936.0)) : 0.0f;
            if (a.out_ap) a.out_ap[c] = ap;
            acc += static_cast<double>(ap);
        }
        red[tid] = acc;
        __syncthreads();
        for (int o = RS_T / 2; o > 0; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        if (tid == 0) {
            float m = static_cast<float>(red[0] / static_cast<double>(C));
            if (a.err && *reinterpret_cast<volatile int32_t *>(a.err) != 0) m = __int_as_float(0x7fc00000);   // failed exchange: NaN, not a number that looks right
            *a.out_map = m;
            if (a.out_n) *a.out_n = n;
        }
    }
}

static int radix_grid(int64_t n_want, size_t smem, int &G)
{
    // occupancy per (device, shared-memory size) is looked up once: the attribute and occupancy calls cost microseconds
    static std::mutex mu;
    static size_t smem_set[64] = {0};
    static int occ_for[64] = {0};
    int dev = 0;
    YH_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    int occ;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > smem_set[dev] || occ_for[dev] == 0) {
            YH_CUDA(cudaFuncSetAttribute(map_radix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            int o = 0;
            YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, map_radix_kernel, RS_T, smem));
            YH_REQUIRE(o >= 1, "map_reduce: the sort kernel does not fit an SM with %zu bytes of shared memory", smem);
            smem_set[dev] = smem;
            occ_for[dev] = o;
        }
        occ = occ_for[dev];
    }
    const int cap = std::min(occ, 2) * sm_count();
    const int64_t want = (n_want + RS_TILE - 1) / RS_TILE;
    G = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(cap, want)));
    if (const char *v = getenv("YH_MAP_GRID")) {                 // experiments: any grid is correct
        const int g = atoi(v);
        if (g >= 1) G = std::min(cap, g);
    }
    return YH_OK;
}

size_t radix_ws_bytes(int64_t n_max, int C)
{
    // two record buffers, hist matrix for the largest grid, epilogue tables
    const size_t G = 2 * 160;
    return align_up(static_cast<size_t>(n_max) * 8, 256) * 2 + align_up(G * RS_BINS * 4, 256) + align_up(G * 4, 256) +
           align_up(static_cast<size_t>(C + 2) * (4 + 8 + 8), 256) + 1024;
}

int radix_launch(ReduceArgs &a, int64_t n_max, void *workspace, size_t ws_bytes, cudaStream_t st)
{
    YH_REQUIRE(n_max >= 0 && n_max < (1ll << 31), "map_reduce: bad record count");
    YH_REQUIRE(workspace != nullptr && ws_bytes >= radix_ws_bytes(n_max, a.C), "map_reduce: workspace of %zu bytes needed",
               radix_ws_bytes(n_max, a.C));
    YH_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "map_reduce: workspace must be 256-byte aligned");
    const size_t smem = sizeof(RadixSmem) + (a.mode == YH_RADIX_AP ? static_cast<size_t>(a.C + 1) * 16 + 16 : 0);
    int G = 1;
    int rc = radix_grid(a.n_hint > 0 ? a.n_hint : n_max, smem, G);
    if (rc != YH_OK) return rc;
    unsigned char *w = static_cast<unsigned char *>(workspace);
    a.buf[0] = reinterpret_cast<unsigned long long *>(w); w += align_up(static_cast<size_t>(n_max) * 8, 256);
    a.buf[1] = reinterpret_cast<unsigned long long *>(w); w += align_up(static_cast<size_t>(n_max) * 8, 256);
    a.hist = reinterpret_cast<uint32_t *>(w); w += align_up(static_cast<size_t>(2 * 160) * RS_BINS * 4, 256);
    a.cta_tp = reinterpret_cast<uint32_t *>(w); w += align_up(static_cast<size_t>(2 * 160) * 4, 256);
    a.class_tp = reinterpret_cast<uint32_t *>(w); w += align_up(static_cast<size_t>(a.C + 2) * 4, 8);
    a.class_start = reinterpret_cast<long long *>(w); w += static_cast<size_t>(a.C + 2) * 8;
    a.apfix = reinterpret_cast<unsigned long long *>(w);
    YH_REQUIRE(G <= 2 * 160, "map_reduce: grid of %d CTAs exceeds the workspace layout", G);
    void *args[] = {&a};
    YH_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(map_radix_kernel), dim3(G), dim3(RS_T), args, smem, st));
    YH_LAUNCH_CHECK("map_radix_kernel");
    return YH_OK;
}

}  // namespace yh

using namespace yh;

namespace yh {
int reduce_impl(int nseg, const uint64_t *const *rec, const int64_t *nrec_max, const int64_t *const *nrec_dev, int n_gt,
                const int32_t *const *gt_parts, int C, float *out_ap, float *out_map, const uint64_t *wait_flags, int wait_n,
                uint64_t wait_epoch, int32_t *err, int64_t n_hint, void *workspace, size_t ws_bytes, void *stream)
{
    YH_REQUIRE(C >= 1 && C <= kMaxMapClasses, "map_reduce: C = %d outside [1, %d]", C, kMaxMapClasses);
    YH_REQUIRE(nseg >= 1 && nseg <= kMaxSegs && nrec_max, "map_reduce: 1..%d record segments", kMaxSegs);
    YH_REQUIRE(gt_parts && n_gt >= 1 && n_gt <= kMaxSegs && out_map, "map_reduce: null pointer");
    ReduceArgs a{};
    a.n_gt = n_gt;
    for (int s = 0; s < n_gt; ++s) {
        YH_REQUIRE(gt_parts[s] != nullptr, "map_reduce: null ground-truth counts");
        a.gt_part[s] = gt_parts[s];
    }
    int64_t n_max = 0;
    a.in.nseg = nseg;
    for (int s = 0; s < nseg; ++s) {
        YH_REQUIRE(nrec_max[s] >= 0 && (nrec_max[s] == 0 || (rec && rec[s])), "map_reduce: bad segment %d", s);
        a.in.ptr[s] = rec ? reinterpret_cast<const unsigned long long *>(rec[s]) : nullptr;
        a.in.cnt_max[s] = nrec_max[s];
        a.in.cnt_dev[s] = nrec_dev ? reinterpret_cast<const long long *>(nrec_dev[s]) : nullptr;
        n_max += nrec_max[s];
    }
    a.bit_lo = 1;
    a.npass = (32 + class_bits(C) + 7) / 8;
    a.mode = YH_RADIX_AP;
    a.C = C;
    a.n_hint = (n_hint > 0 && n_hint < n_max) ? n_hint : n_max;
    a.out_ap = out_ap;
    a.out_map = out_map;
    a.wait_flags = reinterpret_cast<const unsigned long long *>(wait_flags);
    a.wait_n = wait_flags ? wait_n : 0;
    a.wait_epoch = wait_epoch;
    a.wait_cycles = 4000000000ll;          // ~2 s at 1.9 GHz: a peer that never delivers is an error, not a hang
    a.err = err;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AsyncBuf own(st);
    if (!workspace) {
        ws_bytes = radix_ws_bytes(n_max, C);
        YH_CUDA(own.alloc(ws_bytes));
        workspace = own.p;
    }
    return radix_launch(a, n_max, workspace, ws_bytes, st);
}
}  // namespace yh

extern "C" int yh_map_reduce(const uint64_t *rec, int64_t nrec, const int64_t *nrec_dev, int64_t n_hint,
                             const int32_t *gt_per_class, int C, float *out_ap, float *out_map, void *workspace,
                             size_t workspace_bytes, void *stream)
{
    const uint64_t *recs[1] = {rec};
    const int64_t *devs[1] = {nrec_dev};
    const int32_t *gts[1] = {gt_per_class};
    YH_REQUIRE(gt_per_class != nullptr, "map_reduce: null ground-truth counts");
    return reduce_impl(1, recs, &nrec, devs, 1, gts, C, out_ap, out_map, nullptr, 0, 0, nullptr, n_hint, workspace, workspace_bytes,
                       stream);
}
