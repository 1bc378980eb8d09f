// yh_map_internal.cuh - declarations shared by the mAP translation units (yh_map.cu, yh_map_reduce.cu, yh_comm.cu).
#pragma once
#include "yh_common.cuh"

namespace yh {

constexpr int YH_RADIX_SORT_ONLY = 0;
constexpr int YH_RADIX_AP = 1;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct AsyncBuf {   // stream-ordered scratch, freed on scope exit (callers that pass no workspace)
    cudaStream_t st;
    void *p = nullptr;
    explicit AsyncBuf(cudaStream_t s) : st(s) {}
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 16, st); }
    ~AsyncBuf() { if (p) cudaFreeAsync(p, st); }
    template <class T> T *as() { return static_cast<T *>(p); }
};

// the stream-ordered allocator trims its pool at every synchronisation unless a release threshold is set
int keep_pool();

struct SegIn {                                  // the records of all ranks, rank order = image order
    int nseg;
    const unsigned long long *ptr[kMaxSegs];
    const long long *cnt_dev[kMaxSegs];         // device counters (nullable -> cnt_max is the count)
    long long cnt_max[kMaxSegs];
};

struct ReduceArgs {
    SegIn in;
    unsigned long long *buf[2];
    uint32_t *hist;                             // [grid][256]
    int bit_lo, npass, mode;
    // epilogue (YH_RADIX_AP)
    const int32_t *gt_part[kMaxSegs];           // per-class ground-truth counts of the shards (summed in the kernel)
    int n_gt;
    int C;
    long long n_hint;                           // expected record count (sizes the grid; any value is correct)
    uint32_t *cta_tp;                           // [grid]
    uint32_t *class_tp;                         // [C+1]
    long long *class_start;                     // [C+2]
    unsigned long long *apfix;                  // [C]
    float *out_ap, *out_map;
    long long *out_n;                           // nullable: number of records sorted
    // exchange step: wait until wait_flags[0..wait_n) == wait_epoch
    const unsigned long long *wait_flags;
    int wait_n;
    unsigned long long wait_epoch;
    long long wait_cycles;
    int32_t *err;                               // nullable device int: YH_MAP_ERR_* of the last failure
    // counting path (YH_RADIX_AP): AP from ranks counted inside (class, confidence bucket) groups instead of a sort
    int count_nb_log2;                          // log2 of the confidence buckets per class, -1 = path disabled
    long long count_n_max;                      // taken when the DEVICE-side record count is <= this ...
    unsigned long long count_pairs_max;         // ... and the pair count known after the histogram is <= this + count_pairs_per_rec * n
    unsigned long long count_pairs_per_rec;
    unsigned long long *dbg;                    // -DYH_MAP_TIMELINE: phase time stamps of CTA 0 (else null)
};

// ---- single-pass chained scan state --------------------------------------------------------------------------
constexpr int kScanMaxTiles = 1024;
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagIncl = 2ull << 62, kValMask = (1ull << 62) - 1;
struct ScanWs {                       // all zero at rest: the last CTA of a launch cleans up after the others
    unsigned ticket, done, pad0, pad1;
    unsigned long long st[2][kScanMaxTiles];   // [set][tile]: flag << 62 | value;  set 0 = predictions, 1 = ground truth
};

#ifdef __CUDACC__
// exclusive prefix of tile b >= 1 (sum of the totals of the tiles before it), one full warp, 32 tiles per step
__device__ __forceinline__ long long lookback(volatile unsigned long long *st, int b, int lane)
{
    long long prefix = 0;
    int hi = b - 1;
    while (true) {
        const int j = hi - lane;
        unsigned long long w = kFlagIncl;                       // before tile 0: inclusive prefix 0
        if (j >= 0) {
            do { w = st[j]; } while ((w >> 62) == 0);             // tiles with a lower ticket are running or done
        }
        const uint32_t incl = __ballot_sync(0xffffffffu, (w >> 62) == 2);
        const int first = incl ? __ffs(incl) - 1 : 31;            // nearest tile with an inclusive prefix
        long long v = (lane <= first) ? static_cast<long long>(w & kValMask) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        prefix += v;
        if (incl) return prefix;
        hi -= 32;
    }
}
#endif  // __CUDACC__

#ifdef __CUDACC__
// Matching of ONE image by one warp (utils.py:373-422): pb / tb = the image's detections / ground truths as NMS emitted
// them (rows of 6 floats [cls, conf, cx, cy, w, h]; detections in the reference's processing order: confidence
// descending, stable), in global or shared memory.  Best same-class IoU (strict >, first wins, utils.py:386-393), TP iff
// that IoU is > thr and the ground truth is still free (utils.py:395-418): lanes = detections, ground truths broadcast by
// shuffle, claims resolved with match.any inside the warp.  claimed: YH_MAX_CELLS / 32 words of the warp's shared memory;
// hist: the CTA's per-class ground-truth histogram (shared memory); emit(d, record) receives the packed record of
// detection d.  Called by all 32 lanes.
template <class Emit>
__device__ __forceinline__ void match_image(const float *pb, const float *tb, int np_i, int nt_i, int C, float iou_thr,
                                            uint32_t *claimed, int *hist, int lane, Emit emit)
{
    if (lane < YH_MAX_CELLS / 32) claimed[lane] = 0;
    for (int g = lane; g < nt_i; g += 32) {
        const uint32_t gc = class_of(tb[g * 6], C);
        if (gc < static_cast<uint32_t>(C)) atomicAdd(&hist[gc], 1);            // utils.py:330 rows of the class
    }
    __syncwarp();
    for (int d0 = 0; d0 < np_i; d0 += 32) {
        const int d = d0 + lane;
        const bool valid = d < np_i;
        float conf = 0.0f, dx = 0.0f, dy = 0.0f, dw = 0.0f, dh = 0.0f;
        uint32_t dc = static_cast<uint32_t>(C);
        if (valid) {
            const float2 v0 = *reinterpret_cast<const float2 *>(pb + d * 6);
            const float2 v1 = *reinterpret_cast<const float2 *>(pb + d * 6 + 2);
            const float2 v2 = *reinterpret_cast<const float2 *>(pb + d * 6 + 4);
            dc = class_of(v0.x, C);
            conf = v0.y; dx = v1.x; dy = v1.y; dw = v2.x; dh = v2.y;
        }
        float best = 0.0f;                                               // utils.py:382 (unwritten slot reads 0)
        int bj = 0;                                                      // utils.py:383
        for (int g0 = 0; g0 < nt_i; g0 += 32) {
            const int g = g0 + lane;
            uint32_t gc = 0xffffffffu;
            float gx = 0.0f, gy = 0.0f, gw = 0.0f, gh = 0.0f;
            if (g < nt_i) {
                const float2 v0 = *reinterpret_cast<const float2 *>(tb + g * 6);
                const float2 v1 = *reinterpret_cast<const float2 *>(tb + g * 6 + 2);
                const float2 v2 = *reinterpret_cast<const float2 *>(tb + g * 6 + 4);
                gc = class_of(v0.x, C);
                gx = v1.x; gy = v1.y; gw = v2.x; gh = v2.y;
            }
            const int lim = min(32, nt_i - g0);
            for (int k = 0; k < lim; ++k) {                              // utils.py:386: ground truths of the image in row order
                const uint32_t kc = __shfl_sync(0xffffffffu, gc, k);
                const float kx = __shfl_sync(0xffffffffu, gx, k), ky = __shfl_sync(0xffffffffu, gy, k);
                const float kw = __shfl_sync(0xffffffffu, gw, k), kh = __shfl_sync(0xffffffffu, gh, k);
                if (kc == dc && dc < static_cast<uint32_t>(C)) {
                    const float v = iou_ref(dx, dy, dw, dh, kx, ky, kw, kh);         // utils.py:387 (det, gt)
                    if (v > best) { best = v; bj = g0 + k; }                        // utils.py:389
                }
            }
        }
        const bool hit = valid && best > iou_thr;                                  // utils.py:395
        const bool taken = hit && ((claimed[bj >> 5] >> (bj & 31)) & 1u);          // claimed by an earlier chunk
        const uint32_t peers = __match_any_sync(0xffffffffu, (hit && !taken) ? static_cast<uint32_t>(bj) : 0x10000u + lane);
        const bool tp = hit && !taken && lane == __ffs(peers) - 1;                 // utils.py:408-418: first in order claims
        __syncwarp();
        if (tp) atomicOr(&claimed[bj >> 5], 1u << (bj & 31));
        __syncwarp();
        if (valid) emit(d, make_rec(dc, conf, tp ? 1u : 0u));
    }
}
#endif  // __CUDACC__

int scan_ws_for(cudaStream_t st, struct ScanWs **out);

size_t radix_ws_bytes(int64_t n_max, int C);
size_t match_ws_bytes(int64_t nt, int64_t np, bool sorted);
int reduce_impl(int nseg, const uint64_t *const *rec, const int64_t *nrec_max, const int64_t *const *nrec_dev, int n_gt,
                const int32_t *const *gt_parts, int C, float *out_ap, float *out_map, const uint64_t *wait_flags, int wait_n,
                uint64_t wait_epoch, int32_t *err, int64_t n_hint, void *workspace, size_t ws_bytes, void *stream);
int radix_launch(ReduceArgs &a, int64_t n_max, void *workspace, size_t ws_bytes, cudaStream_t st);

}  // namespace yh
