#pragma once
// yh_decode_nms_impl.cuh - K1..K3 (templates shared by yh_decode_nms.cu and yh_decode_nms_half.cu): grid decode, confidence filter + stable rank, per-class greedy
// NMS (bitmask IoU + fixed-point greedy scan), element-wise IoU.  sm_100a.
//
// Reference arithmetic replaced (paths relative to the reference repo, yolo_v1/):
//   decode  utils.py:152-218   NMS  utils.py:79-114   IoU  utils.py:9-43
//   fused loop body of MeanAveragePrecision.update_state  utils.py:470-480
//
// Work decomposition (one warp owns one image at a time):
//   A  each lane decodes cells lane, lane+32, ... (class argmax, best-confidence box,
//      cell offset) and tests conf > thr; survivors are compacted with ballot/popc
//   B  stable descending rank r_i = #{j : s_j > s_i} (+ #{j < i : s_j == s_i} only when a
//      duplicate rank shows a tie exists); candidates are scattered to rank order in smem
//   C  same-class lane masks with match.any, published per class in a small smem table
//   D  every candidate tests only its same-class predecessors -> suppression bit words
//   E  greedy keep flags by fixed-point iteration on ballots (exact: position q depends
//      only on positions < q, so the iteration converges to the sequential result)
//   F  kept rows are written in rank order; count per image
// Kernels: decode_nms_tma_kernel (images <= 12 KB: tiles of whole images through a TMA ring, one warp
// per image), decode_nms_coop_kernel (bigger images: 32-cell chunks through a TMA ring, a team of
// warps per image, one thread per cell), decode_nms_direct_kernel (tails, unaligned inputs, anything
// else), nms_rows_kernel / decode_kernel / iou_kernel (the reference's separate calls).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "yh_common.cuh"

namespace yh {

struct NmsCfg {
    int S, B, C, M, D;       // grid, boxes, classes, cells = S*S, channels = C + 5B
    float inv_s;             // float32(1 / S)                         (utils.py:207)
    float iou_thr, conf_thr;
    int thr_fast;            // iou_thr is a positive normal float: the division-free filter of suppresses() applies
    float band;              // 2^-20 when it does, +inf when the division must always decide
    int score_mode;          // 0 = reference (best box confidence); 1 = confidence x class probability (extension)
    int ws_bytes;            // per-warp workspace bytes
    int tbl_rows;            // rows of the class table (C for fused, 32*NS for row input)
};

constexpr unsigned FULL = 0xffffffffu;

#ifndef YH_RANK_INT
#define YH_RANK_INT 1
#endif
// Bits of (a > b) ? 1.0f : 0.0f, i.e. 127 << 23 or 0: one FSET.BF.  The rank loop adds these bit
// patterns with three-input INTEGER adds (two comparisons per IADD3); the sum is count * 127 * 2^23
// mod 2^32, and count (< 512) is recovered as ((sum >> 23) * 127^-1) mod 512 with 127^-1 = 383 (mod 512).
__device__ __forceinline__ int gt_bits(float a, float b) { return __float_as_int((a > b) ? 1.0f : 0.0f); }

// per-warp workspace carve-up (MP = 32 * NS slots).  Only what other lanes must see lives in
// shared memory: the compact confidences for the rank loop, and per RANK POSITION the box
// corners / area / class key for the IoU tests and the output slot.  Confidence, box and class
// of a cell stay in the registers of the lane that decoded it.
template <int NS, bool kFloatCls>
struct WarpWs {
    static constexpr int MP = 32 * NS;
    float4 *scor;     // [MP]  corners (xmin, xmax, ymin, ymax) of the box at rank position q  (utils.py:24-32)
    float *sarea;     // [MP]  |area| of that box                                               (utils.py:40-41)
    int *smeta;       // [MP]  class key of rank position q (scratch for the duplicate-rank test before)
    float *ckey;      // [MP + 4] compact candidate confidences (source order), -inf padded;
    int *outpos;      //          aliased afterwards: output slot of rank position q, -1 = suppressed
    float *cclsf;     // [MP]  compact float classes          (kFloatCls only)
    unsigned *tbl;    // [tbl_rows * NS] class key -> lanes holding that class, per slot

    __host__ __device__ static int bytes(int tbl_rows)
    {
        int b = MP * 16 + MP * 4 + MP * 4 + (MP + 4) * 4;
        if (kFloatCls) b += MP * 4;
        b += tbl_rows * NS * 4;
        return (b + 15) & ~15;
    }
    __device__ explicit WarpWs(unsigned char *p)
    {
        scor = reinterpret_cast<float4 *>(p);  p += MP * 16;
        sarea = reinterpret_cast<float *>(p);  p += MP * 4;
        smeta = reinterpret_cast<int *>(p);    p += MP * 4;
        ckey = reinterpret_cast<float *>(p);
        outpos = reinterpret_cast<int *>(p);   p += (MP + 4) * 4;
        cclsf = nullptr;
        if (kFloatCls) {
            cclsf = reinterpret_cast<float *>(p);  p += MP * 4;
        }
        tbl = reinterpret_cast<unsigned *>(p);
    }
};

// IoU test of the reference from precomputed corners/areas: p = chosen (earlier) box, q = later
// box.  Same float32 operations in the same order as utils.py:34-43 (min/max/+ commute bit for
// bit), so the decision is the reference's: suppress iff !(fl32(inter / den) < thr) (utils.py:108).
//
// The IEEE division (a ~16-instruction sequence with a slow path) is almost never evaluated.  With
// p = fl(thr * den) and d = fl(inter - p) (den > 0, thr > 0): |d| > 2^-20 p implies that inter / den
// differs from thr by more than a relative 2^-21 (p carries a relative error <= 2^-24 and the
// subtraction is correctly rounded, so its sign and size are right), which is far outside the
// rounding interval of the quotient (relative 2^-24 around thr): the sign of d decides exactly like
// fl(inter / den) < thr.  Inside the band (probability ~1e-6 per test on continuous data), for
// den <= 0 (degenerate negative-extent boxes), NaN/inf operands or an unusual thr, the reference's
// own division decides.
//
// den > 0 needs no test of its own: both areas are absolute values, the clipped extents of the intersection are no
// larger than either box's own (so inter <= min(a1, a2), and inter = 0 as soon as a box has a negative extent), hence
// den >= 1e-6 for finite operands; a NaN or infinite den makes p and d NaN / infinite and the comparison below false.
// cfg.band is 2^-20 when the filter applies and +inf otherwise (YH_EXACT_DIV=1, unusual thr): `|d| > |p| * inf` is false
// (|p| * inf is +inf, or NaN for p = 0).
// The common path is straight-line code (one rarely taken branch to the division).
static __device__ __noinline__ unsigned suppresses_exact(float inter, float den, float thr)   // out of line: taken ~1e-6 of the time
{
    float v = 0.0f;                                          // a zero intersection gives IoU = +0 exactly
    if (inter != 0.0f) v = __fdiv_rn(inter, den);
    return (v < thr) ? 0u : 1u;                              // utils.py:108 keeps iff iou < thr
}

__device__ __forceinline__ unsigned suppresses(const float4 &pc, float pa, const float4 &qc, float qa, const NmsCfg &cfg)
{
    const float iw = clip01(__fsub_rn(fminf(pc.y, qc.y), fmaxf(pc.x, qc.x)));
    const float ih = clip01(__fsub_rn(fminf(pc.w, qc.w), fmaxf(pc.z, qc.z)));
    const float inter = __fmul_rn(iw, ih);
    const float den = __fadd_rn(__fsub_rn(__fadd_rn(pa, qa), inter), 1e-6f);
    const float p = __fmul_rn(cfg.iou_thr, den);
    const float d = __fsub_rn(inter, p);
    unsigned s = (d < 0.0f) ? 0u : 1u;                       // 1 = suppresses
    if (__builtin_expect(!(fabsf(d) > __fmul_rn(fabsf(p), cfg.band)), 0))   // inside the band, or not finite:
        s = suppresses_exact(inter, den, cfg.iou_thr);                       // the reference's division decides
    return s;
}

// ------------------------------------------------------------------------------------------
// Phases B..F for one image held in registers: slot t of this lane is source index
// lane + 32 t.  cls[t] is the class id (fused) or the raw float bits (kFloatCls).
// Returns K (kept rows), identical on all lanes.
// ------------------------------------------------------------------------------------------
template <int NS, bool kFloatCls>
__device__ __forceinline__ int nms_warp(const float (&conf)[NS], const float4 (&box)[NS], const int (&cls)[NS],
                                        const bool (&valid)[NS], const NmsCfg &cfg, WarpWs<NS, kFloatCls> &ws,
                                        float *__restrict__ out_rows, int *__restrict__ out_idx)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;

    // ---- A': compaction of the survivors (utils.py:95, strict >) ----
    unsigned pass_m = 0u;                          // bit t: this lane's slot t survives the filter
    int ci[NS];
    int n = 0;
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        const bool ps = valid[t] && (conf[t] > cfg.conf_thr);
        pass_m |= ps ? (1u << t) : 0u;
        const unsigned b = __ballot_sync(FULL, ps);
        ci[t] = n + __popc(b & lt_mask);
        n += __popc(b);
        if (ps) {
            ws.ckey[ci[t]] = conf[t];
            if (kFloatCls) ws.cclsf[ci[t]] = __int_as_float(cls[t]);
        }
    }
    if (n == 0) return 0;
    if (lane < 4) ws.ckey[n + lane] = -INFINITY;   // pad to a multiple of 4 for the float4 loop
    __syncwarp();

    // ---- B: stable descending rank (utils.py:98): r = #{j : s_j > s}.  Each comparison is one
    //      FSET.BF; two results are folded into the counter by one three-input integer add (see
    //      gt_bits), i.e. 1.5 instructions per comparison; n <= 256 < 512 ----
    int r[NS];
    {
#if YH_RANK_INT
        int acc[NS];
#pragma unroll
        for (int t = 0; t < NS; ++t) acc[t] = 0;
        const float4 *k4p = reinterpret_cast<const float4 *>(ws.ckey);
        const int n4 = (n + 3) >> 2;
        for (int g = 0; g < n4; ++g) {
            const float4 k = k4p[g];
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                const float c = conf[t];
                acc[t] = acc[t] + gt_bits(k.x, c) + gt_bits(k.y, c);
                acc[t] = acc[t] + gt_bits(k.z, c) + gt_bits(k.w, c);
            }
        }
#pragma unroll
        for (int t = 0; t < NS; ++t) r[t] = ((static_cast<unsigned>(acc[t]) >> 23) * 383u) & 511u;
#else
        float rf[NS];
#pragma unroll
        for (int t = 0; t < NS; ++t) rf[t] = 0.0f;
        const float4 *k4p = reinterpret_cast<const float4 *>(ws.ckey);
        const int n4 = (n + 3) >> 2;
        for (int g = 0; g < n4; ++g) {
            const float4 k = k4p[g];
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                const float c = conf[t];
                rf[t] += (k.x > c) ? 1.0f : 0.0f;
                rf[t] += (k.y > c) ? 1.0f : 0.0f;
                rf[t] += (k.z > c) ? 1.0f : 0.0f;
                rf[t] += (k.w > c) ? 1.0f : 0.0f;
            }
        }
#pragma unroll
        for (int t = 0; t < NS; ++t) r[t] = static_cast<int>(rf[t]);
#endif
    }
    // a duplicate rank <=> equal confidences exist: only then pay for the tie-break pass
    {
#pragma unroll
        for (int t = 0; t < NS; ++t)
            if (((pass_m >> t) & 1u)) ws.smeta[r[t]] = ci[t];
        __syncwarp();
        bool dup = false;
#pragma unroll
        for (int t = 0; t < NS; ++t)
            if (((pass_m >> t) & 1u)) dup |= (ws.smeta[r[t]] != ci[t]);
        dup = __any_sync(FULL, dup);
        if (dup) {
            // Equal confidences keep their candidate order (stable sort, utils.py:98).  The members of a tie group
            // all computed the rank G of its first member and own the positions G .. G+g-1: in every round the
            // member with the lowest candidate index still waiting takes the group's next position (shared-memory
            // atomicMin), the others move one position down.  Rounds = size of the largest tie group.
            for (int q = lane; q < n; q += 32) ws.smeta[q] = 0x7fffffff;
            __syncwarp();
            unsigned pend = pass_m;
            for (;;) {
#pragma unroll
                for (int t = 0; t < NS; ++t)
                    if ((pend >> t) & 1u) atomicMin(&ws.smeta[r[t]], ci[t]);
                __syncwarp();
                unsigned lost = 0u;
#pragma unroll
                for (int t = 0; t < NS; ++t) {
                    if (((pend >> t) & 1u) && ws.smeta[r[t]] != ci[t]) {
                        lost |= 1u << t;
                        r[t] += 1;
                    }
                }
                pend = lost;
                if (!__any_sync(FULL, pend != 0u)) break;
            }
        }
    }
    // class key: the class id itself, or (row input) the first candidate holding an equal class
    int key[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) key[t] = kFloatCls ? ci[t] : cls[t];
    if (kFloatCls) {
        // Row input carries the class as a float (utils.py:175).  Usual case: every candidate's class is a
        // small non-negative integer -> the integer is the key, as in the fused kernels.  Otherwise (any
        // float, NaN never equal to anything) fall back to "first candidate holding an equal class".
        bool small_int = true;
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            if ((pass_m >> t) & 1u) {
                const float f = __int_as_float(cls[t]);
                const int v = __float2int_rz(f);
                small_int = small_int && (f == static_cast<float>(v)) && v >= 0 && v < cfg.tbl_rows;
            }
        }
        if (__all_sync(FULL, small_int)) {
#pragma unroll
            for (int t = 0; t < NS; ++t) key[t] = __float2int_rz(__int_as_float(cls[t]));
        } else {
            for (int j = n - 1; j >= 0; --j) {
                const float f = ws.cclsf[j];
#pragma unroll
                for (int t = 0; t < NS; ++t)
                    if (((pass_m >> t) & 1u) && f == __int_as_float(cls[t])) key[t] = j;
            }
        }
    }
    __syncwarp();   // every lane is done with smeta (scratch) and ckey before they are rewritten
    // scatter corners / area / class key to rank order                              (utils.py:24-32,40)
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        if (((pass_m >> t) & 1u)) {
            const int q = r[t];
            const float xn = __fmul_rn(__fsub_rn(box[t].x, box[t].z), 0.5f), xx = __fmul_rn(__fadd_rn(box[t].x, box[t].z), 0.5f);
            const float yn = __fmul_rn(__fsub_rn(box[t].y, box[t].w), 0.5f), yx = __fmul_rn(__fadd_rn(box[t].y, box[t].w), 0.5f);
            ws.scor[q] = make_float4(xn, xx, yn, yx);
            ws.sarea[q] = fabsf(__fmul_rn(__fsub_rn(xx, xn), __fsub_rn(yx, yn)));
            ws.smeta[q] = key[t];
        }
    }
    __syncwarp();

    // ---- C: same-class masks.  Slot t of this lane now is rank position q = lane + 32 t ----
    const int NT = (n + 31) >> 5;
    unsigned act_m = 0u;                           // bit t: rank position lane + 32 t exists
    int qkey[NS];
    unsigned lead = 0;
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        const int q = lane + 32 * t;
        const bool ac = q < n;
        act_m |= ac ? (1u << t) : 0u;
        qkey[t] = ac ? ws.smeta[q] : 0;
    }
    // all MATCH instructions are issued back to back (their latency overlaps), the table is written afterwards
    unsigned mm[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t)
        mm[t] = (t < NT) ? __match_any_sync(FULL, ((act_m >> t) & 1u) ? qkey[t] : (0x7f000000 + lane)) : 0u;
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        if (t < NT && ((act_m >> t) & 1u) && (__ffs(mm[t]) - 1) == lane) {
            ws.tbl[qkey[t] * NS + t] = mm[t];
            lead |= 1u << t;
        }
    }
    __syncwarp();

    // ---- D: suppression bits against same-class predecessors (utils.py:108) ----
    unsigned supp[NS][NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) {
#pragma unroll
        for (int t2 = 0; t2 < NS; ++t2) supp[t][t2] = 0u;
    }
    if constexpr (NS == 2) {
        // One loop per lane over ALL its predecessor bits (slot-1 candidate first, then slot 0), so
        // the warp runs max_lane(total) trips instead of the sum of three per-word maxima.
        unsigned w0 = 0u, cur_lo = 0u, cur_hi = 0u;
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), qc = c0;
        float a0 = 0.f, qa = 0.f;
        unsigned pending0 = 0u, on1 = 0u;
        if ((act_m & 1u)) {
            w0 = ws.tbl[qkey[0] * 2] & lt_mask;
            c0 = ws.scor[lane];
            a0 = ws.sarea[lane];
        }
        if ((act_m & 2u)) {
            cur_lo = ws.tbl[qkey[1] * 2];
            cur_hi = ws.tbl[qkey[1] * 2 + 1] & lt_mask;
        }
        if (cur_lo | cur_hi) {
            qc = ws.scor[lane + 32];
            qa = ws.sarea[lane + 32];
            pending0 = (w0 != 0u) ? 1u : 0u;
            on1 = 1u;
        } else {
            cur_lo = w0;
            qc = c0;
            qa = a0;
        }
        unsigned s_lo = 0u, s_hi = 0u;
        for (;;) {
            if ((cur_lo | cur_hi) == 0u) {
                if (!pending0) break;
                supp[1][0] = s_lo; supp[1][1] = s_hi;      // slot-1 candidate done, switch to slot 0
                s_lo = s_hi = 0u;
                cur_lo = w0; qc = c0; qa = a0;
                pending0 = 0u; on1 = 0u;
            }
            const bool lo = cur_lo != 0u;
            const unsigned w = lo ? cur_lo : cur_hi;
            const unsigned bit = w & (0u - w);
            const int p = (31 - __clz(bit)) + (lo ? 0 : 32);
            const unsigned sb = suppresses(ws.scor[p], ws.sarea[p], qc, qa, cfg) ? bit : 0u;
            if (lo) { cur_lo ^= bit; s_lo |= sb; } else { cur_hi ^= bit; s_hi |= sb; }
        }
        if (on1) { supp[1][0] = s_lo; supp[1][1] = s_hi; } else { supp[0][0] = s_lo; }
    } else {
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            if (t < NT && ((act_m >> t) & 1u)) {
                const unsigned *row = ws.tbl + qkey[t] * NS;
                const float4 qc = ws.scor[lane + 32 * t];
                const float qa = ws.sarea[lane + 32 * t];
#pragma unroll
                for (int t2 = 0; t2 <= t; ++t2) {
                    unsigned w = row[t2];
                    if (t2 == t) w &= lt_mask;
                    while (w) {
                        const int b = __ffs(w) - 1;
                        w &= w - 1;
                        supp[t][t2] |= suppresses(ws.scor[32 * t2 + b], ws.sarea[32 * t2 + b], qc, qa, cfg) << b;
                    }
                }
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NS; ++t)
        if (lead & (1u << t)) ws.tbl[qkey[t] * NS + t] = 0u;   // leave the table zeroed

    // ---- E: greedy keep flags, fixed point of keep[q] = !any(supp[q] & keep) ----
    unsigned alive_m = act_m;                                  // bit t: rank position lane + 32 t is kept
    unsigned kw[NS];
    for (;;) {
#pragma unroll
        for (int t = 0; t < NS; ++t) kw[t] = (t < NT) ? __ballot_sync(FULL, (alive_m >> t) & 1u) : 0u;
        unsigned nm = 0u;
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            unsigned sgot = 0u;
#pragma unroll
            for (int t2 = 0; t2 <= t; ++t2) sgot |= supp[t][t2] & kw[t2];
            nm |= (sgot == 0u) ? (1u << t) : 0u;
        }
        nm &= act_m;
        const bool ch = nm != alive_m;
        alive_m = nm;
        if (!__any_sync(FULL, ch)) break;
    }
    // ---- F: output slot of every rank position, then each cell's lane writes its own row in
    //      pick order (utils.py:112) from the registers it decoded into ----
    int K = 0;
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        if (((act_m >> t) & 1u)) ws.outpos[lane + 32 * t] = ((alive_m >> t) & 1u) ? K + __popc(kw[t] & lt_mask) : -1;
        K += __popc(kw[t]);
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        if (((pass_m >> t) & 1u)) {
            const int pos = ws.outpos[r[t]];
            if (pos >= 0) {
                const float c = kFloatCls ? __int_as_float(cls[t]) : static_cast<float>(cls[t]);   // utils.py:175
                float2 *o = reinterpret_cast<float2 *>(out_rows + 6 * pos);
                o[0] = make_float2(c, conf[t]);
                o[1] = make_float2(box[t].x, box[t].y);
                o[2] = make_float2(box[t].z, box[t].w);
                if (out_idx) out_idx[pos] = lane + 32 * t;
            }
        }
    }
    __syncwarp();   // workspace is reused by the next image
    return K;
}

// ------------------------------------------------------------------------------------------
// Element types of the prediction tensor: float32 (the reference), and - head adapter, SURVEY.md 8f N3 -
// float16 / bfloat16 heads, widened EXACTLY to float32 as they are read, so that every result equals
// the float32 path run on the widened tensor while HBM traffic is halved.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float *p, int j) { return p[j]; }
__device__ __forceinline__ float ldf(const __half *p, int j) { return __half2float(p[j]); }
__device__ __forceinline__ float ldf(const __nv_bfloat16 *p, int j) { return __bfloat162float(p[j]); }
// two consecutive elements from a 4-byte aligned pair
__device__ __forceinline__ float2 ldf2(const float *p, int i) { return reinterpret_cast<const float2 *>(p)[i]; }
__device__ __forceinline__ float2 ldf2(const __half *p, int i) { return __half22float2(reinterpret_cast<const __half2 *>(p)[i]); }
__device__ __forceinline__ float2 ldf2(const __nv_bfloat16 *p, int i)
{
    const uint32_t w = reinterpret_cast<const uint32_t *>(p)[i];       // bf16 -> f32 is a 16-bit left shift
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// Packed 16-bit pairs: the class argmax of a half-precision head is taken on the packed values themselves
// (ordering and equality of fp16 / bf16 values are those of their exact float32 widenings), so the class scores
// are never unpacked: HMNMX2 tree, then one two-predicate packed compare per pair.
template <typename E> struct Pair16;
template <> struct Pair16<__half> {
    using T2 = __half2;
    static __device__ __forceinline__ float2 widen(T2 v) { return __half22float2(v); }
    static __device__ __forceinline__ T2 bcast_max(T2 m) { const __half b = __hmax(__low2half(m), __high2half(m)); return __halves2half2(b, b); }
    static __device__ __forceinline__ float lo(T2 v) { return __low2float(v); }
    // cls = 2i+1 if the high element equals, then 2i if the low one does (lowest index wins)
    static __device__ __forceinline__ void pick(int &cls, T2 v, T2 best, int i)
    {
        asm("{\n\t.reg .pred p, q;\n\tsetp.eq.f16x2 p|q, %1, %2;\n\t@q mov.s32 %0, %3;\n\t@p mov.s32 %0, %4;\n\t}"
            : "+r"(cls) : "r"(*reinterpret_cast<const unsigned *>(&v)), "r"(*reinterpret_cast<const unsigned *>(&best)), "r"(2 * i + 1), "r"(2 * i));
    }
};
template <> struct Pair16<__nv_bfloat16> {
    using T2 = __nv_bfloat162;
    static __device__ __forceinline__ float2 widen(T2 v)
    {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(&v);
        return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
    }
    static __device__ __forceinline__ T2 bcast_max(T2 m) { const __nv_bfloat16 b = __hmax(__low2bfloat16(m), __high2bfloat16(m)); return __halves2bfloat162(b, b); }
    static __device__ __forceinline__ float lo(T2 v) { return __low2float(v); }
    static __device__ __forceinline__ void pick(int &cls, T2 v, T2 best, int i)
    {
        asm("{\n\t.reg .pred p, q;\n\tsetp.eq.bf16x2 p|q, %1, %2;\n\t@q mov.s32 %0, %3;\n\t@p mov.s32 %0, %4;\n\t}"
            : "+r"(cls) : "r"(*reinterpret_cast<const unsigned *>(&v)), "r"(*reinterpret_cast<const unsigned *>(&best)), "r"(2 * i + 1), "r"(2 * i));
    }
};

// ------------------------------------------------------------------------------------------
// Phase A: decode one cell (utils.py:173-208).  p may point to shared or global memory.
// ------------------------------------------------------------------------------------------
template <int CT, int BT, typename E>
__device__ __forceinline__ void decode_cell(const E *__restrict__ p, const NmsCfg &cfg, float colf, float rowf,
                                            int &cls, float &conf, float4 &box)
{
    float bx, by, bw, bh;
    float cbest;                                                   // the winning class score (score_mode 1 only)
    if constexpr (CT > 0 && BT > 0 && ((CT + 5 * BT) % 2 == 0) && (CT % 2 == 0) && !std::is_same<E, float>::value) {
        // half-precision head: class argmax on the packed pairs, only the box values are widened
        using P = Pair16<E>;
        using T2 = typename P::T2;
        constexpr int D = CT + 5 * BT;
        const T2 *p2 = reinterpret_cast<const T2 *>(p);
        T2 w[D / 2];
#pragma unroll
        for (int i = 0; i < D / 2; ++i) w[i] = p2[i];
        T2 m = w[0];
#pragma unroll
        for (int i = 1; i < CT / 2; ++i) m = __hmax2(m, w[i]);
        const T2 bb = P::bcast_max(m);
        cls = CT - 1;
#pragma unroll
        for (int i = CT / 2 - 1; i >= 0; --i) P::pick(cls, w[i], bb, i);
        cbest = P::lo(bb);
        float v[5 * BT];
#pragma unroll
        for (int i = 0; i < (5 * BT) / 2; ++i) {
            const float2 x = P::widen(w[CT / 2 + i]);
            v[2 * i] = x.x;
            v[2 * i + 1] = x.y;
        }
        conf = v[0]; bx = v[1]; by = v[2]; bw = v[3]; bh = v[4];
#pragma unroll
        for (int b = 1; b < BT; ++b) {
            if (v[5 * b] > conf) {                                 // first max over boxes (utils.py:183)
                conf = v[5 * b]; bx = v[5 * b + 1]; by = v[5 * b + 2];
                bw = v[5 * b + 3]; bh = v[5 * b + 4];
            }
        }
    } else if constexpr (CT > 0 && BT > 0 && ((CT + 5 * BT) % 2 == 0)) {
        constexpr int D = CT + 5 * BT;
        float v[D];
#pragma unroll
        for (int i = 0; i < D / 2; ++i) {
            const float2 x = ldf2(p, i);
            v[2 * i] = x.x;
            v[2 * i + 1] = x.y;
        }
        // first maximum (tf.argmax, utils.py:173): the maximum by a tree of 3-input FMNMX, then the lowest index
        // holding it, found by a descending chain of predicated moves (2 instructions per class instead of 3)
        float best = v[0];
#pragma unroll
        for (int j = 1; j + 1 < CT; j += 2) best = fmaxf(fmaxf(best, v[j]), v[j + 1]);
        if constexpr (CT % 2 == 0) best = fmaxf(best, v[CT - 1]);
        cls = CT - 1;
#pragma unroll
        for (int j = CT - 2; j >= 0; --j)
            if (v[j] == best) cls = j;
        cbest = best;
        conf = v[CT]; bx = v[CT + 1]; by = v[CT + 2]; bw = v[CT + 3]; bh = v[CT + 4];
#pragma unroll
        for (int b = 1; b < BT; ++b) {
            if (v[CT + 5 * b] > conf) {                            // first max over boxes (utils.py:183)
                conf = v[CT + 5 * b]; bx = v[CT + 5 * b + 1]; by = v[CT + 5 * b + 2];
                bw = v[CT + 5 * b + 3]; bh = v[CT + 5 * b + 4];
            }
        }
    } else if constexpr (CT > 0 && BT > 0) {
        // odd channel count (rows are only 4-byte aligned): scalar loads with immediate offsets.
        // Class argmax in blocks of 8: block maximum (one FMNMX per score), the FIRST block whose
        // maximum beats the running one (strict >, so ties stay with the earlier block), then the
        // first score of that block equal to the maximum = tf.argmax's first maximum (utils.py:173).
        constexpr int NB = (CT + 7) / 8;
        float best = 0.f;
        int bb = 0;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float m = ldf(p, 8 * b);
#pragma unroll
            for (int i = 1; i < 8; ++i)
                if (8 * b + i < CT) m = fmaxf(m, ldf(p, 8 * b + i));
            if (b == 0) best = m;
            else if (m > best) { best = m; bb = b; }
        }
        cls = 8 * bb;
#pragma unroll
        for (int i = 7; i >= 1; --i) {
            const int j = min(8 * bb + i, CT - 1);
            if (ldf(p, j) == best) cls = j;
        }
        if (ldf(p, 8 * bb) == best) cls = 8 * bb;
        cbest = best;
        int k = 0;
        conf = ldf(p, CT);
#pragma unroll
        for (int b = 1; b < BT; ++b) {
            const float x = ldf(p, CT + 5 * b);
            if (x > conf) { conf = x; k = b; }                    // first max over boxes (utils.py:183)
        }
        const E *q = p + CT + 5 * k;
        bx = ldf(q, 1); by = ldf(q, 2); bw = ldf(q, 3); bh = ldf(q, 4);
    } else {
        const int C = cfg.C, B = cfg.B;
        cls = 0;
        float best = ldf(p, 0);
        for (int j = 1; j < C; ++j) {
            const float x = ldf(p, j);
            if (x > best) { best = x; cls = j; }
        }
        cbest = best;
        int k = 0;
        conf = ldf(p, C);
        for (int b = 1; b < B; ++b) {
            const float x = ldf(p, C + 5 * b);
            if (x > conf) { conf = x; k = b; }
        }
        const E *q = p + C + 5 * k;
        bx = ldf(q, 1); by = ldf(q, 2); bw = ldf(q, 3); bh = ldf(q, 4);
    }
    // Extension, NOT the reference (utils.py:173,183-197 scores a cell by its best box confidence only):
    // score = confidence x winning class probability, the class-specific score of the YOLO paper
    if (cfg.score_mode != 0) conf = __fmul_rn(conf, cbest);
    box.x = __fmul_rn(cfg.inv_s, __fadd_rn(bx, colf));             // utils.py:207 (column)
    box.y = __fmul_rn(cfg.inv_s, __fadd_rn(by, rowf));             // utils.py:208 (row)
    box.z = bw;
    box.w = bh;
}

// ------------------------------------------------------------------------------------------
// Direct kernel: one warp per image, cells read straight from global memory / L1, no CTA-level
// cooperation.  Covers tails, inputs that are only 8-byte aligned and every shape the ring kernels
// cannot take.  Any shape (S*S <= 32 NS), any alignment >= 4 B.  (Two staged variants - per-slot LDG
// staging and a per-warp cp.async.bulk double buffer - were measured slower, because they leave
// fewer resident warps for what is a serial chain per warp, and were removed.)
// ------------------------------------------------------------------------------------------
template <int NS, int CT, int BT, typename E>
__global__ void __launch_bounds__(512) decode_nms_direct_kernel(const E *__restrict__ pred, int64_t n, NmsCfg cfg,
                                                                float *__restrict__ out_boxes,
                                                                int *__restrict__ out_count, int *__restrict__ out_idx)
{
    extern __shared__ uint4 smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    WarpWs<NS, false> ws(reinterpret_cast<unsigned char *>(smem_raw) + static_cast<size_t>(warp) * cfg.ws_bytes);
    for (int i = lane; i < cfg.tbl_rows * NS; i += 32) ws.tbl[i] = 0u;
    __syncwarp();

    float colf[NS], rowf[NS];
    bool valid[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        const int cell = lane + 32 * t;
        valid[t] = cell < cfg.M;
        rowf[t] = static_cast<float>(cell / cfg.S);
        colf[t] = static_cast<float>(cell % cfg.S);
    }
    for (int64_t img = static_cast<int64_t>(blockIdx.x) * wpb + warp; img < n;
         img += static_cast<int64_t>(gridDim.x) * wpb) {
        const E *base = pred + img * cfg.M * cfg.D;
        float conf[NS];
        float4 box[NS];
        int cls[NS];
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            conf[t] = -INFINITY; cls[t] = 0; box[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid[t]) decode_cell<CT, BT>(base + (lane + 32 * t) * cfg.D, cfg, colf[t], rowf[t], cls[t], conf[t], box[t]);
        }
        const int K = nms_warp<NS, false>(conf, box, cls, valid, cfg, ws, out_boxes + img * cfg.M * 6,
                                          out_idx ? out_idx + img * cfg.M : nullptr);
        if (lane == 0) out_count[img] = K;
    }
}

// ------------------------------------------------------------------------------------------
// TMA kernel: persistent CTAs, warp 0 = producer (cp.async.bulk of T-image tiles into an
// ST-deep mbarrier ring), W consumer warps in G = W/T groups; group g takes the CTA's tiles
// g, g+G, ...; warp j of the group owns image j of the tile.  The stage is released as soon
// as the image has been decoded into registers, before the NMS phases.
// ------------------------------------------------------------------------------------------
struct TmaCfg {
    int T, W, ST;           // images per tile, consumer warps, stages
    uint32_t tile_bytes;    // T * 4 * M * D, multiple of 16
    int64_t n_tiles;
};

template <int NS, int CT, int BT, typename E>
__global__ void __launch_bounds__(864, 1) decode_nms_tma_kernel(const E *__restrict__ pred, NmsCfg cfg, TmaCfg tc,
                                                                float *__restrict__ out_boxes,
                                                                int *__restrict__ out_count, int *__restrict__ out_idx)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *tiles = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + static_cast<size_t>(tc.ST) * tc.tile_bytes);
    uint64_t *empty = full + tc.ST;
    unsigned char *ws_base = reinterpret_cast<unsigned char *>(empty + tc.ST);
    ws_base += (16 - (reinterpret_cast<uintptr_t>(ws_base) & 15)) & 15;

    if (threadIdx.x == 0) {
        for (int s = 0; s < tc.ST; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, tc.T);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t my_tiles = (tc.n_tiles > blockIdx.x) ? (tc.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            const unsigned char *src = reinterpret_cast<const unsigned char *>(pred);
            int s = 0;
            uint32_t ph = 0;
            for (int64_t it = 0; it < my_tiles; ++it) {
                mbar_wait_relaxed(empty + s, ph ^ 1u, 256);
                const int64_t tile = blockIdx.x + it * gridDim.x;
                mbar_arrive_expect_tx(full + s, tc.tile_bytes);
                bulk_g2s(tiles + static_cast<size_t>(s) * tc.tile_bytes, src + tile * tc.tile_bytes, tc.tile_bytes,
                         full + s, pol);
                if (++s == tc.ST) { s = 0; ph ^= 1u; }
            }
        }
        return;
    }

    const int cw = warp - 1;              // consumer warp index
    const int G = tc.W / tc.T;
    const int g = cw / tc.T, j = cw % tc.T;
    WarpWs<NS, false> ws(ws_base + cw * cfg.ws_bytes);
    for (int i = lane; i < cfg.tbl_rows * NS; i += 32) ws.tbl[i] = 0u;
    __syncwarp();

    float colf[NS], rowf[NS];
    bool valid[NS];
#pragma unroll
    for (int t = 0; t < NS; ++t) {
        const int cell = lane + 32 * t;
        valid[t] = cell < cfg.M;
        rowf[t] = static_cast<float>(cell / cfg.S);
        colf[t] = static_cast<float>(cell % cfg.S);
    }
    const int img_elems = cfg.M * cfg.D;
    int s = g % tc.ST;                                   // stage / phase of tile `it`, kept incrementally
    uint32_t ph = static_cast<uint32_t>((g / tc.ST) & 1);
    for (int64_t it = g; it < my_tiles; it += G) {
        mbar_wait(full + s, ph);
        const E *base = reinterpret_cast<const E *>(tiles + static_cast<size_t>(s) * tc.tile_bytes) + j * img_elems;
        float conf[NS];
        float4 box[NS];
        int cls[NS];
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            conf[t] = -INFINITY; cls[t] = 0; box[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid[t]) decode_cell<CT, BT>(base + (lane + 32 * t) * cfg.D, cfg, colf[t], rowf[t], cls[t], conf[t], box[t]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);     // image is in registers: hand the slot back
        const int64_t img = (blockIdx.x + it * gridDim.x) * tc.T + j;
        const int K = nms_warp<NS, false>(conf, box, cls, valid, cfg, ws, out_boxes + img * cfg.M * 6,
                                          out_idx ? out_idx + img * cfg.M : nullptr);
        if (lane == 0) out_count[img] = K;
        s += G;
        while (s >= tc.ST) { s -= tc.ST; ph ^= 1u; }
    }
}

// ------------------------------------------------------------------------------------------
// Cooperative kernel for big images: a TEAM of TW = ceil(M/32) warps works on one image, one
// THREAD per grid cell / candidate / rank position, phases separated by named barriers
// (bar.sync id, TW*32); NTEAM teams per CTA run different images so that one team's barrier
// and load latencies are covered by the others.  Compared with one warp per image the serial
// chain per image is TW times shorter, the per-image state is ~11 KB per TEAM (not per warp), and
// the suppression bits of a candidate live in its own registers (no shared-memory bit rows).
//   warp 0           producer: cp.async.bulk of 32-cell chunks, in image order, into an ST-deep ring;
//                    after issuing chunk c it publishes c + 1 in `issued`
//   team T, warp t   consumes chunk t of the images T, T + NTEAM, ... of this CTA.  A consumer first
//                    waits until `issued` > c (then the stage's barrier is in the phase of chunk c and
//                    the parity wait is unambiguous for any ring depth), then on full[c % ST]
// Phases of a team (thread q = 32 * warp + lane is cell q, later rank position q):
//   A  decode own cell, release the stage;  A' ballot compaction across the team's warps
//   B  rank by counting over the compact confidences (+ tie pass if a duplicate rank shows up)
//   C  scatter corners / area / class to rank order; match.any per warp -> class table [class][warp]
//   D  own same-class predecessors (words 0..warp of the class row) -> suppression words in registers
//   E  greedy keep flags: fixed point on the team's ballot words (exact, see nms_warp)
//   F  output slots from the final keep words; every cell thread writes its own row
// ------------------------------------------------------------------------------------------
constexpr int kTeamWarpsMax = 8;
constexpr int kRankBuckets = 64;          // phase B of the team kernel: confidence buckets of the rank

struct CoopCfg {
    int ST, NTEAM, TW;      // ring stages, teams per CTA, warps per team (= chunks per image)
    uint32_t chunk_bytes, last_bytes;
    uint32_t slot_bytes;    // ring slot: chunk_bytes, + 16 when chunks may start off a 16-byte boundary
    int shifted;            // images are not a multiple of 16 bytes: each chunk is copied as the 16-byte aligned range
                            // around it and read at its offset a = address & 15 inside the slot
    int team_bytes;         // shared memory per team
    int64_t n;
};

__device__ __forceinline__ void team_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ bool team_any(int id, int nthreads, bool p)
{
    int r;
    asm volatile(
        "{\n\t.reg .pred q, o;\n\t"
        "setp.ne.s32 q, %3, 0;\n\t"
        "bar.red.or.pred o, %1, %2, q;\n\t"
        "selp.s32 %0, 1, 0, o;\n\t}"
        : "=r"(r)
        : "r"(id), "r"(nthreads), "r"(static_cast<int>(p))
        : "memory");
    return r != 0;
}

template <int CT, int BT, typename E>
__global__ void __launch_bounds__(1024, 1) decode_nms_coop_kernel(const E *__restrict__ pred, NmsCfg cfg, CoopCfg cc,
                                                                  float *__restrict__ out_boxes, int *__restrict__ out_count,
                                                                  int *__restrict__ out_idx)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *ring = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + static_cast<size_t>(cc.ST) * cc.slot_bytes);
    uint64_t *empty = full + cc.ST;
    volatile int64_t *issued = reinterpret_cast<volatile int64_t *>(empty + cc.ST);
    unsigned char *teams = reinterpret_cast<unsigned char *>(const_cast<int64_t *>(issued) + 2);

    if (threadIdx.x == 0) {
        for (int s = 0; s < cc.ST; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        *issued = 0;
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t my_imgs = (cc.n > blockIdx.x) ? (cc.n - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_chunks = my_imgs * cc.TW;
    const int64_t img_bytes = static_cast<int64_t>(sizeof(E)) * cfg.M * cfg.D;

    if (warp == 0) {                                               // ---- producer
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            const unsigned char *src = reinterpret_cast<const unsigned char *>(pred);
            int s = 0, t = 0;
            uint32_t ph = 0;
            int64_t i = 0;
            for (int64_t c = 0; c < n_chunks; ++c) {
                mbar_wait_relaxed(empty + s, ph ^ 1u, 64);
                const uint32_t bytes = (t == cc.TW - 1) ? cc.last_bytes : cc.chunk_bytes;
                const int64_t off = (blockIdx.x + i * gridDim.x) * img_bytes + static_cast<int64_t>(t) * cc.chunk_bytes;
                const uint32_t a = cc.shifted ? static_cast<uint32_t>(off & 15) : 0u;     // pred itself is 16-byte aligned
                const uint32_t cp = (bytes + a + 15u) & ~15u;
                mbar_arrive_expect_tx(full + s, cp);
                bulk_g2s(ring + static_cast<size_t>(s) * cc.slot_bytes, src + off - a, cp, full + s, pol);
                __threadfence_block();
                *issued = c + 1;
                if (++s == cc.ST) { s = 0; ph ^= 1u; }
                if (++t == cc.TW) { t = 0; ++i; }
            }
        }
        return;
    }
    const int team = (warp - 1) / cc.TW, wt = (warp - 1) % cc.TW;   // team, warp within the team
    if (team >= cc.NTEAM) return;
    const int q = 32 * wt + lane;                                   // cell index, later rank position
    const int nthr = cc.TW * 32, bar_id = 1 + team;
    const unsigned lt_mask = (1u << lane) - 1u;

    // team workspace
    unsigned char *tp = teams + static_cast<size_t>(team) * cc.team_bytes;
    const int MPT = cc.TW * 32;
    float4 *scor = reinterpret_cast<float4 *>(tp);           tp += MPT * 16;
    float *sarea = reinterpret_cast<float *>(tp);            tp += MPT * 4;
    int *smeta = reinterpret_cast<int *>(tp);                tp += MPT * 4;
    float *ckey = reinterpret_cast<float *>(tp);             tp += (MPT + 4) * 4;
    int *outpos = reinterpret_cast<int *>(tp);               tp += MPT * 4;
    int *wcnt = reinterpret_cast<int *>(tp);                 tp += kTeamWarpsMax * 4;
    unsigned *kws = reinterpret_cast<unsigned *>(tp);        tp += kTeamWarpsMax * 4;
    unsigned *help = reinterpret_cast<unsigned *>(tp);       tp += MPT * (kTeamWarpsMax / 2) * 4;   // re-dealt words of phase D
    int *bhist = reinterpret_cast<int *>(tp);                tp += kRankBuckets * 4;                 // phase B: candidates per bucket
    int *bbase = reinterpret_cast<int *>(tp);                tp += kRankBuckets * 4;                 //          candidates in higher buckets
    unsigned *tbl = reinterpret_cast<unsigned *>(tp);        // [C][TW]
    for (int i = q; i < cfg.C * cc.TW; i += nthr) tbl[i] = 0u;
    team_sync(bar_id, nthr);

    const bool valid = q < cfg.M;
    const int row = q / cfg.S, col = q - row * cfg.S;
    const float rowf = static_cast<float>(row), colf = static_cast<float>(col);

    for (int64_t i = team; i < my_imgs; i += cc.NTEAM) {
        const int64_t img = blockIdx.x + i * gridDim.x;
        if (q < kRankBuckets) bhist[q] = 0;                        // read last in phase B of the previous image; the A' barrier orders this
        // ---- A: decode own cell from the ring
        const int64_t c = i * cc.TW + wt;
        const int s = static_cast<int>(c % cc.ST);
        const uint32_t ph = static_cast<uint32_t>((c / cc.ST) & 1);
        while (*issued <= c) __nanosleep(32);
        mbar_wait(full + s, ph);
        int cls = 0;
        float conf = -INFINITY;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            const uint32_t a = cc.shifted ? static_cast<uint32_t>((img * img_bytes + static_cast<int64_t>(wt) * cc.chunk_bytes) & 15) : 0u;
            decode_cell<CT, BT>(reinterpret_cast<const E *>(ring + static_cast<size_t>(s) * cc.slot_bytes + a) + lane * cfg.D, cfg,
                                colf, rowf, cls, conf, box);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        // ---- A': compaction (utils.py:95, strict >)
        const bool pass = valid && (conf > cfg.conf_thr);
        const unsigned bal = __ballot_sync(FULL, pass);
        if (lane == 0) wcnt[wt] = __popc(bal);
        team_sync(bar_id, nthr);
        int n = 0, ci = 0;
#pragma unroll
        for (int w2 = 0; w2 < kTeamWarpsMax; ++w2) {
            if (w2 < cc.TW) {
                const int k = wcnt[w2];
                if (w2 < wt) ci += k;
                n += k;
            }
        }
        ci += __popc(bal & lt_mask);
        if (n == 0) {                                              // team-uniform
            if (q == 0) out_count[img] = 0;
            team_sync(bar_id, nthr);                               // wcnt is rewritten by the next image
            continue;
        }
        // ---- B: stable descending rank (utils.py:98): r = #{s_j > s_i} + #{j < i : s_j = s_i}.
        //      Not by comparing every pair (n^2 / 32 comparisons per warp): the candidates are first dropped into
        //      kRankBuckets confidence buckets (monotone in the confidence, so every candidate of a higher bucket ranks
        //      before every candidate of a lower one), a suffix sum over the bucket counts gives the rank of a bucket's
        //      first member, and only the handful of candidates that share a bucket are compared with each other -
        //      (confidence, source index) lexicographically, which is the stable order itself: no tie pass.
        int bkt = 0, slot = 0;
        if (pass) {
            bkt = min(max(static_cast<int>(__fmul_rn(conf, static_cast<float>(kRankBuckets))), 0), kRankBuckets - 1);
            slot = atomicAdd(&bhist[bkt], 1);
        }
        team_sync(bar_id, nthr);
        if (wt == 0) {                                             // suffix sums over 64 buckets: two per lane
            const int h0 = bhist[2 * lane], h1 = bhist[2 * lane + 1];
            int suf = h0 + h1;                                     // inclusive suffix sum over the lanes above
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_down_sync(FULL, suf, o);
                if (lane + o < 32) suf += t;
            }
            const int above = suf - (h0 + h1);
            bbase[2 * lane + 1] = above;
            bbase[2 * lane] = above + h1;
        }
        team_sync(bar_id, nthr);
        if (pass) {
            const int pos = bbase[bkt] + slot;
            ckey[pos] = conf;
            outpos[pos] = ci;
        }
        team_sync(bar_id, nthr);
        int r = 0;
        if (pass) {
            const int st0 = bbase[bkt], nb = bhist[bkt];
            r = st0;
            for (int m = 0; m < nb; ++m) {
                const float k = ckey[st0 + m];
                const int i2 = outpos[st0 + m];
                r += ((k > conf) || (k == conf && i2 < ci)) ? 1 : 0;
            }
        }
        // ---- C: scatter to rank order (utils.py:24-32,40); class masks
        if (pass) {
            const float xn = __fmul_rn(__fsub_rn(box.x, box.z), 0.5f), xx = __fmul_rn(__fadd_rn(box.x, box.z), 0.5f);
            const float yn = __fmul_rn(__fsub_rn(box.y, box.w), 0.5f), yx = __fmul_rn(__fadd_rn(box.y, box.w), 0.5f);
            scor[r] = make_float4(xn, xx, yn, yx);
            sarea[r] = fabsf(__fmul_rn(__fsub_rn(xx, xn), __fsub_rn(yx, yn)));
            smeta[r] = cls;
        }
        team_sync(bar_id, nthr);
        const bool act = q < n;                                    // rank position q exists
        const int qkey = act ? smeta[q] : 0;
        const unsigned mm = __match_any_sync(FULL, act ? qkey : (0x7f000000 + lane));
        const bool lead = act && (__ffs(mm) - 1) == lane;
        if (lead) tbl[qkey * cc.TW + wt] = mm;
        team_sync(bar_id, nthr);
        // ---- D: suppression words against same-class predecessors (utils.py:108).
        //      Rank position q = 32 wt + lane has predecessors in words 0..wt of its class row, so the last
        //      warp of a team has TW times the work of the first and the team would wait for it.  The words
        //      are therefore re-dealt: warp wt > its mirror pw = TW-1-wt gives its lowest (wt - pw) / 2 words
        //      to warp pw, whose lane l tests them for candidate 32 wt + l and hands the bits back through
        //      shared memory (every warp then walks about (TW+1)/2 words).
        unsigned supp[kTeamWarpsMax];
#pragma unroll
        for (int w2 = 0; w2 < kTeamWarpsMax; ++w2) supp[w2] = 0u;
        const int pw = cc.TW - 1 - wt;                             // mirror warp
        const int give = (wt > pw) ? ((wt - pw) >> 1) : 0;         // words 0..give-1 of this warp are done by warp pw
        const int take = (pw > wt) ? ((pw - wt) >> 1) : 0;         // words 0..take-1 of warp pw are done here
        if (act) {
            const unsigned *rowp = tbl + qkey * cc.TW;
            const float4 qc = scor[q];
            const float qa = sarea[q];
#pragma unroll
            for (int w2 = 0; w2 < kTeamWarpsMax; ++w2) {
                if (w2 <= wt && w2 >= give) {
                    unsigned w = rowp[w2];
                    if (w2 == wt) w &= lt_mask;
                    while (w) {
                        const int b = __ffs(w) - 1;
                        w &= w - 1;
                        supp[w2] |= suppresses(scor[32 * w2 + b], sarea[32 * w2 + b], qc, qa, cfg) << b;
                    }
                }
            }
        }
        if (take > 0) {
            const int q2 = 32 * pw + lane;                         // the mirror warp's candidate of this lane
            if (q2 < n) {
                const unsigned *rowp = tbl + smeta[q2] * cc.TW;
                const float4 qc = scor[q2];
                const float qa = sarea[q2];
                for (int w2 = 0; w2 < take; ++w2) {                // all strictly below warp pw: no lane mask
                    unsigned w = rowp[w2], hs = 0u;
                    while (w) {
                        const int b = __ffs(w) - 1;
                        w &= w - 1;
                        hs |= suppresses(scor[32 * w2 + b], sarea[32 * w2 + b], qc, qa, cfg) << b;
                    }
                    help[q2 * (kTeamWarpsMax / 2) + w2] = hs;
                }
            }
        }
        team_sync(bar_id, nthr);
        if (lead) tbl[qkey * cc.TW + wt] = 0u;                     // leave the table zeroed
        if (act) {
#pragma unroll
            for (int w2 = 0; w2 < kTeamWarpsMax / 2; ++w2)
                if (w2 < give) supp[w2] = help[q * (kTeamWarpsMax / 2) + w2];
        }
        // ---- E: greedy keep flags, fixed point of keep[q] = !any(supp[q] & keep)
        bool alive = act;
        for (;;) {
            const unsigned kb = __ballot_sync(FULL, alive);
            if (lane == 0) kws[wt] = kb;
            team_sync(bar_id, nthr);
            unsigned sgot = 0u;
#pragma unroll
            for (int w2 = 0; w2 < kTeamWarpsMax; ++w2)
                if (w2 <= wt) sgot |= supp[w2] & kws[w2];
            const bool na = act && (sgot == 0u);
            const bool ch = team_any(bar_id, nthr, na != alive);   // also: everybody has read kws
            alive = na;
            if (!ch) break;
        }
        // ---- F: output slots (kws holds the final keep words), rows in pick order (utils.py:112)
        int K = 0, below = 0;
#pragma unroll
        for (int w2 = 0; w2 < kTeamWarpsMax; ++w2) {
            if (w2 < cc.TW) {
                const int k = __popc(kws[w2]);
                if (w2 < wt) below += k;
                K += k;
            }
        }
        if (act) outpos[q] = alive ? below + __popc(kws[wt] & lt_mask) : -1;
        team_sync(bar_id, nthr);
        if (pass) {
            const int pos = outpos[r];
            if (pos >= 0) {
                float2 *o = reinterpret_cast<float2 *>(out_boxes + (img * cfg.M + pos) * 6);
                o[0] = make_float2(static_cast<float>(cls), conf);                    // utils.py:175
                o[1] = make_float2(box.x, box.y);
                o[2] = make_float2(box.z, box.w);
                if (out_idx) out_idx[img * cfg.M + pos] = q;
            }
        }
        if (q == 0) out_count[img] = K;
        // the next image's first barrier separates its shared-memory writes (wcnt) from this image's reads
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

static int pick_ns(int M)
{
    const int need = (M + 31) / 32;
    const int avail[] = {1, 2, 4, 7, 8};
    for (int a : avail)
        if (a >= need) return a;
    return 0;
}

// whether the division-free filter of suppresses() may be used for this threshold
static void set_thr(NmsCfg &cfg)
{
    const float t = cfg.iou_thr;
    cfg.thr_fast = (env_int("YH_EXACT_DIV", 0) == 0 && std::isfinite(t) && t > 1e-30f && t < 1e30f) ? 1 : 0;
    cfg.band = cfg.thr_fast ? 9.5367431640625e-07f : INFINITY;
}

static int fill_cfg(NmsCfg &cfg, int S, int B, int C, float iou_thr, float conf_thr)
{
    YH_REQUIRE(S >= 1 && B >= 1 && C >= 1, "decode/nms: S, B, C must be >= 1 (got %d, %d, %d)", S, B, C);
    if (S * S > YH_MAX_CELLS) {
        set_error("decode/nms: S*S = %d exceeds the compiled limit %d", S * S, YH_MAX_CELLS);
        return YH_ERR_UNSUPPORTED;
    }
    if (C >= (1 << 22)) {
        set_error("decode/nms: C = %d too large", C);
        return YH_ERR_UNSUPPORTED;
    }
    cfg.S = S; cfg.B = B; cfg.C = C; cfg.M = S * S; cfg.D = C + 5 * B;
    cfg.inv_s = static_cast<float>(1.0 / static_cast<double>(S));
    cfg.iou_thr = iou_thr; cfg.conf_thr = conf_thr;
    cfg.ws_bytes = 0; cfg.tbl_rows = C; cfg.score_mode = 0;
    set_thr(cfg);
    return YH_OK;
}

template <int NS, int CT, int BT, typename E>
static int launch_direct(const E *pred, int64_t n, NmsCfg cfg, float *out_boxes, int *out_count, int *out_idx,
                         cudaStream_t st)
{
    cfg.ws_bytes = WarpWs<NS, false>::bytes(cfg.tbl_rows);
    const size_t per_warp = static_cast<size_t>(cfg.ws_bytes);
    if (per_warp > 220 * 1024) {
        set_error("decode_nms: per-warp workspace %zu B does not fit shared memory (C = %d too large)", per_warp, cfg.C);
        return YH_ERR_UNSUPPORTED;
    }
    // as many warps as fit one SM, in two blocks when there are enough of them
    int warps_sm = static_cast<int>(std::min<size_t>(32, (224 * 1024) / per_warp));
    int wpb = warps_sm >= 8 ? std::min(16, warps_sm / 2) : warps_sm;
    if (wpb < 1) wpb = 1;
    const size_t smem = static_cast<size_t>(wpb) * per_warp;
    auto kern = decode_nms_direct_kernel<NS, CT, BT, E>;
    YH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 1;
    YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t want = (n + wpb - 1) / wpb;
    const int grid = static_cast<int>(std::min<int64_t>(want, static_cast<int64_t>(sm_count()) * per_sm));
    kern<<<grid, wpb * 32, smem, st>>>(pred, n, cfg, out_boxes, out_count, out_idx);
    YH_LAUNCH_CHECK("decode_nms_direct_kernel");
    return YH_OK;
}

// Big images, cooperative variant (decode_nms_coop_kernel): teams of ceil(M/32) warps per image.
template <int CT, int BT, typename E>
static int launch_coop(const E *pred, int64_t n, NmsCfg cfg, float *out_boxes, int *out_count, int *out_idx,
                       cudaStream_t st, int64_t *done)
{
    *done = 0;
    const int64_t img_bytes = static_cast<int64_t>(sizeof(E)) * cfg.M * cfg.D;
    if (env_int("YH_COOP", 1) == 0 || reinterpret_cast<uintptr_t>(pred) % 16 != 0 || img_bytes % 4 != 0) return YH_OK;
    CoopCfg cc;
    cc.TW = (cfg.M + 31) / 32;
    if (cc.TW > kTeamWarpsMax) return YH_OK;
    cc.chunk_bytes = 32u * cfg.D * static_cast<uint32_t>(sizeof(E));
    cc.last_bytes = static_cast<uint32_t>(img_bytes - static_cast<int64_t>(cc.TW - 1) * cc.chunk_bytes);
    if (cc.chunk_bytes % 16 != 0) return YH_OK;
    // Images that are not a multiple of 16 bytes (odd cell / channel counts, 16-bit heads): bulk copies need 16-byte
    // aligned ranges, so each chunk is fetched as the aligned range around it (at most 15 bytes of its neighbours on
    // either side, all inside the tensor) and read at its offset inside a slot that is 16 bytes longer.  Only the very
    // last chunk of the tensor could reach past its end: if that end is unaligned, the last image is left to the caller.
    cc.shifted = (img_bytes % 16 != 0) ? 1 : 0;
    cc.slot_bytes = cc.chunk_bytes + (cc.shifted ? 16u : 0u);
    const int64_t n_coop = (cc.shifted && (n * img_bytes) % 16 != 0) ? n - 1 : n;
    if (n_coop < 1 || (cc.shifted && env_int("YH_COOP_SHIFTED", 1) == 0)) return YH_OK;
    const int MPT = cc.TW * 32;
    cc.team_bytes = (MPT * 16 + MPT * 4 + MPT * 4 + (MPT + 4) * 4 + MPT * 4 + 2 * kTeamWarpsMax * 4 + MPT * (kTeamWarpsMax / 2) * 4 +
                     2 * kRankBuckets * 4 +
                     cfg.C * cc.TW * 4 + 15) & ~15;
    cc.NTEAM = std::max(1, std::min(std::min(15, 31 / cc.TW), env_int("YH_COOP_TEAMS", 4)));
    cc.ST = std::max(2, std::min(32, env_int("YH_COOP_STAGES", 10)));
    auto need = [&](int teams, int stg) {
        return static_cast<size_t>(stg) * cc.slot_bytes + 2 * static_cast<size_t>(stg) * 8 + 16 + static_cast<size_t>(teams) * cc.team_bytes + 128;
    };
    while (cc.ST > 4 && need(cc.NTEAM, cc.ST) > 227 * 1024) --cc.ST;
    while (cc.NTEAM > 1 && need(cc.NTEAM, cc.ST) > 227 * 1024) --cc.NTEAM;
    while (cc.ST > 2 && need(cc.NTEAM, cc.ST) > 227 * 1024) --cc.ST;
    if (need(cc.NTEAM, cc.ST) > 227 * 1024) return YH_OK;
    cc.n = n_coop;
    const size_t smem = need(cc.NTEAM, cc.ST);
    auto kern = decode_nms_coop_kernel<CT, BT, E>;
    YH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int grid = static_cast<int>(std::min<int64_t>(n_coop, sm_count()));
    kern<<<grid, 32 * (1 + cc.NTEAM * cc.TW), smem, st>>>(pred, cfg, cc, out_boxes, out_count, out_idx);
    YH_LAUNCH_CHECK("decode_nms_coop_kernel");
    *done = n_coop;
    return YH_OK;
}

template <int NS, int CT, int BT, typename E>
static int launch_fused(const E *pred, int64_t n, NmsCfg cfg, float *out_boxes, int *out_count, int *out_idx,
                        cudaStream_t st)
{
    const int64_t img_bytes = static_cast<int64_t>(sizeof(E)) * cfg.M * cfg.D;
    int64_t done = 0;
    // ---- TMA ring for the aligned bulk ----
    const bool tma_on = env_int("YH_TMA", 1) != 0;
    // Images above this size go to the cooperative team kernel instead of the tile ring: 12 KB in general (the ring would
    // hold too few images), 6 KB for grids of more than 64 cells (NS > 2: long per-warp chains and register spills in the
    // one-warp-per-image kernel; measured 2-3x in favour of the team kernel on S = 9..16, profiles/prof_midsize.py)
    const int64_t coop_min = env_int("YH_COOP_MIN_BYTES", NS > 2 ? 6 * 1024 : 12 * 1024);
    if (tma_on && (reinterpret_cast<uintptr_t>(pred) % 16 == 0) && img_bytes <= 12 * 1024 && img_bytes <= coop_min) {
        TmaCfg tc;
        tc.W = env_int("YH_TMA_W", 24);
        tc.T = env_int("YH_TMA_T", 4);
        if (tc.W < 1 || tc.W > 26) tc.W = 24;
        if (tc.T < 1) tc.T = 8;
        while (tc.T > 1 && ((tc.W % tc.T) != 0 || tc.T * img_bytes > 64 * 1024)) tc.T >>= 1;
        if ((tc.W % tc.T) == 0 && (tc.T * img_bytes) % 16 == 0 && n >= tc.T) {
            tc.tile_bytes = static_cast<uint32_t>(tc.T * img_bytes);
            cfg.ws_bytes = WarpWs<NS, false>::bytes(cfg.tbl_rows);
            // Every stage must always be consumed by the same warp group, so that a group's waits on
            // a stage's mbarrier are strictly sequential (parity waits are only unambiguous one
            // phase apart): the stage count is a multiple of the group count G = W / T.
            int G = tc.W / tc.T;
            size_t fixed = 2 * 16 * 8 /*barriers*/ + 16 + static_cast<size_t>(tc.W) * cfg.ws_bytes + 128;
            int cap = std::min(16, env_int("YH_TMA_STAGES", 16));
            auto fit = [&](size_t fx) { return fx >= 227 * 1024 ? 0 : static_cast<int>((227 * 1024 - fx) / tc.tile_bytes); };
            int stages = (std::min(cap, fit(fixed)) / G) * G;
            if (stages < G || stages < 2) {             // not enough room: one group only
                tc.W = tc.T;
                G = 1;
                fixed = 2 * 16 * 8 + 16 + static_cast<size_t>(tc.W) * cfg.ws_bytes + 128;
                stages = std::min(cap, fit(fixed));
            }
            tc.ST = stages;
            tc.n_tiles = n / tc.T;
            const size_t smem = fixed + static_cast<size_t>(tc.ST) * tc.tile_bytes;
            if (tc.ST >= 2 && smem <= 227 * 1024) {
                auto kern = decode_nms_tma_kernel<NS, CT, BT, E>;
                YH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                const int grid = static_cast<int>(std::min<int64_t>(tc.n_tiles, sm_count()));
                kern<<<grid, 32 * (tc.W + 1), smem, st>>>(pred, cfg, tc, out_boxes, out_count, out_idx);
                YH_LAUNCH_CHECK("decode_nms_tma_kernel");
                done = tc.n_tiles * tc.T;
            }
        }
    }
    // ---- images too large for the tile ring: cooperative team kernel ----
    if (done == 0 && (img_bytes > 12 * 1024 || img_bytes > coop_min)) {
        const int rc = launch_coop<CT, BT, E>(pred, n, cfg, out_boxes, out_count, out_idx, st, &done);
        if (rc != YH_OK) return rc;
    }
    // ---- tail / fallback ----
    if (done < n) {
        return launch_direct<NS, CT, BT, E>(pred + done * cfg.M * cfg.D, n - done, cfg, out_boxes + done * cfg.M * 6,
                                         out_count + done, out_idx ? out_idx + done * cfg.M : nullptr, st);
    }
    return YH_OK;
}

// Fused decode + NMS for a prediction tensor of element type E (float, __half, __nv_bfloat16).
template <typename E>
static int decode_nms_typed(const E *pred, int64_t n, int S, int B, int C, float iou_thr, float conf_thr,
                            float *out_boxes, int32_t *out_count, int32_t *out_keep_idx, cudaStream_t st, int score_mode)
{
    NmsCfg cfg;
    int rc = fill_cfg(cfg, S, B, C, iou_thr, conf_thr);
    if (rc != YH_OK) return rc;
    YH_REQUIRE(score_mode == YH_SCORE_CONF || score_mode == YH_SCORE_CONF_X_PROB, "decode_nms: unknown score_mode %d", score_mode);
    cfg.score_mode = score_mode;
    YH_REQUIRE(n >= 0, "decode_nms: n < 0");
    if (n == 0) return YH_OK;
    YH_REQUIRE(pred && out_boxes && out_count, "decode_nms: null pointer");
    YH_REQUIRE(reinterpret_cast<uintptr_t>(pred) % (2 * sizeof(E)) == 0 && reinterpret_cast<uintptr_t>(out_boxes) % 8 == 0,
               "decode_nms: pred must be aligned to two elements and out_boxes to 8 bytes");
    const int ns = pick_ns(cfg.M);
    if (ns == 2 && C == 20 && B == 2) return launch_fused<2, 20, 2>(pred, n, cfg, out_boxes, out_count, out_keep_idx, st);
    if (ns == 7 && C == 80 && B == 3 && env_int("YH_SPECIAL", 1) != 0) return launch_fused<7, 80, 3>(pred, n, cfg, out_boxes, out_count, out_keep_idx, st);
    switch (ns) {
        case 1: return launch_fused<1, 0, 0>(pred, n, cfg, out_boxes, out_count, out_keep_idx, st);
        case 2: return launch_fused<2, 0, 0>(pred, n, cfg, out_boxes, out_count, out_keep_idx, st);
        case 4: return launch_fused<4, 0, 0>(pred, n, cfg, out_boxes, out_count, out_keep_idx, st);
        case 7: return launch_fused<7, 0, 0>(pred, n, cfg, out_boxes, out_count, out_keep_idx, st);
        case 8: return launch_fused<8, 0, 0>(pred, n, cfg, out_boxes, out_count, out_keep_idx, st);
    }
    set_error("decode_nms: unsupported cell count %d", cfg.M);
    return YH_ERR_UNSUPPORTED;
}

}  // namespace yh
