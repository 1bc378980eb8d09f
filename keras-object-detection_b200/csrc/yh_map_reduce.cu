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
#include <cstddef>
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
    long long seg_off[kMaxSegs + 1];      // (first: shared with the counting path, whose tables overlay everything from `stage` on)
    unsigned cta_tp, pad_;
    unsigned long long stage[RS_TILE];    // tile in digit order
    uint32_t wc[RS_W][RS_BINS];           // per-warp digit counters -> exclusive warp offsets
    uint32_t tile_cnt[RS_BINS];
    uint32_t tile_start[RS_BINS];
    uint32_t dig_off[RS_BINS];            // next output position of this CTA per digit
    uint32_t hist[RS_BINS];
    uint32_t wsum[RS_W];
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

// np.trapz term of one TRUE positive (utils.py:430-444) from its position in the class (pos, 1-based: detections of the
// class up to and including it), the true positives up to and including it (tpi) and the class's ground truths, as a
// 2^-56 fixed-point integer.  The term depends on nothing else - which is what lets the counting path below skip the sort.
__device__ __forceinline__ unsigned long long ap_term_fixed(int tpi, int pos, int ngt)
{
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
    return static_cast<unsigned long long>(static_cast<double>(term) * 72057594037927936.0);  // 2^56
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

// ---------------------------------------------------------------------------------------------------------------
// Counting path (small and medium inputs): AP without a sort.
// The np.trapz term of a true positive needs only its position inside the class and the number of true positives
// up to it (ap_term_fixed) - two RANKS, and a rank is a count: pos - 1 = #{same class : (confidence desc, row asc)
// before it}.  So instead of five radix passes (ten grid barriers of dependent L2 round trips - at cfg4's 73,595
// records the sort was 76 us of pure latency) the records are partitioned ONCE, unordered, into bins
// (class, confidence bucket, tp) - buckets are a monotone map of the confidence, a higher bucket always ranks first -
// and every true positive counts, by brute force over its own (class, bucket) group in shared memory, the members that
// precede it: 64-bit keys  ~orderable(conf) << 32 | row  compare exactly like the stable sort's order, so ties need
// no extra pass.  The members of higher buckets are added from the bin totals.  Work = sum over groups of
// TPs x members pairs at 4 instructions each: the kernel decides ON THE DEVICE, after the histogram, whether that is
// cheaper than sorting (a.count_pairs_max; skewed classes or huge inputs fall through to the radix passes).
// Three grid barriers instead of thirteen; the fixed-point accumulation makes the result bit-identical to the radix
// path's whatever the order of the adds.
// ---------------------------------------------------------------------------------------------------------------
constexpr int CT_SB = 4;                  // independent sub-blocks of a CTA in the counting phase (named barriers 1..4)
constexpr int CT_SBW = RS_W / CT_SB;      // warps per sub-block = key slices of a unit
constexpr int CT_SBT = 32 * CT_SBW;       // threads per sub-block
constexpr int CT_CHUNK = 1024;            // keys per shared-memory chunk of a sub-block
constexpr int CT_GROUPS = 512;            // (class, bucket) groups: C * buckets <= 512
constexpr int CT_BINS = 2 * CT_GROUPS;    // bin = 2 * group + (tp ? 0 : 1): a group's true positives come first
constexpr int CT_ROWS = 2 * 160 * RS_BINS / CT_BINS;   // rows of the count matrix that fit the radix hist region at full width

struct CountSmem {
    unsigned long long keys[CT_SB][CT_CHUNK];
    uint32_t bin_start[CT_BINS + 2];      // start of every bin in the partitioned array (+ total)
    uint32_t cur[CT_BINS];                // histogram of this CTA -> its cursors
    uint32_t tp_before[CT_GROUPS];        // true positives of the class in the groups before this one
    uint32_t unit_start[CT_GROUPS + 1];   // first unit of work of every group
    uint32_t red[RS_T][2];                // (true positives, false positives) before, per thread of a unit
    uint32_t wsum[3][RS_W];
    unsigned long long wpairs[RS_W];
    int fallback;
};
static_assert(sizeof(CountSmem) <= sizeof(RadixSmem) - offsetof(RadixSmem, stage), "the counting path's tables overlay the radix buffers from `stage` on");
static_assert(RS_W % CT_SB == 0 && CT_GROUPS <= RS_T, "counting path geometry");

// bin of a record, -1 for rows the reference never selects (class outside [0, C)).  nb = 2^nb_log2 buckets over
// [0, 1]: bucket = floor(conf * nb), everything >= 1 (also +inf / +NaN, which sort first) in the top one, everything
// <= 0 (also negatives / -NaN, which sort last) in the bottom one - monotone in the sort key by construction.
__device__ __forceinline__ int count_bin(unsigned long long rec, int C, int nb_log2)
{
    const uint32_t cls = static_cast<uint32_t>(rec >> kRecClassShift);
    if (cls >= static_cast<uint32_t>(C)) return -1;
    const uint32_t o = ~static_cast<uint32_t>(rec >> 1);                 // orderable(conf)
    const int nb = 1 << nb_log2;
    int bucket;
    if (o >= 0xbf800000u) bucket = nb - 1;                               // conf >= 1
    else if (o <= 0x80000000u) bucket = 0;                               // conf <= 0
    else bucket = min(nb - 1, static_cast<int>(__fmul_rn(__uint_as_float(o & 0x7fffffffu), static_cast<float>(nb))));
    return (((static_cast<int>(cls) << nb_log2) + (nb - 1 - bucket)) << 1) + ((rec & 1ull) ? 0 : 1);
}

// exclusive scans of three values over the CTA's 512 threads at once; totals in tot[3]
__device__ __forceinline__ void scan512x3(uint32_t (&v)[3], uint32_t (&tot)[3], CountSmem &cs)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        x[k] = v[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, x[k], o);
            if (lane >= o) x[k] += t;
        }
        if (lane == 31) cs.wsum[k][warp] = x[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint32_t off = 0, t = 0;
#pragma unroll
        for (int w = 0; w < RS_W; ++w) {
            const uint32_t s = cs.wsum[k][w];
            if (w < warp) off += s;
            t += s;
        }
        tot[k] = t;
        v[k] = off + x[k] - v[k];
    }
    __syncthreads();
}

// number of keys[s .. e) below `key` (every lane reads the same address: a shared-memory broadcast).  Four independent
// counters over 16-byte loads: 3.5 instructions per key (LDS.128 for two, then ISETP.LT.U32 / ISETP.LT.U32.EX / IADD each).
__device__ __forceinline__ uint32_t count_before(const unsigned long long *keys, int s, int e, unsigned long long key)
{
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    int j = s;
    if ((j & 1) && j < e) { c0 += keys[j] < key ? 1u : 0u; ++j; }          // align to 16 bytes
    const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(keys + j);
    const int pairs = (e - j) >> 1;
    int q = 0;
    for (; q + 2 <= pairs; q += 2) {
        const ulonglong2 a = k2[q], b = k2[q + 1];
        c0 += a.x < key ? 1u : 0u;
        c1 += a.y < key ? 1u : 0u;
        c2 += b.x < key ? 1u : 0u;
        c3 += b.y < key ? 1u : 0u;
    }
    if (q < pairs) {
        const ulonglong2 a = k2[q];
        c0 += a.x < key ? 1u : 0u;
        c1 += a.y < key ? 1u : 0u;
        ++q;
    }
    j += 2 * pairs;
    if (j < e) c2 += keys[j] < key ? 1u : 0u;
    return (c0 + c1) + (c2 + c3);
}

// -DYH_MAP_TIMELINE: CTA 0 stamps %globaltimer at the phase boundaries of the counting path and the host prints them
// (YH_MAP_DBG=1) - how profiles/README.md's phase timeline was taken.  Compiled out otherwise.
__device__ __forceinline__ void dbg_stamp(const ReduceArgs &a, int k)
{
#ifdef YH_MAP_TIMELINE
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t) :: "memory");
        a.dbg[k] = t;
    }
#endif
}
// Returns false when sorting is cheaper (nothing has been written but the count matrix): the caller runs the radix passes.
__device__ __noinline__ bool count_path(const ReduceArgs &a, RadixSmem &sm, CountSmem &cs, const int32_t *gt_s, long long n,
                                        cg::grid_group &grid)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;
    // buckets per class: as many (up to the host's limit) as keep a (class, bucket) group near 256 members on average -
    // fewer bins mean a smaller count matrix for everybody to read, more bins mean fewer pairs; n is the device-side count
    int nbl = 0;
    while (nbl < a.count_nb_log2 && ((n / a.C) >> nbl) > 256) ++nbl;
    const int nb = 1 << nbl;
    const int ngroups = a.C << nbl, nbins = 2 * ngroups;
    // the partition is done by as many CTAs as there are 4,096-record pieces (their count matrix is read by everybody)
    const int Gp = static_cast<int>(max(1ll, min(static_cast<long long>(min(G, CT_ROWS)), (n + RS_TILE - 1) / RS_TILE)));
    const long long plo = cta < Gp ? n * cta / Gp : 0, phi = cta < Gp ? n * (cta + 1) / Gp : 0;

    dbg_stamp(a, 1);
    // ---- 1. histogram of this CTA's piece over the bins
    for (int b = tid; b < nbins; b += RS_T) cs.cur[b] = 0;
    __syncthreads();
    dbg_stamp(a, 10);
    // (plain shared-memory atomics: with hundreds of bins nearly every lane of a warp holds a different one, and MATCH.ANY
    // takes one round per distinct value - warp aggregation cost 0.5 us per record and warp here)
    unsigned long long r[8];                                             // the piece's last batch stays in registers for step 3
    long long i_last = plo;
    for (long long i0 = plo;; i0 += RS_T * 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {                                    // all loads of the thread in flight together
            const long long i = i0 + j * RS_T + tid;
            r[j] = i < phi ? load_in(a, sm, nullptr, i) : ~0ull;         // class field all ones: no bin
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int bin = count_bin(r[j], a.C, nbl);
            if (bin >= 0) atomicAdd(&cs.cur[bin], 1u);
        }
        i_last = i0;
        if (i0 + RS_T * 8 >= phi) break;
    }
    dbg_stamp(a, 13);
    __syncthreads();
    if (cta < Gp)
        for (int b = tid; b < nbins; b += RS_T) a.hist[static_cast<size_t>(cta) * nbins + b] = cs.cur[b];
    dbg_stamp(a, 2);
    grid.sync();
    dbg_stamp(a, 3);

    // ---- 2. totals per bin, this CTA's offset inside every bin, starts of the bins, units of work, pair count
    for (int b = tid; b < nbins; b += RS_T) {
        cs.bin_start[b] = 0;
        cs.cur[b] = 0;
    }
    __syncthreads();
    {
        // the Gp x nbins matrix read flat (coalesced, 16 loads in flight per thread: one L2 round trip per 8 K entries)
        const int total = Gp * nbins, step_row = RS_T / nbins, step_bin = RS_T % nbins;
        int row = tid / nbins, bin = tid % nbins;
        for (int e0 = tid; e0 < total; e0 += 16 * RS_T) {
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (e0 + j * RS_T < total) ? __ldcg(a.hist + e0 + j * RS_T) : 0u;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (v[j]) {
                    atomicAdd(&cs.bin_start[bin], v[j]);
                    if (row < cta) atomicAdd(&cs.cur[bin], v[j]);
                }
                bin += step_bin;
                row += step_row;
                if (bin >= nbins) { bin -= nbins; ++row; }
            }
        }
    }
    __syncthreads();
    dbg_stamp(a, 15);
    {
        const bool ing = tid < ngroups;                                  // thread = group
        const uint32_t vT = ing ? cs.bin_start[2 * tid] : 0u, vF = ing ? cs.bin_start[2 * tid + 1] : 0u;
        unsigned long long pairs = static_cast<unsigned long long>(vT) * (vT + vF);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
        if (lane == 0) cs.wpairs[warp] = pairs;
        uint32_t v[3] = {vT + vF, vT, (vT + 31) / 32}, tot[3];
        scan512x3(v, tot, cs);                                           // (syncs: wpairs visible, bin_start reads done)
        if (ing) {
            cs.bin_start[2 * tid] = v[0];
            cs.bin_start[2 * tid + 1] = v[0] + vT;
            cs.tp_before[tid] = v[1];
            cs.unit_start[tid] = v[2];
        }
        if (tid == 0) {
            cs.bin_start[nbins] = tot[0];
            cs.unit_start[ngroups] = tot[2];
            unsigned long long p = 0;
            for (int w = 0; w < RS_W; ++w) p += cs.wpairs[w];
            cs.fallback = p > a.count_pairs_max + a.count_pairs_per_rec * static_cast<unsigned long long>(n) ? 1 : 0;
        }
        __syncthreads();
        const uint32_t cls_base = ing ? cs.tp_before[tid & ~(nb - 1)] : 0u;   // first group of the class
        __syncthreads();
        if (ing) cs.tp_before[tid] -= cls_base;
    }
    if (cs.fallback) {                                                   // same totals on every CTA: uniform
        grid.sync();                                                     // the count matrix shares its memory with the radix histograms
        return false;
    }
    dbg_stamp(a, 4);
    // ---- 3. partition: keys (~orderable(conf) << 32 | row) into their bins, any order inside a bin
    unsigned long long *comp = a.buf[0];
    if (cta < Gp) {
        for (int b = tid; b < nbins; b += RS_T) cs.cur[b] += cs.bin_start[b];
        __syncthreads();
        // the last batch first (still in registers), then the earlier ones (re-read: they are in L2)
        for (long long i0 = i_last;;) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int bin = count_bin(r[j], a.C, nbl);
                if (bin >= 0)
                    comp[atomicAdd(&cs.cur[bin], 1u)] = (static_cast<unsigned long long>(static_cast<uint32_t>(r[j] >> 1)) << 32) |
                                                        static_cast<uint32_t>(i0 + j * RS_T + tid);
            }
            if (i0 == plo) break;
            i0 -= RS_T * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long i = i0 + j * RS_T + tid;
                r[j] = i < phi ? load_in(a, sm, nullptr, i) : ~0ull;
            }
        }
    }
    dbg_stamp(a, 5);
    grid.sync();
    dbg_stamp(a, 6);

    // ---- 4. units of work = (group, 32 of its true positives): count the group's members that precede each of them.
    //      A CTA runs four units at a time, one per sub-block of four warps (named barriers): lane = true positive, the
    //      four warps split the group's keys (staged in shared memory, every read a broadcast) - a unit is a couple of
    //      dependent L2 round trips and a short scan, so what counts is how many are in flight, not how wide one is.
    const uint32_t n_units = cs.unit_start[ngroups];
    const int sb = warp / CT_SBW, w4 = warp % CT_SBW, t_sb = tid % CT_SBT;
    unsigned long long *keys = cs.keys[sb];
    for (uint32_t u = static_cast<uint32_t>(sb) * G + cta; u < n_units; u += CT_SB * G) {
        int gl = 0, gh = ngroups;                                        // last group with unit_start <= u
        while (gh - gl > 1) {
            const int mid = (gl + gh) >> 1;
            if (cs.unit_start[mid] <= u) gl = mid; else gh = mid;
        }
        const int g = gl;
        const uint32_t k_base = cs.bin_start[2 * g], mT = cs.bin_start[2 * g + 1] - k_base;
        const uint32_t mk = cs.bin_start[2 * g + 2] - k_base;            // true positives first, then false positives
        const uint32_t t_off = (u - cs.unit_start[g]) * 32;
        const int nt = static_cast<int>(min(32u, mT - t_off));
        const unsigned long long mykey = lane < nt ? __ldcg(comp + k_base + t_off + lane) : 0ull;
        uint32_t cT = 0, cF = 0;
        for (uint32_t k0 = 0; k0 < mk; k0 += CT_CHUNK) {
            const int len = static_cast<int>(min(static_cast<uint32_t>(CT_CHUNK), mk - k0));
#pragma unroll
            for (int q = 0; q < CT_CHUNK / CT_SBT; ++q) {                // all loads of the thread in flight together
                const int j = q * CT_SBT + t_sb;
                if (j < len) keys[j] = __ldcg(comp + k_base + k0 + j);
            }
            asm volatile("bar.sync %0, %1;" ::"r"(1 + sb), "r"(CT_SBT) : "memory");
            const int L = (len + CT_SBW - 1) / CT_SBW;
            const int s0 = min(w4 * L, len), e = min(s0 + L, len);
            const int mid = static_cast<int>(min(static_cast<long long>(e), max(static_cast<long long>(s0), static_cast<long long>(mT) - static_cast<long long>(k0))));
            cT += count_before(keys, s0, mid, mykey);
            cF += count_before(keys, mid, e, mykey);
            asm volatile("bar.sync %0, %1;" ::"r"(1 + sb), "r"(CT_SBT) : "memory");
        }
        cs.red[tid][0] = cT;
        cs.red[tid][1] = cF;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + sb), "r"(CT_SBT) : "memory");
        if (w4 == 0) {
            unsigned long long fx = 0;
            const int cls = g >> nbl;
            const int ngt = gt_s[cls];
            if (lane < nt && ngt > 0) {                                  // utils.py:334-336
                uint32_t sT = 0, sF = 0;
#pragma unroll
                for (int q = 0; q < CT_SBW; ++q) { sT += cs.red[tid + 32 * q][0]; sF += cs.red[tid + 32 * q][1]; }
                const int tpi = static_cast<int>(cs.tp_before[g] + sT) + 1;
                const int pos = static_cast<int>(k_base - cs.bin_start[2 * (cls << nbl)] + sT + sF) + 1;
                fx = ap_term_fixed(tpi, pos, ngt);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) fx += __shfl_xor_sync(0xffffffffu, fx, o);
            if (lane == 0 && fx) atomicAdd(&a.apfix[cls], fx);
        }
        // (no barrier here: the next unit's first barrier comes before anything of this one is overwritten - red[] by
        // the other warps only after two more barriers, keys[] only after every warp has left the scan above)
    }
    __syncthreads();
    dbg_stamp(a, 7);
    grid.sync();
    dbg_stamp(a, 8);
    return true;
}

// AP per class and the mean over ALL classes (utils.py:456) from the fixed-point sums; CTA 0, after a grid barrier.
__device__ __noinline__ void finish_ap(const ReduceArgs &a, RadixSmem &sm, const int32_t *gt_s, long long n)
{
    const int tid = threadIdx.x, cta = blockIdx.x, C = a.C;
    if (cta == 0) {
        double *red = reinterpret_cast<double *>(sm.stage);
        double acc = 0.0;
        for (int c = tid; c < C; c += RS_T) {
            const float ap = gt_s[c] > 0 ? static_cast<float>(static_cast<double>(a.apfix[c]) * (1.0 / 72057594037927936.0)) : 0.0f;
            if (a.out_ap) a.out_ap[c] = ap;
            acc += static_cast<double>(ap);
        }
        // lanes -> warp (shuffles) -> 16 warp sums added in warp order by thread 0: a fixed order, one barrier
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < RS_W; ++w) tot += red[w];
            float m = static_cast<float>(tot / static_cast<double>(C));
            if (a.err && *reinterpret_cast<volatile int32_t *>(a.err) != 0) m = __int_as_float(0x7fc00000);   // failed exchange: NaN, not a number that looks right
            *a.out_map = m;
            if (a.out_n) *a.out_n = n;
        }
    }
}

__global__ void __launch_bounds__(RS_T, 2) map_radix_kernel(const __grid_constant__ ReduceArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RadixSmem &sm = *reinterpret_cast<RadixSmem *>(smem_raw);
    unsigned long long *ap_s = reinterpret_cast<unsigned long long *>(smem_raw + sizeof(RadixSmem));   // [C+1] fixed-point AP sums
    uint32_t *tp_s = reinterpret_cast<uint32_t *>(ap_s + (a.C + 1));                                   // [C+1] TP counts
    int32_t *gt_s = reinterpret_cast<int32_t *>(tp_s + (a.C + 1));                                     // [C+1] ground truths per class
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;

    dbg_stamp(a, 0);
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
    if (a.mode == YH_RADIX_AP) {                          // (loads in flight together with the counts below)
        for (int c = tid; c <= a.C; c += RS_T) {
            int g = 0;
            if (c < a.C)
                for (int s = 0; s < a.n_gt; ++s) g += a.gt_part[s][c];            // the shards' counts, summed here
            gt_s[c] = g;
        }
    }
    if (tid == RS_T - 1) {
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

    const int C = a.C;
    if (a.mode == YH_RADIX_AP) {
        if (a.count_nb_log2 >= 0 && n <= a.count_n_max && count_path(a, sm, *reinterpret_cast<CountSmem *>(smem_raw + offsetof(RadixSmem, stage)), gt_s, n, grid)) {
            finish_ap(a, sm, gt_s, n);
            dbg_stamp(a, 9);
            return;
        }
    }

    const unsigned long long *src = nullptr;              // pass 0 reads the segments
    unsigned long long *dst = a.buf[0];
    for (int p = 0; p < a.npass; ++p) {
        const int shift = a.bit_lo + 8 * p;
        // ---- histogram of this CTA's range
        if (tid < RS_BINS) sm.hist[tid] = 0;
        __syncthreads();
        for (long long i0 = lo; i0 < hi; i0 += RS_T * RS_IPT) {
            unsigned long long kk[RS_IPT];
#pragma unroll
            for (int j = 0; j < RS_IPT; ++j) {                   // all loads of the thread in flight together (one load per
                const long long i = i0 + j * RS_T + tid;         // iteration left this phase at 1.2 TB/s: 15 % of the samples)
                kk[j] = i < hi ? load_in(a, sm, src, i) : 0ull;
            }
#pragma unroll
            for (int j = 0; j < RS_IPT; ++j) {
                if (i0 + j * RS_T + (tid & ~31) < hi) {          // warp-uniform
                    const bool valid = i0 + j * RS_T + tid < hi;
                    const uint32_t d = valid ? static_cast<uint32_t>(kk[j] >> shift) & 0xffu : 256u + lane;
                    const uint32_t peers = __match_any_sync(0xffffffffu, d);
                    if (valid && lane == __ffs(peers) - 1) atomicAdd(&sm.hist[d], __popc(peers));
                }
            }
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
    for (int c = tid; c <= C; c += RS_T) {
        ap_s[c] = 0;
        tp_s[c] = 0;
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
    for (long long i0 = lo; i0 < hi; i0 += RS_T * RS_IPT) {
        unsigned long long kk[RS_IPT];
#pragma unroll
        for (int j = 0; j < RS_IPT; ++j) {                       // all loads of the thread in flight together
            const long long i = i0 + j * RS_T + tid;
            kk[j] = i < hi ? src[i] : 0ull;                      // (an absent record is no true positive)
        }
#pragma unroll
        for (int j = 0; j < RS_IPT; ++j) {
            if (i0 + j * RS_T + (tid & ~31) < hi) {              // warp-uniform
                const bool tp = kk[j] & 1ull;
                const uint32_t c = tp ? static_cast<uint32_t>(kk[j] >> kRecClassShift) : 0x10000u + lane;
                const uint32_t peers = __match_any_sync(0xffffffffu, c);
                if (tp && lane == __ffs(peers) - 1) {
                    atomicAdd(&tp_s[c], __popc(peers));
                    atomicAdd(&sm.cta_tp, __popc(peers));
                }
            }
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
    for (long long i00 = lo; i00 < hi; i00 += 4 * RS_T) {
      unsigned long long kq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {                              // four iterations' loads in flight together
          const long long i = i00 + q * RS_T + tid;
          kq[q] = i < hi ? src[i] : 0ull;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long i0 = i00 + q * RS_T;
        if (i0 >= hi) break;                                     // CTA-uniform
        const long long i = i0 + tid;
        const bool valid = i < hi;
        const unsigned long long key = kq[q];
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
                fx = ap_term_fixed(tpi, pos, ngt);
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
    }
    __syncthreads();
    for (int c = tid; c < C; c += RS_T)
        if (ap_s[c]) atomicAdd(&a.apfix[c], ap_s[c]);
    grid.sync();

    // ---- epilogue 3: AP per class and the mean over ALL classes (utils.py:456)
    finish_ap(a, sm, gt_s, n);
}

static int radix_grid(int64_t n_want, size_t smem, int &G, bool counting)
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
    if (counting) {
        // the counting path spreads small units of work (a (class, bucket) group x 128 true positives) and has only three
        // grid barriers: one CTA per SM even when the records themselves would fit a few tiles
        static const int env_cg = [] { const char *v = getenv("YH_MAP_COUNT_GRID"); return (v && *v) ? atoi(v) : 0; }();
        G = std::max(G, std::min(cap, env_cg > 0 ? env_cg : sm_count()));
    }
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
    // counting path: up to 16 confidence buckets per class while C * buckets <= CT_GROUPS; YH_MAP_COUNT=0 turns it off,
    // YH_MAP_COUNT_NMAX / YH_MAP_COUNT_PAIRS move the limits (read per call: the tests switch them)
    a.count_nb_log2 = -1;
    a.count_n_max = 0;
    a.count_pairs_max = 0;
    a.count_pairs_per_rec = 0;
    if (a.mode == YH_RADIX_AP && a.C <= CT_GROUPS) {
        const char *v = getenv("YH_MAP_COUNT");
        if (!(v && *v && atoi(v) == 0)) {
            int l = 4;
            while (l > 0 && (a.C << l) > CT_GROUPS) --l;
            if (const char *b = getenv("YH_MAP_COUNT_BUCKETS")) {
                int want = atoi(b), wl = 0;
                while ((2 << wl) <= want) ++wl;
                l = std::min(l, std::max(0, wl));
            }
            a.count_nb_log2 = l;
            const char *nm = getenv("YH_MAP_COUNT_NMAX"), *pm = getenv("YH_MAP_COUNT_PAIRS");
            a.count_n_max = (nm && *nm) ? atoll(nm) : (1ll << 20);
            // measured on B200 (profiles/README.md): the radix passes cost about 68 us + 0.097 us per 1,000 records, the counting
            // path about 20 us + 0.33 us per million pairs -> counting wins below 145 M + 290 n pairs
            a.count_pairs_max = (pm && *pm) ? strtoull(pm, nullptr, 10) : 145000000ull;
            a.count_pairs_per_rec = (pm && *pm) ? 0ull : 290ull;
        }
    }
    const int64_t n_grid = a.n_hint > 0 ? a.n_hint : n_max;
    int rc = radix_grid(n_grid, smem, G, a.count_nb_log2 >= 0 && n_grid <= 4 * a.count_n_max);
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
    a.dbg = nullptr;
#ifdef YH_MAP_TIMELINE
    static unsigned long long *dbg_dev = nullptr;
    if (getenv("YH_MAP_DBG")) {
        if (!dbg_dev) cudaMalloc(&dbg_dev, 256);
        cudaMemsetAsync(dbg_dev, 0, 256, st);
        a.dbg = dbg_dev;
    }
#endif
    void *args[] = {&a};
    YH_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(map_radix_kernel), dim3(G), dim3(RS_T), args, smem, st));
    YH_LAUNCH_CHECK("map_radix_kernel");
#ifdef YH_MAP_TIMELINE
    if (a.dbg) {
        unsigned long long h[32];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, a.dbg, 256, cudaMemcpyDeviceToHost);
        fprintf(stderr, "timeline (us from kernel start, CTA 0 of %d):", G);
        const int order[] = {1, 10, 13, 2, 3, 15, 4, 5, 6, 7, 8, 9};
        const char *name[] = {"prologue", "hist-zero", "hist", "hist-written", "barrier1", "matrix-read", "tables", "partition", "barrier2",
                              "units", "barrier3", "finish"};
        for (int k = 0; k < 12; ++k) fprintf(stderr, " %s %.2f", name[k], h[order[k]] ? (double)(h[order[k]] - h[0]) / 1000.0 : -1.0);
        fprintf(stderr, "\n");
    }
#endif
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
