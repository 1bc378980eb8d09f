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

size_t radix_ws_bytes(int64_t n_max, int C);
size_t match_ws_bytes(int64_t nt, int64_t np, bool sorted);
int reduce_impl(int nseg, const uint64_t *const *rec, const int64_t *nrec_max, const int64_t *const *nrec_dev, int n_gt,
                const int32_t *const *gt_parts, int C, float *out_ap, float *out_map, const uint64_t *wait_flags, int wait_n,
                uint64_t wait_epoch, int32_t *err, int64_t n_hint, void *workspace, size_t ws_bytes, void *stream);
int radix_launch(ReduceArgs &a, int64_t n_max, void *workspace, size_t ws_bytes, cudaStream_t st);

}  // namespace yh
