// yh_common.cuh - shared helpers of libyolohot (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "yolohot.h"

namespace yh {

// ---- host-side error plumbing (yh_abi.cu) -------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void count_launch(int n = 1);

#define YH_CUDA(call)                                                  \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) return ::yh::cuda_fail(e_, #call);      \
    } while (0)

#define YH_REQUIRE(cond, ...)                                          \
    do {                                                               \
        if (!(cond)) {                                                 \
            ::yh::set_error(__VA_ARGS__);                              \
            return YH_ERR_ARG;                                         \
        }                                                              \
    } while (0)

#define YH_LAUNCH_CHECK(name)                                          \
    do {                                                               \
        cudaError_t e_ = cudaGetLastError();                           \
        if (e_ != cudaSuccess) return ::yh::cuda_fail(e_, name);       \
        ::yh::count_launch();                                          \
    } while (0)

// communicator of the single-process multi-GPU API (yh_comm.cu, yh_map.cu); comms are ncclComm_t
constexpr int kMaxPeers = 16;
struct Comm {
    int ndev = 0;
    int devs[kMaxPeers] = {};
    void *comms[kMaxPeers] = {};
    cudaEvent_t ev[kMaxPeers] = {};
    bool p2p = false;       // every device can read/write every other device's memory (NVLink / NVSwitch)
};

// ---- mAP records (yh_map.cu, yh_map_reduce.cu, yh_comm.cu) ------------------------------------
// One detection = one uint64:  class << 33 | ~orderable(conf) << 1 | tp.  Ascending order of the bits above
// bit 0 is (class asc, confidence desc) - the reference's processing order, utils.py:367; a stable sort on
// those bits keeps the row order among equal confidences.  Rows whose class is not an integer in [0, C)
// carry class = C (the reference never selects them, utils.py:329-330) and sort after every class.
constexpr int kRecClassShift = 33;
constexpr int kMaxMapClasses = 4096;          // shared-memory tables of the mAP kernels
constexpr int kMaxSegs = kMaxPeers;           // record segments (one per rank) the reduce stage takes

__host__ __device__ inline int class_bits(int C)
{
    int b = 1;
    while ((1ll << b) <= C) ++b;
    return b;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t orderable(float f)
{
    f = __fadd_rn(f, 0.0f);                       // -0 -> +0 so that equal floats get equal keys
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// class of a row as an integer in [0, C), or C when the reference would never select the row
// (utils.py:329-330 compare the float class with float(c), c = 0..C-1)
__device__ __forceinline__ uint32_t class_of(float cf, int C)
{
    const int ci = static_cast<int>(cf);
    return (cf >= 0.0f && cf < static_cast<float>(C) && static_cast<float>(ci) == cf) ? static_cast<uint32_t>(ci)
                                                                                    : static_cast<uint32_t>(C);
}

__device__ __forceinline__ uint64_t make_rec(uint32_t cls, float conf, uint32_t tp)
{
    return (static_cast<uint64_t>(cls) << kRecClassShift) | (static_cast<uint64_t>(~orderable(conf)) << 1) | tp;
}
#endif

inline int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

#ifdef __CUDACC__
// ---- exact float32 arithmetic of the reference -------------------------------------------
// TF evaluates every op separately in float32 (round-to-nearest, no FMA contraction).  The
// *_rn intrinsics are never contracted by nvcc, whatever --fmad says.

// tf.clip_by_value(v, 0, 1)                                     (utils.py:39)
__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// utils.py:24-43.  a = boxes1 (cx, cy, w, h), b = boxes2.  Corners are (c -/+ w) / 2 (the
// reference halves the centre too), extents clipped to [0, 1], |area|, +1e-6f last.
__device__ __forceinline__ float iou_ref(float ax, float ay, float aw, float ah,
                                         float bx, float by, float bw, float bh)
{
    const float x1n = __fmul_rn(__fsub_rn(ax, aw), 0.5f), y1n = __fmul_rn(__fsub_rn(ay, ah), 0.5f);
    const float x1x = __fmul_rn(__fadd_rn(ax, aw), 0.5f), y1x = __fmul_rn(__fadd_rn(ay, ah), 0.5f);
    const float x2n = __fmul_rn(__fsub_rn(bx, bw), 0.5f), y2n = __fmul_rn(__fsub_rn(by, bh), 0.5f);
    const float x2x = __fmul_rn(__fadd_rn(bx, bw), 0.5f), y2x = __fmul_rn(__fadd_rn(by, bh), 0.5f);
    const float iw = clip01(__fsub_rn(fminf(x1x, x2x), fmaxf(x1n, x2n)));
    const float ih = clip01(__fsub_rn(fminf(y1x, y2x), fmaxf(y1n, y2n)));
    const float inter = __fmul_rn(iw, ih);
    const float a1 = fabsf(__fmul_rn(__fsub_rn(x1x, x1n), __fsub_rn(y1x, y1n)));
    const float a2 = fabsf(__fmul_rn(__fsub_rn(x2x, x2n), __fsub_rn(y2x, y2n)));
    const float den = __fadd_rn(__fsub_rn(__fadd_rn(a1, a2), inter), 1e-6f);
    return __fdiv_rn(inter, den);
}

__device__ __forceinline__ float iou_ref(const float4 &a, const float4 &b)
{
    return iou_ref(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + bulk async copy (TMA, non-tensor form) ------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// same, for a waiter that is not latency critical (the TMA producer runs stages ahead): sleep
// between polls so its spin does not take issue slots from the consumer warps of its SM sub-partition
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity, unsigned ns)
{
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
// global -> shared bulk copy; completion is signalled on `bar` as `bytes` of transaction count.
// dst/src 16-byte aligned, bytes a multiple of 16.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// shared -> global bulk copy (TMA store, bulk async-group completion).  src/dst 16-byte aligned,
// bytes a multiple of 16.  Generic-proxy writes to the source must be made visible to the async
// proxy first (fence.proxy.async.shared::cta by the writers, then a barrier).
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read_all()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_all()
{
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
#endif  // __CUDACC__

}  // namespace yh
