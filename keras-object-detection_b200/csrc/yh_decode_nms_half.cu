// yh_decode_nms_half.cu - float16 / bfloat16 instantiations of the fused decode + NMS kernels
// (yh_decode_nms_impl.cuh): the head adapter of SURVEY.md 8f N3 fused into the hot path.  A mixed-precision
// head's output is widened exactly as it is read, so results equal the float32 path on the widened tensor
// while the kernel reads half the bytes.  Own translation unit so that it compiles next to the float32 one.
#include "yh_decode_nms_impl.cuh"

using namespace yh;

// Note on odd channel counts (e.g. S=14, B=3, C=80: D = 95): rows are then only 2-byte aligned, every element is a
// 16-bit load plus a conversion, and the fused kernel (3.1 ms on cfg5) is no faster than widening first
// (yh_head_to_f32, 0.4 + 1.2 ms of copies around the 1.4 ms float32 kernel) - it only saves the temporary.  With an
// even channel count (VOC: D = 30) cells are read as aligned pairs and the fused path is the faster one.
extern "C" int yh_decode_nms_typed(const void *pred, int dtype, int64_t n, int S, int B, int C, float iou_thr, float conf_thr,
                                   int score_mode, float *out_boxes, int32_t *out_count, int32_t *out_keep_idx, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case YH_DTYPE_F32:
            return yh_decode_nms_ex(static_cast<const float *>(pred), n, S, B, C, iou_thr, conf_thr, score_mode, out_boxes,
                                    out_count, out_keep_idx, stream);
        case YH_DTYPE_F16:
            return decode_nms_typed<__half>(static_cast<const __half *>(pred), n, S, B, C, iou_thr, conf_thr, out_boxes, out_count,
                                            out_keep_idx, st, score_mode);
        case YH_DTYPE_BF16:
            return decode_nms_typed<__nv_bfloat16>(static_cast<const __nv_bfloat16 *>(pred), n, S, B, C, iou_thr, conf_thr, out_boxes,
                                                   out_count, out_keep_idx, st, score_mode);
    }
    set_error("decode_nms_typed: unknown dtype %d", dtype);
    return YH_ERR_ARG;
}
