/* yolohot.h - C-ABI of libyolohot.so: the B200 (sm_100a) implementation of the YOLOv1
 * post-processing / loss / mAP hot path of myungsanglee/Keras-Object-Detection.
 *
 * The reference has no FFI: its boundary is the Python call surface of
 * yolo_v1/utils.py, loss.py and metric.py.  Each entry point below names the reference
 * function (file:line, relative to the reference repo) whose arithmetic it replaces; the
 * Python mirror of that surface (keras-object-detection_b200/yolohot/{utils,loss,metric}.py)
 * binds these symbols with ctypes.  See INTEGRATION.md for the binding a maintainer adds.
 *
 * Conventions
 *   - plain C: pointers + sizes; no C++/torch types.  `stream` is a cudaStream_t passed as
 *     void* (NULL = legacy default stream).  All device entry points are asynchronous on
 *     `stream` and run on the CUDA device that is current when they are called.
 *   - all tensors are float32 / int32, C-contiguous, device-resident unless the name ends
 *     in `_host`.  Layout of a prediction/label tensor is (N, S, S, C + 5*B) with the cell's
 *     channels [class probs (C) | B x (conf, x, y, w, h)]  (yolo_v1/dataset.py:88-112).
 *   - return value: YH_OK or a negative YH_ERR_*; the message is in yh_last_error()
 *     (thread-local).  No exceptions, aborts or CPU fallbacks cross this boundary.
 *   - inputs are expected to be FINITE.  Signed zeros and infinities / NaNs in the values a
 *     result is computed from (the selected box of a cell, its class scores, the confidences)
 *     behave like the reference's float32 arithmetic (tests: test_ref_golden.py, "nonfinite_sel").
 *     One documented deviation: the reference picks a cell's box by multiplying every box by
 *     a 0/1 mask and summing (utils.py:184-197), so an inf / NaN in a box that was NOT
 *     selected turns the cell's row into NaN there; the kernels select, and return the selected
 *     box.  Nothing is raised for non-finite inputs.
 */
#ifndef YOLOHOT_H_
#define YOLOHOT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* exported symbol (the library is built with -fvisibility=hidden) */
#if defined(__GNUC__)
#define YH_API __attribute__((visibility("default")))
#else
#define YH_API
#endif

#define YH_OK 0
#define YH_ERR_ARG (-1)         /* null / shape / dtype / device / contiguity / alignment */
#define YH_ERR_CUDA (-2)        /* a CUDA runtime call or launch failed */
#define YH_ERR_NCCL (-3)        /* libnccl could not be loaded, or an NCCL call failed */
#define YH_ERR_UNSUPPORTED (-4) /* configuration outside the compiled limits (e.g. S*S > 256) */

#define YH_MAX_CELLS 256        /* S*S limit of the NMS kernels (8 candidate slots per lane) */

/* Forward declaration of the DLPack exchange struct (ABI of dlpack.h v0.8 / v1.0 legacy
 * DLManagedTensor); full layout in yh_dlpack.h. */
struct DLManagedTensor;

/* ---- library ------------------------------------------------------------------------- */
YH_API int yh_version(void);                 /* 1000*major + minor */
YH_API const char *yh_last_error(void); /* thread-local, never NULL */
YH_API int yh_device_info(int *sm_count, int *cc_major, int *cc_minor); /* of the current device */
/* number of kernel launches issued by this library since load (all threads); bench.py's
 * "gpu_launches" claim is counted from this. */
YH_API int64_t yh_launch_count(void);

/* ---- IoU: utils.py:9-43 intersection_over_union ---------------------------------------
 * boxes1, boxes2: (n, 4) [cx, cy, w, h];  out: (n) - the reference's (...,4)->(...,1). */
YH_API int yh_iou(const float *boxes1, const float *boxes2, int64_t n, float *out, void *stream);

/* ---- decode: utils.py:152-218 decode_predictions --------------------------------------
 * pred (n,S,S,C+5B) -> out_boxes (n, S*S, 6) rows [cls, conf, cx, cy, w, h]. */
YH_API int yh_decode(const float *pred, int64_t n, int S, int B, int C, float *out_boxes, void *stream);

/* ---- NMS: utils.py:79-114 non_max_suppression, batched --------------------------------
 * boxes (n, M, 6) decoded rows; per image: keep conf > conf_thr, stable descending sort,
 * greedy suppression of same-class boxes with IoU >= iou_thr.
 * out_boxes (n, M, 6): the first out_count[i] rows of image i are the kept rows in pick
 * order, the rest is left untouched.  out_keep_idx (n, M) (nullable): source row index of
 * each kept row. */
YH_API int yh_nms(const float *boxes, int64_t n, int M, float iou_thr, float conf_thr,
           float *out_boxes, int32_t *out_count, int32_t *out_keep_idx, void *stream);

/* ---- fused decode + IoU + NMS: loop body of utils.py:470-480 --------------------------
 * The graded throughput path.  Same outputs as yh_decode followed by yh_nms;
 * out_keep_idx holds the source CELL index r*S+c of each kept row (nullable). */
YH_API int yh_decode_nms(const float *pred, int64_t n, int S, int B, int C,
                  float iou_thr, float conf_thr,
                  float *out_boxes, int32_t *out_count, int32_t *out_keep_idx, void *stream);

/* Same with a score mode.  YH_SCORE_CONF is the reference (a cell's score is its best box
 * confidence, utils.py:173,183-197).  YH_SCORE_CONF_X_PROB is an EXTENSION that the reference
 * does not have: score = best box confidence x winning class probability (the class-specific
 * score of the YOLO paper), used for the threshold, the order and the reported confidence. */
#define YH_SCORE_CONF 0
#define YH_SCORE_CONF_X_PROB 1
YH_API int yh_decode_nms_ex(const float *pred, int64_t n, int S, int B, int C,
                     float iou_thr, float conf_thr, int score_mode,
                     float *out_boxes, int32_t *out_count, int32_t *out_keep_idx, void *stream);

/* Same, HOST buffers in and out (pageable or pinned): chunked H2D copy / kernel / D2H copy
 * pipelined on internal streams of device `device`; returns after the results are on the
 * host.  This is the call bench.py times for its end-to-end ("e2e") figure. */
YH_API int yh_decode_nms_host(const float *pred_host, int64_t n, int S, int B, int C,
                       float iou_thr, float conf_thr,
                       float *out_boxes_host, int32_t *out_count_host,
                       int32_t *out_keep_idx_host /* nullable */, int device);

/* The same with a COMPACT result: only the kept rows come back.  out_rows_host (row_capacity, 7) receives
 * rows [img, cls, conf, cx, cy, w, h] of all images in image order (img = index in the batch, as float32
 * like the reference's accumulator, utils.py:476; within an image the NMS pick order), out_count_host (n)
 * the kept rows per image, *out_total_rows their sum.  D2H bytes: 28 * kept rows + 4 * n instead of
 * 24 * S*S * n.  If the rows do not fit row_capacity the call copies what fits, sets *out_total_rows to
 * the number needed and returns YH_ERR_ARG (S*S * n rows always suffice). */
YH_API int yh_decode_nms_host_rows(const void *pred_host, int dtype, int64_t n, int S, int B, int C,
                            float iou_thr, float conf_thr,
                            float *out_rows_host, int64_t row_capacity, int32_t *out_count_host,
                            int64_t *out_total_rows, int device);

/* The chunk pipeline of the *_host entry points overlaps its copies only with PINNED host memory
 * (pageable buffers work, but the driver stages them and the copies run one after the other).
 * yh_host_alloc = cudaHostAlloc (portable; write_combined = 1 for input staging buffers that the
 * CPU only writes: faster over PCIe on some hosts, very slow to read back on the CPU). */
YH_API int yh_host_alloc(size_t bytes, int write_combined, void **ptr);
YH_API int yh_host_free(void *ptr);

/* ---- confidence filter without NMS: metric.py:35-37, 81 (the stale evaluator's ground truth) ----
 * rows (n, M, 6) -> out_rows (n, M, 6): the rows of image i with conf > conf_thr, cell order kept,
 * packed at the front; out_count (n).  out_rows must not alias rows. */
YH_API int yh_filter_rows(const float *rows, int64_t n, int M, float conf_thr,
                   float *out_rows, int32_t *out_count, void *stream);

/* ---- evaluator rows: utils.py:476-489 (prefix img_idx, append) ------------------------
 * Compacts padded NMS output (n, M, 6) + count (n) into rows [img, cls, conf, cx,cy,w,h]
 * appended at out_rows + 7 * (*row_cursor) in image order; image i gets img index
 * img_base + i (as float32, like the reference).  row_cursor is a DEVICE int64 counter that
 * the call advances by the number of rows of the batch; rows beyond `out_capacity` rows are
 * dropped (the cursor still advances).  One kernel (single-pass chained scan), no host
 * synchronisation.  Calls that share a stream are ordered; calls on different streams of one
 * device may run concurrently. */
YH_API int yh_rows_append(const float *boxes, const int32_t *count, int64_t n, int M, int64_t img_base,
                   float *out_rows, int64_t out_capacity, int64_t *row_cursor, void *stream);

/* ---- evaluator update: the whole accumulation step of MeanAveragePrecision.update_state
 * (utils.py:470-491) after decode + NMS, in ONE kernel: appends the prediction rows and the
 * ground-truth rows of the batch like yh_rows_append (cursors[0] / cursors[1], device int64)
 * and matches every image on the spot (utils.py:373-422): a detection can only claim a ground
 * truth of its own image, and NMS emits an image's detections in the reference's processing
 * order, so the TP / FP decision of a detection is final as soon as its image is seen.
 * rec (pred_capacity) receives one packed record per prediction row, at the row's index:
 *     uint64 = class << 33 | ~orderable(conf) << 1 | tp        (orderable: monotone float -> uint32)
 * gt_per_class (C int32) ACCUMULATES the ground truths per class (zero it when the evaluator
 * restarts).  result() is then yh_map_reduce over rec[0 .. cursors[0]) - no host sync anywhere. */
YH_API int yh_eval_update(const float *pred_boxes, const int32_t *pred_count,
                   const float *true_boxes, const int32_t *true_count,
                   int64_t n, int M, int64_t img_base, int C, float iou_thr,
                   float *pred_rows, int64_t pred_capacity, float *true_rows, int64_t true_capacity,
                   uint64_t *rec, int64_t *cursors, int32_t *gt_per_class, void *stream);

/* ---- evaluator update, fused: MeanAveragePrecision.update_state (utils.py:470-491) in ONE launch.
 * y_true, y_pred (n, S, S, C+5B) float32, 8-byte aligned: per image one warp runs the fused decode + NMS
 * of the prediction cells and of the label cells (utils.py:475 / :480, thresholds nms_iou_thr /
 * nms_conf_thr - the reference hard-codes 0.5 / 0.4), matches the image (match_iou_thr, utils.py:496)
 * while its kept rows are still in shared memory, and appends rows and records exactly like
 * yh_eval_update (same buffers, cursors and record layout; the results are bit-identical to
 * yh_decode_nms x 2 + yh_eval_update).  restart != 0: the evaluator starts over (utils.py:484-486) - the
 * two cursors and gt_per_class are zeroed on the stream before the kernel.  Grids of more than 64 cells, or class counts whose tables do
 * not fit shared memory, return YH_ERR_UNSUPPORTED: use the three calls.  One launch wins where three
 * are latency - up to about 8,000 VOC-sized images per call on a B200 (33 vs 45 us at 5,000); beyond
 * that the TMA tile kernels behind yh_decode_nms stream the cells faster (85 vs 132 us at 20,000). */
YH_API int yh_eval_update_state(const float *y_true, const float *y_pred, int64_t n, int S, int B, int C,
                         float nms_iou_thr, float nms_conf_thr, int64_t img_base, float match_iou_thr,
                         float *pred_rows, int64_t pred_capacity, float *true_rows, int64_t true_capacity,
                         uint64_t *rec, int64_t *cursors, int32_t *gt_per_class, int restart, void *stream);

/* ---- loss: loss.py:120-215 YoloV1Loss.call (+ its autodiff backward) ------------------
 * y_true, y_pred: (n_cells, C+5B) i.e. the (N,S,S,D) tensors flattened over cells.
 * out_terms: 6 floats [xy, wh, obj, noobj, cls, total] (batch sums, loss.py:172-213).
 * out_grad (nullable): d(total)/d(y_pred), same shape as y_pred. */
YH_API int yh_loss(const float *y_true, const float *y_pred, int64_t n_cells, int B, int C,
            float lambda_coord, float lambda_noobj,
            float *out_terms, float *out_grad, void *stream);

/* ---- mAP: utils.py:303-456 mean_average_precision -------------------------------------
 * Stage 1 on arbitrary rows (per image shard, no communication): greedy IoU matching of
 * detections to ground truths.  true_rows (nt,7), pred_rows (np,7) rows
 * [img, cls, conf, cx,cy,w,h] in any order.  nt_dev / np_dev (nullable): DEVICE int64 row
 * counts (<= nt / np, which then only bound the buffers) - e.g. an evaluator's cursors.
 * flags: YH_MAP_TRUE_ROWS_BY_IMAGE = the ground-truth rows are grouped by image with
 * nondecreasing image index (what an evaluator accumulates); otherwise they are first ordered
 * by image with a stable radix sort.
 * out_rec (np): one packed record per detection, in ROW order (layout: yh_eval_update);
 * out_gt_per_class (C int32): ground truths per class.  Rows whose class is not an integer in
 * [0, C) are ignored (the reference never selects them, utils.py:329-330); their record carries
 * class C and sorts after every class.
 * workspace: yh_workspace_bytes(YH_OP_MAP_MATCH, max(nt, np), ..) bytes, 256-byte aligned, or
 * NULL (stream-ordered allocation inside the call). */
#define YH_MAP_TRUE_ROWS_BY_IMAGE 1
YH_API int yh_map_match(const float *true_rows, int64_t nt, const int64_t *nt_dev,
                 const float *pred_rows, int64_t np, const int64_t *np_dev,
                 int C, float iou_thr, int flags,
                 uint64_t *out_rec, int32_t *out_gt_per_class,
                 void *workspace, size_t workspace_bytes, void *stream);

/* Stage 2 (on the records of all shards concatenated in shard order = image order, and the
 * per-class ground-truth counts summed): ONE persistent cooperative kernel - stable radix sort
 * by (class asc, confidence desc), cumulative TP / FP, precision / recall, np.trapz AP per class
 * (out_ap (C), nullable) and the mean over all C classes (out_map (1)).  For small and medium
 * inputs the kernel takes a sort-free path with bit-identical results (the np.trapz term of a
 * true positive needs only two ranks, which are counted inside (class, confidence bucket)
 * groups after one unordered partition); the choice is made on the device from the record and
 * pair counts (YH_MAP_COUNT=0 in the environment forces the sort).  nrec_dev (nullable):
 * DEVICE int64 record count (<= nrec, which then only bounds the buffers); n_hint: expected
 * record count (sizes the grid only, any value is correct; 0 = nrec).  workspace:
 * yh_workspace_bytes(YH_OP_MAP_REDUCE, nrec, ..) bytes, 256-byte aligned, or NULL.  No host
 * synchronisation; capturable in a CUDA graph (the count is re-read at every replay). */
YH_API int yh_map_reduce(const uint64_t *rec, int64_t nrec, const int64_t *nrec_dev, int64_t n_hint,
                  const int32_t *gt_per_class, int C, float *out_ap, float *out_map,
                  void *workspace, size_t workspace_bytes, void *stream);

/* ---- multi-GPU exchange step of the mAP -------------------------------------------------
 *   Between stage 1 on every shard and stage 2.
 *
 * (a) Fused with kernels over peer-mapped memory (NVLink / NVSwitch), no collective call and no
 * host synchronisation.  Every rank owns one exchange buffer of yh_map_exchange_bytes(n, C,
 * capacity) bytes, zero-filled once, writable by all its peers: cudaDeviceEnablePeerAccess in
 * one process (yh_comm_init_all does it), or yh_ipc_alloc / yh_ipc_open between processes.
 * yh_map_exchange on rank `self` stores its records (rec[0 .. *nrec_dev)), their count and its
 * per-class ground-truth counts into slot `self` of EVERY rank's buffer (bufs[0..n)), then
 * releases a per-rank flag = epoch.  yh_map_reduce_exchanged on a rank waits (bounded, ~2 s) for
 * the n flags of `epoch` in its own buffer and reduces the n segments in rank order.  epoch
 * starts at 1 and increases by 1 per exchange, in lockstep on all ranks (double buffered by its
 * parity).  capacity = records per rank and epoch the buffers were sized for (equal on all
 * ranks).  err (nullable device int32) receives YH_MAP_ERR_TIMEOUT / YH_MAP_ERR_OVERFLOW (a
 * rank's shard exceeded `capacity`); the result is then meaningless.  n_hint: expected total
 * record count (sizes the grid only; 0 = n * capacity). */
#define YH_MAP_ERR_TIMEOUT 1
#define YH_MAP_ERR_OVERFLOW 2
YH_API size_t yh_map_exchange_bytes(int n_peers, int C, int64_t capacity);
YH_API int yh_map_exchange(int n_peers, int self, void *const *bufs, int C, int64_t capacity,
                    const uint64_t *rec, int64_t nrec_max, const int64_t *nrec_dev,
                    const int32_t *gt_per_class, uint64_t epoch, void *stream);
YH_API int yh_map_reduce_exchanged(int n_peers, void *own_buf, int C, int64_t capacity, uint64_t epoch, int64_t n_hint,
                            float *out_ap, float *out_map, int32_t *err,
                            void *workspace, size_t workspace_bytes, void *stream);

/* (b) As a collective, single process driving all devices: yh_comm_init_all = ncclCommInitAll
 * over `devs` (NULL = 0..ndev-1; NCCL is loaded at run time from the libnccl.so.2 of the
 * process) + peer access + one event per device.  yh_map_allgather: rec[d] holds nrec[d] records
 * on device d (nrec is a HOST array), gt_per_class[d] (C int32, device d) is all-reduced in
 * place; out_rec[d] (device d, capacity out_capacity records) receives the concatenation in
 * device order; streams[d] (nullable) is device d's stream.  Asynchronous on those streams.
 * yh_comm_p2p: 1 when every device can access every other device's memory (then (a) works on
 * plain cudaMalloc buffers); yh_comm_barrier: cross-device barrier on the streams, no host wait. */
YH_API int yh_comm_init_all(int ndev, const int *devs, void **comm);
YH_API int yh_comm_destroy(void *comm);
YH_API int yh_map_allgather(void *comm, const uint64_t *const *rec, const int64_t *nrec,
                     int32_t *const *gt_per_class, int C, uint64_t *const *out_rec,
                     int64_t out_capacity, void *const *streams);
YH_API int yh_comm_p2p(void *comm);
YH_API int yh_comm_barrier(void *comm, void *const *streams);

/* Device buffers that other processes of the same box can map: yh_ipc_alloc = cudaMalloc (zero-filled) + an opaque
 * YH_IPC_HANDLE_BYTES handle to send to the peers; yh_ipc_open maps a peer's handle (never one's own) with peer access
 * over NVLink; yh_ipc_close unmaps; yh_ipc_free releases an own buffer after every peer closed it. */
#define YH_IPC_HANDLE_BYTES 64
YH_API int yh_ipc_alloc(size_t bytes, void **ptr, unsigned char *handle);
YH_API int yh_ipc_open(const unsigned char *handle, void **ptr);
YH_API int yh_ipc_close(void *ptr);
YH_API int yh_ipc_free(void *ptr);

/* Device scratch (bytes) of an operation for n images / rows.  The mAP stages take it as a caller
 * workspace (or allocate it themselves when passed NULL); the others allocate internally. */
#define YH_OP_DECODE_NMS 1
#define YH_OP_DECODE_NMS_HOST 2
#define YH_OP_LOSS 3
#define YH_OP_MAP_MATCH 4
#define YH_OP_MAP_REDUCE 5
YH_API size_t yh_workspace_bytes(int op, int64_t n, int S, int B, int C);

/* ---- callers either side of the path ------------------------ */

/* N2 - dataset.py:88-112 YoloV1Generator._get_labels, batched.  boxes (total, 5) FLOAT64 rows
 * [cx, cy, w, h, class] (what dataset.py:114-123 _get_boxes / albumentations hand over);
 * offsets (n+1) int64, image i owns rows [offsets[i], offsets[i+1]).  out (n,S,S,C+5B) float32
 * is written completely (zeros + labels): first box wins a cell (:107), x from the column, y from
 * the row (:101-105), float64 arithmetic then the float32 cast of :85.  Python index rules are
 * kept (cell/class indices wrap once when negative); a box whose indices would raise IndexError
 * in the reference is skipped and counted in *out_bad (device int32, nullable). */
YH_API int yh_encode_labels(const double *boxes, const int64_t *offsets, int64_t n, int S, int B, int C,
                     float *out, int32_t *out_bad, void *stream);

/* N3 - head adapter (train.py:208, model.py:107).  A flat (N, S*S*(C+5B)) head output IS the
 * (N,S,S,C+5B) tensor - pass the same pointer to the entry points above.  A half-precision head
 * is widened exactly to float32 first: src_dtype YH_DTYPE_F16 / YH_DTYPE_BF16, n elements. */
#define YH_DTYPE_F32 0
#define YH_DTYPE_F16 1
#define YH_DTYPE_BF16 2
YH_API int yh_head_to_f32(const void *src, int src_dtype, int64_t n, float *dst, void *stream);

/* The same adapter fused into the hot path: yh_decode_nms_ex on a prediction tensor of element type
 * `dtype` (YH_DTYPE_F32 / _F16 / _BF16).  A half-precision head is widened exactly as it is read, so the
 * results equal yh_decode_nms_ex on the widened tensor while the kernel reads half the bytes. */
YH_API int yh_decode_nms_typed(const void *pred, int dtype, int64_t n, int S, int B, int C,
                        float iou_thr, float conf_thr, int score_mode,
                        float *out_boxes, int32_t *out_count, int32_t *out_keep_idx, void *stream);

/* yh_decode_nms_host for a HOST tensor of element type `dtype`: a float16 / bfloat16 head crosses PCIe and HBM at
 * half the bytes and is widened exactly inside the kernel (results = yh_decode_nms_host on the widened tensor). */
YH_API int yh_decode_nms_host_typed(const void *pred_host, int dtype, int64_t n, int S, int B, int C,
                             float iou_thr, float conf_thr,
                             float *out_boxes_host, int32_t *out_count_host,
                             int32_t *out_keep_idx_host /* nullable */, int device);

/* N4 - utils.py:652-655 (get_tagged_img / get_grid_tagged_img): pixel corners of kept rows,
 * xmin = int((cx - w/2) * width), ... in float32, int() truncating toward zero.
 * rows (n, M, 6) as written by yh_nms / yh_decode_nms, count (n) nullable; out (n, M, 4) int32
 * [xmin, ymin, xmax, ymax], rows at or beyond count[i] are set to -1. */
YH_API int yh_pixel_boxes(const float *rows, const int32_t *count, int64_t n, int M, int width, int height,
                   int32_t *out, void *stream);

/* ---- DLPack front ends ----------------------------------------------------------------
 * Same operations taking DLManagedTensor* (what `tensor.__dlpack__()` capsules hold, so
 * torch / TF-Keras / CuPy tensors pass zero-copy).  They validate device (kDLCUDA, the
 * current device), dtype, C-contiguity, shape and 4-byte alignment, then call the pointer
 * entry points above.  A kDLCPU tensor is YH_ERR_ARG: there is no CPU fallback. */
YH_API int yh_iou_dl(const struct DLManagedTensor *boxes1, const struct DLManagedTensor *boxes2,
              struct DLManagedTensor *out, void *stream);
YH_API int yh_decode_dl(const struct DLManagedTensor *pred, int B, int C,
                 struct DLManagedTensor *out_boxes, void *stream);
YH_API int yh_nms_dl(const struct DLManagedTensor *boxes, float iou_thr, float conf_thr,
              struct DLManagedTensor *out_boxes, struct DLManagedTensor *out_count,
              struct DLManagedTensor *out_keep_idx /* nullable */, void *stream);
YH_API int yh_decode_nms_dl(const struct DLManagedTensor *pred, int B, int C,
                     float iou_thr, float conf_thr,
                     struct DLManagedTensor *out_boxes, struct DLManagedTensor *out_count,
                     struct DLManagedTensor *out_keep_idx /* nullable */, void *stream);
YH_API int yh_loss_dl(const struct DLManagedTensor *y_true, const struct DLManagedTensor *y_pred,
               int B, int C, float lambda_coord, float lambda_noobj,
               struct DLManagedTensor *out_terms, struct DLManagedTensor *out_grad /* nullable */,
               void *stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLOHOT_H_ */
