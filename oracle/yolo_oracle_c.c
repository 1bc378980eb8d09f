/* CPU oracle, C restatement ("port") of the YOLOv1 post-processing hot path.
 *
 * TEST INFRASTRUCTURE ONLY - the checker and the timed CPU baseline.  Nothing in
 * the product (keras-object-detection_b200/) links, loads or calls this file.
 *
 * PINNING: the reference (myungsanglee/Keras-Object-Detection) is pure Python on
 * TensorFlow, which is absent from this image, and it holds no tests or golden
 * outputs.  This file restates the reference loops in float32 C (compile with
 * -ffp-contract=off, never -ffast-math).  It is checked against the NumPy
 * restatement oracle/yolo_oracle.py (tests/test_oracle.py) and against
 * tests/golden/ref_golden.npz = outputs of the reference's own utils.py/loss.py run
 * on a NumPy stand-in for TF (tests/test_ref_golden.py).  Real TensorFlow kernels
 * remain unpinned.
 * Citations are relative to /root/reference/yolo_v1/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* utils.py:24-43 : corners (c -/+ w)/2, extents clipped to [0,1], abs areas,
 * ((a1 + a2) - inter) + 1e-6f */
static inline float clip01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

float yo_iou(const float *a, const float *b)
{
    float x1n = (a[0] - a[2]) / 2.0f, y1n = (a[1] - a[3]) / 2.0f;
    float x1x = (a[0] + a[2]) / 2.0f, y1x = (a[1] + a[3]) / 2.0f;
    float x2n = (b[0] - b[2]) / 2.0f, y2n = (b[1] - b[3]) / 2.0f;
    float x2x = (b[0] + b[2]) / 2.0f, y2x = (b[1] + b[3]) / 2.0f;
    float ixn = fmaxf(x1n, x2n), iyn = fmaxf(y1n, y2n);
    float ixx = fminf(x1x, x2x), iyx = fminf(y1x, y2x);
    float inter = clip01(ixx - ixn) * clip01(iyx - iyn);
    float a1 = fabsf((x1x - x1n) * (y1x - y1n));
    float a2 = fabsf((x2x - x2n) * (y2x - y2n));
    return inter / (((a1 + a2) - inter) + 1e-6f);
}

void yo_iou_many(const float *a, const float *b, int64_t n, float *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = yo_iou(a + 4 * i, b + 4 * i);
}

/* utils.py:173-216 : one cell -> [cls, conf, cx, cy, w, h] */
static inline void decode_cell(const float *p, int r, int c, int S, int B, int C, float *o)
{
    int cls = 0;
    float best = p[0];
    for (int j = 1; j < C; ++j)               /* :173 first max */
        if (p[j] > best) { best = p[j]; cls = j; }
    int k = 0;
    float conf = p[C];
    for (int b = 1; b < B; ++b)               /* :183 first max */
        if (p[C + 5 * b] > conf) { conf = p[C + 5 * b]; k = b; }
    const float *q = p + C + 5 * k;
    float inv = (float)(1.0 / (double)S);     /* :207 `1 / 7 *` -> float32 constant */
    o[0] = (float)cls;                        /* :175 */
    o[1] = conf;
    o[2] = inv * (q[1] + (float)c);           /* :207 column */
    o[3] = inv * (q[2] + (float)r);           /* :208 row */
    o[4] = q[3];
    o[5] = q[4];
}

void yo_decode(const float *pred, int64_t n, int S, int B, int C, float *out)
{
    const int D = C + 5 * B, M = S * S;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        for (int cell = 0; cell < M; ++cell)
            decode_cell(pred + (i * M + cell) * D, cell / S, cell % S, S, B, C,
                        out + (i * M + cell) * 6);
}

/* utils.py:91-114 : filter >, stable descending sort, pop/filter greedy loop.
 * rows (M,6); writes K rows to out_rows (if non-null) and their source indices to
 * keep_idx (if non-null); scratch needs 2*M ints. */
int yo_nms_image(const float *rows, int M, float iou_thr, float conf_thr,
                 float *out_rows, int32_t *keep_idx, int32_t *scratch)
{
    int32_t *cur = scratch, *nxt = scratch + M;
    int n = 0;
    for (int i = 0; i < M; ++i)               /* :95 */
        if (rows[6 * i + 1] > conf_thr) cur[n++] = i;
    for (int i = 1; i < n; ++i) {             /* :98 stable insertion sort, descending */
        int32_t v = cur[i];
        float s = rows[6 * v + 1];
        int j = i - 1;
        while (j >= 0 && rows[6 * cur[j] + 1] < s) { cur[j + 1] = cur[j]; --j; }
        cur[j + 1] = v;
    }
    int K = 0;
    while (n >= 1) {                          /* :103 */
        int head = cur[0];                    /* :104 */
        int m = 0;
        for (int t = 1; t < n; ++t) {         /* :106 */
            int j = cur[t];
            if (rows[6 * j] != rows[6 * head] ||
                yo_iou(rows + 6 * head + 2, rows + 6 * j + 2) < iou_thr)   /* :108 */
                nxt[m++] = j;
        }
        int32_t *tmp = cur; cur = nxt; nxt = tmp; n = m;                    /* :110 */
        if (out_rows) memcpy(out_rows + 6 * K, rows + 6 * head, 6 * sizeof(float));
        if (keep_idx) keep_idx[K] = head;
        ++K;                                  /* :112 */
    }
    return K;
}

/* driver of utils.py:471-480 for a batch; padded outputs (N,M,6), (N), (N,M) */
void yo_decode_nms(const float *pred, int64_t n, int S, int B, int C,
                   float iou_thr, float conf_thr,
                   float *out_boxes, int32_t *out_count, int32_t *out_keep_idx, int nthreads)
{
    const int D = C + 5 * B, M = S * S;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        float *dec = (float *)malloc(sizeof(float) * 6 * M);
        int32_t *scr = (int32_t *)malloc(sizeof(int32_t) * 2 * M);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            for (int cell = 0; cell < M; ++cell)
                decode_cell(pred + (i * M + cell) * D, cell / S, cell % S, S, B, C, dec + 6 * cell);
            int K = yo_nms_image(dec, M, iou_thr, conf_thr,
                                 out_boxes ? out_boxes + i * M * 6 : NULL,
                                 out_keep_idx ? out_keep_idx + i * M : NULL, scr);
            if (out_keep_idx)
                for (int t = K; t < M; ++t) out_keep_idx[i * M + t] = -1;
            out_count[i] = K;
        }
        free(dec);
        free(scr);
    }
}

/* loss.py:126-213 forward; element-wise float32, sums in double.
 * terms[6] = xy, wh, obj, noobj, cls, total */
void yo_loss(const float *yt, const float *yp, int64_t cells, int B, int C,
             float lambda_coord, float lambda_noobj, double *terms, int nthreads)
{
    const int D = C + 5 * B;
    double sxy = 0, swh = 0, sob = 0, snb = 0, scl = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : sxy, swh, sob, snb, scl)
    for (int64_t i = 0; i < cells; ++i) {
        const float *t = yt + i * D, *p = yp + i * D;
        int k = 0;
        float u = yo_iou(t + C + 1, p + C + 1);              /* :127-133 */
        for (int b = 1; b < B; ++b) {
            float v = yo_iou(t + C + 1, p + C + 1 + 5 * b);
            if (v > u) { u = v; k = b; }                      /* :136 first max */
        }
        const float *q = p + C + 5 * k;
        float obj = t[C], noobj = 1.0f - obj;                 /* :162-163 */
        float dx = t[C + 1] - q[1], dy = t[C + 2] - q[2];
        sxy += (double)(obj * (dx * dx)) + (double)(obj * (dy * dy));      /* :171 */
        for (int a = 3; a <= 4; ++a) {                        /* :176-178 */
            float pv = q[a];
            float sg = (pv > 0.0f) ? 1.0f : ((pv < 0.0f) ? -1.0f : 0.0f);
            float d = sqrtf(t[C + a]) - sg * sqrtf(fabsf(pv) + 1e-6f);
            swh += (double)(obj * (d * d));
        }
        float e = u - q[0];
        sob += (double)(obj * (e * e));                       /* :189 */
        float z = 0.0f - q[0];
        snb += (double)(noobj * (z * z));                     /* :197 */
        for (int j = 0; j < C; ++j) {                         /* :206 */
            float d = t[j] - p[j];
            scl += (double)(obj * (d * d));
        }
    }
    terms[0] = sxy; terms[1] = swh; terms[2] = sob; terms[3] = snb; terms[4] = scl;
    terms[5] = (double)lambda_coord * (sxy + swh) + sob + (double)lambda_noobj * snb + scl; /* :210-213 */
}

/* utils.py:317-456 : rows [img, cls, conf, cx, cy, w, h]; returns mAP (float32 maths for
 * the PR points, double for the trapezoid sum); ap_out[C] optional. */
typedef struct { float conf; int32_t row; } det_t;

static int det_cmp(const void *a, const void *b)
{
    const det_t *x = (const det_t *)a, *y = (const det_t *)b;
    if (x->conf > y->conf) return -1;
    if (x->conf < y->conf) return 1;
    return (x->row > y->row) - (x->row < y->row);             /* stable: :367 */
}

double yo_map(const float *tb, int64_t nt, const float *pb, int64_t np_, int C,
              float iou_thr, float *ap_out)
{
    double sum_ap = 0.0;
    det_t *det = (det_t *)malloc(sizeof(det_t) * (size_t)(np_ > 0 ? np_ : 1));
    int32_t *gt = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nt > 0 ? nt : 1));
    uint8_t *claimed = (uint8_t *)malloc((size_t)(nt > 0 ? nt : 1));
    for (int c = 0; c < C; ++c) {                              /* :325 */
        int64_t nd = 0, ng = 0;
        for (int64_t i = 0; i < np_; ++i)
            if (pb[7 * i + 1] == (float)c) { det[nd].conf = pb[7 * i + 2]; det[nd].row = (int32_t)i; ++nd; }
        for (int64_t i = 0; i < nt; ++i)
            if (tb[7 * i + 1] == (float)c) gt[ng++] = (int32_t)i;
        float ap = 0.0f;
        if (ng > 0) {                                          /* :334-336 */
            qsort(det, (size_t)nd, sizeof(det_t), det_cmp);
            memset(claimed, 0, (size_t)ng);
            float tpc = 0.0f, fpc = 0.0f, total = (float)ng;
            float prev_r = 0.0f, prev_p = 1.0f;                /* :438-439 */
            double acc = 0.0;
            for (int64_t d = 0; d < nd; ++d) {                 /* :373 */
                const float *dr = pb + 7 * det[d].row;
                float best = 0.0f;
                int64_t bj = -1, first = -1;
                for (int64_t g = 0; g < ng; ++g) {             /* :378,386 */
                    const float *gr = tb + 7 * gt[g];
                    if (gr[0] != dr[0]) continue;
                    if (first < 0) first = g;
                    float v = yo_iou(dr + 3, gr + 3);          /* :387 */
                    if (v > best) { best = v; bj = g; }        /* :389 */
                }
                if (bj < 0) bj = first;                        /* best_gt_idx defaults to 0 (:383) */
                int is_tp = 0;
                if (best > iou_thr && bj >= 0 && !claimed[bj]) { claimed[bj] = 1; is_tp = 1; } /* :395-418 */
                if (is_tp) tpc += 1.0f; else fpc += 1.0f;      /* :430-431 */
                float r = tpc / (total + 1e-6f);               /* :434 */
                float p = tpc / ((tpc + fpc) + 1e-6f);         /* :435 */
                float term = ((r - prev_r) * (p + prev_p)) / 2.0f;  /* np.trapz element (float32) */
                acc += (double)term;
                prev_r = r; prev_p = p;
            }
            ap = (float)acc;                                   /* :444 */
        }
        if (ap_out) ap_out[c] = ap;
        sum_ap += (double)ap;
    }
    free(det); free(gt); free(claimed);
    return sum_ap / (double)C;                                 /* :456 */
}

int yo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
