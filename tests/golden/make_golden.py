"""Regenerates tests/golden/golden.npz: outputs of the ORACLE (oracle/yolo_oracle.py) on the
reference's own fixtures and on small seeded synthetic inputs.

The reference cannot run here (TensorFlow is absent), so these are NOT
observed reference outputs: they pin the oracle against regressions, while the values that can
be derived by hand from the reference source are asserted separately, as
literals, in tests/test_oracle.py.  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import yolo_oracle as O  # noqa: E402
from tests import fixtures as F  # noqa: E402


def main():
    g = {}
    yt, yp = F.utils_demo()
    g["utils_demo_nms_pred"] = O.non_max_suppression(O.decode_predictions(yp, 3, 2)[0])
    g["utils_demo_nms_true"] = O.non_max_suppression(O.decode_predictions(yt, 3, 2)[0])
    ev = O.MeanAveragePrecision(3, 2)
    ev.update_state(yt, yp)
    g["utils_demo_map"] = np.float32(ev.result())
    lt, lp = F.loss_demo()
    L = O.yolo_v1_loss(lt, lp, 3, 2)
    g["loss_demo_terms"] = np.array([L[k + "_f64"] for k in ("xy", "wh", "obj", "noobj", "cls", "total")])
    g["loss_demo_grad"] = O.yolo_v1_loss_grad(lt, lp, 3, 2)
    mt, mp, mp2 = F.metric_demo()
    ev = O.MeanAveragePrecision(20, 2, nms_true=False)
    for i in range(5):
        ev.update_state(mt, mp if i == 0 else mp2)          # metric.py:146-151
    g["metric_demo_map"] = np.float32(ev.result())
    lab = O.encode_labels(F.TEST_TXT_BOXES, 7, 3, 2)[None].astype(np.float32)
    g["test_txt_nms"] = O.non_max_suppression(O.decode_predictions(lab, 3, 2)[0])
    # small seeded synthetic cases
    p = F.synth_dense(8, seed=1234)
    b, c, k = O.decode_nms(p, 20, 2)
    g["dense8_boxes"], g["dense8_count"], g["dense8_idx"] = b, c, k
    p = F.synth_quantised(4, 14, 3, 80, seed=5)
    b, c, k = O.decode_nms(p, 80, 3, 0.5, 0.05)
    g["quant4_count"], g["quant4_idx"] = c, k
    yt = F.synth_labels(16, seed=7)
    yp = F.synth_loss_pred(yt.shape, seed=7)
    L = O.yolo_v1_loss(yt, yp)
    g["loss16_terms"] = np.array([L[k + "_f64"] for k in ("xy", "wh", "obj", "noobj", "cls", "total")])
    yt = F.synth_labels(40, seed=11)
    mp = F.synth_map_pred(yt)
    ev = O.MeanAveragePrecision(20, 2)
    ev.update_state(yt, mp)
    g["map40"] = np.float32(ev.result())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz"), **g)
    for k_, v in g.items():
        print(k_, np.asarray(v).shape)


if __name__ == "__main__":
    main()
