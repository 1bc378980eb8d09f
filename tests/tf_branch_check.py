"""Runs yolohot.loss with a `tensorflow` module importable (the NumPy stand-in of tests/golden/tfshim, extended with
custom_gradient / experimental.dlpack / keras.losses.Loss), in its own process so that the stand-in never leaks into
other tests.  Usage: python tests/tf_branch_check.py cpu|gpu
  cpu: the TF / Keras branch of YoloV1Loss is taken and fails loudly (no CUDA device, no CPU fallback);
  gpu: the same branch end to end on cuda:0 - value and custom gradient against the torch path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "tfshim"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import tensorflow as tf  # noqa: E402  (the stand-in)
from tests import fixtures as F  # noqa: E402
from yolohot import loss as yl  # noqa: E402


def main(mode):
    assert yl._tf is tf and issubclass(yl.YoloV1Loss, tf.keras.losses.Loss)          # loss.py:100
    lf = yl.YoloV1Loss(20, 2)
    assert lf.name == "YoloV1Loss" and lf.lambda_coord == 5 and lf.lambda_noobj == 0.5 and lf.batch_size == 0
    yt = F.synth_labels(16, seed=7)
    yp = F.synth_loss_pred(yt.shape, seed=7)
    t_tf, p_tf = tf.constant(yt), tf.constant(yp)
    if mode == "cpu":
        try:
            lf(t_tf, p_tf)                                                          # Keras __call__ -> call -> custom_gradient
        except RuntimeError as e:
            assert "no CPU fallback" in str(e) or "CUDA" in str(e), e
            print("tf_branch_check ok (cpu): Keras branch taken, refused without a CUDA device")
            return
        raise AssertionError("the TF branch ran without a GPU")
    import torch
    total = lf(t_tf, p_tf)
    assert tf.is_tensor(total) and lf.batch_size == 16
    pt = torch.from_numpy(yp).cuda().requires_grad_(True)
    want = yl.YoloV1Loss(20, 2)(torch.from_numpy(yt).cuda(), pt)
    want.backward()
    assert float(total) == float(want.detach()), (float(total), float(want))
    g = total._grad_fn(tf.constant(np.float32(2.0)))                                # what the tape applies
    assert np.array_equal(np.asarray(g), 2.0 * pt.grad.cpu().numpy())
    flat = lf(t_tf, tf.constant(yp.reshape(16, -1)))                                # flat head output (model.py:107)
    assert float(flat) == float(total) and np.asarray(flat._grad_fn(tf.constant(np.float32(1.0)))).shape == (16, 7 * 7 * 30)
    terms = yl.yolo_v1_loss_terms(t_tf, p_tf)                                       # TF in -> TF out through DLPack
    assert tf.is_tensor(terms) and float(np.asarray(terms)[5]) == float(total)
    print("tf_branch_check ok (gpu): custom_gradient value and gradient equal the torch path")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cpu")
