"""CPU tests of the host-side Python logic that needs no GPU: label-file parsing against what the reference's
`_get_boxes` returned (tests/golden/ref_golden.npz), the head adapter's shape logic, shard ranges, and the
loud failures of the product path without a CUDA device."""
import os

import numpy as np
import pytest
import torch

R = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))


def test_get_boxes_parses_yolo_txt_like_the_reference(tmp_path):
    from yolohot import dataset as yd
    f = tmp_path / "img.txt"
    f.write_text("0 0.756250 0.210417 0.293750 0.179167\n1 0.450000 0.480556 0.582812 0.505556\n2 0.287891 0.661806 0.233594 0.556944\n")
    got = yd.get_boxes(str(f))                                  # the content of the reference's data/test.txt
    assert got.dtype == np.float64 and np.array_equal(got, R["lab_txt_boxes"])
    (tmp_path / "empty.txt").write_text("")
    assert yd.get_boxes(str(tmp_path / "empty.txt")).shape == (0, 5)
    lab = yd.YoloV1Labels(3, 2)
    assert lab.output_shape == (7, 7, 13) and np.array_equal(lab._get_boxes(str(f)), got)


def test_head_adapter_shape_logic():
    from yolohot._tensor import as_grid
    flat = torch.zeros(5, 7 * 7 * 30)
    v, n, S = as_grid(flat, 20, 2)
    assert (n, S) == (5, 7) and tuple(v.shape) == (5, 7, 7, 30) and v.data_ptr() == flat.data_ptr()      # a view, no copy
    v, n, S = as_grid(torch.zeros(2, 14 * 14 * 95), 80, 3, grid=14)
    assert (n, S) == (2, 14)
    for bad in (torch.zeros(2, 7 * 7 * 30 + 1), torch.zeros(2, 7, 6, 30), torch.zeros(2, 7, 7, 31), torch.zeros(7, 7, 30)):
        with pytest.raises(ValueError):
            as_grid(bad, 20, 2)
    with pytest.raises(ValueError):
        as_grid(torch.zeros(2, 7 * 7 * 30), 20, 2, grid=14)


def test_shard_ranges_cover_the_batch():
    from yolohot import dist as yd
    for n in (0, 1, 5, 5000, 1_000_003):
        for w in (1, 2, 3, 8):
            r = [yd.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a CUDA device")
def test_product_path_fails_loudly_without_a_gpu():
    from yolohot import dataset as yd, loss as yl, utils as yu
    p = np.zeros((2, 7, 7, 30), np.float32)
    for call in (lambda: yu.decode_nms(p, 20, 2), lambda: yu.decode_predictions(p, 20, 2),
                 lambda: yu.intersection_over_union(p[0, 0, :, :4], p[0, 0, :, :4]),
                 lambda: yl.YoloV1Loss(20, 2)(p, p), lambda: yd.get_labels([[0.5, 0.5, 0.1, 0.1, 0]]),
                 lambda: yu.MeanAveragePrecision(20, 2).update_state(p, p)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_tf_keras_branch_of_the_loss_is_taken_and_needs_a_gpu():
    """With `tensorflow` importable (the stand-in), YoloV1Loss is a keras.losses.Loss and its call goes through
    tf.custom_gradient + DLPack; without a CUDA device it must refuse, not fall back.  Own process: the stand-in
    must not leak into the other tests."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "tf_branch_check.py"), "cpu"], capture_output=True, text=True,
                       timeout=300, env=env)
    assert r.returncode == 0 and "tf_branch_check ok (cpu)" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
