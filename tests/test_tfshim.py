"""Known-answer tests of the NumPy stand-in for TensorFlow (tests/golden/tfshim) that the reference's own source is
executed on to produce tests/golden/ref_golden.npz: every op family the reference's hot path leans on (SURVEY.md
section 8c, "call sites relying on library semantics") must give the result TensorFlow's public API documentation
shows for it.  The inputs / outputs below are the examples of the tf.* API pages (tf.argsort, tf.math.argmax,
tf.one_hot, tf.clip_by_value, tf.where, tf.unique_with_counts, tf.TensorArray, tf.lookup.experimental.DenseHashTable)
plus the tie / unwritten-slot behaviours the reference depends on (utils.py:98,173,183,367-369,382-389)."""
import os
import sys

import numpy as np
import pytest

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tfshim")


@pytest.fixture()
def tf():
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "tensorflow" or k.startswith("tensorflow.")}
    sys.path.insert(0, SHIM)
    try:
        import tensorflow as tf_mod
        yield tf_mod
    finally:
        sys.path.remove(SHIM)
        for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_argsort_and_argmax(tf):
    v = tf.constant([1, 10, 26.9, 2.8, 166.32, 62.3])
    assert np.array_equal(tf.argsort(v), [0, 3, 1, 2, 5, 4])                                   # tf.argsort page
    assert np.array_equal(tf.argsort(v, direction="DESCENDING"), [4, 5, 2, 1, 3, 0])
    # equal values keep their order in a stable descending sort (utils.py:98, :367)
    assert np.array_equal(tf.argsort(tf.constant([3.0, 1.0, 3.0, 2.0, 1.0]), direction="DESCENDING", stable=True), [0, 2, 3, 1, 4])
    a = tf.constant([2, 20, 30, 3, 6])
    b = tf.constant([[2, 20, 30, 3, 6], [3, 11, 16, 1, 8], [14, 45, 23, 5, 27]])
    assert int(tf.math.argmax(a)) == 2                                                        # tf.math.argmax page
    assert np.array_equal(tf.math.argmax(b, 0), [2, 2, 0, 2, 2])
    assert np.array_equal(tf.math.argmax(b, 1), [2, 2, 1])
    assert int(tf.math.argmax(tf.constant([1.0, 5.0, 5.0, 0.0]))) == 1                        # first maximum (utils.py:173,183)


def test_one_hot_clip_where(tf):
    assert np.array_equal(tf.one_hot([0, 1, 2], 3), np.eye(3, dtype=np.float32))               # tf.one_hot page
    assert np.array_equal(tf.one_hot([0, 2, -1, 1], 3), [[1, 0, 0], [0, 0, 1], [0, 0, 0], [0, 1, 0]])
    t = tf.constant([[-10., -1., 0.], [0., 2., 10.]])
    assert np.array_equal(tf.clip_by_value(t, clip_value_min=-1, clip_value_max=1), [[-1, -1, 0], [0, 1, 1]])   # its page
    assert np.array_equal(tf.where([True, False, False, True]), [[0], [3]])                   # tf.where page
    assert np.array_equal(tf.where([True, False], tf.constant([1, 2]), tf.constant([8, 9])), [1, 9])


def test_unique_with_counts_and_hash_table(tf):
    y, idx, count = tf.unique_with_counts(tf.constant([1, 1, 2, 4, 4, 4, 7, 8, 8]))           # its page
    assert np.array_equal(y, [1, 2, 4, 7, 8])
    assert np.array_equal(idx, [0, 0, 1, 2, 2, 2, 3, 4, 4])
    assert np.array_equal(count, [2, 1, 3, 1, 2])
    table = tf.lookup.experimental.DenseHashTable(key_dtype=tf.float32, value_dtype=tf.int32, default_value=-1,
                                                  empty_key=-1.0, deleted_key=-2.0)
    table.insert(tf.constant([3.0, 5.0]), tf.constant([30, 50]))
    assert np.array_equal(table.lookup(tf.constant([5.0, 4.0, 3.0])), [50, -1, 30])           # default for a missing key


def test_tensor_array_reads_zeros_from_unwritten_slots(tf):
    """utils.py:368-369,382-383: slots never written read as zeros of the element shape."""
    ta = tf.TensorArray(tf.float32, size=0, dynamic_size=True, element_shape=(2,))
    ta = ta.write(2, tf.constant([1.0, 2.0]))
    assert np.array_equal(ta.read(0), [0.0, 0.0]) and np.array_equal(ta.read(2), [1.0, 2.0])
    assert np.array_equal(ta.stack(), [[0, 0], [0, 0], [1, 2]])
    assert int(ta.size()) == 3


def test_single_element_predicates_and_float32_arithmetic(tf):
    """utils.py:108,389,395: a rank-1 single-element tensor is usable as an `if` predicate; arithmetic stays float32."""
    assert bool(tf.constant([0.7]) > 0.5) and not bool(tf.constant([0.2]) > 0.5)
    x = tf.constant([0.1], dtype=tf.float32) + tf.constant([0.2], dtype=tf.float32)
    assert x.dtype == np.float32 and x[0] == np.float32(0.1) + np.float32(0.2)
    assert np.array_equal(tf.math.sign(tf.constant([-2.0, 0.0, 3.0])), [-1.0, 0.0, 1.0])
    assert np.array_equal(tf.cumsum(tf.constant([1.0, 2.0, 3.0])), [1.0, 3.0, 6.0])
    assert float(tf.reduce_sum(tf.constant([[1.0, 2.0], [3.0, 4.0]]))) == 10.0
