"""tensorflow.keras stand-in: `keras.losses.Loss` (base class of YoloV1Loss, loss.py:100) and
`keras.utils.Sequence` (base class of YoloV1Generator, dataset.py:19; a plain object here).
Keras calls `call(y_true, y_pred)` and applies the reduction; on the scalar that loss.py:215
returns, SUM_OVER_BATCH_SIZE is the identity."""
import types


class _Loss:
    def __init__(self, reduction="auto", name=None):
        self.reduction, self.name = reduction, name

    def __call__(self, y_true, y_pred, sample_weight=None):
        return self.call(y_true, y_pred)

    def call(self, y_true, y_pred):
        raise NotImplementedError


losses = types.ModuleType("tensorflow.keras.losses")
losses.Loss = _Loss


utils = types.ModuleType("tensorflow.keras.utils")
utils.Sequence = object
