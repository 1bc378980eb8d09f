"""Small end-to-end pass over every kernel, meant for compute-sanitizer (memcheck / synccheck):
    compute-sanitizer --tool memcheck python profiles/sanitize_small.py
(compute-sanitizer is closed on the build pool's GPU boxes, so in round 1 this only ran plain.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests import fixtures as F  # noqa: E402
from yolohot import dataset as yd, loss as yl, utils as yu  # noqa: E402

dev = torch.device("cuda:0")
n_voc = int(os.environ.get("YH_SAN_VOC", 1500))
n_big = int(os.environ.get("YH_SAN_BIG", 700))
p = torch.from_numpy(F.synth_dense(n_voc, seed=3)).to(dev)
b, c = yu.decode_nms(p, 20, 2)                                   # TMA tile ring + direct tail
ps = torch.from_numpy(F.synth_stress(n_big)).to(dev)
bs, cs = yu.decode_nms(ps, 80, 3, 0.5, 0.05)                     # cooperative team kernel
os.environ["YH_COOP"] = "0"
bs3, cs3 = yu.decode_nms(ps[:64], 80, 3, 0.5, 0.05)              # direct kernel
del os.environ["YH_COOP"]
assert torch.equal(cs[:64], cs3)
yt = torch.from_numpy(F.synth_labels(600, seed=7)).to(dev)
yp = torch.from_numpy(F.synth_loss_pred(tuple(yt.shape), seed=7)).to(dev).requires_grad_(True)
yl.YoloV1Loss(20, 2)(yt, yp).backward()
ev = yu.MeanAveragePrecision(20, 2)
mp = torch.from_numpy(F.synth_map_pred(yt.cpu().numpy())).to(dev)
ev.update_state(yt, mp)
m = float(ev.result())
rng = np.random.default_rng(0)
lists = [np.concatenate([rng.random((k, 2)), 0.1 + 0.5 * rng.random((k, 2)), rng.integers(0, 20, (k, 1))], 1) for k in rng.integers(0, 40, 300)]
lab = yd.encode_labels(lists, None, 7, 20, 2)
px = yu.pixel_boxes(b, 448, 448, count=c)
h = yu.decode_nms(p.to(torch.bfloat16), 20, 2)
d = yu.decode_predictions(p, 20, 2)
i = yu.intersection_over_union(d[0, :, 2:], d[1, :, 2:])
torch.cuda.synchronize()
print("sanitize_small ok: kept", int(c.sum()), int(cs.sum()), "mAP %.4f" % m, "labels", float(lab.sum()))
