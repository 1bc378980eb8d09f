"""Phase timeline of the counting path of map_radix_kernel (build csrc/yh_map_reduce.cu with -DYH_MAP_TIMELINE, run with YH_MAP_DBG=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "keras-object-detection_b200"))
import torch
from tests import fixtures as F
from yolohot import utils as yu
dev = torch.device("cuda:0")
yt0 = F.synth_labels(5000, seed=11)
a, b = torch.from_numpy(yt0).to(dev), torch.from_numpy(F.synth_map_pred(yt0)).to(dev)
ev = yu.MeanAveragePrecision(20, 2)
ev.update_state(a, b)
for _ in range(6):
    m = ev.result()
torch.cuda.synchronize()
