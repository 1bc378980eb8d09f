"""cfg1 (BASELINE configs[0]: batch 64) latency of the fused decode+NMS call: plain C-ABI call, CUDA-graph replay of\nthe same call, and the Python surface; also shows the entry point is capturable."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "keras-object-detection_b200"))
import numpy as np, torch
from yolohot import _lib, utils as yu
from tests import fixtures as F
L=_lib.lib(); dev=torch.device("cuda:0")
p=torch.from_numpy(F.synth_dense(64, seed=1234)).to(dev)
boxes=torch.zeros((64,49,6),device=dev); cnt=torch.zeros(64,dtype=torch.int32,device=dev)
s=torch.cuda.Stream()
with torch.cuda.stream(s):
    sp=ctypes.c_void_p(s.cuda_stream)
    for _ in range(3):
        _lib.check(L.yh_decode_nms(p.data_ptr(),64,7,2,20,0.5,0.4,boxes.data_ptr(),cnt.data_ptr(),None,sp))
    s.synchronize()
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        sp2=ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.yh_decode_nms(p.data_ptr(),64,7,2,20,0.5,0.4,boxes.data_ptr(),cnt.data_ptr(),None,sp2))
boxes.zero_(); cnt.zero_()
g.replay(); torch.cuda.synchronize()
c_graph, b_graph = cnt.clone(), boxes.clone()
boxes.zero_(); cnt.zero_()
_lib.check(L.yh_decode_nms(p.data_ptr(),64,7,2,20,0.5,0.4,boxes.data_ptr(),cnt.data_ptr(),None,ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
print("graph replay == plain call:", bool(torch.equal(c_graph, cnt) and torch.equal(b_graph, boxes)))
def t(fn,n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e6
spd=ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("plain call  us:", t(lambda: L.yh_decode_nms(p.data_ptr(),64,7,2,20,0.5,0.4,boxes.data_ptr(),cnt.data_ptr(),None,spd)))
print("graph replay us:", t(lambda: g.replay()))
print("python surface us:", t(lambda: yu.decode_nms(p,20,2)))
