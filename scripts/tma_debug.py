"""one setData with the TMA bit-planes kernel, for compute-sanitizer"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bpvo_b200 import synth
from bpvo_b200.engine import Context
from conftest import make_params
sc = synth.scene_small(96, 128)
p = make_params("bitplanes", 2, "tukey")
ctx = Context(sc.K, sc.baseline, (sc.rows, sc.cols), p, flags=2)
f = ctx.frame()
img, d = sc.render(0)
f.setData(img, d)
ctx.synchronize()
D = f.descriptor(0)
print("descriptor ok", D.shape, float(D.sum()))
