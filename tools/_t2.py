import os, sys, ctypes
sys.path.insert(0,'eao-fusion_b200')
import numpy as np, eaof
from eaof import synth
W,H,B,NF = 640,480,2,1000
fr = synth.make_frames(B, W, H, tex=synth.base_texture(W,H,seed=5))
ex = eaof.ORBextractor(NF,1.2,8,20,7,width=W,height=H,max_batch=B)
try:
    r = ex.extract_batch(fr); print('ok')
except Exception as e:
    print('FAIL')
L = eaof.lib(); L.eaof_debug_tma_records.restype = ctypes.POINTER(ctypes.c_int)
p = L.eaof_debug_tma_records()
a = np.ctypeslib.as_array(p, shape=(8192,8)).copy()
for i in range(8192):
    if a[i,7]: print('REC', i//8, i%8, *a[i])
