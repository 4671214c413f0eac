import os, sys
sys.path.insert(0,'eao-fusion_b200')
import numpy as np, eaof
from eaof import synth
W,H,B,NF = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
fr = synth.make_frames(B, W, H, tex=synth.base_texture(W,H,seed=5))
ex = eaof.ORBextractor(NF,1.2,8,20,7,width=W,height=H,max_batch=B)
try:
    r = ex.extract_batch(fr)
    print('mode', os.environ.get('EAOF_FAST_TMA'), W,H,B, 'ok', len(r[0][0]), len(r[-1][0]))
except Exception as e:
    print('mode', os.environ.get('EAOF_FAST_TMA'), W,H,B,'FAIL', e)
