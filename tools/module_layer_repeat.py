import sys,time,os
sys.path.insert(0,'/root/repo')
import numpy as np
from cart_slam_b200 import host
from cart_slam_b200.synth import SyntheticSequence
os.environ["CARTB200_QUIET"]="1"
W,H,D=1242,375,128
seq=SyntheticSequence(W,H,D,n_frames=8,tint=True)
fr=[seq.frame(1+i)[:2] for i in range(8)]
n=192
L=np.stack([fr[i%8][0] for i in range(n)]); R=np.stack([fr[i%8][1] for i in range(n)])
modules=[{"type":"superpixels","initial_iterations":24,"iterations":8,"block_size":12,"reset_iterations":64},
 {"type":"disparity","num_disparities":D,"smoothing_radius":2,"smoothing_iterations":1},
 {"type":"disparity_derivative"},
 {"type":"superpixel_disparity_planeseg","parameter_provider":{"type":"histogram_peak"}}]
host.run_config(modules,L[:8],R[:8],sequential=False)
for rep in range(8):
    for sq in (False,True):
        t=time.perf_counter(); host.run_config(modules,L,R,sequential=sq); dt=time.perf_counter()-t
        print(rep,'sequential' if sq else 'inflight12', round(n/dt,1),'fps', flush=True)
