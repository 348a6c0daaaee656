import sys, time, os
sys.path.insert(0,'galileo-sdr-sim_b200'); sys.path.insert(0,'tests')
import numpy as np, e1b200 as E, e1util as U
fs=U.fs_as_reference(2.6e6); n=260000; ne=2999; nc=36
recs=U.synthetic_recs_fast(ne,nc,fs,seed=1000)
s=E.Synth(fs,n,nc)
h_recs=E.PinnedBuffer(recs.nbytes); h_recs.u8[:]=recs.view(np.uint8).reshape(-1)
h_out=E.PinnedBuffer(ne*n*4)
rv=h_recs.view(U.REC_DTYPE).reshape(ne,nc); ov=h_out.view(np.int16)
os.environ.pop("E1B200_TRACE",None)
for i in range(2): s.synth_epochs(rv,ov)
os.environ["E1B200_TRACE"]="1"
t=time.perf_counter(); s.synth_epochs(rv,ov); print("e2e ms",(time.perf_counter()-t)*1e3, file=sys.stderr)
