import sys, importlib, torch
sys.path.insert(0,'.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
W,H,F=1920,1080,256
dev=torch.device('cuda',0)
frames=bench.synth_clip_torch(W,H,F,1,dev)
cap=W*H*F//2+4096
d_stream=torch.zeros(cap,dtype=torch.uint8,device=dev)
d_out=torch.empty_like(frames)
ts=torch.cuda.Stream(); torch.cuda.set_stream(ts); st=ts.cuda_stream
for cube in (4,8):
    c=codec.Codec(W,H,cube)
    c.set_option('reuse_zeroed',1)
    def timeit(label, fn, n=5):
        for _ in range(2): fn()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/n
        print(f'cube={cube} {label:20s} {ms:8.3f} ms  {F/ms*1e3:10.0f} fps')
    end=c.encode_u8_dev(frames,F,d_stream,cap,0,st)
    print('cube',cube,'bits/sample',end/(W*H*F))
    timeit('encode', lambda: c.encode_u8_dev(frames,F,d_stream,cap,0,st))
    timeit('decode', lambda: c.decode_u8_dev(d_stream,end//8+1,F,d_out,0,st))
    c.close()
