import sys, importlib, torch
sys.path.insert(0,'.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
W,H,F=1920,1080,int(sys.argv[1]) if len(sys.argv)>1 else 256
dev=torch.device('cuda',0)
frames=bench.synth_clip_torch(W,H,F,1,dev)
cap=W*H*F//2+4096
d_stream=torch.zeros(cap,dtype=torch.uint8,device=dev)
d_out=torch.empty_like(frames)
c=codec.Codec(W,H,8)
ts=torch.cuda.Stream(); torch.cuda.set_stream(ts); st=ts.cuda_stream
for i in range(3):
    end=c.encode_u8_dev(frames,F,d_stream,cap,0,st)
    c.decode_u8_dev(d_stream,end//8+1,F,d_out,0,st)
torch.cuda.synchronize()
print('ok', end)
