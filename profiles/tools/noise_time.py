import sys, importlib, torch
sys.path.insert(0,'.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
W,H,F=1920,1080,64
dev=torch.device('cuda',0)
ts=torch.cuda.Stream(); torch.cuda.set_stream(ts); st=ts.cuda_stream
for name in ('noise','natural','constant'):
    if name=='noise': frames=torch.randint(0,256,(F,H,W),dtype=torch.uint8,device=dev)
    elif name=='natural': frames=bench.synth_clip_torch(W,H,F,1,dev)
    else: frames=torch.full((F,H,W),128,dtype=torch.uint8,device=dev)
    cap=W*H*F+4096
    d_stream=torch.zeros(cap,dtype=torch.uint8,device=dev)
    d_out=torch.empty_like(frames)
    c=codec.Codec(W,H,8)
    for i in range(3):
        end=c.encode_u8_dev(frames,F,d_stream,cap,0,st)
        c.decode_u8_dev(d_stream,end//8+1,F,d_out,0,st)
    e=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(ts)
    for i in range(5): end=c.encode_u8_dev(frames,F,d_stream,cap,0,st)
    e[1].record(ts)
    for i in range(5): c.decode_u8_dev(d_stream,end//8+1,F,d_out,0,st)
    e[2].record(ts); torch.cuda.synchronize()
    err=(d_out.float()-frames.float()).abs().mean().item()
    print(name,'bits/sample %.2f'%(end/frames.numel()),'enc ms %.3f'%(e[0].elapsed_time(e[1])/5),'dec ms %.3f'%(e[1].elapsed_time(e[2])/5),'launches',c.stat('launches'),'mean abs err %.2f'%err)
    c.close()
