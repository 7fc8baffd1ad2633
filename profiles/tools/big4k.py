import sys, importlib, torch
sys.path.insert(0,'.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
W,H,F=3840,2160,128
dev=torch.device('cuda',0)
frames=bench.synth_clip_torch(W,H,F,3,dev)
N=W*H*F; cap=N//2+4096
d_stream=torch.zeros(cap,dtype=torch.uint8,device=dev); out=torch.empty_like(frames)
c=codec.Codec(W,H,8)
ts=torch.cuda.Stream(); torch.cuda.set_stream(ts); st=ts.cuda_stream
torch.cuda.synchronize()
end=c.encode_u8_dev(frames,F,d_stream,cap,0,st); dend=c.decode_u8_dev(d_stream,end//8+1,F,out,0,st); torch.cuda.synchronize()
print('bits/sample',end/N,'decode end ok',dend==end,'mean abs err',(out.float()-frames.float()).abs().mean().item())
# ranges of 32 frames coded separately must decode to the same pixels, and their bit counts must add up
part=torch.zeros(W*H*32//2+4096,dtype=torch.uint8,device=dev); o2=torch.empty((32,H,W),dtype=torch.uint8,device=dev)
ok=True; tot=0
for i in range(0,F,32):
    e=c.encode_u8_dev(frames[i:i+32],32,part,part.numel(),0,st); c.decode_u8_dev(part,e//8+1,32,o2,0,st); torch.cuda.synchronize()
    ok=ok and bool(torch.equal(o2,out[i:i+32])); tot+=e
print('ranges equal one-shot:',ok,'bits add up:',tot==end)
for _ in range(2):
    e0,e1,e2=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e0.record(); end=c.encode_u8_dev(frames,F,d_stream,cap,0,st); e1.record(); c.decode_u8_dev(d_stream,end//8+1,F,out,0,st); e2.record(); torch.cuda.synchronize()
    print('4K: encode %.2f ms (%.0f fps)  decode %.2f ms (%.0f fps)'%(e0.elapsed_time(e1),F/e0.elapsed_time(e1)*1e3,e1.elapsed_time(e2),F/e1.elapsed_time(e2)*1e3))
