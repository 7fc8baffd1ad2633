# config-5-sized device-resident check: 1920x1080x4096 (8.5 G samples, stream > 2^32 bits)
import sys, importlib, torch, time
sys.path.insert(0,'.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
W,H,F=1920,1080,int(sys.argv[1]) if len(sys.argv)>1 else 4096
dev=torch.device('cuda',0)
base=bench.synth_clip_torch(W,H,256,1,dev)
frames=torch.empty((F,H,W),dtype=torch.uint8,device=dev)
for i in range(0,F,256):
    frames[i:i+256]=base.roll(i//256*8, dims=2) if i else base
N=W*H*F
cap=N//4+4096
d_stream=torch.zeros(cap,dtype=torch.uint8,device=dev)
out=torch.empty_like(frames)
c=codec.Codec(W,H,8)
ts=torch.cuda.Stream(); torch.cuda.set_stream(ts); st=ts.cuda_stream
t0=time.time()
end=c.encode_u8_dev(frames,F,d_stream,cap,0,st)
torch.cuda.synchronize(); t1=time.time()
print('bits',end,'> 2^32:',end>2**32,'bits/sample',end/N,'encode s',t1-t0)
dend=c.decode_u8_dev(d_stream,end//8+1,F,out,0,st)
torch.cuda.synchronize(); t2=time.time()
print('decode end ok',dend==end,'decode s',t2-t1)
# reference: per-256-frame ranges coded independently
part=torch.zeros(W*H*256//2+4096,dtype=torch.uint8,device=dev)
o2=torch.empty((256,H,W),dtype=torch.uint8,device=dev)
ok=True; tot=0
for i in range(0,F,256):
    e=c.encode_u8_dev(frames[i:i+256],256,part,part.numel(),0,st)
    c.decode_u8_dev(part,e//8+1,256,o2,0,st)
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(o2,out[i:i+256])); tot+=e
print('ranges equal one-shot decode:',ok,' sum of range bits == total:',tot==end)
for _ in range(2):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e2=torch.cuda.Event(enable_timing=True)
    e0.record(); end=c.encode_u8_dev(frames,F,d_stream,cap,0,st); e1.record(); c.decode_u8_dev(d_stream,end//8+1,F,out,0,st); e2.record(); torch.cuda.synchronize()
    print('encode %.1f ms (%.0f fps)  decode %.1f ms (%.0f fps)'%(e0.elapsed_time(e1),F/e0.elapsed_time(e1)*1e3,e1.elapsed_time(e2),F/e1.elapsed_time(e2)*1e3))
print('max mem GB', torch.cuda.max_memory_allocated()/1e9)
