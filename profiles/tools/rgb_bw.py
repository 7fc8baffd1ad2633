import sys, importlib, torch
sys.path.insert(0,'.')
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
dev=torch.device('cuda',0)
n=1920*1080*3*64
rgb=torch.randint(0,256,(n,),dtype=torch.uint8,device=dev)
pl=[torch.empty(n//3,dtype=torch.uint8,device=dev) for _ in range(3)]
out=torch.empty_like(rgb)
ts=torch.cuda.Stream(); torch.cuda.set_stream(ts); st=ts.cuda_stream
c=codec.Codec(1920,1080,8)
for name,fn in (('split',lambda: c.rgb_split_dev(rgb,n,*pl,st)),('mix',lambda: c.rgb_mix_dev(*pl,n//3,out,st))):
    for _ in range(3): fn()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for _ in range(10): fn()
    e1.record(ts); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(name,'%.3f ms'%ms,'%.0f GB/s (read+write)'%(2*n/ms/1e6))
print('ok', bool(torch.equal(out,rgb)))
