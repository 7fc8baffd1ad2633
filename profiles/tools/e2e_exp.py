import sys, importlib, time, threading as th, queue, ctypes as C
import torch
sys.path.insert(0,'.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
W,H,F=1920,1080,256
dev=torch.device('cuda',0)
frames=bench.synth_clip_torch(W,H,F,1,dev)
h_frames=torch.empty((F,H,W),dtype=torch.uint8,pin_memory=True); h_frames.copy_(frames)
h_out=torch.empty((F,H,W),dtype=torch.uint8,pin_memory=True)
def run(nchunk,nthr,mode):
    cf=F//nchunk; fsz=W*H*cf; ccap=fsz//2+4096
    hs=[torch.zeros(ccap,dtype=torch.uint8,pin_memory=True) for _ in range(nchunk)]
    encs=[codec.Codec(W,H,8) for _ in range(nthr)]; decs=[codec.Codec(W,H,8) for _ in range(nthr)]
    L=encs[0].L; sizes=[0]*nchunk
    q=None
    def enc(k):
        nb,ny=C.c_uint64(),C.c_size_t()
        for i in range(k,nchunk,nthr):
            assert L.dct3d_encode_u8(encs[k].h,h_frames.data_ptr()+i*fsz,cf,hs[i].data_ptr(),ccap,C.byref(nb),C.byref(ny))==0
            sizes[i]=ny.value
            if q is not None: q.put(i)
    def dec(k):
        if q is None:
            for i in range(k,nchunk,nthr):
                assert L.dct3d_decode_u8(decs[k].h,hs[i].data_ptr(),sizes[i],cf,h_out.data_ptr()+i*fsz)==0
            return
        while True:
            i=q.get()
            if i<0: return
            assert L.dct3d_decode_u8(decs[k].h,hs[i].data_ptr(),sizes[i],cf,h_out.data_ptr()+i*fsz)==0
    res=[]
    # prime the streams so that decode-only / independent modes have something to decode
    for k in range(nthr): enc(k)
    for rep in range(4):
        q=queue.Queue() if mode=='both' else None
        t0=time.perf_counter()
        if mode in ('enc','both','indep'):
            te=[th.Thread(target=enc,args=(k,)) for k in range(nthr)]
        else: te=[]
        td=[th.Thread(target=dec,args=(k,)) for k in range(nthr)] if mode in ('dec','both','indep') else []
        for t in te+td: t.start()
        for t in te: t.join()
        if q is not None:
            for _ in td: q.put(-1)
        for t in td: t.join()
        res.append(time.perf_counter()-t0)
    for c in encs+decs: c.close()
    return min(res[1:])
for nchunk,nthr in ((16,1),(16,2),(32,2)):
    e=run(nchunk,nthr,'enc'); d=run(nchunk,nthr,'dec'); b=run(nchunk,nthr,'both'); i=run(nchunk,nthr,'indep')
    print(f"ranges {nchunk} threads {nthr}: enc-only {e*1e3:.2f} ms ({0.531/e:.1f} GB/s H2D), dec-only {d*1e3:.2f} ms ({0.531/d:.1f} GB/s D2H), both {b*1e3:.2f} ms ({F/b:.0f} fps), independent {i*1e3:.2f} ms")
