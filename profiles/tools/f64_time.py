import sys, importlib, torch
sys.path.insert(0,'.')
import bench
codec = importlib.import_module('3ddctvideoencoding_b200.codec')
W,H,F=1920,1080,64
dev=torch.device('cuda',0)
ts=torch.cuda.Stream(); torch.cuda.set_stream(ts); st=ts.cuda_stream
frames=bench.synth_clip_torch(W,H,F,1,dev)
cap=W*H*F//2+4096
d_stream=torch.zeros(cap,dtype=torch.uint8,device=dev)
d_out=torch.empty_like(frames)
for prec in (32,64):
    c=codec.Codec(W,H,8); c.set_option('precision',prec)
    for i in range(2):
        end=c.encode_u8_dev(frames,F,d_stream,cap,0,st); c.decode_u8_dev(d_stream,end//8+1,F,d_out,0,st)
    e=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(ts)
    for i in range(5): end=c.encode_u8_dev(frames,F,d_stream,cap,0,st)
    e[1].record(ts)
    for i in range(5): c.decode_u8_dev(d_stream,end//8+1,F,d_out,0,st)
    e[2].record(ts); torch.cuda.synchronize()
    print('precision',prec,'bits',end,'enc ms %.3f (%.0f fps)'%(e[0].elapsed_time(e[1])/5, F/(e[0].elapsed_time(e[1])/5)*1e3),'dec ms %.3f (%.0f fps)'%(e[1].elapsed_time(e[2])/5, F/(e[1].elapsed_time(e[2])/5)*1e3))
    c.close()
