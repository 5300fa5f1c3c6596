import sys, time, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, pkgload, oracle_py as O
P = pkgload.load()
w,h = 4000,3000
t=time.time(); img = O.synthetic_image(w,h,seed=0); print('synth %.1fs'%(time.time()-t))
bgra = np.concatenate([img[...,::-1], np.full((h,w,1),255,np.uint8)],axis=2)
for eff in (3,7):
    t=time.time(); data = P.encode_to_memory(bgra, P.EncoderOptions(quality=90, effort=eff)); t1=time.time()-t
    t=time.time(); data = P.encode_to_memory(bgra, P.EncoderOptions(quality=90, effort=eff)); t2=time.time()-t
    print('effort',eff,'bytes',len(data),'bpp %.3f'%(len(data)*8/w/h),'encode wall %.3fs (2nd %.3fs)'%(t1,t2), P.last_stage_times()['total'])
    for i in range(4):
        t=time.time(); out = P.load_image_bgra(data); dt=time.time()-t
        print('  decode wall %.1f ms'%(dt*1e3), {k:round(v,2) for k,v in P.last_stage_times().items()})
    if eff==7:
        t=time.time(); ref = O.decode(data, threads=os.cpu_count()); print('oracle decode %.2fs with %d threads'%(time.time()-t, os.cpu_count()))
        print('max err', int(np.abs(out[...,2::-1].astype(int)-ref.pixels.astype(int)).max()), 'psnr vs src %.2f'%O.psnr(ref.pixels,img))
        open('/root/repo/gpurun_out/sample12mp.jxl','wb').write(data)
