import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/multiple-object-tracking_b200')
import mot_b200 as M
W,H,NS=1920,1080,64
pin=torch.empty((NS,H,W,3),dtype=torch.uint8).pin_memory()
fr=[pin[i].numpy() for i in range(NS)]
ctx=M.Context(W,H,max_tracks=8,n_frame_slots=NS,kind=M.TRACKER_KCF)
for i in range(NS): ctx.upload(i,fr[i])
ctx.sync()
for rep in range(3):
    t0=time.perf_counter()
    for i in range(NS): ctx.upload(i,fr[i])
    t1=time.perf_counter(); ctx.sync(); t2=time.perf_counter()
    print('issue %.2f ms, total %.2f ms'%((t1-t0)*1e3,(t2-t0)*1e3))
# pageable comparison
pg=np.zeros((H,W,3),np.uint8)
t0=time.perf_counter()
for i in range(8): ctx.upload(i,pg)
t1=time.perf_counter(); ctx.sync(); t2=time.perf_counter()
print('pageable x8: issue %.2f ms, total %.2f ms'%((t1-t0)*1e3,(t2-t0)*1e3))
