import sys, time, ctypes as C, threading
sys.path[:0]=['realsense-pointcloud_b200','tools']
import numpy as np, rspcl_b200 as R, gen_scene
from concurrent.futures import ThreadPoolExecutor
W,H=640,480; NPX=W*H; CP=8
L=R.lib()
fr,_=gen_scene.make_sweep(2,CP+1)
g=np.eye(4); g[:3,:3]=gen_scene.rot_y(-0.523599)
kw=dict(max_iterations=50, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300, mse_threshold_absolute=-1.0)
icp=R.icp_params(**kw)
NC=4
ws=[]
for w in range(NC):
    c=R.Context(0); h=c.pinned((CP+1)*NPX*32); h.view(R.PCL32)[:]=np.concatenate([R.to_pcl32(f) for f in fr]); ho=c.pinned(CP*NPX*32)
    ws.append(dict(ctx=c,h=h,ho=ho,fr=c.cloud(CP+1,NPX),out=c.cloud(CP,NPX),oc=np.zeros(CP,np.int32)))
cnts=np.full(CP+1,NPX,np.int32); si=np.arange(1,CP+1,dtype=np.int32); ti=np.arange(0,CP,dtype=np.int32)
def up(w): w['fr'].upload_raw(w['h'].ctypes.data_as(C.c_void_p), cnts, W,H,R.LAYOUT_PCL32); w['ctx'].sync()
def comp(w): R.register_pairs(w['ctx'], w['fr'], si, ti, R.COARSE_ICP, icp=icp, guess=g, out_transformed=w['out'])
def down(w): w['ctx'].check(L.rspcl_cloud_download(w['ctx'].h, w['out'].h, w['ho'].ctypes.data_as(C.c_void_p), R.LAYOUT_PCL32, C.c_longlong(CP*NPX), w['oc'].ctypes.data_as(C.c_void_p)))
for w in ws: up(w); comp(w); down(w)
for w in ws: comp(w)
def t(fn):
    t0=time.perf_counter(); fn(); return (time.perf_counter()-t0)*1e3
print('single: up %.2f comp %.2f down %.2f ms'%(t(lambda: up(ws[0])), t(lambda: comp(ws[0])), t(lambda: down(ws[0]))))
pool=ThreadPoolExecutor(NC)
for n in (1,2,4):
    dt=t(lambda: list(pool.map(comp, ws[:n])))
    print('%d concurrent computes: %.2f ms'%(n,dt))
dt=t(lambda: list(pool.map(lambda i: (up(ws[0]) if i==0 else comp(ws[1])), range(2)))); print('up || comp: %.2f ms'%dt)
dt=t(lambda: list(pool.map(lambda i: (down(ws[0]) if i==0 else comp(ws[1])), range(2)))); print('down || comp: %.2f ms'%dt)
dt=t(lambda: list(pool.map(lambda i: (down(ws[0]) if i==0 else (up(ws[2]) if i==2 else comp(ws[1]))), range(3)))); print('up || comp || down: %.2f ms'%dt)
# ---- replicate the bench pipeline with host timestamps
h2d=threading.Lock(); d2h=threading.Lock()
log=[]
T0=[0.0]
def run(wi, steps, locks):
    w=ws[wi]
    for s in range(steps):
        a=time.perf_counter()
        if locks:
            with h2d: up(w)
        else: up(w)
        b=time.perf_counter(); comp(w); c=time.perf_counter()
        if locks:
            with d2h: down(w)
        else: down(w)
        d=time.perf_counter()
        log.append((wi,s,(a-T0[0])*1e3,(b-T0[0])*1e3,(c-T0[0])*1e3,(d-T0[0])*1e3))
for locks in (False, True):
    log.clear(); T0[0]=time.perf_counter()
    list(pool.map(lambda i: run(i,4,locks), range(NC)))
    tot=(time.perf_counter()-T0[0])*1e3
    print('locks',locks,'total %.1f ms for 4 steps -> %.2f ms/step'%(tot,tot/4))
    for r in sorted(log, key=lambda r:(r[0],r[1]))[:8]:
        print('  w%d s%d start %.1f up_done %.1f comp_done %.1f down_done %.1f'%r)
