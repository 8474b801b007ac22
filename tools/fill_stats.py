"""Workload statistics behind fine's cost model (CPU only, through the oracle): fills and commands per tile before and after the
occlusion start, segments and pixel touches per fill. usage: python tools/fill_stats.py"""
import sys, numpy as np, time
sys.path.insert(0, '.')
from vello_b200 import scenes
from vello_b200.encoding import resolve, BLACK
from oracle.vbo import Oracle
packed = resolve(scenes.paris_like(30000, 4096, seed=30000).encoding)
o = Oracle(threads=8)
o.bind(packed, 4096, 4096, BLACK.premul_rgba8_u32(), 2)
t=time.time(); o.run("pathtag","path_tiling"); print("stages", round(time.time()-t,1),"s")
ptcl = o.buffer("ptcl"); segs = o.buffer("segments")
print(segs.dtype)
n_tiles = 256*256
CMD_END,CMD_FILL,CMD_SOLID,CMD_COLOR,CMD_JUMP=0,1,3,5,12
size = {1:4,3:1,5:2,6:3,7:3,8:3,9:2,10:1,11:3,13:3}
fills_all=[]; fills_exec=[]; cmds_exec=0; solid_color_exec=0
seg_p0 = segs["p0"]; seg_p1 = segs["p1"]
def span(a,b): return np.maximum(np.ceil(np.maximum(a,b))-np.floor(np.minimum(a,b)),1.0)
touch_per_seg = (span(seg_p0[:,0],seg_p1[:,0]) + span(seg_p0[:,1],seg_p1[:,1]) - 1).astype(np.int64)
rng = np.random.default_rng(0)
sample = rng.choice(n_tiles, 6000, replace=False)
for t in sample:
    ix = t*64+1
    seq=[]  # (tag, pos, nseg, segix)
    while True:
        tag = int(ptcl[ix])
        if tag==CMD_END: break
        if tag==CMD_JUMP: ix=int(ptcl[ix+1]); continue
        if tag==CMD_FILL: seq.append((tag,int(ptcl[ix+1])>>1,int(ptcl[ix+2])))
        elif tag==CMD_COLOR: seq.append((tag,int(ptcl[ix+1])>>24,0))
        else: seq.append((tag,0,0))
        ix += size.get(tag,1)
    # cull start: last SOLID followed by COLOR alpha 255 (no clips in this scene)
    start=0
    for i in range(len(seq)-1):
        if seq[i][0]==CMD_SOLID and seq[i+1][0]==CMD_COLOR and seq[i+1][1]==255: start=i
    for i,(tag,a,b) in enumerate(seq):
        if tag==CMD_FILL:
            tch = int(touch_per_seg[b:b+a].sum())
            fills_all.append((a,tch))
            if i>=start: fills_exec.append((a,tch))
    cmds_exec += len(seq)-start
fa=np.array(fills_all); fe=np.array(fills_exec)
print("tiles sampled", len(sample), "fills/tile all %.1f exec %.2f; cmds exec/tile %.1f" % (len(fa)/len(sample), len(fe)/len(sample), cmds_exec/len(sample)))
for name,f in (("all",fa),("executed",fe)):
    print(name, "segments/fill mean %.2f median %d p90 %d max %d | touches/fill mean %.1f median %d p90 %d" % (f[:,0].mean(), np.median(f[:,0]), np.percentile(f[:,0],90), f[:,0].max(), f[:,1].mean(), np.median(f[:,1]), np.percentile(f[:,1],90)))
    for k in (1,2,4,8,16,32):
        print("   fills with <= %2d segments: %.1f %%   touches <= %3d: %.1f %%" % (k, 100*(f[:,0]<=k).mean(), k*8, 100*(f[:,1]<=k*8).mean()))
