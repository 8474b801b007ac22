import numpy as np, sys
sys.path.insert(0,'.')
from vello_b200 import scenes
from vello_b200.config import *
from vello_b200.encoding import *
from vello_b200.shapes import *
from vello_b200.renderer import Renderer
from oracle.vbo import Oracle
r=Renderer(); o=Oracle(threads=4)
def run(s,w,h,aa=0,label=''):
    packed=resolve(s.encoding)
    img=r.render_to_texture(packed, RenderParams(BLACK,w,h,aa))
    ref=o.render(packed,w,h,BLACK.premul_rgba8_u32(),aa)
    d=np.abs(img.astype(int)-ref.astype(int))
    print(label,'maxdiff',d.max(),'ndiff px',(d.max(axis=2)>0).sum(), r.last_stats.as_dict()['blend'], r.last_stats.as_dict()['retries'])
    ys,xs=np.where(d.max(axis=2)>0)
    for k in range(0,len(ys),max(1,len(ys)//3)):
        print('   ',ys[k],xs[k],'gpu',img[ys[k],xs[k]],'ref',ref[ys[k],xs[k]])
    return img,ref
for N in range(1,8):
    s=Scene()
    s.fill(FILL_NON_ZERO,Affine.IDENTITY,Color.from_rgba8(240,240,240),None,Rect(0,0,64,64))
    for k in range(N):
        s.push_clip_layer(FILL_NON_ZERO,Affine.IDENTITY,Rect(2+k,2+k,60-k,60-k))
    s.fill(FILL_NON_ZERO,Affine.IDENTITY,Color.from_rgba8(200,0,0),None,Rect(0,0,64,20))
    for k in range(N): s.pop_layer()
    run(s,64,64,0,'nest%d'%N)
