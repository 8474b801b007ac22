import numpy as np, sys
sys.path.insert(0,'.')
from vello_b200.config import *
from vello_b200.encoding import *
from vello_b200.shapes import *
from vello_b200.renderer import Renderer
r=Renderer()
N=2
s=Scene()
s.fill(FILL_NON_ZERO,Affine.IDENTITY,Color.from_rgba8(240,240,240),None,Rect(0,0,64,64))
for k in range(N):
    s.push_clip_layer(FILL_NON_ZERO,Affine.IDENTITY,Rect(2+k,2+k,60-k,60-k))
s.fill(FILL_NON_ZERO,Affine.IDENTITY,Color.from_rgba8(200,0,0),None,Rect(0,0,64,20))
for k in range(N): s.pop_layer()
img=r.render_to_texture(resolve(s.encoding), RenderParams(BLACK,64,64,0))
print(img[0,0])
