import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, vkvg_b200 as v
from tests import scenes
dev=v.Device(4)
def render(cs, mode, size=256):
    s=v.Surface(dev,size,size); c=v.Context(s)
    if mode == 2:
        assert c.replay(*cs.arrays())==0; c.flush()
    else:
        v.lib().vkvg_b200_set_submit_decoder(mode)
        st=c.submit(*cs.arrays2()); assert st==0, st
    img=s.pixels(); c.close(); s.close(); return img
polys, cols = scenes.polygons_c2(40, 256, 1)
cs=v.CommandStream(); cs.set_fill_rule(1)
for p,c in zip(polys[:5],cols[:5]):
    cs.set_source_rgba(*[float(x) for x in c]); cs.polyline(p); cs.close_path(); cs.fill()
cm,ar=cs.arrays2(); print(cm[:6], [hex(int(x)) for x in cm[:6]], ar[:8])
for order in ((2,1,0),(1,),(0,1,2)):
    print(order, [int(render(cs,m)[...,3].astype(bool).sum()) for m in order], v.submit_counts())
