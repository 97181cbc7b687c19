import sys, ctypes, numpy as np, torch, json
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,R+'/tetra-nerf_b200'); os.makedirs(R+'/gpurun_out',exist_ok=True)
import bench
from tetranerf import cpp
from tetranerf.b200 import synthetic as syn
from tetranerf.b200.render import FusedRenderer, RenderSettings, _lib
dev=torch.device('cuda:0')
V,C,field=bench.make_workload(); params=bench.mlp_params()
tr=cpp.TetrahedraTracer(dev); dV,dC=torch.from_numpy(V).to(dev),torch.from_numpy(C).to(dev); tr.load_tetrahedra(dV,dC)
fr=FusedRenderer(tr); fr.set_field(torch.from_numpy(field).to(dev)); fr.set_weights(params)
st=RenderSettings.tetra_nerf()
o,d=syn.camera_rays(4096,seed=5); o=torch.from_numpy(o).to(dev); d=torch.from_numpy(d).to(dev)
for _ in range(3): fr.render(o,d,st)
torch.cuda.synchronize()
buf=torch.zeros(65001,dtype=torch.int64,device=dev)
_lib.tn_debug_set_timeline(ctypes.c_void_p(buf.data_ptr()))
fr.render(o,d,st); torch.cuda.synchronize()
_lib.tn_debug_set_timeline(ctypes.c_void_p(0))
b=buf.cpu().numpy().astype(np.uint64); n=int(b[0]); rec=b[1:1+n]
tag=(rec>>np.uint64(40)).astype(np.int64); clk=(rec&np.uint64(0xFFFFFFFFFF)).astype(np.int64)
warp=tag&31; ev=(tag>>5)&15; it=(tag>>9)&255; layer=(tag>>17)&7
t0=clk.min(); clk=clk-t0
cta=b[1000:1000+8*148].reshape(148,8).astype(np.int64); g0=cta[:,0].min()
dur=(cta[:,1]-cta[:,0])/1e3
print('per-CTA duration us pct', np.percentile(dur,[0,25,50,75,100]), 'tiles', np.unique(cta[:,3]))
o=np.argsort(dur)
print('slot-0 tiles by smid:', [(int(a),int(b)) for a,b in sorted(zip(cta[:,4],cta[:,3]))])
print('SM clock during kernel (GHz) pct:', np.percentile(cta[:,7]/(dur*1e3),[0,50,100]))
print('smid sorted by duration (fast->slow):', cta[o,4].tolist())
print('duration:', np.round(dur[o]).astype(int).tolist())
print('gather busy kcycles:', (cta[o,5]//1000).tolist())
print('gather wait kcycles:', (cta[o,6]//1000).tolist())
np.save('gpurun_out/timeline.npy', np.stack([warp,ev,it,layer,clk],1))
print("records",n,"span cycles",clk.max())
# steady-state phase durations (cycles) from the recorded (last = fine) pass
import collections
def summarize(w):
    sel=np.where(warp==w)[0]; sel=sel[np.argsort(clk[sel])]
    dur=collections.defaultdict(list); prev=None
    for i in sel:
        k=(int(ev[i]),int(layer[i]))
        if prev is not None: dur[(prev[0],prev[1],k[0],k[1])].append(int(clk[i]-prev[2]))
        prev=(k[0],k[1],int(clk[i]))
    return {f"{names.get(a,a)} L{b} -> {names.get(c,c)} L{d}": (int(np.median(v)), len(v)) for (a,b,c,d),v in sorted(dur.items())}
names={1:'mma:a_ready s0',2:'mma:a_ready s1',3:'mma:commit s0',4:'mma:commit s1',5:'gather start',6:'A0 done',7:'wait D',8:'D seen',9:'epi done',10:'mma:kb0 issued'}
for w in (0,8,17,20):
    print("warp",w)
    for k,v in summarize(w).items(): print("   ",k,v)
