import sys, time, os, json
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import debigulator_b200 as dbg
from debigulator_b200 import corpus
from bench import make_unique, pack
W=int(sys.argv[1]); H=int(sys.argv[2]); N=int(sys.argv[3]); U=int(sys.argv[4])
def gen(i):
    from debigulator_b200 import corpus
    f = os.environ.get('PNG_FILT', '4')   # forced Paeth by default; 'mix' = config 3's i % 6 - 1
    return corpus.png_cfg3(i if f == 'mix' else i*6+int(f)+1, W, H)
import bench
bench._gen=gen
def _g(i): return gen(i)
uniq=[gen(i) for i in range(U)]
dev=torch.device('cuda',0); ctx=dbg.Context(0)
offs,sizes,total=pack([u[0] for u in uniq],N)
h=np.zeros(total+64,np.uint8)
for i in range(N):
    b=uniq[i%U][0]; h[offs[i]:offs[i]+len(b)]=np.frombuffer(b,np.uint8)
rgba=W*H*4
s=torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
d_in=torch.from_numpy(h).to(dev); d_out=torch.zeros(N*rgba,dtype=torch.uint8,device=dev)
i64=lambda a: torch.from_numpy(np.asarray(a,np.uint64).view(np.int64)).to(dev)
a_off,a_sz=i64(offs),i64(sizes); o_off,o_cap=i64(np.arange(N,dtype=np.uint64)*np.uint64(rgba)),i64(np.full(N,rgba,np.uint64))
st=torch.zeros(N,dtype=torch.int32,device=dev)
def step(): ctx.png_device(d_in,a_off,a_sz,d_out,o_off,o_cap,st,int(sum(sizes)),N*rgba,stream=s.cuda_stream)
step(); torch.cuda.synchronize()
assert int(st.abs().sum())==0, st.tolist()[:8]
exp=torch.stack([torch.from_numpy(np.frombuffer(u[1],np.uint8).copy()) for u in uniq]).to(dev)
got=d_out.view(N,rgba)
for i in range(N): assert torch.equal(got[i],exp[i%U]), i
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record(); 
for _ in range(2): step()
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/2
ctx.profile_enable(True); step(); torch.cuda.synchronize()
groups={k:round(ctx.profile_read_tag(t)[0],3) for k,t in (('fx_sizes',ctx.PROF_FX_SIZES),('fx_expand',ctx.PROF_FX_EXPAND),('png_scan',ctx.PROF_PNG_SCAN),('png_unfilter',ctx.PROF_PNG_UNFILTER),('inflate',ctx.PROF_INFLATE))}
ctx.profile_enable(False)
print(json.dumps({'W':W,'H':H,'N':N,'filt':os.environ.get('PNG_FILT','4'),'groups_ms':groups,'split_max':os.environ.get('DBG_SPLIT_MAX_STREAMS','default'),'ms':ms,'Mpix_s':N*W*H/ms/1e3,'GBps':N*rgba/ms/1e6}))
