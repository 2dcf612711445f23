import importlib, sys, torch
sys.path.insert(0,'/root/repo')
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
dev=torch.device("cuda"); B,N=64,1024
src,tgt,R,t = dv.synthetic.make_batch("modelnet", list(range(B)), N)
src,tgt,R,t = src.to(dev),tgt.to(dev),R.to(dev),t.view(B,3,1).to(dev)
torch.manual_seed(0)
model = dv.DeepVCP(use_normal=True, npoint=N, r=0.8, s=0.4).to(dev).eval()
g=torch.Generator().manual_seed(7)
starts=(torch.randint(0,N,(B,),generator=g),torch.randint(0,64,(B,),generator=g),torch.randint(0,N,(B,),generator=g))
for _ in range(3): model(src,tgt,R,torch.zeros(1,3),starts=starts)
model.profile=True
acc={}
for _ in range(5):
    model(src,tgt,R,torch.zeros(1,3),starts=starts); torch.cuda.synchronize()
    for k,v in model.stage_times_ms().items(): acc[k]=acc.get(k,0)+v/5
print({k:round(v,3) for k,v in acc.items()}, "sum", round(sum(acc.values()),3))
