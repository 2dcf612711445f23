import importlib, sys, time, torch
sys.path.insert(0,'/root/repo')
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
dev=torch.device("cuda")
for N in (20000, 40000):
    src,tgt,R,t = dv.synthetic.make_batch("kitti",[0],N)
    torch.manual_seed(0)
    model = dv.DeepVCP(use_normal=False, npoint=N, r=2.0, s=0.4).to(dev).eval()
    st=(torch.tensor([1]),torch.tensor([2]),torch.tensor([3]))
    torch.cuda.synchronize(); t0=time.time()
    kp,vcp = model(src.to(dev),tgt.to(dev),R.to(dev),torch.zeros(1,3),starts=st,keep_stages=True)
    torch.cuda.synchronize(); dt=time.time()-t0
    L=model.last
    ok = torch.equal(L["src_fps"].cpu().long().sort(dim=1)[0], torch.arange(N).view(1,-1))
    d=L["knn_dist"].cpu()
    print(N, "forward s", round(dt,3), "fps perm", ok, "knn sorted", bool((d[...,1:]>=d[...,:-1]).all()), "vcp finite", bool(torch.isfinite(vcp).all()))
