import torch, json
dev=torch.device("cuda",0)
out={}
for n in (512, 768):
    t=torch.empty((n,n,n),dtype=torch.float64,device=dev)
    for name,fn in (("fill",lambda: t.fill_(1.5)),("zero",lambda: t.zero_())):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/10
        out[f"{name}_{n}"]={"ms":ms,"gbs":8.0*n**3/ms/1e6}
    del t
print(json.dumps(out))
