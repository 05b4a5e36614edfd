import torch, sys
from pathlib import Path; sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof
dev=torch.device("cuda:0"); n=1<<28
g=torch.Generator(device=dev).manual_seed(2024)
x=torch.randn(n,device=dev,generator=g)*torch.exp2(torch.randint(-12,13,(n,),device=dev,generator=g).float())
for name,d,mode in (("e4m3 RNE",lof.create_p3109_params(8,4,bias=7),0),("e4m3 RoundUp(exact)",lof.create_p3109_params(8,4,bias=7),6),("e2m1 RNE",lof.create_p3109_params(4,2,bias=1),0),("half RNE",lof.create_half_params(),0)):
    fn=lambda: lof.virtual_round_strict(x,d,mode)
    fn(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(5): fn()
    b.record(); torch.cuda.synchronize()
    print("strict",name,"%.0f GB/s"%(8*n*5/(a.elapsed_time(b)*1e-3)/1e9))
