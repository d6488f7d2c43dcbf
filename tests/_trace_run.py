import sys, os
sys.path.insert(0, 'adapting-2d-vits-for-3d-point-cloud-understanding_b200')
import torch
from p3tok import ops
dev = torch.device('cuda:0')
for (M,K,N,relu,mx) in ((524288,256,512,True,False),(524288,384,768,True,False),(524288,768,384,False,True)):
    a = (torch.randn(M,K,device=dev)*0.5).bfloat16(); w=(torch.randn(N,K,device=dev)*0.1).bfloat16(); b=torch.randn(N,device=dev)
    ops.linear_bf16(a,w,b,relu,mx); torch.cuda.synchronize()
