import sys, torch
sys.path.insert(0,'/root/repo')
from zpc_b200 import api
pol=api.cuda_exec()
n=1<<(int(sys.argv[1]) if len(sys.argv)>1 else 26)
keys=torch.randint(-2**31,2**31-1,(n,),device='cuda',dtype=torch.int32); vals=torch.arange(n,device='cuda',dtype=torch.int32)
ko,vo=torch.empty_like(keys),torch.empty_like(vals)
for _ in range(2): pol.radix_sort_pair(keys,vals,ko,vo,kind='i32')
ones=torch.ones(n,device='cuda',dtype=torch.int32); out=torch.empty_like(ones)
for _ in range(2): pol.exclusive_scan(ones,out)
torch.cuda.synchronize()
