import torch, sys
sys.path.insert(0,'/root/repo')
from hybrid_fem_lssvr_b200 import batch
E=10**7
nodes=batch.mesh_linspace(-1.0,1.0,E+1); u=torch.sin(3.141592653589793*nodes)
coef=torch.empty((E,9),dtype=torch.float64,device='cuda'); fine=torch.empty((E,32),dtype=torch.float64,device='cuda')
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
print('coef only   ms', t(lambda: batch.lssvr_primal_batch(nodes,u,9,1e4,N=12,F=0,want_coef=True,coef_out=coef)))
print('coef + fine ms', t(lambda: batch.lssvr_primal_batch(nodes,u,9,1e4,N=12,F=32,want_coef=True,coef_out=coef,want_fine=True,fine_out=fine)))
print('fine only   ms', t(lambda: batch.lssvr_primal_batch(nodes,u,9,1e4,N=12,F=32,want_coef=False,want_fine=True,fine_out=fine)))
x=torch.rand(10**7,dtype=torch.float64,device='cuda')*2-1
print('evaluate_points 1e7 pts ms', t(lambda: batch.evaluate_points(nodes,coef,x),n=3))
