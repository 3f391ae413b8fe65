import torch, sys
sys.path.insert(0,'/root/repo')
from hybrid_fem_lssvr_b200 import batch
E=10**7
nodes=batch.mesh_linspace(-1.0,1.0,E+1); u=torch.sin(3.141592653589793*nodes)
fine=torch.empty((E,32),dtype=torch.float64,device="cuda")
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
for store in (3,5):
  batch.set_option('primal_store',store)
  for dbg in ((0,1,2) if store==3 else (0,1,2,3)):
    batch.set_option('primal_debug',dbg)
    print('store',store,'debug',dbg,'ms',t(lambda: batch.lssvr_primal_batch(nodes,u,9,1e4,N=12,F=32,want_coef=False,want_fine=True,fine_out=fine)))
batch.set_option('primal_debug',0)
err3=batch.new_error_accumulator()
for store in (3,5):
    batch.set_option('primal_store',store)
    print('store',store,'fused err ms',t(lambda: batch.lssvr_primal_batch(nodes,u,9,1e4,N=12,F=32,want_coef=False,want_fine=True,fine_out=fine,err3=err3)))
batch.set_option('primal_store',0)
print('fill ms', t(lambda: fine.fill_(1.0)))
