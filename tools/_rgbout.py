import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import dev_tensor, new_product_store, seed_backward
from oracle import ref_ops as R
from test_ops_gpu import _opt
from biggan_tensorflow_b200 import ops, _lib
np.set_printoptions(linewidth=250)
for (B,H,Ci,tanh) in [(2,8,32,0),(4,64,32,0),(4,64,32,1),(4,32,32,0),(4,16,32,0),(1,64,16,0)]:
    rng=np.random.default_rng(0)
    x=rng.standard_normal((B,H,H,Ci)).astype(np.float32)
    po,ro=_opt()
    st=new_product_store(torch.float32)
    tx=dev_tensor(x)
    y=ops.conv(tx,3,po,kernel=3,stride=1,pad=1,use_bias=False,scope="G_logit",_act=_lib.ACT_TANH if tanh else _lib.ACT_NONE)
    dy=rng.standard_normal(y.shape).astype(np.float32)
    seed_backward(y,dy)
    rs=R.RefStore(torch.float64)
    xr=torch.tensor(x,dtype=torch.float64,requires_grad=True)
    yr=R.conv(rs,xr,3,ro,kernel=3,stride=1,pad=1,use_bias=False,scope="G_logit")
    if tanh: yr=torch.tanh(yr)
    w=rs.vars["G_logit/kernel"]
    gx,gw=torch.autograd.grad((yr*torch.as_tensor(dy,dtype=torch.float64)).sum(),[xr,w])
    e=np.abs(tx.grad.float().cpu().numpy()-gx.numpy())
    print((B,H,Ci,tanh),"fwd",np.abs(y.numpy()-yr.detach().numpy()).max()/np.abs(yr.detach().numpy()).max(),"gx",e.max()/np.abs(gx.numpy()).max(),
          "gw",np.abs(st.vars["G_logit/kernel"].grad.float().cpu().numpy()-gw.numpy()).max()/np.abs(gw.numpy()).max())
    em=e.max(axis=(0,3))/np.abs(gx.numpy()).max()
    if H==64: print((em[::4,::4]*1e6).astype(int))
