import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import *
from gpu_utils import make_opt
np.set_printoptions(linewidth=200, precision=4, suppress=True)
n,t,d,h,o = 256,3,16,64,1
x,y,w = synthetic_problem(n,t,d,h,o,seed=h)
_, opt = make_opt(w,x,y,GOOGLE,"admm",use_tensor_cores=True)
s = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
opt._state['h'][1:] = torch.rand_like(opt._state['h'][1:])
opt.state_changed()
tt=2
ndbg = 2*49152//4 + 128*256 + 16
buf = torch.full((4*h*opt.ldn + ndbg,), 7.0, device="cuda")
opt._call("admm_debug_preact", opt._pp, tt, buf.data_ptr(), 2, s)
torch.cuda.synchronize()
z = buf[:4*h*opt.ldn].view(4,h,opt.ldn).cpu().numpy().astype(np.float64)   # [g][j][n]
hs = opt._state['h'][tt-1].cpu().numpy().astype(np.float64)   # [H][ldn]
xs = opt._x[tt-1].cpu().numpy().astype(np.float64)            # [D][ldn]
wx = opt._wx.cpu().numpy().astype(np.float64); wh = opt._wh.cpu().numpy().astype(np.float64)
full = np.einsum('kn,gkj->gjn', xs, wx) + np.einsum('kn,gkj->gjn', hs, wh)
hpart = np.einsum('kn,gkj->gjn', hs, wh)
xpart = np.einsum('kn,gkj->gjn', xs, wx)
last8 = np.einsum('kn,gkj->gjn', hs[56:], wh[:,56:])
last16 = np.einsum('kn,gkj->gjn', hs[48:], wh[:,48:])
def rel(a,b): return np.max(np.abs(a-b))/np.max(np.abs(b))
print("z stats", z.min(), z.max())
for name, ref in (("full",full),("hpart",hpart),("xpart",xpart),("last8",last8),("last16",last16)):
    print(name, "rel err", rel(z, ref), " corr", np.corrcoef(z.ravel(), ref.ravel())[0,1])
print("z[0,:3,:6]\n", z[0,:3,:6]); print("full[0,:3,:6]\n", full[0,:3,:6])
