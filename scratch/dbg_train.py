import sys, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from _helpers import golden
from oracle import clair_oracle as orc
import clair_torch_b200 as ct
from clair_torch_b200.training import linearity_loss_and_table_grad
np.set_printoptions(linewidth=200, precision=6)
for name in sys.argv[1:]:
    z=golden(name)
    val=torch.from_numpy(z['val']).cuda(); std=torch.from_numpy(z['std']).cuda() if 'std' in z else None
    i,j,r=ct.common.get_valid_exposure_pairs(torch.from_numpy(z['exposure']), float(z['thr']))
    print(name, i.tolist(), j.tolist(), r.tolist())
    for step in range(int(z['n_steps'])):
        theta = z['theta0'] if step==0 else z[f'theta_after_{step-1}']
        lin,sp,g=linearity_loss_and_table_grad(val,std,i,j,r,torch.from_numpy(theta).cuda(),1/255,254/255,bool(z['rel']),bool(z['unc']))
        print(step,'lin',lin.cpu().numpy(), z[f'linloss_{step}'])
        d=np.abs(sp.cpu().numpy()-z[f'spatial_{step}'])/np.abs(z[f'spatial_{step}'])
        print(' spatial relerr per pair', d.max(axis=1))
        print(' theta min/max', theta.min(), theta.max())
