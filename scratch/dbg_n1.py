import numpy as np, torch, sys
sys.path.insert(0, "tests")
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
from oracle import c_oracle as corc
for n in (1, 2, 3):
    val, std, _ = ct.synthetic.make_stack(n, 3, 24, 40, bits=16, seed=n)
    t = 1e-3 * 1.19 ** np.arange(n)
    theta = ct.synthetic.reference_curve(3)
    rad, sig = kernels.hdr_merge_update(kernels.HdrMergeState(), val.cuda(), std.cuda(), t, theta.cuda(), True, True)
    o_rad, o_sig = corc.hdr_merge(val.numpy(), std.numpy(), t, theta.numpy(), True)
    r = np.abs(rad.cpu().numpy() - o_rad) / np.abs(o_rad); s = np.abs(sig.cpu().numpy() - o_sig) / np.abs(o_sig)
    i = np.unravel_index(np.argmax(s), s.shape)
    print(n, "rad", r.max(), "sig", s.max(), "at", i, "x", val[0][i].item(), "sig", sig.cpu().numpy()[i], o_sig[i])
