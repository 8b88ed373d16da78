import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
from _helpers import max_rel
from oracle import c_oracle as corc, clair_oracle as orc
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
from clair_torch_b200.inference.measure_linearity import spatial_statistics
DEV = "cuda:0"
for channels, n, h, w in ((1, 13, 9, 11), (4, 5, 10, 13), (3, 13, 16, 16), (3, 13, 9, 11), (3, 5, 9, 11)):
    val, std, _ = ct.synthetic.make_stack(n, channels, h, w, bits=16, seed=n + channels)
    t = 1e-3 * 1.3 ** np.arange(n)
    theta = ct.synthetic.reference_curve(channels)
    i, j, r = orc.exposure_pairs(t, 0.0)
    sums = kernels.pair_stats(val.to(DEV), std.to(DEV), i, j, r, theta.to(DEV), 1 / 255, 254 / 255, True, True)
    mean, sd, err = spatial_statistics(sums.cpu(), True)
    o_mean, o_sd, o_err = corc.pair_stats(val.numpy(), std.numpy(), i, j, r, theta.numpy())
    em = np.abs(mean.numpy() - o_mean) / np.maximum(np.abs(o_mean), 1e-30)
    print(channels, n, h, w, 'P', len(i), 'mean', em.max(), 'argmax pair', np.unravel_index(em.argmax(), em.shape), 'sd', max_rel(sd.numpy(), o_sd), 'err', max_rel(err.numpy(), o_err))
    cnt = sums[..., 4].cpu().numpy(); print('   min count', cnt.min(), 'worst pair count', cnt[np.unravel_index(em.argmax(), em.shape)], 'o_mean there', o_mean[np.unravel_index(em.argmax(), em.shape)])
# training trajectory
from clair_torch_b200.datasets import ExposureStackDataset, custom_collate
val, std, t = ct.synthetic.make_stack(6, 3, 64, 96, bits=8, seed=2)
model = ct.ICRFModelDirect(256, 3, initial_power=2.5).to(DEV)
opts = [torch.optim.Adam(model.channel_params(c), lr=1e-3) for c in range(3)]
for k in range(60):
    loss = ct.train_icrf_step(model, opts, val.to(DEV), std.to(DEV), torch.from_numpy(t), use_uncertainty_weighting=False, alpha=10.0, exposure_ratio_threshold=0.25)
    if k % 6 == 0: print(k, loss.tolist())
