"""ICRF training — call-compatible with clair_torch/training/icrf_training.py:18-186.

One optimisation step = two fused kernel passes over the device-resident exposure stack:
  1. clair_pair_stats: per (pair, channel) spatial sums -> weighted means m[p, c] (icrf_training.py:105-133)
  2. clair_pair_grad : d(sum_c sqrt(sum_p m[p,c]^2)) / d table, in closed form     (icrf_training.py:136,148-149)
The (P, C)-sized algebra between them, the four curve penalties (768 numbers), Adam and update_icrf stay in torch.
"""
from typing import Optional

import torch
from torch.optim import Optimizer
from torch.utils.data import DataLoader

from .. import _native, kernels
from ..common.general_functions import get_valid_exposure_pairs
from ..models.base import ICRFModelBase
from ..inference._common import NATIVE_MODE, as_device, stage_batch
from ..common.errors import ArgumentTypeError


def linearity_loss_and_table_grad(images, stds, i_idx, j_idx, ratio_pairs, table, lower_valid_threshold,
                                  upper_valid_threshold, use_relative_linearity_loss, use_uncertainty_weighting,
                                  want_grad=True, row_base=None, reduce_fn=None, interp_mode=_native.INTERP_LINEAR):
    """Linearity loss per channel (C,) float64, spatial means (P, C), and d(sum_c loss_c)/d table (C, L) float64.

    `reduce_fn`, if given, is applied in place to the (P, C, 5) sums and to the (C, L) gradient: the data-parallel
    trainer passes an NCCL all-reduce here (spatial shards of one image add their sums).
    """
    if want_grad and table is not None and kernels.can_fuse_pair(images, stds, len(ratio_pairs), use_uncertainty_weighting):
        # one exposure pair, weights independent of the uncertainty: statistics and gradient in ONE pass over the stack and,
        # when sharded, ONE all-reduce (the upstream factor is a scalar per channel and is applied afterwards)
        fused = kernels.pair_fused(images, i_idx, j_idx, ratio_pairs, table, lower_valid_threshold, upper_valid_threshold,
                                   use_relative_linearity_loss, row_base=row_base, interp_mode=interp_mode)
        if reduce_fn is not None:
            reduce_fn(fused)
        return kernels.pair_fused_combine(fused, table.shape[0], table.shape[1])
    sums = kernels.pair_stats(images, stds, i_idx, j_idx, ratio_pairs, table, lower_valid_threshold,
                              upper_valid_threshold, use_relative_linearity_loss, use_uncertainty_weighting,
                              row_base=row_base, means_only=True, interp_mode=interp_mode)
    if reduce_fn is not None:
        reduce_fn(sums)
    linearity_loss, spatial, upstream, mean_for_grad = kernels.pair_upstream(sums)
    grad = None
    if want_grad:
        grad = kernels.pair_grad(images, stds, i_idx, j_idx, ratio_pairs, table, lower_valid_threshold,
                                 upper_valid_threshold, use_relative_linearity_loss, use_uncertainty_weighting,
                                 upstream, mean_for_grad, row_base=row_base, interp_mode=interp_mode)
        if reduce_fn is not None:
            reduce_fn(grad)
    return linearity_loss, spatial, grad


def train_icrf_step(icrf_model: ICRFModelBase, optimizers: list[Optimizer], images: torch.Tensor,
                    stds: Optional[torch.Tensor], exposures: torch.Tensor, *, use_relative_linearity_loss=True,
                    use_uncertainty_weighting=True, alpha=1.0, beta=1.0, gamma=1.0, delta=1.0,
                    lower_valid_threshold=1 / 255, upper_valid_threshold=254 / 255, exposure_ratio_threshold=0.1,
                    row_base=None, reduce_fn=None):
    """Body of the batch loop, icrf_training.py:105-156, for one device-resident batch.  Returns the detached
    per-channel loss (C,) (a 0-dim sum when a single optimiser is used, :145-146).

    Launch sequence: pair means -> upstream factors -> table gradient (+ finalize) -> curve penalties (value and
    gradient) -> one autograd edge from the table to the model parameters -> optimisers -> update_icrf.
    """
    i_idx, j_idx, ratio_pairs = get_valid_exposure_pairs(exposures, exposure_ratio_threshold)
    for optimizer in optimizers:
        optimizer.zero_grad()
    return _fused_train_step(icrf_model, optimizers, images, stds, i_idx, j_idx, ratio_pairs, use_relative_linearity_loss,
                              use_uncertainty_weighting, alpha, beta, gamma, delta, lower_valid_threshold,
                              upper_valid_threshold, row_base, reduce_fn)


def _fused_train_step(icrf_model, optimizers, images, stds, i_idx, j_idx, ratio_pairs, use_relative_linearity_loss,
                      use_uncertainty_weighting, alpha, beta, gamma, delta, lower_valid_threshold, upper_valid_threshold,
                      row_base=None, reduce_fn=None, table=None):
    """The step on the fused kernels, for a model in any InterpMode.  `table` overrides where the kernels read the curve
    from (the captured step reads a static copy, see GraphedTrainStep); the autograd edge is always icrf_model.icrf."""
    curve = icrf_model.icrf                                        # (C, L); a function of the parameters after update_icrf
    if table is None:
        table = curve.detach()
    connected = curve.requires_grad                                # False on the very first step
    linearity_loss, _, grad = linearity_loss_and_table_grad(
        images, stds, i_idx, j_idx, ratio_pairs, table, lower_valid_threshold, upper_valid_threshold,
        use_relative_linearity_loss, use_uncertainty_weighting, want_grad=connected, row_base=row_base,
        reduce_fn=reduce_fn, interp_mode=NATIVE_MODE[icrf_model.interpolation_mode])
    if grad is None:
        grad = torch.zeros(tuple(table.shape), dtype=torch.float64, device=table.device)
    penalties = kernels.curve_penalties(table, alpha, beta, gamma, delta, grad)     # adds d penalties / d table to grad
    loss = linearity_loss + penalties
    if connected:
        # sum_c loss_c back-propagated once: the C backward() calls of :148-149 accumulate exactly this
        curve.backward(grad.to(curve.dtype))
    for optimizer in optimizers:
        optimizer.step()
    icrf_model.update_icrf()
    if len(optimizers) == 1:
        loss = torch.sum(loss)
    return loss.detach()


def _capturable(optimizers) -> bool:
    """Optimisers whose step keeps all of its state on the device (torch's `capturable=True` family)."""
    return bool(optimizers) and all(all(pg.get("capturable", False) for pg in opt.param_groups) for opt in optimizers)


def _hyper_key(optimizers):
    """Every python-side hyper-parameter a captured optimiser step bakes into its kernels (a scheduler changes lr)."""
    return tuple(tuple((k, v if isinstance(v, (int, float, bool, tuple, type(None))) else id(v))
                       for k, v in sorted(pg.items()) if k != "params") for opt in optimizers for pg in opt.param_groups)


class GraphedTrainStep:
    """train_icrf_step for ONE device-resident batch, captured once as a CUDA graph and replayed.

    A step is ~35 launches (2 large kernels, 4 small ones, the autograd edge table -> parameters, the optimisers,
    update_icrf) and the GPU finishes them faster than Python can enqueue them: replaying the captured sequence makes the
    step GPU-bound.  Requirements: optimisers created with `capturable=True`, the model already
    connected to its parameters (at least one eager step has run).  `row_base` / `reduce_fn` are the
    row-band arguments of train_icrf_step: an NCCL all-reduce passed as `reduce_fn` is captured with the kernels.
    The kernels read the curve from a static copy that the captured sequence refreshes after update_icrf, and the
    autograd edge is rebuilt inside the capture so that its backward runs on the capturing stream.
    """

    def __init__(self, icrf_model: ICRFModelBase, optimizers: list[Optimizer], images: torch.Tensor,
                 stds: Optional[torch.Tensor], exposures: torch.Tensor, *, use_relative_linearity_loss=True,
                 use_uncertainty_weighting=True, alpha=1.0, beta=1.0, gamma=1.0, delta=1.0, lower_valid_threshold=1 / 255,
                 upper_valid_threshold=254 / 255, exposure_ratio_threshold=0.1, row_base=None, reduce_fn=None):
        if not _capturable(optimizers):
            raise ValueError("GraphedTrainStep needs optimisers created with capturable=True")
        if not icrf_model.icrf.requires_grad:
            raise RuntimeError("GraphedTrainStep: run one eager train_icrf_step first (the first step only connects the "
                               "table to the parameters)")
        # no reference to the batch is kept: the step replays only while a batch sits at the captured address (make_key)
        self.key = self.make_key(optimizers, images, stds, exposures)
        i_idx, j_idx, ratio_pairs = get_valid_exposure_pairs(exposures, exposure_ratio_threshold)
        self._table = icrf_model.icrf.detach().clone()
        args = (use_relative_linearity_loss, use_uncertainty_weighting, alpha, beta, gamma, delta, lower_valid_threshold,
                upper_valid_threshold)
        self.graph = torch.cuda.CUDAGraph()
        # let go of the eager autograd edge: its AccumulateGrad nodes are bound to the eager stream, and they would be
        # reused (and invalidate the capture) for as long as anything keeps them alive
        icrf_model._icrf = icrf_model.icrf.detach()
        with torch.cuda.graph(self.graph):
            icrf_model.update_icrf()                               # same values; a fresh edge on the capturing stream
            for optimizer in optimizers:
                optimizer.zero_grad()
            loss = _fused_train_step(icrf_model, optimizers, images, stds, i_idx, j_idx, ratio_pairs, *args,
                                      row_base=row_base, reduce_fn=reduce_fn, table=self._table)
            self._table.copy_(icrf_model.icrf.detach())
            self._loss = loss
            self._curve = icrf_model.icrf                          # lives in the graph's memory, refreshed by every replay
        # capture does not execute anything: give the model a valid table again until the first replay
        self._model = icrf_model
        icrf_model.update_icrf()

    @staticmethod
    def make_key(optimizers, images, stds, exposures):
        exp = tuple(float(t) for t in torch.as_tensor(exposures).reshape(-1).tolist())
        return (images.data_ptr(), tuple(images.shape), None if stds is None else stds.data_ptr(), exp, _hyper_key(optimizers))

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        self._model._icrf = self._curve
        return self._loss.clone()


def train_icrf(dataloader: DataLoader, batch_size: int, device, icrf_model: ICRFModelBase,
               optimizers: Optional[list[Optimizer]] = None, schedulers: Optional[list] = None,
               use_relative_linearity_loss: bool = True, use_uncertainty_weighting: bool = True, epochs: int = 150,
               patience: int = 300, alpha: float = 1.0, beta: float = 1.0, gamma: float = 1.0, delta: float = 1.0,
               lower_valid_threshold: float = 1 / 255, upper_valid_threshold: float = 254 / 255,
               exposure_ratio_threshold: float = 0.1, *, verbose: bool = True, use_cuda_graph: bool = True,
               code_max: Optional[float] = None) -> ICRFModelBase:
    """Training loop with the reference's signature, defaults, early stopping and scheduler handling.

    `use_cuda_graph`: a batch that stays at the same device address from one epoch to the next (a device-resident
    dataset) with capturable optimisers (the default ones are) is stepped through GraphedTrainStep from its third visit
    on; anything else takes the eager step.

    Batches may carry raw uint8 / uint16 camera codes (and a `datasets.StdSpec` for the std): they are normalised on the
    device (`code_max` defaults to 255 / 65535), see measure_linearity."""
    if not isinstance(dataloader, DataLoader):
        raise ArgumentTypeError(f"dataloader must be a torch DataLoader, got {type(dataloader)}")
    if not isinstance(icrf_model, ICRFModelBase):
        raise ArgumentTypeError(f"icrf_model must be an ICRFModelBase, got {type(icrf_model)}")
    if not isinstance(batch_size, int):
        raise ArgumentTypeError("batch_size must be an int")
    dev = as_device(device)
    channels = icrf_model.channels
    if batch_size == 1:
        raise ValueError("Batch size must be larger than 1.")
    if optimizers is None:
        # the reference's per-channel Adam(lr=1e-3); fused = one kernel per optimiser step, capturable = state on the device
        optimizers = [torch.optim.Adam(icrf_model.channel_params(c), lr=1e-3, amsgrad=False, capturable=True, fused=True)
                      for c in range(channels)]
    previous_lrs = [pg["lr"] for opt in optimizers for pg in opt.param_groups]
    if schedulers is None:
        schedulers = [None] * len(optimizers)
    if len(schedulers) != len(optimizers):
        raise ValueError(f"Mismatched number of optimizers: {len(optimizers)} and schedulers: {len(schedulers)}.")
    best_losses = [float("inf")] * channels
    epochs_without_improvement = [0] * channels
    icrf_model.train()
    icrf_model.plot_icrf()
    step_kw = dict(use_relative_linearity_loss=use_relative_linearity_loss, use_uncertainty_weighting=use_uncertainty_weighting,
                   alpha=alpha, beta=beta, gamma=gamma, delta=delta, lower_valid_threshold=lower_valid_threshold,
                   upper_valid_threshold=upper_valid_threshold, exposure_ratio_threshold=exposure_ratio_threshold)
    graphable = use_cuda_graph and _capturable(optimizers)
    seen, graphs = {}, {}                                          # batch key -> eager visits / captured step
    for epoch in range(epochs):
        running_loss = torch.zeros(channels if len(optimizers) > 1 else (), dtype=torch.float64, device=dev)
        for _, val_batch, std_batch, meta_batch in dataloader:
            images, stds = stage_batch(val_batch, std_batch, dev, code_max=code_max, expand_codes=True)
            if images.shape[0] < 2:
                if verbose:
                    print("Skipped batch due to single image.")
                continue
            exposures = meta_batch["exposure_time"]
            if graphable:
                key = GraphedTrainStep.make_key(optimizers, images, stds, exposures)
                if key not in graphs and seen.get(key, 0) >= 2 and icrf_model.icrf.requires_grad and len(graphs) < 8:
                    graphs[key] = GraphedTrainStep(icrf_model, optimizers, images, stds, exposures, **step_kw)
                if key in graphs:
                    running_loss += graphs[key]()
                    continue
                seen[key] = seen.get(key, 0) + 1
            running_loss += train_icrf_step(icrf_model, optimizers, images, stds, exposures, **step_kw)
        avg_loss = (running_loss / len(dataloader)).cpu().numpy().reshape(-1)
        if verbose:
            print(f"Epoch {epoch + 1} Loss: {avg_loss}")
        for c in range(min(channels, len(avg_loss))):
            if avg_loss[c] < best_losses[c]:
                best_losses[c] = avg_loss[c]
                epochs_without_improvement[c] = 0
            else:
                epochs_without_improvement[c] += 1
        if all(epochs_without_improvement[c] >= patience for c in range(min(channels, len(avg_loss)))):
            if verbose:
                print(f"Early stopping triggered for all channels (patience = {patience} epochs).")
            break
        for c, scheduler in enumerate(schedulers):
            if scheduler is not None:
                scheduler.step(avg_loss[min(c, len(avg_loss) - 1)])
        for i, optimizer in enumerate(optimizers):
            current_lr = optimizer.param_groups[0]["lr"]
            if current_lr != previous_lrs[i] and verbose:
                print(f"Optimizer {i} learning rate changed to: {current_lr}")
            previous_lrs[i] = current_lr
    return icrf_model
