"""CPU-side checks (no GPU): the C-ABI library builds, loads and exports every symbol include/clair_b200.h
declares; argument validation answers before any CUDA call; the host-side mirror of the reference interface
behaves like the reference (known answers, error conventions); the product path refuses to run without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from _helpers import golden, max_rel
from oracle import clair_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from clair_torch_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        _native.build()
    return _native


def test_header_symbols_are_exported(native):
    header = open(os.path.join(ROOT, "include", "clair_b200.h")).read()
    declared = set(re.findall(r"CLAIR_API\s+[\w\s\*]+?\b(clair_\w+)\s*\(", header))
    assert len(declared) >= 10
    lib = native.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(native.EXPORTED_SYMBOLS)
    assert lib.clair_abi_version() == native.ABI_VERSION
    for macro, value in (("CLAIR_MAX_FRAMES", native.MAX_FRAMES), ("CLAIR_MAX_CHANNELS", native.MAX_CHANNELS),
                         ("CLAIR_MAX_LUT", native.MAX_LUT), ("CLAIR_MAX_PAIRS", native.MAX_PAIRS),
                         ("CLAIR_INTERP_LOOKUP", native.INTERP_LOOKUP), ("CLAIR_INTERP_LINEAR", native.INTERP_LINEAR),
                         ("CLAIR_INTERP_CATMULL", native.INTERP_CATMULL)):
        assert int(re.search(rf"#define {macro} (\d+)", header).group(1)) == value


def test_argument_validation_without_gpu(native):
    lib = native.load()
    assert lib.clair_grad_workspace_bytes(3, 256) == 4 * 1024 * 2 * 3 * 258
    # null buffers are rejected before any CUDA call
    rc = lib.clair_hdr_merge_update(None, None, None, 5, None, 3, 256, 100, None, 1, None, None, None, 1, 1, None, 0, None, None)
    assert rc == -1 and b"null" in lib.clair_last_error()
    t = np.ones(70)
    buf = ctypes.create_string_buffer(64)
    addr = ctypes.addressof(buf)
    rc = lib.clair_hdr_merge_update(addr, None, t.ctypes.data_as(ctypes.c_void_p), 70, None, 3, 256, 4, None, 1, None, None,
                                    None, 1, 1, addr, 0, None, None)
    assert rc == -2 and b"limit" in lib.clair_last_error()          # more than CLAIR_MAX_FRAMES frames per batch
    rc = lib.clair_icrf_forward(addr, addr, addr, None, 1, 3, 4, 256, 7, None, None)
    assert rc == -3                                                  # unknown interpolation mode
    rc = lib.clair_icrf_forward(addr, addr, addr, None, 1, 9, 4, 256, 2, None, None)
    assert rc == -2                                                  # more than CLAIR_MAX_CHANNELS
    with pytest.raises(ValueError):
        native.check(rc, "clair_icrf_forward")
    assert lib.clair_copy_band_h2d(None, addr, 1, 64, 0, 64, None) == -1 and lib.clair_copy_band_h2d(addr, addr, 1, 64, 32, 64, None) == -1
    assert lib.clair_set_tuning(b"no_such_knob", 1) == -1
    import re
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "clair_b200.h")).read()
    listed = re.search(r"developer knob for kernel tuning experiments \(keys: ([^)]*)\)", header, re.S)
    keys = [k.strip(" */\n") for k in listed.group(1).replace("\n", " ").split(",")]
    assert {"hdr_waves", "stats_waves", "grad_waves", "aux_waves", "fwd_blocks", "dark_strip", "dark_rows"} <= set(keys)
    for key in keys:                                                 # every key the header lists
        assert lib.clair_set_tuning(key.encode(), 0) == 0, key


def test_descriptor_entry_points_validate_without_gpu(native):
    lib = native.load()
    assert lib.clair_hdr_merge(None, None) == -1 and b"null descriptor" in lib.clair_last_error()
    d = native.MergeDesc()
    d.struct_bytes = 8                                               # a caller built against another layout
    assert lib.clair_hdr_merge(ctypes.byref(d), None) == -1 and b"bytes" in lib.clair_last_error()
    d.struct_bytes = ctypes.sizeof(native.MergeDesc)
    d.code_bytes = 3
    assert lib.clair_hdr_merge(ctypes.byref(d), None) == -3
    d.code_bytes = 0
    assert lib.clair_hdr_merge(ctypes.byref(d), None) == -1          # null val / exposure
    buf = ctypes.create_string_buffer(4096)
    t = np.ones(4)
    d.val_dev = d.std_dev = d.radiance_dev = d.sigma_dev = ctypes.addressof(buf)
    d.exposure_host, d.n_frames, d.n_channels, d.lut_size, d.plane = t.ctypes.data, 4, 3, 256, 16
    d.theta_dev, d.interp_mode, d.is_first, d.is_final = ctypes.addressof(buf), 9, 1, 1
    assert lib.clair_hdr_merge(ctypes.byref(d), None) == -3 and b"interp_mode" in lib.clair_last_error()
    d.interp_mode, d.plane_stride = native.INTERP_LINEAR, 8
    assert lib.clair_hdr_merge(ctypes.byref(d), None) == -1 and b"plane_stride" in lib.clair_last_error()
    d.plane_stride = 0
    d.dark_dev, d.dark_std_dev, d.height, d.width = ctypes.addressof(buf), ctypes.addressof(buf), 3, 5    # odd width, H*W != plane
    assert lib.clair_hdr_merge(ctypes.byref(d), None) == -3 and b"clair_dark_field_mix" in lib.clair_last_error()
    d.dark_dev = None
    assert lib.clair_hdr_merge_staged(ctypes.byref(d), None, None, 4, None, None) == -1                    # no host stack
    assert lib.clair_hdr_merge_staged(ctypes.byref(d), ctypes.addressof(buf), ctypes.addressof(buf), 4, None, None) == -1
    assert b"copy_stream" in lib.clair_last_error()
    assert lib.clair_linearize(ctypes.addressof(buf), ctypes.addressof(buf), ctypes.addressof(buf), ctypes.addressof(buf),
                               ctypes.addressof(buf), 1, 3, 16, 256, native.INTERP_LOOKUP, None, None) == -3   # LOOKUP with std


def test_descriptor_layout_matches_the_header(native, tmp_path):
    """The header is plain C, and ctypes' MergeDesc has the size and field offsets a C compiler gives clair_merge_desc."""
    import subprocess
    fields = [name for name, _ in native.MergeDesc._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "clair_b200.h"\nint main(void) {\n'
                   '    printf("%zu\\n", sizeof(clair_merge_desc));\n'
                   + "".join(f'    printf("%zu\\n", offsetof(clair_merge_desc, {f}));\n' for f in fields)
                   + "    return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(native.MergeDesc)
    assert out[1:] == [getattr(native.MergeDesc, f).offset for f in fields]


def test_product_path_refuses_cpu():
    import clair_torch_b200 as ct
    from clair_torch_b200.datasets import ExposureStackDataset, custom_collate
    val, std, t = ct.synthetic.make_stack(3, 3, 8, 8)
    loader = DataLoader(ExposureStackDataset(list(val), list(std), list(t)), batch_size=3, collate_fn=custom_collate)
    with pytest.raises(RuntimeError, match="CUDA"):
        ct.compute_hdr_image(loader, "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        ct.measure_linearity(loader, torch.device("cpu"))
    with pytest.raises(RuntimeError, match="CUDA"):
        ct.kernels.icrf_forward(val, ct.synthetic.reference_curve(3))
    with pytest.raises(RuntimeError, match="CUDA"):
        ct.ICRFModelDirect()(val)
    src = open(os.path.join(ROOT, "clair_torch_b200", "kernels.py")).read() + open(
        os.path.join(ROOT, "clair_torch_b200", "_native.py")).read()
    assert "oracle" not in src.replace("no CPU or eager fallback", "")   # the product never imports the checker


def test_exposure_pairs_known_answers():
    import clair_torch_b200 as ct
    z = golden("known_answers")
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(z["pairs_exposure"]))
    assert i.tolist() == [0, 0, 1] and j.tolist() == [1, 2, 2] and torch.allclose(r, torch.tensor([0.5, 0.25, 0.5]))
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(z["pairs_exposure"]), 0.4)
    assert i.tolist() == [0, 1] and j.tolist() == [1, 2]
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(z["pairs6_exposure"]), 0.2)
    assert np.array_equal(i.numpy(), z["pairs6_i"]) and np.array_equal(r.numpy(), z["pairs6_r"]) and r.dtype == torch.float64
    with pytest.raises(TypeError):
        ct.common.get_valid_exposure_pairs("bad", 0.1)
    with pytest.raises(TypeError):
        ct.common.get_valid_exposure_pairs(torch.ones(5), "bad")


def test_compat_helpers_match_reference_fixtures():
    import clair_torch_b200 as ct
    from clair_torch_b200 import training
    z = golden("known_answers")
    m = ct.common.get_pairwise_valid_pixel_mask(torch.from_numpy(z["mask_stack"]), torch.from_numpy(z["mask_i"]),
                                                torch.from_numpy(z["mask_j"]), val_lower=0.1, val_upper=1.0)
    assert np.array_equal(m.numpy(), z["mask_expected"])
    with pytest.raises(ValueError):
        ct.common.get_pairwise_valid_pixel_mask(torch.ones(2, 2), torch.tensor([0]), torch.tensor([0]), val_lower=0.5,
                                                val_upper=0.1)
    assert torch.equal(training.gaussian_value_weights(torch.from_numpy(z["gw_x"])), torch.from_numpy(z["gw_30"]))
    w, _ = ct.common.weighted_mean_and_std(torch.from_numpy(z["wm_values"]), torch.from_numpy(z["wm_weights"]), None, dim=1,
                                           compute_std=False)
    assert torch.allclose(w, torch.from_numpy(z["wm_weighted"]))
    # the compat per-pixel loss functions reproduce a measure_linearity fixture end to end on the CPU
    g = golden("linearity_rel1_unc1_std0_model0")
    val = torch.from_numpy(g["val"])
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(g["exposure"]), 0.2)
    mask = ct.common.get_pairwise_valid_pixel_mask(val, i, j, None, 1 / 255, 254 / 255)
    loss, err = training.pixelwise_linearity_loss(val, i, j, r, None, True)
    mean, std, e = training.compute_spatial_linearity_loss(loss, err, training.combined_gaussian_pair_weights(val, i, j), mask)
    assert max_rel(mean.numpy(), g["mean"]) < 1e-12 and max_rel(std.numpy(), g["stddev"]) < 1e-12 and e is None


def test_curve_penalties_and_their_gradients():
    from clair_torch_b200 import training
    rng = np.random.default_rng(0)
    theta = (np.linspace(0, 1, 64)[None] ** np.array([[2.0], [2.4], [1.7]]) + rng.normal(0, 0.02, (3, 64))).astype(np.float32)
    th = torch.from_numpy(theta).to(torch.float64).requires_grad_(True)
    pens = [training.compute_monotonicity_penalty(th, per_channel=True), training.compute_range_penalty(th, per_channel=True),
            training.compute_endpoint_penalty(th, per_channel=True), training.compute_smoothness_penalty(th, per_channel=True)]
    want, gwant = orc.curve_penalties(theta)
    for p, w, gw in zip(pens, want, gwant):
        assert max_rel(p.detach().numpy(), w, 1e-12) < 1e-9
        (g,) = torch.autograd.grad(p.sum(), th, retain_graph=True)
        assert np.max(np.abs(g.numpy() - gw)) < 1e-9
    # values pinned by the reference's (transposed-layout) unit tests, tests/unit/training/test_losses.py:92-217
    mono = torch.tensor([[0.0, 0.5, 0.4, 1.0]])
    assert float(training.compute_monotonicity_penalty(mono)) == pytest.approx(0.01, rel=1e-5)
    assert float(training.compute_endpoint_penalty(torch.tensor([[0.1, 0.5, 0.9]]))) == pytest.approx(0.02, rel=1e-5)
    assert float(training.compute_range_penalty(torch.tensor([[-0.5, 0.5, 1.5]]))) == pytest.approx(1.0)
    assert float(training.compute_smoothness_penalty(torch.tensor([[0.0, 1.0, 0.0]]))) == pytest.approx(4.0)


def test_collate_sorts_by_exposure_and_handles_missing_std():
    from clair_torch_b200.datasets import ExposureStackDataset, custom_collate
    vals = [torch.full((3, 2, 2), float(k)) for k in range(3)]
    ds = ExposureStackDataset(vals, [v * 0.05 for v in vals], [0.4, 0.1, 0.2])
    idx, val, std, meta = custom_collate([ds[0], ds[1], ds[2]])
    assert idx.tolist() == [1, 2, 0] and meta["exposure_time"].tolist() == [0.1, 0.2, 0.4]
    assert meta["exposure_time"].dtype == torch.float64 and val.shape == (3, 3, 2, 2) and std is not None
    assert val[:, 0, 0, 0].tolist() == [1.0, 2.0, 0.0]
    ds2 = ExposureStackDataset(vals, None, [0.4, 0.1, 0.2])
    assert custom_collate([ds2[0], ds2[1]])[2] is None
    with pytest.raises(ValueError):
        ExposureStackDataset(vals, None, [0.1])


def test_error_conventions_match_reference():
    import clair_torch_b200 as ct
    from clair_torch_b200.datasets import ExposureStackDataset, custom_collate
    val, std, t = ct.synthetic.make_stack(3, 3, 8, 8)
    ds = ExposureStackDataset(list(val), list(std), list(t))
    model = ct.ICRFModelDirect()
    with pytest.raises(ValueError, match="batch_size of 1"):
        list(ct.linearize_dataset_generator(DataLoader(ds, batch_size=2, collate_fn=custom_collate), "cuda", model))
    with pytest.raises(ValueError, match="larger than 1"):
        ct.train_icrf(DataLoader(ds, batch_size=1, collate_fn=custom_collate), 1, "cuda", model)
    with pytest.raises(TypeError):
        ct.compute_hdr_image("not a loader", "cuda")
    # the reference's functions are @typechecked: the error a reference user catches is typeguard's (not a TypeError)
    import typeguard
    for bad_call in (lambda: ct.compute_hdr_image("not a loader", "cuda"),
                     lambda: ct.measure_linearity(DataLoader(ds, batch_size=3, collate_fn=custom_collate), "cuda", "yes"),
                     lambda: ct.train_icrf("not a loader", 2, "cuda", model),
                     lambda: ct.ICRFModelDirect(interpolation_mode="LINEAR")):
        with pytest.raises(typeguard.TypeCheckError):
            bad_call()
    with pytest.raises(TypeError):
        ct.ICRFModelDirect(interpolation_mode="LINEAR")
    with pytest.raises(TypeError):
        ct.compute_hdr_image(DataLoader(ds, batch_size=3, collate_fn=custom_collate), "cuda", model, None, object())
    m = ct.ICRFModelDirect(icrf=torch.zeros(2, 17))
    assert m.channels == 2 and m.n_points == 17 and m.icrf.shape == (2, 17)
    assert torch.equal(ct.ICRFModelDirect(8, 2, initial_power=2.0).icrf,
                       (torch.linspace(0, 1, 8) ** 2.0).unsqueeze(0).repeat(2, 1))
    assert [e.value for e in ct.InterpMode] == [1, 2, 3]


def test_spatial_statistics_from_sums():
    """(P, C, 5) sums -> mean / std / errmean, including the clamp(min=1e-8) corner of an all-masked pair."""
    from clair_torch_b200.inference.measure_linearity import spatial_statistics
    g = golden("linearity_rel1_unc1_std1_model1")
    i, j, r = orc.exposure_pairs(g["exposure"], 0.2)
    tm = orc._pair_terms(g["val"], g["std"], i, j, r, g["theta"], 1 / 255, 254 / 255, True, True, 0)
    mw = tm["mask"] * tm["wt"]
    sums = np.stack([mw.sum((2, 3)), (mw * tm["ell"]).sum((2, 3)), (mw * tm["ell"] ** 2).sum((2, 3)),
                     (tm["mask"] * tm["err"]).sum((2, 3)), tm["mask"].sum((2, 3))], axis=-1)
    sums[0, 0] = 0.0
    mean, std, err = spatial_statistics(torch.from_numpy(sums), True)
    assert max_rel(mean.numpy()[1:], g["mean"][1:]) < 1e-9 and max_rel(std.numpy()[1:], g["stddev"][1:]) < 1e-6
    assert max_rel(err.numpy()[1:], g["errmean"][1:]) < 1e-9
    assert mean[0, 0] == 0 and std[0, 0] == 0 and err[0, 0] == 0


def test_shard_row_base_matches_flat_index_rule():
    from clair_torch_b200 import kernels
    for c, h, w, r0 in ((3, 10, 7, 4), (3, 9, 12, 0), (4, 6, 5, 3), (2, 5, 5, 1), (1, 4, 4, 2)):
        rows_full = orc.curve_rows((1, c, h, w))
        base = kernels.shard_row_base(c, h, w, r0)
        for ch in range(c):
            assert base[ch] == rows_full[0, ch, r0, 0]


def test_cpu_baseline_threads_survive_torchrun_environment():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arms of bench.py must still use every host core
    (round 1's reference arm ran on one thread at N >= 2).  Checked in a child process with that environment."""
    import os
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); from oracle import c_oracle, reference_runner as rr; import torch; "
            "print(c_oracle.use_all_host_threads(), rr.use_all_host_threads(), len(os.sched_getaffinity(0)))" % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    omp, torch_threads, cpus = (int(x) for x in out.stdout.split())
    assert omp == cpus and torch_threads == cpus and cpus == (os.cpu_count() or 1)
