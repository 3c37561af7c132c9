"""C-ABI surface checks that need no GPU: the library builds/loads, exports every symbol include/mli_b200.h declares,
refuses to compute without a B200 (no CPU fallback), and the host-side mirror keeps the reference's error behaviour."""
import ctypes
import re

import pytest
import torch

from mli_nerf_b200 import _lib, config
from mli_nerf_b200.model import Model


def test_every_declared_symbol_is_exported(mli_lib):
    text = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)
    names = set(re.findall(r"\b(mli_\w+)\s*\(", text))
    assert len(names) >= 35
    lib = mli_lib.load()
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/mli_b200.h but not exported by libmli_b200.so"
    assert lib.mli_abi_version() == 4


def test_signature_table_covers_header(mli_lib):
    assert set(_lib._SIGS) >= {"mli_encode_rays", "mli_tc_linear", "mli_tc_wgrad", "mli_composite_bwd", "mli_losses_fwd_bwd"}
    # every compute entry point takes a stream last
    assert all(sig.endswith("s") for name, sig in _lib._SIGS.items() if name not in _lib.HOST_ONLY)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(mli_lib):
    lib = mli_lib.load()
    assert lib.mli_device_ok() == 0
    # a compute entry point must fail with MLI_ENODEV (-3) and a message, not compute on the host
    lib.mli_sample_coarse.restype = ctypes.c_int
    rc = lib.mli_sample_coarse(None, None, None, ctypes.c_int64(0), 1, None, ctypes.c_int64(1), None)
    assert rc == -3 and b"no CPU fallback" in lib.mli_last_error()
    cfg = config.experiment("syn_hotdog_b", dict_size=14)
    model = Model(cfg.model, cfg.data)  # constructing (e.g. to inspect a checkpoint) is fine ...
    with pytest.raises(_lib.MliError):  # ... rendering is not
        model(dict(pose=torch.eye(3, 4)[None], intr=torch.eye(3)[None], pose_light=torch.eye(3, 4)[None],
                   ray_idx=torch.zeros(1, 8, dtype=torch.long), idx=torch.zeros(1, dtype=torch.long)))
    with pytest.raises(_lib.MliError):
        _lib.call("mli_sample_coarse", torch.zeros(4), torch.zeros(4), None, 4, 1, torch.zeros(4), 1, 0)


def test_host_mirror_error_behaviour():
    cfg = config.experiment("syn_hotdog_b", dict_size=14)
    cfg.model.object.sdf.gradient.taps = 5
    with pytest.raises(ValueError, match="Only support 4 or 6 taps"):  # modules.py:177
        Model(cfg.model, cfg.data)
    cfg = config.experiment("syn_hotdog_b", dict_size=14)
    cfg.model.background.enabled = True
    with pytest.raises(NotImplementedError):  # NeuralLumen/model.py:247-249
        Model(cfg.model, cfg.data)
    cfg = config.experiment("syn_hotdog_b", dict_size=14)
    cfg.model.object.rgb.network_mode = "bogus"
    with pytest.raises(NotImplementedError):
        Model(cfg.model, cfg.data)


def test_trainer_facing_attributes():
    cfg = config.experiment("syn_hotdog_a", dict_size=14)
    m = Model(cfg.model, cfg.data)
    sdf = m.neural_sdf
    sdf.warm_up_end = 5000
    sdf.set_active_levels(5000 + 5000 * 9)
    sdf.set_normal_epsilon()
    assert sdf.anneal_levels == 9 and sdf.active_levels == 9 and abs(sdf.normal_eps - 1.0 / sdf.resolutions[8]) < 1e-12
    sdf.set_active_levels(0)
    assert sdf.anneal_levels == 1 and sdf.active_levels == 8
    assert hasattr(m, "s_var") and m.progress == 1.0 and abs(float(sdf.growth_rate) - 2 ** 0.4) < 1e-9
    cfgb = config.experiment("rene_savannah_b", dict_size=14)
    mb = Model(cfgb.model, cfgb.data)
    assert mb.bounding_type == "box" and tuple(mb.bounding_box_aabb.shape) == (6,)
    groups = mb.get_param_groups(cfgb.optim)
    names = {n for n, p in mb.named_parameters() if any(p is q for q in groups)}
    assert names and all("neural_rgb" in n for n in names)


def test_fused_adamw_has_no_cpu_path():
    """The optimizer is part of the product: parameters that are not contiguous fp32 CUDA tensors are refused, not
    updated on the host."""
    from mli_nerf_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    with pytest.raises(_lib.MliError):
        FusedAdamW([p], lr=1e-3).step()
    assert float(p.detach().abs().max()) == 0.0
    with pytest.raises(ValueError):
        FusedAdamW([p], lr=-1.0)


def test_neural_sdf_query_delegates_and_model_copies_cleanly():
    """`model.neural_sdf.sdf(x)` is what the reference's mesh script calls (scripts/extract_mesh.py:101): it must reach the
    owning Model's engine (here: refuse on CPU, no fallback); the back-reference must survive deepcopy / pickling and must
    not leak into the state_dict."""
    import copy
    import io
    cfg = config.experiment("syn_hotdog_b", dict_size=14)
    model = Model(cfg.model, cfg.data)
    assert not [k for k in model.state_dict() if "owner" in k]
    if not torch.cuda.is_available():
        with pytest.raises(_lib.MliError):
            model.neural_sdf.sdf(torch.zeros(4, 3))
    clone = copy.deepcopy(model)
    assert clone.neural_sdf.__dict__["_owner"]() is clone and model.neural_sdf.__dict__["_owner"]() is model
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    again = torch.load(buf, weights_only=False)
    assert again.neural_sdf.__dict__["_owner"]() is again
    assert all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), again.state_dict().values()))
