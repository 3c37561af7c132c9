"""FusedAdamW <-> torch.optim.AdamW checkpoint exchange (CPU: state handling only; the update itself is a CUDA kernel,
tests/test_gpu_parity.py::test_fused_adamw_matches_torch)."""
import pytest
import torch

from mli_nerf_b200.optim import FusedAdamW


def _params():
    torch.manual_seed(0)
    return [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]


def test_load_torch_adamw_checkpoint():
    ref_p = _params()
    ref = torch.optim.AdamW(ref_p, lr=2e-3, weight_decay=5e-2)
    for p in ref_p:
        p.grad = torch.randn_like(p)
    ref.step()
    sd = ref.state_dict()
    ours = FusedAdamW(_params(), lr=1e-3, weight_decay=1e-2, grad_scale=0.5)
    ours.load_state_dict(sd)
    g = ours.param_groups[0]
    assert g["lr"] == 2e-3 and g["weight_decay"] == 5e-2 and g["grad_scale"] == 0.5
    for p, q in zip(ours.param_groups[0]["params"], ref_p):
        st = ours.state[p]
        assert int(st["step"]) == 1
        assert torch.equal(st["exp_avg"], ref.state[q]["exp_avg"]) and torch.equal(st["exp_avg_sq"], ref.state[q]["exp_avg_sq"])
    # step() reaches the kernel call with every group key present (no KeyError); without a GPU it must refuse loudly
    for p in ours.param_groups[0]["params"]:
        p.grad = torch.randn_like(p)
    if not torch.cuda.is_available():
        with pytest.raises(Exception) as e:
            ours.step()
        assert not isinstance(e.value, KeyError)


def test_torch_adamw_loads_our_checkpoint():
    ours = FusedAdamW(_params(), lr=3e-3, weight_decay=2e-2)
    for p in ours.param_groups[0]["params"]:  # state as step() creates it
        ours.state[p] = dict(step=torch.tensor(4.0), exp_avg=torch.randn_like(p), exp_avg_sq=torch.rand_like(p))
    sd = ours.state_dict()
    ref_p = _params()
    ref = torch.optim.AdamW(ref_p, lr=1e-3)
    ref.load_state_dict(sd)
    assert ref.param_groups[0]["lr"] == 3e-3 and ref.param_groups[0]["weight_decay"] == 2e-2
    for p in ref_p:
        p.grad = torch.randn_like(p)
    ref.step()  # torch's update runs with our groups: all of its keys are present
    assert int(ref.state[ref_p[0]]["step"]) == 5


@pytest.mark.parametrize("flag", ["amsgrad", "maximize"])
def test_unsupported_checkpoint_flags_raise(flag):
    ref = torch.optim.AdamW(_params(), **{flag: True})
    ours = FusedAdamW(_params())
    with pytest.raises(ValueError):
        ours.load_state_dict(ref.state_dict())
