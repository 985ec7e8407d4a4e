"""CPU: host-side mirror of the reference interface -- argument checks, drop-in installation,
shard arithmetic.  No CUDA calls."""
import sys
import types

import pytest
import torch

import smoe_stereo_b200 as sb
from smoe_stereo_b200 import sharding


def test_blocks_refuse_cpu_tensors_loudly():
    f = torch.zeros(1, 8, 2, 8)
    for cls in (sb.CorrBlock1D, sb.CorrBlockFast1D, sb.PytorchAlternateCorrBlock1D):
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            cls(f, f, num_levels=4, radius=4)
    with pytest.raises(RuntimeError, match="volume must be a CUDA tensor"):
        sb.corr_sampler.forward(torch.zeros(1, 2, 8, 8), torch.zeros(1, 1, 2, 8), 4)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        sb.corr_sampler.backward(torch.zeros(1, 2, 8, 8), torch.zeros(1, 1, 2, 8),
                                 torch.zeros(1, 9, 2, 8), 4)
    with pytest.raises(NotImplementedError):       # dead in the reference too (core/corr.py:161)
        sb.AlternateCorrBlock(f, f)


def test_reference_signatures():
    import inspect
    for cls in (sb.CorrBlock1D, sb.CorrBlockFast1D, sb.PytorchAlternateCorrBlock1D):
        ps = list(inspect.signature(cls.__init__).parameters.values())
        assert [p.name for p in ps] == ["self", "fmap1", "fmap2", "num_levels", "radius"]
        assert ps[3].default == 4 and ps[4].default == 4
        assert callable(cls.corr)
    assert [p for p in inspect.signature(sb.corr_sampler.forward).parameters] == \
        ["volume", "coords", "radius"]
    assert [p for p in inspect.signature(sb.corr_sampler.backward).parameters] == \
        ["volume", "coords", "corr_grad", "radius"]


def test_install_rebinds_reference_modules():
    fake = types.ModuleType("core.raft_stereo")
    fake.CorrBlock1D = object
    fake.CorrBlockFast1D = object
    fake.PytorchAlternateCorrBlock1D = object
    sys.modules["core.raft_stereo"] = fake
    try:
        done = sb.install()
        assert "core.raft_stereo.CorrBlock1D" in done
        assert fake.CorrBlock1D is sb.CorrBlock1D and fake.CorrBlockFast1D is sb.CorrBlockFast1D
        import corr_sampler            # the bare import of core/corr.py:5-8 now resolves to ours
        assert corr_sampler is sb.corr_sampler
        sb.uninstall()
        assert fake.CorrBlock1D is object and "corr_sampler" not in sys.modules
    finally:
        sys.modules.pop("core.raft_stereo", None)
        sb.uninstall()


def test_shard_bounds():
    assert [sharding.shard_bounds(500, 8, r) for r in range(8)] == \
        [(0, 63), (63, 126), (126, 189), (189, 252), (252, 314), (314, 376), (376, 438), (438, 500)]
    assert sharding.shard_sizes(8, 8) == [1] * 8
    assert sharding.shard_sizes(3, 4) == [1, 1, 1, 0]
    for n in (0, 1, 7, 135, 500):
        for w in (1, 2, 4, 8):
            b = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)
    t = torch.arange(2 * 3 * 5 * 4).view(2, 3, 5, 4)
    assert torch.equal(sharding.take_shard(t, "rows", 2, 1), t[:, :, 3:])
    assert torch.equal(sharding.take_shard(t, "batch", 2, 0), t[:1])


def _ref_models():
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from baseline import ref_models
    return ref_models


@pytest.mark.skipif(not _ref_models().available(), reason="baseline/_ref/core not installed (tools/install_ref.sh)")
def test_install_rebinds_every_reference_binding():
    """All eight names the reference binds by value point at this package after
    install(), and are restored by uninstall()."""
    import smoe_stereo_b200 as sb
    import sys
    R = _ref_models()
    R.import_core(with_ref_sampler=False)
    import core.corr as c
    import core.raft_stereo as rs
    import core.SMoEStereo as ss
    sb.uninstall()
    stock = c.CorrBlock1D
    sb.install()
    try:
        for mod in (c, rs, ss):
            for n in ("CorrBlock1D", "CorrBlockFast1D", "PytorchAlternateCorrBlock1D"):
                assert getattr(mod, n) is getattr(sb, n), (mod.__name__, n)
        assert c.CorrSampler is sb.CorrSampler
        assert sys.modules["corr_sampler"] is sb.corr_sampler and c.corr_sampler is sb.corr_sampler
    finally:
        sb.uninstall()
    assert c.CorrBlock1D is stock and rs.CorrBlock1D is stock


def test_install_output_ring_sets_and_clears_the_default():
    """install(output_ring=n) is the opt-in switch for the lookup output ring; uninstall() clears it."""
    import smoe_stereo_b200 as sb
    import smoe_stereo_b200.corr as c
    assert c._default_output_ring == 0
    try:
        sb.install(output_ring=3)
        assert c._default_output_ring == 3
        sb.install()                         # a plain install() resets the option
        assert c._default_output_ring == 0
        sb.set_default_output_ring(-5)       # negative depths mean "off"
        assert c._default_output_ring == 0
        sb.set_default_output_ring(2)
        assert c._default_output_ring == 2
    finally:
        sb.uninstall()
    assert c._default_output_ring == 0
