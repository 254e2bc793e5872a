"""CPU: the drop-in module surface (mnasnet-pytorch_b200/models) against the oracle's state_dict contract
(SURVEY.md T0), the header <-> library symbol contract, and host-side planning logic."""
import ctypes
import re

import pytest
import torch

from oracle import mnasnet_oracle as O


def _model(cfg='512', nc=1000, seed=42):
    from models.classifiers import FineTuneModelPool, load_model
    torch.manual_seed(seed)
    return FineTuneModelPool(load_model('mnasnet'), 'mnasnet', nc, cfg)


@pytest.mark.parametrize("cfg,nc", [('512', 1000), ('256', 74), ('512_256', 1000), ('320', 10)])
def test_state_dict_matches_reference_layout(cfg, nc):
    m = _model(cfg, nc)
    torch.manual_seed(42)
    sd = O.init_state_dict(nc, cfg)
    msd = m.state_dict()
    assert list(msd.keys()) == list(sd.keys())
    for k in sd:
        assert msd[k].shape == sd[k].shape
        assert torch.equal(msd[k], sd[k]), k          # same RNG stream -> bit-identical init

    def groups(d):
        g = {}
        for k, v in d.items():
            g.setdefault(v.data_ptr(), []).append(k)
        return sorted(tuple(v) for v in g.values())
    assert groups({k: v for k, v in msd.items() if v.numel() > 1}) == \
        groups({k: v for k, v in sd.items() if v.numel() > 1})
    assert [n for n, _ in m.named_parameters()] == O.unique_param_names(sd)


def test_param_count_and_aliases():
    m = _model()
    assert sum(p.numel() for p in m.parameters()) == 2218400
    assert len(list(m.parameters())) == 112
    seq = m.features[2].sequence
    assert seq[0] is seq[1] is seq[2]
    from models import mnasnet
    assert mnasnet._InvertedResidual is mnasnet.MBConv_block
    mm = mnasnet.MnasNet(10, 320)
    assert mm.classifier[1].out_features == 10


def test_load_state_dict_roundtrip_with_module_prefix():
    m = _model(seed=1)
    torch.manual_seed(42)
    sd = O.init_state_dict()
    m.load_state_dict(sd)
    assert torch.equal(m.features[0].conv.weight, sd["features.0.conv.weight"])
    wrapped = torch.nn.DataParallel(m) if False else None   # DataParallel needs CUDA to construct
    pref = {"module." + k: v for k, v in m.state_dict().items()}
    m2 = _model(seed=3)
    m2.load_state_dict({k[len("module."):]: v for k, v in pref.items()})
    assert torch.equal(m2.classifier[4].weight, m.classifier[4].weight)


def test_cut_channels_first_variant():
    from models.mnasnet import Mnasnet
    n = Mnasnet(cut_channels_first=True)
    assert sum(p.numel() for p in n.parameters()) == 2799728
    assert len(n.state_dict()) == 399


def test_freeze_unfreeze():
    m = _model()
    m.freeze()
    assert not any(p.requires_grad for p in m.features.parameters())
    assert all(p.requires_grad for p in m.classifier.parameters())
    m.unfreeze()
    assert all(p.requires_grad for p in m.parameters())


def test_library_exports_every_header_symbol():
    from mnb200 import _lib
    assert len(_lib.DECLS) >= 26
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.DECLS:
        assert hasattr(lib, name), name
    assert _lib.lib.mnb_version() >= 100
    assert _lib.lib.mnb_last_error() is not None


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = _model()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 3, 64, 64))


def test_product_does_not_import_oracle():
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mnasnet-pytorch_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle", src, re.M), \
                    os.path.join(dp, f)
