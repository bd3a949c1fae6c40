"""CPU-only checks of the host side: drop-in module tree / state_dict parity, C-ABI library exports, and the
no-fallback rule (the product path must raise without a CUDA device)."""
import ctypes
import os
import re

import pytest
import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pcss-unet_b200", "libnsm_b200.so")
HEADER = os.path.join(ROOT, "include", "nsm_b200.h")


def test_state_dict_matches_reference_layout():
    from Unetmodel import Unet
    torch.manual_seed(42)
    net = Unet()
    sd = net.state_dict()
    P = oracle.init_params(42)
    assert list(sd.keys()) == list(P.keys())          # 114 keys, reference order (golden-pinned in the oracle test)
    for k in sd:
        assert sd[k].shape == P[k].shape and sd[k].dtype == P[k].dtype, k
        assert torch.equal(sd[k], P[k]), k              # same default init under the same seed
    assert [n for n, _ in net.named_parameters()] == oracle.param_names()
    names = [n for n, _ in net.named_children()]
    assert names == ["conv2", "pool2", "conv3", "pool3", "conv4", "pool4", "conv5", "up6", "conv6", "up7", "conv7",
                     "up8", "conv8", "up9", "conv9", "conv10"]
    assert net.conv9.conv[3].p == pytest.approx(0.1) and net.conv2.conv[3].p == pytest.approx(0.2)
    # load_state_dict(strict) round trip, incl. checkpoint dict form used by infer.py:36-41
    net2 = Unet()
    net2.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build first: python pcss-unet_b200/build.py"
    declared = set(re.findall(r"\b(nsm_[a-z0-9_]+)\s*\(", open(HEADER).read()))
    declared -= {"nsm_conv_args"}
    lib = ctypes.CDLL(LIB)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    import nsm
    assert set(nsm.EXPORTS) == declared
    l = nsm.lib()
    assert l.nsm_version() >= 100
    # pure host-side size queries work without a GPU
    assert l.nsm_unet_packed_bytes(nsm.MODE_FP32) > l.nsm_unet_packed_bytes(nsm.MODE_BF16) > 15_000_000 * 2
    assert l.nsm_unet_workspace_bytes(1, 1080, 1920, nsm.MODE_FP32) > 0
    assert l.nsm_unet_workspace_bytes(1, 8, 8, nsm.MODE_FP32) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks behaviour WITHOUT a GPU")
def test_no_cpu_fallback():
    from Unetmodel import Unet
    import nsm
    net = Unet().eval()
    with pytest.raises(nsm.NsmError):
        net(torch.zeros(1, 4, 32, 32))
    from customLoss import CustomLoss
    with pytest.raises(nsm.NsmError):
        CustomLoss("cpu")(torch.rand(1, 1, 8, 8), torch.rand(1, 1, 8, 8), None)
    from customLoss import EnhancedCustomLoss
    with pytest.raises(nsm.NsmError):
        EnhancedCustomLoss("cpu", vgg_loss=None)(lambda x: x[:, :1], torch.rand(1, 1, 8, 8), torch.rand(1, 1, 8, 8),
                                                 torch.rand(1, 4, 8, 8))


def test_loss_classes_keep_the_reference_interfaces():
    """Constructor arguments and the attributes the reference's trainer reads (main.py:215,265-277,938-944;
    customLoss.py:196-201) -- host-side only, nothing is launched."""
    import inspect
    from customLoss import CustomLoss, EnhancedCustomLoss as EnhancedMSE
    from pert_loss import EnhancedCustomLoss, PerturbationLoss
    assert list(inspect.signature(CustomLoss.__init__).parameters)[:3] == ["self", "device", "alpha"]
    assert list(inspect.signature(EnhancedMSE.__init__).parameters)[:4] == ["self", "device", "alpha", "beta"]
    assert list(inspect.signature(EnhancedMSE.forward).parameters) == ["self", "model", "output", "target", "inputs"]
    c = CustomLoss("cpu", alpha=0.8, vgg_loss=None)
    assert c.alpha == 0.8 and callable(c.l1)
    e = EnhancedMSE("cpu", alpha=0.7, beta=0.02, vgg_loss=None)
    assert (e.alpha, e.beta) == (0.7, 0.02) and callable(e.l1) and e.vgg_loss is None
    p = EnhancedCustomLoss("cpu", alpha=0.9, perturb_weight=0.1)
    assert isinstance(p.perturbation_loss, PerturbationLoss)
    assert list(inspect.signature(p.forward).parameters)[:4] == ["model", "output", "target", "inputs"]
