"""Shared by tests/golden/make_golden_vgg.py (reference side) and the VGG parity tests (oracle / B200 side): the frozen
VGG19 both sides use -- ImageNet weights cannot be downloaded here, so a seeded random init stands in -- and
the seeded output / target pairs."""
import torch


def seeded_vgg19(weights=None, seed=1234, **kw):
    from torchvision.models.vgg import vgg19 as tv_vgg19
    dev = torch.get_default_device()
    torch.set_default_device("cpu")
    try:
        state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        net = tv_vgg19(weights=None, **kw)
        with torch.no_grad():
            for m in net.features:
                if isinstance(m, torch.nn.Conv2d):
                    m.bias.uniform_(-0.1, 0.1)          # torchvision zero-initialises biases: exercise the bias path
        torch.random.set_rng_state(state)
    finally:
        torch.set_default_device(dev)
    return net


def cases():
    g = torch.Generator().manual_seed(77)
    out = {}
    for tag, shape in (("a", (2, 1, 32, 48)), ("b", (1, 1, 64, 80))):
        o = torch.rand(*shape, generator=g)
        t = (o + 0.15 * torch.randn(*shape, generator=g)).clamp(0, 1)
        out[tag] = (o, t)
    return out
