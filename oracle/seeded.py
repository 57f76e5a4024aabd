"""TEST INFRASTRUCTURE ONLY.

Deterministic (seeded, CPU generator) synthetic weights and inputs shared by the golden
fixture generator, the parity tests, `smoke()` and `bench.py`.  Key names and shapes follow
the reference's state dict; `oracle/make_golden.py` asserts they are
identical to what the reference model actually creates.
"""
import torch


def _conv_bn(entries, key, cin, cout, transposed=False, bias=False, bn=True, k=3):
    shape = (cin, cout, k, k, k) if transposed else (cout, cin, k, k, k)
    entries.append((key + ".0.weight", shape, "conv"))
    if bias:
        entries.append((key + ".0.bias", (cout,), "bias"))
    if bn:
        entries.append((key + ".1.weight", (cout,), "bn_w"))
        entries.append((key + ".1.bias", (cout,), "bn_b"))
        entries.append((key + ".1.running_mean", (cout,), "bn_m"))
        entries.append((key + ".1.running_var", (cout,), "bn_v"))
        entries.append((key + ".1.num_batches_tracked", (), "count"))


def _hourglass(entries, key, c, bias):
    _conv_bn(entries, key + ".conv1", c, 2 * c, bias=bias)
    _conv_bn(entries, key + ".conv2", 2 * c, 2 * c, bias=bias)
    _conv_bn(entries, key + ".conv3", 2 * c, 2 * c, bias=bias)
    _conv_bn(entries, key + ".conv4", 2 * c, 2 * c, bias=bias)
    _conv_bn(entries, key + ".conv5", 2 * c, 2 * c, transposed=True, bias=bias)
    _conv_bn(entries, key + ".conv6", 2 * c, c, transposed=True, bias=bias)


def aggregator_entries(kind="PSMNet", in_planes=64, prefix=""):
    """(key, shape, role) list of a PSMAggregator / AcfAggregator state dict.
    PSMNet: every conv bias=False (aggregators/PSMNet.py:30-53).  AcfNet: trunk convs keep
    the factory default bias=True, hourglass convs bias=False, plus deconv1..3
    (aggregators/AcfNet.py:30-57)."""
    acf = kind == "AcfNet"
    e = []
    p = prefix
    _conv_bn(e, p + "dres0.0", in_planes, 32, bias=acf)
    _conv_bn(e, p + "dres0.1", 32, 32, bias=acf)
    _conv_bn(e, p + "dres1.0", 32, 32, bias=acf)
    _conv_bn(e, p + "dres1.1", 32, 32, bias=acf)
    for name in ("dres2", "dres3", "dres4"):
        _hourglass(e, p + name, 32, bias=False)
    for name in ("classif1", "classif2", "classif3"):
        _conv_bn(e, p + name + ".0", 32, 32, bias=acf)
        e.append((p + name + ".1.weight", (1, 32, 3, 3, 3), "conv"))
    if acf:
        for name in ("deconv1", "deconv2", "deconv3"):
            e.append((p + name + ".weight", (1, 1, 8, 8, 8), "deconv_up"))
    return e


def module_entries(module):
    """(key, shape, role) list of ANY module made of Conv3d / ConvTranspose3d / BatchNorm3d children, in state-dict
    order -- the reference's GCAggregator / StereoNetAggregator and our mirrors yield the same list (same keys, same
    registration order), so one seed gives both the same weights without storing them in a fixture."""
    e = []
    for name, m in module.named_modules():
        p = name + "." if name else ""
        if isinstance(m, (torch.nn.Conv3d, torch.nn.ConvTranspose3d)):
            e.append((p + "weight", tuple(m.weight.shape), "conv"))
            if m.bias is not None:
                e.append((p + "bias", tuple(m.bias.shape), "bias"))
        elif isinstance(m, torch.nn.BatchNorm3d):
            for suffix, role in (("weight", "bn_w"), ("bias", "bn_b"), ("running_mean", "bn_m"), ("running_var", "bn_v")):
                e.append((p + suffix, (m.num_features,), role))
            e.append((p + "num_batches_tracked", (), "count"))
    return e


def seeded_state_dict(entries, seed=0, sharpen=1.0):
    """Fill `entries` with seeded values.  Conv weights ~ N(0, 2/fan) (so activations keep an
    O(1) scale through the stack), BN affine/statistics randomised so that folding is
    exercised, the final 32->1 classifier weights multiplied by
    `sharpen` (SURVEY.md section 7 'degenerate parity')."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, role in entries:
        if role == "conv":
            fan = shape[1] * shape[2] * shape[3] * shape[4]
            w = torch.randn(shape, generator=g) * (2.0 / fan) ** 0.5
            if shape[0] == 1 and len(shape) == 5 and ".1.weight" in key and "classif" in key:
                w = w * sharpen
            sd[key] = w
        elif role == "deconv_up":
            sd[key] = torch.randn(shape, generator=g) * 0.05 + 1.0 / 8.0
        elif role == "bias":
            sd[key] = torch.randn(shape, generator=g) * 0.05
        elif role == "bn_w":
            sd[key] = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif role == "bn_b":
            sd[key] = torch.randn(shape, generator=g) * 0.1
        elif role == "bn_m":
            sd[key] = torch.randn(shape, generator=g) * 0.1
        elif role == "bn_v":
            sd[key] = torch.rand(shape, generator=g) + 0.5
        elif role == "count":
            sd[key] = torch.tensor(1, dtype=torch.long)
        else:
            raise KeyError(role)
    return sd


def feature_pair(B, C, H, W, seed=0, scale=1.0, shift=None):
    """Synthetic left/right feature maps.  `shift=None`: independent N(0,scale).  Otherwise
    right = left shifted by `shift` pixels (a structured pair with known disparity) plus
    small noise."""
    g = torch.Generator().manual_seed(seed)
    left = torch.randn(B, C, H, W, generator=g) * scale
    if shift is None:
        right = torch.randn(B, C, H, W, generator=g) * scale
    else:
        right = torch.zeros_like(left)
        if shift > 0:
            right[..., :W - shift] = left[..., shift:]
        else:
            right = left.clone()
        right = right + torch.randn(B, C, H, W, generator=g) * scale * 0.01
    return left, right


def checksum(sd):
    """Order-independent float checksum of a state dict (guards RNG drift between boxes)."""
    tot = 0.0
    for k in sorted(sd):
        tot += float(sd[k].double().abs().sum())
    return tot
