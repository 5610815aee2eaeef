"""ORACLE - TEST INFRASTRUCTURE ONLY.

The oracle's OWN statement of the reference state-dict schema (model/smap.py, 268 conv_bn_relu units in registration
order) and of the seeded weight / input generators.  Deliberately a separate file from the
product's smap_b200/schema.py: the checker must not import from the thing it checks.  tests/test_shims_schema_cpu.py pins
BOTH against the reference modules themselves (key set, shapes, construction-order RNG stream) and against each other
(bit-identical tensors), so "random-init SMAP weights" and "synthetic frame" mean the same bytes on both sides."""
import math

import torch

LAYERS = (3, 4, 6, 3)  # model/smap.py:300  (ResNet-50)
PLANES = (64, 128, 256, 512)
UP_IN = (2048, 1024, 512, 256)  # model/smap.py:249
BN_EPS = 1e-5


# ----------------------------------------------------------------------------
# schema
# ----------------------------------------------------------------------------
def unit_specs(stage_num=3, chl=256, kpt_paf=43, paf=14):
    """Ordered list of (prefix, cin, cout, k, stride, pad, relu, encoder) for all
    conv_bn_relu units, in the registration order of model/smap.py."""
    specs = [("top.conv", 3, 64, 7, 2, 3, True, False)]
    for s in range(stage_num):
        gen_skip = s != stage_num - 1
        pre = "stage%d." % s
        in_planes = 64
        for li, (planes, nblk) in enumerate(zip(PLANES, LAYERS)):
            stride = 1 if li == 0 else 2
            for b in range(nblk):
                p = "%sdownsample.layer%d.%d." % (pre, li + 1, b)
                st = stride if b == 0 else 1
                specs.append((p + "conv_bn_relu1", in_planes, planes, 1, 1, 0, True, True))
                specs.append((p + "conv_bn_relu2", planes, planes, 3, st, 1, True, True))
                specs.append((p + "conv_bn_relu3", planes, planes * 4, 1, 1, 0, False, True))
                if b == 0:
                    specs.append((p + "downsample", in_planes, planes * 4, 1, st, 0, False, True))
                in_planes = planes * 4
        for ind in range(4):
            p = "%supsample.up%d." % (pre, ind + 1)
            cin = UP_IN[ind]
            specs.append((p + "u_skip", cin, chl, 1, 1, 0, False, False))
            if ind > 0:
                specs.append((p + "up_conv", chl, chl, 1, 1, 0, False, False))
            if gen_skip:
                specs.append((p + "skip1", cin, cin, 1, 1, 0, True, False))
                specs.append((p + "skip2", chl, cin, 1, 1, 0, True, False))
            if ind == 3 and gen_skip:
                specs.append((p + "cross_conv", chl, 64, 1, 1, 0, True, False))
            specs.append((p + "res_conv1", chl, chl, 1, 1, 0, True, False))
            specs.append((p + "res_conv2", chl, kpt_paf, 3, 1, 1, False, False))
            specs.append((p + "res_d_conv1", chl, chl, 1, 1, 0, True, False))
            specs.append((p + "res_d_conv2", chl, paf, 3, 1, 1, False, False))
            specs.append((p + "res_rd_conv1", chl, chl, 1, 1, 0, True, False))
            specs.append((p + "res_rd_conv2", chl, 1, 3, 1, 1, False, False))
    return specs


def make_state_dict(seed=0, bn="identity", stage_num=3):
    """Deterministic random weights in the reference schema.

    bn="identity": gamma=1, beta=0, mean=0, var=1 everywhere: what SMAP(cfg)
        holds right after construction (model/smap.py:111-117 for the encoder,
        nn.BatchNorm2d defaults elsewhere).
    bn="random": non-trivial running stats and affine terms, to exercise folding.
    Encoder convs: kaiming_normal_(fan_out, relu) (model/smap.py:113-114), bias
    keeps nn.Conv2d's default U(+-1/sqrt(fan_in)); all other convs keep the
    nn.Conv2d defaults (kaiming_uniform_(a=sqrt(5)) == U(+-1/sqrt(fan_in))).
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for (name, cin, cout, k, _s, _p, _relu, enc) in unit_specs(stage_num):
        fan_in, fan_out = cin * k * k, cout * k * k
        if enc:
            w = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / fan_out)
        else:
            bound = 1.0 / math.sqrt(fan_in)
            w = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        bound = 1.0 / math.sqrt(fan_in)
        b = (torch.rand(cout, generator=g) * 2 - 1) * bound
        sd[name + ".conv.weight"] = w
        sd[name + ".conv.bias"] = b
        if bn == "identity":
            sd[name + ".bn.weight"] = torch.ones(cout)
            sd[name + ".bn.bias"] = torch.zeros(cout)
            sd[name + ".bn.running_mean"] = torch.zeros(cout)
            sd[name + ".bn.running_var"] = torch.ones(cout)
        else:
            sd[name + ".bn.weight"] = torch.rand(cout, generator=g) * 0.5 + 0.5
            sd[name + ".bn.bias"] = torch.randn(cout, generator=g) * 0.1
            sd[name + ".bn.running_mean"] = torch.randn(cout, generator=g) * 0.1
            sd[name + ".bn.running_var"] = torch.rand(cout, generator=g) + 0.5
        sd[name + ".bn.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd


def make_input(batch, h=512, w=832, seed=1):
    """uniform RGB in [0,1) normalised with the
    BGR mean/std of exps/stage3_root2/config.py:34-35."""
    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor([0.406, 0.456, 0.485]).view(1, 3, 1, 1)
    std = torch.tensor([0.225, 0.224, 0.229]).view(1, 3, 1, 1)
    return (torch.rand(batch, 3, h, w, generator=g) - mean) / std


