"""ORACLE - TEST INFRASTRUCTURE ONLY.

Plain-PyTorch fp32 restatement of the reference backbone forward
(model/smap.py:403-419, inference branch) driven directly by the reference's
state-dict schema (SURVEY.md section 8(a), 1876 keys).  It is written as a
functional graph walker, not as a module tree, so it shares no structure with
model/smap.py; parity with the real reference module is pinned by
tests/golden/make_golden.py (which imports /root/reference/model/smap.py in the
build container, runs it on seeded inputs and commits the outputs) and
tests/test_oracle_backbone.py.

Also restates the reference's random initialisation (model/smap.py:111-117 +
PyTorch defaults for every other conv) with an explicit generator so that
"random-init SMAP weights" means the same tensors on every box.
"""
import math

import torch
import torch.nn.functional as F

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


# ----------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------
def _unit(sd, name, x, stride=1, pad=0, relu=True):
    """conv_bn_relu (model/smap.py:13-45): conv(bias) -> BN(eval) -> optional ReLU."""
    y = F.conv2d(x, sd[name + ".conv.weight"], sd[name + ".conv.bias"], stride=stride, padding=pad)
    y = F.batch_norm(y, sd[name + ".bn.running_mean"], sd[name + ".bn.running_var"],
                     sd[name + ".bn.weight"], sd[name + ".bn.bias"], False, 0.0, BN_EPS)
    return F.relu(y) if relu else y


def _bottleneck(sd, p, x, stride, has_ds):
    """model/smap.py:66-77"""
    out = _unit(sd, p + "conv_bn_relu1", x)
    out = _unit(sd, p + "conv_bn_relu2", out, stride=stride, pad=1)
    out = _unit(sd, p + "conv_bn_relu3", out, relu=False)
    if has_ds:
        x = _unit(sd, p + "downsample", x, stride=stride, relu=False)
    return F.relu(out + x)


def _up(t, size):
    return F.interpolate(t, size=size, mode="bilinear", align_corners=True)


@torch.no_grad()
def smap_forward(sd, imgs, stage_num=3, return_all=False):
    """Inference branch of SMAP.forward (model/smap.py:403-419).

    sd: state dict in the reference schema (tensors on imgs.device, fp32).
    Returns (heatmap_2d [B,43,H/4,W/4], det_d [B,14,..], root_d [B,1,..]).
    """
    B, _, H, W = imgs.shape
    out_shape = (H // 4, W // 4)
    oh, ow = out_shape
    up_sizes = [(oh // 8, ow // 8), (oh // 4, ow // 4), (oh // 2, ow // 2), (oh, ow)]
    x = _unit(sd, "top.conv", imgs, stride=2, pad=3)  # model/smap.py:89
    x = F.max_pool2d(x, 3, 2, 1)  # model/smap.py:90
    skip1 = skip2 = None
    heads = None
    for s in range(stage_num):
        pre = "stage%d." % s
        gen_skip = s != stage_num - 1
        feats = []
        t = x
        for li, nblk in enumerate(LAYERS):  # model/smap.py:140-154
            for b in range(nblk):
                p = "%sdownsample.layer%d.%d." % (pre, li + 1, b)
                t = _bottleneck(sd, p, t, (1 if li == 0 else 2) if b == 0 else 1, b == 0)
            if s > 0:
                t = t + skip1[li] + skip2[li]
            feats.append(t)
        xs = feats[::-1]  # x4, x3, x2, x1
        up_x = None
        res, res_d, res_rd, sk1, sk2 = [], [], [], [], []
        cross = None
        for ind in range(4):  # model/smap.py:210-241
            p = "%supsample.up%d." % (pre, ind + 1)
            out = _unit(sd, p + "u_skip", xs[ind], relu=False)
            if ind > 0:
                u = _up(up_x, up_sizes[ind])
                out = out + _unit(sd, p + "up_conv", u, relu=False)
            out = F.relu(out)
            res.append(_up(_unit(sd, p + "res_conv2", _unit(sd, p + "res_conv1", out), pad=1, relu=False), out_shape))
            res_d.append(_up(_unit(sd, p + "res_d_conv2", _unit(sd, p + "res_d_conv1", out), pad=1, relu=False), out_shape))
            res_rd.append(_up(_unit(sd, p + "res_rd_conv2", _unit(sd, p + "res_rd_conv1", out), pad=1, relu=False), out_shape))
            if gen_skip:
                sk1.append(_unit(sd, p + "skip1", xs[ind]))
                sk2.append(_unit(sd, p + "skip2", out))
                if ind == 3:
                    cross = _unit(sd, p + "cross_conv", out)
            up_x = out
        skip1, skip2 = sk1[::-1], sk2[::-1]  # model/smap.py:281-282 (finest first)
        x = cross
        heads = (res, res_d, res_rd)
    res, res_d, res_rd = heads
    outputs_2d = res[3] + res[2] + res[1]  # model/smap.py:418
    if return_all:
        return outputs_2d, res_d[3], res_rd[3], heads
    return outputs_2d, res_d[3], res_rd[3]


def flip_merge(o2d, o2d_flip):
    """Flip-TTA merge, exps/stage3_root2/test.py:55-70.
    o2d_flip = model(flip(imgs)) BEFORE un-flipping; o2d is modified in place."""
    flip_order = [0, 1, 2, 9, 10, 11, 12, 13, 14, 3, 4, 5, 6, 7, 8]  # data_settings.py:22
    flip_channel = [0, 1, 2, 3, 10, 11, 12, 13, 14, 15, 4, 5, 6, 7, 8, 9,
                    22, 23, 24, 25, 26, 27, 16, 17, 18, 19, 20, 21]  # data_settings.py:33-34
    f = torch.flip(o2d_flip, dims=[-1])
    pair = flip_order + [15 + c for c in flip_channel]
    for i in range(len(pair)):
        if i >= 15 and (i - 15) % 2 == 0:
            o2d[:, i] += f[:, pair[i]] * -1
        else:
            o2d[:, i] += f[:, pair[i]]
    o2d[:, 15:] *= 0.5
    return o2d
