"""ORACLE - TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs; never by the
product path).

CPU restatement of the RefineNet post-processing step:
  * model/refinenet.py:5-38      RefineNet_base: 4 x (Linear -> BatchNorm1d(eval) -> ReLU) + Linear, 75 -> 160 -> 256 ->
                                 256 -> 128 -> 45, driven by the reference's state-dict keys (block.layerN.{0,1}.*, block.layer5.*)
  * exps/stage3_root2/test_util.py:102-131  lift_and_refine_3d_pose: root-relative 2D/3D input assembly (fp64 numpy, cast to
                                 fp32), network, root re-addition in float32, score column.
Pinned by tests/golden/refine_cases.npz, produced by the unmodified reference code (tests/golden/make_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

LAYERS = [(75, 160), (160, 256), (256, 256), (256, 128), (128, 45)]


def refine_keys():
    """[(key, shape)] of the reference RefineNet state dict, in registration order (model/refinenet.py:8-17)."""
    out = []
    for i, (k, n) in enumerate(LAYERS[:4], start=1):
        p = "block.layer%d." % i
        out += [(p + "0.weight", (n, k)), (p + "0.bias", (n,)), (p + "1.weight", (n,)), (p + "1.bias", (n,)),
                (p + "1.running_mean", (n,)), (p + "1.running_var", (n,)), (p + "1.num_batches_tracked", ())]
    out += [("block.layer5.weight", (45, 128)), ("block.layer5.bias", (45,))]
    return out


def mlp(sd, x):
    """model/refinenet.py:19-26 in eval mode.  x: float32 [n,75] -> float32 [n,45]."""
    for i in range(1, 5):
        p = "block.layer%d." % i
        x = F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"])
        x = F.batch_norm(x, sd[p + "1.running_mean"], sd[p + "1.running_var"], sd[p + "1.weight"], sd[p + "1.bias"],
                         False, 0.0, 1e-5)
        x = F.relu(x)
    return F.linear(x, sd["block.layer5.weight"], sd["block.layer5.bias"])


def refine_inputs(pred2d, pred3d, root_n=2):
    """test_util.py:103-114 -> float32 [n,75]."""
    n = pred3d.shape[0]
    inp = np.zeros((n, 15, 5), np.float64)
    inp[:, root_n, :2] = pred2d[:, root_n, :2]
    inp[:, root_n, 2:] = pred3d[:, root_n, :3]
    for i in range(n):
        for j in range(15):
            if j != root_n and pred3d[i, j, 3] > 0:
                inp[i, j, :2] = pred2d[i, j, :2] - pred2d[i, root_n, :2]
                inp[i, j, 2:] = pred3d[i, j, :3] - pred3d[i, root_n, :3]
    return inp.reshape(n, 75).astype(np.float32)


def refine(pred2d, pred3d, sd, root_n=2):
    """test_util.py:102-131.  pred2d float32 [n,15,4], pred3d float64 [n,15,4] -> float64 [n,15,4]."""
    n = pred3d.shape[0]
    if n == 0:
        return np.zeros((0, 15, 4), np.float64)
    score = np.ones((n, 15, 1), np.float64)
    score[pred3d[:, root_n, 3] == 0] = 0
    with torch.no_grad():
        pred = mlp(sd, torch.from_numpy(refine_inputs(pred2d, pred3d, root_n))).numpy().reshape(n, 15, 3)
    for i in range(n):
        for j in range(15):
            if j != root_n:
                pred[i, j] += pred3d[i, root_n, :3]  # float32 += float64 -> computed in double, stored as float32
            else:
                pred[i, j] = pred3d[i, j, :3]
    return np.concatenate([pred, score], axis=2)
