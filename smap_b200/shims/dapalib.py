"""Drop-in replacement for the reference `dapalib` extension module (extensions/association.cpp:236-241).

    import dapalib
    dapalib.connect(hmsIn, rDepth, rootIdx=2, distFlag=True) -> Tensor
    dapalib.extract(hmsIn) -> (list[Tensor], list[Tensor])

Same names, argument meaning and return types as the reference: `hmsIn` is a contiguous CUDA fp32 tensor
[43,128,208] already divided by 255/127 (exps/stage3_root2/test.py:111-112), `rDepth` a fp32 [128,208] tensor
(the reference passes a CPU tensor; CUDA is accepted too).  Returns CPU tensors: `connect` -> [P,15,4]
(x, y, 0, score) in heat-map pixels sorted by ascending root depth, or a 1-D empty tensor when no root peak exists
(association.cpp:133-136); `extract` -> (15 tensors [n_j,3], 14 tensors [nA,nB]).

Put this directory in front of the reference's `extensions/` on PYTHONPATH.  Everything runs in
libsmap_b200.so on the tensor's device; there is no CPU fallback.
"""
import torch

from smap_b200.engine import Engine

_HM_SHAPE = (43, 128, 208)  # heatmapDim, extensions/association.cpp:21
_PAIRS = [0, 1, 0, 2, 0, 9, 9, 10, 10, 11, 0, 3, 3, 4, 4, 5, 2, 12, 12, 13, 13, 14, 2, 6, 6, 7, 7, 8]
_engines = {}


def _engine(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _engines:
        _engines[idx] = Engine(idx, max_batch=1, in_h=_HM_SHAPE[1] * 4, in_w=_HM_SHAPE[2] * 4)
    return _engines[idx]


def _check(hmsIn):
    if not (torch.is_tensor(hmsIn) and hmsIn.is_cuda and hmsIn.dtype == torch.float32):
        raise RuntimeError("dapalib: hmsIn must be a CUDA float32 tensor")  # the reference would crash in cudaMemcpy
    if tuple(hmsIn.shape) != _HM_SHAPE:
        raise RuntimeError("dapalib: hmsIn must have shape %s (heatmapDim, association.cpp:21)" % (_HM_SHAPE,))
    return hmsIn.contiguous()


def extract(hmsIn):
    hms = _check(hmsIn)
    with torch.cuda.device(hms.device):
        peaks, scores = _engine(hms.device).extract(hms[None])
        peaks, scores = peaks[0].cpu(), scores[0].cpu()
    pose = []
    for j in range(15):
        n = int(peaks[j, 0, 0])
        pose.append(peaks[j, 1:n + 1].clone())
    paf = []
    for l in range(14):
        nA, nB = pose[_PAIRS[2 * l]].shape[0], pose[_PAIRS[2 * l + 1]].shape[0]
        paf.append(scores[l, :nA, :nB].clone())
    return pose, paf


def connect(hmsIn, rDepth, rootIdx=2, distFlag=True):
    hms = _check(hmsIn)
    with torch.cuda.device(hms.device):
        rd = rDepth.to(hms.device, torch.float32).contiguous()
        bodies, counts = _engine(hms.device).connect(hms[None], rd[None], int(rootIdx), bool(distFlag))
        n = int(counts[0])
        if n == 0:
            return torch.empty(0)
        return bodies[0, :n].cpu()
