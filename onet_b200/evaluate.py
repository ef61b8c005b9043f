"""Evaluation step next to the hot path: `test_simclutter`'s per-batch
`predict_label -> re_assign_label -> evaluate_nau_segmentation_v2` (Train_Onet_on_simclutter_20250407.py:109-147,
utils_20231218.py:100-234, 410-453) as ONE on-device reduction to the 2 x 2 confusion counts plus host arithmetic on four
integers.  The reference makes ~12 reductions with `.item()` host syncs per batch; here one 32-byte read-back.
"""
import numpy as np
import torch

from ._lib import call, ptr

EPS = float(np.spacing(1))


def confusion_counts(Vt, Vd, gt):
    """int64[4] device tensor, counts[pred * 2 + gt] with pred = (Vd > Vt), gt in {0, 1} (any non-zero gt counts as 1)."""
    if not Vt.is_cuda:
        raise RuntimeError("onet_b200.evaluate has no CPU path")
    gt = gt.to(device=Vt.device, dtype=torch.long).contiguous()
    Vt, Vd = Vt.contiguous().float(), Vd.contiguous().float()
    assert Vt.numel() == Vd.numel() == gt.numel()
    counts = torch.zeros(4, dtype=torch.long, device=Vt.device)
    call("onet_eval_confusion", ptr(Vt), ptr(Vd), ptr(gt), Vt.numel(), ptr(counts), torch.cuda.current_stream(Vt.device).cuda_stream)
    return counts


def _metrics(c00, c01, c10, c11):
    """(acc, miou, dr, far, t_iou) of utils_20231218.py from counts c[pred][gt]."""
    n = c00 + c01 + c10 + c11
    acc = (c00 + c11) / float(n)                                     # _acc :116
    miou, nums = 0.0, 0
    for inter, gt_n, pd_n in ((c00, c00 + c10, c00 + c01), (c11, c01 + c11, c10 + c11)):     # _miou :119-148, classes 0 and 1
        if gt_n == 0 and pd_n == 0:
            miou += 1.0
        elif gt_n == 0 or pd_n == 0:
            miou += 0.0
        else:
            miou += inter / float(gt_n + pd_n - inter)
        nums += 1
    miou /= nums
    t_iou = c11 / (float(c01 + c10 + c11) + EPS)                     # _target_iou :156-172
    dr = c11 / (float(c01 + c11) + EPS)                              # _detection_rate :174-185
    far = c10 / (float(c00 + c10) + EPS)                             # _false_alarm_rate :187-192
    return acc, miou, dr, far, t_iou


def segmentation_metrics(counts, reassign=True):
    """`counts`: the int64[4] tensor / sequence of `confusion_counts` (summed over as many batches as wanted).
    reassign=True applies re_assign_label (:410-453): the prediction is flipped when that raises the pixel accuracy.
    Returns dict(acc, miou, dr, far, t_iou, flipped)."""
    c00, c01, c10, c11 = (int(v) for v in (counts.tolist() if hasattr(counts, "tolist") else counts))
    flipped = False
    if reassign:
        n = c00 + c01 + c10 + c11
        if (c00 + c11) / float(n) < (c10 + c01) / float(n):         # pred_acc < reor_acc with reordered = 1 - pred
            c00, c01, c10, c11 = c10, c11, c00, c01
            flipped = True
    acc, miou, dr, far, t_iou = _metrics(c00, c01, c10, c11)
    return dict(acc=acc, miou=miou, dr=dr, far=far, t_iou=t_iou, flipped=flipped)
