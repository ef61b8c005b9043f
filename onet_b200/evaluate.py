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
    call("onet_eval_confusion", ptr(Vt), ptr(Vd), ptr(gt), Vt.numel(), ptr(counts), torch.cuda.current_stream(Vt.device).cuda_stream, device=Vt.device)
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


def normalize_per_frame(x):
    """`tensor_normal_per_frame` (utils_20231218.py:673-689) on the device: every (image, channel) frame of the 4-D fp32
    tensor scaled to [0,1] by its own min / max.  One C-ABI call (min/max reduction + scaling kernels)."""
    assert x.dim() == 4
    if not x.is_cuda:
        raise RuntimeError("onet_b200.evaluate has no CPU path")
    x = x.contiguous().float()
    nb, nc, h, w = x.shape
    out = torch.empty_like(x)
    work = torch.empty(2 * nb * nc, dtype=torch.int32, device=x.device)
    call("onet_normalize_per_frame", ptr(x), nb * nc, h * w, ptr(work), ptr(out), torch.cuda.current_stream(x.device).cuda_stream, device=x.device)
    return out


def _rows_from_counts(counts):
    """[(acc, miou, dr, far, t_iou)] per batch from the stacked [n_batches, 4] confusion counts: ONE device -> host copy
    for the whole evaluation loop (the reference syncs ~12 times per batch, Train_Onet_on_simclutter_20250407.py:109-147)."""
    rows = []
    for c in torch.stack(counts).tolist():
        m = segmentation_metrics(c, reassign=True)
        rows.append((m["acc"], m["miou"], m["dr"], m["far"], m["t_iou"]))
    return rows


def _device_of(config, onet):
    dev = getattr(config, "device", None)
    return torch.device(dev) if dev is not None else next(onet.parameters()).device


def test_simclutter(str_txt, config, onet, test_loader, verbose=0, measure_snr=False):
    """Same arguments and return value as the reference's `test_simclutter`
    (Train_Onet_on_simclutter_20250407.py:97-172): eval-mode forward of every (X, label, psnr) batch, label =
    predict_label re-assigned against the ground truth, five metrics per batch, means over the batches ->
    (acc, miou, dr, far, tiou).  Plotting / saving of a random batch (`verbose`) is not part of the path."""
    import numpy as np
    onet.eval()
    counts = []
    dev = _device_of(config, onet)
    with torch.no_grad():
        for X, label, _ in test_loader:
            _, Vt, _, Vd, _ = onet(X.to(dev, non_blocking=True))
            counts.append(confusion_counts(Vt, Vd, label.to(dev, non_blocking=True)))      # stays on the device
    rows = _rows_from_counts(counts)
    return tuple(float(v) for v in np.array(rows, dtype=np.float64).mean(axis=0))


def test_2nd_stage_simclutter(str_txt, config, onet, onet2nd, test_loader, verbose=0, return_all=False):
    """Two-stage cascade, same arguments and return value as the reference's `test_2nd_stage_simclutter`
    (Train_Onet_on_simclutter_20250407.py:296-390).  Per batch: stage-1 forward; the response map that represents the
    foreground (Vd1 when `re_assign_label` kept the labels, Vt1 when it flipped them, :328-331) is normalised per frame
    on the device and fed to the second Onet; both stages are scored against the ground truth.  Returns
    (acc2, miou2, dr2, far2, tiou1) like the reference (:390); return_all=True returns both stages' five means."""
    import numpy as np
    onet.eval()
    onet2nd.eval()
    counts1, counts2 = [], []
    dev = _device_of(config, onet)
    with torch.no_grad():
        for X1, label, _ in test_loader:
            label = label.to(dev, non_blocking=True)
            _, Vt1, _, Vd1, _ = onet(X1.to(dev, non_blocking=True))
            c1 = confusion_counts(Vt1, Vd1, label)
            counts1.append(c1)
            # re_assign_label flips the prediction when that raises the pixel accuracy (:410-453); the flip decides which
            # response map is the foreground one - taken on the device, no host round trip inside the loop
            flipped = (c1[0] + c1[3]) < (c1[2] + c1[1])
            X2 = normalize_per_frame(torch.where(flipped, Vt1, Vd1))
            _, Vt2, _, Vd2, _ = onet2nd(X2)
            counts2.append(confusion_counts(Vt2, Vd2, label))
    rows1, rows2 = _rows_from_counts(counts1), _rows_from_counts(counts2)
    s1 = tuple(float(v) for v in np.array(rows1, dtype=np.float64).mean(axis=0))
    s2 = tuple(float(v) for v in np.array(rows2, dtype=np.float64).mean(axis=0))
    if return_all:
        return s1, s2
    return s2[0], s2[1], s2[2], s2[3], s1[4]
