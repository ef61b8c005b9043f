"""CPU oracle of the evaluation step — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates `re_assign_label` + `evaluate_nau_segmentation_v2` of the reference
(/root/reference/source_code/utils_20231218.py:100-234, 410-453) with plain torch reductions on CPU, exactly as the
reference computes them (element-wise comparisons and sums, `np.spacing(1)` in the denominators).  Pinned against the
unmodified reference functions by tests/golden/make_eval_golden.py -> tests/golden/eval_kat.npz
(tests/test_oracle_golden.py::test_eval_oracle_matches_reference_golden)."""
import numpy as np
import torch

EPS = np.spacing(1)


def acc(preds, targets):                                   # _acc :100-117
    return (preds == targets).sum().item() / float(preds.numel())


def miou(preds, targets, num_k=2):                         # _miou :119-154
    total, nums = 0.0, 0
    for k in range(num_k):
        gt_elem, pd_elem = targets == k, preds == k
        if gt_elem.sum() == 0 and pd_elem.sum() == 0:
            total += 1.0
        elif gt_elem.sum() == 0 or pd_elem.sum() == 0:
            total += 0.0
        else:
            total += float(torch.logical_and(gt_elem, pd_elem).sum()) / float(torch.logical_or(gt_elem, pd_elem).sum())
        nums += 1
    return total / nums


def target_iou(preds, targets):                            # _target_iou :156-172
    return float(torch.logical_and(targets, preds).sum() / (torch.logical_or(targets, preds).sum() + EPS))


def detection_rate(preds, targets):                        # _detection_rate :174-185
    return float(torch.sum((targets == 1) * (preds == 1)) / (torch.sum(targets == 1) + EPS))


def false_alarm_rate(preds, targets):                      # _false_alarm_rate :187-192
    return float(torch.sum((targets == 0) * (preds == 1)) / (torch.sum(targets == 0) + EPS))


def re_assign_label(predict_label, gt_label):              # re_assign_label :410-453 (accuracy criterion)
    reordered = 1 - predict_label
    return reordered if acc(predict_label, gt_label) < acc(reordered, gt_label) else predict_label


def evaluate(predict_label, gt_label):                     # evaluate_nau_segmentation_v2 :213-234
    p, g = predict_label.reshape(-1), gt_label.reshape(-1)
    return acc(p, g), miou(p, g), detection_rate(p, g), false_alarm_rate(p, g), target_iou(p, g)


# ---- evaluation loops of Train_Onet_on_simclutter_20250407.py, restated over the oracle forward -----------------------
def _means(rows):
    return tuple(float(v) for v in np.array(rows, dtype=np.float64).mean(axis=0))


def test_simclutter(st, loader):                           # test_simclutter :97-172 (verbose=0)
    from oracle import onet_oracle as orc
    rows = []
    with torch.no_grad():
        for X, label, _ in loader:
            _, _, _, _, S = orc.onet_forward(st, X, training=False)
            rows.append(evaluate(re_assign_label(orc.predict_label(S), label), label))
    return _means(rows)


def test_2nd_stage_simclutter(st1, st2, loader):           # test_2nd_stage_simclutter :296-390 (verbose=0)
    """Returns (stage-1 means, stage-2 means); the reference returns (acc2, miou2, dr2, far2, tiou1)."""
    from oracle import onet_oracle as orc
    rows1, rows2 = [], []
    with torch.no_grad():
        for X1, label, _ in loader:
            _, Vt1, _, Vd1, S1 = orc.onet_forward(st1, X1, training=False)
            raw1 = orc.predict_label(S1)
            pred1 = re_assign_label(raw1, label)
            rows1.append(evaluate(pred1, label))
            X2 = Vd1 if torch.equal(raw1, pred1) else Vt1                    # :328-331
            X2 = orc.tensor_normal_per_frame(X2)
            _, _, _, _, S2 = orc.onet_forward(st2, X2, training=False)
            rows2.append(evaluate(re_assign_label(orc.predict_label(S2), label), label))
    return _means(rows1), _means(rows2)
