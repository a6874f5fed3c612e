"""Drop-in for u2pl/utils/loss_helper.py: same names, positional signatures, return values and
side effects (target mutation, memory-bank mutation) as the reference, computed by the sm_100a
kernels of libu2pl_b200.so.  Reference lines are cited per function."""
import torch
import torch.nn as nn

from u2pl_b200 import contra as _contra
from u2pl_b200 import ops as _ops


def compute_unsupervised_loss(predict, target, percent, pred_teacher):
    """Reference loss_helper.py:30-48.  `target` is rewritten in place (unreliable pixels -> 255)."""
    return _ops.unsup_loss(predict, target, percent, pred_teacher, ignore=255)


def compute_contra_memobank_loss(rep, label_l, label_u, prob_l, prob_u, low_mask, high_mask, cfg, memobank,
                                 queue_prtlis, queue_size, rep_teacher, momentum_prototype=None, i_iter=0):
    """Reference loss_helper.py:51-235.  Returns (new_keys, loss); mutates memobank / queue_prtlis."""
    return _contra.compute_contra_memobank_loss(rep, label_l, label_u, prob_l, prob_u, low_mask, high_mask, cfg,
                                                memobank, queue_prtlis, queue_size, rep_teacher,
                                                momentum_prototype, i_iter)


def get_criterion(cfg):
    """Reference loss_helper.py:238-255."""
    cfg_criterion = cfg["criterion"]
    aux_weight = cfg["net"]["aux_loss"]["loss_weight"] if cfg["net"].get("aux_loss", False) else 0
    ignore_index = cfg["dataset"]["ignore_label"]
    cls = CriterionOhem if cfg_criterion["type"] == "ohem" else Criterion
    return cls(aux_weight, ignore_index=ignore_index, **cfg_criterion["kwargs"])


def _check_aux(preds, target):
    main_pred, aux_pred = preds
    h, w = target.size(1), target.size(2)
    assert len(preds) == 2 and main_pred.shape[2:] == aux_pred.shape[2:] == (h, w)
    return main_pred, aux_pred


class Criterion(nn.Module):
    """Mean CE with ignore_index (+ aux head) -- reference loss_helper.py:258-320."""

    def __init__(self, aux_weight, ignore_index=255, use_weight=False):
        super().__init__()
        if use_weight:
            raise NotImplementedError("Criterion(use_weight=True) (loss_helper.py:266-293) is not built; "
                                      "every shipped config sets use_weight: False")
        self._aux_weight = aux_weight
        self._ignore_index = ignore_index
        self.use_weight = use_weight

    def forward(self, preds, target):
        if self._aux_weight > 0:
            main_pred, aux_pred = _check_aux(preds, target)
            return (_ops.cross_entropy_mean(main_pred, target, self._ignore_index)
                    + self._aux_weight * _ops.cross_entropy_mean(aux_pred, target, self._ignore_index))
        assert preds.shape[2:] == target.shape[1:]
        return _ops.cross_entropy_mean(preds, target, self._ignore_index)

    def forward_lowres(self, preds_low, target):
        """Extension used by u2pl_b200.step.SemiStep only: the same loss from the LOW-resolution logits, the x4 bilinear
        up-sampling of train_semi.py:344-358 fused into the loss kernels (the 354 MB up-sampled tensor is never built)."""
        if self._aux_weight > 0:
            main_low, aux_low = preds_low
            return (_ops.upsampled_ce_mean(main_low, target, self._ignore_index)
                    + self._aux_weight * _ops.upsampled_ce_mean(aux_low, target, self._ignore_index))
        return _ops.upsampled_ce_mean(preds_low, target, self._ignore_index)


class OhemCrossEntropy2dTensor(nn.Module):
    """Online hard example mining CE -- reference loss_helper.py:451-531 (use_weight=False, reduce=False)."""

    def __init__(self, ignore_index=255, thresh=0.7, min_kept=256, use_weight=False, reduce=False):
        super().__init__()
        if use_weight or reduce:
            raise NotImplementedError("OHEM class weights / reduce=True are not built")
        self.ignore_index, self.thresh, self.min_kept = ignore_index, float(thresh), int(min_kept)

    def forward(self, pred, target):
        return _ops.ohem_cross_entropy(pred, target, self.thresh, self.min_kept, self.ignore_index)


class CriterionOhem(nn.Module):
    """Reference loss_helper.py:323-360."""

    def __init__(self, aux_weight, thresh=0.7, min_kept=100000, ignore_index=255, use_weight=False):
        super().__init__()
        self._aux_weight = aux_weight
        self._criterion1 = OhemCrossEntropy2dTensor(ignore_index, thresh, min_kept, use_weight)
        self._criterion2 = OhemCrossEntropy2dTensor(ignore_index, thresh, min_kept)

    def forward(self, preds, target):
        if self._aux_weight > 0:
            main_pred, aux_pred = _check_aux(preds, target)
            return self._criterion1(main_pred, target) + self._aux_weight * self._criterion2(aux_pred, target)
        assert preds.shape[2:] == target.shape[1:]
        return self._criterion1(preds, target)
