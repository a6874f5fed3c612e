"""oracle/step_port.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

CPU restatement of one semi-supervised training step of the reference driver
(train_semi.py:272-561, branch epoch >= sup_only_epoch), built from oracle/model_port.py
(network, plain torch fp32) and oracle/port.py (losses).  It is (a) the checker for
u2pl_b200/step.py in tests and (b) what `bench.py --impl reference` / the cpu_baseline leg time
on the host cores.  Nothing under u2pl_b200/ is imported here.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import port
from .model_port import Net


def _up(t, size):
    return F.interpolate(t, size, mode="bilinear", align_corners=True)



def _sup_criterion(cfg, pred, target, ignore=255):
    """get_criterion's per-head loss (loss_helper.py:238-262): mean CE, or OhemCrossEntropy2dTensor.forward
    (loss_helper.py:502-531: kept = mask_prob <= max(thresh, min_kept-th smallest mask_prob)) restated in torch."""
    crit = cfg.get("criterion", {"type": "CELoss"})
    if crit["type"] != "ohem":
        return F.cross_entropy(pred, target, ignore_index=ignore)
    thresh, min_kept = float(crit["kwargs"]["thresh"]), int(crit["kwargs"]["min_kept"])
    b, c, h, w = pred.shape
    t = target.reshape(-1)
    valid = t.ne(ignore)                                                        # :505
    t = t * valid.long()                                                        # :506
    num_valid = int(valid.sum())
    if not (min_kept > num_valid) and num_valid > 0:                            # :512-514
        with torch.no_grad():
            prob = F.softmax(pred, dim=1).transpose(0, 1).reshape(c, -1)        # :509-510
            prob = prob.masked_fill(~valid, 1)                                  # :516
            mask_prob = prob[t, torch.arange(t.numel(), device=t.device)]       # :517
            threshold = thresh
            if min_kept > 0:
                _, index = mask_prob.sort()                                     # :520
                kth = mask_prob[index[min(index.numel(), min_kept) - 1]]        # :521
                if kth > thresh:                                                # :522-523
                    threshold = kth
                kept = mask_prob.le(threshold)                                  # :524
                t = t * kept.long()
                valid = valid & kept                                            # :526
    t = t.masked_fill(~valid, ignore).view(b, h, w)                             # :528-529
    return F.cross_entropy(pred, t, ignore_index=ignore)

class ReferenceStep:
    def __init__(self, student_state, teacher_state, cfg, arch="resnet101", lr=0.001, momentum=0.9,
                 weight_decay=1e-4, head_lr_mult=10, bank_dim=256):
        net = cfg["net"]
        C = net["num_classes"]
        aux = bool(net.get("aux_loss", False))
        self.cfg, self.C, self.aux = cfg, C, aux
        self.student = Net(student_state, arch, C, aux)
        self.teacher = Net(teacher_state, arch, C, aux)
        enc = [v for k, v in student_state.items() if v.requires_grad and k.startswith("encoder.")]
        head = [v for k, v in student_state.items() if v.requires_grad and not k.startswith("encoder.")]
        self.opt = torch.optim.SGD([dict(params=enc, lr=lr), dict(params=head, lr=lr * head_lr_mult)],
                                   lr=lr, momentum=momentum, weight_decay=weight_decay)      # train_semi.py:97-112
        self.memobank = [[np.zeros((0, bank_dim), np.float32)] for _ in range(C)]             # :161-169
        self.queue_ptr = [[0] for _ in range(C)]
        self.queue_size = [30000] * C
        self.queue_size[0] = 50000

    def step(self, image_l, label_l, image_u, epoch, i_iter, len_loader):
        cfg, tr = self.cfg, self.cfg["trainer"]
        h, w = label_l.shape[1:]
        S, T = self.student, self.teacher
        sup_only = tr.get("sup_only_epoch", 1)
        # T1 (:317-324)
        T.training = False
        with torch.no_grad():
            p = _up(T.forward(image_u)["pred"], (h, w))
            logits_u_aug, label_u_aug = torch.max(F.softmax(p, dim=1), dim=1)
        # strong augmentation (:326-337)
        if np.random.uniform(0, 1) < 0.5 and tr["unsupervised"].get("apply_aug", False):
            d, t, l = port.generate_unsup_data(image_u.numpy(), label_u_aug.numpy(), logits_u_aug.numpy(),
                                               mode=tr["unsupervised"]["apply_aug"])
            image_u_aug, label_u_aug = torch.from_numpy(d), torch.from_numpy(t)
        else:
            image_u_aug = image_u
        # S (:339-350)
        nl = len(image_l)
        S.training = True
        image_all = torch.cat((image_l, image_u_aug))
        outs = S.forward(image_all)
        pred_all, rep_all = outs["pred"], outs["rep"]
        pred_l_large, pred_u_large = _up(pred_all[:nl], (h, w)), _up(pred_all[nl:], (h, w))
        # supervised loss (:352-358)
        sup = _sup_criterion(cfg, pred_l_large, label_l)
        if self.aux:
            sup = sup + cfg["net"]["aux_loss"]["loss_weight"] * _sup_criterion(cfg, _up(outs["aux"][:nl], (h, w)), label_l)
        # T2 (:360-374)
        T.training = True
        with torch.no_grad():
            out_t = T.forward(image_all)
            pred_all_t, rep_all_t = out_t["pred"], out_t["rep"]
            prob_all_t = F.softmax(pred_all_t, dim=1)
            pred_u_large_t = _up(pred_all_t[nl:], (h, w))
        # unsupervised loss (:376-388, loss_helper.py:30-48)
        drop = tr["unsupervised"].get("drop_percent", 100)
        drop = 100 - (100 - drop) * (1 - epoch / tr["epochs"])
        target = label_u_aug.numpy().copy()
        u = port.compute_unsupervised_loss(pred_u_large.detach().numpy(), target, drop, pred_u_large_t.numpy())
        weight = target.size / max(u["n_kept"], 1)
        unsup = weight * F.cross_entropy(pred_u_large, torch.from_numpy(target), ignore_index=255) \
            * tr["unsupervised"].get("loss_weight", 1)
        # contrastive loss (:390-519, loss_helper.py:51-235)
        cc = tr.get("contrastive", False)
        contra_val = 0.0
        surrogate = rep_all.sum() * 0
        if cc:
            alpha_t = cc["low_entropy_threshold"] * (1 - epoch / tr["epochs"])
            prep = port.contra_prep(pred_u_large_t.numpy(), label_l.numpy(), label_u_aug.numpy(), alpha_t, self.C,
                                    tuple(pred_all.shape[2:]), cc.get("negative_high_entropy", True))
            out = port.compute_contra_memobank_loss(
                rep_all.detach().numpy(), prep["label_l_small"].astype(np.int64), prep["label_u_small"].astype(np.int64),
                prob_all_t[:nl].numpy(), prob_all_t[nl:].numpy(), prep["low_mask_all"], prep["high_mask_all"], cc,
                self.memobank, self.queue_ptr, self.queue_size, rep_all_t.numpy(), want_grad=True)
            contra_val = float(out["loss"]) * cc.get("loss_weight", 1)
            surrogate = (rep_all * torch.from_numpy(out["rep_grad"] * cc.get("loss_weight", 1))).sum()
        # backward + SGD (:524-528); `surrogate` has exactly the contrastive loss's gradient w.r.t. rep_all
        self.opt.zero_grad()
        (sup + unsup + surrogate).backward()
        self.opt.step()
        # EMA of the parameters (:531-548)
        with torch.no_grad():
            d = min(1 - 1 / (i_iter - len_loader * sup_only + 1), cfg["net"]["ema_decay"])
            for k, v in self.teacher.s.items():
                if v.requires_grad:
                    v.mul_(d).add_(self.student.s[k].detach(), alpha=1 - d)
        return float(sup.detach()), float(unsup.detach()), contra_val
