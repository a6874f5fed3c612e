"""oracle/eager_step.py -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT.

The reference step the way the reference executes it on a GPU box: plain torch ops on `device` for the
network and the per-pixel math, **numpy on the host for the three percentiles**, **the memory banks as CPU
tensors** that are copied to the device once per valid class, and one Python iteration per class
(train_semi.py:272-561; loss_helper.py:30-48, 51-235; utils.py:16-59).  It is the "reference torch-eager on
the same B200" comparator of SURVEY.md section 8(d)(i) (`bench.py --impl eager`): /root/reference cannot
travel to the GPU box, so what is timed there is this restatement, which keeps the reference's device
placement, host synchronisations and launch structure (it does not tidy any of them up).

It differs from oracle/step_port.py only in *where* things run (step_port does the loss math in numpy and
the C arithmetic contract so that it can be the bit-exact checker); on `device="cpu"` the two agree to
fp32 rounding, which tests/test_eager_step_cpu.py asserts.  Nothing under u2pl_b200/ is imported.

Single-process only: with world size 1 the reference's `dist.barrier()` + `all_gather_object` of the keys
(utils.py:17-25) degenerate to a pickle round trip, which is left out here (in the reference's favour).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import port
from .model_port import Net

IGNORE = 255


def _up(t, size):
    return F.interpolate(t, size, mode="bilinear", align_corners=True)


def _entropy(logits):
    p = torch.softmax(logits, dim=1)
    return -torch.sum(p * torch.log(p + 1e-10), dim=1)


def _host_percentile(values, q):
    return np.percentile(values.detach().cpu().numpy().flatten(), q)             # D2H + host partition (blocking)


def unsup_loss(predict, target, percent, pred_teacher):
    """loss_helper.py:30-48; mutates `target`."""
    B, _, h, w = predict.shape
    with torch.no_grad():
        ent = _entropy(pred_teacher)
        thresh = _host_percentile(ent[target != IGNORE], percent)
        drop = ent.ge(thresh).bool() * (target != IGNORE).bool()
        target[drop] = IGNORE
        weight = B * h * w / torch.sum(target != IGNORE)
    return weight * F.cross_entropy(predict, target, ignore_index=IGNORE)


def label_onehot(labels, C):
    """utils.py:50-59, same five tensor ops (so quirk Q8 -- a [B,1,H,W] index scattered along dim 0 of a
    [C,B,H,W] buffer -- is reproduced by construction, not by emulation)."""
    B, H, W = labels.shape
    buf = torch.zeros((C, B, H, W), device=labels.device)
    idx = labels.clone()
    idx[labels == IGNORE] = 0
    buf.scatter_(0, idx.unsqueeze(1), 1.0)
    buf[:, labels == IGNORE] = 0
    return buf.permute(1, 0, 2, 3)


def contra_prep(pred_u_large_t, label_l, label_u_aug, alpha_t, C, out_hw, negative_high_entropy=True):
    """train_semi.py:401-465."""
    with torch.no_grad():
        ent = _entropy(pred_u_large_t)
        valid = label_u_aug != IGNORE
        low_mask = ent.le(_host_percentile(ent[valid], alpha_t)).float() * valid.bool()
        high_mask = ent.ge(_host_percentile(ent[valid], 100 - alpha_t)).float() * valid.bool()
        labelled = (label_l.unsqueeze(1) != IGNORE).float()
        low_all = F.interpolate(torch.cat((labelled, low_mask.unsqueeze(1))), size=out_hw, mode="nearest")
        second = high_mask if negative_high_entropy else torch.ones_like(low_mask)
        high_all = F.interpolate(torch.cat((labelled, second.unsqueeze(1))), size=out_hw, mode="nearest")
        l_small = F.interpolate(label_onehot(label_l, C), size=out_hw, mode="nearest")
        u_small = F.interpolate(label_onehot(label_u_aug, C), size=out_hw, mode="nearest")
    return low_all, high_all, l_small.long(), u_small.long()


def enqueue(keys, bank, ptr, capacity):
    """utils.py:27-47 at world size 1: device keys -> host -> device -> host (the reference's own round trip),
    host concat, keep the newest `capacity` rows."""
    dev = keys.device
    keys = torch.cat([keys.detach().clone().cpu()], dim=0).to(dev)
    n = keys.shape[0]
    bank[0] = torch.cat((bank[0], keys.cpu()), dim=0)
    if bank[0].shape[0] >= capacity:
        bank[0] = bank[0][-capacity:, :]
        ptr[0] = capacity
    else:
        ptr[0] = (int(ptr[0]) + n) % capacity
    return n


def contra_loss(rep, label_l, label_u, prob_l, prob_u, low_mask, high_mask, cc, banks, ptrs, capacities, rep_teacher):
    """loss_helper.py:51-235 (momentum_prototype=None).  Returns (new_keys, loss)."""
    dev = rep.device
    p_thr, n_thr = cc["current_class_threshold"], cc["current_class_negative_threshold"]
    lo_rank, hi_rank = cc["low_rank"], cc["high_rank"]
    nq, nneg, temp = cc["num_queries"], cc["num_negatives"], cc["temperature"]
    D, n_lab, C = rep.shape[1], label_l.shape[0], label_l.shape[1]

    labels = torch.cat((label_l, label_u), dim=0)
    low_valid, high_valid = labels * low_mask, labels * high_mask
    rep = rep.permute(0, 2, 3, 1)
    rep_teacher = rep_teacher.permute(0, 2, 3, 1)
    order_l = torch.sort(prob_l, 1, True)[1].permute(0, 2, 3, 1)
    order_u = torch.sort(prob_u, 1, True)[1].permute(0, 2, 3, 1)
    prob = torch.cat((prob_l, prob_u), dim=0)

    anchors, counts, protos, present, new_keys = [], [], [], [], []
    unused = []                                                       # Q4: gathered with grad, never read
    for c in range(C):
        low_c, high_c = low_valid[:, c], high_valid[:, c]
        p_c = prob[:, c, :, :]
        anchor_mask = (p_c > p_thr) * low_c.bool()
        neg_mask = (p_c < n_thr) * high_c.bool()
        unused.append(rep[low_c.bool()])
        anchors.append(rep[anchor_mask])
        protos.append(torch.mean(rep_teacher[low_c.bool()].detach(), dim=0, keepdim=True))
        in_window_u = torch.sum(order_u[:, :, :, lo_rank:hi_rank].eq(c), dim=3).bool()
        in_top_l = torch.sum(order_l[:, :, :, :lo_rank].eq(c), dim=3).bool()
        class_mask = torch.cat((in_top_l * (label_l[:, c] == 0), in_window_u), dim=0)
        keys = rep_teacher[neg_mask * class_mask].detach()
        new_keys.append(enqueue(keys, banks[c], ptrs[c], capacities[c]))
        if low_c.sum() > 0:                                           # host sync
            counts.append(int(low_c.sum().item()))                    # host sync
            present.append(c)

    if len(counts) <= 1:
        return new_keys, torch.tensor(0.0, device=dev) * rep.sum()
    loss = torch.tensor(0.0, device=dev)
    proto = torch.cat(protos)
    n_present = len(counts)
    for j in range(n_present):
        bank = banks[present[j]][0]                                   # Q1: position j vs class present[j]
        if len(anchors[j]) == 0 or bank.shape[0] == 0:
            loss = loss + 0 * rep.sum()
            continue
        pick = torch.randint(len(anchors[j]), size=(nq,))             # CPU generator
        anchor = anchors[j][pick].clone().to(dev)
        with torch.no_grad():
            neg = bank.clone().to(dev)                                # H2D of the whole bank
            neg_pick = torch.randint(len(neg), size=(nq * nneg,))
            neg = neg[neg_pick].reshape(nq, nneg, D)
            pos = proto[j].unsqueeze(0).unsqueeze(0).repeat(nq, 1, 1).to(dev)
            both = torch.cat((pos, neg), dim=1)
        logits = torch.cosine_similarity(anchor.unsqueeze(1), both, dim=2)
        loss = loss + F.cross_entropy(logits / temp, torch.zeros(nq, device=dev).long())
    return new_keys, loss / n_present



def _sup_criterion(cfg, pred, target, ignore=IGNORE):
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

class EagerStep:
    """Same constructor and `step` signature as oracle/step_port.ReferenceStep; tensors of `*_state` decide the
    device of the network, the banks always live on the host."""

    def __init__(self, student_state, teacher_state, cfg, arch="resnet101", lr=0.001, momentum=0.9,
                 weight_decay=1e-4, head_lr_mult=10, bank_dim=256):
        net = cfg["net"]
        C = net["num_classes"]
        self.cfg, self.C, self.aux = cfg, C, bool(net.get("aux_loss", False))
        self.student = Net(student_state, arch, C, self.aux)
        self.teacher = Net(teacher_state, arch, C, self.aux)
        self.device = next(iter(student_state.values())).device
        enc = [v for k, v in student_state.items() if v.requires_grad and k.startswith("encoder.")]
        head = [v for k, v in student_state.items() if v.requires_grad and not k.startswith("encoder.")]
        self.opt = torch.optim.SGD([dict(params=enc, lr=lr), dict(params=head, lr=lr * head_lr_mult)],
                                   lr=lr, momentum=momentum, weight_decay=weight_decay)      # train_semi.py:97-112
        self.memobank = [[torch.zeros(0, bank_dim)] for _ in range(C)]                        # :161-169 (CPU)
        self.queue_ptr = [torch.zeros(1).long() for _ in range(C)]
        self.queue_size = [30000] * C
        self.queue_size[0] = 50000

    def step(self, image_l, label_l, image_u, epoch, i_iter, len_loader):
        cfg, tr = self.cfg, self.cfg["trainer"]
        dev = self.device
        image_l, label_l, image_u = image_l.to(dev), label_l.to(dev), image_u.to(dev)         # :283,287
        h, w = label_l.shape[1:]
        S, T = self.student, self.teacher
        sup_only = tr.get("sup_only_epoch", 1)
        T.training = False                                                                    # T1 :317-324
        with torch.no_grad():
            p = _up(T.forward(image_u)["pred"], (h, w))
            logits_u_aug, label_u_aug = torch.max(F.softmax(p, dim=1), dim=1)
        if np.random.uniform(0, 1) < 0.5 and tr["unsupervised"].get("apply_aug", False):      # :326-337
            image_u_aug, label_u_aug, logits_u_aug = self._strong_aug(image_u, label_u_aug, logits_u_aug,
                                                                      tr["unsupervised"]["apply_aug"])
        else:
            image_u_aug = image_u
        nl = len(image_l)                                                                     # S :339-350
        S.training = True
        image_all = torch.cat((image_l, image_u_aug))
        outs = S.forward(image_all)
        pred_all, rep_all = outs["pred"], outs["rep"]
        pred_l_large, pred_u_large = _up(pred_all[:nl], (h, w)), _up(pred_all[nl:], (h, w))
        sup = _sup_criterion(cfg, pred_l_large, label_l)                                      # :352-358
        if self.aux:
            sup = sup + cfg["net"]["aux_loss"]["loss_weight"] * _sup_criterion(cfg, _up(outs["aux"][:nl], (h, w)), label_l)
        T.training = True                                                                     # T2 :360-374
        with torch.no_grad():
            out_t = T.forward(image_all)
            pred_all_t, rep_all_t = out_t["pred"], out_t["rep"]
            prob_all_t = F.softmax(pred_all_t, dim=1)
            prob_l_t, prob_u_t = prob_all_t[:nl], prob_all_t[nl:]
            pred_u_large_t = _up(pred_all_t[nl:], (h, w))
        drop = tr["unsupervised"].get("drop_percent", 100)                                    # :376-388
        drop = 100 - (100 - drop) * (1 - epoch / tr["epochs"])
        unsup = unsup_loss(pred_u_large, label_u_aug.clone(), drop, pred_u_large_t.detach()) \
            * tr["unsupervised"].get("loss_weight", 1)
        cc = tr.get("contrastive", False)                                                     # :390-519
        contra = 0 * rep_all.sum()
        if cc:
            alpha_t = cc["low_entropy_threshold"] * (1 - epoch / tr["epochs"])
            low_all, high_all, l_small, u_small = contra_prep(pred_u_large_t, label_l, label_u_aug, alpha_t, self.C,
                                                              tuple(pred_all.shape[2:]),
                                                              cc.get("negative_high_entropy", True))
            _, contra = contra_loss(rep_all, l_small, u_small, prob_l_t.detach(), prob_u_t.detach(), low_all, high_all,
                                    cc, self.memobank, self.queue_ptr, self.queue_size, rep_all_t.detach())
            contra = contra * cc.get("loss_weight", 1)
        self.opt.zero_grad()                                                                  # :524-528
        (sup + unsup + contra).backward()
        self.opt.step()
        with torch.no_grad():                                                                 # EMA :531-548
            d = min(1 - 1 / (i_iter - len_loader * sup_only + 1), cfg["net"]["ema_decay"])
            for k, v in self.teacher.s.items():
                if v.requires_grad:
                    v.data = d * v.data + (1 - d) * self.student.s[k].data
        return sup.item(), unsup.item(), float(contra.detach())                                        # host syncs :553-561

    @staticmethod
    def _strong_aug(image, label, conf, mode):
        """augmentation.py:498-541 through the port's numpy-RNG mask generator; tensors stay on their device."""
        B, _, H, W = image.shape
        if mode not in ("cutout", "cutmix"):
            raise NotImplementedError(mode)
        new_i, new_l, new_c = [], [], []
        for b in range(B):
            m = torch.from_numpy(port.generate_cutout_mask([H, W])).to(image.device)
            if mode == "cutout":
                label[b][(1 - m).bool()] = IGNORE
                new_i.append((image[b] * m).unsqueeze(0))
                new_l.append(label[b].unsqueeze(0))
                new_c.append((conf[b] * m).unsqueeze(0))
                continue
            o = (b + 1) % B
            new_i.append((image[b] * m + image[o] * (1 - m)).unsqueeze(0))
            new_l.append((label[b] * m + label[o] * (1 - m)).unsqueeze(0))
            new_c.append((conf[b] * m + conf[o] * (1 - m)).unsqueeze(0))
        return torch.cat(new_i), torch.cat(new_l).long(), torch.cat(new_c)
