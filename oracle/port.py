"""oracle/port.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

CPU restatement (numpy + torch-CPU fp32 for the floating point tail) of the U2PL
per-step loss path, function by function, each citing the reference lines it
follows.  It is the checker for the CUDA path in u2pl_b200/; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  The product never does.

Pinned against the real reference: oracle/make_golden.py imports /root/reference
(in the build container only), runs the reference functions on seeded inputs and
commits inputs + outputs under tests/golden/; tests/test_oracle_golden.py replays
them through this file.  The ordering-sensitive arithmetic (softmax entropy,
percentile, ranks) lives in oracle/u2pl_oracle.c under the arithmetic contract of
DESIGN.md section 3.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libu2pl_oracle.so")
_lib = None


def build():
    """gcc-compile oracle/u2pl_oracle.c (test infrastructure)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def _c():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        f32p, i64 = ctypes.POINTER(ctypes.c_float), ctypes.c_int64
        L.u2pl_oracle_expf.restype = ctypes.c_float
        L.u2pl_oracle_expf.argtypes = [ctypes.c_float]
        L.u2pl_oracle_logf.restype = ctypes.c_float
        L.u2pl_oracle_logf.argtypes = [ctypes.c_float]
        L.u2pl_oracle_entropy.restype = None
        L.u2pl_oracle_entropy.argtypes = [f32p, i64, i64, i64, f32p]
        L.u2pl_oracle_softmax.restype = None
        L.u2pl_oracle_softmax.argtypes = [f32p, i64, i64, i64, f32p]
        L.u2pl_oracle_percentile.restype = ctypes.c_int
        L.u2pl_oracle_percentile.argtypes = [f32p, i64, ctypes.c_float, f32p,
                                             ctypes.POINTER(i64), ctypes.POINTER(i64)]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


# ----------------------------------------------------------------------------- entropy / percentile
def entropy(logits):
    """-sum(softmax * log(softmax + 1e-10), dim=1) -- loss_helper.py:35-36, train_semi.py:402-403."""
    x = np.ascontiguousarray(logits, dtype=np.float32)
    B, C = x.shape[0], x.shape[1]
    HW = int(np.prod(x.shape[2:]))
    out = np.empty((B,) + tuple(x.shape[2:]), dtype=np.float32)
    _c().u2pl_oracle_entropy(_fp(x), B, C, HW, _fp(out))
    return out


def softmax(logits):
    x = np.ascontiguousarray(logits, dtype=np.float32)
    B, C = x.shape[0], x.shape[1]
    HW = int(np.prod(x.shape[2:]))
    out = np.empty_like(x)
    _c().u2pl_oracle_softmax(_fp(x), B, C, HW, _fp(out))
    return out


def percentile(vals, q):
    """np.percentile(float32 data, q) restated (numpy 2.x float32 index arithmetic, two-sided lerp).
    Call sites: loss_helper.py:38-40, train_semi.py:405-407,412-415."""
    v = np.ascontiguousarray(vals, dtype=np.float32).ravel()
    if v.size == 0:
        return np.float32(np.nan)
    out = ctypes.c_float()
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    rc = _c().u2pl_oracle_percentile(_fp(v), v.size, np.float32(q), ctypes.byref(out), ctypes.byref(lo), ctypes.byref(hi))
    assert rc == 0
    return np.float32(out.value)


# ----------------------------------------------------------------------------- A6
def _log_softmax64(x):
    x = x.astype(np.float64)
    m = x.max(axis=1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=1, keepdims=True))


def cross_entropy_sum(predict, target, ignore=255):
    """sum over non-ignored pixels of -log_softmax(predict)[target], and their count (float64 reference)."""
    ls = _log_softmax64(np.asarray(predict))
    t = np.asarray(target)
    valid = t != ignore
    tt = np.where(valid, t, 0)
    nll = -np.take_along_axis(ls, tt[:, None], axis=1)[:, 0]
    return float((nll * valid).sum()), int(valid.sum())


def compute_unsupervised_loss(predict, target, percent, pred_teacher, ignore=255):
    """loss_helper.py:30-48.  Mutates `target` in place like the reference.
    Returns dict(loss, entropy, thresh, drop_mask, n_kept)."""
    predict = np.asarray(predict, dtype=np.float32)
    B, C, H, W = predict.shape
    ent = entropy(pred_teacher)                                      # :35-36
    thresh = percentile(ent[target != ignore], percent)              # :38-40
    drop = (ent >= thresh) & (target != ignore)                      # :41
    target[drop] = ignore                                            # :43
    n_kept = int((target != ignore).sum())
    nll_sum, n_used = cross_entropy_sum(predict, target, ignore)
    assert n_used == n_kept
    weight = np.float32(B * H * W) / np.float32(n_kept) if n_kept else np.float32(np.nan)   # :44
    loss = np.float32(weight) * np.float32(nll_sum / n_kept if n_kept else np.nan)          # :46
    return dict(loss=np.float32(loss), entropy=ent, thresh=thresh, drop_mask=drop, n_kept=n_kept)


def unsup_grad(predict, target_after, ignore=255, upstream=1.0):
    """d loss / d predict for compute_unsupervised_loss given the rewritten target."""
    predict = np.asarray(predict, dtype=np.float32)
    B, C, H, W = predict.shape
    valid = target_after != ignore
    n = int(valid.sum())
    p = np.exp(_log_softmax64(predict))
    oh = np.zeros_like(p)
    tt = np.where(valid, target_after, 0)
    np.put_along_axis(oh, tt[:, None], 1.0, axis=1)
    g = (p - oh) * valid[:, None] * (upstream * (B * H * W) / n / n)
    return g.astype(np.float32)


# ----------------------------------------------------------------------------- A8
def nearest_src_index(dst, src):
    """F.interpolate(mode='nearest') source index: min(floor(d * float32(src/dst)), src-1)."""
    scale = np.float32(src) / np.float32(dst)
    idx = np.floor(np.arange(dst, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, src - 1)


def label_onehot(inputs, num_segments, ignore=255):
    """utils.py:50-59 -> [B, C, H, W] float32, INCLUDING the reference's scatter quirk (Q8):
    `outputs.scatter_(0, inputs_temp.unsqueeze(1), 1.0)` on a [C, B, H, W] buffer with a
    [B, 1, H, W] index writes outputs[label[b, y, x], 0, y, x] for EVERY b, so batch slot 0
    receives the union (multi-hot) of all images' classes -- ignored pixels of any image count as
    class 0 (:54-55) -- and slots b > 0 stay all-zero; `outputs[:, inputs == 255] = 0` (:57) then
    clears slot 0 only where image 0 itself is ignored."""
    inputs = np.asarray(inputs)
    B, H, W = inputs.shape
    out = np.zeros((B, num_segments, H, W), dtype=np.float32)
    t = np.where(inputs == ignore, 0, inputs)                       # :54-55
    for b in range(B):                                              # :56 (dim-1 index is always 0)
        np.put_along_axis(out[0], t[b][None], 1.0, axis=0)
    out[0] *= (inputs[0] != ignore)[None]                           # :57
    return out


def contra_prep(pred_u_large_teacher, label_l, label_u_aug, alpha_t, num_classes, out_hw,
                negative_high_entropy=True, ignore=255):
    """train_semi.py:401-465: entropy, two percentiles, low/high masks, nearest down-sample,
    one-hot labels.  Returns dict(low_mask_all, high_mask_all, label_l_small, label_u_small,
    low_thresh, high_thresh, entropy)."""
    ent = entropy(pred_u_large_teacher)                                                # :402-403
    valid_u = label_u_aug != ignore
    low_thresh = percentile(ent[valid_u], alpha_t)                                     # :405-407
    low_entropy_mask = (ent <= low_thresh).astype(np.float32) * valid_u                # :408-410
    high_thresh = percentile(ent[valid_u], 100 - alpha_t)                              # :412-415
    high_entropy_mask = (ent >= high_thresh).astype(np.float32) * valid_u              # :416-418
    lab_valid = (label_l != ignore).astype(np.float32)
    h, w = out_hw
    H, W = label_l.shape[1:]
    sy, sx = nearest_src_index(h, H), nearest_src_index(w, W)

    def down(a):                                                                       # :427-429 nearest
        return a[..., sy[:, None], sx[None, :]]

    low_mask_all = down(np.concatenate([lab_valid, low_entropy_mask])[:, None])        # :420-429
    if negative_high_entropy:
        high_mask_all = down(np.concatenate([lab_valid, high_entropy_mask])[:, None])  # :432-439,451-453
    else:
        high_mask_all = down(np.concatenate([lab_valid, np.ones_like(high_entropy_mask)])[:, None])
    label_l_small = down(label_onehot(label_l, num_classes, ignore))                   # :456-460
    label_u_small = down(label_onehot(label_u_aug, num_classes, ignore))               # :461-465
    return dict(low_mask_all=low_mask_all.astype(np.float32), high_mask_all=high_mask_all.astype(np.float32),
                label_l_small=label_l_small, label_u_small=label_u_small,
                low_thresh=low_thresh, high_thresh=high_thresh, entropy=ent)


# ----------------------------------------------------------------------------- A10
def dequeue_and_enqueue(keys_per_rank, queue, queue_ptr, queue_size):
    """utils.py:28-47.  keys_per_rank: list (rank order) of [k_r, D] arrays -- the result of
    gather_together (:16-24).  queue: 1-element list holding the [n, D] bank; queue_ptr: 1-element list."""
    keys = np.concatenate(keys_per_rank, axis=0) if len(keys_per_rank) else np.zeros((0, queue[0].shape[1]), np.float32)
    batch = keys.shape[0]
    ptr = int(queue_ptr[0])
    queue[0] = np.concatenate([queue[0], keys.astype(np.float32)], axis=0)     # :38
    if queue[0].shape[0] >= queue_size:                                        # :39-41
        queue[0] = queue[0][-queue_size:, :]
        ptr = queue_size
    else:
        ptr = (ptr + batch) % queue_size                                       # :43
    queue_ptr[0] = ptr
    return batch


# ----------------------------------------------------------------------------- A9
def rank_desc_stable(prob):
    """rank[b, c, y, x] of class c in a stable descending sort over the class axis
    (= position of c in torch.sort(prob, 1, True) indices, loss_helper.py:91-97)."""
    p = np.asarray(prob)
    C = p.shape[1]
    gt = (p[:, None, :] > p[:, :, None])                               # [B, c, j, ...] : p_j > p_c
    eq = (p[:, None, :] == p[:, :, None])
    jlt = (np.arange(C)[None, :] < np.arange(C)[:, None])              # j < c
    jlt = jlt.reshape((1, C, C) + (1,) * (p.ndim - 2))
    return (gt | (eq & jlt)).sum(axis=2)


def contra_select(label_l, label_u, prob_l, prob_u, low_mask, high_mask, cfg):
    """Loop 1 of compute_contra_memobank_loss (loss_helper.py:80-154) without touching features:
    per class the row-major flat pixel indices (over [2B, h, w]) of anchors, low-valid pixels and
    negative keys, plus valid_classes."""
    thr, nthr = cfg["current_class_threshold"], cfg["current_class_negative_threshold"]
    low_rank, high_rank = cfg["low_rank"], cfg["high_rank"]
    label = np.concatenate([label_l, label_u], axis=0)
    num_labeled = label_l.shape[0]
    C = label_l.shape[1]
    low_valid = label * low_mask                                        # :80
    high_valid = label * high_mask                                      # :81
    prob = np.concatenate([prob_l, prob_u], axis=0)                     # :99
    rank_l, rank_u = rank_desc_stable(prob_l), rank_desc_stable(prob_u)
    anchors, lowvalid, negs, valid_classes = [], [], [], []
    for i in range(C):
        lv = low_valid[:, i].astype(bool)
        hv = high_valid[:, i].astype(bool)
        m_low = (prob[:, i] > thr) & lv                                 # :108-110
        m_high = (prob[:, i] < nthr) & hv                               # :111-113
        cm_u = (rank_u[:, i] >= low_rank) & (rank_u[:, i] < high_rank)  # :127-129
        cm_l = (rank_l[:, i] < low_rank) & (label_l[:, i] == 0)         # :134,137
        cm = np.concatenate([cm_l, cm_u], axis=0)
        neg = m_high & cm                                               # :140
        anchors.append(np.flatnonzero(m_low))
        lowvalid.append(np.flatnonzero(lv))
        negs.append(np.flatnonzero(neg))
        if lv.sum() > 0:                                                # :152-154
            valid_classes.append(i)
    return dict(anchors=anchors, lowvalid=lowvalid, negs=negs, valid_classes=valid_classes,
                num_labeled=num_labeled)


def compute_contra_memobank_loss(rep, label_l, label_u, prob_l, prob_u, low_mask, high_mask, cfg,
                                 memobank, queue_prtlis, queue_size, rep_teacher,
                                 other_rank_keys=None, want_grad=False):
    """loss_helper.py:51-235 (momentum_prototype=None branch).

    rep / rep_teacher: [2B, D, h, w] float32 arrays.  memobank: list of 1-element lists of [n, D]
    arrays, mutated in place; queue_prtlis: list of 1-element lists.  other_rank_keys: optional
    callable(class_index, local_keys) -> list of per-rank key arrays in rank order (world > 1).
    Consumes the torch CPU default generator exactly like the reference (:179-181,194-196).
    Returns dict(new_keys, loss (np.float32), rep_grad (if want_grad), sel)."""
    temp, nq, nn_ = cfg["temperature"], cfg["num_queries"], cfg["num_negatives"]
    rep = np.asarray(rep, dtype=np.float32)
    rep_teacher = np.asarray(rep_teacher, dtype=np.float32)
    D = rep.shape[1]
    rep_rows = torch.from_numpy(np.ascontiguousarray(rep.transpose(0, 2, 3, 1)).reshape(-1, D)).clone()   # :83
    rep_rows.requires_grad_(want_grad)
    rept_rows = np.ascontiguousarray(rep_teacher.transpose(0, 2, 3, 1)).reshape(-1, D)                     # :84
    sel = contra_select(label_l, label_u, prob_l, prob_u, low_mask, high_mask, cfg)
    C = label_l.shape[1]
    new_keys, protos = [], []
    for i in range(C):
        lv = sel["lowvalid"][i]
        with np.errstate(invalid="ignore", divide="ignore"):
            protos.append(rept_rows[lv].mean(axis=0, dtype=np.float32, keepdims=True) if lv.size
                          else np.full((1, D), np.nan, np.float32))                                         # :119-123
        keys = rept_rows[sel["negs"][i]]                                                                    # :142
        per_rank = other_rank_keys(i, keys) if other_rank_keys is not None else [keys]
        new_keys.append(dequeue_and_enqueue(per_rank, memobank[i], queue_prtlis[i], queue_size[i]))         # :143-150
    valid_classes = sel["valid_classes"]
    valid_seg = len(valid_classes)
    out = dict(new_keys=new_keys, sel=sel, valid_classes=valid_classes, sampled=[])
    if valid_seg <= 1:                                                                                      # :156-162
        out["loss"] = np.float32(0.0)
        out["rep_grad"] = np.zeros_like(rep) if want_grad else None
        return out
    loss = torch.zeros(())
    for j in range(valid_seg):                                                                              # :173
        anc = sel["anchors"][j]                                # Q1: list position j, not valid_classes[j]
        bank = memobank[valid_classes[j]][0]
        if not (len(anc) > 0 and bank.shape[0] > 0):                                                        # :174-188
            continue
        a_idx = torch.randint(len(anc), size=(nq,))                                                         # :179-181
        anchor = rep_rows[torch.from_numpy(anc)[a_idx]]                                                     # :182-184
        n_idx = torch.randint(bank.shape[0], size=(nq * nn_,))                                              # :194-196
        neg = torch.from_numpy(bank)[n_idx].reshape(nq, nn_, D)                                             # :197-200
        pos = torch.from_numpy(protos[j]).reshape(1, 1, D).repeat(nq, 1, 1)                                 # :201-207
        allf = torch.cat((pos, neg), dim=1)                                                                 # :220-222
        logits = torch.cosine_similarity(anchor.unsqueeze(1), allf, dim=2)                                  # :224-226
        loss = loss + torch.nn.functional.cross_entropy(logits / temp, torch.zeros(nq, dtype=torch.long))   # :228-230
        out["sampled"].append(dict(j=j, cls=valid_classes[j], a_idx=a_idx.numpy(), n_idx=n_idx.numpy()))
    loss = loss / valid_seg                                                                                 # :233
    if want_grad and loss.requires_grad:
        loss.backward()
        B2, _, h, w = rep.shape
        out["rep_grad"] = rep_rows.grad.numpy().reshape(B2, h, w, D).transpose(0, 3, 1, 2).copy()
    elif want_grad:
        out["rep_grad"] = np.zeros_like(rep)
    out["loss"] = np.float32(loss.detach().numpy())
    return out


# ----------------------------------------------------------------------------- A12
def criterion_ce(pred, target, ignore=255):
    """Criterion.forward without aux (loss_helper.py:316-320): mean CE over non-ignored pixels."""
    s, n = cross_entropy_sum(pred, target, ignore)
    return np.float32(s / n) if n else np.float32(np.nan)


def ohem_ce(pred, target, thresh=0.7, min_kept=100000, ignore=255):
    """OhemCrossEntropy2dTensor.forward (loss_helper.py:502-531).  Returns (loss, kept_mask [B,H,W])."""
    pred = np.asarray(pred, dtype=np.float32)
    b, c, h, w = pred.shape
    t = np.asarray(target).reshape(-1).copy()
    valid = t != ignore                                                         # :505
    t = t * valid                                                               # :506
    num_valid = int(valid.sum())
    prob = softmax(pred).transpose(1, 0, 2, 3).reshape(c, -1)                   # :509-510
    if min_kept > num_valid:                                                    # :512
        pass
    elif num_valid > 0:
        prob = np.where(valid[None], prob, np.float32(1))                       # :516
        mask_prob = prob[t, np.arange(t.size)]                                  # :517
        threshold = np.float32(thresh)
        if min_kept > 0:
            srt = np.sort(mask_prob, kind="stable")                             # :520
            kth = srt[min(srt.size, min_kept) - 1]                              # :521
            if kth > thresh:                                                    # :522-523
                threshold = kth
            kept = mask_prob <= threshold                                       # :524
            t = t * kept
            valid = valid & kept                                                # :526
    t = np.where(valid, t, ignore).reshape(b, h, w)                             # :528-529
    return criterion_ce(pred, t, ignore), valid.reshape(b, h, w)


# ----------------------------------------------------------------------------- A13 / A14
def ema_update(teacher, student, ema_decay_origin, i_iter, len_loader, sup_only_epoch):
    """train_semi.py:531-548 over lists of arrays (parameters only)."""
    d = min(1 - 1 / (i_iter - len_loader * sup_only_epoch + 1), ema_decay_origin)
    return [np.float32(d) * t + np.float32(1 - d) * s for t, s in zip(teacher, student)], d


def generate_cutout_mask(img_size, ratio=2):
    """augmentation.py:471-485 (consumes np.random exactly like the reference)."""
    cutout_area = img_size[0] * img_size[1] / ratio
    w = np.random.randint(img_size[1] / ratio + 1, img_size[1])
    h = np.round(cutout_area / w)
    x_start = np.random.randint(0, img_size[1] - w + 1)
    y_start = np.random.randint(0, img_size[0] - h + 1)
    x_end, y_end = int(x_start + w), int(y_start + h)
    mask = np.ones(img_size, dtype=np.int64)
    mask[int(y_start):y_end, x_start:x_end] = 0
    return mask


def generate_unsup_data(data, target, logits, mode="cutout"):
    """augmentation.py:498-541 for cutout / cutmix (classmix draws torch.randperm; not restated)."""
    B = data.shape[0]
    nd, nt, nl = [], [], []
    target = target.copy()
    for i in range(B):
        m = generate_cutout_mask(list(data.shape[2:]), ratio=2)
        if mode == "cutout":
            target[i][(1 - m).astype(bool)] = 255
            nd.append(data[i] * m)
            nt.append(target[i])
            nl.append(logits[i] * m)
            continue
        assert mode == "cutmix"
        j = (i + 1) % B
        nd.append(data[i] * m + data[j] * (1 - m))
        nt.append(target[i] * m + target[j] * (1 - m))
        nl.append(logits[i] * m + logits[j] * (1 - m))
    return np.stack(nd).astype(np.float32), np.stack(nt).astype(np.int64), np.stack(nl).astype(np.float32)
