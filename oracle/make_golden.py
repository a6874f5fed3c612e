"""oracle/make_golden.py -- regenerates tests/golden/*.npz by RUNNING THE REFERENCE.

Runs only in the build container (needs /root/reference, which never travels to the
GPU box).  It imports the unmodified reference package with three non-invasive shims
(SURVEY.md section 8c): a stub `skimage.measure`, `Tensor.cuda -> identity`, and a
1-rank gloo process group so utils.dequeue_and_enqueue's barrier/all_gather_object work.
Inline driver code (train_semi.py:401-465, no importable name) is executed from the
reference file by line range at run time -- nothing from the reference is copied here.

    python oracle/make_golden.py            # writes tests/golden/*.npz

Fixtures are small (few hundred KB each) and committed; the script is committed with them.
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch

REF = os.environ.get("U2PL_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def install_shims():
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.measure")
    skm.label = skm.regionprops = lambda *a, **k: None
    sk.measure = skm
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.measure", skm)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29731")
        dist.init_process_group("gloo", rank=0, world_size=1)
    sys.path.insert(0, REF)


def smooth_logits(rng, B, C, h, w, H, W, scale=3.0):
    """SURVEY 8d kernel-level inputs: low-res N(0, scale^2), box-smoothed, bilinear x(H/h) align_corners."""
    x = torch.from_numpy(rng.standard_normal((B, C, h, w)).astype(np.float32)) * scale
    x = torch.nn.functional.avg_pool2d(x, 5, stride=1, padding=2, count_include_pad=False)
    return torch.nn.functional.interpolate(x, (H, W), mode="bilinear", align_corners=True)


def blocky_labels(rng, B, H, W, C, block=8, border=0):
    hb, wb = (H + block - 1) // block, (W + block - 1) // block
    lab = rng.integers(0, C, size=(B, hb, wb))
    lab = np.repeat(np.repeat(lab, block, axis=1), block, axis=2)[:, :H, :W].astype(np.int64)
    if border:
        lab[:, :border] = 255
        lab[:, -border:] = 255
        lab[:, :, :border] = 255
        lab[:, :, -border:] = 255
    return lab


def gen_unsup(name, seed, B, C, H, W, percent, with_ignore):
    from u2pl.utils.loss_helper import compute_unsupervised_loss
    rng = np.random.default_rng(seed)
    pred_teacher = smooth_logits(rng, B, C, (H + 3) // 4, (W + 3) // 4, H, W)
    predict = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).requires_grad_(True)
    target = pred_teacher.argmax(1)
    if with_ignore:
        target[:, H // 4: H // 2, W // 3: W - 2] = 255
    target_in = target.clone()
    loss = compute_unsupervised_loss(predict, target, percent, pred_teacher)
    loss.backward()
    prob = torch.softmax(pred_teacher, dim=1)
    ent = -torch.sum(prob * torch.log(prob + 1e-10), dim=1)
    thresh = np.percentile(ent[target_in != 255].numpy().flatten(), percent)
    np.savez_compressed(os.path.join(OUT, name), predict=predict.detach().numpy(), pred_teacher=pred_teacher.numpy(),
                        target_in=target_in.numpy().astype(np.int16), target_out=target.numpy().astype(np.int16),
                        percent=np.float64(percent), loss=loss.detach().numpy(), grad=predict.grad.numpy(),
                        ref_entropy=ent.numpy(), ref_thresh=np.float32(thresh))
    print(name, "loss", float(loss), "thresh", float(thresh), "kept", int((target != 255).sum()))


def exec_driver_lines(first, last, ns):
    """Execute train_semi.py[first:last] (1-based, inclusive) from the reference file in namespace ns."""
    with open(os.path.join(REF, "train_semi.py")) as fh:
        lines = fh.readlines()[first - 1:last]
    exec(compile(textwrap.dedent("".join(lines)), f"train_semi.py:{first}-{last}", "exec"), ns)


def gen_prep(name, seed, B, C, H, W, h, w, epoch, epochs, cutout):
    from u2pl.utils.utils import label_onehot
    rng = np.random.default_rng(seed)
    pred_u_large_teacher = smooth_logits(rng, B, C, h, w, H, W)
    label_l = torch.from_numpy(blocky_labels(rng, B, H, W, C, border=3))
    label_u_aug = smooth_logits(rng, B, C, h, w, H, W).argmax(1)
    if cutout:
        label_u_aug[:, 5: H // 2, 7: W // 2] = 255
    cfg = {"trainer": {"epochs": epochs, "contrastive": {"low_entropy_threshold": 20, "negative_high_entropy": True,
                                                         "low_rank": 3, "high_rank": 20}},
           "net": {"num_classes": C}}
    ns = dict(torch=torch, np=np, F=torch.nn.functional, label_onehot=label_onehot, cfg=cfg,
              cfg_contra=cfg["trainer"]["contrastive"], epoch=epoch, pred_u_large_teacher=pred_u_large_teacher,
              label_l=label_l, label_u_aug=label_u_aug, logits_u_aug=torch.zeros(B, H, W),
              pred_all=torch.zeros(2 * B, C, h, w), contra_flag="")
    exec_driver_lines(397, 399, ns)          # alpha_t
    exec_driver_lines(401, 465, ns)          # the no_grad block
    np.savez_compressed(os.path.join(OUT, name), pred_u_large_teacher=pred_u_large_teacher.numpy(),
                        label_l=label_l.numpy().astype(np.int16), label_u_aug=label_u_aug.numpy().astype(np.int16),
                        alpha_t=np.float64(ns["alpha_t"]), h=h, w=w, C=C,
                        low_mask_all=ns["low_mask_all"].numpy().astype(np.uint8),
                        high_mask_all=ns["high_mask_all"].numpy().astype(np.uint8),
                        label_l_small=ns["label_l_small"].numpy().astype(np.uint8),
                        label_u_small=ns["label_u_small"].numpy().astype(np.uint8),
                        low_thresh=np.float32(ns["low_thresh"]), high_thresh=np.float32(ns["high_thresh"]),
                        ref_entropy=ns["entropy"].numpy())
    print(name, "alpha_t", ns["alpha_t"], "low", float(ns["low_thresh"]), "high", float(ns["high_thresh"]),
          "nlow", int(ns["low_mask_all"].sum()), "nhigh", int(ns["high_mask_all"].sum()))


def contra_inputs(rng, Bl, Bu, C, D, h, w, missing_class=None, driver_onehot=False):
    """Inputs with the statistics the loss needs: smooth teacher probabilities, labels that mostly
    agree with them but not always (so rank windows fire), random masks."""
    import torch.nn.functional as F
    logit_l = smooth_logits(rng, Bl, C, (h + 1) // 2, (w + 1) // 2, h, w, scale=7.0)
    logit_u = smooth_logits(rng, Bu, C, (h + 1) // 2, (w + 1) // 2, h, w, scale=7.0)
    prob_l, prob_u = torch.softmax(logit_l, 1), torch.softmax(logit_u, 1)
    lab_l = prob_l.argmax(1)
    flip = torch.from_numpy(rng.random((Bl, h, w)) < 0.3)
    lab_l = torch.where(flip, torch.from_numpy(rng.integers(0, C, (Bl, h, w))), lab_l)
    # pseudo labels: a mid-ranked class on 40% of the pixels (unreliable pixels), argmax elsewhere
    order = prob_u.argsort(1, descending=True)
    pick = torch.from_numpy(rng.integers(0, min(C, 8), (Bu, 1, h, w)))
    lab_u = torch.where(torch.from_numpy(rng.random((Bu, h, w)) < 0.4), order.gather(1, pick)[:, 0], prob_u.argmax(1))
    if missing_class is not None:
        lab_l[lab_l == missing_class] = (missing_class + 1) % C
        lab_u[lab_u == missing_class] = (missing_class + 1) % C
    lab_l[:, :2] = 255
    onehot = lambda lab: (F.one_hot(torch.where(lab == 255, 0, lab), C).permute(0, 3, 1, 2) * (lab != 255)[:, None]).long()
    if driver_onehot:        # what train_semi.py:456-465 really feeds the loss: the reference's own label_onehot (quirk Q8)
        from u2pl.utils.utils import label_onehot
        onehot = lambda lab: label_onehot(lab, C).long()
    low_mask = torch.cat(((lab_l != 255).float(), torch.from_numpy((rng.random((Bu, h, w)) < 0.6).astype(np.float32))))[:, None]
    high_mask = torch.cat(((lab_l != 255).float(), torch.from_numpy((rng.random((Bu, h, w)) < 0.5).astype(np.float32))))[:, None]
    rep = torch.from_numpy(rng.standard_normal((Bl + Bu, D, h, w)).astype(np.float32))
    rep_t = torch.from_numpy(rng.standard_normal((Bl + Bu, D, h, w)).astype(np.float32))
    return dict(rep=rep, rep_teacher=rep_t, label_l=onehot(lab_l), label_u=onehot(lab_u), prob_l=prob_l, prob_u=prob_u,
                low_mask=low_mask, high_mask=high_mask)


def gen_contra(name, seed, Bl, Bu, C, D, h, w, steps, qsize, missing_class=None, nq=16, nneg=5, driver_onehot=False):
    from u2pl.utils.loss_helper import compute_contra_memobank_loss
    rng = np.random.default_rng(seed)
    cfg = dict(negative_high_entropy=True, low_rank=3, high_rank=min(20, C), current_class_threshold=0.3,
               current_class_negative_threshold=1, num_negatives=nneg, num_queries=nq, temperature=0.5)
    memobank = [[torch.zeros(0, D)] for _ in range(C)]
    queue_ptrlis = [torch.zeros(1, dtype=torch.long) for _ in range(C)]
    queue_size = [qsize] * C
    queue_size[0] = qsize + 7
    torch.manual_seed(seed)
    save = dict(cfg_keys=np.array(list(cfg.keys())), cfg_vals=np.array([float(v) for v in cfg.values()]),
                queue_size=np.array(queue_size), steps=steps, seed=seed)
    for s in range(steps):
        inp = contra_inputs(rng, Bl, Bu, C, D, h, w, missing_class if s == steps - 1 else None, driver_onehot)
        rep = inp["rep"].clone().requires_grad_(True)
        new_keys, loss = compute_contra_memobank_loss(rep, inp["label_l"], inp["label_u"], inp["prob_l"], inp["prob_u"],
                                                      inp["low_mask"], inp["high_mask"], cfg, memobank, queue_ptrlis,
                                                      queue_size, inp["rep_teacher"])
        loss.backward()
        for k, v in inp.items():
            a = v.numpy()
            save[f"s{s}_{k}"] = a.astype(np.uint8) if k in ("label_l", "label_u", "low_mask", "high_mask") else a
        save[f"s{s}_new_keys"] = np.array(new_keys)
        save[f"s{s}_loss"] = loss.detach().numpy()
        save[f"s{s}_grad"] = rep.grad.numpy() if rep.grad is not None else np.zeros_like(rep.detach().numpy())
        save[f"s{s}_bank_len"] = np.array([m[0].shape[0] for m in memobank])
        save[f"s{s}_ptr"] = np.array([int(p[0]) for p in queue_ptrlis])
        print(name, "step", s, "loss", float(loss), "new_keys", new_keys[:8], "bank", save[f"s{s}_bank_len"][:8])
    for c in range(C):
        save[f"bank_{c}"] = memobank[c][0].numpy()
    np.savez_compressed(os.path.join(OUT, name), **save)


def gen_ohem(name, seed, B, C, H, W, min_kept, thresh):
    from u2pl.utils.loss_helper import OhemCrossEntropy2dTensor, Criterion
    rng = np.random.default_rng(seed)
    pred = (smooth_logits(rng, B, C, H // 4, W // 4, H, W) * 0.7).requires_grad_(True)
    target = torch.from_numpy(blocky_labels(rng, B, H, W, C, border=2))
    crit = OhemCrossEntropy2dTensor(255, thresh, min_kept)
    loss = crit(pred, target.clone())
    loss.backward()
    ce = Criterion(0, ignore_index=255)(pred.detach(), target)
    np.savez_compressed(os.path.join(OUT, name), pred=pred.detach().numpy(), target=target.numpy().astype(np.int16),
                        min_kept=min_kept, thresh=thresh, loss=loss.detach().numpy(), grad=pred.grad.numpy(),
                        ce_loss=ce.numpy())
    print(name, "ohem", float(loss), "ce", float(ce))


def gen_aug(name, seed, B, H, W, mode):
    from u2pl.dataset.augmentation import generate_unsup_data
    rng = np.random.default_rng(seed)
    data = torch.from_numpy(rng.standard_normal((B, 3, H, W)).astype(np.float32))
    target = torch.from_numpy(rng.integers(0, 21, (B, H, W)))
    logits = torch.from_numpy(rng.random((B, H, W)).astype(np.float32))
    np.random.seed(seed)
    torch.manual_seed(seed)                                   # classmix draws torch.randperm (augmentation.py:488-497)
    nd, nt, nl = generate_unsup_data(data, target.clone(), logits.clone(), mode=mode)
    np.savez_compressed(os.path.join(OUT, name), data=data.numpy(), target=target.numpy().astype(np.int16),
                        logits=logits.numpy(), seed=seed, mode=mode, new_data=nd.numpy(),
                        new_target=nt.numpy().astype(np.int16), new_logits=nl.numpy())
    print(name, mode, "changed", int((nt != target).sum()))


def gen_model(name, seed, arch, C, B, H, W, aux):
    """Seeded construction (ModelBuilder consumes the torch generator deterministically) + forward in
    train and eval mode.  Only input + outputs are stored; the checker rebuilds weights from the seed."""
    from u2pl.models.model_helper import ModelBuilder
    net = {"num_classes": C, "sync_bn": False, "ema_decay": 0.99,
           "encoder": {"type": f"u2pl.models.resnet.{arch}",
                       "kwargs": {"multi_grid": True, "zero_init_residual": True, "fpn": True,
                                  "replace_stride_with_dilation": [False, True, True], "pretrained": False}},
           "decoder": {"type": "u2pl.models.decoder.dec_deeplabv3_plus",
                       "kwargs": {"inner_planes": 256, "dilations": [12, 24, 36]}}}
    if aux:
        net["aux_loss"] = {"aux_plane": 1024, "loss_weight": 0.4}
    torch.manual_seed(seed)
    model = ModelBuilder(net)
    x = torch.randn(B, 3, H, W)
    names = [n for n, _ in model.named_parameters()]
    model.eval()
    with torch.no_grad():
        out_eval = model(x)
    psum = np.array([float(p.double().sum()) for p in model.parameters()])
    save = dict(x=x.numpy(), seed=seed, arch=arch, C=C, aux=aux, param_names=np.array(names), param_sums=psum,
                n_params=sum(p.numel() for p in model.parameters()))
    for k, v in out_eval.items():
        save[f"eval_{k}"] = v.numpy()
    model.train()
    torch.manual_seed(seed + 1)               # Dropout2d masks
    out_tr = model(x)
    for k, v in out_tr.items():
        save[f"train_{k}"] = v.detach().numpy()
    np.savez_compressed(os.path.join(OUT, name), **save)
    print(name, arch, "params", save["n_params"], {k: tuple(v.shape) for k, v in out_eval.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    install_shims()
    gen_unsup("unsup_c21", 11, 2, 21, 33, 33, 90.0, False)
    gen_unsup("unsup_c19_ignore", 12, 2, 19, 31, 37, 83.5, True)
    gen_unsup("unsup_c5_p100", 13, 1, 5, 24, 24, 100.0, False)
    gen_prep("prep_c21", 21, 2, 21, 33, 33, 9, 9, 40, 80, False)
    gen_prep("prep_c19_cutout", 22, 2, 19, 41, 41, 11, 11, 13, 200, True)
    gen_contra("contra_c21", 31, 2, 2, 21, 16, 13, 13, steps=3, qsize=6)
    gen_contra("contra_c19_missing", 32, 2, 2, 19, 16, 12, 14, steps=3, qsize=40, missing_class=2)
    gen_contra("contra_c21_driver_onehot", 33, 3, 3, 21, 16, 12, 12, steps=3, qsize=25, driver_onehot=True)
    gen_ohem("ohem_c19", 41, 2, 19, 32, 32, min_kept=300, thresh=0.7)
    gen_ohem("ohem_c19_kth", 42, 2, 19, 32, 32, min_kept=1500, thresh=0.05)
    gen_aug("aug_cutmix", 51, 4, 33, 33, "cutmix")
    gen_aug("aug_cutout", 52, 3, 29, 35, "cutout")
    gen_aug("aug_classmix", 53, 4, 31, 27, "classmix")
    gen_model("model_r50_c21", 61, "resnet50", 21, 2, 33, 33, False)
    gen_model("model_r50_c19_aux", 62, "resnet50", 19, 2, 41, 41, True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "aug_classmix":       # added in round 2: regenerate this fixture alone
        os.makedirs(OUT, exist_ok=True)
        install_shims()
        gen_aug("aug_classmix", 53, 4, 31, 27, "classmix")
    else:
        main()
