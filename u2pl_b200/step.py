"""One semi-supervised U2PL training step on the GPU -- the fused counterpart of the reference
driver's loop body (train_semi.py:272-561).  Same order of operations and same RNG consumption;
what changes is where the work runs:

  reference (per step)                                   here
  -----------------------------------------------------  ------------------------------------------
  softmax/entropy of the teacher logits computed TWICE,   ONE u2pl_entropy_thresholds call resolves the
  three D2H copies + np.percentile on the host            three percentiles on device (drop_percent,
  (loss_helper.py:35-40, train_semi.py:402-415)           alpha_t, 100-alpha_t)
  two [B,C,H,W] fp32 one-hot temporaries + 4 nearest       u2pl_contra_prep_lowres -> class bitmasks +
  interpolations (train_semi.py:420-465)                  low/high masks at 1/4 resolution
  2*C python iterations, >= 2*C host syncs, C pickled      one count read-back, banks resident on device
  all-gathers, whole-bank H2D copies (loss_helper.py)     (contra.py)
  360 x 3 tiny kernels for the EMA (train_semi.py:543)    torch._foreach (multi-tensor) update
  3 x (all_reduce + .item()) for logging (:551-561)       one all_reduce of a 3-vector, no host sync

The network runs channels-last under bf16 autocast (fp32 master weights, fp32 BN statistics) when
`amp=True`; every loss is computed in fp32 from the fp32-cast logits, as in the reference.
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import contra, ops


def _world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class SemiStep:
    def __init__(self, model, model_teacher, optimizer, sup_loss_fn, cfg, memobank, queue_ptrlis, queue_size,
                 generate_unsup_data=None, amp=True, channels_last=True):
        self.model, self.teacher, self.optimizer, self.sup_loss_fn = model, model_teacher, optimizer, sup_loss_fn
        self.cfg = cfg
        self.memobank, self.queue_ptrlis, self.queue_size = memobank, queue_ptrlis, queue_size
        self.amp, self.channels_last = amp, channels_last
        if generate_unsup_data is None:
            from .u2pl.dataset.augmentation import generate_unsup_data
        self.generate_unsup_data = generate_unsup_data
        self.last = {}
        # x4 bilinear up-sampling fused into its consumers (csrc/upsample_ce.cu) when the class count is specialised;
        # U2PL_FUSED_UP=0 restores the ATen interpolate + full-resolution kernels (A/B runs, parity tests)
        import os
        self.fused_up = (os.environ.get("U2PL_FUSED_UP", "1") == "1"
                         and ops.upsample_fused_supported(cfg["net"]["num_classes"]))
        # SGD + EMA as one multi-tensor kernel (csrc/sgd_ema.cu); U2PL_FUSED_OPT=0 restores optimizer.step() + _foreach EMA
        self.fused_opt = None
        if (os.environ.get("U2PL_FUSED_OPT", "1") == "1" and isinstance(optimizer, torch.optim.SGD)
                and all(g.get("momentum", 0) > 0 and not g.get("nesterov", False) for g in optimizer.param_groups)):
            from .optim import FusedSGDEMA
            self.fused_opt = FusedSGDEMA(optimizer, list(model.parameters()), list(model_teacher.parameters()))

    # ------------------------------------------------------------------ helpers
    def _net(self, net, x):
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
            return net(x)

    PHASES = ("t1", "aug", "student_fwd", "sup_loss", "t2", "losses", "backward", "optim_ema")

    def _phase(self, start, name):
        """Close phase `name` (if it is being timed) and open the next one."""
        self._event(start, name)
        return self._event()

    def _event(self, start=None, name=None):
        """CUDA-event bracket on the current stream (only when a `timers` dict is installed, e.g. by bench.py)."""
        timers = getattr(self, "timers", None)
        if not timers or (name is not None and name not in timers):
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if start is not None:
            timers[name].append((start, ev))
        return ev

    @staticmethod
    def _up(t, size):
        # low-res logits leave the network channels-last; convert the small tensor (22 MB) to NCHW before
        # the x4 up-sampling so the 354 MB full-resolution tensors reach the loss kernels in the layout the
        # reference API defines (no hidden full-res transposes later)
        return F.interpolate(t.float().contiguous(), size, mode="bilinear", align_corners=True)

    # ------------------------------------------------------------------ the step
    def __call__(self, image_l, label_l, image_u, epoch, i_iter, len_loader):
        cfg = self.cfg
        trainer = cfg["trainer"]
        model, teacher = self.model, self.teacher
        sup_only_epoch = trainer.get("sup_only_epoch", 1)
        h, w = label_l.shape[1:]
        has_aux = "aux_loss" in cfg["net"].keys()
        model.train()

        if epoch < sup_only_epoch:                                            # train_semi.py:288-307
            outs = self._net(model, image_l)
            pred = self._up(outs["pred"], (h, w))
            sup_loss = self.sup_loss_fn([pred, self._up(outs["aux"], (h, w))] if has_aux else pred, label_l)
            teacher.train()
            with torch.no_grad():
                self._net(teacher, image_l)
            unsup_loss = contra_loss = outs["rep"].float().sum() * 0
        else:
            if epoch == sup_only_epoch:                                       # :309-315 (aliases, like the reference)
                with torch.no_grad():
                    for t_p, s_p in zip(teacher.parameters(), model.parameters()):
                        t_p.data = s_p.data
            # ---- T1: pseudo labels from the eval-mode teacher (:317-324)
            ph = self._event()
            teacher.eval()
            with torch.no_grad():
                dec = getattr(getattr(teacher, "module", teacher), "decoder", None)
                if dec is not None:
                    dec.skip_rep = True                                       # only "pred" is read here
                try:
                    pred_t1 = self._net(teacher, image_u)["pred"]
                finally:
                    if dec is not None:
                        dec.skip_rep = False
                if self.fused_up:                                             # bilinear + softmax + max in one kernel
                    logits_u_aug, label_u_aug = ops.up_softmax_max(pred_t1, (h, w))
                else:
                    pred_u_teacher = self._up(pred_t1, (h, w))
                    logits_u_aug, label_u_aug = torch.max(F.softmax(pred_u_teacher, dim=1), dim=1)
            ph = self._phase(ph, "t1")
            # ---- strong augmentation (:326-337)
            if np.random.uniform(0, 1) < 0.5 and trainer["unsupervised"].get("apply_aug", False):
                image_u_aug, label_u_aug, logits_u_aug = self.generate_unsup_data(
                    image_u, label_u_aug.clone(), logits_u_aug.clone(), mode=trainer["unsupervised"]["apply_aug"])
            else:
                image_u_aug = image_u
            ph = self._phase(ph, "aug")
            # ---- S: student forward on labelled + augmented unlabelled (:339-350)
            num_labeled = len(image_l)
            outs = self._net(model, torch.cat((image_l, image_u_aug)))
            pred_all, rep_all = outs["pred"], outs["rep"]
            sup_lowres = self.fused_up and hasattr(self.sup_loss_fn, "forward_lowres")
            if not sup_lowres:
                pred_l_large = self._up(pred_all[:num_labeled], (h, w))
            if not self.fused_up:
                pred_u_large = self._up(pred_all[num_labeled:], (h, w))
            ph = self._phase(ph, "student_fwd")
            # ---- supervised loss (:352-358)
            if sup_lowres:                                                    # CE (+aux): up-sampling fused into the loss
                sup_loss = self.sup_loss_fn.forward_lowres(
                    [pred_all[:num_labeled], outs["aux"][:num_labeled]] if has_aux else pred_all[:num_labeled], label_l)
            elif has_aux:
                aux = self._up(outs["aux"][:num_labeled], (h, w))
                sup_loss = self.sup_loss_fn([pred_l_large, aux], label_l.clone())
            else:
                sup_loss = self.sup_loss_fn(pred_l_large, label_l.clone())
            ph = self._phase(ph, "sup_loss")
            # ---- T2: train-mode teacher forward, no grad (:360-374)
            teacher.train()
            with torch.no_grad():
                out_t = self._net(teacher, torch.cat((image_l, image_u_aug)))
                pred_all_teacher, rep_all_teacher = out_t["pred"].float().contiguous(), out_t["rep"]
                prob_all_teacher = F.softmax(pred_all_teacher, dim=1)
                pred_u_large_teacher = self._up(pred_all_teacher[num_labeled:], (h, w))
            ph = self._phase(ph, "t2")
            # ---- unsupervised + contrastive losses, one entropy pass (:376-519)
            drop_percent = trainer["unsupervised"].get("drop_percent", 100)
            drop_percent = 100 - (100 - drop_percent) * (1 - epoch / trainer["epochs"])
            cfg_contra = trainer.get("contrastive", False)
            percents = [drop_percent]
            if cfg_contra:
                alpha_t = cfg_contra["low_entropy_threshold"] * (1 - epoch / trainer["epochs"])
                percents += [alpha_t, 100 - alpha_t]
            label_u_aug = label_u_aug.contiguous()
            ev = self._event()                                                 # one launch: loss_helper.py:35-44 + :402-415
            ent, thresh, _, target, n_kept, _ = ops.entropy_partition(pred_u_large_teacher, label_u_aug, percents, 0)
            self._event(ev, "entropy_partition")
            if self.fused_up:
                unsup_loss = ops.upsampled_unsup_ce(pred_all[num_labeled:], target, n_kept)
            else:
                unsup_loss = ops.unsup_ce(pred_u_large, target, n_kept)
            unsup_loss = unsup_loss * trainer["unsupervised"].get("loss_weight", 1)
            if cfg_contra:
                if cfg_contra.get("binary", False):
                    raise NotImplementedError("compute_binary_memobank_loss is undefined in the reference itself "
                                              "(train_semi.py:469)")
                label_bits, low_mask_all, high_mask_all = ops.contra_prep_lowres(
                    label_l, label_u_aug, ent, thresh, 1, 2, tuple(pred_all.shape[2:]), cfg["net"]["num_classes"],
                    cfg_contra.get("negative_high_entropy", True))
                new_keys, contra_loss = contra.compute_contra_memobank_loss(
                    rep_all, None, None, prob_all_teacher[:num_labeled], prob_all_teacher[num_labeled:],
                    low_mask_all, high_mask_all, cfg_contra, self.memobank, self.queue_ptrlis, self.queue_size,
                    rep_all_teacher, label_bits=label_bits)
                world = _world_size()
                if world > 1:                                                  # :514 all_reduce(value), local gradient
                    total = contra_loss.detach().clone()
                    dist.all_reduce(total)
                    contra_loss = contra_loss + (total - contra_loss.detach())
                contra_loss = contra_loss / world * cfg_contra.get("loss_weight", 1)
                self.last["new_keys"] = new_keys
            else:
                contra_loss = rep_all.float().sum() * 0

        ph = self._phase(ph, "losses") if epoch >= sup_only_epoch else self._event()
        loss = sup_loss + unsup_loss + contra_loss                            # :524-528
        self.optimizer.zero_grad()
        loss.backward()
        ph = self._phase(ph, "backward")
        if epoch > sup_only_epoch and self.fused_opt is not None:
            # SGD step and EMA (train_semi.py:531-548) in one multi-tensor kernel
            ema_decay = min(1 - 1 / (i_iter - len_loader * sup_only_epoch + 1), cfg["net"]["ema_decay"])
            self.fused_opt.step(ema_decay)
        else:
            self.optimizer.step()
            if epoch >= sup_only_epoch:                                       # :531-548 EMA (parameters only)
                with torch.no_grad():
                    ema_decay = min(1 - 1 / (i_iter - len_loader * sup_only_epoch + 1), cfg["net"]["ema_decay"])
                    t_params = [p.data for p in teacher.parameters()]
                    s_params = [p.data for p in model.parameters()]
                    if epoch == sup_only_epoch:
                        # the reference re-binds t.data to a fresh tensor (:546-548); keep that un-aliasing
                        new = torch._foreach_mul(t_params, ema_decay)
                        torch._foreach_add_(new, s_params, alpha=1 - ema_decay)
                        for p, n in zip(teacher.parameters(), new):
                            p.data = n
                    else:
                        torch._foreach_mul_(t_params, ema_decay)
                        torch._foreach_add_(t_params, s_params, alpha=1 - ema_decay)

        self._phase(ph, "optim_ema")
        losses = torch.stack([sup_loss.detach().float(), unsup_loss.detach().float(), contra_loss.detach().float()])
        if _world_size() > 1:                                                 # :551-561, one collective, no .item()
            dist.all_reduce(losses)
        self.last["losses"] = losses
        return losses
