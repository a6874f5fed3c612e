"""Drop-in for the one augmentation entry point the semi driver imports
(u2pl/dataset/augmentation.py:471-541): CutOut / CutMix / ClassMix on device tensors.
Rectangles come from numpy's global RNG in the reference's order so a seeded run cuts the
same boxes."""
import numpy as np
import torch


def generate_cutout_mask(img_size, ratio=2):
    area = img_size[0] * img_size[1] / ratio
    w = np.random.randint(img_size[1] / ratio + 1, img_size[1])
    h = np.round(area / w)
    x0 = np.random.randint(0, img_size[1] - w + 1)
    y0 = np.random.randint(0, img_size[0] - h + 1)
    return int(y0), int(y0 + h), int(x0), int(x0 + w)


def generate_class_mask(pseudo_labels):
    labels = torch.unique(pseudo_labels)
    chosen = labels[torch.randperm(len(labels))][: len(labels) // 2]
    return (pseudo_labels.unsqueeze(-1) == chosen).any(-1).float()


def generate_unsup_data(data, target, logits, mode="cutout"):
    B, _, H, W = data.shape
    keep = torch.ones((B, H, W), dtype=data.dtype, device=data.device)      # 1 = keep own pixel
    for i in range(B):
        if mode == "classmix":
            keep[i] = generate_class_mask(target[i]).to(data.device)
        else:
            y0, y1, x0, x1 = generate_cutout_mask([H, W], ratio=2)
            keep[i, y0:y1, x0:x1] = 0
    if mode == "cutout":
        target = target.clone()
        target[keep == 0] = 255
        return data * keep[:, None], target.long(), logits * keep
    nxt = torch.roll(torch.arange(B, device=data.device), -1)               # partner (i + 1) % B
    k_long = keep.long()
    new_data = data * keep[:, None] + data[nxt] * (1 - keep[:, None])
    new_target = target * k_long + target[nxt] * (1 - k_long)
    new_logits = logits * keep + logits[nxt] * (1 - keep)
    return new_data, new_target.long(), new_logits
