"""Drop-in for the reference's model/CE/classes.py: `from classes import ViTSegmentationModel, LightningViTModel`.

LightningViTModel mirrors model/CE/classes.py:264-297 (ctor, .model, .loss_fn, _resize_target, training_step,
validation_step, configure_optimizers).  The training/validation steps take the fused path: low-resolution logits ->
fused bilinear-upsample + cross-entropy kernel, numerically the same loss as
nn.CrossEntropyLoss()(model(x), y) without materialising [B,C,224,224]."""
import torch
import torch.nn.functional as F
from torch import nn

from .._lightning import LightningModule
from ..losses import upsample_cross_entropy
from ..model import ViTSegmentationModel

__all__ = ["ViTSegmentationModel", "LightningViTModel"]


class LightningViTModel(LightningModule):
    def __init__(self, num_classes, patch_size, hidden_size, num_hidden_layers, num_attention_heads, **kwargs):
        super().__init__()
        self.model = ViTSegmentationModel(num_classes, patch_size, hidden_size, num_hidden_layers,
                                          num_attention_heads, **kwargs)
        self.loss_fn = nn.CrossEntropyLoss()  # kept for API parity (model/CE/classes.py:268); see _loss()

    def forward(self, x):
        return self.model(x)

    def _resize_target(self, y, size):
        # model/CE/classes.py:273-274 (legacy 'nearest' on the int64 label map; index glue, not arithmetic)
        return F.interpolate(y.unsqueeze(1).float(), size=size, mode='nearest').squeeze(1).long()

    def _loss(self, x, y):
        # training_step of the reference: y = self._resize_target(y, size=(224, 224)); loss_fn(self(x), y).  Here the
        # nearest-neighbour index map of _resize_target is applied inside the loss kernel's label read
        S = x.shape[-1]
        low = self.model.forward_lowres(x)
        return upsample_cross_entropy(low, y, S)

    def training_step(self, batch, batch_idx):
        x, y = batch
        loss = self._loss(x, y)
        self.log("train_loss", loss, prog_bar=True, on_epoch=True, logger=True)
        return loss

    def validation_step(self, batch, batch_idx):
        x, y = batch
        loss = self._loss(x, y)
        self.log("valid_loss", loss, prog_bar=True, on_epoch=True, logger=True)
        return loss

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=1e-5)
