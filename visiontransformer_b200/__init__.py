"""visiontransformer_b200 — B200 (sm_100a) native implementation of the ViT-segmentation hot path of
mtumalan/VisionTransformer (model/CE and model/PAED training and inference).

    from visiontransformer_b200.ce.classes import ViTSegmentationModel, LightningViTModel
    from visiontransformer_b200.paed.classes import PAEDTrainer, paed_loss_multiclass_soft

Same constructor signatures, attribute tree and state_dict keys as the reference classes; every stage underneath is a
hand-written CUDA kernel in libvitseg.so reached through the C ABI in include/vitseg.h.  There is no CPU fallback."""
from .model import ViTSegmentationModel, flops_per_image  # noqa: F401

__version__ = "0.1.0"
