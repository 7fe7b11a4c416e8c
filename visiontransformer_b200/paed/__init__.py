from .classes import LightningViTModel, PAEDTrainer, ViTSegmentationModel, paed_loss_multiclass_soft  # noqa: F401
from .segmentation import compute_sdf, dice_score, intersection_over_union, pixel_accuracy  # noqa: F401
