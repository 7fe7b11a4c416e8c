from .classes import LightningViTModel, ViTSegmentationModel  # noqa: F401
