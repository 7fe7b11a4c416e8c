"""LightningModule base when `lightning` is importable, else a plain nn.Module with the logging no-ops the
reference wrappers call (self.log / self.log_dict) — Lightning itself is not part of the hot path."""
from torch import nn

try:  # pragma: no cover - depends on the environment
    import lightning as L  # type: ignore

    LightningModule = L.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):  # type: ignore
        def log(self, *args, **kwargs):
            return None

        def log_dict(self, *args, **kwargs):
            return None
