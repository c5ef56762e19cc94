"""Stub of the `lightning` surface the reference's models import (base class + Trainer name only)."""
from types import SimpleNamespace

__version__ = "0.0-stub"

from torch import nn


class _HParams(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


class LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self._hparams = _HParams()

    def save_hyperparameters(self, hp=None, **_):
        if hp:
            self._hparams.update(dict(hp))

    @property
    def hparams(self):
        return self._hparams


class Trainer:  # pragma: no cover - never used for arithmetic
    def __init__(self, *a, **k):
        pass

    def predict(self, module, dataloader):
        import torch

        module.eval()
        with torch.inference_mode():
            return [module.predict_step(b) for b in dataloader]


pytorch = SimpleNamespace()
