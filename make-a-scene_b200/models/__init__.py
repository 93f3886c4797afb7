"""Drop-in for the reference's `models` package (models/__init__.py:1): `_target_: models.VQBASE` resolves here
when `make-a-scene_b200/` is on sys.path in place of the reference root."""
from .vqvae import VQBASE  # noqa: F401
