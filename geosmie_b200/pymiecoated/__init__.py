"""Drop-in mirror of the reference package `pymiecoated` (src/pymiecoated/pymiecoated/__init__.py:1) on sm_100a CUDA."""
from .mie_coated import Mie, MultipleMie  # noqa: F401
