"""Compat mirrors of the reference's training/ surface files that do not import on Python >= 3.7."""
from .batch_processor import batch_processor  # noqa: F401
