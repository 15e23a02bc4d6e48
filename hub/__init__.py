"""`from hub import ClipCompressor` keeps working exactly as in the reference (hub/__init__.py)."""
from lossyless_b200.hub import ClipCompressor

__all__ = ["ClipCompressor"]
