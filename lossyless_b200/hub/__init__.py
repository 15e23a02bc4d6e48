from .compressor import ClipCompressor

__all__ = ["ClipCompressor"]
