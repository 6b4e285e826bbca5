"""video_3d_pipeline -- B200-native drop-in for the depth hot path of jabberjabberjabber/video-3d-pipeline.

Export list follows the reference's __init__.py:5-16 (including the `IGEVStereoDepthExtractor`
name that run_pipeline.py:12 imports and the reference never defines).
"""
__version__ = "0.1.0"

from .align import VideoAligner
from .depth import HybridStereoDepthExtractor, IGEVStereoDepthExtractor
from .upscale import SimpleDepthUpscaler
from .utils import get_video_info, extract_audio, verify_video_compatibility

__all__ = [
    "VideoAligner",
    "HybridStereoDepthExtractor",
    "IGEVStereoDepthExtractor",
    "SimpleDepthUpscaler",
    "get_video_info",
    "extract_audio",
    "verify_video_compatibility",
]
