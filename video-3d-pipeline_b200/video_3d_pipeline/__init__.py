"""video_3d_pipeline -- B200-native drop-in for the depth hot path of jabberjabberjabber/video-3d-pipeline.

The public names are the ones the reference package exports (its __init__.py:5-16) plus the
`IGEVStereoDepthExtractor` name that run_pipeline.py:12 imports and the reference never defines.
They resolve lazily (PEP 562): `import video_3d_pipeline` does not load torch, cv2 or libv3d.so until
one of them is touched, so tools that only need `shard.frame_ranges` or `synthetic` start fast.
"""
import importlib

__version__ = "0.1.0"

_EXPORTS = {
    "align": ("VideoAligner",),
    "depth": ("HybridStereoDepthExtractor", "IGEVStereoDepthExtractor"),
    "upscale": ("SimpleDepthUpscaler",),
    "utils": ("get_video_info", "extract_audio", "verify_video_compatibility"),
}
_HOME = {name: mod for mod, names in _EXPORTS.items() for name in names}
__all__ = sorted(_HOME)


def __getattr__(name):
    mod = _HOME.get(name)
    if mod is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    value = getattr(importlib.import_module(f".{mod}", __name__), name)
    globals()[name] = value
    return value


def __dir__():
    return sorted(set(globals()) | set(__all__))
