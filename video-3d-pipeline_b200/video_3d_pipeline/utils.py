"""Shared utilities the depth / upscale steps call (mirror of the reference's utils.py surface).

Only `get_video_info` (utils.py:17-38) and `create_work_directory` (utils.py:292-296) sit on the
hot path's call stack (depth.py:17,34-35,147,418; upscale.py:9,87).  The audio helpers belong to
the alignment step, which is outside the accelerated path; they are kept as thin, dependency-gated
functions so that `from video_3d_pipeline import ...` (reference __init__.py:5-8) keeps working.
"""
import shutil
import subprocess
from fractions import Fraction
from pathlib import Path
from typing import Dict, Optional

import cv2


def _probe_with_ffprobe(video_path: str) -> Optional[Dict]:
    """utils.py:17-38 asks ffprobe (through ffmpeg-python); same fields, no eval()."""
    exe = shutil.which("ffprobe")
    if exe is None:
        return None
    import json
    try:
        out = subprocess.run([exe, "-v", "error", "-print_format", "json", "-show_streams", str(video_path)],
                             capture_output=True, check=True, text=True).stdout
        streams = json.loads(out).get("streams", [])
    except Exception:
        return None
    vs = next((s for s in streams if s.get("codec_type") == "video"), None)
    if not vs:
        return None
    try:
        return {
            "width": int(vs["width"]),
            "height": int(vs["height"]),
            "fps": float(Fraction(vs["r_frame_rate"])),
            "duration": float(vs["duration"]),
            "frames": int(vs.get("nb_frames", 0)),
        }
    except Exception:
        return None


def _probe_with_opencv(video_path: str) -> Optional[Dict]:
    cap = cv2.VideoCapture(str(video_path))
    try:
        if not cap.isOpened():
            return None
        fps = float(cap.get(cv2.CAP_PROP_FPS)) or 0.0
        frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        info = {
            "width": int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)),
            "height": int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)),
            "fps": fps,
            "duration": (frames / fps) if fps > 0 else 0.0,
            "frames": max(frames, 0),
        }
        return info if info["width"] > 0 and info["height"] > 0 else None
    finally:
        cap.release()


def get_video_info(video_path: str) -> Optional[Dict]:
    """{'width','height','fps','duration','frames'} or None (utils.py:17-38).

    ffprobe when the binary exists (what the reference uses), else cv2.VideoCapture.
    """
    try:
        info = _probe_with_ffprobe(video_path) or _probe_with_opencv(video_path)
    except Exception as e:   # the reference prints and returns None (utils.py:36-38)
        print(f"Error getting video info: {e}")
        return None
    if info is None:
        print(f"Error getting video info: could not probe {video_path}")
    return info


def create_work_directory(base_path: str = "temp_pipeline") -> Path:
    """utils.py:292-296."""
    work_dir = Path(base_path)
    work_dir.mkdir(exist_ok=True)
    return work_dir


def verify_video_compatibility(video1_path: str, video2_path: str) -> Dict:
    """Coarse duration / fps comparison of two videos (utils.py:228-259 surface)."""
    a, b = get_video_info(video1_path), get_video_info(video2_path)
    if not a or not b:
        return {"compatible": False, "reason": "could not read video info", "video1": a, "video2": b}
    dur = abs(a["duration"] - b["duration"])
    fps = abs(a["fps"] - b["fps"])
    return {
        "compatible": dur < 60.0 and fps < 0.1,
        "duration_diff": dur,
        "fps_diff": fps,
        "video1": a,
        "video2": b,
    }


def extract_audio(video_path: str, work_dir: Path, duration_seconds: float = 600, sample_rate: int = 22050) -> str:
    """Mono WAV extraction for the alignment step, cached in work_dir (utils.py:41-119: same positional
    signature, cache file name and ValueError sites).  Needs the ffmpeg binary."""
    import hashlib
    import os
    exe = shutil.which("ffmpeg")
    if exe is None:
        raise RuntimeError("extract_audio needs the ffmpeg binary, which is not installed; "
                           "audio alignment is outside the accelerated depth path")
    if not get_video_info(video_path):
        raise ValueError(f"Could not read video info for {video_path}")       # utils.py:47-49
    video_hash = hashlib.md5(f"{video_path}_{duration_seconds}_{sample_rate}".encode()).hexdigest()[:16]
    wav = Path(work_dir) / f"audio_cache_{video_hash}.wav"                      # utils.py:63-64
    if wav.exists() and os.path.getmtime(wav) > os.path.getmtime(video_path):   # utils.py:67-72
        print(f"Using cached audio: {wav}")
        return str(wav)
    cmd = [exe, "-y", "-v", "error", "-t", str(duration_seconds), "-i", str(video_path), "-vn", "-acodec", "pcm_s16le",
           "-ac", "1", "-ar", str(sample_rate), str(wav)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 or not wav.exists():
        raise ValueError(f"Could not extract audio from {video_path}")         # utils.py:105, 108
    if wav.stat().st_size < 1000:
        raise ValueError("Audio extraction produced unusually small file")     # utils.py:112-113
    return str(wav)


def apply_alignment_offset(alignment_file: str, target_video_path: str, base_start_time: float = 0) -> float:
    """Start time in `target_video_path` after the stored audio offset (utils.py:299-327): video1 (the SBS
    clip) is the time reference, video2 (the 4K clip) is shifted by time_offset_seconds, never below 0."""
    import json
    with open(alignment_file) as f:
        data = json.load(f)
    offset = float(data["time_offset_seconds"])
    if str(target_video_path) == data.get("video1_path"):
        start = float(base_start_time)
    elif str(target_video_path) == data.get("video2_path"):
        start = float(base_start_time) + offset
    else:
        raise ValueError(f"Video {target_video_path} not found in alignment data")
    if start < 0:
        print(f"Warning: Adjusted start time {start:.3f}s < 0, using 0")
        start = 0.0
    return start


def guide_start_frame(alignment_file: str, video_4k_path: str, fps: float) -> int:
    """First 4K frame that belongs to depth map 0: the alignment offset in frames of the 4K clip."""
    return int(round(apply_alignment_offset(alignment_file, video_4k_path, 0.0) * float(fps)))
