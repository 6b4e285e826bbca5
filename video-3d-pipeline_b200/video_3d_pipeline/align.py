"""Audio-based temporal alignment (reference: align.py) -- OUTSIDE the accelerated path.

One cross-correlation per title is seconds of CPU work and not per-frame, so nothing here touches
the GPU.  The class is provided so that run_pipeline.py:11,41-43 imports and runs unmodified; it
needs the ffmpeg binary for audio extraction and raises a clear error when that is missing.
"""
import json
import wave
from typing import Dict

import numpy as np

from .utils import create_work_directory, extract_audio


def _read_wav_mono(path: str):
    with wave.open(path, "rb") as w:
        rate, n, width = w.getframerate(), w.getnframes(), w.getsampwidth()
        raw = w.readframes(n)
    dt = {1: np.uint8, 2: np.int16, 4: np.int32}[width]
    a = np.frombuffer(raw, dtype=dt).astype(np.float64)
    if width == 1:
        a -= 128.0
    return a, rate


class VideoAligner:
    """find_alignment / assess_alignment_quality surface of align.py:13-116."""

    def __init__(self, video1_path: str, video2_path: str, work_dir: str = "temp_alignment"):
        self.video1_path = video1_path          # align.py:16-19: video 1 = the SBS clip, the time reference
        self.video2_path = video2_path
        self.work_dir = create_work_directory(work_dir)

    def find_alignment(self, max_audio_length: float = 300) -> Dict:
        from scipy import signal
        a_path = extract_audio(self.video1_path, self.work_dir, max_audio_length)      # align.py:41-42
        b_path = extract_audio(self.video2_path, self.work_dir, max_audio_length)
        a, rate = _read_wav_mono(a_path)
        b, rate_b = _read_wav_mono(b_path)
        if rate != rate_b:
            raise ValueError("audio sample rates differ")
        a = (a - a.mean()) / (a.std() + 1e-12)
        b = (b - b.mean()) / (b.std() + 1e-12)
        corr = signal.correlate(b, a, mode="full", method="fft")
        lag = int(np.argmax(corr)) - (len(a) - 1)
        peak = float(corr.max() / max(min(len(a), len(b)), 1))
        from .utils import get_video_info
        i1, i2 = get_video_info(self.video1_path) or {}, get_video_info(self.video2_path) or {}
        fps1 = float(i1.get("fps") or 0.0)
        offset = lag / float(rate)
        # same keys as the reference's alignment_data.json (align.py:65-76)
        data = {"video1_path": str(self.video1_path), "video2_path": str(self.video2_path),
                "time_offset_seconds": float(offset), "offset_frames": float(offset * fps1) if fps1 else 0.0,
                "correlation_strength": peak, "frame_duration": (1.0 / fps1) if fps1 else 0.0,
                "video1_fps": fps1, "video2_fps": float(i2.get("fps") or 0.0), "sample_rate": int(rate),
                "audio_length_analyzed": float(max_audio_length)}
        with open(self.work_dir / "alignment_data.json", "w") as f:
            json.dump(data, f, indent=2)
        return data

    def assess_alignment_quality(self, alignment_data: Dict, tolerance_frames: float = 2.0) -> str:
        """align.py:87-116: same thresholds and labels."""
        offset = float(alignment_data["time_offset_seconds"])
        c = float(alignment_data.get("correlation_strength", 0.0))
        if abs(offset) < float(alignment_data.get("frame_duration", 0.0)) * tolerance_frames:
            return "EXCELLENT"
        if c > 0.8:
            return "GOOD"
        if c > 0.6:
            return "MODERATE"
        return "POOR"
