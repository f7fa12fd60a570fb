"""chrono_photo_b200 -- B200 (sm_100a) implementation of chrono-photo's frame-stack compositing path
(`--mode outlier | darker | lighter`) behind the reference's own processor interface. See DESIGN.md.

(The directory is named chrono_photo_b200 because `chrono-photo_b200` is not an importable Python package name.)
"""
from .options import (BackgroundMode, Fade, FadeMode, FrameRange, OutlierSelectionMode, ParseEnumError, ParseOptionError,
                      SelectionMode, ShakeAnchor, ShakeParams, Threshold)
from .processors import (Context, FrameStack, OutlierProcessor, ShakeAnalyzer, SimpleProcessor, crop_create, decode_jpeg, encode_jpeg, fetch_last, fetch_last_device,
                         sample_positions, set_tuning, synth_frame_host, video_windows)

__all__ = ["BackgroundMode", "Fade", "FadeMode", "FrameRange", "OutlierSelectionMode", "ParseEnumError", "ParseOptionError",
           "SelectionMode", "ShakeAnchor", "ShakeParams", "Threshold", "Context", "ShakeAnalyzer", "FrameStack", "OutlierProcessor", "SimpleProcessor", "crop_create",
           "decode_jpeg", "encode_jpeg", "fetch_last", "fetch_last_device", "set_tuning", "sample_positions", "synth_frame_host", "video_windows"]
