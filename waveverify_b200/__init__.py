"""waveverify_b200 - B200-native (sm_100a) drop-in for WaveVerify's embed / detect / locate path.

Same names and signatures as the reference's `model.Generator / Detector / Locator /
AudioWatermarking` and `waveverify.WaveVerify`; all compute runs in hand-written CUDA behind the
C ABI in include/wv_b200.h.  No CPU fallback.
"""
from .audio import AudioSignal
from .models import (AudioWatermarking, Detector, Generator, Locator, ber_miou, metric_counters)
from .api import (WaveVerify, embed_streaming, load_audio, message_to_tensor, save_audio, tensor_to_message)
from .params import NetConfig, config_from_kwargs, fixture_state_dict, param_spec
from .watermark_id import WatermarkID

__all__ = [
    "AudioSignal", "AudioWatermarking", "Detector", "Generator", "Locator", "NetConfig",
    "WaveVerify", "WatermarkID", "embed_streaming", "load_audio", "message_to_tensor", "save_audio",
    "tensor_to_message", "ber_miou", "config_from_kwargs", "fixture_state_dict", "metric_counters", "param_spec",
]
