"""Audio front end for speech models (Whisper): decode / resample to 16 kHz mono, log-mel spectrogram.

Parity: reference ``multimodal/processors/whisper.py`` (wraps the HF ``WhisperFeatureExtractor``).  Own implementation (torch.stft + a Slaney
mel filterbank) so that serving needs neither librosa nor torchaudio; ``tests/test_whisper.py`` checks it against the HF extractor.
"""
from __future__ import annotations

import base64
import io
import os
import wave
from typing import Any, List, Tuple

import numpy as np
import torch

SAMPLE_RATE = 16000


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3)
    logstep = np.log(6.4) / 27.0
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / logstep, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = m * (200.0 / 3)
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), lin)


def mel_filterbank(n_mels: int = 80, n_fft: int = 400, sr: int = SAMPLE_RATE) -> np.ndarray:
    """[n_mels, n_fft/2 + 1] triangular filters on the Slaney mel scale, area-normalised (librosa.filters.mel defaults)"""
    fft_f = np.linspace(0, sr / 2, n_fft // 2 + 1)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    lower, upper = -ramps[:-2] / fdiff[:-1, None], ramps[2:] / fdiff[1:, None]
    w = np.maximum(0, np.minimum(lower, upper))
    return (w * (2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels]))[:, None]).astype(np.float32)


def log_mel_spectrogram(wave_f32: np.ndarray, n_mels: int = 80, n_fft: int = 400, hop: int = 160, chunk_s: int = 30) -> np.ndarray:
    """-> [n_mels, chunk_s * 100] float32 (zero-padded / truncated to the 30 s window Whisper was trained on)"""
    n = chunk_s * SAMPLE_RATE
    x = np.zeros(n, dtype=np.float32)
    m = min(n, wave_f32.shape[0])
    x[:m] = wave_f32[:m]
    t = torch.from_numpy(x)
    st = torch.stft(t, n_fft, hop, window=torch.hann_window(n_fft), center=True, pad_mode="reflect", return_complex=True)
    mag = st[..., :-1].abs().pow(2)
    mel = torch.from_numpy(mel_filterbank(n_mels, n_fft)) @ mag
    lg = torch.clamp(mel, min=1e-10).log10()
    lg = torch.maximum(lg, lg.max() - 8.0)
    return ((lg + 4.0) / 4.0).numpy().astype(np.float32)


def _resample(x: np.ndarray, sr: int) -> np.ndarray:
    if sr == SAMPLE_RATE:
        return x
    n = int(round(x.shape[0] * SAMPLE_RATE / sr))
    # band-limited enough for speech: linear interpolation after a short moving-average low-pass when decimating
    if sr > SAMPLE_RATE:
        k = max(1, int(sr // SAMPLE_RATE))
        x = np.convolve(x, np.ones(k, dtype=np.float32) / k, mode="same")
    return np.interp(np.linspace(0, x.shape[0] - 1, n), np.arange(x.shape[0]), x).astype(np.float32)


def load_audio(data: Any) -> np.ndarray:
    """waveform array (assumed 16 kHz) | (array, sample_rate) | WAV bytes | path | base64 / data-URL of a WAV  ->  float32 mono 16 kHz"""
    if isinstance(data, tuple) and len(data) == 2:
        return _resample(np.asarray(data[0], dtype=np.float32).reshape(-1), int(data[1]))
    if isinstance(data, (np.ndarray, list)):
        return np.asarray(data, dtype=np.float32).reshape(-1)
    if isinstance(data, torch.Tensor):
        return data.detach().cpu().float().numpy().reshape(-1)
    if isinstance(data, str):
        if data.startswith("data:"):
            data = base64.b64decode(data.split(",", 1)[1])
        elif os.path.exists(data):
            with open(data, "rb") as f:
                data = f.read()
        else:
            data = base64.b64decode(data)
    if not isinstance(data, (bytes, bytearray)):
        raise TypeError(f"unsupported audio input {type(data)}")
    with wave.open(io.BytesIO(bytes(data)), "rb") as w:
        sr, ch, sw, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if sw == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif sw == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif sw == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported WAV sample width {sw}")
    if ch > 1:
        x = x.reshape(-1, ch).mean(1)
    return _resample(x, sr)


def process_audio(audio_data: Any, cfg: Any) -> dict:
    a = cfg.audio
    feats = log_mel_spectrogram(load_audio(audio_data), a["num_mel_bins"], chunk_s=a.get("chunk_length", 30))
    return {"input_features": feats[:, : 2 * a["max_source_positions"]]}


def process_audio_llm(input_ids: List[int], audio_data: Any, cfg: Any) -> Tuple[List[int], dict]:
    """audio-understanding LLMs (Qwen2-Audio): every audio placeholder token expands to the encoder's output length for its clip; the clips'
    padded log-mel windows and true frame counts travel in the request (reference multimodal/processors/qwen_audio.py)"""
    import hashlib

    a = cfg.audio
    clips = audio_data if isinstance(audio_data, (list, tuple)) and not (len(audio_data) == 2 and isinstance(audio_data[1], (int, float))) else [audio_data]
    tok, V = a["audio_token_id"], cfg.vocab_size
    window = 2 * a["max_source_positions"]
    feats, frames, per_clip = [], [], []
    fixed = int(a.get("tokens_per_window") or 0)          # Voxtral: every window yields the same number of tokens, long audio = several windows
    for c in clips:
        w = load_audio(c)
        pieces = [w] if not fixed else [w[i:i + window * 160] for i in range(0, max(1, w.shape[0]), window * 160)]
        first = len(feats)
        for piece in pieces:
            f = log_mel_spectrogram(piece, a["num_mel_bins"], chunk_s=a.get("chunk_length", 30))[:, :window]
            if f.shape[1] < window:
                f = np.pad(f, ((0, 0), (0, window - f.shape[1])))
            feats.append(f)
            frames.append(max(2, min(window, piece.shape[0] // 160)))
        per_clip.append((first, len(feats)))
    out, offsets, counts, k = [], [], [], 0
    for t in input_ids:
        if t == tok and k < len(clips):
            lo, hi = per_clip[k]
            if fixed:
                n = fixed * (hi - lo)
            else:
                l1 = (frames[lo] - 1) // 2 + 1
                n = (l1 - 2) // 2 + 1
            h = int.from_bytes(hashlib.sha256(np.stack(feats[lo:hi]).tobytes()).digest()[:8], "little")
            offsets.append(len(out))
            counts.append(n)
            out.extend(int(h % max(1, V - 1)) if (j & 1) == 0 else int((h >> 17) % max(1, V - 1)) for j in range(n))
            k += 1
        else:
            out.append(t)
    if k < len(clips):
        raise ValueError(f"{len(clips)} audio clips but only {k} audio placeholder tokens in the prompt")
    return out, {"image_offsets": offsets, "image_token_counts": counts, "num_image_tokens": 0,
                 "audio_features": np.stack(feats) if feats else np.zeros((0, a["num_mel_bins"], window), np.float32), "audio_frames": frames}
