"""Minimal RIFF/WAVE reader (PCM 8/16/24/32-bit and IEEE float32/64) -> float32 [channels, samples] in [-1, 1).

Stands in for ``torchaudio.load`` at reference feature_extractor.py:43 (integer PCM is scaled by 2^-(bits-1),
as torchaudio does); the DCASE recordings are 24 kHz 4-channel int16.
"""
import struct

import numpy as np
import torch


def load_wav(path: str):
    with open(path, 'rb') as fh:
        data = fh.read()
    if data[:4] != b'RIFF' or data[8:12] != b'WAVE':
        raise ValueError(f'{path}: not a RIFF/WAVE file')
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        tag, size = data[pos:pos + 4], struct.unpack('<I', data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if tag == b'fmt ':
            fmt = struct.unpack('<HHIIHH', body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:          # WAVE_FORMAT_EXTENSIBLE: real tag in the sub-format GUID
                fmt = (struct.unpack('<H', body[24:26])[0],) + fmt[1:]
        elif tag == b'data':
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None:
        raise ValueError(f'{path}: missing fmt or data chunk')
    tag, n_chan, rate, _, _, bits = fmt
    if tag == 1:
        if bits == 8:
            x = (np.frombuffer(payload, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(payload, dtype='<i2').astype(np.float32) / 32768.0
        elif bits == 24:
            raw = np.frombuffer(payload[:len(payload) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)
            v = np.where(v & 0x800000, v - 0x1000000, v)
            x = v.astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(payload, dtype='<i4').astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise ValueError(f'{path}: unsupported PCM width {bits}')
    elif tag == 3:
        x = np.frombuffer(payload, dtype='<f4' if bits == 32 else '<f8').astype(np.float32)
    else:
        raise ValueError(f'{path}: unsupported WAVE format tag {tag}')
    n = x.size // n_chan
    wav = np.ascontiguousarray(x[:n * n_chan].reshape(n, n_chan).T)
    return torch.from_numpy(wav), int(rate)
