"""Measured max-abs errors of the CUDA extractor against the committed golden outputs of the unmodified reference
(tests/golden) -- the numbers quoted in README.md / DESIGN.md.

    python tools/report_errors.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from cases import CASES, case_input, check_features, input_matches_golden, load_golden  # noqa: E402

from seld_b200 import feature_extractor as fe  # noqa: E402

worst = {'foa': [0.0, 0.0], 'mic': [0.0, 0.0]}
for name in CASES:
    wav, sr, n_mels, kw = case_input(name)
    g = load_golden(name)
    if not input_matches_golden(wav, g):
        print(json.dumps({'case': name, 'skipped': 'input differs from fixture'}))
        continue
    for mode in ('foa', 'mic'):
        got = fe.extract_features(wav, sr, mode=mode, n_mels=n_mels, **kw)
        e_mel, e_rest = check_features(got, g[mode], mode, f'{name}/{mode}')
        worst[mode][0] = max(worst[mode][0], e_mel)
        worst[mode][1] = max(worst[mode][1], e_rest)
        print(json.dumps({'case': name, 'mode': mode, 'logmel_max_abs_dB': e_mel, ('iv' if mode == 'foa' else 'gcc') + '_max_abs': e_rest}))
print(json.dumps({'worst': {'logmel_dB': max(worst['foa'][0], worst['mic'][0]), 'iv': worst['foa'][1], 'gcc_tensor_core_path': worst['mic'][1]},
                  'tolerances': {'logmel_dB': 1e-4, 'iv': 1e-3, 'gcc': 1e-3}}))
