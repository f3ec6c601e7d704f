#!/usr/bin/env python
"""Randomised parity sweep of the padded-tensor convolution paths added late in round 2 (weight-stationary 256-row tiles, image
tiles with a 9-pixel descriptor group stride, the 4-D store of the 1x1 stride-2 layers), against torch fp32 - the body of
tests/test_gpu_conv.py::test_padded_conv_matches_torch over batch sizes and channel widths the fixed cases do not hold.
    python scripts/conv_fuzz.py [cases]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_conv as T  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(20261019)
geoms = [  # H, W, k, s, (cin choices), (cout choices)
    (16, 8, 3, 1, (64, 128, 256), (64, 96, 128, 256)),     # image tiles
    (64, 32, 3, 1, (64,), (64,)),                           # 256-row tiles, weight-stationary pairs
    (32, 16, 3, 1, (64, 128), (64, 128)),                   # pairs / flat raster
    (8, 4, 3, 1, (64, 256), (64, 512)),
    (64, 32, 1, 2, (64,), (128,)), (32, 16, 1, 2, (128,), (256,)), (16, 8, 1, 2, (256,), (512,)),   # 4-D store of whole rows / images
    (64, 32, 3, 2, (64,), (128,)), (16, 8, 3, 2, (256,), (512,)),
]
done = 0
for i in range(n):
    H, W, k, s, cins, couts = geoms[int(rng.integers(len(geoms)))]
    B = int(rng.choice([1, 2, 3, 5, 17, 40, 131, 300]))
    cin, cout = int(rng.choice(cins)), int(rng.choice(couts))
    act = int(rng.integers(0, 3))
    res_mode = int(rng.integers(0, 3)) if (s == 1 and cin == cout or s == 1) else 0
    if res_mode and k == 1:
        res_mode = 0
    kind = int(rng.integers(1, 3))
    case = ("fuzz%d_%dx%d_b%d_c%d_%d_k%ds%d_a%d_r%d" % (i, H, W, B, cin, cout, k, s, act, res_mode), B, H, W, cin, cout, k, s, act, res_mode, 1, 1)
    T.test_padded_conv_matches_torch(case, kind)
    done += 1
print("conv_fuzz: %d randomised padded-convolution cases match torch fp32 (both border kinds, batches 1 - 300)" % done)
