"""Seeded clips and parameters of the encoder soaks (tools/encode_soak.py on the GPU, tools/soak_cpu.py encode on the CPU)."""
import numpy as np


def make(args):
    seed, i = args
    rng = np.random.default_rng([seed, i])
    n = int(rng.integers(100, 60000))
    t = np.arange(n)
    kind = i % 6
    if kind == 0:
        x = rng.standard_normal(n) * rng.uniform(0.001, 0.5)
    elif kind == 1:
        x = rng.uniform(0.05, 1.0) * np.sin(t * rng.uniform(0.001, 3.0))
    elif kind == 2:
        x = 0.5 * np.sin(t * t * rng.uniform(1e-6, 1e-4))
    elif kind == 3:
        x = (rng.standard_normal(n) * 0.4) * (np.sin(t * 0.002) > 0.7)
    elif kind == 4:
        x = rng.standard_normal(n) * 1e-4
    else:
        x = np.clip(rng.standard_normal(n) * 1.5, -1.0, 1.0)
    p = (int(rng.integers(0, 2)), int(rng.choice([0, 3])), int(rng.choice([8000, 32000, 64000, 96000, 128000, 192000, 256000, 512000])),
         float(rng.choice([0.5, 0.9, 0.97, 1.0])), float(rng.choice([1.0, 10.0, 100.0])) / 32768.0, float(rng.choice([0.0, 10.0, 200.0])) / 32768.0)
    fmt = int(rng.choice([0x9400, 0x9400, 0x9302, 0x9301]))
    if fmt == 0x9400:
        if i % 11 == 0:
            p = (-1, -1) + p[2:]                # the wildcard: every format tried, the first of the smallest kept
    elif fmt == 0x9302:
        p = (int(rng.choice([0, 1, -1])), 0) + p[2:]
    else:
        p = (0, 0) + p[2:]                      # $9301: stream type 0 (nobody encodes OS93a type 1)
    return x.astype(np.float32), p + (fmt,)
