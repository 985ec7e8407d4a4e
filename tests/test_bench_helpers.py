"""CPU: the byte/flop accounting bench.py reports equals SURVEY.md 8(d) / BASELINE.md section 5."""
import numpy as np

import bench


def test_algorithmic_bytes_match_survey_table():
    build, lookup, flops = bench.algorithmic_bytes(bench.WORKLOADS["cfg1"])
    assert abs(build / 1e6 - 124.7) < 0.1 and abs(lookup / 1e6 - 9.98) < 0.01 and abs(flops / 1e9 - 3.98) < 0.01
    build, lookup, flops = bench.algorithmic_bytes(bench.WORKLOADS["cfg3"])
    assert abs(build / 1e6 - 1051.4) < 0.5 and abs(lookup / 1e6 - 73.80) < 0.05 and abs(flops / 1e9 - 38.28) < 0.05


def test_synthetic_coords_cover_out_of_range():
    wl = dict(bench.WORKLOADS["cfg1"], H=8, iters=4, C=4)
    f1, f2, coords = bench.synth_inputs(wl, seed=0)
    assert coords.shape == (4, 1, 2, 8, 240) and f1.dtype == np.float32
    x = coords[:, :, 0]
    assert (x < 0).mean() >= 0.01 and (x > 239).mean() >= 0.01      # >=1 % each side
    assert np.array_equal(coords[0, 0, 1, :, 0], np.arange(8, dtype=np.float32))
