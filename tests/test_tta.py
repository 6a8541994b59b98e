"""TTA semantics: D4 codes against numpy, and the cumulative in-place quirk against the oracle restatement and
against golden vectors produced by executing the reference's own apply_tta / transform_prob."""
import json
import os

import numpy as np
import pytest

from digipathai_b200 import tta
from oracle import pipeline_ref

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ['FLIP_LEFT_RIGHT', 'ROTATE_90', 'ROTATE_180', 'ROTATE_270']


def test_codes_match_numpy():
    a = np.arange(7 * 7 * 3).reshape(7, 7, 3)
    assert np.array_equal(tta.apply(tta.code_of('FLIP_LEFT_RIGHT'), a), np.fliplr(a))
    for k, nm in ((1, 'ROTATE_90'), (2, 'ROTATE_180'), (3, 'ROTATE_270')):
        assert np.array_equal(tta.apply(tta.code_of(nm), a), np.rot90(a, k))
    for g in range(8):
        assert np.array_equal(tta.apply(tta.inverse(g), tta.apply(g, a)), a)


def test_unknown_names_are_identity_passes():
    # utils.py:490-499: anything else falls into `else: img = img` but still counts as a pass
    assert tta.code_of('flip_lr') == 0 and tta.code_of('flip_ud') == 0 and tta.code_of('rot90') == 0
    assert tta.pass_codes(['flip_lr', 'rot90']) == [(0, 0), (0, 0), (0, 0)]


@pytest.mark.parametrize("lst", [None, ['FLIP_LEFT_RIGHT'], ['FLIP_LEFT_RIGHT', 'ROTATE_90'],
                                 ['ROTATE_90', 'ROTATE_90', 'FLIP_LEFT_RIGHT'], NAMES, ['bogus', 'ROTATE_270']])
def test_pass_codes_reproduce_the_cumulative_inplace_loop(lst):
    """Drive the oracle's literal loop (Segmentation.py:150-158) with an identity 'model' and compare."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 6, 6, 2)).astype(np.float32)
    names = ['DEFAULT'] + (lst or [])
    patches = x.copy()
    want_in, want_out = [], []
    for nm in names:
        patches = pipeline_ref.apply_tta(patches, nm)
        want_in.append(patches.copy())
        want_out.append(pipeline_ref.transform_prob(patches.copy(), nm))
    for k, (cin, cout) in enumerate(tta.pass_codes(lst)):
        got_in = np.stack([tta.apply(cin, t) for t in x])
        assert np.array_equal(got_in, want_in[k])
        got_out = np.stack([tta.apply(tta.inverse(cout), t) for t in got_in])
        assert np.array_equal(got_out, want_out[k])


def test_golden_vectors_from_the_reference_functions():
    path = os.path.join(HERE, "golden", "tta_golden.npz")
    z = np.load(path)
    x = z["x"]
    seqs = json.loads(str(z["sequences"]))
    for si, seq in enumerate(seqs):
        patches = x.copy()
        for k, nm in enumerate(['DEFAULT'] + seq):
            patches = pipeline_ref.apply_tta(patches, nm)
            assert np.array_equal(patches, z[f"in_{si}_{k}"])
            assert np.array_equal(pipeline_ref.transform_prob(patches.copy(), nm), z[f"out_{si}_{k}"])
        for k, (cin, cout) in enumerate(tta.pass_codes(seq)):
            gi = np.stack([tta.apply(cin, t) for t in x])
            assert np.array_equal(gi, z[f"in_{si}_{k}"])
            assert np.array_equal(np.stack([tta.apply(tta.inverse(cout), t) for t in gi]), z[f"out_{si}_{k}"])
