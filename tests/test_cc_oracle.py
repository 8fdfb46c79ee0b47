"""Pins the C restatement of the connected-components / hole-filling path (oracle/cc_oracle.c, which
follows sam2/csrc/connected_components.cu:213-282 and sam2/utils/misc.py:365-393) against an
independent implementation: scipy.ndimage.label with 8-connectivity.  The reference's own kernel is a
CUDA/ATen extension that cannot run on CPU (and is silently skipped there, misc.py:389-391)."""
import numpy as np
import pytest
from scipy import ndimage

from oracle import cc_oracle

EIGHT = np.ones((3, 3), dtype=np.int32)


def _areas_scipy(m):
    lab, n = ndimage.label(m, structure=EIGHT)
    cnt = np.bincount(lab.ravel(), minlength=n + 1)
    area = cnt[lab]
    area[lab == 0] = 0
    return lab, area


@pytest.mark.parametrize("shape,density,seed", [((1, 1, 32, 32), 0.5, 0), ((3, 1, 64, 48), 0.3, 1),
                                                ((2, 1, 256, 256), 0.6, 2), ((1, 1, 7, 5), 0.9, 3)])
def test_component_areas_match_scipy(shape, density, seed):
    rng = np.random.default_rng(seed)
    m = (rng.random(shape) < density).astype(np.uint8)
    labels, counts = cc_oracle.connected_components(m)
    for n in range(shape[0]):
        lab_s, area_s = _areas_scipy(m[n, 0])
        np.testing.assert_array_equal(counts[n, 0], area_s)
        # same partition: one oracle label per scipy label and vice versa
        fg = m[n, 0] > 0
        pairs = set(zip(labels[n, 0][fg].tolist(), lab_s[fg].tolist()))
        assert len(pairs) == len({a for a, _ in pairs}) == len({b for _, b in pairs})
        assert (labels[n, 0][~fg] == 0).all() or (counts[n, 0][~fg] == 0).all()


def test_empty_and_full_masks():
    z = np.zeros((2, 1, 16, 16), np.uint8)
    _, c = cc_oracle.connected_components(z)
    assert (c == 0).all()
    o = np.ones((1, 1, 16, 16), np.uint8)
    _, c = cc_oracle.connected_components(o)
    assert (c == 256).all()


def test_fill_holes_semantics():
    """misc.py:365-393: background (score <= 0) components of area <= max_area become +0.1."""
    s = np.full((1, 20, 20), 5.0, np.float32)
    s[0, 2:4, 2:4] = -1.0          # 4-pixel hole -> filled
    s[0, 10:13, 10:13] = -2.0      # 9-pixel hole -> kept (area > 8)
    s[0, 5, 5] = 0.0               # score == 0 counts as background
    s[0, 6, 6] = -3.0              # diagonal neighbour: same 8-connected component (area 2) -> filled
    out = cc_oracle.fill_holes(s, 8)
    assert np.allclose(out[0, 2:4, 2:4], 0.1)
    assert np.allclose(out[0, 10:13, 10:13], -2.0)
    assert np.isclose(out[0, 5, 5], 0.1) and np.isclose(out[0, 6, 6], 0.1)
    fg = s > 0
    assert np.array_equal(out[fg], s[fg])
