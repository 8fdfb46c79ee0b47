"""Input-independent tables the engine builds on the host (no GPU needed)."""
import torch


def test_rope_axial_table_equals_full_table():
    """engine._rope_axial is engine._rope_table read two ways, bit for bit (host arithmetic only)."""
    from detsam2_b200.engine import _rope_axial, _rope_table
    for side in (64, 32, 16):
        full, ax = _rope_table(256, side, 10000.0), _rope_axial(256, side, 10000.0)   # [128, side^2, 2], [64, side, 2]
        pos = torch.arange(side * side)
        assert torch.equal(full[:64], ax[:, pos % side])
        assert torch.equal(full[64:], ax[:, pos // side])
