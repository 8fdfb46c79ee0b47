"""Hole filling pinned against the REFERENCE KERNEL ITSELF: oracle/_ref/sam2_ref_C.so is the reference's
sam2/csrc/connected_components.cu compiled unmodified for sm_100a by oracle/build_ref.py (in the build container, from
the sources under /root/reference; the built module travels with the repo snapshot).

  * ds2_connected_components (the drop-in for sam2._C.get_connected_componnets, csrc/connected_components.cu:213-289):
    the same partition into 8-connected components (labels may be numbered differently) and the same per-pixel areas;
  * ds2_fill_holes against misc.py:365-393 evaluated with the reference kernel: bit-identical scores;
  * oracle/cc_oracle.c (the CPU checker used everywhere else) against the reference kernel too.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref():
    from oracle import build_ref
    if build_ref.built_path() is None:
        pytest.skip("oracle/_ref/sam2_ref_C.so was not built (python -m oracle.build_ref in the build container)")
    return build_ref.load()


def _masks(seed, N, H, W, kind):
    g = torch.Generator().manual_seed(seed)
    if kind == "noise":            # salt-and-pepper: thousands of tiny components, many of area <= 8
        m = torch.rand(N, 1, H, W, generator=g) < 0.55
    elif kind == "blobs":          # smooth field thresholded: a few large components with small holes
        x = torch.randn(N, 1, H // 8, W // 8, generator=g)
        x = torch.nn.functional.interpolate(x, size=(H, W), mode="bilinear", align_corners=False)
        m = (x + 0.15 * torch.randn(N, 1, H, W, generator=g)) > 0
    elif kind == "stripes":        # long thin diagonal structures: deep union-find chains, 8-connectivity matters
        yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        m = (((yy + xx) % 7) < 2)[None, None].expand(N, 1, H, W).clone()
        m ^= torch.rand(N, 1, H, W, generator=g) < 0.02
    else:                          # empty / full
        m = torch.zeros(N, 1, H, W, dtype=torch.bool)
        m[N // 2:] = True
    return m.to(torch.uint8).to(DEV)


def _same_partition(la, lb):
    """Two labelings describe the same partition iff the label pairs are in bijection."""
    la, lb = la.reshape(-1).long(), lb.reshape(-1).long()
    assert torch.equal(la > 0, lb > 0)
    pairs = torch.unique(torch.stack([la, lb], 1), dim=0)
    return len(torch.unique(pairs[:, 0])) == len(pairs) == len(torch.unique(pairs[:, 1]))


@pytest.mark.parametrize("kind", ["noise", "blobs", "stripes", "flat"])
@pytest.mark.parametrize("N,H,W", [(16, 256, 256), (3, 64, 96), (2, 128, 30)])
def test_connected_components_equal_reference_kernel(ref, kind, N, H, W):
    from detsam2_b200 import ops
    m = _masks(7, N, H, W, kind)
    rl, rc = ref.get_connected_componnets(m)
    ol, oc = ops.connected_components(m)
    torch.cuda.synchronize()
    assert torch.equal(oc, rc)                                   # per-pixel component area: identical integers
    for n in range(N):
        assert _same_partition(ol[n], rl[n]), n


@pytest.mark.parametrize("kind", ["noise", "blobs", "stripes"])
def test_fill_holes_equals_reference_formula_on_reference_kernel(ref, kind):
    """misc.py:365-393: labels, areas = cc(mask <= 0); mask = where((labels > 0) & (areas <= 8), 0.1, mask)."""
    from detsam2_b200 import ops
    from oracle import cc_oracle
    N, H, W = 16, 256, 256
    torch.manual_seed(3)
    sign = _masks(11, N, H, W, kind).float() * 2 - 1
    scores = (sign * (torch.rand(N, 1, H, W, device=DEV) * 5 + 0.01)).contiguous()
    labels, areas = ref.get_connected_componnets((scores <= 0).to(torch.uint8))
    want = torch.where((labels > 0) & (areas <= 8), torch.full_like(scores, 0.1), scores)
    got = scores.clone()
    lab = torch.empty((N, H, W), dtype=torch.int32, device=DEV)
    cnt = torch.empty_like(lab)
    ops.fill_holes(got, lab, cnt, N, H, W, 8)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    if kind != "stripes":                                        # (the stripe pattern leaves one connected background)
        assert int((got != scores).sum()) > 0                    # the case does fill something
    # the CPU checker used by the other tests agrees with the reference kernel as well
    cpu = cc_oracle.fill_holes(scores.cpu().numpy(), 8)
    assert np.array_equal(cpu.reshape(want.shape), want.cpu().numpy())
