"""SURVEY.md section 8a row a14, pinned: csrc/scans.cu (and the CPU oracle's restatement) against the REFERENCE's own
SPN kernels (dmb/ops/spn/src/gaterecurrent2dnoind_kernel.cu:130-685), compiled from the reference tree into the
test-only oracle/_ref/libspn_ref.so by oracle/build_ref.py (built in the build container, shipped with the snapshot;
/root/reference is not read here).

The reference Function hands zero-initialised output / gradient buffers to the kernels
(functions/gaterecurrent2dnoind.py:13,31-34); so does this test.  Gates are signed and drawn at two magnitudes: with
|G1|+|G2|+|G3| < 1 everywhere (what AnyNet's caller guarantees by normalising them) and with sums up to 1.5, where
the recurrence is no longer a convex combination -- the kernels contain no clamping, and both must match.
"""
import ctypes
import os

import pytest
import torch

import dmb_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libspn_ref.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(LIB):
        pytest.fail("oracle/_ref/libspn_ref.so is missing: run `python oracle/build_ref.py` in the build container "
                    "(__graft_entry__.build() does) so that it ships with the snapshot")
    lib = ctypes.CDLL(LIB)
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.spn_ref_forward.argtypes = [I, I] + [P] * 5 + [I] * 4
    lib.spn_ref_backward.argtypes = [I, I] + [P] * 10 + [I] * 4
    return lib


def _ref_forward(lib, X, G, horizontal, reverse):
    out = torch.zeros_like(X)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    torch.cuda.synchronize()
    rc = lib.spn_ref_forward(int(horizontal), int(reverse), p(X), p(G[0]), p(G[1]), p(G[2]), p(out), *X.shape)
    assert rc == 0
    return out


def _ref_backward(lib, X, G, out, go, horizontal, reverse):
    go = go.clone()        # the reference accumulates the adjoint IN PLACE in grad_output (kernel.cu:318-322)
    grads = [torch.zeros_like(X) for _ in range(4)]
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    torch.cuda.synchronize()
    rc = lib.spn_ref_backward(int(horizontal), int(reverse), p(out), p(go), p(X), p(G[0]), p(G[1]), p(G[2]),
                              p(grads[0]), p(grads[1]), p(grads[2]), p(grads[3]), *X.shape)
    assert rc == 0
    return grads


@pytest.mark.parametrize("shape", [(2, 3, 5, 6), (1, 8, 34, 60), (1, 2, 1, 7), (1, 2, 7, 1), (1, 8, 64, 128), (2, 4, 37, 53)])
@pytest.mark.parametrize("gate_scale", [0.3, 0.5])
def test_spn_vs_reference_kernel(ref, shape, gate_scale):
    from densematchingbenchmark_b200.ops import GateRecurrent2dnoind
    N, C, H, W = shape
    g = torch.Generator().manual_seed(H * 131 + W)
    X = torch.randn(N, C, H, W, generator=g).to(DEV)
    G = [((torch.rand(N, C, H, W, generator=g) * 2 - 1) * gate_scale).to(DEV) for _ in range(3)]
    go = torch.randn(N, C, H, W, generator=g).to(DEV)
    for horizontal in (True, False):
        for reverse in (False, True):
            want = _ref_forward(ref, X, G, horizontal, reverse)
            wg = _ref_backward(ref, X, G, want, go, horizontal, reverse)
            Xg = X.clone().requires_grad_(True)
            Gg = [t.clone().requires_grad_(True) for t in G]
            out = GateRecurrent2dnoind(horizontal, reverse)(Xg, *Gg)
            # the same fp32 operations in the same order along the scan: agreement to rounding of fused multiply-adds
            torch.testing.assert_close(out.detach(), want, atol=2e-5, rtol=2e-5)
            out.backward(go)
            for name, got, w in zip("X G1 G2 G3".split(), [Xg.grad] + [t.grad for t in Gg], wg):
                torch.testing.assert_close(got, w, atol=5e-5, rtol=2e-4, msg=lambda m, n=name: "%s h=%s r=%s: %s" % (n, horizontal, reverse, m))
            # ... and the CPU oracle's restatement is pinned to the same reference outputs (removes its "unpinned" status)
            if N * C * H * W <= 20000:
                Xc, Gc = X.cpu(), [t.cpu() for t in G]
                o_cpu = O.spn_scan(Xc, *Gc, horizontal, reverse)
                torch.testing.assert_close(o_cpu, want.cpu(), atol=2e-5, rtol=2e-5)
                og = O.spn_scan_backward(Xc, *Gc, o_cpu, go.cpu(), horizontal, reverse)
                for got, w in zip(og, wg):
                    torch.testing.assert_close(got, w.cpu(), atol=5e-5, rtol=2e-4)
