"""The one-shot peer exchange kernels (csrc/peer_sync.cu) with a single-rank context: the protocol (push, sequence
flags, parity alternation, epoch counter) runs exactly as with N ranks, the sum is over one slot. The N>1 behaviour is
checked on real GPUs by tools/dp_check.py (torchrun, 2+ ranks)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ctx():
    from gan_playground_b200 import ops

    buf = torch.zeros(ops.peer_buffer_bytes() // 4, device="cuda")
    epoch = torch.zeros(1, dtype=torch.int32, device="cuda")
    return ops.make_peer_ctx([buf.data_ptr()], 0, epoch), buf, epoch


def test_peer_allreduce_single_rank_is_identity_and_advances_epoch():
    from gan_playground_b200 import ops

    ctx, buf, epoch = _ctx()
    for n in (1, 7, 256, 2048, 4096, 33):
        t = torch.randn(n, device="cuda")
        ref = t.clone()
        ops.peer_allreduce_sum_(ctx, t)
        assert torch.equal(t, ref)
    assert epoch.item() == 6
    with pytest.raises(ops._lib.GpError):
        ops.peer_allreduce_sum_(ctx, torch.zeros(4097, device="cuda"))


def test_bn_finalize_peer_matches_bn_finalize():
    from gan_playground_b200 import ops

    ctx, buf, epoch = _ctx()
    C, count = 384, 4096.0
    x = torch.randn(4096, C, device="cuda") * 2 + 0.5
    st = torch.stack([x.sum(0), (x * x).sum(0)]).contiguous()
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    rm1, rv1, n1 = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
    rm2, rv2, n2 = rm1.clone(), rv1.clone(), n1.clone()
    a = ops.bn_finalize(st.clone(), count, gamma, beta, rm1, rv1, n1)
    b = ops.bn_finalize_peer(ctx, st.clone(), count, gamma, beta, rm2, rv2, n2)
    assert torch.equal(a, b)
    assert torch.equal(rm1, rm2) and torch.equal(rv1, rv2) and n1.item() == n2.item() == 1
