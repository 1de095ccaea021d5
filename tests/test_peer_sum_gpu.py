"""GPU: bplx_peer_sum, the peer-memory sum that finishes the sharded predictive grid.  On one GPU the "ranks" are buffers
of the same process and their kernels run on separate streams: the flag protocol and the rank-ordered sums are the same;
the NVLink path itself is exercised by scripts/multi_gpu_check.py and the bench sub-record under torchrun."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 5])
def test_peer_sum_flags_and_rank_ordered_sums(n):
    import torch
    from bpl_next_b200 import _abi

    lib = _abi.lib()
    FLAG = 64
    N0, N1 = 4 * 121 * 37, 3 * 37 + 2  # (a vectorised range and a ragged scalar one)
    g = torch.Generator(device="cuda").manual_seed(n)
    bufs = [torch.zeros(FLAG + N0 + N1 + 8, device="cuda") for _ in range(n)]
    ptrs = (C.c_void_p * n)(*[b.data_ptr() for b in bufs])
    streams = [torch.cuda.Stream() for _ in range(n)]
    outs = [(torch.empty(N0, device="cuda"), torch.empty(N1, device="cuda")) for _ in range(n)]
    for epoch in (1, 2, 3):
        for b in bufs:
            b[FLAG:].copy_(torch.randn(b.numel() - FLAG, generator=g, device="cuda") * 10 ** float(epoch))
        torch.cuda.synchronize()
        for r in range(n):  # every "rank" launches on its own stream: each kernel waits for all the others' flags
            with torch.cuda.stream(streams[r]):
                _abi.check(lib.bplx_peer_sum(ptrs, n, r, FLAG * 4, 0, N0, outs[r][0].data_ptr(), N0 + 1, N1,
                                             outs[r][1].data_ptr(), epoch, streams[r].cuda_stream))
        torch.cuda.synchronize()
        want0 = bufs[0][FLAG:FLAG + N0].clone()
        want1 = bufs[0][FLAG + N0 + 1:FLAG + N0 + 1 + N1].clone()
        for q in range(1, n):  # float32, rank order: the kernel's sums bit for bit
            want0 += bufs[q][FLAG:FLAG + N0]
            want1 += bufs[q][FLAG + N0 + 1:FLAG + N0 + 1 + N1]
        for r in range(n):
            assert torch.equal(outs[r][0], want0) and torch.equal(outs[r][1], want1)
        flags = torch.stack([b[:n].view(torch.int32) for b in bufs]).cpu().numpy()
        assert (flags == epoch).all()


def test_peer_sum_rejects_bad_arguments():
    import torch
    from bpl_next_b200 import _abi

    lib = _abi.lib()
    b = torch.zeros(256, device="cuda")
    ptrs = (C.c_void_p * 1)(b.data_ptr())
    out = torch.empty(16, device="cuda")
    assert lib.bplx_peer_sum(ptrs, 1, 1, 256, 0, 16, out.data_ptr(), 0, 0, None, 1, None) < 0     # rank outside the group
    assert lib.bplx_peer_sum(ptrs, 17, 0, 256, 0, 16, out.data_ptr(), 0, 0, None, 1, None) < 0    # too many ranks
    assert lib.bplx_peer_sum(ptrs, 1, 0, 2, 0, 16, out.data_ptr(), 0, 0, None, 1, None) < 0       # flag area too small
    assert lib.bplx_peer_sum(ptrs, 1, 0, 256, 0, 16, None, 0, 0, None, 1, None) < 0               # no output
