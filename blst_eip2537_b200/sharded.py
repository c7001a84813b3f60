"""Multi-GPU MULTIEXP: points sharded across ranks, partial sums all-gathered, summed and encoded.

One process per GPU (torch.distributed, NCCL over NVLink on the B200 box; gloo in CPU tests).
The sum over pairs is associative and commutative, so each rank runs the whole single-GPU
pipeline (decode -> Pippenger) on its contiguous slice and contributes one XYZZ partial sum
(192 B for G1, 384 B for G2).  The only exchange step is an all-gather of those partials plus a
min-reduction of the "first failing pair" key, which reproduces the reference's sequential error
precedence (/root/reference/src/eip2537.c:580-592, :650-668) across shards.

The compute backend is pluggable so the host logic can be tested without a GPU (tests plug in
the CPU oracle); the default backend is the CUDA engine and there is no CPU fallback in it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

STATUS_OK = (1 << 63) - 1   # "no error" as a signed 64-bit key so MIN picks real errors


def shard_range(n_pairs: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of the pair index space owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class CudaBackend:
    """Partial MSM / combine on the current CUDA device through the C ABI (include/eip2537_b200.h)."""

    def __init__(self, group: int):
        from . import _native
        self.L = _native.lib()
        self.group = group
        self.xy = int(self.L.bls12_b200_partial_bytes(group))
        self.plen = 128 if group == 1 else 256
        self.device = torch.device("cuda", torch.cuda.current_device())

    def partial(self, d_in: torch.Tensor, n: int, index_base: int):
        part = torch.zeros(self.xy, dtype=torch.uint8, device=self.device)
        status = torch.full((1,), -1, dtype=torch.int64, device=self.device)
        if n:
            s = torch.cuda.current_stream().cuda_stream
            rc = self.L.bls12_b200_msm_partial_device(self.group, d_in.data_ptr(), n, index_base, part.data_ptr(), status.data_ptr(), s)
            if rc != 0:
                raise RuntimeError("msm_partial_device: %d %s" % (rc, self.L.bls12_b200_last_error()))
        status = torch.where(status < 0, torch.full_like(status, STATUS_OK), status)
        return part, status

    def combine(self, parts: torch.Tensor, count: int) -> bytes:
        out = torch.zeros(self.plen, dtype=torch.uint8, device=self.device)
        s = torch.cuda.current_stream().cuda_stream
        rc = self.L.bls12_b200_msm_combine_device(self.group, parts.data_ptr(), count, out.data_ptr(), s)
        if rc != 0:
            raise RuntimeError("msm_combine_device: %d" % rc)
        return bytes(out.cpu().numpy())


def sharded_multiexp(local_pairs: torch.Tensor, n_local: int, index_base: int, backend, group=None):
    """Every rank passes ITS slice (wire bytes, already on the backend's device).

    Returns (code, out_bytes_or_None) on every rank -- identical to what the single-call
    bls12_g{1,2}multiexp would return for the concatenated input.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    part, status = backend.partial(local_pairs, n_local, index_base)
    if world > 1:
        gathered = torch.zeros(world * part.numel(), dtype=part.dtype, device=part.device)
        dist.all_gather_into_tensor(gathered, part, group=group)
        dist.all_reduce(status, op=dist.ReduceOp.MIN, group=group)
    else:
        gathered = part
    key = int(status.item())
    if key != STATUS_OK:
        return key & 0xFF, None
    return 0, backend.combine(gathered, world)
