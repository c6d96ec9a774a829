"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed): sharding, the init-time table broadcast and the
end-of-step gather of accept bitmaps / MSM partial sums.  Pure tensor code, so the same functions run under the gloo
backend on CPU tensors in the tests and under NCCL on the GPU box."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous index range [lo, hi) of rank `rank` ([r*N/W, (r+1)*N/W))"""
    return (n * rank) // world, (n * (rank + 1)) // world


_WEIGHTS = {}


def pack_bitmap(ok):
    """uint8 0/1 vector (length multiple of 8) -> packed bits, LSB first"""
    if ok.numel() % 8:
        raise ValueError("bitmap length must be a multiple of 8")
    key = (ok.device.type, ok.device.index)
    if key not in _WEIGHTS:
        _WEIGHTS[key] = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=ok.device)
    return (ok.view(-1, 8) * _WEIGHTS[key]).sum(dim=1, dtype=torch.uint8)


def unpack_bitmap(packed):
    shifts = torch.arange(8, device=packed.device, dtype=torch.uint8)
    return ((packed.view(-1, 1) >> shifts) & 1).reshape(-1).to(torch.uint8)


def allgather_bitmaps(packed, out=None):
    """every rank contributes its packed bitmap (same length); returns the concatenation in rank order"""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return packed
    if out is None:
        out = torch.empty(world * packed.numel(), dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(out, packed)
    return out


def broadcast_table(table, src=0):
    """ONE broadcast of the generator comb table (uint8 tensor) from rank `src`"""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(table, src=src)
    return table


def replicate_engine_table(engine, rank, device, stream):
    """rank 0 exports its built table, everyone else imports the broadcast copy"""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return
    tab = torch.empty(engine.table_bytes, dtype=torch.uint8, device=device)
    if rank == 0:
        engine.table_export_dev(tab.data_ptr(), stream)
    broadcast_table(tab, 0)
    if rank != 0:
        engine.table_import_dev(tab.data_ptr(), stream)
    torch.cuda.synchronize()


def allgather_points(xy64):
    """MSM partial sums: 64-byte affine points, one per rank -> (world, 64)"""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return xy64.view(1, 64)
    out = torch.empty((world, 64), dtype=torch.uint8, device=xy64.device)
    dist.all_gather_into_tensor(out.view(-1), xy64.view(-1))
    return out


def allgather_partials(p96, out=None):
    """MSM shard sums: 96-byte Jacobian points (secp256k1_b200_ecmult_multi_partial_dev), one per rank -> (world, 96)"""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if out is None:
        out = torch.empty((world, 96), dtype=torch.uint8, device=p96.device)
    if world == 1:
        out.view(-1).copy_(p96.view(-1))
        return out
    dist.all_gather_into_tensor(out.view(-1), p96.view(-1))
    return out
