"""Host-side logic of the constraint-sharded solve (one process per rank).  Pure Python + torch.distributed for the
rendezvous; the per-pivot exchanges are done by libb2s's kernels over peer memory (or by NCCL calls it issues when peer
memory is switched off).  Two bootstraps: NCCL (init_sharded_solver: one rank per GPU) and the host layer
(init_sharded_solver_host: any torch.distributed backend, ranks may share a GPU).  Everything here also runs under the gloo
backend on CPU (tests)."""
from . import _lib as L

STAGE1_BLOCK = 512  # reference stage-1 block (src/reduction.cu:6): slabs are multiples of it


def slab(rank, world, m):
    """Constraints [lo, hi) owned by `rank`; m must split into equal multiples of 512."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    if m % (world * STAGE1_BLOCK) != 0:
        raise ValueError(f"constraints ({m}) must be a multiple of world*{STAGE1_BLOCK} = {world * STAGE1_BLOCK}")
    width = m // world
    return rank * width, (rank + 1) * width


def owner_of(p, world, m):
    """Rank that stores constraint p."""
    return p // (m // world)


def stage1_blocks(rank, world, m):
    """Global ids of the reference stage-1 blocks whose winners this rank contributes to the all-gather."""
    lo, hi = slab(rank, world, m)
    return range(lo // STAGE1_BLOCK, hi // STAGE1_BLOCK)


def exchange_unique_id(dist, make_id):
    """Rank 0 creates the NCCL unique id (make_id()), everybody receives the same 128 bytes."""
    box = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != L.NCCL_ID_BYTES:
        raise RuntimeError("unique id exchange failed")
    return bytes(uid)


def init_sharded_solver(solver, dist):
    """Create the solver's NCCL communicator across the ranks of the torch.distributed job."""
    from .solver import dist_unique_id
    uid = exchange_unique_id(dist, dist_unique_id)
    solver.dist_init(dist.get_rank(), dist.get_world_size(), uid)
    return solver


def make_host_allgather(dist, group=None):
    """A b2s_allgather_fn over torch.distributed CPU tensors: every rank's `nbytes` bytes, in rank order, into recv."""
    import ctypes
    import numpy as np
    import torch
    world = dist.get_world_size(group)

    def allgather(send, recv, nbytes):
        src = np.ctypeslib.as_array((ctypes.c_ubyte * nbytes).from_address(send))
        mine = torch.from_numpy(src.copy())
        outs = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(outs, mine, group=group)
        dst = np.ctypeslib.as_array((ctypes.c_ubyte * (nbytes * world)).from_address(recv))
        dst[:] = torch.cat(outs).numpy()
        return 0
    return allgather


def init_sharded_solver_host(solver, dist, group=None):
    """Sharded solver without NCCL: CUDA-IPC handles, barriers, the price-out chain and the solution travel through
    torch.distributed CPU collectives (needs a group whose backend takes CPU tensors, e.g. gloo); the per-pivot
    exchanges are the library's peer-memory kernels.  Several ranks may use the same GPU."""
    if group is None and dist.get_backend() != "gloo":
        group = dist.new_group(backend="gloo")
    solver.dist_init_host(dist.get_rank(), dist.get_world_size(), make_host_allgather(dist, group))
    return solver
