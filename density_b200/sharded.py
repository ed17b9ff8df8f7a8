"""Multi-GPU sharding of ONE bit-exact Chameleon stream: one process per GPU, torch.distributed for the
only exchange step the path has — the 256 KiB last-writer tables.

Rank r owns bytes [r*S, (r+1)*S) of the stream (S a multiple of 256 so block grids line up).
  phase 1 (local)   flag pass with unknown carry-in; exports the shard's last-writer table   (density_b200_shard_phase1)
  exchange          all_gather of the tables (world * 256 KiB), left fold of ranks < r        (this file)
  phase 2 (local)   resolve first touches against the carried-in dictionary, scan, emit       (density_b200_shard_phase2)
The concatenation of the per-rank outputs is byte-identical to one chameleon_encode call over the whole buffer as long as
the protection automaton stays quiet (flags bit 0 reports otherwise). Outputs stay where they were produced; a caller
that wants them on one rank gathers them with the sizes returned here.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib

TABLE_ENTRIES = 65536
TOUCHED = 0x10000


def initial_table(device):
    """Dictionary state at the stream start: only bucket 0 'holds quad 0' (chameleon.rs:41,89-91)."""
    t = torch.zeros(TABLE_ENTRIES, dtype=torch.int32, device=device)
    t[0] = TOUCHED
    return t


def fold_tables(gathered, rank):
    """carry-in of `rank` = left fold of the tables of ranks < rank over the initial state.
    gathered: int32 [world, 65536] (bit 16 = touched, low 16 bits = fingerprint). Pure tensor ops: runs on CPU (gloo
    tests) and CUDA alike."""
    carry = initial_table(gathered.device)
    for r in range(rank):
        t = gathered[r]
        carry = torch.where((t & TOUCHED) != 0, t, carry)
    return carry


def exchange_tables(table, group=None):
    world = dist.get_world_size(group)
    gathered = torch.empty((world, TABLE_ENTRIES), dtype=torch.int32, device=table.device)
    dist.all_gather_into_tensor(gathered.view(-1), table.contiguous(), group=group)
    return gathered


class ShardedChameleonEncoder:
    def __init__(self):
        self._lib = _lib.load()
        self._h = self._lib.density_b200_shard_create()
        if not self._h:
            raise _lib.DensityB200Error(_lib.last_error())

    def close(self):
        if self._h:
            self._lib.density_b200_shard_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, d_in, d_out, d_size, d_flags, group=None):
        """d_in / d_out: CUDA uint8 tensors (this rank's shard / its output buffer); d_size: int64[1]; d_flags: int32[1].
        Everything is enqueued on torch's current stream; the all_gather is the only collective."""
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        table = torch.empty(TABLE_ENTRIES, dtype=torch.int32, device=d_in.device)
        rc = self._lib.density_b200_shard_phase1(self._h, d_in.data_ptr(), d_in.numel(), int(rank == world - 1),
                                                 table.data_ptr(), stream)
        if rc:
            raise _lib.DensityB200Error(f"shard_phase1 rc={rc}: {_lib.last_error()}")
        carry_ptr = None
        if world > 1:
            gathered = exchange_tables(table, group)
            if rank > 0:
                self._carry = fold_tables(gathered, rank)
                carry_ptr = self._carry.data_ptr()
        rc = self._lib.density_b200_shard_phase2(self._h, carry_ptr, d_out.data_ptr(), d_out.numel(), d_size.data_ptr(),
                                                 d_flags.data_ptr(), stream)
        if rc:
            raise _lib.DensityB200Error(f"shard_phase2 rc={rc}: {_lib.last_error()}")


class ShardedEncoder:
    """The C++ multi-GPU path (`density_b200_encode_sharded`, include/density_b200.h): one process per GPU; the library owns its NCCL
    communicator (the 128-byte id travels once through torch.distributed), the table all-gather, the single fold kernel, the exact
    seam verdict and the optional variable-length gather of the pieces to one rank. torch.distributed is only used to hand out the id.
    """

    def __init__(self, device, group=None):
        self._lib = _lib.load()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.world > 1:
            if self.rank == 0:
                buf = (ctypes.c_uint8 * 128)()
                rc = self._lib.density_b200_sharded_unique_id(buf)
                if rc:
                    raise _lib.DensityB200Error(f"sharded_unique_id rc={rc}: {_lib.last_error()}")
                ident = torch.tensor(list(buf), dtype=torch.uint8)
            t = ident.to(device) if dist.get_backend(group) == "nccl" else ident
            dist.broadcast(t, src=0, group=group)
            ident = t.cpu()
        self._id = ident.contiguous()
        self._h = self._lib.density_b200_sharded_create(self._id.data_ptr() if self.world > 1 else None, self.rank, self.world)
        if not self._h:
            raise _lib.DensityB200Error(_lib.last_error())
        self.d_total = torch.zeros(1, dtype=torch.int64, device=device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.density_b200_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, d_in, d_out, d_size, d_flags, gather_root=-1, d_gather=None):
        """Enqueue on torch's current stream. d_size int64[1]: this rank's piece; d_flags int32[1]: != 0 -> the stream is not quiet and
        the pieces are void; self.d_total int64[1]: stream length. gather_root >= 0: pieces gathered into d_gather on that rank (blocks)."""
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = self._lib.density_b200_encode_sharded(self._h, d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), d_out.numel(), d_size.data_ptr(),
                                                   d_flags.data_ptr(), self.d_total.data_ptr(), int(gather_root),
                                                   d_gather.data_ptr() if d_gather is not None else None,
                                                   d_gather.numel() if d_gather is not None else 0, stream)
        if rc:
            raise _lib.DensityB200Error(f"encode_sharded rc={rc}: {_lib.last_error()}")

    def profile(self):
        """stage times (ms) of the last call: flag pass, table exchange + fold, carry / resolve / sizes / scan, emit, seams + gather"""
        out = (ctypes.c_float * 5)()
        rc = self._lib.density_b200_sharded_profile(self._h, out)
        if rc:
            raise _lib.DensityB200Error(f"sharded_profile rc={rc}: {_lib.last_error()}")
        return [float(x) for x in out]
