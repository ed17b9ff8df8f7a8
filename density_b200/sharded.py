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
