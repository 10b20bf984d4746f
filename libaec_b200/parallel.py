"""Multi-GPU sharding of one AEC stream.

One process per GPU.  The input is split into contiguous ranges of whole RSIs;
the predictor and the zero-run logic reset at every RSI, so the only things
that cross a shard boundary are the bit position and the split position k of
the previous block.  Protocol per encode:

  1. every rank codes its shard from carry (0 bits, k = 0)          [CUDA, no comm]
  2. ONE all_gather of (bits, klo, khi, last 64 bits) per shard -- 32 bytes
     per rank                                                        [NCCL]
  3. a rank whose true incoming k is not 0 re-codes its leading tiles
     (the lengths do not depend on k, only the ids / split bits do)  [CUDA]
  4. every rank moves its stream to its bit offset in the global stream
     (aecb200_place_bits_device, a funnel-shift copy) and completes its
     first word with the predecessor's tail bits                     [CUDA]

`ShardedCodec.step_enqueue` runs steps 1-4 without the host in between: the
encoder's summary kernel writes the 32 bytes into device memory, NCCL gathers
them on the codec's stream, a one-thread kernel works out the plan
(aecb200_shard_plan_device), the repair launch reads its incoming k and tile
count from that plan and the placement kernel its bit offset.

After step 4 rank r holds exactly the bytes [4*word_lo, ...) of the single
stream it owns; the concatenation over ranks is byte-identical to the stream a
single GPU (or the CPU reference) produces for the whole input.  Decode needs
no exchange at all: each rank decodes its RSIs from its own segment.

The pure functions at the top carry the host-side logic and are what the gloo
CPU tests exercise; `ShardedCodec` wires them to the device layer.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


REPAIR_TILES = 64        # AECB200_REPAIR_TILES of include/aec_b200.h


def shard_range(total_samples: int, rsi_samples: int, rank: int, world: int):
    """Contiguous RSI-aligned shard [start, start + count) of `rank`."""
    nrsi = (total_samples + rsi_samples - 1) // rsi_samples
    per = (nrsi + world - 1) // world
    s = min(rank * per, nrsi) * rsi_samples
    e = min(min((rank + 1) * per, nrsi) * rsi_samples, total_samples)
    return s, max(e - s, 0)


def clamp(v: int, lo: int, hi: int) -> int:
    return lo if v < lo else (hi if v > hi else v)


@dataclass
class ShardPlan:
    bit_offset: int      # where this shard's first bit sits in the global stream
    k_in: int            # k of the last coded block before this shard
    end_bit: int         # bit_offset + bits
    word_lo: int         # first 32-bit word of the global stream this rank owns
    word_hi: int         # one past the last word it owns
    total_bits: int
    head_or: int         # predecessor bits that share this shard's first word (already in place)


def boundary_bits(prev_tail64: int, bit_offset: int) -> int:
    """The bit_offset % 32 bits in front of a shard inside its first word: the
    last bits of the stream so far, taken from the predecessor's last 64 bits."""
    n = bit_offset & 31
    if n == 0:
        return 0
    return ((prev_tail64 & ((1 << n) - 1)) << (32 - n)) & 0xFFFFFFFF


def plan_shards(infos, seed_k: int = 0):
    """infos: per rank (bits, klo, khi[, tail64]).  Exclusive scan of the bit
    lengths and the clamp chain of k (k_out = clamp(k_in, klo, khi)).
    Shards must be at least 64 bits long unless they are empty."""
    plans = []
    off, k = 0, seed_k
    total = sum(int(i[0]) for i in infos)
    prev_tail = 0
    for r, info in enumerate(infos):
        bits, lo, hi = int(info[0]), int(info[1]), int(info[2])
        end = off + bits
        last = r == len(infos) - 1
        plans.append(ShardPlan(off, k, end, off >> 5, ((end + 31) >> 5) if last else (end >> 5), total,
                               boundary_bits(prev_tail, off)))
        k = clamp(k, lo, hi)
        off = end
        if len(info) > 3 and bits:
            prev_tail = int(info[3]) & 0xFFFFFFFFFFFFFFFF
    return plans


def place_bits_host(stream: np.ndarray, nbits: int, dst_bit: int, head_or: int = 0) -> np.ndarray:
    """numpy model of aecb200_place_bits_device (CPU tests): returns the bytes of
    the 32-bit words [dst_bit >> 5, (dst_bit + nbits + 31) >> 5) with the
    stream's bits at their place, head_or in front and zeros behind."""
    bits = np.unpackbits(stream)[:nbits]
    w0 = dst_bit >> 5
    nw = ((dst_bit + nbits + 31) >> 5) - w0
    out = np.zeros(nw * 32, dtype=np.uint8)
    s = dst_bit - (w0 << 5)
    if s:
        out[:s] = np.unpackbits(np.array([head_or], dtype=">u4").view(np.uint8))[:s]
    out[s:s + nbits] = bits
    return np.packbits(out)


def tail64_host(stream: np.ndarray, nbits: int) -> int:
    """Last 64 bits of a stream, right-aligned (numpy model of the shard summary)."""
    bits = np.unpackbits(stream)[:nbits][-64:]
    v = 0
    for b in bits.tolist():
        v = (v << 1) | int(b)
    return v


class ShardedCodec:
    """Device-side sharded encoder/decoder for one rank."""

    def __init__(self, params, rank: int, world: int, device: int, group=None, stream=None):
        import ctypes
        import torch
        self._C = ctypes
        from .api import DeviceCodec, encode_bound
        self.torch = torch
        self.p = params
        self.rank, self.world = rank, world
        self.group = group
        self.codec = DeviceCodec(device=device, stream=stream)
        self.codec.set_shard_mode(True)
        self._bound = encode_bound
        self.local = None
        self.placed = None
        self.offsets = None
        self.grp = None
        self.plan = None
        self._mine = torch.zeros(4, dtype=torch.int64, device="cuda")
        self._all = torch.zeros(4 * world, dtype=torch.int64, device="cuda")
        self._mine_h = torch.zeros(4, dtype=torch.int64).pin_memory()
        self._all_h = torch.zeros(4 * world, dtype=torch.int64).pin_memory()
        # the exchange runs on a stream of its own, next to whatever the caller enqueues on the
        # codec's stream after encode_local (the decode of the local shard needs no exchange)
        self._xstream = torch.cuda.Stream()
        self._xdone = torch.cuda.Event()
        self._xpending = False
        # device-resident protocol: the summary kernel writes here, the plan kernel's result lands in _plan_d
        self._info_d = torch.zeros(4, dtype=torch.int64, device="cuda")
        self._plan_d = torch.zeros(8, dtype=torch.int64, device="cuda")
        self._pstream = torch.cuda.Stream()
        self._placed_ev = torch.cuda.Event()
        self._async_pending = False
        self.codec.set_shard_out(self._info_d)

    def close(self):
        self.codec.close()

    def _ensure(self, nbytes: int, nrsi: int):
        torch = self.torch
        cap = (self._bound(self.p, nbytes) + 64 + 3) // 4 * 4
        if self.local is None or self.local.numel() < cap:
            self.local = torch.empty(cap, dtype=torch.uint8, device="cuda")
            self.placed = torch.empty(cap + 8, dtype=torch.uint8, device="cuda")
        if self.offsets is None or self.offsets.numel() < nrsi:
            self.offsets = torch.empty(max(nrsi, 1), dtype=torch.int64, device="cuda")
            self.grp = torch.zeros(max(nrsi, 1) * 32, dtype=torch.int64, device="cuda")

    def encode_local(self, d_raw, nbytes: int):
        """Step 1: code this rank's shard on its own (no communication)."""
        p = self.p
        R = p.rsi * p.block_size
        nrsi = (nbytes // p.bytes_per_sample + R - 1) // R
        self._ensure(nbytes, nrsi)
        self._raw, self._nbytes = d_raw, nbytes
        # also records the RSI and group indexes for the decoder
        self.codec.encode_enqueue(p, d_raw, nbytes, self.local, self.offsets, d_grp=self.grp)
        st, bits, kend = self.codec.encode_finish()
        assert st == 0
        self.bits = bits
        self._info = self.codec.shard_info()
        return bits

    def exchange_begin(self):
        """Step 2, first half: start the only exchange (32 bytes per rank) on the side stream.  The
        host does not wait here; work enqueued on the codec's stream meanwhile overlaps it."""
        import torch.distributed as dist
        if self.world <= 1:
            return
        torch = self.torch
        klo, khi, first_const, tail64 = self._info
        h = self._mine_h
        h[0], h[1], h[2] = self.bits, klo, khi
        h[3] = tail64 - (1 << 64) if tail64 >= (1 << 63) else tail64
        with torch.cuda.stream(self._xstream):
            self._mine.copy_(h, non_blocking=True)
            dist.all_gather_into_tensor(self._all, self._mine, group=self.group)
            self._all_h.copy_(self._all, non_blocking=True)
            self._xdone.record(self._xstream)
        self._xpending = True

    def stitch(self):
        """Steps 2-4: the 32-byte exchange (finished here), k repair, placement at the global bit
        phase.  Independent of decoding the local shard: callers call exchange_begin, enqueue the
        decode, then stitch, so that the exchange and the host's part of it hide behind the decode."""
        from .api import Carry
        p = self.p
        bits = self.bits
        klo, khi, first_const, tail64 = self._info
        if self.world > 1:
            if not self._xpending:
                self.exchange_begin()
            self._xdone.synchronize()                  # the side stream only, not the codec's stream
            self._xpending = False
            allv = self._all_h.view(self.world, 4).tolist()
            infos = [(v[0], v[1], v[2], v[3] & 0xFFFFFFFFFFFFFFFF) for v in allv]
        else:
            infos = [(bits, klo, khi, tail64)]
        plan = plan_shards(infos)[self.rank]
        # 3. k repair of the leading tiles
        if plan.k_in != 0 and self._nbytes:
            self.codec.set_tile_limit(first_const + 1)
            self.codec.encode_enqueue(p, self._raw, self._nbytes, self.local, None, Carry(0, plan.k_in, 0))
        # 4. move to the global bit phase (placed[0] is global word floor(bit_offset / 32)) and
        #    complete the first word with the predecessor's tail
        self.codec.place_bits(self.local, bits, self.placed, plan.bit_offset & 31, plan.head_or)
        self.plan = plan
        return plan

    def encode(self, d_raw, nbytes: int):
        """Steps 1-4 for this rank's shard `d_raw` (uint8 CUDA tensor).  Returns
        the ShardPlan; self.placed then holds this rank's words of the global
        stream starting at word plan.word_lo."""
        self.encode_local(d_raw, nbytes)
        plan = self.stitch()
        self.torch.cuda.current_stream().synchronize()
        return plan

    # ---- the same protocol with nothing but enqueues (no host round trip inside a step) ----
    def step_enqueue(self, d_raw, nbytes: int, gather_ptr: int | None = None, gather_cap: int = 0):
        """Steps 1-4 for this rank's shard, all on the device: returns at once.  The placement runs on a
        side stream (it only reads the shard, like the decode that may follow on the codec's stream).
        gather_ptr: base address of a buffer for the WHOLE stream (this rank's own memory or a peer's
        mapped over NVLink): the placement then writes the words this shard owns straight there."""
        import torch.distributed as dist
        torch = self.torch
        p = self.p
        R = p.rsi * p.block_size
        nrsi = (nbytes // p.bytes_per_sample + R - 1) // R
        self._ensure(nbytes, nrsi)
        self._raw, self._nbytes = d_raw, nbytes
        self._gather = (gather_ptr, gather_cap) if gather_ptr is not None else None
        self.bits = None                                # known on the device only until step_finish
        cur = torch.cuda.current_stream()
        if self._async_pending:
            cur.wait_event(self._placed_ev)            # the previous placement still reads self.local
        self.codec.encode_enqueue(p, d_raw, nbytes, self.local, self.offsets, d_grp=self.grp)
        if self.world > 1:
            dist.all_gather_into_tensor(self._all, self._info_d, group=self.group)
            src = self._all
        else:
            src = self._info_d
        self.codec.shard_plan(src, self.world, self.rank, self._plan_d)
        self.codec.encode_repair(p, d_raw, nbytes, self.local)
        ev = torch.cuda.Event()
        ev.record(cur)
        # the placement reads the repaired shard; run it next to whatever follows on the codec's stream
        self._pstream.wait_event(ev)
        self.codec.lib.aecb200_ctx_set_stream(self.codec.ctx, self._C.c_void_p(self._pstream.cuda_stream))
        try:
            if gather_ptr is not None:
                self.codec.place_planned(self.local, None, dst_ptr=gather_ptr, dst_cap=gather_cap,
                                         global_stream=True, last_rank=self.rank == self.world - 1)
            else:
                self.codec.place_planned(self.local, self.placed)
        finally:
            self.codec.lib.aecb200_ctx_set_stream(self.codec.ctx, self._C.c_void_p(cur.cuda_stream))
        self._placed_ev.record(self._pstream)
        self._async_pending = True

    def join(self):
        """Make the codec's stream wait for the placement of the last step_enqueue (no host wait)."""
        if self._async_pending:
            self.torch.cuda.current_stream().wait_event(self._placed_ev)

    def step_finish(self):
        """Wait for the last step_enqueue and read the plan back (host)."""
        self._placed_ev.synchronize()
        self.torch.cuda.current_stream().synchronize()
        v = self._plan_d.tolist()
        bits = int(v[5])
        off = int(v[2])
        total = int(v[4])
        end = off + bits
        last = self.rank == self.world - 1
        self.bits = bits
        self.plan = ShardPlan(off, int(v[0]), end, off >> 5, ((end + 31) >> 5) if last else (end >> 5), total,
                              int(v[3]) & 0xFFFFFFFF)
        self._async_pending = False
        if int(v[1]) > REPAIR_TILES and self._nbytes:
            # rare: k depends on the incoming k for more leading tiles than the device-side repair covers
            # (long runs of all-zero and plateau blocks): finish with the host-driven repair and place again
            from .api import Carry
            self.codec.set_tile_limit(int(v[1]))
            self.codec.encode_enqueue(self.p, self._raw, self._nbytes, self.local, None, Carry(0, self.plan.k_in, 0))
            if self._gather is not None:
                self.codec.place_planned(self.local, None, dst_ptr=self._gather[0], dst_cap=self._gather[1],
                                         global_stream=True, last_rank=last)
            else:
                self.codec.place_planned(self.local, self.placed)
            self.torch.cuda.current_stream().synchronize()
        return self.plan

    def owned_bytes(self):
        """This rank's bytes of the global stream (device tensor view)."""
        plan = self.plan
        nbytes = (plan.word_hi - plan.word_lo) * 4
        if self.rank == self.world - 1:
            nbytes = (plan.total_bits + 7) // 8 - plan.word_lo * 4
        return self.placed[:max(nbytes, 0)]

    def decode_enqueue(self, d_out, nbytes: int):
        """Decode this rank's shard from its own stream: no exchange needed."""
        p = self.p
        R = p.rsi * p.block_size
        nrsi = (nbytes // p.bytes_per_sample + R - 1) // R
        # the stream's length may still be known on the device only: the buffer's size bounds the reads then
        nb = (self.bits + 7) // 8 if self.bits is not None else self.local.numel() - 8
        return self.codec.decode_enqueue(p, self.local, nb, self.offsets, nrsi, d_out, nbytes, d_grp=self.grp)

    def decode(self, d_out, nbytes: int):
        self.decode_enqueue(d_out, nbytes)
        return self.codec.decode_finish()
