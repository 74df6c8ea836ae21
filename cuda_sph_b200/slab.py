"""x-slab domain decomposition for multi-GPU runs (one process per GPU, torch.distributed).

No reference counterpart -- the reference is single-GPU.  Interactions reach one cell
(voxel_size >= INF_R), so rank g owns the cell columns [X_g, X_{g+1}) and needs, every step,

  1. halo:      copies of the particles in the two columns on either side of its slab.  Two columns, not one,
                because the density of a first-column ghost is recomputed locally and that needs the ghost's own
                27 cells (saves the second exchange of the step, the one for rho);
  2. local step (sph_slab_step): hash / sort / in-cell order by GLOBAL id / density / forces on owned + ghosts,
                integrating owned particles only;
  3. migration: owned particles whose new column belongs to another rank move there (any rank: the reference's
                physics produces speeds of 1e3+ cells per step; the pipe outlet -> inlet recycle is a last -> first
                migration), with their xoroshiro state in PIPE mode.

Everything that crosses ranks is one variable-size all_to_all_single of packed byte rows (plus one for the counts).
The partition / packing / routing logic in `SlabRunner` is device-agnostic: tests/test_slab_gloo.py drives it on CPU
tensors over gloo with the fp64 oracle as the local step and demands bitwise equality with the single-domain oracle;
`GpuSlabRunner` binds the same logic to the device buffers of libsph_b200.so (zero copy) over NCCL.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

HALO = 2  # ghost columns per side


def equal_count_bounds(col_hist: np.ndarray, world: int, min_width: int = HALO) -> List[int]:
    """Column boundaries X_0 = 0 < X_1 < ... < X_world = W with (nearly) equal particle counts per slab
    (SURVEY.md section 7 'load balance').  Every slab is at least `min_width` columns wide so that a two-column halo
    only ever comes from the adjacent ranks."""
    w = int(len(col_hist))
    if w < world * min_width:
        raise ValueError(f"{w} cell columns cannot be split into {world} slabs of >= {min_width} columns")
    cum = np.concatenate([[0], np.cumsum(col_hist, dtype=np.int64)])
    total = int(cum[-1])
    bounds = [0]
    for g in range(1, world):
        target = total * g / world
        x = int(np.searchsorted(cum, target, side="left"))
        lo = bounds[-1] + min_width
        hi = w - (world - g) * min_width
        bounds.append(min(max(x, lo), hi))
    bounds.append(w)
    return bounds


def balanced_bounds(col_hist: np.ndarray, world: int, ghost_weight: float = 1.25, min_width: int = HALO) -> List[int]:
    """Column boundaries that balance the WORK of a step rather than the owned count: a rank hashes, sorts and stages its
    ghosts too (two columns per interior side), so the cost of slab [lo, hi) is
        sum(hist[lo:hi]) + ghost_weight * (hist[lo - HALO:lo] if lo > 0) + ghost_weight * (hist[hi:hi + HALO] if hi < W).
    Minimises the maximum cost over ranks (binary search on the bound, greedy sweep).  With whole-column slabs this puts
    the wider slabs at the domain ends, where there is only one halo.

    ghost_weight: measured, not guessed -- per-rank work of a box32m step (bench line `work_ms_per_rank`): 4 GPUs, slabs
    41|40|40|41 columns: 4.46 ms at the ends, 4.63 ms inside; 8 GPUs, 21|20x6|21: 2.47 / 2.64 ms.  Both give a ghost column
    the cost of 1.2-1.3 owned columns (it is hashed, sorted, staged, gets a density if it is the inner one, and is routed
    and unpacked on top)."""
    hist = np.asarray(col_hist, np.float64)
    w = len(hist)
    if w < world * min_width:
        raise ValueError(f"{w} cell columns cannot be split into {world} slabs of >= {min_width} columns")
    cum = np.concatenate([[0.0], np.cumsum(hist)])

    def cost(lo, hi):
        c = cum[hi] - cum[lo]
        if lo > 0:
            c += ghost_weight * (cum[lo] - cum[max(lo - HALO, 0)])
        if hi < w:
            c += ghost_weight * (cum[min(hi + HALO, w)] - cum[hi])
        return c

    def sweep(limit):
        bounds, lo = [0], 0
        for r in range(world):
            rest = world - 1 - r                      # slabs still to be placed after this one
            hi_max = w - rest * min_width
            if r == world - 1:
                hi = w
                if cost(lo, hi) > limit:
                    return None
            else:
                hi = lo + min_width
                if hi > hi_max or cost(lo, hi) > limit:
                    return None
                while hi < hi_max and cost(lo, hi + 1) <= limit:
                    hi += 1
            bounds.append(hi)
            lo = hi
        return bounds

    lo_t, hi_t = 0.0, float(cum[-1]) * (1.0 + 2.0 * ghost_weight) + 1.0
    best = sweep(hi_t)
    if best is None:
        return equal_count_bounds(col_hist, world, min_width)
    for _ in range(60):
        mid = 0.5 * (lo_t + hi_t)
        b = sweep(mid)
        if b is None:
            lo_t = mid
        else:
            best, hi_t = b, mid
    return [int(x) for x in best]


class SlabRunner:
    """Device-agnostic slab logic.  Sub-classes provide the storage tensors and the local step.

    Storage (rows = local particle slots, capacity `cap`):
        P [cap, 4]  x, y, z, density      V [cap, 4]  vx, vy, vz, 0      G [cap] int32 global particle id
        R [n_global, 2] int64 xoroshiro states indexed by global id (PIPE mode) or None
    Owned particles live in rows [0, n_own); ghosts of the current step in [n_own, n_local).
    """

    def __init__(self, n_cols: int, voxel_x: float, bounds: Sequence[int], group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_cols = int(n_cols)
        self.voxel_x = float(voxel_x)
        self.bounds = [int(b) for b in bounds]
        assert len(self.bounds) == self.world + 1 and self.bounds[0] == 0 and self.bounds[-1] == self.n_cols
        for g in range(self.world):
            assert self.bounds[g + 1] - self.bounds[g] >= HALO, "slab narrower than the halo"
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.n_own = 0
        self.n_local = 0
        self.P = self.V = self.G = self.R = None
        self.stats = {"halo_sent": 0, "migrated": 0, "steps": 0}

    # ------------------------------------------------------------------ to be provided by the backend
    def _local_step(self, n_own: int, n_local: int) -> None:
        raise NotImplementedError

    # ------------------------------------------------------------------ partition helpers
    def columns(self, x: torch.Tensor) -> torch.Tensor:
        """Cell column of a position: int32(x / voxel) with fp64 division and C truncation, as the hash kernel does
        (voxel_kernels.py:9-12).  Non-finite -> -1."""
        q = x.to(torch.float64) / self.voxel_x
        fin = torch.isfinite(q) & (q.abs() < 2147483648.0)
        col = torch.where(fin, q, torch.full_like(q, -1.0)).to(torch.int64)
        return torch.where(fin, col, torch.full_like(col, -1))

    def owner_of(self, col: torch.Tensor) -> torch.Tensor:
        """Rank owning a column (columns outside [0, W) are clamped: such particles are dead on their owner)."""
        b = torch.tensor(self.bounds[1:-1], dtype=torch.int64, device=col.device)
        return torch.bucketize(col.clamp(0, self.n_cols - 1), b, right=True)

    # ------------------------------------------------------------------ packed variable-size all-to-all
    def _exchange(self, rows: torch.Tensor, dest: torch.Tensor) -> torch.Tensor:
        """rows: [k, B] uint8, dest: [k] rank per row.  Returns the rows received (grouped by source rank)."""
        if self.world == 1:
            return rows[:0]
        order = torch.argsort(dest, stable=True)
        rows = rows[order].contiguous()
        send = torch.bincount(dest, minlength=self.world).to(torch.int64)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        send_l, recv_l = send.tolist(), recv.tolist()           # host sync: sizes of the payload exchange
        out = torch.empty((sum(recv_l), rows.shape[1]), dtype=torch.uint8, device=rows.device)
        dist.all_to_all_single(out, rows, output_split_sizes=recv_l, input_split_sizes=send_l, group=self.group)
        return out

    def _pack(self, idx: torch.Tensor, with_rng: bool) -> torch.Tensor:
        k, pb = len(idx), self.P.element_size() * 4          # explicit widths: k may be 0
        parts = [self.P[idx].contiguous().view(torch.uint8).reshape(k, pb),
                 self.V[idx].contiguous().view(torch.uint8).reshape(k, pb),
                 self.G[idx].contiguous().view(torch.uint8).reshape(k, 4)]
        if with_rng and self.R is not None:
            parts.append(self.R[self.G[idx].long()].contiguous().view(torch.uint8).reshape(k, 16))
        return torch.cat(parts, dim=1)

    def _unpack(self, rows: torch.Tensor, at: int, with_rng: bool) -> int:
        k = rows.shape[0]
        if k == 0:
            return 0
        if at + k > self.P.shape[0]:
            raise RuntimeError(f"slab capacity exceeded on rank {self.rank}: {at + k} > {self.P.shape[0]}")
        pb = self.P.element_size() * 4
        o = 0   # (.reshape(-1).clone(): a one-row slice keeps the wide row stride through .contiguous(), which view() rejects)
        self.P[at:at + k] = rows[:, o:o + pb].reshape(-1).clone().view(self.P.dtype).reshape(k, 4)
        o += pb
        self.V[at:at + k] = rows[:, o:o + pb].reshape(-1).clone().view(self.V.dtype).reshape(k, 4)
        o += pb
        gid = rows[:, o:o + 4].reshape(-1).clone().view(torch.int32).reshape(k)
        self.G[at:at + k] = gid
        o += 4
        if with_rng and self.R is not None:
            self.R[gid.long()] = rows[:, o:o + 16].reshape(-1).clone().view(torch.int64).reshape(k, 2)
        return k

    # ------------------------------------------------------------------ one step
    def exchange_halo(self) -> None:
        n = self.n_own
        col = self.columns(self.P[:n, 0])
        to_left = (col >= self.lo) & (col < self.lo + HALO) if self.rank > 0 else torch.zeros_like(col, dtype=torch.bool)
        to_right = (col >= self.hi - HALO) & (col < self.hi) if self.rank < self.world - 1 \
            else torch.zeros_like(col, dtype=torch.bool)
        il, ir = torch.nonzero(to_left).flatten(), torch.nonzero(to_right).flatten()
        idx = torch.cat([il, ir])
        dest = torch.cat([torch.full_like(il, self.rank - 1), torch.full_like(ir, self.rank + 1)])
        got = self._exchange(self._pack(idx, with_rng=False), dest)
        self.n_local = n + self._unpack(got, n, with_rng=False)
        self.stats["halo_sent"] += int(len(idx))

    def migrate(self) -> None:
        n = self.n_own
        col = self.columns(self.P[:n, 0])
        dest = torch.where(col >= 0, self.owner_of(col), torch.full_like(col, self.rank))
        leave = torch.nonzero(dest != self.rank).flatten()
        got = self._exchange(self._pack(leave, with_rng=True), dest[leave])
        if len(leave):
            keep = torch.nonzero(dest == self.rank).flatten()
            k = len(keep)
            self.P[:k] = self.P[keep]
            self.V[:k] = self.V[keep]
            self.G[:k] = self.G[keep]
            n = k
        self.n_own = n + self._unpack(got, n, with_rng=True)
        self.n_local = self.n_own
        self.stats["migrated"] += int(len(leave))

    def step(self, n_steps: int = 1) -> None:
        for _ in range(n_steps):
            self.exchange_halo()
            self._local_step(self.n_own, self.n_local)
            self.migrate()
            self.stats["steps"] += 1

    # ------------------------------------------------------------------ load balance
    def column_histogram(self) -> np.ndarray:
        """Owned particles per cell column, summed over the ranks."""
        col = self.columns(self.P[:self.n_own, 0]).clamp(0, self.n_cols - 1)
        hist = torch.bincount(col.cpu(), minlength=self.n_cols).to(torch.int64)
        if self.world > 1:
            dist.all_reduce(hist, group=self.group)
        return hist.numpy()

    def _set_bounds(self, bounds: Sequence[int]) -> None:
        self.bounds = [int(b) for b in bounds]
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]

    def rebalance(self) -> bool:
        """Between steps: re-evaluate the slab boundaries from the current distribution (balanced_bounds) and hand every
        particle to its new owner.  This protocol-level version moves the boundaries in place -- the particles whose
        column changed hands are ordinary migrants; NativeSlabRunner.rebalanced() rebuilds the fixed-capacity device
        layout instead.  Returns True if the boundaries moved."""
        new_bounds = [int(b) for b in balanced_bounds(self.column_histogram(), self.world)]
        if new_bounds == self.bounds:
            return False
        self._set_bounds(new_bounds)
        self._after_rebound()
        return True

    def _after_rebound(self) -> None:
        self.migrate()          # at rest every particle sits with its owner again

    # ------------------------------------------------------------------ loading / gathering
    def load_global(self, position: np.ndarray, velocity: np.ndarray) -> None:
        """Every rank passes the same full start state and keeps the particles of its slab (global id = row)."""
        pos = torch.as_tensor(np.ascontiguousarray(position))
        col = self.columns(pos[:, 0])
        mine = torch.nonzero(torch.where(col >= 0, self.owner_of(col), torch.zeros_like(col)) == self.rank).flatten()
        k = len(mine)
        if k > self.P.shape[0]:
            raise RuntimeError(f"slab capacity exceeded on rank {self.rank}: {k} > {self.P.shape[0]}")
        dev, dt = self.P.device, self.P.dtype
        self.P[:k, :3] = pos[mine].to(dt).to(dev)
        self.P[:k, 3] = 0
        self.V[:k, :3] = torch.as_tensor(np.ascontiguousarray(velocity))[mine].to(dt).to(dev)
        self.V[:k, 3] = 0
        self.G[:k] = mine.to(torch.int32).to(dev)
        self.n_own = self.n_local = k

    def gather_global(self, n_global: int):
        """All ranks -> every rank: (position, velocity, density) fp64 arrays in global-id order."""
        n = self.n_own
        rows = torch.cat([self.P[:n].to(torch.float64), self.V[:n, :3].to(torch.float64),
                          self.G[:n].to(torch.float64)[:, None]], dim=1).contiguous()
        if self.world > 1:
            counts = torch.tensor([n], dtype=torch.int64, device=rows.device)
            allc = [torch.empty_like(counts) for _ in range(self.world)]
            dist.all_gather(allc, counts, group=self.group)
            sizes = [int(c.item()) for c in allc]
            pad = torch.zeros((max(sizes), 8), dtype=torch.float64, device=rows.device)   # equal sizes for all_gather
            pad[:n] = rows
            bufs = [torch.empty_like(pad) for _ in allc]
            dist.all_gather(bufs, pad, group=self.group)
            rows = torch.cat([b[:k] for b, k in zip(bufs, sizes)])
        rows = rows.cpu().numpy()
        gid = rows[:, 7].astype(np.int64)
        assert len(gid) == n_global and len(np.unique(gid)) == n_global, "particles lost or duplicated"
        pos, vel, rho = np.empty((n_global, 3)), np.empty((n_global, 3)), np.empty(n_global)
        pos[gid], rho[gid], vel[gid] = rows[:, :3], rows[:, 3], rows[:, 4:7]
        return pos, vel, rho

    def count_global(self) -> int:
        c = torch.tensor([self.n_own], dtype=torch.int64, device=self.P.device)
        if self.world > 1:
            dist.all_reduce(c, group=self.group)
        return int(c.item())


class SingleExchangeSlabRunner(SlabRunner):
    """The protocol of the native path (csrc/slab_exchange.cuh) in torch ops: ghosts AND migrants of a step travel in one
    exchange, because the sender of a particle knows its final owner o and therefore also which neighbour of o needs it
    as a ghost (o - 1 if it sits in o's first two columns, o + 1 if in its last two).  Same routing rule, expression
    for expression, as slab_route_kernel; tests/test_slab_gloo.py runs it over gloo with the oracle as the local step."""

    def route_and_exchange(self) -> None:
        n = self.n_own
        dev = self.P.device
        col = self.columns(self.P[:n, 0])
        o = torch.where(col >= 0, self.owner_of(col), torch.full_like(col, self.rank))
        b = torch.tensor(self.bounds, dtype=torch.int64, device=dev)
        lo, hi = b[o], b[o + 1]
        leaves = o != self.rank
        ghost_l = (o > 0) & (col >= lo) & (col < lo + HALO)
        ghost_r = (o < self.world - 1) & (col >= hi - HALO) & (col < hi)
        im, il, ir = (torch.nonzero(m).flatten() for m in (leaves, ghost_l, ghost_r))
        idx = torch.cat([im, il, ir])
        dest = torch.cat([o[im], o[il] - 1, o[ir] + 1])
        kind = torch.cat([torch.zeros_like(im), torch.ones_like(il), torch.ones_like(ir)]).to(torch.uint8)
        rows = torch.cat([self._pack(idx, with_rng=True), kind[:, None]], dim=1)
        if self.world > 1:
            got = self._exchange(rows, dest)
        else:
            got = rows
        # the emigrants' slots are closed
        keep = torch.nonzero(~leaves).flatten()
        k = len(keep)
        if k != n:
            self.P[:k], self.V[:k], self.G[:k] = self.P[keep], self.V[keep], self.G[keep]
        is_ghost = got[:, -1] == 1
        mig, gho = got[~is_ghost][:, :-1], got[is_ghost][:, :-1]
        self.n_own = k + self._unpack(mig, k, with_rng=True)
        self.n_local = self.n_own + self._unpack(gho, self.n_own, with_rng=False)
        self.stats["halo_sent"] += int(len(il) + len(ir))
        self.stats["migrated"] += int(len(im))

    def step(self, n_steps: int = 1) -> None:
        for _ in range(n_steps):
            self.route_and_exchange()
            self._local_step(self.n_own, self.n_local)
            self.stats["steps"] += 1

    def _after_rebound(self) -> None:
        pass                    # the next step's routing sends everybody to the owner under the new boundaries


class _CudaBuffer:
    """Zero-copy view of a raw device pointer for torch.as_tensor (__cuda_array_interface__ v2)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class GpuSlabRunner(SlabRunner):
    """SlabRunner over the device buffers of one libsph_b200 handle in x-slab mode."""

    def __init__(self, params, constants=None, *, capacity: int, bounds: Sequence[int], device: int = 0,
                 cuda_stream: Optional[int] = None, group=None):
        from . import _lib
        from .strategy import SphConstants
        self._lib = _lib.load()
        self._chk = _lib.check
        cst = constants or SphConstants()
        space = np.asarray(params.space_size, np.float64).reshape(3)
        voxel = np.asarray(params.voxel_size, np.float64).reshape(3)
        n_cols = int(np.ceil(space[0] / voxel[0]))
        super().__init__(n_cols, float(voxel[0]), bounds, group)
        self.n_global = int(params.particle_count)
        p = _lib.SphParams()
        p.particle_count = int(capacity)
        p.mode = _lib.MODE_PIPE if cst.mode.upper() == "PIPE" else _lib.MODE_BOX
        p.h, p.mass, p.rho0, p.k, p.visc, p.damp = cst.h, cst.mass, cst.rho0, cst.k, cst.visc, cst.damp
        p.dt = 1 / params.fps
        ext = np.asarray(params.external_force, np.float64).reshape(3)
        for d in range(3):
            p.external_force[d], p.space_size[d], p.voxel_size[d] = ext[d], space[d], voxel[d]
        p.max_neighbours = cst.max_neighbours
        p.flags = _lib.FLAG_SLAB | _lib.FLAG_NO_GRAPH
        p.rng_seed = cst.rng_seed
        self._h = C.c_void_p()
        self._chk(self._lib.sph_create(C.byref(p), int(device), C.byref(self._h)))
        torch.cuda.set_device(device)
        if cuda_stream is None:   # the torch ops of the exchange and the engine kernels must share one stream
            cuda_stream = torch.cuda.current_stream().cuda_stream
        self._chk(self._lib.sph_set_stream(self._h, C.c_void_p(int(cuda_stream))))
        if p.mode == _lib.MODE_PIPE:
            table = np.ascontiguousarray(params.pipe.to_numpy(), dtype=np.float64)
            self._chk(self._lib.sph_set_pipe(self._h, table.ctypes.data, table.shape[0]))
        self._chk(self._lib.sph_slab_configure(self._h, self.lo, self.hi, self.n_global))
        dev = torch.device("cuda", device)
        cap = int(capacity)

        def view(which, shape, typestr, dtype):
            ptr, _ = C.c_void_p(), None
            cnt = C.c_int64()
            self._chk(self._lib.sph_device_ptr(self._h, which, C.byref(ptr), C.byref(cnt)))
            return torch.as_tensor(_CudaBuffer(ptr.value, shape, typestr), device=dev).view(dtype)

        # master arrays: one 32-byte record (position float4 | velocity float4) per slot -> two strided views of it
        M = view(0, (cap, 8), "<f4", torch.float32)
        self.P, self.V = M[:, 0:4], M[:, 4:8]
        self.G = view(4, (cap,), "<i4", torch.int32)
        self.R = view(5, (self.n_global, 2), "<i8", torch.int64) if p.mode == _lib.MODE_PIPE else None

    def _local_step(self, n_own: int, n_local: int) -> None:
        self._chk(self._lib.sph_slab_step(self._h, int(n_own), int(n_local)))

    def launch_count(self) -> int:
        return int(self._lib.sph_launch_count(self._h))

    def synchronize(self):
        self._chk(self._lib.sph_sync(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.P = self.V = self.G = self.R = None
            self._lib.sph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def plan_capacities(col_hist, bounds: Sequence[int], rank: int, n_global: int, pipe_mode: bool, *,
                    own_slack: float = 1.08, ghost_slack: float = 1.4, migrant_frac: float = 0.02,
                    far_frac: float = 0.005, migrant_floor: int = 4096) -> dict:
    """Slot and block capacities of rank `rank` for the native exchange, derived from the GLOBAL column histogram of the
    start state so that every rank computes the same block sizes: the block rank a sends to rank b has the size of the
    block b sends to a (an all_to_all with static split sizes needs that).

    cap_m[r] / cap_g[r]: migrant / ghost records in the block exchanged with rank r.  Adjacent ranks get
    ghost_slack x (their two boundary columns) ghosts and migrant_frac x (particles per rank) migrants; the others
    far_frac x (particles per rank) migrants for the rare long jump (in PIPE mode the last <-> first pair carries the
    outlet -> inlet recycle and gets a quarter of a rank); the block to oneself holds the ghosts a rank keeps of its own
    emigrants.  own_cap: owned region (own_slack x the fullest rank); capacity: owned + ghost region."""
    world = len(bounds) - 1
    hist = np.asarray(col_hist, np.int64)
    assert len(hist) == bounds[-1] and bounds[0] == 0
    own = [int(hist[bounds[r]:bounds[r + 1]].sum()) for r in range(world)]
    per_rank = max(max(own), n_global // world)
    m_adj = int(migrant_frac * per_rank) + int(migrant_floor)

    def band(r, side):   # particles rank r sends as ghosts to its left (0) / right (1) neighbour at the start
        lo, hi = bounds[r], bounds[r + 1]
        return int(hist[lo:lo + HALO].sum()) if side == 0 else int(hist[max(hi - HALO, lo):hi].sum())

    def caps(a, b):      # block a -> b, symmetric in (a, b)
        if a == b:       # ghosts for oneself: emigrants that stop inside the neighbour's boundary band (most do)
            return 0, 2 * m_adj
        a, b = min(a, b), max(a, b)
        wrap = pipe_mode and a == 0 and b == world - 1
        if b - a == 1:
            m, g = m_adj, int(ghost_slack * max(band(a, 1), band(b, 0))) + 4096
        else:
            m, g = int(far_frac * per_rank) + 2048, 1024
        if wrap:
            m = max(m, per_rank // 4)
        return m, g

    cap_m = np.asarray([caps(rank, r)[0] for r in range(world)], np.int32)
    cap_g = np.asarray([caps(rank, r)[1] for r in range(world)], np.int32)
    # the ghost REGION is sized for the expected halo (the blocks carry more slack: they are only wire volume)
    expected = sum(band(rank + d, 1 if d < 0 else 0) for d in (-1, 1) if 0 <= rank + d < world)
    ghost_cap = int(1.4 * expected) + 2 * m_adj + 8192
    own_cap = int(own_slack * per_rank) + 8192
    return {"cap_m": cap_m, "cap_g": cap_g, "own_cap": own_cap, "ghost_cap": ghost_cap,
            "capacity": own_cap + ghost_cap, "per_rank": per_rank}


class NativeSlabRunner:
    """x-slab runner on the native exchange path of libsph_b200.so (csrc/slab_exchange.cuh).

    Same decomposition as `SlabRunner` (two-column halos, ownership by cell column, migration to any rank, in-cell order
    by global id), but routing, packing and unpacking are CUDA kernels working on a fixed slot layout, and ghosts and
    migrants of one step travel together in ONE fixed-size all_to_all: nothing in the step loop synchronises with the
    host.  `SlabRunner` remains the executable specification of the protocol (tests/test_slab_gloo.py, CPU / gloo);
    tests/test_gpu_slab.py demands that this class reproduces the single-GPU engine bitwise.

    Slots [0, own_cap) hold owned particles (holes and unused slots have global id -1), [own_cap, capacity) ghosts.
    """

    def __init__(self, params, constants=None, *, col_hist: np.ndarray, bounds: Sequence[int], device: int = 0,
                 group=None, own_slack: float = 1.08, ghost_slack: float = 1.4, migrant_frac: float = 0.02,
                 far_frac: float = 0.005, compact_every: int = 4, poll_every: int = 16, p2p: bool | None = None,
                 migrant_floor: int = 4096):
        from . import _lib
        from .strategy import SphConstants
        self._lib = _lib.load()
        self._chk = _lib.check
        cst = constants or SphConstants()
        # what rebalanced() needs to build the successor of this runner
        self._ctor = dict(params=params, constants=constants, device=device, group=group, own_slack=own_slack,
                          ghost_slack=ghost_slack, migrant_frac=migrant_frac, far_frac=far_frac,
                          compact_every=compact_every, poll_every=poll_every, p2p=p2p, migrant_floor=migrant_floor)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        space = np.asarray(params.space_size, np.float64).reshape(3)
        voxel = np.asarray(params.voxel_size, np.float64).reshape(3)
        self.n_cols = int(np.ceil(space[0] / voxel[0]))
        self.voxel_x = float(voxel[0])
        self.bounds = [int(b) for b in bounds]
        assert len(self.bounds) == self.world + 1 and self.bounds[0] == 0 and self.bounds[-1] == self.n_cols
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.n_global = int(params.particle_count)
        # holes left by emigrants are closed every `compact_every` steps (device-side; inbound migrants take new slots
        # above the high-water mark in between, so this bounds how far it can creep towards own_cap)
        self.compact_every = int(compact_every)
        # the overflow word is reduced over the ranks and read back every `poll_every` steps WITHOUT blocking the step
        # loop: the copy issued at one poll is looked at when the next one comes round
        self.poll_every = int(poll_every)
        self._poll = None
        pipe_mode = cst.mode.upper() == "PIPE"

        plan = plan_capacities(col_hist, self.bounds, self.rank, self.n_global, pipe_mode, own_slack=own_slack,
                               ghost_slack=ghost_slack, migrant_frac=migrant_frac, far_frac=far_frac,
                               migrant_floor=migrant_floor)
        cap_m, cap_g = plan["cap_m"], plan["cap_g"]
        self.own_cap, self.capacity = plan["own_cap"], plan["capacity"]

        p = _lib.SphParams()
        p.particle_count = int(self.capacity)
        p.mode = _lib.MODE_PIPE if pipe_mode else _lib.MODE_BOX
        p.h, p.mass, p.rho0, p.k, p.visc, p.damp = cst.h, cst.mass, cst.rho0, cst.k, cst.visc, cst.damp
        p.dt = 1 / params.fps
        ext = np.asarray(params.external_force, np.float64).reshape(3)
        for d in range(3):
            p.external_force[d], p.space_size[d], p.voxel_size[d] = ext[d], space[d], voxel[d]
        p.max_neighbours = cst.max_neighbours
        p.flags = _lib.FLAG_SLAB | _lib.FLAG_NO_GRAPH
        p.rng_seed = cst.rng_seed
        self._h = C.c_void_p()
        self._chk(self._lib.sph_create(C.byref(p), int(device), C.byref(self._h)))
        torch.cuda.set_device(device)
        self._chk(self._lib.sph_set_stream(self._h, C.c_void_p(int(torch.cuda.current_stream().cuda_stream))))
        if pipe_mode:
            table = np.ascontiguousarray(params.pipe.to_numpy(), dtype=np.float64)
            self._chk(self._lib.sph_set_pipe(self._h, table.ctypes.data, table.shape[0]))
        self._chk(self._lib.sph_slab_configure(self._h, self.lo, self.hi, self.n_global))
        b = np.asarray(self.bounds, np.int32)
        sptr, rptr = C.c_void_p(), C.c_void_p()
        sizes = np.zeros(self.world, np.int64)
        self._chk(self._lib.sph_slab_exchange_init(self._h, self.world, self.rank, b.ctypes.data, self.own_cap,
                                                   cap_m.ctypes.data, cap_g.ctypes.data, C.byref(sptr), C.byref(rptr),
                                                   sizes.ctypes.data))
        self.block_bytes = [int(s) for s in sizes]
        dev = torch.device("cuda", device)
        total = int(sizes.sum())
        stride = (total + 255) & ~255          # the receive buffer is double-buffered by exchange parity (sph_b200.h)
        self.send = torch.as_tensor(_CudaBuffer(sptr.value, (total,), "|u1"), device=dev)
        self.recv2 = [torch.as_tensor(_CudaBuffer(rptr.value + q * stride, (total,), "|u1"), device=dev) for q in (0, 1)]
        # exchange over peer memory: every rank maps the receive allocations of all ranks (CUDA IPC) and pushes its
        # ghost / migrant records straight into them (route -> push -> flag barrier kernels): a step needs no all_to_all
        if p2p is None:
            p2p = os.environ.get("SPH_SLAB_P2P", "1") != "0"
        self.p2p = False
        if p2p:
            # every rank must end up on the same path: if CUDA IPC is unavailable on ANY rank (containers without a shared
            # IPC namespace, peer access disabled) all ranks take the NCCL all_to_all path, loudly
            err = None
            try:
                self._open_peers(col_hist, pipe_mode, dict(own_slack=own_slack, ghost_slack=ghost_slack,
                                                           migrant_frac=migrant_frac, far_frac=far_frac,
                                                           migrant_floor=migrant_floor), dev)
            except Exception as exc:   # noqa: BLE001 -- reported below, then agreed on by all ranks
                err = exc
            ok = torch.tensor([0 if err else 1], device=dev, dtype=torch.int32)
            if self.world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                self.p2p = False
                if err is not None or self.rank == 0:
                    print(f"[slab rank {self.rank}] peer-memory exchange unavailable ({err}); using the NCCL all_to_all "
                          "path on every rank", file=sys.stderr, flush=True)

        def view(which, shape, typestr, dtype):
            ptr, cnt = C.c_void_p(), C.c_int64()
            self._chk(self._lib.sph_device_ptr(self._h, which, C.byref(ptr), C.byref(cnt)))
            return torch.as_tensor(_CudaBuffer(ptr.value, shape, typestr), device=dev).view(dtype)

        cap = self.capacity
        # master arrays: one 32-byte record (position float4 | velocity float4) per slot -> two strided views of it
        M = view(0, (cap, 8), "<f4", torch.float32)
        self.P, self.V = M[:, 0:4], M[:, 4:8]
        self.G = view(4, (cap,), "<i4", torch.int32)
        self.counters = view(6, (8,), "<i4", torch.int32)
        self.R = view(5, (self.n_global, 2), "<i8", torch.int64) if pipe_mode else None   # xoroshiro128+ states by global id
        self.steps = 0

    def _open_peers(self, col_hist, pipe_mode, slack, dev) -> None:
        handle = np.zeros(64, np.uint8)
        self._chk(self._lib.sph_slab_ipc_handle(self._h, handle.ctypes.data))
        if self.world > 1:
            mine = torch.as_tensor(handle).to(dev)
            allh = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allh, mine, group=self.group)
            handles = np.ascontiguousarray(torch.stack(allh).cpu().numpy())
        else:
            handles = handle[None, :].copy()
        # offset of the block "from me" inside rank k's receive buffer: rank k lays it out by ITS block sizes
        remote = np.zeros(2 * self.world, np.int64)
        for k in range(self.world):
            pk = plan_capacities(col_hist, self.bounds, k, self.n_global, pipe_mode, **slack)
            rec = 48 if pipe_mode else 32
            sizes_k = 16 + (pk["cap_m"].astype(np.int64) + pk["cap_g"].astype(np.int64)) * rec
            stride_k = (int(sizes_k.sum()) + 255) & ~255       # rank k's second receive buffer starts here
            remote[2 * k] = int(sizes_k[:self.rank].sum())
            remote[2 * k + 1] = stride_k + remote[2 * k]
            assert int(sizes_k[self.rank]) == self.block_bytes[k], "block sizes must be symmetric"
        self._chk(self._lib.sph_slab_open_peers(self._h, handles.ctypes.data, remote.ctypes.data))
        self.p2p = True

    # ------------------------------------------------------------------ loading / stepping
    def _parity(self) -> int:
        q = C.c_int32()
        self._chk(self._lib.sph_slab_parity(self._h, C.byref(q)))
        return int(q.value)

    def exchange(self) -> None:
        """Deliver migrants and rebuild the halos from the owned regions as they are.  Peer memory: route -> push into the
        receivers' buffers -> flag barrier (three kernels, no collective); otherwise route -> one all_to_all.  Then unpack."""
        if self.p2p:
            self._chk(self._lib.sph_slab_exchange_p2p(self._h))
            self._chk(self._lib.sph_slab_unpack(self._h))
            return
        self._chk(self._lib.sph_slab_route(self._h))
        recv = self.recv2[self._parity()]
        if self.world > 1:
            dist.all_to_all_single(recv, self.send, output_split_sizes=self.block_bytes,
                                   input_split_sizes=self.block_bytes, group=self.group)
        else:
            recv.copy_(self.send)
        self._chk(self._lib.sph_slab_unpack(self._h))

    def load_global(self, position: np.ndarray, velocity: np.ndarray) -> None:
        """Every rank passes the same full start state and keeps the particles of its slab (global id = row)."""
        q = np.asarray(position[:, 0], np.float64) / self.voxel_x
        fin = np.isfinite(q) & (np.abs(q) < 2147483648.0)
        col = np.where(fin, q, -1.0).astype(np.int64)
        owner = np.searchsorted(np.asarray(self.bounds[1:-1]), np.clip(col, 0, self.n_cols - 1), side="right")
        owner = np.where(col >= 0, owner, 0)
        mine = np.nonzero(owner == self.rank)[0]
        k = len(mine)
        if k > self.own_cap:
            raise RuntimeError(f"slab capacity exceeded on rank {self.rank}: {k} > {self.own_cap}")
        dev = self.P.device
        self.G.fill_(-1)
        self.P.fill_(float("nan"))
        self.P[:k, :3] = torch.as_tensor(np.ascontiguousarray(position[mine], np.float32)).to(dev)
        self.P[:k, 3] = 0
        self.V[:k, :3] = torch.as_tensor(np.ascontiguousarray(velocity[mine], np.float32)).to(dev)
        self.V[:k, 3] = 0
        self.G[:k] = torch.as_tensor(mine.astype(np.int32)).to(dev)
        self.counters.zero_()
        self.counters[0] = k
        self.steps = 0
        self.exchange()

    # ------------------------------------------------------------------ load balance
    def column_histogram(self) -> np.ndarray:
        """Particles per cell column over ALL ranks, from the current owned positions."""
        own = torch.nonzero(self.G[:self.own_cap] >= 0).flatten()
        x = self.P[own, 0].to(torch.float64) / self.voxel_x
        col = torch.where(torch.isfinite(x), x, torch.zeros_like(x)).to(torch.int64).clamp_(0, self.n_cols - 1)
        hist = torch.bincount(col, minlength=self.n_cols).to(torch.int64)
        if self.world > 1:
            dist.all_reduce(hist, group=self.group)
        return hist.cpu().numpy()

    def rebalanced(self, min_gain: float = 0.03) -> "NativeSlabRunner":
        """Slab boundaries re-evaluated from the CURRENT particle distribution (a dam break spreads out; the boundaries of
        the start state then leave most of the work to one rank).  Returns this runner if the work-balanced boundaries
        (balanced_bounds) would lower the maximum per-rank cost by less than `min_gain`; otherwise a NEW runner that owns
        the same particles (same global ids, positions, velocities, xoroshiro states) under the new boundaries -- this one
        is closed.  Stop-the-world and host-mediated (gather by global id, reload): meant for every k >> 1 steps, off the
        step path; the particles' state is carried in fp32 exactly, so the run continues bit for bit."""
        hist = self.column_histogram()
        new_bounds = [int(b) for b in balanced_bounds(hist, self.world)]
        if new_bounds == self.bounds:
            return self

        def max_cost(bounds):
            cum = np.concatenate([[0], np.cumsum(hist)])
            cost = []
            for r in range(self.world):
                lo, hi = bounds[r], bounds[r + 1]
                c = cum[hi] - cum[lo]
                c += 1.25 * (cum[lo] - cum[max(lo - HALO, 0)]) + 1.25 * (cum[min(hi + HALO, self.n_cols)] - cum[hi])
                cost.append(c)
            return max(cost)
        if max_cost(new_bounds) > (1.0 - min_gain) * max_cost(self.bounds):
            return self
        pos, vel, _ = self.gather_global(self.n_global)
        rng = None
        if self.R is not None:   # the xoroshiro state of a particle lives with its owner: collect by global id
            own = torch.nonzero(self.G[:self.own_cap] >= 0).flatten()
            gids = self.G[own].to(torch.int64)
            rng = torch.zeros_like(self.R)
            rng[gids] = self.R[gids]
            if self.world > 1:
                dist.all_reduce(rng, group=self.group)
        kw = dict(self._ctor)
        params, constants = kw.pop("params"), kw.pop("constants")
        steps = self.steps
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)   # nobody is still pushing into a buffer that is about to be unmapped
        self.close()
        if self.world > 1:
            dist.barrier(group=self.group)
        new = NativeSlabRunner(params, constants, col_hist=hist, bounds=new_bounds, **kw)
        new.load_global(pos, vel)
        if rng is not None:
            new.R.copy_(rng)
        new.steps = steps
        return new

    def run(self, n_steps: int, rebalance_every: int = 0, min_gain: float = 0.03) -> "NativeSlabRunner":
        """n_steps steps with the slab boundaries re-evaluated every `rebalance_every` steps (0: never).  Returns the
        runner that holds the state afterwards (rebalanced() replaces the runner when the boundaries move)."""
        runner, done = self, 0
        while done < n_steps:
            g = min(rebalance_every or n_steps, n_steps - done)
            runner.step(g)
            done += g
            if rebalance_every and done < n_steps:
                runner = runner.rebalanced(min_gain)
        return runner

    def step(self, n_steps: int = 1) -> None:
        """Between steps the state is AT REST: every particle sits with its owner and the ghost region holds the halos of
        the current positions.  A step = local step -> exchange (peer memory or all_to_all) -> unpack."""
        for _ in range(n_steps):
            self._chk(self._lib.sph_slab_step_all(self._h))
            self.exchange()
            self.steps += 1
            if self.compact_every and self.steps % self.compact_every == 0:
                self._chk(self._lib.sph_slab_compact(self._h))
            if self.poll_every and self.steps % self.poll_every == 0:
                self._poll_overflow()

    def _poll_overflow(self) -> None:
        """Sticky overflow check off the critical path.  A lost particle cannot be recovered, so this raises (on every
        rank: the flag is max-reduced first) at the latest `poll_every` steps after the step that dropped it."""
        if self._poll is not None:
            host, ev = self._poll
            ev.synchronize()          # recorded poll_every steps ago: long done
            if int(host.item()):
                raise RuntimeError(f"x-slab exchange overflow (flags {int(host.item())}, rank {self.rank}): particles "
                                   "were dropped; raise own_slack / ghost_slack / migrant_frac / far_frac")
        flag = self.counters[2:3].to(torch.int64)
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        host = torch.empty(1, dtype=torch.int64, pin_memory=True)
        host.copy_(flag, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._poll = (host, ev)

    def step_timed(self) -> dict:
        """One step with CUDA events around every phase (synchronises; for the breakdown in bench.py, not for `value`)."""
        from . import _lib
        t = _lib.SphTimings()
        self._chk(self._lib.sph_slab_step_all_timed(self._h, C.byref(t)))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        ms3 = None
        if self.p2p:
            ev[1].record()
            ms3 = np.zeros(3, np.float32)
            self._chk(self._lib.sph_slab_exchange_p2p_timed(self._h, ms3.ctypes.data))
        else:
            self._chk(self._lib.sph_slab_route(self._h))
            ev[1].record()
            recv = self.recv2[self._parity()]
            if self.world > 1:
                dist.all_to_all_single(recv, self.send, output_split_sizes=self.block_bytes,
                                       input_split_sizes=self.block_bytes, group=self.group)
            else:
                recv.copy_(self.send)
        ev[2].record()
        self._chk(self._lib.sph_slab_unpack(self._h))
        ev[3].record()
        torch.cuda.synchronize()
        self.steps += 1
        # exchange_ms: the all_to_all, or (peer memory) the flag barrier = waiting for the slowest neighbour
        out = {"route_ms": ev[0].elapsed_time(ev[1]), "all_to_all_ms": ev[1].elapsed_time(ev[2]),
               "unpack_ms": ev[2].elapsed_time(ev[3])}
        if ms3 is not None:   # peer memory: no all_to_all; the exchange is three kernels
            out.update(route_ms=float(ms3[0]), all_to_all_ms=0.0, push_ms=float(ms3[1]), flag_barrier_ms=float(ms3[2]))
        out.update({k: getattr(t, k) for k in ("hash_ms", "sort_ms", "reorder_ms", "density_ms", "force_ms")})
        return out

    # ------------------------------------------------------------------ snapshots (bench windows)
    def snapshot(self):
        """Owned region + counters (+ the xoroshiro states in PIPE mode: they travel with the particles)."""
        return (self.P[:self.own_cap].clone(), self.V[:self.own_cap].clone(), self.G[:self.own_cap].clone(),
                self.counters.clone(), self.R.clone() if self.R is not None else None)

    def restore(self, snap) -> None:
        self.P[:self.own_cap], self.V[:self.own_cap], self.G[:self.own_cap] = snap[0], snap[1], snap[2]
        overflow = self.counters[2].clone()     # sticky: a restore must not hide an overflow that already happened
        self.counters.copy_(snap[3])
        self.counters[2] = torch.maximum(overflow, snap[3][2])
        if snap[4] is not None:
            self.R.copy_(snap[4])
        self.exchange()      # the ghost region belongs to another state: rebuild the halos (collective)

    # ------------------------------------------------------------------ inspection (these synchronise)
    def status(self) -> dict:
        out = np.zeros(5, np.int32)
        self._chk(self._lib.sph_slab_counters(self._h, out.ctypes.data))
        st = {"hwm": int(out[0]), "ghosts": int(out[1]), "overflow": int(out[2]), "live": int(out[3]),
              "own_cap": int(out[4]), "capacity": self.capacity}
        return st

    def check(self) -> dict:
        """Collective: raises on EVERY rank if any rank overflowed a send block (1), its owned region (2) or its ghost
        region (4) -- a rank must never leave the others waiting in a collective."""
        st = self.status()
        flags = torch.tensor([st["overflow"]], dtype=torch.int64, device=self.P.device)
        if self.world > 1:
            dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=self.group)
        if int(flags.item()):
            raise RuntimeError(f"x-slab exchange overflow (flags {int(flags.item())}; rank {self.rank}: {st}): "
                               "raise own_slack / ghost_slack / migrant_frac / far_frac")
        return st

    def count_global(self) -> int:
        c = torch.tensor([self.check()["live"]], dtype=torch.int64, device=self.P.device)
        if self.world > 1:
            dist.all_reduce(c, group=self.group)
        return int(c.item())

    def gather_global(self, n_global: int):
        """All ranks -> every rank: (position, velocity, density) fp64 arrays in global-id order."""
        self.check()
        own = torch.nonzero(self.G[:self.own_cap] >= 0).flatten()
        n = len(own)
        rows = torch.cat([self.P[own].to(torch.float64), self.V[own, :3].to(torch.float64),
                          self.G[own].to(torch.float64)[:, None]], dim=1).contiguous()
        if self.world > 1:
            counts = torch.tensor([n], dtype=torch.int64, device=rows.device)
            allc = [torch.empty_like(counts) for _ in range(self.world)]
            dist.all_gather(allc, counts, group=self.group)
            sizes = [int(c.item()) for c in allc]
            pad = torch.zeros((max(sizes), 8), dtype=torch.float64, device=rows.device)
            pad[:n] = rows
            bufs = [torch.empty_like(pad) for _ in allc]
            dist.all_gather(bufs, pad, group=self.group)
            rows = torch.cat([b_[:k] for b_, k in zip(bufs, sizes)])
        rows = rows.cpu().numpy()
        gid = rows[:, 7].astype(np.int64)
        assert len(gid) == n_global and len(np.unique(gid)) == n_global, "particles lost or duplicated"
        pos, vel, rho = np.empty((n_global, 3)), np.empty((n_global, 3)), np.empty(n_global)
        pos[gid], rho[gid], vel[gid] = rows[:, :3], rows[:, 3], rows[:, 4:7]
        return pos, vel, rho

    def launch_count(self) -> int:
        return int(self._lib.sph_launch_count(self._h))

    def synchronize(self):
        self._chk(self._lib.sph_sync(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.P = self.V = self.G = self.R = self.counters = self.send = self.recv2 = None
            self._lib.sph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
