"""Multi-GPU fidelity Gram: one process per GPU (torch.distributed), sharded by row blocks.

The reference has no multi-device path at all; this module is the data-parallel
extension the north star asks for, built on the same single-GPU entry points:

  1. rank r simulates its slice of the data points            (no communication)
  2. each rank receives the statevector blocks its share of the Gram needs (NCCL point-to-point
     over NVLink: floor(P/2) of the P-1 remote blocks), overlapped with its diagonal block
  3. the P x P grid of Gram blocks is covered using Hermitian symmetry: rank r computes the
     diagonal block (r, r) with the symmetric kernel and, in one launch, the rectangular blocks
     (r, (r+s) mod P),
     s = 1 .. floor(P/2)  (for even P the s = P/2 blocks are split between the two owners), i.e.
     P(P+1)/2 blocks in total, balanced to within one block
  4. every off-diagonal block is transposed and sent to the rank that owns the mirrored position,
     so each rank ends up with its complete row block K[r0:r1, :].

Single states are never split across GPUs (a 20-qubit state is 16 MiB).
"""

import os
from typing import List, Optional, Tuple

import numpy as np


def _mark(name):  # phase hook for tools/dist_breakdown.py (no-op in production)
    pass


def row_partition(n: int, parts: int) -> List[Tuple[int, int]]:
    """Balanced contiguous row ranges [(start, stop)] of n rows over `parts` ranks."""
    base, rem = divmod(n, parts)
    out, start = [], 0
    for r in range(parts):
        size = base + (1 if r < rem else 0)
        out.append((start, start + size))
        start += size
    return out


def symmetric_block_schedule(parts: int) -> List[List[int]]:
    """For each rank r the column blocks c it computes so that every unordered pair {r, c} is
    computed exactly once and the load is balanced (a half ring)."""
    sched = [[r] for r in range(parts)]
    for r in range(parts):
        for s in range(1, parts // 2 + 1):
            c = (r + s) % parts
            if parts % 2 == 0 and s == parts // 2 and r >= parts // 2:
                continue  # the antipodal pair is computed by the lower rank only
            sched[r].append(c)
    return sched


def gram_jobs(parts: List[Tuple[int, int]]) -> List[List[dict]]:
    """Per rank, the rectangular pieces of the symmetric Gram matrix it computes.

    Every job is ``{"rows": (a, b), "cols": (a, b), "owner": rank owning the column states,
    "diag": bool}`` in global indices; ``rows`` is a sub-range of the rank's own row block.
    Regular half-ring blocks use the full row block; for even P the antipodal block (r, r + P/2)
    is split between its two owners — the lower rank takes the first half of its rows against the
    whole remote block, the upper rank takes its full row block against the second half of the
    lower rank's states — so every rank carries exactly P/2 block units."""
    world = len(parts)
    sched = symmetric_block_schedule(world)
    jobs: List[List[dict]] = [[] for _ in range(world)]
    for r in range(world):
        r0, r1 = parts[r]
        if r1 > r0:
            jobs[r].append({"rows": (r0, r1), "cols": (r0, r1), "owner": r, "diag": True})
        for c in sched[r]:
            if c == r:
                continue
            c0, c1 = parts[c]
            if c1 == c0 or r1 == r0:
                continue
            antipodal = world % 2 == 0 and (c - r) % world == world // 2
            if not antipodal:
                jobs[r].append({"rows": (r0, r1), "cols": (c0, c1), "owner": c, "diag": False})
            else:  # r is the lower rank of the pair (the schedule lists the pair only there)
                mid = r0 + (r1 - r0 + 1) // 2
                jobs[r].append({"rows": (r0, mid), "cols": (c0, c1), "owner": c, "diag": False})
                if r1 > mid:
                    jobs[c].append({"rows": (c0, c1), "cols": (mid, r1), "owner": r, "diag": False})
    return jobs


class ShardedFidelityKernel:
    """Fidelity Gram matrix of ``x`` across all ranks of a process group.

    ``kernel`` is a :class:`squlearn_b200.kernel.FidelityKernel` (or any object with
    ``states(x) -> device tensor`` and ``_executor.engine.gram``); ``group`` the process group
    (default: the world).  Works with NCCL (GPU tensors) and, for the CPU tests of the sharding
    logic, with gloo and a stand-in engine.
    """

    def __init__(self, kernel, group=None, diagonal: str = "one", precision: Optional[str] = None):
        import torch.distributed as dist

        self._dist = dist
        self._blocks = {}
        self.precision = precision  # Gram arithmetic: None = the engine's default policy ("auto")
        self._gram_kw = {} if precision in (None, "auto") else {"precision": precision}
        self.kernel = kernel
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.diagonal = diagonal

    def evaluate_rows(self, x):
        """This rank's row block K[r0:r1, :] of the symmetric Gram matrix (device tensor)."""
        import torch

        dist = self._dist
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = x.shape[0]
        parts = row_partition(n, self.world)
        r0, r1 = parts[self.rank]
        eng = self.kernel._executor.engine
        psi_local = self.kernel.states(x[r0:r1]).contiguous()
        _mark("states")
        dim = psi_local.shape[1]
        dev = psi_local.device
        f64 = torch.float64

        jobs = gram_jobs(parts)
        mine = [j for j in jobs[self.rank] if not j["diag"]]
        # every off-diagonal piece of mine: column states come from the owner of those columns
        mine.sort(key=lambda j: (j["rows"], j["owner"]))
        total_cols = sum(j["cols"][1] - j["cols"][0] for j in mine)
        n_local = r1 - r0
        rows = torch.zeros((n_local, n), dtype=f64, device=dev)
        use_planes = (hasattr(eng, "gram_from_planes")
                      and eng.gram_precision(n, n, dim, self.precision) == "int8")  # same decision on every rank

        if use_planes:
            # ---- split-integer path: every rank slices its own states ONCE into int8 digit planes and the
            # planes travel (12 instead of 16 bytes per amplitude, no re-slicing on the receiver).  A block of
            # rows is one contiguous buffer [12 planes | row exponents], so the exchange is a single NCCL
            # message per peer, all posted in one batch and overlapped with the diagonal block.
            equal = all(b - a == n_local for a, b in parts)
            block_bytes = 12 * n_local * 2 * dim + 4 * n_local
            mode = os.environ.get("B200SQ_DIST_EXCHANGE", "auto")   # "allgather" | "p2p" | "auto"
            # default: point-to-point (measured faster: 15.4 vs 16.7 ms per C3 step on 4 GPUs; it moves only the
            # floor(P/2) blocks a rank needs); the all-gather is kept as an option
            use_allgather = mode == "allgather" and equal and n_local > 0 \
                and n_local % 4 == 0 and self.world * block_bytes <= (32 << 30)   # 16-byte aligned blocks, <= 32 GiB
            diag_after = os.environ.get("B200SQ_DIST_DIAG_AFTER", "0") == "1"
            if use_allgather:
                # ---- one NCCL all-gather of the plane blocks (NVSwitch collective bandwidth; a rank uses
                # floor(P/2) of the P - 1 remote blocks it receives, which is still faster than P2P send/recv:
                # measured 250 GB/s per rank for the point-to-point batch).  Block q of `gathered` is rank q's
                # [12 planes | row exponents] buffer; my own block is sliced in place.
                gathered = self._gather_buffer(eng, self.world * block_bytes)
                nb = 12 * n_local * 2 * dim
                view = lambda q: (gathered[q * block_bytes: q * block_bytes + nb].view(12, n_local, 2 * dim),
                                  gathered[q * block_bytes + nb: (q + 1) * block_bytes].view(torch.int32))
                planes, expo = view(self.rank)
                eng.slice_states(psi_local, planes, expo, 0)
                del psi_local
                work = dist.all_gather_into_tensor(gathered, gathered[self.rank * block_bytes:(self.rank + 1) * block_bytes],
                                                   group=self.group, async_op=True)
                if not diag_after:
                    rows[:, r0:r1] = eng.gram_from_planes(planes, expo, 0, n_local, diagonal=self.diagonal)
                _mark("diag")
                work.wait()
                _mark("exchange")
                if diag_after:
                    rows[:, r0:r1] = eng.gram_from_planes(planes, expo, 0, n_local, diagonal=self.diagonal)
                for j in mine:
                    ra, rb = j["rows"]
                    (a, b), c0 = j["cols"], parts[j["owner"]][0]
                    pq, eq = view(j["owner"])
                    blk = eng.gram_from_planes(planes, expo, ra - r0, rb - ra, pq, eq, a - c0, b - a)
                    rows[ra - r0:rb - r0, a:b] = blk
                    j["blk"] = blk
            else:
                buf_local, planes, expo = self._plane_block(eng, "local", n_local, dim)
                eng.slice_states(psi_local, planes, expo, 0)
                ops, keep = [], []
                for q in range(self.world):      # sends: rows of mine that other ranks' jobs need
                    if q == self.rank:
                        continue
                    for j in jobs[q]:
                        if j["owner"] == self.rank and not j["diag"]:
                            a, b = j["cols"]
                            if (a, b) == (r0, r1):
                                msg = buf_local
                            else:  # a sub-range of my rows (antipodal split): stage it contiguously
                                msg, mp, me = self._plane_block(eng, ("stage", q), b - a, dim)
                                mp.copy_(planes[:, a - r0:b - r0])
                                me.copy_(expo[a - r0:b - r0])
                                keep.append(msg)
                            ops.append(dist.P2POp(dist.isend, msg, self._global_rank(q), self.group))
                for j in mine:
                    w = j["cols"][1] - j["cols"][0]
                    j["buf"], j["planes"], j["expo"] = self._plane_block(eng, ("recv", j["owner"]), w, dim)
                    ops.append(dist.P2POp(dist.irecv, j["buf"], self._global_rank(j["owner"]), self.group))
                reqs = dist.batch_isend_irecv(ops) if ops else []
                del psi_local  # only the planes are needed from here on (16 GiB per rank at n = 20)
                # The diagonal block needs no remote data.  By default it runs while the planes are in flight;
                # B200SQ_DIST_DIAG_AFTER=1 runs it after the exchange.
                if n_local and not diag_after:
                    rows[:, r0:r1] = eng.gram_from_planes(planes, expo, 0, n_local, diagonal=self.diagonal)
                _mark("diag")
                for req in reqs:
                    req.wait()
                _mark("exchange")
                if n_local and diag_after:
                    rows[:, r0:r1] = eng.gram_from_planes(planes, expo, 0, n_local, diagonal=self.diagonal)
                for j in mine:
                    ra, rb = j["rows"]
                    w = j["cols"][1] - j["cols"][0]
                    blk = eng.gram_from_planes(planes, expo, ra - r0, rb - ra, j["planes"], j["expo"], 0, w)
                    rows[ra - r0:rb - r0, j["cols"][0]:j["cols"][1]] = blk
                    j["blk"] = blk
            mine_done = True
        else:
            # ---- exchange only the statevector rows the jobs need (NCCL point-to-point over NVLink:
            # every peer is one NVSwitch hop away, so there is no ring-order penalty).  The receive
            # buffer holds the column states of all my jobs back to back, so that jobs sharing a row
            # range become ONE rectangular Gram launch.
            recv = torch.empty((max(total_cols, 1), dim), dtype=psi_local.dtype, device=dev)
            ops, off = [], 0
            for q in range(self.world):          # sends: rows of mine that other ranks' jobs need
                if q == self.rank:
                    continue
                for j in jobs[q]:
                    if j["owner"] == self.rank and not j["diag"]:
                        a, b = j["cols"]
                        ops.append(dist.P2POp(dist.isend, psi_local[a - r0:b - r0].view(f64),
                                              self._global_rank(q), self.group))
            for j in mine:
                w = j["cols"][1] - j["cols"][0]
                j["off"] = off
                ops.append(dist.P2POp(dist.irecv, recv[off:off + w].view(f64),
                                      self._global_rank(j["owner"]), self.group))
                off += w
            reqs = dist.batch_isend_irecv(ops) if ops else []

            # ---- the diagonal block needs no remote data: it runs while the blocks are in flight
            if n_local:
                rows[:, r0:r1] = eng.gram(psi_local, None, diagonal=self.diagonal, **self._gram_kw)
            _mark("diag")
            for req in reqs:
                req.wait()
            _mark("exchange")

            def rect(ra, rb, c_lo, c_hi):
                return eng.gram(psi_local[ra - r0:rb - r0], recv[c_lo:c_hi], **self._gram_kw)
            mine_done = False

        # ---- one launch per distinct row range
        i = len(mine) if mine_done else 0
        while i < len(mine):
            k = i
            while k < len(mine) and mine[k]["rows"] == mine[i]["rows"]:
                k += 1
            ra, rb = mine[i]["rows"]
            c_lo, c_hi = mine[i]["off"], mine[k - 1]["off"] + (mine[k - 1]["cols"][1] - mine[k - 1]["cols"][0])
            big = rect(ra, rb, c_lo, c_hi)
            for j in mine[i:k]:
                w = j["cols"][1] - j["cols"][0]
                blk = big[:, j["off"] - c_lo: j["off"] - c_lo + w]
                rows[ra - r0:rb - r0, j["cols"][0]:j["cols"][1]] = blk
                j["blk"] = blk
            i = k

        _mark("rect")
        # ---- mirror: every off-diagonal piece, transposed, belongs to the owner of its columns
        ops, placements = [], []
        for j in mine:
            ops.append(dist.P2POp(dist.isend, j["blk"].t().contiguous(), self._global_rank(j["owner"]),
                                  self.group))
        for q in range(self.world):
            if q == self.rank:
                continue
            for j in jobs[q]:
                if j["owner"] == self.rank and not j["diag"]:
                    (a, b), (ra, rb) = j["cols"], j["rows"]
                    buf = torch.empty((b - a, rb - ra), dtype=f64, device=dev)
                    placements.append((a - r0, b - r0, ra, rb, buf))
                    ops.append(dist.P2POp(dist.irecv, buf, self._global_rank(q), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for a, b, ra, rb, buf in placements:
            rows[a:b, ra:rb] = buf
        _mark("mirror")
        return rows, (r0, r1)

    def _gather_buffer(self, eng, nbytes: int):
        buf = self._blocks.get("gathered")
        if buf is None or buf.numel() != nbytes:
            self._blocks.pop("gathered", None)
            buf = self._blocks["gathered"] = eng.alloc_plane_block(0, 1)[0].new_empty((nbytes,))
        return buf

    def _plane_block(self, eng, tag, rows: int, dim: int):
        """Plane buffers are kept between calls (tens of GiB per rank at n = 20: re-allocating them every
        evaluation fragments the caching allocator at the edge of the 180 GB)."""
        key = (tag, rows, dim)
        blk = self._blocks.get(key)
        if blk is None:
            for k in [k for k in self._blocks if isinstance(k, tuple) and k[0] == tag]:   # shape changed: drop the old one
                del self._blocks[k]
            blk = self._blocks[key] = eng.alloc_plane_block(rows, dim)
        return blk

    def _global_rank(self, group_rank: int) -> int:
        if self.group is None:
            return group_rank
        return self._dist.get_global_rank(self.group, group_rank)

    def evaluate(self, x, dst: int = 0) -> Optional[np.ndarray]:
        """Full (N, N) Gram matrix as a NumPy array on rank ``dst`` (None elsewhere)."""
        import torch

        dist = self._dist
        rows, (r0, r1) = self.evaluate_rows(x)
        n = rows.shape[1]
        parts = row_partition(n, self.world)
        if self.rank == dst:
            full = torch.empty((n, n), dtype=torch.float64, device=rows.device)
            full[r0:r1] = rows
            ops = []
            for c, (c0, c1) in enumerate(parts):
                if c != dst and c1 > c0:
                    ops.append(dist.P2POp(dist.irecv, full[c0:c1], self._global_rank(c), self.group))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            eng = self.kernel._executor.engine
            return eng.to_host(full) if hasattr(eng, "to_host") else full.cpu().numpy()
        if r1 > r0:
            for req in dist.batch_isend_irecv(
                    [dist.P2POp(dist.isend, rows, self._global_rank(dst), self.group)]):
                req.wait()
        return None


def sharded_parameter_shift(circuit, executor, param, x, param_obs, group=None) -> np.ndarray:
    """d<O>/d param with the shifted-variant batch sharded over the ranks of ``group``; every rank
    keeps the full batch of samples and returns the complete (N, n_obs, P) gradient
    (sum all-reduced).  Shards by shifted-parameter batch as the north star asks."""
    import torch
    import torch.distributed as dist

    from .executor import b200_gradient

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_var = len(circuit.program.shift_variants(np.atleast_2d(np.asarray(x, dtype=np.float64)))[0])
    a, b = row_partition(n_var, world)[rank]
    part = b200_gradient(circuit, executor, param=param, x=x, param_obs=param_obs,
                         variant_slice=slice(a, b))
    dev = getattr(executor.engine, "device", "cpu")
    t = torch.from_numpy(part).to(dev)
    dist.all_reduce(t, group=group)
    return t.cpu().numpy()
