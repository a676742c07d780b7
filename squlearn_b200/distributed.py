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


# Sub-block pipelining of the residue exchange (see ShardedFidelityKernel._rows_planes) is OFF by default: measured
# on 4 and 8 GPUs (profiles/r2_timeline_n{4,8}_variants.log) the per-copy overhead of the many small pushes and the
# less efficient small simulation batches cost more (8 GPUs: 6.3 - 6.7 ms) than the earlier start saves (5.7 ms with
# one push per block).
_DEFAULT_SUBBLOCKS = 1
_DEFAULT_PUSH_STREAMS = "1"
# Order of the exchanged blocks: "ring" (nearest owner first) or "narrow" (narrowest block first); a second push stream
# (B200SQ_DIST_PUSH_STREAMS=2) sends alternate blocks concurrently.  Measured on 4 / 8 GPUs
# (profiles/r2_timeline_n{4,8}_order.log): ring + one stream 8.9 / 5.75 ms, narrow 10.9 / 7.4 ms, two streams 10.7 / 7.3 ms.
_DEFAULT_ORDER = "ring"
_DEFAULT_PIPELINE = "full"


def _mark(name, stream=None):  # phase hook for tools/dist_breakdown.py / dist_timeline.py (no-op in production)
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


def ring_distance(rank: int, owner: int, world: int) -> int:
    """Steps from ``rank`` forward to ``owner`` on the ring of ranks."""
    return (owner - rank) % world


def remote_jobs(jobs: List[List[dict]], rank: int, world: int) -> List[dict]:
    """The off-diagonal jobs of ``rank`` in the order their column planes arrive: nearest owner first (every rank
    pushes its planes to rank-1, rank-2, ... in that order, so the block a rank needs first lands first; see
    ``_job_order`` for the measured alternative)."""
    mine = [j for j in jobs[rank] if not j["diag"]]
    mine.sort(key=lambda j: _job_order(j, rank, world))
    return mine


def _job_order(j: dict, rank: int, world: int):
    if os.environ.get("B200SQ_DIST_ORDER", _DEFAULT_ORDER).lower() == "ring":   # nearest owner first (A/B measurements)
        return (0, ring_distance(rank, j["owner"], world), j["rows"])
    return (j["cols"][1] - j["cols"][0], ring_distance(rank, j["owner"], world), j["rows"])


def arena_layout(jobs_of_rank: List[dict], rank: int, world: int, dim: int, n_local: int = 0, n: int = 0,
                 align: int = 1024, n_planes: int = 6):
    """Receive arena of a rank: one slot [n_planes nat planes | row exponents] per off-diagonal job (6 digit planes,
    or the s residue planes when residues travel), in arrival order,
    then the arrival flags, the "mirrored blocks stored" flags and the rank's row block K[r0:r1, :] itself (peers
    store the transposes of their blocks straight into it).  Returns ([{"owner", "cols", "rows", "off",
    "expo_off"}], {"flags", "done", "rows", "total"} byte offsets); every rank computes the layout of every other
    rank from the job table alone (a sender needs the receiver's offsets)."""
    slots, off = [], 0
    mine = [j for j in jobs_of_rank if not j["diag"]]
    mine.sort(key=lambda j: _job_order(j, rank, world))
    for j in mine:
        w = j["cols"][1] - j["cols"][0]
        nb = n_planes * w * 2 * dim
        slots.append({"owner": j["owner"], "cols": j["cols"], "rows": w, "off": off, "expo_off": off + nb})
        off = (off + nb + 4 * w + align - 1) // align * align
    flags_off = off
    done_off = flags_off + (4 * world + 127) // 128 * 128
    rows_off = (done_off + 4 * world + align - 1) // align * align
    total = rows_off + 8 * n_local * n + align
    return slots, {"flags": flags_off, "done": done_off, "rows": rows_off, "total": total}


def push_schedule(pushes: List[dict], bounds: List[Tuple[int, int]], fill: int = 1) -> List[List[tuple]]:
    """Order of the copy-engine pushes when the local rows become available sub-block by sub-block.

    ``pushes`` (priority order: nearest consumer first) each cover local rows [a, b); ``bounds`` are the sub-blocks in
    the order they are sliced.  A chunk = (push index, lo, hi, last): rows [lo, hi) of that push, ``last`` when it
    completes the push (its exponents and arrival flag follow).  Stage s lists the chunks issued right after
    sub-block s has been sliced: the ``1 + fill`` most urgent chunks whose rows exist by then, everything that is left
    after the last sub-block.  Every row of every push is issued exactly once, a push's chunks in row order."""
    chunks = []   # (priority, sub, push, lo, hi)
    for pi, p in enumerate(pushes):
        for sub, (a, b) in enumerate(bounds):
            lo, hi = max(a, p["a"]), min(b, p["b"])
            if hi > lo:
                chunks.append((pi, sub, lo, hi))
    left = {pi: sum(1 for c in chunks if c[0] == pi) for pi in range(len(pushes))}
    done, stages = set(), []
    for s in range(len(bounds)):
        ready = [c for c in chunks if c[1] <= s and c not in done]   # sorted by (push priority, sub-block)
        take = ready if s == len(bounds) - 1 else ready[:1 + fill]
        stage = []
        for c in take:
            done.add(c)
            left[c[0]] -= 1
            stage.append((c[0], c[2], c[3], left[c[0]] == 0))
        stages.append(stage)
    return stages


class ShardedFidelityKernel:
    """Fidelity Gram matrix of ``x`` across all ranks of a process group.

    ``kernel`` is a :class:`squlearn_b200.kernel.FidelityKernel` (or any object with
    ``states(x) -> device tensor`` and ``_executor.engine.gram``); ``group`` the process group
    (default: the world).  Works with NCCL (GPU tensors) and, for the CPU tests of the sharding
    logic, with gloo and a stand-in engine.

    INT8 (digit-plane) path, exchange modes (``B200SQ_DIST_EXCHANGE``):

    * ``ipc`` (default on CUDA): every rank PUSHES the nat planes of its row block (CRT arithmetic: the s residue
      planes; digit form: the 6 digit planes) into the receive arenas of the ranks that need them with copy-engine
      copies over NVLink (CUDA IPC mappings; no SM is spent on the
      exchange) and then writes an epoch flag; the receiver's single job-table Gram launch multiplies each
      block as soon as its flag is visible (``b200sq_gram_jobs``), so the exchange hides behind the diagonal
      block and the earlier blocks.
    * ``p2p``: the same blocks as one NCCL send/recv batch, waited for before the (single) Gram launch.
    * ``allgather``: one NCCL all-gather of the equal-sized blocks (what the north star names).
    """

    def __init__(self, kernel, group=None, diagonal: str = "one", precision: Optional[str] = None):
        import torch.distributed as dist

        self._dist = dist
        self._blocks = {}
        self._peer = None          # state of the ipc exchange (arena, peer mappings, epoch)
        self._host = None          # shared pinned host buffers of evaluate()
        self.precision = precision  # Gram arithmetic: None = the engine's default policy ("auto")
        self._gram_kw = {} if precision in (None, "auto") else {"precision": precision}
        self.kernel = kernel
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.diagonal = diagonal
        self._borrow_rows = False   # evaluate() reads the shared row block in place (it copies it to the host)

    # ------------------------------------------------------------------ digit-plane path
    def _exchange_mode(self, eng, equal: bool, n_local: int, block_bytes: int) -> str:
        mode = os.environ.get("B200SQ_DIST_EXCHANGE", "auto").lower()
        if mode == "allgather" and not (equal and n_local > 0 and n_local % 4 == 0
                                        and self.world * block_bytes <= (32 << 30)):
            mode = "auto"   # the all-gather needs equal, 16-byte aligned blocks that fit 32 GiB
        if mode in ("auto", "ipc"):
            can_ipc = hasattr(eng, "ipc_export") and self._dist.get_backend(self.group) == "nccl"
            mode = "ipc" if can_ipc else "p2p"
        if mode not in ("ipc", "p2p", "allgather"):
            raise ValueError("B200SQ_DIST_EXCHANGE must be auto, ipc, p2p or allgather")
        return mode

    def _peer_setup(self, eng, parts, jobs, dim, n_planes=6):
        """Arena, flags, the shared row block and the peer mappings of the ipc exchange (cached while the shapes
        stay the same).  ``n_planes``: nat planes per exchanged block (6 digit planes or the s residue planes)."""
        import torch

        dist = self._dist
        key = (tuple(parts), dim, n_planes, os.environ.get("B200SQ_DIST_ORDER", _DEFAULT_ORDER).lower())
        if self._peer is not None and self._peer["key"] == key:
            return self._peer
        self._peer_close()
        n = parts[-1][1]
        lay = lambda q: arena_layout(jobs[q], q, self.world, dim, parts[q][1] - parts[q][0], n, n_planes=n_planes)
        slots, offs = lay(self.rank)
        arena = eng.alloc_bytes(offs["total"])
        arena[offs["flags"]:offs["rows"]].zero_()
        handle, off = eng.ipc_export(arena)
        infos = [None] * self.world
        dist.all_gather_object(infos, (handle, off), group=self.group)
        r0, r1 = parts[self.rank]
        bases, layouts = {}, {}

        def base_of(q):
            if q not in bases:
                bases[q] = eng.ipc_open(infos[q][0])
                layouts[q] = lay(q)
            return bases[q] + infos[q][1]

        pushes = []    # my planes -> the arenas of the ranks whose jobs read them, nearest consumer first
        writers = []   # ranks that store mirrored blocks into my rows (they signal my "done" flags)
        for q in range(self.world):
            if q == self.rank:
                continue
            q_slots, q_offs = lay(q)
            for sl in q_slots:
                if sl["owner"] != self.rank:
                    continue
                base = base_of(q)
                a, b = sl["cols"]
                pushes.append({"dist": ring_distance(q, self.rank, self.world), "q": q, "a": a - r0, "b": b - r0,
                               "dst": base + sl["off"], "dst_expo": base + sl["expo_off"],
                               "dst_flag": base + q_offs["flags"] + 4 * self.rank})
                if q not in writers:
                    writers.append(q)
        narrow = os.environ.get("B200SQ_DIST_ORDER", _DEFAULT_ORDER).lower() != "ring"
        pushes.sort(key=lambda p: (p["b"] - p["a"] if narrow else 0, p["dist"], p["q"]))   # the consumers' order (_job_order)
        owners = {}    # owner of the columns of one of my jobs -> (its rows pointer, its "done" flag for me)
        for sl in slots:
            q = sl["owner"]
            base = base_of(q)
            owners[q] = (base + layouts[q][1]["rows"], base + layouts[q][1]["done"] + 4 * self.rank, parts[q][0])
        torch.cuda.synchronize(eng.device)
        dist.barrier(group=self.group)   # every arena is mapped and zeroed before the first push
        n_local = r1 - r0
        rows = arena[offs["rows"]:offs["rows"] + 8 * n_local * n].view(torch.float64).view(n_local, n)
        self._peer = {"key": key, "arena": arena, "slots": slots, "offs": offs, "pushes": pushes, "writers": writers,
                      "owners": owners, "rows": rows, "bases": bases, "epoch": 0,
                      "stream": torch.cuda.Stream(device=eng.device), "stream2": torch.cuda.Stream(device=eng.device),
                      "sliced": torch.cuda.Event(), "pushed": None,
                      "eng": eng}
        return self._peer

    def _peer_close(self):
        if self._peer is None:
            return
        try:
            import torch

            torch.cuda.synchronize(self._peer["eng"].device)
            for base in self._peer["bases"].values():
                self._peer["eng"].ipc_close(base)
        except Exception:   # interpreter / process-group shutdown
            pass
        self._peer = None

    def __del__(self):
        self._peer_close()
        self._host_close()

    def _rows_planes(self, eng, x, parts, jobs, dim):
        """The digit-plane (INT8) path: slice once, exchange the nat planes, ONE job-table Gram launch.
        Returns (rows, jobs left for the NCCL mirror exchange or None when the blocks were mirrored by peer stores)."""
        import torch

        dist = self._dist
        r0, r1 = parts[self.rank]
        n_local = r1 - r0
        mine = remote_jobs(jobs, self.rank, self.world)
        equal = all(b - a == n_local for a, b in parts)
        nat_bytes = 6 * n_local * 2 * dim
        mode = self._exchange_mode(eng, equal, n_local, nat_bytes + 4 * n_local)

        # Arithmetic of the blocks: digit slices ("int8": the 6 nat digit planes travel) or residues ("crt": fewer
        # tensor-core products; every rank turns its own states into residue planes directly).
        scheme = eng.gram_precision(parts[-1][1], parts[-1][1], dim, self.precision)
        crt = scheme == "crt" and hasattr(eng, "digits_to_crt")
        # CRT + ipc: by default the s nat RESIDUE planes travel instead of the 6 digit planes.  That is s / 6 times the
        # bytes but no conversion at the receiver: measured on 2, 4 and 8 GPUs (profiles/r2_timeline_*), a conversion
        # that runs next to the persistent Gram kernel is starved by it (3 - 10 times its stand-alone time) and was the
        # critical path of every step, while the pushes ride on otherwise idle copy engines
        # (B200SQ_DIST_CRT_EXCHANGE = auto | digits | residues; auto = residues).
        xch = os.environ.get("B200SQ_DIST_CRT_EXCHANGE", "auto").lower()
        if xch not in ("auto", "digits", "residues"):
            raise ValueError("B200SQ_DIST_CRT_EXCHANGE must be auto, digits or residues")
        residues = crt and mode == "ipc" and xch != "digits"
        n_xch = eng.plane_count(dim, "crt") if residues else 6
        if crt:   # residue planes for the local rows (+ with the digit exchange their 6 nat digit planes, from ONE read)
            _, a_planes, a_expo = self._plane_block(eng, "local_crt", n_local, dim, "crt")
            if residues:
                buf_local, planes, expo = None, a_planes, a_expo
            else:
                buf_local, planes, expo = self._nat_block(eng, "local_nat", n_local, dim)
        else:
            buf_local, planes, expo = self._plane_block(eng, "local", n_local, dim)
            a_planes, a_expo = planes, expo      # row operand of the Gram jobs
        if mode == "ipc":
            peer = self._peer_setup(eng, parts, jobs, dim, n_xch)
            for ev in peer["pushed"] or ():   # my previous pushes read the local planes
                torch.cuda.current_stream(eng.device).wait_event(ev)
            rows = peer["rows"]              # peers store their mirrored blocks into this (shared) row block
        else:
            rows = torch.zeros((n_local, parts[-1][1]), dtype=torch.float64, device=eng.device
                               if hasattr(eng, "device") else "cpu")
        # Optional (B200SQ_DIST_SUBBLOCKS > 1): the local rows are simulated, sliced and pushed in SUB-BLOCKS, so that
        # the copy engines start while the later sub-blocks are still being computed (see _DEFAULT_SUBBLOCKS).
        n_sub, sub_sim = 1, True
        if residues:
            env_sub = os.environ.get("B200SQ_DIST_SUBBLOCKS", "")
            n_sub = int(env_sub) if env_sub else _DEFAULT_SUBBLOCKS
            n_sub = max(1, min(n_sub, n_local // 64 if n_local >= 64 else 1))
            # "full": simulate, slice and push sub-block by sub-block; "slice": simulate all local rows in one batch
            # (small batches lose efficiency in the gate passes), slice and push in sub-blocks
            sub_sim = os.environ.get("B200SQ_DIST_PIPELINE", _DEFAULT_PIPELINE).lower() != "slice"
        if mode == "ipc":
            peer["epoch"] += 1
            epoch = peer["epoch"]
            arena, st = peer["arena"], peer["stream"]
            epoch_t = torch.full((1,), epoch, dtype=torch.int32, device=eng.device)
            kb = 2 * dim
        if n_sub > 1:
            bounds = [ab for ab in row_partition(n_local, n_sub) if ab[1] > ab[0]]
            stages = push_schedule(peer["pushes"], bounds, fill=1)
            main = torch.cuda.current_stream(eng.device)
            if not sub_sim:
                psi_local = self.kernel.states(x[r0:r1]).contiguous()
                stats = eng.row_stats_of(psi_local)
            for (a, b), stage in zip(bounds, stages):
                if sub_sim:
                    psi_sub = self.kernel.states(x[r0 + a:r0 + b]).contiguous()
                    eng.slice_states(psi_sub, a_planes, a_expo, a, "crt")
                    del psi_sub
                else:
                    eng.slice_states(psi_local[a:b], a_planes, a_expo, a, "crt",
                                     stats=stats[a:b] if stats is not None else None)
                ev = torch.cuda.Event()
                ev.record(main)
                st.wait_event(ev)
                for pi, lo, hi, last in stage:
                    p = peer["pushes"][pi]
                    w = p["b"] - p["a"]
                    eng.copy_planes_async(p["dst"] + (lo - p["a"]) * kb, w * kb, planes[0, lo].data_ptr(), n_local * kb,
                                          (hi - lo) * kb, n_xch, st)
                    if last:
                        eng.copy_async(p["dst_expo"], expo[p["a"]:].data_ptr(), 4 * w, st)
                        eng.copy_async(p["dst_flag"], epoch_t.data_ptr(), 4, st)
            psi_local = None
            _mark("states")
            _mark("sliced")
        else:
            psi_local = self.kernel.states(x[r0:r1]).contiguous()
            _mark("states")
            if crt:
                eng.slice_states(psi_local, a_planes, a_expo, 0, "crt", digits=None if residues else planes)
                expo = a_expo
            else:
                eng.slice_states(psi_local, planes, expo, 0)
            del psi_local  # only the planes are needed from here on (16 GiB per rank at n = 20)
            _mark("sliced")

        if mode == "ipc":
            if n_sub == 1:
                peer["sliced"].record(torch.cuda.current_stream(eng.device))
                st.wait_event(peer["sliced"])
                # B200SQ_DIST_PUSH_STREAMS=2: alternate blocks go out on a second stream (two copy engines at a time)
                two = os.environ.get("B200SQ_DIST_PUSH_STREAMS", _DEFAULT_PUSH_STREAMS) == "2"
                if two:
                    peer["stream2"].wait_event(peer["sliced"])
                for pk, p in enumerate(peer["pushes"]):     # the consumers' order
                    st = peer["stream2"] if (two and pk % 2 == 1) else peer["stream"]
                    w = p["b"] - p["a"]
                    if w == n_local:         # the whole block: its nat planes are contiguous
                        eng.copy_async(p["dst"], planes.data_ptr(), n_xch * n_local * kb, st)
                    else:                    # a row range (antipodal split): one copy per plane
                        eng.copy_planes_async(p["dst"], w * kb, planes[0, p["a"]].data_ptr(), n_local * kb, w * kb, n_xch, st)
                    eng.copy_async(p["dst_expo"], expo[p["a"]:].data_ptr(), 4 * w, st)
                    eng.copy_async(p["dst_flag"], epoch_t.data_ptr(), 4, st)
            peer["pushed"] = []
            for st in (peer["stream"], peer["stream2"]):
                ev = torch.cuda.Event()
                ev.record(st)
                peer["pushed"].append(ev)
                epoch_t.record_stream(st)
            _mark("pushed", peer["stream"])
            _mark("pushed2", peer["stream2"])
            flags_ptr = arena.data_ptr() + peer["offs"]["flags"]
            n = rows.shape[1]
            store_mirror = os.environ.get("B200SQ_DIST_MIRROR", "store").lower() != "nccl"
            for j, sl in zip(mine, peer["slots"]):
                w = sl["rows"]
                j["planes_y"], j["expo_y"] = arena.data_ptr() + sl["off"], arena.data_ptr() + sl["expo_off"]
                j["rows_y"], j["y0"], j["ready"] = w, 0, (flags_ptr + 4 * j["owner"], epoch)
                # the transposed block goes straight into the owner's row block (peer stores from the epilogue);
                # B200SQ_DIST_MIRROR=nccl keeps the NCCL send/recv mirror exchange (A/B measurements)
                if store_mirror:
                    o_rows, _, c0 = peer["owners"][j["owner"]]
                    j["mirror"] = (o_rows + 8 * ((j["cols"][0] - c0) * n + j["rows"][0]), n)
            _mark("diag")
            _mark("exchange")
        elif mode == "allgather":
            # one NCCL all-gather of the [6 nat planes | exponents] blocks; a job reads its owner's block in place
            block = nat_bytes + 4 * n_local
            gathered = self._bytes(eng, "gathered", self.world * block)
            send = self._bytes(eng, "send", block)
            send[:nat_bytes].copy_(buf_local[:nat_bytes])
            send[nat_bytes:].copy_(expo.contiguous().view(torch.int8))
            dist.all_gather_into_tensor(gathered, send, group=self.group)
            _mark("diag")
            _mark("exchange")
            for j in mine:
                q, c0 = j["owner"], parts[j["owner"]][0]
                j["planes_y"] = gathered[q * block: q * block + nat_bytes].view(6, n_local, 2 * dim)
                j["expo_y"] = gathered[q * block + nat_bytes: (q + 1) * block].view(torch.int32)
                j["rows_y"], j["y0"], j["ready"] = n_local, j["cols"][0] - c0, None
        else:
            # one NCCL batch: [6 nat planes | exponents] of exactly the rows every consumer needs
            ops, keep = [], []
            for q in range(self.world):
                if q == self.rank:
                    continue
                for j in jobs[q]:
                    if j["owner"] == self.rank and not j["diag"]:
                        a, b = j["cols"]
                        msg, mp, me = self._nat_block(eng, ("stage", q), b - a, dim)
                        mp.copy_(planes[:6, a - r0:b - r0])
                        me.copy_(expo[a - r0:b - r0])
                        keep.append(msg)
                        ops.append(dist.P2POp(dist.isend, msg, self._global_rank(q), self.group))
            for j in mine:
                w = j["cols"][1] - j["cols"][0]
                buf, j["planes_y"], j["expo_y"] = self._nat_block(eng, ("recv", j["owner"]), w, dim)
                j["rows_y"], j["y0"], j["ready"] = w, 0, None
                ops.append(dist.P2POp(dist.irecv, buf, self._global_rank(j["owner"]), self.group))
            reqs = dist.batch_isend_irecv(ops) if ops else []
            _mark("diag")
            for req in reqs:
                req.wait()
            _mark("exchange")
        glist = []
        if n_local:
            glist.append({"x0": 0, "nx": n_local, "planes_y": a_planes, "expo_y": a_expo, "rows_y": n_local, "y0": 0,
                          "ny": n_local, "out": rows[:, r0:r1], "symmetric": True, "diagonal": self.diagonal})
        if crt and mine and not residues:
            self._convert_blocks(eng, mine, dim, mode, peer if mode == "ipc" else None,
                                 epoch_t if mode == "ipc" else None)
        for j in mine:
            ra, rb = j["rows"]
            w = j["cols"][1] - j["cols"][0]
            j["blk"] = rows[ra - r0:rb - r0, j["cols"][0]:j["cols"][1]]
            glist.append({"x0": ra - r0, "nx": rb - ra, "planes_y": j["planes_y"], "expo_y": j["expo_y"],
                          "rows_y": j["rows_y"], "y0": j["y0"], "ny": w, "out": j["blk"], "ready": j["ready"],
                          "mirror": j.get("mirror")})
        if glist:
            eng.gram_jobs(a_planes, a_expo, glist)
        _mark("gram")
        if mode == "ipc" and store_mirror:
            # tell every owner that its mirrored blocks are stored; wait for the ranks that store into my rows
            for q in sorted({j["owner"] for j in mine}):
                eng.copy_async(peer["owners"][q][1], epoch_t.data_ptr(), 4)
            if peer["writers"]:
                eng.wait_flags(arena.data_ptr() + peer["offs"]["done"], peer["writers"], epoch)
            return rows, None       # nothing left for the NCCL mirror exchange
        return rows, mine

    def _convert_blocks(self, eng, mine, dim, mode, peer, epoch_t):
        """CRT arithmetic: residue planes of every received digit block (b200sq_gram_digits_to_crt).  With the ipc
        exchange the conversions run on their own stream, each behind a wait for its block's arrival flag and
        followed by a flag of its own that the block's Gram job waits for — so a block is converted while earlier
        blocks are multiplied and later ones are still in flight.  Otherwise (NCCL exchanges: all blocks are there)
        they simply run in stream order."""
        import torch

        n_mod = eng.plane_count(dim, "crt")
        ptr = lambda t: int(t) if isinstance(t, int) else t.data_ptr()
        if mode == "ipc":
            if "conv_flags" not in peer:
                peer["conv_flags"] = torch.zeros(self.world, dtype=torch.int32, device=eng.device)
                peer["conv_stream"] = torch.cuda.Stream(device=eng.device, priority=-1)   # ahead of the Gram's CTAs
            cs, cflags = peer["conv_stream"], peer["conv_flags"]
            cs.wait_stream(torch.cuda.current_stream(eng.device))   # the epoch scalar, and last step's Gram is behind us
            flags_ptr = peer["arena"].data_ptr() + peer["offs"]["flags"]
        for j in mine:
            w = j["cols"][1] - j["cols"][0]
            conv = self._bytes(eng, ("conv", j["owner"], j["cols"]), n_mod * w * 2 * dim)
            expo_ptr = ptr(j["expo_y"]) + 4 * j["y0"]
            if mode == "ipc":
                with torch.cuda.stream(cs):
                    eng.wait_flags(flags_ptr, [j["owner"]], j["ready"][1])
                    eng.digits_to_crt(j["planes_y"], j["rows_y"], j["y0"], w, dim, conv, w, 0)
                    eng.copy_async(cflags.data_ptr() + 4 * j["owner"], epoch_t.data_ptr(), 4, cs)
                j["ready"] = (cflags.data_ptr() + 4 * j["owner"], j["ready"][1])
                _mark("converted", cs)
            else:
                eng.digits_to_crt(j["planes_y"], j["rows_y"], j["y0"], w, dim, conv, w, 0)
            j["planes_y"], j["expo_y"], j["rows_y"], j["y0"] = conv, expo_ptr, w, 0
        if mode == "ipc":
            epoch_t.record_stream(cs)

    def _rows_states(self, eng, x, parts, jobs, rows, dim):
        """The FP64 path: exchange only the statevector rows the jobs need (NCCL point-to-point over NVLink:
        every peer is one NVSwitch hop away, so there is no ring-order penalty).  The receive buffer holds the
        column states of all my jobs back to back, so that jobs sharing a row range become ONE rectangular
        Gram launch."""
        import torch

        dist = self._dist
        f64 = torch.float64
        r0, r1 = parts[self.rank]
        n_local = r1 - r0
        psi_local = self.kernel.states(x[r0:r1]).contiguous()
        _mark("states")
        dev = psi_local.device
        mine = [j for j in jobs[self.rank] if not j["diag"]]
        mine.sort(key=lambda j: (j["rows"], j["owner"]))
        total_cols = sum(j["cols"][1] - j["cols"][0] for j in mine)
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
        # ---- one launch per distinct row range
        i = 0
        while i < len(mine):
            k = i
            while k < len(mine) and mine[k]["rows"] == mine[i]["rows"]:
                k += 1
            ra, rb = mine[i]["rows"]
            c_lo, c_hi = mine[i]["off"], mine[k - 1]["off"] + (mine[k - 1]["cols"][1] - mine[k - 1]["cols"][0])
            big = eng.gram(psi_local[ra - r0:rb - r0], recv[c_lo:c_hi], **self._gram_kw)
            for j in mine[i:k]:
                w = j["cols"][1] - j["cols"][0]
                blk = big[:, j["off"] - c_lo: j["off"] - c_lo + w]
                rows[ra - r0:rb - r0, j["cols"][0]:j["cols"][1]] = blk
                j["blk"] = blk
            i = k
        return mine

    def evaluate_rows(self, x):
        """This rank's row block K[r0:r1, :] of the symmetric Gram matrix (device tensor)."""
        import torch

        dist = self._dist
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = x.shape[0]
        parts = row_partition(n, self.world)
        r0, r1 = parts[self.rank]
        eng = self.kernel._executor.engine
        if hasattr(self.kernel, "num_qubits"):
            dim = 1 << int(self.kernel.num_qubits)
        else:   # stand-in kernels: ask an (empty) batch for its width
            dim = int(self.kernel.states(x[:0]).shape[1])
        jobs = gram_jobs(parts)
        n_local = r1 - r0
        dev = getattr(eng, "device", "cpu")
        use_planes = (hasattr(eng, "gram_jobs")      # same decision on every rank
                      and eng.gram_precision(n, n, dim, self.precision) in ("int8", "crt"))
        if use_planes:
            rows, mine = self._rows_planes(eng, x, parts, jobs, dim)
        else:
            rows = torch.zeros((n_local, n), dtype=torch.float64, device=dev)
            mine = self._rows_states(eng, x, parts, jobs, rows, dim)
        _mark("rect")
        if mine is None:   # ipc exchange: the row block is complete (and shared with the peers: hand out a copy)
            _mark("mirror")
            return (rows if self._borrow_rows else rows.clone()), (r0, r1)
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
                    buf = torch.empty((b - a, rb - ra), dtype=torch.float64, device=dev)
                    placements.append((a - r0, b - r0, ra, rb, buf))
                    ops.append(dist.P2POp(dist.irecv, buf, self._global_rank(q), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for a, b, ra, rb, buf in placements:
            rows[a:b, ra:rb] = buf
        _mark("mirror")
        return rows, (r0, r1)

    def _bytes(self, eng, tag, nbytes: int):
        buf = self._blocks.get(tag)
        if buf is None or buf.numel() != nbytes:
            self._blocks.pop(tag, None)
            buf = self._blocks[tag] = eng.alloc_bytes(nbytes)
        return buf

    def _nat_block(self, eng, tag, rows: int, dim: int):
        """[6 nat planes | row exponents] of ``rows`` states in one contiguous buffer (one message)."""
        import torch

        nb = 6 * rows * 2 * dim
        buf = self._bytes(eng, tag, nb + 4 * rows)
        return buf, buf[:nb].view(6, rows, 2 * dim), buf[nb:].view(torch.int32)

    def _plane_block(self, eng, tag, rows: int, dim: int, scheme: str = "int8"):
        """Plane buffers are kept between calls (tens of GiB per rank at n = 20: re-allocating them every
        evaluation fragments the caching allocator at the edge of the 180 GB)."""
        key = (tag, rows, dim)
        blk = self._blocks.get(key)
        if blk is None:
            for k in [k for k in self._blocks if isinstance(k, tuple) and k[0] == tag]:   # shape changed: drop the old one
                del self._blocks[k]
            blk = self._blocks[key] = eng.alloc_plane_block(rows, dim, scheme) if scheme != "int8" \
                else eng.alloc_plane_block(rows, dim)
        return blk

    def _global_rank(self, group_rank: int) -> int:
        if self.group is None:
            return group_rank
        return self._dist.get_global_rank(self.group, group_rank)

    # ------------------------------------------------------------------ host result
    def _host_setup(self, n: int, dev):
        """Two (n, n) float64 buffers in shared memory (/dev/shm), page-locked in every rank, so that each rank
        copies its row block device -> host over its own PCIe link instead of funnelling K through one GPU."""
        import mmap

        import torch

        dist = self._dist
        if self._host is not None and self._host["n"] == n:
            return self._host
        self._host_close()
        nbytes = max(n * n * 8, 8)
        ok = False
        if getattr(dev, "type", "cpu") == "cuda" and os.environ.get("B200SQ_DIST_HOST", "shm") == "shm":
            try:
                st = os.statvfs("/dev/shm")
                ok = st.f_bavail * st.f_frsize > 2 * nbytes + (64 << 20)
            except OSError:
                ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self._host = {"n": n, "maps": None}
            return self._host
        paths = [None, None]
        if self.rank == 0:
            paths = [f"/dev/shm/b200sq_{os.getpid()}_{id(self):x}_{k}" for k in range(2)]
            for path in paths:
                fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                os.ftruncate(fd, nbytes)
                os.close(fd)
        dist.broadcast_object_list(paths, src=self._global_rank(0), group=self.group)
        maps, views, regs = [], [], []
        rt = torch.cuda.cudart()
        for path in paths:
            fd = os.open(path, os.O_RDWR)
            mm = mmap.mmap(fd, nbytes)
            os.close(fd)
            t = torch.frombuffer(mm, dtype=torch.float64, count=n * n).view(n, n)
            err = rt.cudaHostRegister(t.data_ptr(), nbytes, 0)
            regs.append(t.data_ptr() if int(err) == 0 else None)
            maps.append(mm)
            views.append(t)
        dist.barrier(group=self.group)
        if self.rank == 0:   # every rank holds its mapping: the names can go (nothing to leak on a crash)
            for path in paths:
                os.unlink(path)
        self._host = {"n": n, "maps": maps, "views": views, "regs": regs, "turn": 0}
        return self._host

    def _host_close(self):
        h, self._host = self._host, None
        if not h or not h.get("maps"):
            return
        try:
            import torch

            rt = torch.cuda.cudart()
            for p in h["regs"]:
                if p is not None:
                    rt.cudaHostUnregister(p)
        except Exception:   # interpreter shutdown
            pass
        h["views"] = None   # the mappings themselves go when the last array referencing them does

    def evaluate(self, x, dst: int = 0) -> Optional[np.ndarray]:
        """Full (N, N) Gram matrix as a NumPy array on rank ``dst`` (None elsewhere).

        On CUDA every rank copies its row block straight into a shared page-locked host buffer (two buffers are
        used in turn: the array returned by a call stays valid until the call after the next one)."""
        import torch

        dist = self._dist
        self._borrow_rows = True
        try:
            rows, (r0, r1) = self.evaluate_rows(x)
        finally:
            self._borrow_rows = False
        n = rows.shape[1]
        parts = row_partition(n, self.world)
        host = self._host_setup(n, rows.device) if dst == 0 else {"maps": None}
        if host.get("maps"):
            view = host["views"][host["turn"]]
            host["turn"] ^= 1
            if r1 > r0:
                view[r0:r1].copy_(rows, non_blocking=True)
            torch.cuda.current_stream(rows.device).synchronize()
            dist.barrier(group=self.group)
            return view.numpy() if self.rank == dst else None
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
