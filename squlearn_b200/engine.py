"""Device-side engine: thin Python wrapper over the C ABI, with PyTorch used only for device
memory, streams and host<->device copies.

All arithmetic happens in libb200sq.so's CUDA kernels.  There is no CPU path: constructing an
:class:`Engine` without a CUDA device (or without the built library) raises.
"""

import ctypes
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .observables import pauli_masks
from .program import Program, as_uint32_array

def gram_i8_ok(dim: int) -> bool:
    """The INT8 Gram's constraint on the state dimension (csrc/gram_i8.cu: gram_i8_supported)."""
    return dim >= 64 and dim % 8 == 0


_DIAG_FLAGS = {"one": _lib.GRAM_DIAG_ONE, "zero": _lib.GRAM_DIAG_ZERO, "computed": 0}


class CompiledProgram:
    """A planned gate list living behind a b200sq_program handle."""

    def __init__(self, program: Program, params: Optional[np.ndarray], tile_bits: int = 0):
        self._lib = _lib.load()
        self.program = program
        self._handle = ctypes.c_void_p()
        gates = program.c_gates(params)
        _lib.check(self._lib.b200sq_program_create(program.num_qubits, len(program), gates,
                                                   int(tile_bits), ctypes.byref(self._handle)))
        self.params = None if params is None else np.array(params, dtype=np.float64)

    @property
    def handle(self):
        return self._handle

    @property
    def num_passes(self) -> int:
        return self._lib.b200sq_program_num_passes(self._handle)

    def set_params(self, params: Optional[np.ndarray]):
        """New circuit parameters: only the angle coefficients change, the plan is kept."""
        if params is not None and self.params is not None and np.array_equal(np.asarray(params, dtype=np.float64), self.params):
            return   # unchanged (every sub-block of a sharded Gram, every step of a benchmark loop)
        c0, c1 = self.program.coefficient_arrays(params)
        n = max(len(c0), 1)
        a0 = (ctypes.c_double * n)(*c0)
        a1 = (ctypes.c_double * n)(*c1)
        _lib.check(self._lib.b200sq_program_set_coeffs(self._handle, a0, a1))
        self.params = None if params is None else np.array(params, dtype=np.float64)

    def pass_layouts(self):
        """[(in_qubits, out_qubits)] per pass: the active-qubit counts of the layout each pass reads and
        writes (a pass moves 16 * (2^in + 2^out) bytes per state; in = 0 means nothing is read)."""
        w = self.dump()
        n_pass = int(w[2])
        pos, out = 5, []
        for _ in range(n_pass):
            T, a_in, a_out = int(w[pos]), int(w[pos + 1]) & 0xffffffff, int(w[pos + 2]) & 0xffffffff
            out.append((bin(a_in).count("1"), bin(a_out).count("1")))
            pos += 11 + T
        return out

    def fp64_instructions_per_state(self):
        """Estimated FP64 instructions (DMUL / DFMA) per state and pass, from the plan: 4 per amplitude for a
        one-qubit gate slot, a phase-table or stray diagonal multiply and each product table of the expansion,
        8 for a general / controlled 2x2, 32 for a 4x4; only the tiles that are not identically zero count
        (tile qubits plus the outer qubits that are already active in the pass's input layout)."""
        w = self.dump()
        n, _, n_pass, n_rounds, n_steps = (int(v) for v in w[:5])
        pos, passes = 5, []
        for _ in range(n_pass):
            Tp, a_in, a_out, last, r0, r1, s0, s1, mat, nfresh, ntabs = (int(v) for v in w[pos:pos + 11])
            tile_mask = 0
            for q in w[pos + 11: pos + 11 + Tp]:
                tile_mask |= 1 << int(q)
            passes.append((Tp, a_in & 0xffffffff, r0, r1, nfresh, tile_mask))
            pos += 11 + Tp
        rounds = [tuple(int(v) for v in w[pos + 6 * i: pos + 6 * i + 6]) for i in range(n_rounds)]
        pos += 6 * n_rounds
        steps = [tuple(int(v) for v in w[pos + 7 * i: pos + 7 * i + 7]) for i in range(n_steps)]
        cost = {6: 4.0, 7: 4.0, 8: 8.0, 10: 2.0, 11: 2.0}   # layer steps: per active register slot
        out = []
        for Tp, a_in, r0, r1, nfresh, tile_mask in passes:
            per_amp = 0.0
            if nfresh:
                per_amp += (4.0 if a_in else 0.0) + (4.0 if nfresh > 6 else 0.0)
            for r in range(r0, r1):
                for kind, _, a, _, _, _, _ in steps[rounds[r][4]:rounds[r][5]]:
                    if kind in cost:
                        per_amp += cost[kind] * bin(a).count("1")
                    else:
                        per_amp += {9: 4.0, 2: 4.0, 5: 8.0}.get(kind, 32.0)
            live_bits = Tp + bin(a_in & ~tile_mask).count("1")
            out.append(per_amp * 2.0 ** min(live_bits, n))
        return out

    def dump(self) -> np.ndarray:
        n = ctypes.c_int(0)
        _lib.check(self._lib.b200sq_program_dump(self._handle, None, 0, ctypes.byref(n)))
        buf = (ctypes.c_int32 * n.value)()
        _lib.check(self._lib.b200sq_program_dump(self._handle, buf, n.value, ctypes.byref(n)))
        return np.frombuffer(buf, dtype=np.int32).copy()

    def __del__(self):
        try:
            if self._handle:
                self._lib.b200sq_program_destroy(self._handle)
                self._handle = ctypes.c_void_p()
        except Exception:  # interpreter shutdown
            pass


class Engine:
    """One engine per process / GPU (``device`` = CUDA device index, default: current)."""

    #: cap of the scratch buffer for expectation values of states that do not fit one tile
    workspace_bytes = 2 << 30

    def __init__(self, device: Optional[int] = None):
        import torch

        self._torch = torch
        self._lib = _lib.load()
        if not torch.cuda.is_available() or self._lib.b200sq_device_count() < 1:
            raise RuntimeError(
                "squlearn_b200 needs a CUDA device (built for sm_100a / B200); there is no CPU "
                "fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", int(device))
        self._workspace = None

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def to_device(self, arr, dtype=None):
        torch = self._torch
        if isinstance(arr, torch.Tensor):
            t = arr.to(self.device, dtype=dtype or arr.dtype, non_blocking=True)
        else:
            t = torch.as_tensor(np.ascontiguousarray(arr), dtype=dtype).to(self.device,
                                                                           non_blocking=True)
        return t.contiguous()

    def _features(self, program: Program, x):
        torch = self._torch
        xd = self.to_device(x, torch.float64)
        if xd.dim() == 1:
            xd = xd.reshape(-1, max(program.num_features, 1))
        if xd.dim() != 2 or (program.num_features and xd.shape[1] != program.num_features):
            raise ValueError(
                f"features must have shape (N, {program.num_features}), got {tuple(xd.shape)}")
        return xd

    def _table(self, cprog: CompiledProgram, x):
        prog = cprog.program
        if not prog.table_gates:
            return None
        xh = x.detach().cpu().numpy() if isinstance(x, self._torch.Tensor) else np.asarray(x)
        return self.to_device(prog.angle_table(np.asarray(xh, dtype=np.float64), cprog.params),
                              self._torch.float64)

    @staticmethod
    def _variants(variants):
        """(n, gate ptr, shift ptr, keep-alive[, second gate ptr, second shift ptr]) of a variant spec
        ``(gate_idx, shift)`` or ``(gate_idx, shift, gate_idx2, shift2)`` (two shifts per variant)."""
        if variants is None:
            return 1, None, None, None
        arrs = [np.ascontiguousarray(a, dtype=np.int32 if k % 2 == 0 else np.float64) for k, a in enumerate(variants)]
        if len(arrs) not in (2, 4) or any(len(a) != len(arrs[0]) for a in arrs):
            raise ValueError("variants must be (gate_idx, shift) or (gate_idx, shift, gate_idx2, shift2)")
        ptrs = [a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32 if k % 2 == 0 else ctypes.c_double))
                for k, a in enumerate(arrs)]
        return (len(arrs[0]), ptrs[0], ptrs[1], arrs) + tuple(ptrs[2:])

    @staticmethod
    def _masks(terms, n_qubits):
        xs, zs = [], []
        for t in terms:
            if isinstance(t, str):
                if len(t) != n_qubits:
                    raise ValueError("Pauli string length does not match the number of qubits")
                xm, zm = pauli_masks(t)
            else:
                xm, zm = t
            xs.append(xm)
            zs.append(zm)
        return as_uint32_array(xs), as_uint32_array(zs)

    # ------------------------------------------------------------------ entry points
    def compile(self, program: Program, params=None, tile_bits: int = 0) -> CompiledProgram:
        return CompiledProgram(program, params, tile_bits)

    def simulate(self, cprog: CompiledProgram, x, variants=None):
        """States of every data point: complex128 tensor (N, 2^n), or (V, N, 2^n) with variants."""
        torch = self._torch
        prog = cprog.program
        xd = self._features(prog, x)
        n_samples = xd.shape[0]
        nv, vg, vs, keep = self._variants(variants)[:4]
        if variants is not None and len(variants) > 2:
            raise ValueError("states of doubly shifted circuits are not offered; use simulate_expvals")
        table = self._table(cprog, x)
        with torch.cuda.device(self.device):
            psi = torch.empty((nv, n_samples, 1 << prog.num_qubits), dtype=torch.complex128,
                              device=self.device)
            # the last gate pass also reduces the row statistics the INT8 Gram's slicing needs (b200sq_row_stats, 16
            # bytes per state: row maximum for the digit form, 2-norm for the CRT form): a later gram() /
            # slice_states() of exactly this tensor then reads the states once, not twice
            row_max = torch.empty((nv, n_samples, 4), dtype=torch.int32, device=self.device) \
                if gram_i8_ok(1 << prog.num_qubits) else None
            _lib.check(self._lib.b200sq_simulate_rowmax(
                cprog.handle, n_samples, xd.shape[1], xd.data_ptr(),
                table.data_ptr() if table is not None else None, nv, vg, vs, psi.data_ptr(),
                row_max.data_ptr() if row_max is not None else None, self._stream()))
        if variants is not None:
            return psi
        out = psi[0]
        if row_max is not None:
            out._b200sq_row_max = (row_max[0], out._version)
        return out

    def row_stats_of(self, psi):
        """Row statistics ([N, 4] int32 view of b200sq_row_stats) simulate() attached to ``psi``, or None."""
        return self._row_max_of(psi) if psi.is_contiguous() else None

    @staticmethod
    def _row_max_of(psi):
        """The row statistics simulate() attached to this very tensor, if it has not been written to since."""
        tag = getattr(psi, "_b200sq_row_max", None)
        if tag is not None and tag[1] == psi._version and tag[0].shape[0] == psi.shape[0]:
            return tag[0]
        return None

    def simulate_expvals(self, cprog: CompiledProgram, x, terms: Sequence, variants=None):
        """Pauli expectation values <P_t> of every data point: float64 (N, n_terms) or (V, N, n_terms)."""
        torch = self._torch
        prog = cprog.program
        xd = self._features(prog, x)
        n_samples = xd.shape[0]
        nv, vg, vs, keep, *second = self._variants(variants)
        vg2, vs2 = second if second else (None, None)
        xm, zm = self._masks(terms, prog.num_qubits)
        table = self._table(cprog, x)
        with torch.cuda.device(self.device):
            out = torch.empty((nv, n_samples, len(terms)), dtype=torch.float64, device=self.device)
            ws_ptr, ws_bytes = None, 0
            if cprog.num_passes > 1 or prog.num_qubits > 13:
                state_bytes = 16 << prog.num_qubits
                want = min(max(self.workspace_bytes // state_bytes, 1), n_samples) * state_bytes
                if self._workspace is None or self._workspace.numel() < want:
                    self._workspace = None
                    self._workspace = torch.empty(want, dtype=torch.uint8, device=self.device)
                ws_ptr, ws_bytes = self._workspace.data_ptr(), self._workspace.numel()
            _lib.check(self._lib.b200sq_simulate_expvals2(
                cprog.handle, n_samples, xd.shape[1], xd.data_ptr(),
                table.data_ptr() if table is not None else None, nv, vg, vs, vg2, vs2, len(terms), xm, zm,
                out.data_ptr(), ws_ptr, ws_bytes, self._stream()))
        return out if variants is not None else out[0]

    def expvals(self, psi, n_qubits: int, terms: Sequence):
        torch = self._torch
        psi = psi.contiguous()
        if psi.dtype != torch.complex128 or psi.shape[-1] != (1 << n_qubits):
            raise ValueError("psi must be complex128 with 2^n amplitudes per state")
        n_states = psi.numel() >> n_qubits
        xm, zm = self._masks(terms, n_qubits)
        with torch.cuda.device(self.device):
            out = torch.empty((n_states, len(terms)), dtype=torch.float64, device=self.device)
            _lib.check(self._lib.b200sq_expvals(psi.data_ptr(), n_states, n_qubits, len(terms), xm,
                                                zm, out.data_ptr(), self._stream()))
        return out.reshape(tuple(psi.shape[:-1]) + (len(terms),))

    def gram(self, psi_x, psi_y=None, diagonal: str = "one", out=None, three_m: Optional[bool] = None,
             precision: Optional[str] = None):
        """K[i, j] = |<psi_x[i]|psi_y[j]>|^2.  ``psi_y=None`` means the symmetric Gram of psi_x;
        ``diagonal`` in {"one", "zero", "computed"} then sets the diagonal policy.
        ``precision``: "fp64" (FP64 DMMA), "int8" (split-integer FP64 emulation on the INT8 tcgen05 tensor
        cores: 22 digit-pair products, |dK| ~ 1e-12, north-star bound 1e-10), "crt" (the residue-number form of
        the same idea: 11 products per entry at 2^16 amplitudes, CRT reconstruction, same bound) or "auto" (crt /
        int8 where they are faster: 2^n >= 4096 and at least 1024 x 512 entries); default from
        B200SQ_GRAM_PRECISION, else "auto"."""
        torch = self._torch
        rm_x = self._row_max_of(psi_x) if psi_x.is_contiguous() else None
        rm_y = rm_x if psi_y is None else (self._row_max_of(psi_y) if psi_y.is_contiguous() else None)
        psi_x = psi_x.contiguous()
        sym = psi_y is None
        psi_y = psi_x if sym else psi_y.contiguous()
        if psi_x.dtype != torch.complex128 or psi_y.dtype != torch.complex128:
            raise ValueError("states must be complex128")
        if psi_x.dim() != 2 or psi_y.dim() != 2 or psi_x.shape[1] != psi_y.shape[1]:
            raise ValueError("states must be (N, 2^n) arrays with equal 2^n")
        nx, ny, dim = psi_x.shape[0], psi_y.shape[0], psi_x.shape[1]
        flags = (_lib.GRAM_SYMMETRIC | _DIAG_FLAGS[diagonal]) if sym else 0
        if three_m is None:  # default: 3M (Gauss) complex product; B200SQ_GRAM_MODE=4m selects 4M
            three_m = os.environ.get("B200SQ_GRAM_MODE", "3m").lower() != "4m"
        if three_m:
            flags |= _lib.GRAM_3M
            if os.environ.get("B200SQ_GRAM_MODE", "").lower() == "3m-classic" or three_m == "classic":
                flags |= _lib.GRAM_CLASSIC
        scheme = self.gram_precision(nx, ny, dim, precision)
        if scheme in ("int8", "crt"):
            flags |= _lib.GRAM_INT8 | (_lib.GRAM_CRT if scheme == "crt" else 0)
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty((nx, ny), dtype=torch.float64, device=self.device)
            _lib.check(self._lib.b200sq_gram_rowmax(psi_x.data_ptr(), nx, psi_y.data_ptr(), ny, dim,
                                                    out.data_ptr(), out.stride(0), flags,
                                                    rm_x.data_ptr() if rm_x is not None else None,
                                                    rm_y.data_ptr() if rm_y is not None else None, self._stream()))
        return out

    def alloc_planes(self, plane_rows: int, dim: int):
        """Empty digit planes ([12, plane_rows, 2 * dim] int8) and row exponents ([plane_rows] int32)."""
        torch = self._torch
        with torch.cuda.device(self.device):
            return (torch.empty((12, plane_rows, 2 * dim), dtype=torch.int8, device=self.device),
                    torch.zeros((plane_rows,), dtype=torch.int32, device=self.device))

    def plane_count(self, dim: int, scheme: str = "int8") -> int:
        """Planes of ONE operand half (nat; the rot planes double it): 6 digit slices, or the CRT moduli."""
        if scheme == "crt":
            s = int(self._lib.b200sq_gram_crt_moduli(int(dim)))
            if s <= 0:
                raise ValueError("the CRT Gram does not support this state dimension")
            return s
        return 6

    def alloc_plane_block(self, rows: int, dim: int, scheme: str = "int8"):
        """One contiguous int8 buffer holding the planes of ``rows`` states ([2 s][rows][2 dim]: nat then rot;
        s = 6 digit slices or the CRT moduli) followed by their int32 row exponents.  Returns (buffer, planes view,
        expo view)."""
        torch = self._torch
        n_pl = 2 * self.plane_count(dim, scheme)
        nb = n_pl * rows * 2 * dim
        with torch.cuda.device(self.device):
            buf = torch.empty((nb + 4 * rows,), dtype=torch.int8, device=self.device)
        return buf, buf[:nb].view(n_pl, rows, 2 * dim), buf[nb:].view(torch.int32)

    def alloc_bytes(self, nbytes: int):
        """Uninitialised int8 device buffer (exchange arenas and staging blocks of the multi-GPU Gram)."""
        torch = self._torch
        with torch.cuda.device(self.device):
            return torch.empty((int(nbytes),), dtype=torch.int8, device=self.device)

    def slice_states(self, psi, planes, expo, row0: int = 0, scheme: Optional[str] = None, digits=None, stats=None):
        """Write the planes of the states ``psi`` into rows [row0, row0 + len(psi)) of ``planes`` (digit slices, or
        CRT residues when ``scheme == "crt"`` / the plane count says so).  ``digits`` (CRT only): a [6][N][2 dim]
        int8 tensor that receives the nat digit planes of the same states (the multi-GPU exchange format).
        ``stats``: the row statistics of exactly these states (rows of what ``row_stats_of`` returned for the tensor
        ``psi`` is a row range of); default: the statistics simulate() attached to ``psi`` itself."""
        torch = self._torch
        rm = stats if stats is not None else (self._row_max_of(psi) if psi.is_contiguous() else None)
        if rm is not None and (rm.shape[0] != psi.shape[0] or not rm.is_contiguous()):
            raise ValueError("row statistics do not match the states")
        psi = psi.contiguous()
        if psi.dtype != torch.complex128 or psi.dim() != 2:
            raise ValueError("states must be a (N, 2^n) complex128 array")
        if scheme is None:
            scheme = "int8" if planes.shape[0] == 12 else "crt"
        rmp = rm.data_ptr() if rm is not None else None
        with torch.cuda.device(self.device):
            if scheme == "crt":
                _lib.check(self._lib.b200sq_gram_slice_crt_digits(
                    psi.data_ptr(), psi.shape[0], psi.shape[1], planes.data_ptr(), planes.shape[1], int(row0),
                    expo.data_ptr(), rmp, digits.data_ptr() if digits is not None else None,
                    digits.shape[1] if digits is not None else 0, 0, self._stream()))
            else:
                _lib.check(self._lib.b200sq_gram_slice_rowmax(
                    psi.data_ptr(), psi.shape[0], psi.shape[1], planes.data_ptr(), planes.shape[1], int(row0),
                    expo.data_ptr(), rmp, self._stream()))

    def digits_to_crt(self, digits, rows_d: int, row0_d: int, n: int, dim: int, out, rows_o: int, row0_o: int = 0):
        """Residue planes ([s][rows_o][2 dim], nat only) of ``n`` rows of digit planes ([>= 6][rows_d][2 dim]); the
        operands may be tensors or raw device pointers (received blocks of the multi-GPU exchange)."""
        ptr = lambda t: int(t) if isinstance(t, int) else t.data_ptr()
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b200sq_gram_digits_to_crt(ptr(digits), int(rows_d), int(row0_d), int(n), int(dim),
                                                           ptr(out), int(rows_o), int(row0_o), self._stream()))

    def gram_from_planes(self, planes_x, expo_x, x0: int, nx: int, planes_y=None, expo_y=None, y0: int = 0,
                         ny: int = 0, diagonal: str = "one", out=None, scheme: Optional[str] = None):
        """Gram of rows [x0, x0 + nx) of planes_x against rows [y0, y0 + ny) of planes_y
        (``planes_y=None``: the symmetric Gram of the x rows)."""
        torch = self._torch
        sym = planes_y is None
        if sym:
            planes_y, expo_y, y0, ny = planes_x, expo_x, x0, nx
        dim = planes_x.shape[2] // 2
        flags = (_lib.GRAM_SYMMETRIC | _DIAG_FLAGS[diagonal]) if sym else 0
        if scheme is None:
            scheme = "int8" if planes_x.shape[0] == 12 else "crt"
        if scheme == "crt":
            flags |= _lib.GRAM_CRT
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty((nx, ny), dtype=torch.float64, device=self.device)
            _lib.check(self._lib.b200sq_gram_from_planes(
                planes_x.data_ptr(), expo_x.data_ptr(), planes_x.shape[1], int(x0), int(nx), planes_y.data_ptr(),
                expo_y.data_ptr(), planes_y.shape[1], int(y0), int(ny), dim, out.data_ptr(), out.stride(0), flags,
                self._stream()))
        return out

    def gram_with_alignment(self, psi, y, diagonal: str = "one", precision: Optional[str] = None):
        """Symmetric Gram matrix of ``psi`` plus the kernel-target-alignment sums (sum_ij K_ij y_i y_j, sum_ij K_ij^2)
        of src/squlearn/kernel/loss/target_alignment.py:75-90.  On the INT8 path the sums come out of the |.|^2
        epilogue of the Gram kernel (K is never re-read); on the FP64 DMMA path (small problems) they are two device
        reductions over K.  Returns (K, sums) as device tensors."""
        torch = self._torch
        psi = psi.contiguous()
        n, dim = psi.shape
        yd = self.to_device(y, torch.float64).reshape(-1)
        if yd.shape[0] != n:
            raise ValueError("labels and states differ in length")
        scheme = self.gram_precision(n, n, dim, precision)
        if scheme in ("int8", "crt"):
            _, planes, expo = self.alloc_plane_block(n, dim, scheme)
            self.slice_states(psi, planes, expo, 0, scheme)
            with torch.cuda.device(self.device):
                k = torch.empty((n, n), dtype=torch.float64, device=self.device)
                sums = torch.zeros(2, dtype=torch.float64, device=self.device)
            self.gram_jobs(planes, expo, [{"x0": 0, "nx": n, "planes_y": planes, "expo_y": expo, "rows_y": n, "y0": 0,
                                           "ny": n, "out": k, "symmetric": True, "diagonal": diagonal,
                                           "align": (yd, yd, sums, 1.0)}])
            return k, sums
        k = self.gram(psi, None, diagonal=diagonal, precision="fp64")
        return k, torch.stack([torch.dot(yd, k @ yd), torch.sum(k * k)])

    def gram_jobs(self, planes_x, expo_x, jobs: Sequence[dict]):
        """Several Gram blocks sharing the row planes in ONE persistent launch (b200sq_gram_jobs).

        Every job is a dict: ``x0, nx`` (rows of planes_x), ``planes_y, expo_y`` (tensors or raw device pointers of
        the column planes: only the 6 nat planes are read), ``rows_y`` (rows per plane of planes_y), ``y0, ny``,
        ``out`` (2-D float64 view the block is written to), optional ``symmetric`` + ``diagonal``, optional
        ``ready=(flag_ptr, value)`` (start the block when the uint32 at flag_ptr is >= value), optional
        ``align=(yx, yy, out2, weight)`` (fused target-alignment sums, see include/b200sq.h) and optional
        ``mirror=(ptr, ld)`` (also store the transposed block there: the owner of the columns, maybe a peer)."""
        torch = self._torch
        ptr = lambda t: int(t) if isinstance(t, int) else t.data_ptr()
        crt = planes_x.shape[0] != 12     # the row planes tell the scheme: 12 = digit slices, else 2 s CRT residue planes
        arr = (_lib.GramJob * max(len(jobs), 1))()
        for k, j in enumerate(jobs):
            out = j["out"]
            if out.dtype != torch.float64 or out.dim() != 2 or (out.shape[1] > 1 and out.stride(1) != 1):
                raise ValueError("job output must be a row-major float64 2-D view")
            if tuple(out.shape) != (j["nx"], j["ny"]):
                raise ValueError("job output shape does not match (nx, ny)")
            g = arr[k]
            g.x0, g.nx, g.y0, g.ny = int(j["x0"]), int(j["nx"]), int(j["y0"]), int(j["ny"])
            g.planes_y_dev, g.expo_y_dev = ptr(j["planes_y"]), ptr(j["expo_y"])
            g.plane_rows_y = int(j["rows_y"])
            g.K_dev, g.ldk = out.data_ptr(), int(out.stride(0)) if out.shape[0] > 1 else max(int(out.stride(0)), out.shape[1])
            g.flags = (_lib.GRAM_SYMMETRIC | _DIAG_FLAGS[j.get("diagonal", "one")]) if j.get("symmetric") else 0
            if crt:
                g.flags |= _lib.GRAM_CRT
            if j.get("ready") is not None:
                g.ready_flag_dev, g.ready_value = int(j["ready"][0]), int(j["ready"][1])
            if j.get("align") is not None:
                yx, yy, out2, weight = j["align"]
                g.align_yx_dev, g.align_yy_dev, g.align_out_dev = ptr(yx), ptr(yy), ptr(out2)
                g.align_weight = float(weight)
            if j.get("mirror") is not None:      # (pointer of the transposed block, its leading dimension)
                g.K_mirror_dev, g.ldk_mirror = int(j["mirror"][0]), int(j["mirror"][1])
        dim = planes_x.shape[2] // 2
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b200sq_gram_jobs(planes_x.data_ptr(), expo_x.data_ptr(), planes_x.shape[1], dim,
                                                  len(jobs), arr, self._stream()))

    # ------------------------------------------------------------------ peer plumbing (multi-GPU Gram)
    def ipc_export(self, tensor):
        """(64-byte CUDA IPC handle of the allocation holding ``tensor``, offset of the tensor inside it)."""
        handle = ctypes.create_string_buffer(64)
        off = ctypes.c_int64(0)
        _lib.check(self._lib.b200sq_ipc_export(tensor.data_ptr(), handle, ctypes.byref(off)))
        return handle.raw, int(off.value)

    def ipc_open(self, handle: bytes) -> int:
        base = ctypes.c_void_p()
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b200sq_ipc_open(handle, ctypes.byref(base)))
        return int(base.value)

    def ipc_close(self, base: int):
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b200sq_ipc_close(ctypes.c_void_p(base)))

    def wait_flags(self, flags_ptr: int, which: Sequence[int], value: int):
        """Enqueue a wait on the current stream until the uint32 flags flags_ptr[which[k]] are all >= value."""
        arr = (ctypes.c_int32 * max(len(which), 1))(*[int(w) for w in which])
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b200sq_wait_flags(ctypes.c_void_p(flags_ptr), arr, len(which), int(value),
                                                   self._stream()))

    def ready_timeouts(self) -> int:
        """Gram jobs whose ready flag never arrived (synchronises; non-zero means a peer's push was lost)."""
        with self._torch.cuda.device(self.device):
            return int(self._lib.b200sq_gram_ready_timeouts())

    def copy_async(self, dst_ptr: int, src_ptr: int, nbytes: int, stream=None):
        """Copy-engine device-to-device copy (dst may be a peer mapping) on ``stream`` (default: current)."""
        st = self._stream() if stream is None else ctypes.c_void_p(stream.cuda_stream)
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b200sq_copy_async(ctypes.c_void_p(dst_ptr), ctypes.c_void_p(src_ptr), int(nbytes), st))

    def copy_planes_async(self, dst_ptr: int, dst_pitch: int, src_ptr: int, src_pitch: int, nbytes: int, n_planes: int,
                          stream=None):
        """``n_planes`` copy-engine copies of ``nbytes`` each (plane p: src + p * src_pitch -> dst + p * dst_pitch)."""
        st = self._stream() if stream is None else ctypes.c_void_p(stream.cuda_stream)
        with self._torch.cuda.device(self.device):
            _lib.check(self._lib.b200sq_copy_planes_async(ctypes.c_void_p(dst_ptr), int(dst_pitch), ctypes.c_void_p(src_ptr),
                                                          int(src_pitch), int(nbytes), int(n_planes), st))

    @staticmethod
    def gram_precision(nx: int, ny: int, dim: int, precision: Optional[str] = None) -> str:
        """Resolve the Gram arithmetic: "fp64" (DMMA), "int8" (digit-slice tcgen05 path) or "crt" (residue-number
        tcgen05 path: fewer products, more plane memory — 2 x 11..12 bytes per amplitude component instead of 2 x 6)."""
        if precision is None:
            precision = os.environ.get("B200SQ_GRAM_PRECISION", "auto").lower()
        if precision not in ("fp64", "int8", "crt", "auto"):
            raise ValueError("precision must be 'fp64', 'int8', 'crt' or 'auto'")
        if precision == "auto":
            if dim >= 4096 and nx * ny >= 1024 * 512:
                # CRT while its planes (2 * 12 * 2 bytes per amplitude and operand) stay below ~48 GiB
                precision = "crt" if 48.0 * dim * (nx + ny) < 48 * 2.0 ** 30 and dim <= 1 << 18 else "int8"
            else:
                precision = "fp64"
        if precision in ("int8", "crt") and (dim < 64 or dim % 8):
            precision = "fp64"
        if precision == "crt" and dim > 1 << 21:
            precision = "int8"
        return precision

    def rbf(self, fx, fy, gamma: float):
        torch = self._torch
        fx = self.to_device(fx, torch.float64)
        fy = self.to_device(fy, torch.float64)
        with torch.cuda.device(self.device):
            out = torch.empty((fx.shape[0], fy.shape[0]), dtype=torch.float64, device=self.device)
            _lib.check(self._lib.b200sq_rbf(fx.data_ptr(), fx.shape[0], fy.data_ptr(), fy.shape[0],
                                            fx.shape[1], float(gamma), out.data_ptr(),
                                            out.stride(0), self._stream()))
        return out

    def to_host(self, t) -> np.ndarray:
        """Device -> host through pinned memory (full PCIe rate); the NumPy array owns the buffer."""
        torch = self._torch
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host.numpy()

    def launch_count(self) -> int:
        return int(self._lib.b200sq_launch_count())
