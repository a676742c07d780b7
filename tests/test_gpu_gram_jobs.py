"""GPU tests (-m gpu) of the job-table INT8 Gram (b200sq_gram_jobs): several blocks that share the row planes in
one persistent launch, blocks whose column planes arrive late (ready flags written by a copy-engine copy on another
stream, exactly what a peer's push does in the multi-GPU exchange) and the fused target-alignment sums.
Reference semantics: the pair loop of fidelity_kernel_statevector.py:358-383 through the oracle."""

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

INT8_TOL = 1e-10


@pytest.fixture(scope="module")
def eng():
    import squlearn_b200 as sq

    return sq.Executor("b200").engine


def _random_states(rng, n, dim):
    a = rng.normal(size=(n, dim)) + 1j * rng.normal(size=(n, dim))
    return a / np.linalg.norm(a, axis=1, keepdims=True)


def _planes(eng, states):
    import torch

    t = torch.from_numpy(states).cuda()
    buf, planes, expo = eng.alloc_plane_block(states.shape[0], states.shape[1])
    eng.slice_states(t, planes, expo, 0)
    return buf, planes, expo


@pytest.mark.parametrize("dim,n_local,widths", [(4096, 300, (300, 150, 77)), (16384, 512, (512, 512, 256)),
                                                (64, 37, (5,)), (1024, 129, (260, 1))])
def test_job_table_matches_the_oracle_block_by_block(eng, dim, n_local, widths):
    """Diagonal block + rectangular blocks against received [6 planes | exponents] buffers of different widths,
    one launch; sub-ranges of the row operand (the antipodal split) included."""
    import torch

    rng = np.random.default_rng(dim + n_local)
    a = _random_states(rng, n_local, dim)
    _, pa, ea = _planes(eng, a)
    n_total = n_local + sum(widths)
    rows = torch.full((n_local, n_total), -7.0, dtype=torch.float64, device="cuda")
    jobs = [{"x0": 0, "nx": n_local, "planes_y": pa, "expo_y": ea, "rows_y": n_local, "y0": 0, "ny": n_local,
             "out": rows[:, :n_local], "symmetric": True, "diagonal": "one"}]
    cols, keep, off = [], [], n_local
    for k, w in enumerate(widths):
        b = _random_states(rng, w, dim)
        _, pb, eb = _planes(eng, b)
        nat = pb[:6].contiguous()                       # a received block carries only the nat planes
        keep.append((nat, eb))
        x0, nx = (0, n_local) if k % 2 == 0 else (n_local // 3, n_local - n_local // 3)
        jobs.append({"x0": x0, "nx": nx, "planes_y": nat, "expo_y": eb, "rows_y": w, "y0": 0, "ny": w,
                     "out": rows[x0:x0 + nx, off:off + w]})
        cols.append((b, x0, nx, off, w))
        off += w
    eng.gram_jobs(pa, ea, jobs)
    got = rows.cpu().numpy()
    want = oracle.fidelity_gram_zgemm(a, a, True, "off_diagonal")
    assert np.max(np.abs(got[:, :n_local] - want)) < INT8_TOL
    assert np.array_equal(got[:, :n_local], got[:, :n_local].T)
    for b, x0, nx, off, w in cols:
        want = oracle.fidelity_gram_zgemm(a[x0:x0 + nx], b, False)
        assert np.max(np.abs(got[x0:x0 + nx, off:off + w] - want)) < INT8_TOL
        untouched = np.ones(n_local, dtype=bool)
        untouched[x0:x0 + nx] = False
        assert np.all(got[untouched, off:off + w] == -7.0)     # rows outside the job's range are not written


def test_ready_flags_written_late_by_the_copy_engine(eng):
    """The column planes of two jobs are copied in on a second stream AFTER the Gram kernel has started (a long
    sleep kernel delays them); each copy is followed by a 4-byte flag copy, as in the peer push.  The persistent
    kernel must pick the blocks up when their flags appear — and must not read them earlier."""
    import torch

    dim, n_local, w = 8192, 256, 256
    rng = np.random.default_rng(3)
    a, b1, b2 = (_random_states(rng, m, dim) for m in (n_local, w, w))
    _, pa, ea = _planes(eng, a)
    _, p1, e1 = _planes(eng, b1)
    _, p2, e2 = _planes(eng, b2)
    src = [(p1[:6].contiguous(), e1), (p2[:6].contiguous(), e2)]
    arena = [(torch.zeros_like(s), torch.zeros_like(e)) for s, e in src]     # zeros: a premature read gives K = 0
    flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    epoch = torch.full((1,), 5, dtype=torch.int32, device="cuda")
    rows = torch.zeros((n_local, n_local + 2 * w), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        torch.cuda._sleep(int(2e8))                                            # ~0.1 s
        for k, ((s, e), (d, de)) in enumerate(zip(src, arena)):
            eng.copy_async(d.data_ptr(), s.data_ptr(), s.numel(), side)
            eng.copy_async(de.data_ptr(), e.data_ptr(), 4 * e.numel(), side)
            eng.copy_async(flags.data_ptr() + 4 * (k + 1), epoch.data_ptr(), 4, side)
            torch.cuda._sleep(int(4e7))
    jobs = [{"x0": 0, "nx": n_local, "planes_y": pa, "expo_y": ea, "rows_y": n_local, "y0": 0, "ny": n_local,
             "out": rows[:, :n_local], "symmetric": True, "diagonal": "computed"}]
    for k, (d, de) in enumerate(arena):
        jobs.append({"x0": 0, "nx": n_local, "planes_y": d, "expo_y": de, "rows_y": w, "y0": 0, "ny": w,
                     "out": rows[:, n_local + k * w: n_local + (k + 1) * w],
                     "ready": (flags.data_ptr() + 4 * (k + 1), 5)})
    eng.gram_jobs(pa, ea, jobs)
    torch.cuda.synchronize()
    got = rows.cpu().numpy()
    for k, b in enumerate((b1, b2)):
        want = oracle.fidelity_gram_zgemm(a, b, False)
        assert np.max(np.abs(got[:, n_local + k * w: n_local + (k + 1) * w] - want)) < INT8_TOL
    assert np.max(np.abs(np.diag(got[:, :n_local]) - 1.0)) < INT8_TOL


def test_fused_target_alignment_sums(eng):
    """sum K_ij y_i y_j and sum K_ij^2 from the |.|^2 epilogue (target_alignment.py:75-90) equal the NumPy sums
    over the full matrix: symmetric block (mirror counted) and a rectangular block with weight 2."""
    import torch

    dim, n, w = 4096, 700, 300
    rng = np.random.default_rng(11)
    a, b = _random_states(rng, n, dim), _random_states(rng, w, dim)
    y = rng.choice([-1.0, 1.0], size=n)
    yb = rng.choice([-1.0, 1.0], size=w)
    _, pa, ea = _planes(eng, a)
    _, pb, eb = _planes(eng, b)
    yt, ybt = torch.from_numpy(y).cuda(), torch.from_numpy(yb).cuda()
    out_s = torch.zeros(2, dtype=torch.float64, device="cuda")
    out_r = torch.zeros(2, dtype=torch.float64, device="cuda")
    ks = torch.empty((n, n), dtype=torch.float64, device="cuda")
    kr = torch.empty((n, w), dtype=torch.float64, device="cuda")
    eng.gram_jobs(pa, ea, [
        {"x0": 0, "nx": n, "planes_y": pa, "expo_y": ea, "rows_y": n, "y0": 0, "ny": n, "out": ks,
         "symmetric": True, "diagonal": "one", "align": (yt, yt, out_s, 1.0)},
        {"x0": 0, "nx": n, "planes_y": pb, "expo_y": eb, "rows_y": w, "y0": 0, "ny": w, "out": kr,
         "align": (yt, ybt, out_r, 2.0)}])
    k1, k2 = ks.cpu().numpy(), kr.cpu().numpy()
    s, r = out_s.cpu().numpy(), out_r.cpu().numpy()
    assert abs(s[0] - y @ k1 @ y) < 1e-9 * n and abs(s[1] - np.sum(k1 * k1)) < 1e-9 * n
    assert abs(r[0] - 2 * (y @ k2 @ yb)) < 1e-9 * n and abs(r[1] - 2 * np.sum(k2 * k2)) < 1e-9 * n


def test_split_launch_when_there_are_more_jobs_than_one_table_holds(eng):
    import torch

    dim, n_local, w, n_jobs = 256, 64, 40, 15
    rng = np.random.default_rng(2)
    a = _random_states(rng, n_local, dim)
    _, pa, ea = _planes(eng, a)
    rows = torch.zeros((n_local, n_jobs * w), dtype=torch.float64, device="cuda")
    jobs, bs, keep = [], [], []
    for k in range(n_jobs):
        b = _random_states(rng, w, dim)
        _, pb, eb = _planes(eng, b)
        keep.append((pb, eb))
        bs.append(b)
        jobs.append({"x0": 0, "nx": n_local, "planes_y": pb, "expo_y": eb, "rows_y": w, "y0": 0, "ny": w,
                     "out": rows[:, k * w:(k + 1) * w]})
    eng.gram_jobs(pa, ea, jobs)
    got = rows.cpu().numpy()
    for k, b in enumerate(bs):
        assert np.max(np.abs(got[:, k * w:(k + 1) * w] - oracle.fidelity_gram_zgemm(a, b, False))) < INT8_TOL


def test_crt_planes_from_digits_agree_between_sender_and_receiver(eng):
    """The multi-GPU exchange sends 6 nat digit planes; a CRT receiver turns them into residue planes
    (b200sq_gram_digits_to_crt).  The sender's own residue planes, sliced together with the digits
    (b200sq_gram_slice_crt_digits), are residues of the SAME integers: the nat halves agree bit for bit, and a Gram block
    from converted planes equals the oracle."""
    import torch
    from tests.helpers import crt_digits, crt_slice

    dim, n = 2048, 70
    rng = np.random.default_rng(21)
    a = _random_states(rng, n, dim)
    a[:8] *= np.array([3.0, 0.01, 0.7, 1.9, 1.0, 1e-5, 40.0, 0.26])[:, None]   # rows of any norm
    t = torch.from_numpy(a).cuda()
    _, cplanes, cexpo = eng.alloc_plane_block(n, dim, "crt")
    digits = torch.empty((6, n, 2 * dim), dtype=torch.int8, device="cuda")
    eng.slice_states(t, cplanes, cexpo, 0, "crt", digits=digits)
    want_digits, want_expo = crt_digits(a)
    assert np.array_equal(cexpo.cpu().numpy(), want_expo) and np.array_equal(digits.cpu().numpy(), want_digits)
    assert np.array_equal(cplanes.cpu().numpy(), crt_slice(a)[0])
    s_ = cplanes.shape[0] // 2
    conv = torch.empty((s_, n, 2 * dim), dtype=torch.int8, device="cuda")
    eng.digits_to_crt(digits, n, 0, n, dim, conv, n, 0)
    assert torch.equal(conv, cplanes[:s_])
    b = _random_states(rng, 33, dim)
    _, bplanes, be = eng.alloc_plane_block(33, dim, "crt")          # a "remote" block: only its digits travel
    bd = torch.empty((6, 33, 2 * dim), dtype=torch.int8, device="cuda")
    eng.slice_states(torch.from_numpy(b).cuda(), bplanes, be, 0, "crt", digits=bd)
    bconv = torch.empty((s_, 33, 2 * dim), dtype=torch.int8, device="cuda")
    eng.digits_to_crt(bd, 33, 0, 33, dim, bconv, 33, 0)
    k = torch.empty((n, 33), dtype=torch.float64, device="cuda")
    eng.gram_jobs(cplanes, cexpo, [{"x0": 0, "nx": n, "planes_y": bconv, "expo_y": be, "rows_y": 33, "y0": 0, "ny": 33,
                                    "out": k}])
    want = oracle.fidelity_gram_zgemm(a, b, False)
    scale = np.sum(np.abs(a) ** 2, axis=1)[:, None]                  # |<x|y>|^2 scales with |x|^2 (|y| = 1)
    assert np.max(np.abs(k.cpu().numpy() - want) / scale) < INT8_TOL
