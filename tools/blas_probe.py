"""cuBLAS FP64 baselines on this GPU: DGEMM 8192^3 and the ZGEMM of the C3 Gram (4096 x 4096 x 65536)."""
import torch, time
dev = torch.device("cuda")
def bench(f, reps=3):
    f(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); f(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
m = 8192
a = torch.randn(m, m, dtype=torch.float64, device=dev); b = torch.randn(m, m, dtype=torch.float64, device=dev)
ms = bench(lambda: a @ b)
print(f"DGEMM {m}^3: {ms:.2f} ms  {2*m**3/ms/1e9:.2f} TFLOP/s")
del a, b
N, D = 4096, 65536
x = torch.randn(N, D, dtype=torch.complex128, device=dev)
ms = bench(lambda: (x.conj() @ x.T).abs_().square_())
print(f"ZGEMM+abs^2 {N}x{N}x{D}: {ms:.2f} ms  {8*N*N*D/ms/1e9:.2f} TFLOP/s (8 flop per complex MAC)  {N*N/ms*1e3:.3e} entries/s")
ms = bench(lambda: torch.mm(x.conj(), x.T))
print(f"ZGEMM only: {ms:.2f} ms  {8*N*N*D/ms/1e9:.2f} TFLOP/s")
xr = torch.view_as_real(x).reshape(N, 2 * D)
ms = bench(lambda: xr @ xr.T)
print(f"DGEMM NT {N}x{N}x{2*D}: {ms:.2f} ms  {2*N*N*2*D/ms/1e9:.2f} TFLOP/s")
