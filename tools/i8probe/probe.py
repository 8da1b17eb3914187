import ctypes as C, os, torch, time
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libi8probe.so"))
lib.i8gemm.restype = C.c_int
lib.i8gemm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
dev = "cuda"
ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
def run(M, N, K, lda=None, check=False, reps=5):
    lda = lda or K
    A = torch.randint(-127, 128, (M, lda), dtype=torch.int8, device=dev)
    B = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)     # column-major K x N == row-major N x K
    D = torch.empty((M, N), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.i8gemm(A.data_ptr(), lda, B.data_ptr(), K, D.data_ptr(), N, M, N, K, ws.data_ptr(), ws.numel(), st)
    torch.cuda.synchronize()
    assert rc == 0, rc
    if check:
        ref = (A[:512, :K].double() @ B.double().T)
        err = (D[:512].double() - ref).abs().max().item()
        print("  check max abs err", err)
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib.i8gemm(A.data_ptr(), lda, B.data_ptr(), K, D.data_ptr(), N, M, N, K, ws.data_ptr(), ws.numel(), st); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"M={M} N={N} K={K} lda={lda}: {best:.3f} ms  {2.0*M*N*K/best/1e9:.1f} TOP/s", flush=True)
run(4096, 1024, 1024, check=True)
run(1_000_000, 1024, 1024, check=True)
for lev in (2, 4, 8):
    run(1_000_000, 1024, 1024 * lev, lda=8192)
run(8192, 8192, 8192)
