"""Engine cross-check for psgd_gemm on the GPU box: tcgen05 3xTF32 vs float64 torch matmul (diagnostic output)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgd_tf_b200 as psgd
from psgd_tf_b200._lib import check

torch.manual_seed(0)
ctx = psgd.get_context()


def run(M, N, K, ta, tb, engine=2, triu=0, a_tri=0, b_tri=0, bn=128, mode=1, pair=1):
    ctx.set_option("tc_bn", bn)
    ctx.set_option("tc_mode", mode)
    ctx.set_option("tc_pair", pair)
    A = torch.randn((K, M) if ta else (M, K), device="cuda")
    B = torch.randn((N, K) if tb else (K, N), device="cuda")
    if a_tri:
        opA = A.t() if ta else A
        opA = torch.triu(opA) if a_tri == 1 else torch.tril(opA)
        A = (opA.t() if ta else opA).contiguous()
    if b_tri:
        opB = B.t() if tb else B
        opB = torch.triu(opB) if b_tri == 1 else torch.tril(opB)
        B = (opB.t() if tb else opB).contiguous()
    Cm = torch.full((M, N), float("nan"), device="cuda")
    check(ctx.lib.psgd_gemm(ctx.handle, engine, M, N, K, C.c_void_p(A.data_ptr()), A.shape[1], ta,
                            C.c_void_p(B.data_ptr()), B.shape[1], tb, C.c_void_p(Cm.data_ptr()), N, triu, a_tri, b_tri))
    torch.cuda.synchronize()
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    if triu:
        ref = torch.triu(ref)
    err = (Cm.double() - ref)
    rel = (err.norm() / ref.norm()).item()
    nan = torch.isnan(Cm).sum().item()
    msg = f"M={M} N={N} K={K} ta={ta} tb={tb} eng={engine} mode={'TS' if mode else 'SS'}{'+pair' if (pair and mode) else ''} triu={triu} tri=({a_tri},{b_tri}): rel={rel:.3e} nan={nan}"
    if not (rel < 1e-5):
        # where are the errors? per 32x32 block
        e = torch.nan_to_num(err, nan=1e3).abs()
        mb, nb = (M + 31) // 32, (N + 31) // 32
        pad = torch.zeros(mb * 32, nb * 32, device="cuda", dtype=torch.double)
        pad[:M, :N] = e
        blk = pad.view(mb, 32, nb, 32).amax(dim=(1, 3))
        msg += "\n  block max-abs-err (32x32 blocks, first 8x8):\n" + str((blk[:8, :8]).cpu().numpy().round(3))
        msg += f"\n  C[0,:4]={Cm[0,:4].tolist()} ref={ref[0,:4].tolist()}"
    print(msg, flush=True)
    return rel


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "basic"
    if which == "basic":
        run(256, 256, 256, 0, 1, engine=1)
        for mode in (1, 0):
            for ta, tb in ((0, 1), (1, 1), (0, 0), (1, 0)):
                run(256, 256, 256, ta, tb, mode=mode)
            for ta, tb in ((0, 1), (1, 0)):
                run(384, 640, 320, ta, tb, mode=mode)
                run(1000, 520, 264, ta, tb, mode=mode)
            run(512, 512, 512, 0, 1, triu=1, mode=mode)
            run(512, 512, 512, 0, 0, a_tri=1, b_tri=1, triu=1, mode=mode)
            run(100, 4096, 2048, 1, 0, mode=mode)
            run(4096, 4096, 4096, 0, 1, mode=mode)
    elif which == "pairbasic":
        # CTA-pair kernel (cta_group::2): every operand layout, ragged edges, triangular hints, one tiny case first
        run(256, 128, 32, 0, 1)
        for ta, tb in ((0, 1), (1, 1), (0, 0), (1, 0)):
            run(256, 256, 256, ta, tb)
            run(384, 640, 320, ta, tb)
            run(1000, 520, 264, ta, tb)
        run(512, 512, 512, 0, 1, triu=1)
        run(1024, 1024, 1024, 0, 0, a_tri=1, b_tri=1, triu=1)
        run(1024, 1024, 1024, 1, 0, a_tri=2, b_tri=1)
        run(4096, 4096, 4096, 0, 1)
    elif which == "pairperf":
        def timeit(args, n=20):
            for _ in range(3):
                check(ctx.lib.psgd_gemm(*args))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                check(ctx.lib.psgd_gemm(*args))
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        for (M, N, K) in ((4096, 4096, 4096), (4096, 4096, 1024), (4096, 1024, 1024), (1024, 4096, 1024)):
            for (ta, tb) in ((0, 1), (0, 0), (1, 0)):
                A = torch.randn((K, M) if ta else (M, K), device="cuda"); B = torch.randn((N, K) if tb else (K, N), device="cuda")
                Cm = torch.empty(M, N, device="cuda")
                args = (ctx.handle, 2, M, N, K, C.c_void_p(A.data_ptr()), A.shape[1], ta, C.c_void_p(B.data_ptr()), B.shape[1], tb,
                        C.c_void_p(Cm.data_ptr()), N, 0, 0, 0)
                out = []
                for pair in (0, 1):
                    ctx.set_option("tc_pair", pair)
                    ms = timeit(args)
                    out.append(f"{'pair  ' if pair else 'single'} {ms:.3f} ms {2 * M * N * K / ms / 1e9:6.1f} TF")
                print(f"M={M} N={N} K={K} ta={ta} tb={tb}: " + " | ".join(out), flush=True)
        ctx.set_option("tc_pair", 1)
    elif which == "ksweep":
        # per-tile fixed cost vs per-stage cost of the single-CTA kernel: time(K) at M = N = 4096, with and without the
        # epilogue's global stores (debug bit 8; results wrong)
        ctx.set_option("tc_pair", 0)
        M = N = 4096
        for dbg, epi in ((0, 2), (8, 2)):
            ctx.set_option("tc_debug", dbg)
            ctx.set_option("tc_epi", epi)
            for K in (128, 256, 512, 1024, 2048, 4096, 8192):
                A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); Cm = torch.empty(M, N, device="cuda")
                args = (ctx.handle, 2, M, N, K, C.c_void_p(A.data_ptr()), K, 0, C.c_void_p(B.data_ptr()), K, 1,
                        C.c_void_p(Cm.data_ptr()), N, 0, 0, 0)
                for _ in range(5):
                    check(ctx.lib.psgd_gemm(*args))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50):
                    check(ctx.lib.psgd_gemm(*args))
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 50
                print(f"debug={dbg} epi={epi} K={K:5d}: {ms * 1e3:8.1f} us  {2 * M * N * K / ms / 1e9:6.1f} TF  per-tile(7 tiles/CTA) {ms * 1e3 / 7:6.1f} us", flush=True)
        ctx.set_option("tc_debug", 0)
        ctx.set_option("tc_epi", 2)
    elif which == "pairablate":
        # where does the pair kernel's time go?  (results are wrong with debug bits set)
        M = N = K = 4096
        A = torch.randn(M, K, device="cuda"); B = torch.randn(K, N, device="cuda"); Cm = torch.empty(M, N, device="cuda")
        args = (ctx.handle, 2, M, N, K, C.c_void_p(A.data_ptr()), K, 0, C.c_void_p(B.data_ptr()), N, 1,
                C.c_void_p(Cm.data_ptr()), N, 0, 0, 0)
        def loop(n=200):
            for _ in range(5):
                check(ctx.lib.psgd_gemm(*args))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                check(ctx.lib.psgd_gemm(*args))
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        for pair in (1, 0):
            ctx.set_option("tc_pair", pair)
            for dbg, name in ((0, "full"), (1, "no B split"), (2, "no A split"), (3, "no split"), (4, "no MMA"), (7, "TMA + barriers only")):
                ctx.set_option("tc_debug", dbg)
                ms = loop()
                print(f"{'pair  ' if pair else 'single'} {name:22s}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:6.1f} TFLOP/s-equiv", flush=True)
            ctx.set_option("tc_debug", 0)
        ctx.set_option("tc_pair", 1)
    elif which == "ablate":
        # where does the time go?  (results are wrong with debug bits set)  + clocks/power under a sustained loop
        import subprocess, time
        M = N = K = 4096
        A = torch.randn(M, K, device="cuda"); B = torch.randn(K, N, device="cuda"); Cm = torch.empty(M, N, device="cuda")
        args = (ctx.handle, 2, M, N, K, C.c_void_p(A.data_ptr()), K, 0, C.c_void_p(B.data_ptr()), N, 1,
                C.c_void_p(Cm.data_ptr()), N, 0, 0, 0)
        def loop(fn, n):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            q = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                                 stdout=subprocess.PIPE, text=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record(); torch.cuda.synchronize()
            q.terminate()
            lines = [l.split(",") for l in q.stdout.read().strip().splitlines() if "," in l]
            clk = sorted(float(l[0]) for l in lines); pw = sorted(float(l[1]) for l in lines)
            return e0.elapsed_time(e1) / n, (clk[len(clk) // 2] if clk else 0), (pw[len(pw) // 2] if pw else 0)
        for mode in (1, 0):
            ctx.set_option("tc_mode", mode)
            for dbg, name in ((0, "full"), (1, "no B split"), (2, "no A split"), (3, "no split"), (4, "no MMA"), (7, "TMA + barriers only")):
                ctx.set_option("tc_debug", dbg)
                ms, clk, pw = loop(lambda: check(ctx.lib.psgd_gemm(*args)), 1500)
                print(f"mode={'TS' if mode else 'SS'} {name:22s}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:6.1f} TFLOP/s-equiv   sm {clk:.0f} MHz  {pw:.0f} W", flush=True)
            ctx.set_option("tc_debug", 0)
        torch.backends.cuda.matmul.allow_tf32 = True
        ms, clk, pw = loop(lambda: A @ B, 1500)
        print(f"cuBLAS tf32 4096^3: {ms:.3f} ms {2 * M * N * K / ms / 1e9:.1f} TFLOP/s  sm {clk:.0f} MHz {pw:.0f} W", flush=True)
    elif which == "perf":
        import time
        for mode in (1, 0):
            ctx.set_option("tc_mode", mode)
            bn = "TS" if mode else "SS"
            for (ta, tb) in ((0, 1), (0, 0), (1, 0)):
                M = N = K = 4096
                A = torch.randn(M, K, device="cuda"); B = torch.randn(K, N, device="cuda"); Cm = torch.empty(M, N, device="cuda")
                args = (ctx.handle, 2, M, N, K, C.c_void_p(A.data_ptr()), K, ta, C.c_void_p(B.data_ptr()), N, tb,
                        C.c_void_p(Cm.data_ptr()), N, 0, 0, 0)
                for _ in range(3):
                    check(ctx.lib.psgd_gemm(*args))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    check(ctx.lib.psgd_gemm(*args))
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print(f"mode={bn} ta={ta} tb={tb}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s fp32-equivalent "
                      f"({3 * 2 * M * N * K / ms / 1e9:.1f} TF32 TFLOP/s issued)", flush=True)
        # cuBLAS TF32 calibration (roofline denominator; calibration only, not on the product path)
        torch.backends.cuda.matmul.allow_tf32 = True
        A = torch.randn(8192, 8192, device="cuda"); B = torch.randn(8192, 8192, device="cuda")
        for _ in range(3):
            A @ B
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            A @ B
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"cuBLAS tf32 8192^3: {ms:.3f} ms {2 * 8192**3 / ms / 1e9:.1f} TFLOP/s", flush=True)
