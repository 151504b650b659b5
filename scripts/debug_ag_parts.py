"""Loop-back AG->GEMM at bench size: where do mismatches sit (which launch, rank, 128-row block)?  A/B of the part flags."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_tp_loopback_gpu as T

def run(world, M, N, K, b_mn, pairs, parts, launches=4, use_xfull=False):
    os.environ["QN_TP_AG_PARTS"] = str(parts)
    lib = T._lib()
    rows = M // world
    gather = [[torch.zeros(M * K, dtype=torch.bfloat16, device="cuda") for _ in range(world)] for _ in range(2)]
    flags = [torch.zeros(1 << 18, dtype=torch.int32, device="cuda") for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    cap = (148 // world) & ~1
    T._set_env(pairs, "pump")
    torch.cuda.synchronize()
    for seq in range(1, launches + 1):
        xs = [T._rnd(rows, K) for _ in range(world)]
        ws = [T._rnd(K, N, scale=K ** -0.5) if b_mn else T._rnd(N, K, scale=K ** -0.5) for _ in range(world)]
        ys = [torch.empty(M, N, dtype=torch.bfloat16, device="cuda") for _ in range(world)]
        par = seq & 1
        xfull = [torch.zeros(M, K, dtype=torch.bfloat16, device="cuda") if (use_xfull and seq % 2 == 1) else None for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            rc = lib.qn_gemm_ag_bf16(xs[r].data_ptr(), T._ptrs(gather[par]), T._ptrs(flags), r, world, seq, ws[r].data_ptr(),
                                     ys[r].data_ptr(), M, N, K, ws[r].stride(0), N, int(b_mn), 0, None, None, 0,
                                     0, cap, xfull[r].data_ptr() if xfull[r] is not None else None, streams[r].cuda_stream)
            assert rc == 0, rc
        torch.cuda.synchronize()
        xf = torch.cat(xs).float()
        for r in range(world):
            w = ws[r].float()
            ref = xf @ (w if b_mn else w.t())
            bad = ((ys[r].float() - ref).abs() > 6e-2 + 3e-2 * ref.abs()).view(M // 128, 128, N).any(dim=2).any(dim=1)
            gsrc = xfull[r] if xfull[r] is not None else gather[par][r].view(M, K)
            gbad = (gsrc.float() != xf).view(M // 128, 128 * K).any(dim=1)
            print(f"xfull={use_xfull} pairs={pairs} parts={parts} seq={seq} rank={r}: bad y blocks {bad.nonzero().flatten().tolist()[:40]} "
                  f"(n={int(bad.sum())}); bad gather blocks n={int(gbad.sum())} {gbad.nonzero().flatten().tolist()[:20]}", flush=True)

for pairs in (True, False):
    for parts in (4, 1):
        run(2, 16384, 512, 1024, True, pairs, parts)
run(2, 16384, 1536, 1024, False, True, 4)
for parts in (4, 1):
    run(2, 16384, 512, 1024, True, True, parts, use_xfull=True)
