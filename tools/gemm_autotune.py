"""Exhaustive tile / K-split / CTA-pair sweep over the conv_gemm shapes of one denoise step at a given view count.
    python tools/gemm_autotune.py VIEWS [--min-gain 0.05]
Step 1 (subprocess, MD_TRACE=1): one eager step, the library prints every GEMM launch with the configuration its
heuristics chose.  Step 2: every distinct shape is timed with the library default and with every forced combination of
BN in {64,128,160,256}, ksplit in {1,2,3,4,6,8,12,16} and single / pair tiles: weights cold (L2 flushed), activations
re-touched after the flush (inside the step they were just written by the previous kernel).  Prints, per shape, the default
time, the best forced time and the launch-weighted gain."""
import collections
import os
import re
import subprocess
import sys

import torch

sys.path.insert(0, ".")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 2 and sys.argv[2] == "--trace":
    from morphablediffusion_b200 import synth
    from morphablediffusion_b200.engine import Engine
    n = int(sys.argv[1])
    eng = Engine(max_views_per_call=n)
    eng.load_state_dict(synth.make_state_dict())
    eng.bind(synth.make_batch(n), "perspective")
    x_t, x_input, clip = synth.make_inputs(n)
    eng.denoise_step(x_t[0].cuda().contiguous(), x_input[0].cuda().contiguous(), clip[0, 0].cuda().contiguous(), 40, 2.0, seed=1)
    torch.cuda.synchronize()
    sys.exit(0)

from morphablediffusion_b200 import _native as nat  # noqa: E402

views = int(sys.argv[1]) if len(sys.argv) > 1 else 2
env = dict(os.environ, MD_TRACE="1", MD_NO_GRAPH="1")
p = subprocess.run([sys.executable, __file__, str(views), "--trace"], capture_output=True, text=True, env=env, cwd=ROOT)
pat = re.compile(r"conv_gemm B=(\d+) D=(\d+) H=(\d+) W=(\d+) Cin=(\d+) taps=(\d+) N=(\d+) BN=(\d+) tiles=(\d+)x(\d+) ks=(\d+) "
                 r"split_tiles=(\d+) cg2=(\d+) act=(\d+) f32=(\d+) bf16=(\d+) res=(\d+) stats=(\d+) rowvec=(\d+) tail=(\d+) stride=(\d+)")
shapes = collections.OrderedDict()
for line in p.stderr.splitlines():
    m = pat.search(line)
    if not m:
        continue
    v = tuple(int(x) for x in m.groups())
    key = v[:7] + v[13:]
    if key not in shapes:
        shapes[key] = [0, v[7], v[10], v[12], v[11]]
    shapes[key][0] += 1
print(f"views={views}: {sum(s[0] for s in shapes.values())} GEMM launches, {len(shapes)} distinct shapes", flush=True)

ACTS = {0: "none", 1: "silu", 2: "relu", 3: "geglu", 4: "gelu", 5: "quickgelu"}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, A, reps=5):
    ts = []
    for r in range(reps):
        flush.fill_(r)
        A.add_(0)   # activations back into L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


tot_def = tot_best = 0.0
rows = []
for key, (count, BN0, ks0, cg0, st0) in shapes.items():
    B, D, H, W, K, ntaps, N, act, f32, bf16, res, stats, rowvec, tail, stride = key
    if stride or ntaps not in (1, 9, 27):
        continue
    M = B * D * H * W
    A = torch.randn(B, D, H, W, K, device="cuda").to(torch.bfloat16)
    if ntaps == 9:
        taps = [(dx, dy, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    elif ntaps == 27:
        taps = [(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    else:
        taps = [(0, 0, 0)]
    Wt = (torch.randn(N, K * ntaps, device="cuda") / (K * ntaps) ** 0.5).to(torch.bfloat16)
    n_out = N // 2 if act == 3 else N
    kw = dict(B=B, D=D, H=H, W=W, Cin=K, N=N, taps=taps, bias=torch.randn(N, device="cuda"), act=ACTS[act])
    if f32:
        kw["out_f32"] = torch.zeros(M, n_out, device="cuda")
    if bf16:
        kw["out_bf16"] = torch.zeros(M, n_out, device="cuda", dtype=torch.bfloat16)
    if res == 1:
        kw["res_f32"] = torch.randn(M, n_out, device="cuda")
    if res == 2:
        kw["res_bf16"] = torch.randn(M, n_out, device="cuda").to(torch.bfloat16)
    if stats:
        kw["col_stats"] = torch.zeros(B, n_out, 2, device="cuda")
    if rowvec:
        kw["rowvec"] = torch.randn(B, n_out, device="cuda")
    try:
        t_def = timed(lambda: nat.conv_gemm(A, Wt, **kw), A)
    except nat.MdiffError as e:
        print("skip", key, str(e)[:80])
        continue
    best = (t_def, "default")
    for BN in (64, 128, 160, 256):
        for ks in (1, 2, 3, 4, 6, 8, 12, 16):
            for pair in (-1, 1):
                if pair == 1 and (BN not in (160, 256) or ks != 1):
                    continue
                try:
                    t = timed(lambda: nat.conv_gemm(A, Wt, BN=BN, ksplit=ks, cta_pair=pair, **kw), A, reps=3)
                except nat.MdiffError:
                    continue
                if t < best[0]:
                    best = (t, f"BN={BN} ks={ks} pair={pair}")
    tot_def += t_def * count
    tot_best += best[0] * count
    rows.append(((t_def - best[0]) * count, key, count, BN0, ks0, cg0, t_def, best))
    print(f"M={M:6d} K={K:5d}x{ntaps:2d} N={N:5d} act={act} res={res} st={stats} n={count:2d} default(BN={BN0} ks={ks0} cg2={cg0}) "
          f"{t_def:6.1f} us  best {best[0]:6.1f} us [{best[1]}]  gain {100 * (1 - best[0] / t_def):4.1f}%", flush=True)
    del A, Wt, kw
print(f"TOTAL default {tot_def / 1e3:.3f} ms  best {tot_best / 1e3:.3f} ms  ({100 * (1 - tot_best / tot_def):.1f}% of GEMM time)")
rows.sort(reverse=True)
print("top launch-weighted gains:")
for g, key, count, BN0, ks0, cg0, t_def, best in rows[:25]:
    B, D, H, W, K, ntaps, N = key[:7]
    print(f"  {g:7.1f} us/step  M={B * D * H * W:6d} K={K}x{ntaps} N={N} n={count} default(BN={BN0} ks={ks0} cg2={cg0}) {t_def:.1f} -> {best[0]:.1f} [{best[1]}]")
