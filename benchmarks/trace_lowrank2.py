#!/usr/bin/env python
"""Pipeline timeline of the fused scheme-2 kernel: per-tile clock stamps of CTA 0 (cna_lowrank2_fused_forward_traced).

    python benchmarks/trace_lowrank2.py [--C 64 --N 64 --K 64 --H 224 --batch 64]
Prints, per tile, cycles relative to the previous tile's same event, and the waits between roles.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "convnet-approximater_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch  # noqa: E402

EV = ["x_issued", "g1_turn", "x_landed", "g1_issued", "g2_turn", "t_ready", "tepi_start", "t_stored", "yepi_start", "y_store",
      "t_loaded", "a_free", "a_issued", "y_loaded", "stg_free", "y_staged", "g2_acc", "g2_issued", "y_sts_done", "y_fenced"]


def main():
    ap = argparse.ArgumentParser()
    for k, v in dict(C=64, N=64, K=64, H=224, batch=64, tiles=48, quiet=0).items():
        ap.add_argument(f"--{k}", type=int, default=v)
    a = ap.parse_args()
    from approx import backend as B
    torch.manual_seed(0)
    v_w = torch.randn(a.K, a.C, 3, 1, device="cuda") / (3 * a.C) ** 0.5
    h_w = torch.randn(a.N, a.K, 1, 3, device="cuda") / (3 * a.K) ** 0.5
    wv, wh = B.lowrank2_pack(v_w, h_w)
    x = torch.randn(a.batch, a.H, a.H, B.round_up(a.C, 16), device="cuda").bfloat16()
    for _ in range(2):
        y, tr = B.lowrank2_fused_traced(x, wv, wh, None, a.tiles)
    torch.cuda.synchronize()
    t = tr.cpu().numpy().astype("int64")
    wall = t[a.tiles]
    t = t[:a.tiles]
    cyc, ns = wall[2] - wall[0], wall[3] - wall[1]
    ntile = a.batch * -(-a.H // 8) * -(-a.H // 14)
    print(f"CTA 0: {cyc} cycles in {ns} ns -> effective SM clock {cyc / ns * 1e3:.0f} MHz; {ntile} tiles, "
          f"{-(-ntile // 148)} per CTA -> {cyc / -(-ntile // 148):.0f} cycles/tile incl. prologue")
    t0 = t[t > 0].min()
    t = t - t0
    print("tile " + " ".join(f"{e:>10s}" for e in EV))
    for i in range(0 if not a.quiet else a.tiles, a.tiles):
        print(f"{i:4d} " + " ".join(f"{v:10d}" for v in t[i]))
    import numpy as np
    lo = a.tiles // 3
    per = np.diff(t[lo:, :], axis=0).mean(axis=0)
    print("steady-state cycles per tile by event:", dict(zip(EV, per.round().astype(int))))
    d = t[lo:]
    print("mean waits (cycles):")
    print("  g1: turn -> x landed        ", (d[:, 2] - d[:, 1]).mean().round())
    print("  g1: landed -> issued        ", (d[:, 3] - d[:, 2]).mean().round())
    print("  g1 issued -> tepi start     ", (d[:, 6] - d[:, 3]).mean().round(), "(GEMM1 execution + commit latency)")
    print("  tepi start -> T stored      ", (d[:, 7] - d[:, 6]).mean().round())
    print("  T stored -> g2 sees T ready ", (d[:, 5] - d[:, 7]).mean().round(), "(negative: MMA warp arrived later)")
    print("  g2: turn -> T ready         ", (d[:, 5] - d[:, 4]).mean().round())
    print("  g2 ready -> yepi start      ", (d[:, 8] - d[:, 5]).mean().round(), "(GEMM2 execution + commit latency)")
    print("  yepi start -> store issued  ", (d[:, 9] - d[:, 8]).mean().round())
    print("  T-epilogue: start -> TMEM loaded -> A buffer free -> copies issued -> stored:",
          (d[:, 10] - d[:, 6]).mean().round(), (d[:, 11] - d[:, 10]).mean().round(), (d[:, 12] - d[:, 11]).mean().round(), (d[:, 7] - d[:, 12]).mean().round())
    print("  GEMM2 warp: T ready -> accumulator free -> issued (-> next turn):", (d[:, 16] - d[:, 5]).mean().round(), (d[:, 17] - d[:, 16]).mean().round(),
          (d[1:, 4] - d[:-1, 17]).mean().round())
    print("  Y-epilogue thread 0: staging free -> its stores issued -> fence done -> group barrier:", (d[:, 18] - d[:, 14]).mean().round(),
          (d[:, 19] - d[:, 18]).mean().round(), (d[:, 15] - d[:, 19]).mean().round())
    print("  Y-epilogue: start -> TMEM loaded -> staging free -> staged -> store issued:",
          (d[:, 13] - d[:, 8]).mean().round(), (d[:, 14] - d[:, 13]).mean().round(), (d[:, 15] - d[:, 14]).mean().round(), (d[:, 9] - d[:, 15]).mean().round())


if __name__ == "__main__":
    main()
