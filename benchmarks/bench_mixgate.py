#!/usr/bin/env python
"""Times the tcgen05 1x1 conv + bias + gate (MSCA.channel_mix and the `x *` gate, reference layers/msca.py:23,26) and the
fused MSCA depthwise chain on the MSCAN-T shapes; reports achieved HBM GB/s on the algorithmic bytes (a + gate + y for the
mix, x + s for the chain) against the measured copy peak.    python benchmarks/bench_mixgate.py [--batch 32]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "convnet-approximater_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch  # noqa: E402


def t(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    from approx import backend as B
    from approx.layers import MSCA
    from approx.layers import _msca_fused
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    for (C, H) in [(32, 128), (64, 64), (160, 32), (256, 16)]:
        torch.manual_seed(0)
        m = MSCA(C, 5, (7, 11, 21)).cuda().bfloat16().eval()
        x = torch.randn(a.batch, H, H, C, device="cuda").bfloat16()
        elems = x.numel()
        with torch.no_grad():
            s = _msca_fused.conv0_sd_forward(m, x)
            w_nc = B.conv1x1_pack(m.channel_mix.weight)
            bias = m.channel_mix.bias.float()
            ms_mix = t(lambda: B.conv1x1_tc(s, w_nc, bias, gate=x))
            ms_sd = t(lambda: _msca_fused.conv0_sd_forward(m, x))
        print(json.dumps({"bench": "mix_gate_tcgen05", "C": C, "H": H, "batch": a.batch, "ms": round(ms_mix, 4),
                          "hbm_GBps": round(3 * elems * 2 / ms_mix / 1e6, 1), "frac_hbm_peak": round(3 * elems * 2 / ms_mix / 1e6 / hbm, 3)}))
        print(json.dumps({"bench": "msca_sd_fused_chain", "C": C, "H": H, "batch": a.batch, "ms": round(ms_sd, 4),
                          "hbm_GBps": round(2 * elems * 2 / ms_sd / 1e6, 1), "frac_hbm_peak": round(2 * elems * 2 / ms_sd / 1e6 / hbm, 3),
                          "TFMA_per_s": round(elems * 103 / ms_sd / 1e9, 2)}))


if __name__ == "__main__":
    main()
