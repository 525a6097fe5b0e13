#!/usr/bin/env python
"""One forward of an MSCA block in its three forms (original, MscaRep d0+fix, d1+fix) — the command line ncu profiles for the
MSCA kernels' launch list.    python benchmarks/run_msca_once.py [--C 64 --H 64 --batch 32]"""
import argparse
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "convnet-approximater_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    for k, v in dict(C=64, H=64, batch=32).items():
        ap.add_argument(f"--{k}", type=int, default=v)
    a = ap.parse_args()
    from approx.core import MscaRep
    from approx.layers import MSCA
    torch.manual_seed(0)
    src = MSCA(a.C, 5, (7, 11, 21))
    forms = {"original": src}
    for name, (d, f) in {"rep_d0_fix": (0, True), "rep_d1_fix": (1, True)}.items():
        app = MscaRep(decomp=d, fix=f)
        sub = app.initialize(src)
        with contextlib.redirect_stdout(io.StringIO()):
            app.optimize(sub)
        forms[name] = app.postprocess(sub)
    x = torch.randn(a.batch, a.H, a.H, a.C, device="cuda").bfloat16()
    with torch.no_grad():
        for name, m in forms.items():
            m = m.cuda().bfloat16().eval()
            for _ in range(2):
                y = m.forward_nhwc(x)
            torch.cuda.synchronize()
            print(name, float(y.float().abs().mean()))


if __name__ == "__main__":
    main()
