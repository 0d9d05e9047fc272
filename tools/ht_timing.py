"""Halftone timing on device-resident frames (CUDA events), old kernels against the v2 kernels and
their register-allocation variants; also checks that v2 and v1 agree bit for bit at full size.

    gpurun -- 'python tools/ht_timing.py'"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from dither_pie_b200 import _capi, engine, synth
    _capi.ensure_device()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    pal = engine.get_palette(synth.hex_palette(synth.PICO8))
    for (label, h, w, nf) in (("1080p", 1080, 1920, 64), ("4k", 2160, 3840, 16)):
        frames = np.stack([synth.frame(h, w, t) for t in range(2)])
        src = torch.from_numpy(np.concatenate([frames] * (nf // 2))).to(dev)
        dst = torch.empty_like(src)
        idx = torch.empty((nf, h, w), dtype=torch.uint8, device=dev)
        for params in ({}, {"cell_size": 5, "angle": 30.0, "shape": "diamond"}, {"cell_size": 12, "angle": 0.0}):
            plan = engine.Plan("halftone", params, h, w)
            os.environ["DP_HT_V1"] = "1"
            ms1 = timed(lambda: plan.run(pal, src.data_ptr(), nf, dst.data_ptr(), idx.data_ptr(), sp))
            ref, ref_i = dst.clone(), idx.clone()
            ms1c = timed(lambda: plan.run(pal, src.data_ptr(), nf, dst.data_ptr(), None, sp))
            del os.environ["DP_HT_V1"]
            line = f"{label} {params}: v1 {ms1c:.3f} ms (+idx {ms1:.3f})"
            for occ in ("64", "44", "46", "66"):
                os.environ["DP_HT_OCC"] = occ
                dst.zero_()
                idx.zero_()
                ms2 = timed(lambda: plan.run(pal, src.data_ptr(), nf, dst.data_ptr(), idx.data_ptr(), sp))
                ok = bool(torch.equal(dst, ref)) and bool(torch.equal(idx, ref_i))
                ms2c = timed(lambda: plan.run(pal, src.data_ptr(), nf, dst.data_ptr(), None, sp))
                ms2i = timed(lambda: plan.run(pal, src.data_ptr(), nf, None, idx.data_ptr(), sp))
                frac = 6 * nf * h * w / (ms2c * 1e-3) / 1e9 / 6535.1
                line += f" | occ{occ} {ms2c:.3f} ({frac:.3f}) +idx {ms2:.3f} idx-only {ms2i:.3f} {'same' if ok else 'DIFFERENT'}"
            del os.environ["DP_HT_OCC"]
            print(line, flush=True)


if __name__ == "__main__":
    main()
