"""GPU parity tests proper: the CUDA path (through the ctypes C ABI) against
  (1) the golden vectors generated from the live reference, and
  (2) the oracle on fresh seeded inputs (sizes the oracle finishes in seconds), and
  (3) size-independent properties at BASELINE.json's full sizes.
Bit-exact for every image mode; k-means centres within 1e-3 (north_star's tolerance)."""
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu

import dither_pie_b200 as dp  # noqa: E402
from dither_pie_b200 import engine, kmeans, synth  # noqa: E402
from dither_pie_b200.dithering_lib import DitherMode  # noqa: E402
from dither_pie_b200.video_processor import (VideoProcessor, pixelize_regular_array,  # noqa: E402
                                             _resample_array)
from oracle import dither_oracle as O  # noqa: E402  (the checker)


def gpu(img, pal, mode, params=None, gamma=False, **kw):
    return engine.dither_frames(img, pal, mode, params, use_gamma=gamma, **kw)


def mismatch(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return int((a != b).any(axis=-1).sum())


# ------------------------------------------------------------------ golden (reference outputs)
def test_golden_dither_cases(golden_cases):
    data, meta = golden_cases
    bad = []
    for n, m in enumerate(meta):
        out = gpu(data["img_" + m["image"]], data["pal_" + m["palette"]], m["mode"], m["params"],
                  m["gamma"])
        k = mismatch(out, data[f"out_{n}"])
        if k:
            bad.append((n, m["image"], m["palette"], m["mode"], m["params"], m["gamma"], k))
    assert not bad, (len(bad), bad[:8])


def test_golden_diffusion_big():
    g = load_golden("diffusion_big.npz")
    for v in ("floyd_steinberg", "atkinson", "jjn", "sierra"):
        out = gpu(g["img"], g["pal"], "error_diffusion", {"variant": v})
        assert mismatch(out, g["ed_" + v]) == 0, v
    assert mismatch(gpu(g["img"], g["pal16"], "ostromoukhov"), g["ostro"]) == 0


def test_golden_hybrid_cases():
    """hybrid mode against outputs of the live reference (numba path), incl. gamma."""
    g = load_golden("hybrid_cases.npz")
    meta = json.load(open(os.path.join(GOLDEN, "hybrid_cases.json")))
    for n, m in enumerate(meta):
        out = gpu(g["img_" + m["image"]], g["pal_" + m["palette"]], "hybrid", m["params"], m["gamma"])
        assert mismatch(out, g[f"out_{n}"]) == 0, (n, m)


def test_golden_perceptual_cases():
    """perceptual mode against outputs of the live reference (pure-Python loop), incl. gamma."""
    g = load_golden("perceptual_cases.npz")
    meta = json.load(open(os.path.join(GOLDEN, "perceptual_cases.json")))
    for n, m in enumerate(meta):
        out = gpu(g["img_" + m["image"]], g["pal_" + m["palette"]], "perceptual", {}, m["gamma"])
        assert mismatch(out, g[f"out_{n}"]) == 0, (n, m)


def test_golden_adaptive_cases():
    """adaptive_variance against outputs of the live reference (pure Python + scipy filter)."""
    g = load_golden("adaptive_cases.npz")
    meta = json.load(open(os.path.join(GOLDEN, "adaptive_cases.json")))
    for n, m in enumerate(meta):
        out = gpu(g["img_" + m["image"]], g["pal_" + m["palette"]], "adaptive_variance", m["params"],
                  m["gamma"])
        assert mismatch(out, g[f"out_{n}"]) == 0, (n, m)


def test_golden_pixelize_and_final_resize():
    g = load_golden("pixelize.npz")
    small = g["small"]
    for m in (2, 3, 5):
        h, w, _ = small.shape
        nh, nw = h * m + (h * m) % 2, w * m + (w * m) % 2
        assert np.array_equal(_resample_array(small, nh, nw), g[f"small_x{m}_even"])
        assert np.array_equal(_resample_array(small, h * m, w * m), g[f"small_x{m}_cli"])
    meta = json.load(open(os.path.join(GOLDEN, "pixelize.json")))
    for mm in meta[:12]:
        w, h = mm["w"], mm["h"]
        xs = np.arange(w)[None, :].repeat(h, 0)
        ys = np.arange(h)[:, None].repeat(w, 1)
        img = np.stack([xs % 256, ys % 256, (xs // 256) * 16 + (ys // 256)], 2).astype(np.uint8)
        out = pixelize_regular_array(img, mm["max_size"])
        key = f"{w}x{h}_{mm['max_size']}"
        assert out.shape[:2] == (mm["th"], mm["tw"])
        xt = out[0, :, 0].astype(np.int32) + 256 * (out[0, :, 2].astype(np.int32) // 16)
        yt = out[:, 0, 1].astype(np.int32) + 256 * (out[:, 0, 2].astype(np.int32) % 16)
        assert np.array_equal(xt, g["xt_" + key]) and np.array_equal(yt, g["yt_" + key])


def test_golden_kmeans():
    g = load_golden("kmeans.npz")
    for t, k in enumerate((16, 8, 5)):
        c, it = kmeans.kmeans_fit(g[f"sample_{t}"], k, 42)
        assert np.abs(c - g[f"centers_{t}"]).max() <= 1e-3, (t, np.abs(c - g[f"centers_{t}"]).max())
        assert it == int(g[f"niter_{t}"])
        assert np.array_equal(c.astype(int), g[f"palette_{t}"])
    img = synth.frame(300, 400, 3)
    random.seed(7)
    pal = dp.ColorReducer.generate_kmeans_palette(__import__("PIL.Image").Image.fromarray(img), 8)
    assert [tuple(int(v) for v in r) for r in pal] == [tuple(int(v) for v in r) for r in g["palette_1"]]


# ------------------------------------------------------------------ oracle on fresh inputs
PALS = {
    "pico8": synth.hex_palette(synth.PICO8), "c64": synth.hex_palette(synth.C64),
    "gb4": synth.hex_palette(synth.GB_POCKET), "r16": synth.random_palette(16),
    "r64": synth.random_palette(64), "r256": synth.random_palette(256),
    "lat27": synth.lattice_palette(27, 1, 127), "lat64": synth.lattice_palette(64, 3, 51),
    "one": np.array([[12, 200, 77]]), "two": np.array([[0, 0, 0], [255, 255, 255]]),
    "dup": np.array([[10, 10, 10], [200, 50, 50], [10, 10, 10], [200, 50, 50], [90, 90, 200]]),
}
THRESH_MODES = [("none", {}), ("bayer", {"size": "8x8"}), ("bayer", {"size": "2x2"}),
                ("bayer", {"size": "16x16"}), ("blue_noise", {"size": 32, "seed": 5}),
                ("IGN", {}), ("IGN", {"scale": 0.3, "seed": 9999}), ("polka_dot", {}),
                ("polka_dot", {"tile_size": 13, "gamma": 3.0})]


@pytest.mark.parametrize("pname", list(PALS))
def test_threshold_family_vs_oracle(pname):
    pal = PALS[pname]
    imgs = [synth.frame(67, 131, 11), synth.noise_frame(50, 77, 12),
            synth.blocks_frame(64, 96, 13, 8, 6), synth.blocks_frame(33, 47, 14, 4, 4)]
    for img in imgs:
        for mode, params in THRESH_MODES:
            ref = O.apply_dithering(img, pal, mode, params)
            assert mismatch(gpu(img, pal, mode, params), ref) == 0, (pname, mode, params, img.shape)


@pytest.mark.parametrize("k", [2, 3, 16, 27, 30, 31])
def test_threshold_v4_kernel_vs_oracle(k):
    """Widths that are multiples of 16 take k_thresh_v4 (K <= 30, shared-memory 32^3 table);
    K = 31 and odd widths take the older kernels.  Tiles that straddle frames, ragged last
    tiles, non-power-of-two widened matrices, cells with more than four candidates (noise and
    lattice palettes), index output."""
    pal = synth.random_palette(k, seed=100 + k) if k not in (16, 27) else (
        PALS["pico8"] if k == 16 else PALS["lat27"])
    modes = THRESH_MODES + [("bayer", {"size": "psx4x4"}), ("polka_dot", {"tile_size": 6}),
                            ("blue_noise", {"size": 64, "seed": 42})]
    for (h, w, kind) in [(16, 48, "noise"), (40, 64, "frame"), (33, 112, "blocks"), (96, 160, "noise")]:
        if kind == "noise":
            frames = np.stack([synth.noise_frame(h, w, 50 + t) for t in range(3)])
        elif kind == "blocks":
            frames = np.stack([synth.blocks_frame(h, w, 60 + t, 4, 5) for t in range(3)])
        else:
            frames = np.stack([synth.frame(h, w, 70 + t) for t in range(3)])
        for mode, params in modes:
            out, idx = engine.dither_frames(frames, pal, mode, params, return_indices=True)
            for t in range(3):
                ref = O.apply_dithering(frames[t], pal, mode, params)
                assert mismatch(out[t], ref) == 0, (k, h, w, kind, mode, params, t)
                # the index plane names a row with the output colour
                assert np.array_equal(np.asarray(pal, np.uint8)[idx[t]], out[t]), (k, mode)


@pytest.mark.parametrize("pname", ["pico8", "r64", "lat27"])
def test_threshold_family_gamma_vs_oracle(pname):
    pal = PALS[pname]
    img = synth.frame(45, 83, 21)
    for mode, params in THRESH_MODES[:6]:
        ref = O.apply_dithering(img, pal, mode, params, True)
        assert mismatch(gpu(img, pal, mode, params, True), ref) == 0, (pname, mode)


def test_batch_of_frames_and_ragged_sizes():
    pal = PALS["c64"]
    for (h, w) in [(1, 1), (1, 37), (29, 1), (3, 5), (64, 64), (65, 127), (128, 341)]:
        frames = np.stack([synth.frame(h, w, 30 + t) for t in range(3)])
        out = gpu(frames, pal, "bayer", {"size": "4x4"})
        for t in range(3):
            assert mismatch(out[t], O.apply_dithering(frames[t], pal, "bayer", {"size": "4x4"})) == 0
    assert gpu(np.zeros((0, 8, 8, 3), np.uint8), pal, "none").shape == (0, 8, 8, 3)


@pytest.mark.parametrize("variant", list(O.ED_KERNELS))
def test_error_diffusion_vs_oracle(variant):
    for pname, (h, w) in [("pico8", (70, 90)), ("r256", (45, 140)), ("lat27", (97, 33)),
                          ("two", (40, 40)), ("one", (9, 9))]:
        img = synth.frame(h, w, 40)
        pal = PALS[pname]
        for serp in ("false", "true"):
            p = {"variant": variant, "serpentine": serp}
            ref = O.apply_dithering(img, pal, "error_diffusion", p)
            assert mismatch(gpu(img, pal, "error_diffusion", p), ref) == 0, (variant, pname, serp)


def test_error_diffusion_multi_band_multi_frame():
    pal = PALS["r64"]
    frames = np.stack([synth.frame(150, 211, 50 + t) for t in range(3)])
    for v in ("floyd_steinberg", "jjn", "atkinson", "sierra", "burkes"):
        out = gpu(frames, pal, "error_diffusion", {"variant": v})
        for t in range(3):
            ref = O.apply_dithering(frames[t], pal, "error_diffusion", {"variant": v})
            assert mismatch(out[t], ref) == 0, (v, t)
    out = gpu(synth.noise_frame(64, 300, 5), pal, "error_diffusion", {"variant": "stucki"}, True)
    assert mismatch(out, O.apply_dithering(synth.noise_frame(64, 300, 5), pal, "error_diffusion",
                                           {"variant": "stucki"}, True)) == 0


def test_hybrid_vs_oracle():
    """Multi-band, multi-frame, single-row (serial kernel), 1080p; factor pairs incl. the
    degenerate ones (no colour error, no luminance error, plain Floyd-Steinberg)."""
    for pname, (h, w) in [("pico8", (70, 90)), ("r256", (45, 140)), ("lat27", (97, 33)),
                          ("two", (40, 40)), ("one", (9, 9)), ("c64", (1, 77)), ("r64", (150, 211))]:
        frames = np.stack([synth.frame(h, w, 90 + t) for t in range(2)])
        for params in ({}, {"lum_factor": 0.0, "col_factor": 1.5}, {"lum_factor": 1.0, "col_factor": 1.0},
                       {"lum_factor": 1.7, "col_factor": 0.0}):
            out = gpu(frames, PALS[pname], "hybrid", params)
            for t in range(2):
                ref = O.apply_dithering(frames[t], PALS[pname], "hybrid", params)
                assert mismatch(out[t], ref) == 0, (pname, h, w, params, t)
    img = synth.noise_frame(64, 300, 6)
    assert mismatch(gpu(img, PALS["r64"], "hybrid", {}, True),
                    O.apply_dithering(img, PALS["r64"], "hybrid", {}, True)) == 0
    # lum_factor = col_factor = 1 re-assembles the error only approximately (l + (e - l) rounds),
    # so it is NOT required to equal Floyd-Steinberg; the strategy-level API must agree though
    flat = img.reshape(-1, 3).astype(np.float32)
    out = dp.HybridDitherStrategy(0.8, 0.4).dither(flat, PALS["pico8"].astype(np.float32), (64, 300))
    ref = O.apply_dithering(img, PALS["pico8"], "hybrid", {"lum_factor": 0.8, "col_factor": 0.4})
    assert np.array_equal(out.reshape(64, 300, 3).astype(np.uint8), ref)
    big = synth.frame(1080, 1920, 3)
    assert mismatch(gpu(big, PALS["pico8"], "hybrid"), O.apply_dithering(big, PALS["pico8"], "hybrid", {})) == 0


def test_perceptual_vs_oracle():
    """Multi-band, multi-frame, single-row (serial kernel), values leaving the colour cube
    (two-colour and single-colour palettes push the unclamped work values far outside), 1080p."""
    for pname, (h, w) in [("pico8", (70, 90)), ("r256", (45, 140)), ("lat27", (97, 33)),
                          ("two", (40, 40)), ("one", (9, 9)), ("c64", (1, 77)), ("r64", (150, 211)),
                          ("gb4", (64, 64))]:
        frames = np.stack([synth.frame(h, w, 95 + t) for t in range(2)])
        out = gpu(frames, PALS[pname], "perceptual")
        for t in range(2):
            ref = O.apply_dithering(frames[t], PALS[pname], "perceptual", {})
            assert mismatch(out[t], ref) == 0, (pname, h, w, t)
    img = synth.noise_frame(64, 300, 6)
    assert mismatch(gpu(img, PALS["r64"], "perceptual", {}, True),
                    O.apply_dithering(img, PALS["r64"], "perceptual", {}, True)) == 0
    flat = img.reshape(-1, 3).astype(np.float32)
    out = dp.PerceptualDitherStrategy().dither(flat, PALS["pico8"].astype(np.float32), (64, 300))
    assert np.array_equal(out.reshape(64, 300, 3).astype(np.uint8),
                          O.apply_dithering(img, PALS["pico8"], "perceptual", {}))
    big = synth.frame(1080, 1920, 4)
    assert mismatch(gpu(big, PALS["pico8"], "perceptual"), O.apply_dithering(big, PALS["pico8"], "perceptual", {})) == 0


def test_adaptive_variance_vs_oracle():
    """Gate plane (running-sum filter replay, radii 1..5, lines shorter than the window, thresholds
    that are not f32 numbers) and the gated diffusion; multi-band, multi-frame, single row, 1080p."""
    cases = [{}, {"var_threshold": 0.0}, {"var_threshold": 33.3, "window_radius": 2},
             {"var_threshold": 500.1, "window_radius": 5}, {"var_threshold": 1e9}]
    for pname, (h, w) in [("pico8", (70, 90)), ("r256", (45, 140)), ("lat27", (97, 33)), ("two", (40, 40)),
                          ("one", (9, 9)), ("c64", (1, 77)), ("r64", (150, 211)), ("gb4", (3, 4))]:
        frames = np.stack([synth.frame(h, w, 97 + t) if t == 0 else synth.noise_frame(h, w, 97 + t)
                           for t in range(2)])
        for params in cases:
            out = gpu(frames, PALS[pname], "adaptive_variance", params)
            for t in range(2):
                ref = O.apply_dithering(frames[t], PALS[pname], "adaptive_variance", params)
                assert mismatch(out[t], ref) == 0, (pname, h, w, params, t)
    img = synth.noise_frame(64, 300, 6)
    assert mismatch(gpu(img, PALS["r64"], "adaptive_variance", {"var_threshold": 900.0}, True),
                    O.apply_dithering(img, PALS["r64"], "adaptive_variance", {"var_threshold": 900.0}, True)) == 0
    big = synth.frame(1080, 1920, 4)
    assert mismatch(gpu(big, PALS["pico8"], "adaptive_variance", {"var_threshold": 60.0}),
                    O.apply_dithering(big, PALS["pico8"], "adaptive_variance", {"var_threshold": 60.0})) == 0


def test_ostromoukhov_vs_oracle():
    for pname, (h, w) in [("pico8", (70, 90)), ("lat27", (66, 50)), ("r64", (40, 130)),
                          ("gb4", (33, 33))]:
        img = synth.frame(h, w, 60)
        for serp in ("false", "true"):
            ref = O.apply_dithering(img, PALS[pname], "ostromoukhov", {"serpentine": serp})
            assert mismatch(gpu(img, PALS[pname], "ostromoukhov", {"serpentine": serp}), ref) == 0, (pname, serp)
    img = synth.blocks_frame(48, 64, 2, 8, 6)
    ref = O.apply_dithering(img, PALS["lat27"], "ostromoukhov", {})
    assert mismatch(gpu(img, PALS["lat27"], "ostromoukhov"), ref) == 0


def test_halftone_vs_oracle():
    cases = [{}, {"cell_size": 3, "angle": 90.0}, {"shape": "diamond", "angle": 30.0, "cell_size": 6},
             {"shape": "square", "angle": 0.0, "dot_gain": 1.7, "sharpness": 1.0,
              "min_dot_size": 0.1, "max_dot_size": 0.9}, {"cell_size": 32, "angle": 15.0}]
    for pname in ("pico8", "lat27", "r64", "one"):
        for img in (synth.frame(70, 101, 70), synth.blocks_frame(64, 64, 71, 8, 6)):
            for params in cases:
                ref = O.apply_dithering(img, PALS[pname], "halftone", params)
                assert mismatch(gpu(img, PALS[pname], "halftone", params), ref) == 0, (pname, params)
    frames = np.stack([synth.frame(50, 60, 80 + t) for t in range(2)])
    out = gpu(frames, PALS["c64"], "halftone", {}, True)
    for t in range(2):
        assert mismatch(out[t], O.apply_dithering(frames[t], PALS["c64"], "halftone", {}, True)) == 0


def test_fused_pixelize_dither_upscale_vs_oracle_composition():
    pal = PALS["pico8"]
    frames = np.stack([synth.frame(216, 384, 1000 + t) for t in range(2)])
    for mode, params in (("blue_noise", {"size": 32, "seed": 5}), ("IGN", {}), ("none", {}),
                         ("error_diffusion", {"variant": "sierra"})):
        out = gpu(frames, pal, mode, params, pixelize_max_size=54, final_multiplier=4)
        for t in range(2):
            ref = O.final_resize(O.apply_dithering(O.pixelize_regular(frames[t], 54), pal, mode,
                                                   params), 4, True)
            assert mismatch(out[t], ref) == 0, (mode, t)
    # odd sizes: final resize bumps to even (non-multiple output) -> separate resample path
    img = synth.frame(45, 31, 3)
    out = gpu(img, pal, "bayer", {}, final_multiplier=3)
    ref = O.final_resize(O.apply_dithering(img, pal, "bayer", {}), 3, True)
    assert mismatch(out, ref) == 0


# ------------------------------------------------------------------ drop-in API level
def test_strategy_level_api_and_image_wrapper():
    from PIL import Image
    img = synth.frame(40, 56, 0)
    pal = PALS["pico8"]
    flat = img.reshape(-1, 3).astype(np.float32)
    pal_f = pal.astype(np.float32)
    for strat, mode, params in [(dp.NoDitherStrategy(), "none", {}),
                                (dp.BayerDitherStrategy("8x8"), "bayer", {"size": "8x8"}),
                                (dp.PolkaDotDitherStrategy(), "polka_dot", {}),
                                (dp.InterleavedGradientNoiseDitherStrategy(2.5, 17), "IGN",
                                 {"scale": 2.5, "seed": 17}),
                                (dp.ErrorDiffusionDitherStrategy("jjn"), "error_diffusion",
                                 {"variant": "jjn"}),
                                (dp.HalftoneDitherStrategy(), "halftone", {}),
                                (dp.MatrixDitherStrategy(engine.bayer_matrix("16x16")), "bayer",
                                 {"size": "16x16"})]:
        out = strat.dither(flat, pal_f, (40, 56))
        assert out.shape == (40 * 56, 3)
        ref = O.apply_dithering(img, pal, mode, params)
        assert np.array_equal(out.reshape(40, 56, 3).astype(np.uint8), ref), mode
    d = dp.ImageDitherer(16, DitherMode.BAYER, [tuple(int(v) for v in c) for c in pal],
                         dither_params={"size": "8x8"})
    res = d.apply_dithering(Image.fromarray(img, "RGB"))
    assert res.mode == "RGB" and np.array_equal(np.array(res), O.apply_dithering(img, pal, "bayer", {"size": "8x8"}))
    with pytest.raises(ValueError):
        dp.NoDitherStrategy().dither(flat + 0.5, pal_f, (40, 56))


def test_video_processor_frames():
    pal = [tuple(int(v) for v in c) for c in PALS["pico8"]]
    frames = np.stack([synth.frame(108, 192, 1000 + t) for t in range(5)])
    d = dp.ImageDitherer(16, DitherMode.INTERLEAVED_GRADIENT_NOISE, pal)
    seen = []
    vp = VideoProcessor(progress_callback=lambda f, m: seen.append(f))
    out = vp.process_frames(frames, d, ("regular", 54), batch_size=2, final_resize_multiplier=2)
    assert out.shape == (5, 108, 192, 3) and seen
    for t in range(5):
        ref = O.final_resize(O.apply_dithering(O.pixelize_regular(frames[t], 54), PALS["pico8"],
                                               "IGN", {}), 2, True)
        assert mismatch(out[t], ref) == 0


# ------------------------------------------------------------------ full-size checks
def reference_hash(key):
    """SHA-256 of the LIVE reference's output for a full-size BASELINE case
    (tests/golden/baseline_hashes.json, tools/make_golden.py --hashes)."""
    import json
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "baseline_hashes.json")))[key]


def sha(arr):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def test_full_size_1080p_bayer_vs_oracle_and_properties():
    img = synth.frame(1080, 1920, 0)
    pal = PALS["pico8"]
    out, idx = gpu(img, pal, "bayer", {"size": "8x8"}, return_indices=True)
    assert mismatch(out, O.apply_dithering(img, pal, "bayer", {"size": "8x8"})) == 0
    assert sha(out) == reference_hash("config1_bayer8x8")        # the reference's own bytes
    assert np.array_equal(pal[idx].astype(np.uint8), out)       # every pixel is a palette row
    near = gpu(img, pal, "none")
    assert np.array_equal(gpu(near, pal, "none"), near)          # nearest-colour is idempotent
    # ordered output is one of the two exactly-nearest rows (integer distances)
    d = ((img[:, :, None, :].astype(np.int64)[::8, ::8] - pal[None, None]) ** 2).sum(-1)
    two = np.sort(d, axis=-1)[..., 1]
    chosen = np.take_along_axis(d, idx[::8, ::8, None].astype(np.int64), -1)[..., 0]
    assert (chosen <= two).all()


def test_full_size_4k_error_diffusion_256_colours_vs_oracle():
    img = synth.frame(2160, 3840, 1)
    pal = PALS["r256"]
    out = gpu(img, pal, "error_diffusion", {"variant": "floyd_steinberg"})
    ref = O.apply_dithering(img, pal, "error_diffusion", {"variant": "floyd_steinberg"})
    assert mismatch(out, ref) == 0
    assert sha(out) == reference_hash("config2_floyd_steinberg")   # the reference's own bytes
    # a constant image whose colour is in the palette is a fixed point of error diffusion
    flat = np.broadcast_to(pal[7].astype(np.uint8), (64, 64, 3)).copy()
    assert np.array_equal(gpu(flat, pal, "error_diffusion", {"variant": "jjn"}), flat)


def test_kmeans_full_image_sharding_invariance():
    """Integer centroid sums: 1, 2 and 3 shards give bit-identical centres (single GPU, shards
    accumulated one after the other into the same sums -- what an all-reduce would produce)."""
    import ctypes as C
    from dither_pie_b200._capi import DeviceBuffer, check, lib, sync
    img = synth.frame(270, 480, 2).reshape(-1, 3)
    n = img.shape[0]
    init = img[np.random.RandomState(1).choice(n, 16, replace=False)].astype(np.float64)
    buf = DeviceBuffer(img.nbytes).upload(np.ascontiguousarray(img))
    results = []
    for shards in (1, 2, 3):
        cdev = DeviceBuffer(16 * 3 * 8).upload(init.copy())
        sums = DeviceBuffer(16 * 4 * 8)
        sh = DeviceBuffer(8)
        for _ in range(5):
            check(lib().dp_memset(sums.ptr, 0, 16 * 4 * 8, None))
            for s in range(shards):
                lo, hi = s * n // shards, (s + 1) * n // shards
                check(lib().dp_kmeans_accumulate(buf.ptr + 3 * lo, hi - lo, cdev.ptr, 16, sums.ptr, None))
            check(lib().dp_kmeans_update(sums.ptr, 16, cdev.ptr, sh.ptr, None))
        c = np.empty((16, 3), np.float64)
        cdev.download(c)
        sync()
        results.append(c)
    assert np.array_equal(results[0], results[1]) and np.array_equal(results[0], results[2])
    ref, _ = O.lloyd(img.astype(np.float64), init, -1.0, 5)
    assert np.abs(ref - results[0]).max() < 1e-9


# ------------------------------------------------------------------ randomised sweep vs the oracle
def test_random_sweep_vs_oracle():
    """Seeded random palettes / sizes / modes (both kernel families, every code path that depends
    on K, on the width and on the palette being integral)."""
    rs = np.random.RandomState(20260118)
    modes = THRESH_MODES + [("error_diffusion", {"variant": v}) for v in O.ED_KERNELS] + \
        [("error_diffusion", {"variant": "floyd_steinberg", "serpentine": "true"}),
         ("ostromoukhov", {}), ("halftone", {}), ("halftone", {"cell_size": 5, "angle": 30.0, "shape": "diamond"})]
    # (hybrid has its own test: appending it here would re-seat the seeded draws of this sweep)
    for trial in range(40):
        k = int(rs.choice([2, 3, 4, 7, 16, 29, 30, 31, 40, 100]))
        pal = synth.random_palette(k, seed=int(rs.randint(1, 10 ** 6)))
        w = int(rs.choice([16, 32, 48, 80, 37, 53]))
        h = int(rs.choice([8, 17, 33, 40]))
        kind = int(rs.randint(3))
        img = (synth.frame(h, w, trial) if kind == 0 else synth.noise_frame(h, w, trial) if kind == 1
               else synth.blocks_frame(h, w, trial, 4, 5))
        mode, params = modes[int(rs.randint(len(modes)))]
        gamma = bool(rs.randint(4) == 0)
        ref = O.apply_dithering(img, pal, mode, params, gamma)
        got = gpu(img, pal, mode, params, gamma)
        assert mismatch(got, ref) == 0, (trial, k, h, w, kind, mode, params, gamma)


def test_concurrent_callers_get_the_sequential_results():
    """The GUI calls apply_dithering from several daemon threads (dither_pie_gui.py:989-1010):
    the library, the palette/table caches and the buffer cache must be re-entrant."""
    import threading
    pal = PALS["pico8"]
    jobs = [("bayer", {"size": "8x8"}), ("none", {}), ("error_diffusion", {"variant": "atkinson"}),
            ("halftone", {}), ("IGN", {}), ("ostromoukhov", {})]
    imgs = [synth.frame(96, 160, 200 + i) for i in range(len(jobs))]
    want = [gpu(img, pal, m, p) for img, (m, p) in zip(imgs, jobs)]
    got = [None] * len(jobs)
    errs = []

    def work(i):
        try:
            for _ in range(5):
                got[i] = gpu(imgs[i], pal, *jobs[i])
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for i in range(len(jobs)):
        assert np.array_equal(got[i], want[i]), jobs[i]
