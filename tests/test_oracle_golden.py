"""The oracle (oracle/) against the golden vectors generated from the live reference
(tools/make_golden.py).  CPU only.  Bit-exact everywhere except k-means (1e-3, the tolerance
BASELINE.json's north_star states)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import dither_oracle as O


def test_versions_match_golden():
    """Results depend on the numeric libraries; flag drift loudly instead of failing mysteriously."""
    import PIL
    import scipy
    import sklearn
    v = json.load(open(os.path.join(GOLDEN, "VERSIONS.json")))
    assert v["numba_path_used"] is True
    assert scipy.__version__ == v["scipy"]
    assert np.__version__ == v["numpy"]
    assert PIL.__version__ == v["Pillow"]
    assert sklearn.__version__ == v["scikit-learn"]


def test_threshold_sources():
    g = load_golden("threshold_sources.npz")
    for s in ("2x2", "4x4", "8x8", "16x16", "psx4x4"):
        assert np.array_equal(O.bayer_matrix(s), g["bayer_" + s]), s
    assert np.array_equal(O.bayer_matrix("psx"), g["bayer_psx4x4"])
    assert np.array_equal(O.bayer_matrix("bogus"), g["bayer_4x4"])
    assert np.array_equal(O.blue_noise_matrix(64, 42), g["blue_64_42"])
    assert np.array_equal(O.blue_noise_matrix(32, 5), g["blue_32_5"])
    assert np.array_equal(O.polka_dot_matrix(8, 1.5), g["polka_8_1.5"])
    assert np.array_equal(O.polka_dot_matrix(5, 0.7), g["polka_5_0.7"])
    assert np.array_equal(O.ign_thresholds(33, 47, 1.0, 0), g["ign_1_0"])
    assert np.array_equal(O.ign_thresholds(33, 47, 2.5, 17), g["ign_2.5_17"])
    assert np.array_equal(O.ostromoukhov_coeffs(), g["ostro_coeffs"])
    assert np.array_equal(O.halftone_screen(40, 56)[0], g["halftone_default_screen"])
    assert np.array_equal(O.halftone_screen(40, 56, cell_size=6, angle=30.0, shape="diamond")[0],
                          g["halftone_diamond_screen"])


def test_kdtree_c_restatement_matches_scipy_golden():
    g = load_golden("kdtree_queries.npz")
    t = 0
    while f"pal_{t}" in g:
        pal = g[f"pal_{t}"]
        pts = g[f"pts_{t}"].astype(np.float64)
        tree = O.export_kdtree(pal)
        d2a, i1 = O.kdtree_query_c(tree, pts, 1)
        d2b, i2 = O.kdtree_query_c(tree, pts, 2)
        assert np.array_equal(i1[:, 0], g[f"i1_{t}"]), t
        assert np.array_equal(i2, g[f"i2_{t}"]), t
        assert np.array_equal(np.sqrt(d2b), g[f"d2_{t}"]), t
        t += 1
    assert t >= 10


def test_dither_cases_bit_exact(golden_cases):
    data, meta = golden_cases
    bad = []
    for n, m in enumerate(meta):
        out = O.apply_dithering(data["img_" + m["image"]], data["pal_" + m["palette"]],
                                m["mode"], m["params"], m["gamma"])
        if not np.array_equal(out, data[f"out_{n}"]):
            bad.append((n, m))
    assert not bad, bad[:5]


def test_diffusion_big():
    g = load_golden("diffusion_big.npz")
    for v in ("floyd_steinberg", "atkinson", "jjn", "sierra"):
        out = O.apply_dithering(g["img"], g["pal"], "error_diffusion", {"variant": v})
        assert np.array_equal(out, g["ed_" + v]), v
    out = O.apply_dithering(g["img"], g["pal16"], "ostromoukhov", {})
    assert np.array_equal(out, g["ostro"])


def test_hybrid_cases_bit_exact():
    """hybrid mode (_hybrid_numba, dithering_lib.py:1396-1494) incl. gamma and extreme factors."""
    g = load_golden("hybrid_cases.npz")
    meta = json.load(open(os.path.join(GOLDEN, "hybrid_cases.json")))
    assert len(meta) >= 14
    for n, m in enumerate(meta):
        out = O.apply_dithering(g["img_" + m["image"]], g["pal_" + m["palette"]], "hybrid",
                                m["params"], m["gamma"])
        assert np.array_equal(out, g[f"out_{n}"]), (n, m)


def test_perceptual_cases_bit_exact():
    """perceptual mode (pure-Python reference, :1030-1066): unclamped KD-tree lookups, f32."""
    g = load_golden("perceptual_cases.npz")
    meta = json.load(open(os.path.join(GOLDEN, "perceptual_cases.json")))
    assert len(meta) >= 14
    for n, m in enumerate(meta):
        out = O.apply_dithering(g["img_" + m["image"]], g["pal_" + m["palette"]], "perceptual", {}, m["gamma"])
        assert np.array_equal(out, g[f"out_{n}"]), (n, m)


def test_adaptive_cases_bit_exact():
    """adaptive_variance mode (pure-Python reference, :946-1025) incl. gamma, radii 1..5,
    thresholds that are not f32 numbers."""
    g = load_golden("adaptive_cases.npz")
    meta = json.load(open(os.path.join(GOLDEN, "adaptive_cases.json")))
    assert len(meta) >= 14
    for n, m in enumerate(meta):
        out = O.apply_dithering(g["img_" + m["image"]], g["pal_" + m["palette"]], "adaptive_variance",
                                m["params"], m["gamma"])
        assert np.array_equal(out, g[f"out_{n}"]), (n, m)


def test_uniform_filter_restatement_equals_scipy():
    """The running-sum restatement of scipy.ndimage.uniform_filter (the arithmetic the CUDA gate
    kernel replays) against scipy itself, bit for bit."""
    from scipy.ndimage import uniform_filter
    rs = np.random.RandomState(3)
    for (h, w) in ((1, 1), (1, 9), (7, 1), (13, 17), (40, 56), (3, 200)):
        for scale in (1.0, 255.0, 65025.0, 1e-3):
            a = (rs.rand(h, w) * scale).astype(np.float32)
            for size in (1, 3, 5, 7, 11):
                assert np.array_equal(O.uniform_filter_nearest(a, size),
                                      uniform_filter(a, size=size, mode='nearest')), (h, w, scale, size)


def test_pixelize_tables():
    g = load_golden("pixelize.npz")
    meta = json.load(open(os.path.join(GOLDEN, "pixelize.json")))
    for m in meta:
        tw, th = O.even_dimensions(m["w"], m["h"], m["max_size"])
        assert (tw, th) == (m["tw"], m["th"]), m
        key = f"{m['w']}x{m['h']}_{m['max_size']}"
        assert np.array_equal(O.nearest_table(m["w"], tw), g["xt_" + key]), key
        assert np.array_equal(O.nearest_table(m["h"], th), g["yt_" + key]), key
    small = g["small"]
    for mult in (2, 3, 5):
        assert np.array_equal(O.final_resize(small, mult, True), g[f"small_x{mult}_even"])
        assert np.array_equal(O.final_resize(small, mult, False), g[f"small_x{mult}_cli"])


def test_kmeans_centers():
    g = load_golden("kmeans.npz")
    for t, k in enumerate((16, 8, 5)):
        c = O.kmeans_centers(g[f"sample_{t}"], k, 42)
        assert np.abs(c - g[f"centers_{t}"]).max() <= 1e-3  # north_star tolerance
        assert np.array_equal(c.astype(int), g[f"palette_{t}"])


@pytest.mark.parametrize("k", [1, 2])
def test_kdtree_single_colour_palette(k):
    pal = np.array([[10, 20, 30]], np.float32)
    d2, idx = O.kdtree_query_c(O.export_kdtree(pal), np.array([[1.0, 2.0, 3.0]]), k)
    assert idx[0, 0] == 0 and d2[0, 0] == 81 + 324 + 729
    if k == 2:
        assert idx[0, 1] == 1 and np.isinf(d2[0, 1])  # scipy: index n, distance inf


def test_big_cases_from_the_live_reference():
    """540x960 outputs of the reference itself (tools/make_golden.py --big): the oracle pinned at
    the BASELINE configs' kernels and palette sizes, beyond the small cases."""
    g = load_golden("big_cases.npz")
    img = g["img"]
    for key, pk, mode, params in (("ed_jjn", "pal256", "error_diffusion", {"variant": "jjn"}),
                                  ("ed_atkinson", "pal256", "error_diffusion", {"variant": "atkinson"}),
                                  ("ed_sierra", "pal64", "error_diffusion", {"variant": "sierra"}),
                                  ("bayer8", "pal16", "bayer", {"size": "8x8"}),
                                  ("blue", "pal16", "blue_noise", {"size": 64, "seed": 42})):
        assert np.array_equal(O.apply_dithering(img, g[pk], mode, params), g[key]), key
    assert np.array_equal(O.apply_dithering(g["ostro_img"], g["pal64"], "ostromoukhov"), g["ostro"])


BIG2 = [("none16", "pal16", "none", {}, None), ("none256", "pal256", "none", {}, None),
        ("bayer8_256", "pal256", "bayer", {"size": "8x8"}, None),
        ("bayer16_64", "pal64", "bayer", {"size": "16x16"}, None),
        ("ign16", "pal16", "IGN", {"scale": 1.0, "seed": 0}, None),
        ("ign256", "pal256", "IGN", {"scale": 2.5, "seed": 17}, None),
        ("blue256", "pal256", "blue_noise", {"size": 64, "seed": 42}, None),
        ("polka16", "pal16", "polka_dot", {"tile_size": 8, "gamma": 1.5}, None),
        ("halftone16", "pal16", "halftone", {}, None),
        ("halftone64", "pal64", "halftone", {"cell_size": 5, "angle": 30.0, "shape": "diamond"}, None),
        ("ed_fs256", "pal256", "error_diffusion", {"variant": "floyd_steinberg"}, None),
        ("ed_stucki64", "pal64", "error_diffusion", {"variant": "stucki"}, None),
        ("ed_burkes16", "pal16", "error_diffusion", {"variant": "burkes"}, None),
        ("ed_two_row64", "pal64", "error_diffusion", {"variant": "sierra_two_row"}, None),
        ("ed_lite256", "pal256", "error_diffusion", {"variant": "sierra_lite"}, None),
        ("ed_fs64_serp", "pal64", "error_diffusion", {"variant": "floyd_steinberg", "serpentine": "true"}, (270, 480)),
        ("hybrid64", "pal64", "hybrid", {}, None)]


def test_big_cases2_from_the_live_reference():
    """The rest of the mode list at 540x960 from the reference itself (tools/make_golden.py --big2):
    nearest colour, Bayer / IGN / blue noise / polka dot at 16-256 colours, halftone, the other
    diffusion kernels, serpentine, hybrid -- stored as palette-index planes."""
    g = load_golden("big_cases2.npz")
    img = g["img"]
    for key, pk, mode, params, crop in BIG2:
        arr = img if crop is None else np.ascontiguousarray(img[:crop[0], :crop[1]])
        out = O.apply_dithering(arr, g[pk], mode, params)
        assert np.array_equal(out, np.asarray(g[pk], np.uint8)[g[key]]), key


def test_big_cases3_perceptual_and_adaptive_variance_from_the_live_reference():
    """The reference's two pure-Python diffusion modes on a 270x480 crop (tools/make_golden.py
    --big3): nine row bands, unclamped look-ups, the variance map."""
    g = load_golden("big_cases3.npz")
    img = g["img"]
    for key, pk, mode, params in (("perceptual64", "pal64", "perceptual", {}),
                                  ("perceptual16", "pal16", "perceptual", {}),
                                  ("adaptive16", "pal16", "adaptive_variance", {}),
                                  ("adaptive64", "pal64", "adaptive_variance", {"var_threshold": 60.0})):
        out = O.apply_dithering(img, g[pk], mode, params)
        assert np.array_equal(out, np.asarray(g[pk], np.uint8)[g[key]]), key


def test_kmeans_config3_4k_frame_from_the_live_reference():
    """BASELINE configs[2]: the reference's own k-means palette of the seed-2 4K frame
    (random.seed(7) subsample, k=16, random_state=42; tools/make_golden.py --kmeans4k)."""
    import random
    from dither_pie_b200 import synth
    g = load_golden("kmeans_4k.npz")
    flat = synth.frame(2160, 3840, 2).reshape(-1, 3)
    random.seed(7)
    sample = flat[random.sample(range(len(flat)), 10000)]
    assert np.array_equal(sample, g["sample"])          # the subsample is the reference's
    c = O.kmeans_centers(sample, 16, 42)
    assert np.abs(c - g["centers"]).max() <= 1e-3       # north_star tolerance
    assert np.array_equal(c.astype(int), g["palette"])


def test_config4_as_the_reference_cli_runs_it():
    """BASELINE configs[3] from the live reference (tools/make_golden.py --config4): palette by
    median cut from the FULL first 1080p frame (dither_cli.py:619-654), then pixelize 270 -> blue
    noise (64, 42) / IGN (1.0, 0) with that palette -> x4.  The drop-in ColorReducer (host code of the
    library) must find the same palette, the oracle the same frames."""
    from PIL import Image
    import dither_pie_b200 as dp
    from dither_pie_b200 import synth
    g = load_golden("config4.npz")
    frames = [synth.frame(1080, 1920, 1000 + t) for t in range(2)]
    pal = dp.ColorReducer.reduce_colors(Image.fromarray(frames[0], "RGB"), 16)
    assert [tuple(int(v) for v in c) for c in pal] == [tuple(int(v) for v in c) for c in g["palette"]]
    pal_u8 = np.asarray(g["palette"], np.uint8)
    for name, mode, params in (("blue", "blue_noise", {"size": 64, "seed": 42}),
                               ("ign", "IGN", {"scale": 1.0, "seed": 0})):
        for t, f in enumerate(frames):
            small = O.pixelize_regular(f, 270)
            out = O.apply_dithering(small, g["palette"], mode, params)
            assert np.array_equal(out, pal_u8[g[f"{name}_{t}"]]), (name, t)
            if name == "blue" and t == 1:
                assert np.array_equal(O.final_resize(out, 4, True), g["blue_1_x4"])


def test_baseline_configs_at_full_size_hashes_of_the_live_reference():
    """SHA-256 of the reference's own outputs at the BASELINE configs' FULL sizes (tools/make_golden.py
    --hashes): 4K Floyd-Steinberg / Atkinson / JJN with 256 colours (configs[1]), 4K Sierra with 64
    colours (configs[4]), 1080p Bayer 8x8 with 16 colours (configs[0]), 4K nearest colour with config
    3's k-means palette, and the threshold family / nearest colour at 1080p.  The oracle must
    produce the same bytes; the GPU tests compare the CUDA path with the oracle at these sizes."""
    import hashlib
    import json
    from concurrent.futures import ThreadPoolExecutor
    from dither_pie_b200 import synth
    want = json.load(open(os.path.join(GOLDEN, "baseline_hashes.json")))
    img4k = synth.frame(2160, 3840, 1)
    jobs = {f"config2_{v}": (img4k, synth.random_palette(256), "error_diffusion", {"variant": v})
            for v in ("floyd_steinberg", "atkinson", "jjn")}
    jobs["config5_sierra_frame2000"] = (synth.frame(2160, 3840, 2000), synth.random_palette(64),
                                        "error_diffusion", {"variant": "sierra"})
    f1080, pico, r256 = synth.frame(1080, 1920, 0), synth.hex_palette(synth.PICO8), synth.random_palette(256)
    jobs["config1_bayer8x8"] = (f1080, pico, "bayer", {"size": "8x8"})
    jobs["1080p_halftone_pico8"] = (f1080, pico, "halftone", {})
    jobs["1080p_ign_pico8"] = (f1080, pico, "IGN", {"scale": 1.0, "seed": 0})
    jobs["1080p_blue_noise_pico8"] = (f1080, pico, "blue_noise", {"size": 64, "seed": 42})
    jobs["1080p_none_pico8"] = (f1080, pico, "none", {})
    jobs["1080p_none_r256"] = (f1080, r256, "none", {})
    jobs["1080p_bayer8x8_r256"] = (f1080, r256, "bayer", {"size": "8x8"})
    jobs["config3_4k_none_kmeans_palette"] = (synth.frame(2160, 3840, 2), load_golden("kmeans_4k.npz")["palette"],
                                              "none", {})
    f4k = synth.frame(2160, 3840, 2001)
    for v in ("stucki", "burkes", "sierra_two_row", "sierra_lite"):
        jobs[f"4k_{v}_r64"] = (f4k, synth.random_palette(64), "error_diffusion", {"variant": v})
    jobs["1080p_fs_serpentine_r64"] = (f1080, synth.random_palette(64), "error_diffusion",
                                       {"variant": "floyd_steinberg", "serpentine": "true"})
    jobs["540p_ostromoukhov_r64"] = (np.ascontiguousarray(synth.frame(1080, 1920, 6)[:540, :960]),
                                     synth.random_palette(64), "ostromoukhov", {})
    g540, r64 = np.ascontiguousarray(synth.frame(1080, 1920, 5)[:540, :960]), synth.random_palette(64)
    jobs["gamma_540p_bayer8x8_pico8"] = (g540, pico, "bayer", {"size": "8x8"}, True)      # use_gamma=True
    jobs["gamma_540p_none_r64"] = (g540, r64, "none", {}, True)
    jobs["gamma_540p_halftone_pico8"] = (g540, pico, "halftone", {}, True)
    jobs["gamma_540p_fs_r64"] = (g540, r64, "error_diffusion", {"variant": "floyd_steinberg"}, True)
    jobs["gamma_540p_jjn_pico8"] = (g540, pico, "error_diffusion", {"variant": "jjn"}, True)
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:   # the C port releases the GIL
        got = dict(zip(jobs, ex.map(lambda j: hashlib.sha256(
            np.ascontiguousarray(O.apply_dithering(*j)).tobytes()).hexdigest(), jobs.values())))
    assert got == want
