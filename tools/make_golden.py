"""Generate the golden vectors under tests/golden/ from the LIVE reference.

    python tools/make_golden.py

Runs only in the build container (needs /root/reference).  The reference ships no tests or
fixtures of its own, so these files -- outputs of the unmodified reference
on deterministic synthetic inputs -- are what pins both the oracle and the CUDA path.
Library versions are recorded in tests/golden/VERSIONS.json because the results depend on them.
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.ref_loader import load  # noqa: E402
from dither_pie_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def images():
    return {
        "frame": synth.frame(40, 56, 0),
        "noise": synth.noise_frame(36, 48, 1),
        "blocks": synth.blocks_frame(40, 56, 2, 8, 6),
    }


def palettes():
    return {
        "pico8": synth.hex_palette(synth.PICO8),
        "c64": synth.hex_palette(synth.C64),
        "gb4": synth.hex_palette(synth.GB_POCKET),
        "r16": synth.random_palette(16),
        "r64": synth.random_palette(64),
        "r256": synth.random_palette(256),
        "lat27": synth.lattice_palette(27, 1, 127),
        "one": np.array([[12, 200, 77]]),
    }


MODES = [
    ("none", {}),
    ("bayer", {"size": "2x2"}), ("bayer", {"size": "4x4"}), ("bayer", {"size": "8x8"}),
    ("bayer", {"size": "16x16"}), ("bayer", {"size": "psx4x4"}),
    ("blue_noise", {"size": 32, "seed": 5}),
    ("IGN", {}), ("IGN", {"scale": 2.5, "seed": 17}),
    ("polka_dot", {}), ("polka_dot", {"tile_size": 5, "gamma": 0.7}),
    ("halftone", {}), ("halftone", {"cell_size": 3, "angle": 90.0}),
    ("halftone", {"shape": "diamond", "angle": 30.0, "cell_size": 6}),
    ("halftone", {"shape": "square", "angle": 0.0, "dot_gain": 1.7, "sharpness": 1.0,
                  "min_dot_size": 0.1, "max_dot_size": 0.9}),
    ("error_diffusion", {}),
    ("error_diffusion", {"variant": "floyd_steinberg"}),
    ("error_diffusion", {"variant": "jjn"}),
    ("error_diffusion", {"variant": "stucki"}),
    ("error_diffusion", {"variant": "burkes"}),
    ("error_diffusion", {"variant": "sierra"}),
    ("error_diffusion", {"variant": "sierra_two_row"}),
    ("error_diffusion", {"variant": "sierra_lite"}),
    ("error_diffusion", {"variant": "jjn", "serpentine": "true"}),
    ("error_diffusion", {"variant": "floyd_steinberg", "serpentine": "true"}),
    ("ostromoukhov", {}),
    ("ostromoukhov", {"serpentine": "true"}),
]


def case_list():
    """(image, palette, mode, params, gamma) -- a spread that keeps the file small."""
    cases = []
    for iname in ("frame", "noise", "blocks"):
        for pname in ("pico8", "c64", "gb4", "r16", "r64", "r256", "lat27", "one"):
            for mi, (mode, params) in enumerate(MODES):
                # thin the cross product: every mode on 'frame'; a rotating subset elsewhere
                if iname != "frame" and (mi + len(pname)) % 3 != 0 and pname not in ("pico8", "lat27"):
                    continue
                if mode == "ostromoukhov" and pname not in ("pico8", "lat27", "r64", "one"):
                    continue
                cases.append((iname, pname, mode, params, False))
                if pname in ("pico8", "r64") and iname == "frame" and \
                        params.get("size", "8x8") == "8x8" and params.get("variant", "jjn") == "jjn":
                    cases.append((iname, pname, mode, params, True))
    return cases


def main():
    dl, vp = load()
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["hybrid"]:
        return hybrid_golden(dl)
    if sys.argv[1:] == ["perceptual"]:
        return perceptual_golden(dl)
    if sys.argv[1:] == ["adaptive"]:
        return adaptive_golden(dl)
    import numba
    import PIL
    import scipy
    import sklearn
    versions = {"numpy": np.__version__, "scipy": scipy.__version__,
                "scikit-learn": sklearn.__version__, "Pillow": PIL.__version__,
                "numba": numba.__version__, "numba_path_used": bool(dl._NUMBA_AVAILABLE)}
    json.dump(versions, open(os.path.join(OUT, "VERSIONS.json"), "w"), indent=1)

    imgs, pals = images(), palettes()

    # ---- 1. whole-image dithering through ImageDitherer.apply_dithering
    store = {}
    meta = []
    for n, (iname, pname, mode, params, gamma) in enumerate(case_list()):
        d = dl.ImageDitherer(num_colors=len(pals[pname]), dither_mode=dl.DitherMode(mode),
                             palette=[tuple(int(v) for v in c) for c in pals[pname]],
                             use_gamma=gamma, dither_params=dict(params))
        out = np.array(d.apply_dithering(Image.fromarray(imgs[iname], "RGB")))
        store[f"out_{n}"] = out
        meta.append({"image": iname, "palette": pname, "mode": mode, "params": params,
                     "gamma": gamma})
    for k, v in imgs.items():
        store[f"img_{k}"] = v
    for k, v in pals.items():
        store[f"pal_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "dither_cases.npz"), **store)
    json.dump(meta, open(os.path.join(OUT, "dither_cases.json"), "w"))
    print("dither cases:", len(meta))

    # ---- 2. a larger error-diffusion / ostromoukhov image (long-range error propagation)
    big = synth.frame(96, 160, 5)
    store = {"img": big, "pal": pals["r64"], "pal16": pals["pico8"]}
    for variant in ("floyd_steinberg", "atkinson", "jjn", "sierra"):
        d = dl.ImageDitherer(dither_mode=dl.DitherMode.ERROR_DIFFUSION,
                             palette=[tuple(int(v) for v in c) for c in pals["r64"]],
                             dither_params={"variant": variant})
        store[f"ed_{variant}"] = np.array(d.apply_dithering(Image.fromarray(big, "RGB")))
    d = dl.ImageDitherer(dither_mode=dl.DitherMode.OSTROMOUKHOV,
                         palette=[tuple(int(v) for v in c) for c in pals["pico8"]])
    store["ostro"] = np.array(d.apply_dithering(Image.fromarray(big, "RGB")))
    np.savez_compressed(os.path.join(OUT, "diffusion_big.npz"), **store)

    # ---- 3. threshold sources
    U = dl.DitherUtils
    store = {"bayer_2x2": U.BAYER2x2, "bayer_4x4": U.BAYER4x4, "bayer_8x8": U.BAYER8x8,
             "bayer_16x16": U.BAYER16x16, "bayer_psx4x4": U.PSX4x4,
             "blue_64_42": dl.generate_blue_noise(64, 42),
             "blue_32_5": dl.generate_blue_noise(32, 5),
             "polka_8_1.5": dl.PolkaDotDitherStrategy(8, 1.5).threshold_matrix,
             "polka_5_0.7": dl.PolkaDotDitherStrategy(5, 0.7).threshold_matrix,
             "ign_1_0": dl.InterleavedGradientNoiseDitherStrategy(1.0, 0)._generate_thresholds((33, 47)),
             "ign_2.5_17": dl.InterleavedGradientNoiseDitherStrategy(2.5, 17)._generate_thresholds((33, 47)),
             "ostro_coeffs": np.asarray(dl.OstromoukhovDitherStrategy.COEFFS_TABLE, np.int32)}
    h = dl.HalftoneDitherStrategy()
    s, c = h._generate_halftone_screen_with_cells(40, 56)
    store["halftone_default_screen"] = s
    h = dl.HalftoneDitherStrategy(cell_size=6, angle=30.0, shape="diamond")
    s, c = h._generate_halftone_screen_with_cells(40, 56)
    store["halftone_diamond_screen"] = s
    np.savez_compressed(os.path.join(OUT, "threshold_sources.npz"), **store)

    # ---- 4. scipy KD-tree queries on tie-heavy lattices (k=1 and k=2)
    from scipy.spatial import KDTree
    store = {}
    kd_meta = []
    for t, (k_pal, step_pal, step_pts) in enumerate(
            [(4, 85, 17), (8, 51, 51), (10, 51, 17), (11, 51, 17), (16, 51, 17), (20, 51, 51),
             (27, 127, 1), (40, 17, 17), (64, 51, 17), (100, 17, 17), (256, 17, 17)]):
        pal = synth.lattice_palette(k_pal, seed=t, step=step_pal).astype(np.float32)
        rs = np.random.RandomState(100 + t)
        pts = (rs.randint(0, 255 // step_pts + 1, (600, 3)) * step_pts).astype(np.float64)
        tree = KDTree(pal)
        d1, i1 = tree.query(pts, k=1)
        d2, i2 = tree.query(pts, k=2)
        store[f"pal_{t}"] = pal
        store[f"pts_{t}"] = pts.astype(np.uint8)
        store[f"i1_{t}"] = i1.astype(np.int32)
        store[f"i2_{t}"] = i2.astype(np.int32)
        store[f"d2_{t}"] = d2
        kd_meta.append(t)
    np.savez_compressed(os.path.join(OUT, "kdtree_queries.npz"), **store)

    # ---- 5. Pillow NEAREST tables + pixelize sizes
    store = {}
    pix_meta = []
    for (w, hh) in [(1920, 1080), (3840, 2160), (640, 480), (333, 517), (1000, 999), (57, 31),
                    (720, 1280), (100, 100)]:
        # an image whose red/green channels encode x mod 256 / y mod 256 and blue the high bits
        xs = np.arange(w)[None, :].repeat(hh, 0)
        ys = np.arange(hh)[:, None].repeat(w, 1)
        img = np.stack([xs % 256, ys % 256, (xs // 256) * 16 + (ys // 256)], 2).astype(np.uint8)
        for ms in (128, 270, 64, 33, 17, 7):
            out = np.array(vp.pixelize_regular(Image.fromarray(img, "RGB"), ms))
            xt = out[0, :, 0].astype(np.int32) + 256 * (out[0, :, 2].astype(np.int32) // 16)
            yt = out[:, 0, 1].astype(np.int32) + 256 * (out[:, 0, 2].astype(np.int32) % 16)
            key = f"{w}x{hh}_{ms}"
            store["xt_" + key] = xt
            store["yt_" + key] = yt
            pix_meta.append({"w": w, "h": hh, "max_size": ms, "tw": int(out.shape[1]),
                             "th": int(out.shape[0])})
    small = synth.noise_frame(31, 45, 9)
    store["small"] = small
    for m in (2, 3, 5):
        store[f"small_x{m}_even"] = np.array(
            vp._apply_final_resize_to_frame(Image.fromarray(small, "RGB"), m))
        store[f"small_x{m}_cli"] = np.array(
            Image.fromarray(small, "RGB").resize((45 * m, 31 * m), Image.Resampling.NEAREST))
    np.savez_compressed(os.path.join(OUT, "pixelize.npz"), **store)
    json.dump(pix_meta, open(os.path.join(OUT, "pixelize.json"), "w"))

    # ---- 6. k-means (sklearn) on fixed sub-samples
    from sklearn.cluster import KMeans
    store = {}
    for t, (hh, w, k) in enumerate([(1080, 1920, 16), (300, 400, 8), (90, 100, 5)]):
        img = synth.frame(hh, w, 2 + t)
        pix = img.reshape(-1, 3)
        if len(pix) > 10000:
            random.seed(7)
            pix = pix[random.sample(range(len(pix)), 10000)]
        km = KMeans(n_clusters=k, random_state=42).fit(pix)
        random.seed(7)
        pal = dl.ColorReducer.generate_kmeans_palette(Image.fromarray(img, "RGB"), k, 42)
        store[f"sample_{t}"] = pix
        store[f"centers_{t}"] = km.cluster_centers_
        store[f"palette_{t}"] = np.asarray(pal, np.int64)
        store[f"niter_{t}"] = np.asarray(km.n_iter_)
    np.savez_compressed(os.path.join(OUT, "kmeans.npz"), **store)
    median_cut_golden(dl)
    hybrid_golden(dl)
    perceptual_golden(dl)
    adaptive_golden(dl)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def hybrid_golden(dl):
    """---- 8. hybrid mode (HybridDitherStrategy :1071-1155 -> _hybrid_numba :1396-1494) through
    ImageDitherer.apply_dithering; own file so that it can be regenerated alone
    (python tools/make_golden.py hybrid)."""
    assert dl._NUMBA_AVAILABLE
    imgs, pals = images(), palettes()
    imgs["big"] = synth.frame(96, 160, 5)
    store, meta = {}, []
    n = 0
    for iname, pname, params, gamma in [
            ("frame", "pico8", {}, False), ("frame", "pico8", {}, True),
            ("frame", "r64", {"lum_factor": 0.7, "col_factor": 0.9}, False),
            ("frame", "r256", {"lum_factor": 1.3, "col_factor": 0.0}, False),
            ("frame", "gb4", {"lum_factor": 0.0, "col_factor": 2.0}, False),
            ("frame", "lat27", {}, False), ("frame", "one", {}, False),
            ("noise", "c64", {}, False), ("noise", "r16", {"lum_factor": 2.0, "col_factor": 1.0}, False),
            ("noise", "r64", {}, True), ("blocks", "pico8", {}, False),
            ("blocks", "lat27", {"lum_factor": 1.0, "col_factor": 1.0}, False),
            ("big", "r64", {}, False), ("big", "pico8", {"lum_factor": 0.5, "col_factor": 0.5}, False)]:
        d = dl.ImageDitherer(num_colors=len(pals[pname]), dither_mode=dl.DitherMode.HYBRID,
                             palette=[tuple(int(v) for v in c) for c in pals[pname]],
                             use_gamma=gamma, dither_params=dict(params))
        store[f"out_{n}"] = np.array(d.apply_dithering(Image.fromarray(imgs[iname], "RGB")))
        meta.append({"image": iname, "palette": pname, "params": params, "gamma": gamma})
        n += 1
    for k, v in imgs.items():
        store[f"img_{k}"] = v
    for k, v in pals.items():
        store[f"pal_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "hybrid_cases.npz"), **store)
    json.dump(meta, open(os.path.join(OUT, "hybrid_cases.json"), "w"))
    print("hybrid cases:", len(meta))


def perceptual_golden(dl):
    """---- 9. perceptual mode (PerceptualDitherStrategy :1030-1066, pure Python + KDTree) through
    ImageDitherer.apply_dithering; own file (python tools/make_golden.py perceptual)."""
    imgs, pals = images(), palettes()
    imgs["wide"] = synth.frame(70, 90, 5)      # three row bands
    store, meta = {}, []
    n = 0
    for iname, pname, gamma in [
            ("frame", "pico8", False), ("frame", "pico8", True), ("frame", "r64", False),
            ("frame", "r256", False), ("frame", "gb4", False), ("frame", "lat27", False),
            ("frame", "one", False), ("noise", "c64", False), ("noise", "r16", False),
            ("noise", "r64", True), ("blocks", "pico8", False), ("blocks", "lat27", False),
            ("wide", "r64", False), ("wide", "gb4", False)]:
        d = dl.ImageDitherer(num_colors=len(pals[pname]), dither_mode=dl.DitherMode.PERCEPTUAL,
                             palette=[tuple(int(v) for v in c) for c in pals[pname]], use_gamma=gamma)
        store[f"out_{n}"] = np.array(d.apply_dithering(Image.fromarray(imgs[iname], "RGB")))
        meta.append({"image": iname, "palette": pname, "gamma": gamma})
        n += 1
    for k, v in imgs.items():
        store[f"img_{k}"] = v
    for k, v in pals.items():
        store[f"pal_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "perceptual_cases.npz"), **store)
    json.dump(meta, open(os.path.join(OUT, "perceptual_cases.json"), "w"))
    print("perceptual cases:", len(meta))


def adaptive_golden(dl):
    """---- 10. adaptive_variance mode (AdaptiveVarianceDitherStrategy :946-1025, pure Python +
    KDTree + scipy.ndimage.uniform_filter); own file (python tools/make_golden.py adaptive)."""
    imgs, pals = images(), palettes()
    imgs["wide"] = synth.frame(70, 90, 5)
    store, meta = {}, []
    n = 0
    for iname, pname, params, gamma in [
            ("frame", "pico8", {}, False), ("frame", "pico8", {}, True),
            ("frame", "r64", {"var_threshold": 50.0}, False),
            ("frame", "r256", {"var_threshold": 0.0}, False),
            ("frame", "gb4", {"var_threshold": 120.5, "window_radius": 2}, False),
            ("frame", "lat27", {"var_threshold": 77.7, "window_radius": 5}, False),
            ("frame", "one", {}, False), ("noise", "c64", {}, False),
            ("noise", "r16", {"var_threshold": 2500.0}, False),
            ("noise", "r64", {"var_threshold": 4000.3, "window_radius": 3}, True),
            ("blocks", "pico8", {"var_threshold": 10.0}, False),
            ("blocks", "lat27", {"var_threshold": 1000.0, "window_radius": 4}, False),
            ("wide", "r64", {"var_threshold": 60.0}, False), ("wide", "gb4", {"var_threshold": 0.1}, False)]:
        d = dl.ImageDitherer(num_colors=len(pals[pname]), dither_mode=dl.DitherMode.ADAPTIVE_VARIANCE,
                             palette=[tuple(int(v) for v in c) for c in pals[pname]], use_gamma=gamma,
                             dither_params=dict(params))
        store[f"out_{n}"] = np.array(d.apply_dithering(Image.fromarray(imgs[iname], "RGB")))
        meta.append({"image": iname, "palette": pname, "params": params, "gamma": gamma})
        n += 1
    for k, v in imgs.items():
        store[f"img_{k}"] = v
    for k, v in pals.items():
        store[f"pal_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "adaptive_cases.npz"), **store)
    json.dump(meta, open(os.path.join(OUT, "adaptive_cases.json"), "w"))
    print("adaptive cases:", len(meta))


def median_cut_golden(dl):
    """---- 7. median cut / uniform palettes (ColorReducer.reduce_colors :1834-1843,
    generate_uniform_palette :1859-1872) on seeded frames; small, so stored as JSON."""
    cases = []
    for (kind, hh, w, seed, nc) in [("frame", 120, 160, 5, 16), ("frame", 120, 160, 5, 12),
                                    ("noise", 60, 80, 6, 8), ("noise", 60, 80, 6, 2),
                                    ("blocks", 64, 64, 7, 16), ("frame", 7, 9, 8, 256),
                                    ("frame", 120, 160, 9, 1)]:
        img = {"frame": synth.frame, "noise": synth.noise_frame}.get(kind, None)
        arr = img(hh, w, seed) if img else synth.blocks_frame(hh, w, seed, 8, 6)
        pal = dl.ColorReducer.reduce_colors(Image.fromarray(arr, "RGB"), nc)
        cases.append({"kind": kind, "h": hh, "w": w, "seed": seed, "num_colors": nc,
                      "palette": [list(map(int, c)) for c in pal]})
    uniform = {str(n): [list(map(int, c)) for c in dl.ColorReducer.generate_uniform_palette(n)]
               for n in (1, 2, 8, 16, 27, 30)}
    json.dump({"median_cut": cases, "uniform": uniform},
              open(os.path.join(OUT, "median_cut.json"), "w"))


def big_golden(dl):
    """540x960 outputs of the live reference (a quarter of a 1080p frame; 17 row bands per frame
    for the wavefront kernel): the BASELINE configs' kernels and palette sizes, so that the oracle
    itself is pinned beyond the small cases.  Ostromoukhov (pure Python in the reference, ~14 kpx/s)
    is stored on a 270x480 crop.  tests/golden/big_cases.npz."""
    img = synth.frame(1080, 1920, 1)[:540, :960].copy()
    pal16 = synth.hex_palette(synth.PICO8)
    pal64, pal256 = synth.random_palette(64), synth.random_palette(256)
    store = {"img": img, "pal16": pal16, "pal64": pal64, "pal256": pal256}

    def run(arr, pal, mode, params):
        d = dl.ImageDitherer(num_colors=len(pal), dither_mode=dl.DitherMode(mode),
                             palette=[tuple(int(v) for v in c) for c in pal], dither_params=dict(params))
        return np.array(d.apply_dithering(Image.fromarray(arr, "RGB")))

    store["ed_jjn"] = run(img, pal256, "error_diffusion", {"variant": "jjn"})
    store["ed_atkinson"] = run(img, pal256, "error_diffusion", {"variant": "atkinson"})
    store["ed_sierra"] = run(img, pal64, "error_diffusion", {"variant": "sierra"})
    store["bayer8"] = run(img, pal16, "bayer", {"size": "8x8"})
    store["blue"] = run(img, pal16, "blue_noise", {"size": 64, "seed": 42})
    crop = synth.frame(2160, 3840, 2000)[:270, :480].copy()
    store["ostro_img"] = crop
    store["ostro"] = run(crop, pal64, "ostromoukhov", {})
    np.savez_compressed(os.path.join(OUT, "big_cases.npz"), **store)
    print("big_cases.npz:", os.path.getsize(os.path.join(OUT, "big_cases.npz")), "bytes")


BIG2_CASES = [
    # key, palette key, mode, params, crop (rows, cols) of the 540x960 image
    ("none16", "pal16", "none", {}, None),
    ("none256", "pal256", "none", {}, None),
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
    ("hybrid64", "pal64", "hybrid", {}, None),
]


def big2_golden(dl):
    """The rest of the mode list at 540x960 from the live reference (big_cases.npz holds the BASELINE
    configs' kernels): nearest colour, Bayer / IGN / blue noise / polka dot at 16-256 colours,
    halftone, the other diffusion kernels, serpentine, hybrid.  Outputs are stored as palette-index
    planes (the palettes are duplicate-free).  tests/golden/big_cases2.npz."""
    img = synth.frame(1080, 1920, 3)[:540, :960].copy()
    pals = {"pal16": synth.hex_palette(synth.PICO8), "pal64": synth.random_palette(64),
            "pal256": synth.random_palette(256)}
    store = {"img": img, **pals}
    for key, pk, mode, params, crop in BIG2_CASES:
        pal = pals[pk]
        arr = img if crop is None else img[:crop[0], :crop[1]].copy()
        d = dl.ImageDitherer(num_colors=len(pal), dither_mode=dl.DitherMode(mode),
                             palette=[tuple(int(v) for v in c) for c in pal], dither_params=dict(params))
        out = np.array(d.apply_dithering(Image.fromarray(arr, "RGB")))
        code = {tuple(int(v) for v in c): i for i, c in enumerate(np.asarray(pal, np.uint8))}
        assert len(code) == len(pal)
        packed = out[..., 0].astype(np.int32) | (out[..., 1].astype(np.int32) << 8) | (out[..., 2].astype(np.int32) << 16)
        lut = {(r | (g << 8) | (b << 16)): i for (r, g, b), i in code.items()}
        idx = np.vectorize(lut.__getitem__, otypes=[np.uint8])(packed)
        assert np.array_equal(np.asarray(pal, np.uint8)[idx], out)
        store[key] = idx
        print(key, out.shape)
    np.savez_compressed(os.path.join(OUT, "big_cases2.npz"), **store)
    print("big_cases2.npz:", os.path.getsize(os.path.join(OUT, "big_cases2.npz")), "bytes")


BIG3_CASES = [("perceptual64", "pal64", "perceptual", {}), ("perceptual16", "pal16", "perceptual", {}),
              ("adaptive16", "pal16", "adaptive_variance", {}),
              ("adaptive64", "pal64", "adaptive_variance", {"var_threshold": 60.0})]


def big3_golden(dl):
    """The two pure-Python diffusion modes of the reference (~14 kpx/s) on a 270x480 crop: nine
    row bands, the unclamped look-ups of `perceptual`, the variance map of `adaptive_variance`.
    Index planes.  tests/golden/big_cases3.npz."""
    img = synth.frame(1080, 1920, 4)[:270, :480].copy()
    pals = {"pal16": synth.hex_palette(synth.PICO8), "pal64": synth.random_palette(64)}
    store = {"img": img, **pals}
    for key, pk, mode, params in BIG3_CASES:
        pal = pals[pk]
        d = dl.ImageDitherer(num_colors=len(pal), dither_mode=dl.DitherMode(mode),
                             palette=[tuple(int(v) for v in c) for c in pal], dither_params=dict(params))
        out = np.array(d.apply_dithering(Image.fromarray(img, "RGB")))
        lut = {(int(c[0]) | (int(c[1]) << 8) | (int(c[2]) << 16)): i for i, c in enumerate(np.asarray(pal, np.uint8))}
        packed = out[..., 0].astype(np.int32) | (out[..., 1].astype(np.int32) << 8) | (out[..., 2].astype(np.int32) << 16)
        idx = np.vectorize(lut.__getitem__, otypes=[np.uint8])(packed)
        assert np.array_equal(np.asarray(pal, np.uint8)[idx], out)
        store[key] = idx
        print(key, out.shape)
    np.savez_compressed(os.path.join(OUT, "big_cases3.npz"), **store)
    print("big_cases3.npz:", os.path.getsize(os.path.join(OUT, "big_cases3.npz")), "bytes")


def kmeans4k_golden(dl):
    """BASELINE configs[2] exactly: k-means (k=16, random_state=42) on the seed-2 3840x2160 frame
    with random.seed(7) for the reference's 10 000-pixel subsample (dithering_lib.py:1845-1857).
    tests/golden/kmeans_4k.npz."""
    from sklearn.cluster import KMeans
    img = synth.frame(2160, 3840, 2)
    pix = img.reshape(-1, 3)
    random.seed(7)
    sample = pix[random.sample(range(len(pix)), 10000)]
    km = KMeans(n_clusters=16, random_state=42).fit(sample)
    random.seed(7)
    pal = dl.ColorReducer.generate_kmeans_palette(Image.fromarray(img, "RGB"), 16, 42)
    np.savez_compressed(os.path.join(OUT, "kmeans_4k.npz"), sample=sample, centers=km.cluster_centers_,
                        palette=np.asarray(pal, np.int64), niter=np.asarray(km.n_iter_))
    print("kmeans_4k.npz: n_iter", km.n_iter_, "palette", [tuple(map(int, c)) for c in pal][:3], "...")


def config4_golden(dl):
    """BASELINE configs[3] as the reference's CLI runs it (dither_cli.py:619-690): the palette by
    median cut from the FULL first frame (1080p, seed 1000), then per frame pixelize_regular(270)
    -> dither with the fixed palette -> x4 final resize; frames 0 and 1, blue noise (64, 42) and
    IGN (1.0, 0).  Stored: the palette, the dithered 480x270 frames as index planes and one
    up-scaled output.  tests/golden/config4.npz."""
    vp = load()[1]
    frames = [synth.frame(1080, 1920, 1000 + t) for t in range(2)]
    pal = dl.ColorReducer.reduce_colors(Image.fromarray(frames[0], "RGB"), 16)
    pal_u8 = np.asarray(pal, np.uint8)
    assert len({tuple(c) for c in pal}) == len(pal)
    lut = {(int(c[0]) | (int(c[1]) << 8) | (int(c[2]) << 16)): i for i, c in enumerate(pal_u8)}
    store = {"palette": np.asarray(pal, np.int64)}
    for name, mode, params in (("blue", "blue_noise", {"size": 64, "seed": 42}),
                               ("ign", "IGN", {"scale": 1.0, "seed": 0})):
        d = dl.ImageDitherer(num_colors=len(pal), dither_mode=dl.DitherMode(mode), palette=pal,
                             dither_params=dict(params))
        for t, f in enumerate(frames):
            small = vp.pixelize_regular(Image.fromarray(f, "RGB"), 270)
            out = d.apply_dithering(small)
            arr = np.array(out)
            packed = arr[..., 0].astype(np.int32) | (arr[..., 1].astype(np.int32) << 8) | (arr[..., 2].astype(np.int32) << 16)
            store[f"{name}_{t}"] = np.vectorize(lut.__getitem__, otypes=[np.uint8])(packed)
            if t == 1 and name == "blue":
                store["blue_1_x4"] = np.array(vp._apply_final_resize_to_frame(out, 4))
            print(name, t, arr.shape)
    np.savez_compressed(os.path.join(OUT, "config4.npz"), **store)
    print("config4.npz:", os.path.getsize(os.path.join(OUT, "config4.npz")), "bytes")


def baseline_hashes(dl):
    """SHA-256 of the live reference's outputs at the BASELINE configurations' FULL sizes (too large
    to store): configs[1] 3840x2160 seed 1, 256 colours, Floyd-Steinberg / Atkinson / JJN;
    configs[4] 3840x2160 seed 2000, 64 colours, Sierra; configs[0] 1920x1080 seed 0, Bayer 8x8,
    PICO-8.  tests/golden/baseline_hashes.json."""
    import hashlib
    res = {}

    def run(img, pal, mode, params):
        d = dl.ImageDitherer(num_colors=len(pal), dither_mode=dl.DitherMode(mode),
                             palette=[tuple(int(v) for v in c) for c in pal], dither_params=dict(params))
        out = np.ascontiguousarray(np.array(d.apply_dithering(Image.fromarray(img, "RGB"))))
        return hashlib.sha256(out.tobytes()).hexdigest()

    img4k = synth.frame(2160, 3840, 1)
    for v in ("floyd_steinberg", "atkinson", "jjn"):
        res[f"config2_{v}"] = run(img4k, synth.random_palette(256), "error_diffusion", {"variant": v})
        print(v, res[f"config2_{v}"][:16])
    res["config5_sierra_frame2000"] = run(synth.frame(2160, 3840, 2000), synth.random_palette(64),
                                          "error_diffusion", {"variant": "sierra"})
    res["config1_bayer8x8"] = run(synth.frame(1080, 1920, 0), synth.hex_palette(synth.PICO8), "bayer",
                                  {"size": "8x8"})
    # the threshold family and nearest colour at full size
    f1080 = synth.frame(1080, 1920, 0)
    pico, r256 = synth.hex_palette(synth.PICO8), synth.random_palette(256)
    res["1080p_halftone_pico8"] = run(f1080, pico, "halftone", {})
    res["1080p_ign_pico8"] = run(f1080, pico, "IGN", {"scale": 1.0, "seed": 0})
    res["1080p_blue_noise_pico8"] = run(f1080, pico, "blue_noise", {"size": 64, "seed": 42})
    res["1080p_none_pico8"] = run(f1080, pico, "none", {})
    res["1080p_none_r256"] = run(f1080, r256, "none", {})
    res["1080p_bayer8x8_r256"] = run(f1080, r256, "bayer", {"size": "8x8"})
    km = np.load(os.path.join(OUT, "kmeans_4k.npz"))["palette"]
    res["config3_4k_none_kmeans_palette"] = run(synth.frame(2160, 3840, 2), km, "none", {})
    # the other diffusion kernels at 4K (64 colours), serpentine at 1080p, Ostromoukhov at 540x960
    r64b = synth.random_palette(64)
    f4k = synth.frame(2160, 3840, 2001)
    for v in ("stucki", "burkes", "sierra_two_row", "sierra_lite"):
        res[f"4k_{v}_r64"] = run(f4k, r64b, "error_diffusion", {"variant": v})
    res["1080p_fs_serpentine_r64"] = run(f1080, r64b, "error_diffusion",
                                         {"variant": "floyd_steinberg", "serpentine": "true"})
    res["540p_ostromoukhov_r64"] = run(synth.frame(1080, 1920, 6)[:540, :960].copy(), r64b, "ostromoukhov", {})
    # gamma correction (use_gamma=True) at 540x960
    def run_gamma(img, pal, mode, params):
        d = dl.ImageDitherer(num_colors=len(pal), dither_mode=dl.DitherMode(mode), use_gamma=True,
                             palette=[tuple(int(v) for v in c) for c in pal], dither_params=dict(params))
        out = np.ascontiguousarray(np.array(d.apply_dithering(Image.fromarray(img, "RGB"))))
        return hashlib.sha256(out.tobytes()).hexdigest()

    g540 = synth.frame(1080, 1920, 5)[:540, :960].copy()
    r64 = synth.random_palette(64)
    res["gamma_540p_bayer8x8_pico8"] = run_gamma(g540, pico, "bayer", {"size": "8x8"})
    res["gamma_540p_none_r64"] = run_gamma(g540, r64, "none", {})
    res["gamma_540p_halftone_pico8"] = run_gamma(g540, pico, "halftone", {})
    res["gamma_540p_fs_r64"] = run_gamma(g540, r64, "error_diffusion", {"variant": "floyd_steinberg"})
    res["gamma_540p_jjn_pico8"] = run_gamma(g540, pico, "error_diffusion", {"variant": "jjn"})
    json.dump(res, open(os.path.join(OUT, "baseline_hashes.json"), "w"), indent=1)
    print(res)


if __name__ == "__main__":
    if "--hashes" in sys.argv:            # adds baseline_hashes.json without touching the other files
        baseline_hashes(load()[0])
    elif "--config4" in sys.argv:           # adds config4.npz without touching the other files
        config4_golden(load()[0])
    elif "--kmeans4k" in sys.argv:          # adds kmeans_4k.npz without touching the other files
        kmeans4k_golden(load()[0])
    elif "--big3" in sys.argv:              # adds big_cases3.npz without touching the other files
        big3_golden(load()[0])
    elif "--big2" in sys.argv:            # adds big_cases2.npz without touching the other files
        big2_golden(load()[0])
    elif "--big" in sys.argv:             # adds big_cases.npz without touching the other files
        big_golden(load()[0])
    elif "--median-cut-only" in sys.argv:   # adds median_cut.json without touching the other files
        median_cut_golden(load()[0])
    else:
        main()
