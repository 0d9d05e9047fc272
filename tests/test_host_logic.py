"""CPU: host-side logic of the package (tables, geometry, sharding, API surface)."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
import dither_pie_b200 as dp
from dither_pie_b200 import engine, kmeans, synth
from dither_pie_b200.video_processor import shard_frames


def test_threshold_tables_match_reference_golden():
    g = load_golden("threshold_sources.npz")
    for s in ("2x2", "4x4", "8x8", "16x16", "psx4x4"):
        assert np.array_equal(engine.bayer_matrix(s), g["bayer_" + s])
    assert np.array_equal(engine.blue_noise_matrix(64, 42), g["blue_64_42"])
    assert np.array_equal(dp.generate_blue_noise(32, 5), g["blue_32_5"])
    assert np.array_equal(engine.polka_dot_matrix(8, 1.5), g["polka_8_1.5"])
    assert np.array_equal(engine.ostromoukhov_coeffs(), g["ostro_coeffs"])
    assert np.array_equal(dp.DitherUtils.BAYER8x8, g["bayer_8x8"])


def test_pixelize_geometry_matches_pillow_golden():
    g = load_golden("pixelize.npz")
    meta = json.load(open(os.path.join(GOLDEN, "pixelize.json")))
    for m in meta:
        tw, th = engine.even_dimensions(m["w"], m["h"], m["max_size"])
        assert (tw, th) == (m["tw"], m["th"])
        key = f"{m['w']}x{m['h']}_{m['max_size']}"
        assert np.array_equal(engine.nearest_table(m["w"], tw), g["xt_" + key])
        assert np.array_equal(engine.nearest_table(m["h"], th), g["yt_" + key])


def test_gamma_luts_are_the_elementwise_chain():
    lut = engine.gamma_in_lut()
    v = np.arange(256, dtype=np.uint8)
    ref = np.clip(engine.srgb_to_linear(v.astype(np.float32) / 255.0) * 255.0, 0, 255).astype(np.uint8)
    assert np.array_equal(lut, ref) and lut[0] == 0 and lut[255] == 255


def test_kmeans_plusplus_matches_sklearn_public_function():
    from sklearn.cluster import kmeans_plusplus
    g = load_golden("kmeans.npz")
    X = g["sample_1"].astype(np.float64)
    Xc = X - X.mean(axis=0)
    mine = kmeans.kmeans_plusplus(Xc, 8, 42)
    ref, _ = kmeans_plusplus(Xc, 8, random_state=np.random.RandomState(42))
    assert np.allclose(mine, ref, atol=1e-9)


def test_shard_frames_partitions_contiguously():
    for n in (0, 1, 7, 600, 301):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_frames(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_api_surface_matches_reference_names():
    ref_all = ['DitherMode', 'PixelizeMethod', 'PaletteSource', 'ImageDitherer', 'ColorReducer',
               'DitherUtils', 'BaseDitherStrategy', 'ErrorDiffusionKernel', 'NoDitherStrategy',
               'MatrixDitherStrategy', 'BayerDitherStrategy', 'BlueNoiseDitherStrategy',
               'InterleavedGradientNoiseDitherStrategy', 'ErrorDiffusionDitherStrategy',
               'OstromoukhovDitherStrategy', 'RiemersmaDitherStrategy', 'PolkaDotDitherStrategy',
               'WaveletDitherStrategy', 'AdaptiveVarianceDitherStrategy',
               'PerceptualDitherStrategy', 'HybridDitherStrategy', 'HalftoneDitherStrategy',
               'generate_blue_noise']
    from dither_pie_b200 import dithering_lib as dl
    assert sorted(dl.__all__) == sorted(ref_all)
    for name in ref_all:
        assert hasattr(dl, name)
    assert [m.value for m in dl.DitherMode] == [
        "none", "bayer", "error_diffusion", "riemersma", "blue_noise", "IGN", "polka_dot",
        "wavelet", "adaptive_variance", "perceptual", "hybrid", "halftone", "ostromoukhov"]
    assert dl.ErrorDiffusionDitherStrategy.get_parameter_info()["variant"]["default"] == "atkinson"
    assert dl.ErrorDiffusionDitherStrategy().get_current_parameters() == {
        "variant": "atkinson", "serpentine": "false"}
    assert dl.ErrorDiffusionKernel.get_kernel("nope") is dl.ErrorDiffusionKernel.FLOYD_STEINBERG
    assert dl.ErrorDiffusionKernel.JJN["divisor"] == 48 and len(dl.ErrorDiffusionKernel.SIERRA["weights"]) == 10
    with pytest.raises(NotImplementedError):
        dl.RiemersmaDitherStrategy()
    assert dl.AdaptiveVarianceDitherStrategy().get_current_parameters() == {
        "var_threshold": 300.0, "window_radius": 1}
    with pytest.raises(NotImplementedError):
        dl.PerceptualDitherStrategy(base_weights=[(1, 0, 1.0)])
    assert dl.PerceptualDitherStrategy().base_weights[0] == (1, 0, 7 / 16)
    assert dl.ImageDitherer.get_mode_parameters(dl.DitherMode.PERCEPTUAL) is None
    assert dl.HybridDitherStrategy().get_current_parameters() == {"lum_factor": 1.0, "col_factor": 0.2}
    assert dl.ImageDitherer.get_mode_parameters(dl.DitherMode.HYBRID)["col_factor"]["default"] == 0.2
    with pytest.raises(TypeError):
        dl.ImageDitherer(palette=[(0, 0, 0)], dither_params={"bogus": 1})._get_dither_strategy(
            dl.DitherMode.BAYER)
    import pickle
    d = dl.ImageDitherer(8, dl.DitherMode.HALFTONE, [(1, 2, 3)], True, {"cell_size": 4})
    d2 = pickle.loads(pickle.dumps(d))
    assert d2.palette == [(1, 2, 3)] and d2.dither_params == {"cell_size": 4}


def test_median_cut_and_uniform_palettes_match_reference_golden():
    """reduce_colors (array path for > 64 unique colours, list path below) and the uniform cube
    against palettes produced by the live reference (tools/make_golden.py, median_cut.json)."""
    import json
    from PIL import Image
    g = json.load(open(os.path.join(GOLDEN, "median_cut.json")))
    for c in g["median_cut"]:
        gen = {"frame": synth.frame, "noise": synth.noise_frame}.get(c["kind"])
        arr = gen(c["h"], c["w"], c["seed"]) if gen else synth.blocks_frame(c["h"], c["w"], c["seed"], 8, 6)
        pal = dp.ColorReducer.reduce_colors(Image.fromarray(arr, "RGB"), c["num_colors"])
        assert [list(map(int, p)) for p in pal] == c["palette"], c
    for n, pal in g["uniform"].items():
        assert [list(map(int, p)) for p in dp.ColorReducer.generate_uniform_palette(int(n))] == pal, n


def test_unique_colours_come_out_in_cpython_set_order():
    """csrc/dp_pyset.cu replays CPython's tuple hash and set table: the order must equal the
    running interpreter's own ``list(set(image.getdata()))`` (dithering_lib.py:1837) -- across
    the growth policy's regimes (x4 below 50 000 entries, x2 above), duplicates, tiny inputs."""
    from PIL import Image
    rs = np.random.RandomState(5)
    cases = [synth.frame(24, 32, 3), synth.noise_frame(60, 80, 6), synth.blocks_frame(64, 64, 7, 8, 6),
             synth.frame(270, 480, 1), synth.noise_frame(300, 400, 2),          # > 50 000 unique colours
             np.zeros((3, 5, 3), np.uint8), rs.randint(0, 4, (50, 50, 3)).astype(np.uint8),
             rs.randint(0, 256, (1, 1, 3)).astype(np.uint8),
             np.concatenate([synth.noise_frame(200, 300, 9)] * 2)]
    for arr in cases:
        want = list(set(Image.fromarray(arr, "RGB").getdata()))
        got = dp.ColorReducer.unique_colors_in_set_order(arr)
        assert got.shape == (len(want), 3)
        assert [tuple(int(v) for v in c) for c in got] == want, arr.shape


def test_median_cut_and_uniform_palettes_match_reference_semantics():
    from PIL import Image
    img = Image.fromarray(synth.frame(24, 32, 3), "RGB")
    pal = dp.ColorReducer.reduce_colors(img, 12)   # depth = int(log2(12)) = 3 -> 8 colours
    assert len(pal) == 8 and all(len(c) == 3 for c in pal)
    assert dp.ColorReducer.generate_uniform_palette(8)[-1] == (255, 255, 255)
    assert dp.ColorReducer.generate_uniform_palette(1) == [(128, 128, 128)]


def test_native_blue_noise_equals_the_restated_loop():
    """dp_blue_noise_from_order (native farthest-point replay) against the oracle's numpy
    restatement of generate_blue_noise (:381-399) on sizes / seeds the golden file does not hold."""
    from oracle import dither_oracle as O
    for size, seed in ((1, 0), (2, 3), (7, 11), (16, 1), (24, 99), (48, 42)):
        assert np.array_equal(engine.blue_noise_matrix(size, seed), O.blue_noise_matrix(size, seed)), (size, seed)


def test_array_median_cut_equals_the_list_restatement():
    """The planar array cut (stable radix argsort) against the literal list restatement of
    ColorReducer.median_cut (:1822-1832) fed with the interpreter's own set order."""
    from PIL import Image
    rs = np.random.RandomState(11)
    for arr, nc in ((synth.frame(60, 90, 2), 16), (synth.noise_frame(40, 50, 3), 8),
                    (rs.randint(0, 6, (80, 80, 3)).astype(np.uint8) * 50, 32),
                    (synth.blocks_frame(64, 96, 4, 4, 5), 4), (synth.frame(100, 150, 5), 256)):
        img = Image.fromarray(arr, "RGB")
        uniq = list(set(img.getdata()))
        assert len(uniq) >= 64
        depth = int(np.log2(nc))
        want = dp.ColorReducer.median_cut(list(uniq), depth)
        got = dp.ColorReducer.reduce_colors(img, nc)
        assert [tuple(int(v) for v in c) for c in got] == [tuple(int(v) for v in c) for c in want], nc


def test_fix_failed_frames_copies_the_nearest_good_frame():
    """video_processor.py:53-96: previous frames first, then following ones."""
    from dither_pie_b200.video_processor import VideoProcessor
    out = np.arange(6, dtype=np.uint8).reshape(6, 1, 1, 1).repeat(3, axis=3).copy()
    lost = VideoProcessor._fix_failed_frames([0, 1, 4], out)
    assert lost == []
    assert out[:, 0, 0, 0].tolist() == [2, 2, 2, 3, 3, 5]
    out2 = np.zeros((2, 1, 1, 3), np.uint8)
    assert VideoProcessor._fix_failed_frames([0, 1], out2) == [0, 1]


def test_video_processor_worker_and_rank_bookkeeping(monkeypatch):
    from dither_pie_b200.video_processor import NeuralPixelizer, VideoProcessor, _unpack_pixelize
    monkeypatch.setenv("RANK", "3")
    monkeypatch.setenv("WORLD_SIZE", "8")
    vp = VideoProcessor()
    assert (vp.rank, vp.world, vp.num_workers) == (3, 8, 8)
    assert VideoProcessor(num_workers=2).num_workers == 2
    assert _unpack_pixelize(None) is None
    assert _unpack_pixelize(("regular", 64)) == 64
    assert _unpack_pixelize((dp.PixelizeMethod.REGULAR, 65)) == 65
    assert _unpack_pixelize(("none", 64)) is None
    with pytest.raises(NotImplementedError):
        _unpack_pixelize(("neural", 64))
    assert NeuralPixelizer._compute_even_dimensions(1920, 1080, 270) == (480, 270)
    with pytest.raises(NotImplementedError):
        NeuralPixelizer().pixelize(None, 64)


def test_video_geometry_matches_the_reference_rules():
    # pixelize to even dims, x m, odd sizes bumped to even only on the video path
    assert engine.video_geometry(1080, 1920, 270, 4) == ((270, 480), (1080, 1920), 4)
    assert engine.video_geometry(90, 150, 31, 3, True) == ((30, 50), (90, 150), 3)
    assert engine.video_geometry(45, 75, None, 3, True) == ((45, 75), (136, 226), 3)
    assert engine.video_geometry(45, 75, None, 3, False) == ((45, 75), (135, 225), 3)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout")
def test_reference_cli_imports_against_the_drop_in_modules():
    """the reference's front end loads with the two hot-path modules
    swapped for this package (it imports VideoProcessor, NeuralPixelizer, pixelize_regular and the
    dithering_lib names unconditionally, dither_cli.py:26)."""
    import subprocess
    import sys
    code = (
        "import sys, types\n"
        "sys.path.insert(0, %r); sys.path.insert(1, '/root/reference')\n"
        "sys.modules.setdefault('pywt', types.ModuleType('pywt'))\n"
        "import dither_pie_b200.dithering_lib as dl, dither_pie_b200.video_processor as vp\n"
        "sys.modules['dithering_lib'] = dl; sys.modules['video_processor'] = vp\n"
        "import dither_cli\n"
        "assert dither_cli.VideoProcessor is vp.VideoProcessor\n"
        "assert dither_cli.ImageDitherer is dl.ImageDitherer\n"
        "import pathlib\n"
        "cfg = dither_cli.validate_config({'input': 'a.png', 'output': 'b.png'}, pathlib.Path('cfg.json'), True)\n"
        "print('ok', sorted(cfg)[:3])\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.stdout, r.stderr[-800:])


def test_bench_reference_arm_contract_on_cpu():
    """`bench.py --impl reference` (the oracle's C port on the host cores) runs without a GPU and
    prints ONE JSON line with the keys the driver reads: same metric / unit / config as the GPU arm,
    `impl`, a `cpu_baseline` describing the run and an `e2e` that repeats the line's value."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpixels/s" and d["unit"] == "Mpx/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "Mpx/s", "h2d_bytes_per_step": 0,
                                           "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "configs[1]" in d["config"]["workload"]
