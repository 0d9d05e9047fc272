"""GPU parity tests, second batch: the BASELINE.json configurations at their own sizes, the
index-plane-only output of every entry point, the host-buffer frame pipeline, the drop-in
VideoProcessor (sharding, failure contract, raw-frame pipes) and a stress test of the wavefront's
inter-band hand-off.  Everything goes through the ctypes C ABI and is compared bit for bit with
the oracle (the checker)."""
import ctypes as C
import os
import random
import stat
import sys
import textwrap
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import dither_pie_b200 as dp  # noqa: E402
from dither_pie_b200 import _capi, engine, kmeans, pipeline, synth  # noqa: E402
from dither_pie_b200.dithering_lib import DitherMode  # noqa: E402
from dither_pie_b200.video_processor import VideoProcessor  # noqa: E402
from oracle import dither_oracle as O  # noqa: E402  (the checker)

PICO = synth.hex_palette(synth.PICO8)


def mismatch(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return int((a != b).any(axis=-1).sum()) if a.ndim >= 3 and a.shape[-1] == 3 else int((a != b).sum())


def oracle_many(jobs, threads=None):
    """[(img, pal, mode, params)] -> outputs; the C port releases the GIL, one job per core."""
    with ThreadPoolExecutor(max_workers=threads or min(len(jobs), os.cpu_count() or 1)) as ex:
        return list(ex.map(lambda j: O.apply_dithering(*j), jobs))


# ------------------------------------------------------------------ BASELINE configs at size
def test_config2_4k_atkinson_and_jjn_256_colours_vs_oracle():
    """configs[1]: 3840x2160, seed 1, K=256 -- the two kernels of the headline step that had no
    full-size check (Floyd-Steinberg has one in test_gpu_parity.py)."""
    img = synth.frame(2160, 3840, 1)
    pal = synth.random_palette(256)
    jobs = [(img, pal, "error_diffusion", {"variant": v}) for v in ("atkinson", "jjn")]
    refs = oracle_many(jobs)
    import hashlib
    import json
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "baseline_hashes.json")))
    for (_, _, mode, params), ref in zip(jobs, refs):
        out = engine.dither_frames(img, pal, mode, params)
        assert mismatch(out, ref) == 0, params
        # ... and the LIVE reference's own bytes (tools/make_golden.py --hashes)
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == want["config2_" + params["variant"]]


def test_config5_4k_ostromoukhov_and_sierra_64_colours_multi_frame_vs_oracle():
    """configs[4]: 3840x2160 frames (seeds 2000+t), K=64, Ostromoukhov and Sierra, three frames
    in ONE call so that bands of different frames interleave in the ready queue."""
    frames = np.stack([synth.frame(2160, 3840, 2000 + t) for t in range(3)])
    pal = synth.random_palette(64)
    for mode, params in (("ostromoukhov", {}), ("error_diffusion", {"variant": "sierra"})):
        refs = oracle_many([(f, pal, mode, params) for f in frames])
        out = engine.dither_frames(frames, pal, mode, params)
        for t in range(3):
            assert mismatch(out[t], refs[t]) == 0, (mode, t)
        if mode == "error_diffusion":      # frame 2000 under Sierra: the LIVE reference's own bytes
            import hashlib
            import json
            want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "baseline_hashes.json")))
            assert hashlib.sha256(np.ascontiguousarray(out[0]).tobytes()).hexdigest() == want["config5_sierra_frame2000"]


def test_config4_1080p_video_pixelize270_blue_noise_ign_x4_as_written():
    """configs[3]: 1920x1080 frames (seeds 1000+t), pixelize_regular max_size=270 -> 480x270,
    blue noise (64, 42) and IGN (1.0, 0), 16 colours fixed from frame 0 (median cut, as
    dither_cli.py:619-654), final resize x4, through VideoProcessor.process_frames."""
    frames = np.stack([synth.frame(1080, 1920, 1000 + t) for t in range(4)])
    first_small = O.pixelize_regular(frames[0], 270)
    assert first_small.shape == (270, 480, 3)
    for mode, params in ((DitherMode.BLUE_NOISE, {"size": 64, "seed": 42}),
                         (DitherMode.INTERLEAVED_GRADIENT_NOISE, {"scale": 1.0, "seed": 0})):
        d = dp.ImageDitherer(num_colors=16, dither_mode=mode, palette=None, dither_params=params)
        out = VideoProcessor().process_frames(frames, d, ("regular", 270), batch_size=3,
                                              final_resize_multiplier=4)
        assert out.shape == (4, 1080, 1920, 3)
        assert len(d.palette) == 16          # derived from frame 0, stored on the ditherer
        for t in range(4):
            ref = O.final_resize(O.apply_dithering(O.pixelize_regular(frames[t], 270), d.palette,
                                                   mode.value, params), 4, True)
            assert mismatch(out[t], ref) == 0, (mode, t)
        idx = VideoProcessor().process_frames(frames, d, ("regular", 270), batch_size=3,
                                              final_resize_multiplier=4, output="index")
        assert idx.shape == (4, 270, 480)
        pal_u8 = np.asarray(d.palette, np.uint8)
        assert np.array_equal(np.repeat(np.repeat(pal_u8[idx], 4, axis=1), 4, axis=2), out)


def test_config3_4k_kmeans_palette_then_nearest_colour():
    """configs[2]: 3840x2160 seed 2, random.seed(7), k-means k=16 random_state=42 (centres within
    1e-3 of the oracle's Lloyd from the same subsample and k-means++ init), then `none` with the
    shared palette at 4K."""
    from PIL import Image
    img = synth.frame(2160, 3840, 2)
    random.seed(7)
    pal = dp.ColorReducer.generate_kmeans_palette(Image.fromarray(img), 16, random_state=42)
    random.seed(7)
    flat = img.reshape(-1, 3)
    sample = flat[random.sample(range(len(flat)), kmeans.SAMPLE)]
    got, it = kmeans.kmeans_fit(sample, 16, 42)
    ref = O.kmeans_centers(sample, 16, 42)
    assert np.abs(got - ref).max() <= 1e-3, np.abs(got - ref).max()
    assert [tuple(int(v) for v in r) for r in pal] == [tuple(int(v) for v in r) for r in ref.astype(int)]
    out, idx = engine.dither_frames(img, pal, "none", return_indices=True)
    assert mismatch(out, O.apply_dithering(img, pal, "none")) == 0
    assert np.array_equal(np.asarray(pal, np.uint8)[idx], out)


def test_golden_big_cases_from_the_live_reference():
    """540x960 outputs of the reference itself (tools/make_golden.py --big): pins the oracle AND
    the CUDA path beyond the three row bands of diffusion_big.npz."""
    path = os.path.join(os.path.dirname(__file__), "golden", "big_cases.npz")
    g = np.load(path)
    img = g["img"]
    assert img.shape[0] >= 540 and img.shape[1] >= 960
    cases = [("ed_jjn", "pal256", "error_diffusion", {"variant": "jjn"}),
             ("ed_atkinson", "pal256", "error_diffusion", {"variant": "atkinson"}),
             ("ed_sierra", "pal64", "error_diffusion", {"variant": "sierra"}),
             ("bayer8", "pal16", "bayer", {"size": "8x8"}),
             ("blue", "pal16", "blue_noise", {"size": 64, "seed": 42})]
    for key, pk, mode, params in cases:
        out = engine.dither_frames(img, g[pk], mode, params)
        assert mismatch(out, g[key]) == 0, key
    crop = g["ostro_img"]
    assert mismatch(engine.dither_frames(crop, g["pal64"], "ostromoukhov"), g["ostro"]) == 0


# ------------------------------------------------------------------ index-plane-only output
IDX_CASES = [("none", {}), ("bayer", {"size": "8x8"}), ("IGN", {}), ("blue_noise", {"size": 32, "seed": 5}),
             ("halftone", {}), ("error_diffusion", {"variant": "floyd_steinberg"}),
             ("error_diffusion", {"variant": "jjn"}),
             ("error_diffusion", {"variant": "atkinson", "serpentine": "true"}),
             ("ostromoukhov", {}), ("hybrid", {}), ("perceptual", {}),
             ("adaptive_variance", {"var_threshold": 60.0})]


@pytest.mark.parametrize("mode,params", IDX_CASES)
def test_index_only_output_matches_the_colour_output(mode, params):
    """dst_rgb == NULL: only the palette-index plane is produced; it must be the rows whose
    colours the full call writes (and the oracle's rows, for unique palettes)."""
    for (h, w), pal in (((70, 112), PICO), ((45, 83), synth.random_palette(64)),
                        ((96, 160), synth.random_palette(256))):
        frames = np.stack([synth.frame(h, w, 40 + t) for t in range(3)])
        rgb, idx_both = engine.dither_frames(frames, pal, mode, params, return_indices=True)
        idx = engine.dither_frames(frames, pal, mode, params, indices_only=True)
        assert idx.shape == (3, h, w) and idx.dtype == np.uint8
        assert np.array_equal(idx, idx_both)
        assert np.array_equal(np.asarray(pal, np.uint8)[idx], rgb)
        ref = O.apply_dithering(frames[1], pal, mode, params)
        assert mismatch(np.asarray(pal, np.uint8)[idx[1]], ref) == 0


def test_index_only_with_fused_geometry_and_null_outputs_refused():
    frames = np.stack([synth.frame(216, 384, 7 + t) for t in range(2)])
    idx = engine.dither_frames(frames, PICO, "bayer", {"size": "4x4"}, pixelize_max_size=54,
                               final_multiplier=4, indices_only=True)
    assert idx.shape == (2, 54, 96)
    for t in range(2):
        ref = O.apply_dithering(O.pixelize_regular(frames[t], 54), PICO, "bayer", {"size": "4x4"})
        assert mismatch(np.asarray(PICO, np.uint8)[idx[t]], ref) == 0
    pal = engine.get_palette(PICO)
    plan = engine.Plan("none", {}, 8, 16)
    buf = _capi.DeviceBuffer(8 * 16 * 3)
    with pytest.raises(_capi.DitherPieError):
        plan.run(pal, buf.ptr, 1, None, None)
    buf.free()


def test_halftone_v2_kernels_multi_tile_gamma_and_old_kernels_agree(monkeypatch):
    """Rows of a multiple of 16 pixels take the v2 halftone kernels (tile-local cell maps, one u64 of
    sums per cell, colour table in shared memory): several tiles in both directions, partial tiles,
    cell sizes up to the 16-bit limit, gamma LUT, colour + index output -- against the oracle, and
    against the old kernels (DP_HT_V1)."""
    cases = [{}, {"cell_size": 5, "angle": 30.0, "shape": "diamond"},
             {"cell_size": 14, "angle": 75.0, "shape": "square"}, {"cell_size": 3, "angle": 90.0},
             {"cell_size": 8, "angle": 45.0, "dot_gain": 1.3, "min_dot_size": 0.1, "max_dot_size": 0.9},
             {"cell_size": 2, "angle": 10.0}, {"cell_size": 20, "angle": 60.0}]
    frames = np.stack([synth.frame(200, 272, 90 + t) if t != 1 else synth.noise_frame(200, 272, 91)
                       for t in range(3)])
    for pal, gamma in ((PICO, False), (synth.random_palette(64), False), (PICO, True)):
        rows = np.asarray(pal, np.uint8)
        for params in cases:
            rgb, idx = engine.dither_frames(frames, pal, "halftone", params, use_gamma=gamma, return_indices=True)
            only = engine.dither_frames(frames, pal, "halftone", params, use_gamma=gamma, indices_only=True)
            assert np.array_equal(idx, only), (params, gamma)
            if not gamma:
                assert np.array_equal(rows[idx], rgb), (params, gamma)
            refs = oracle_many([(frames[t], pal, "halftone", params, gamma) for t in range(3)])
            for t in range(3):
                assert mismatch(rgb[t], refs[t]) == 0, (params, gamma, t)
            monkeypatch.setenv("DP_HT_V1", "1")
            old = engine.dither_frames(frames, pal, "halftone", params, use_gamma=gamma)
            monkeypatch.delenv("DP_HT_V1")
            assert np.array_equal(old, rgb), (params, gamma)


# ------------------------------------------------------------------ host-buffer entry point
def test_threshold_dither_host_entry_point():
    """dp_threshold_dither_host -- the call INTEGRATION.md's example binding uses."""
    L = _capi.lib()
    frames = np.stack([synth.frame(64, 96, 3 + t) for t in range(3)])
    pal = engine.get_palette(PICO)
    for kind, mat, ign, mode, params in (
            (0, None, (0.0, 0.0, 1.0), "none", {}),
            (1, engine.bayer_matrix("8x8"), (0.0, 0.0, 1.0), "bayer", {"size": "8x8"}),
            (2, None, (float(np.float32(17 * 0.37)), float(np.float32(17 * 0.73)), 2.5), "IGN",
             {"scale": 2.5, "seed": 17})):
        out = np.zeros_like(frames)
        mh, mw = (mat.shape if mat is not None else (0, 0))
        _capi.check(L.dp_threshold_dither_host(pal.handle, frames.ctypes.data, 3, 64, 96, kind,
                                               mat.ctypes.data if mat is not None else None, mh, mw,
                                               ign[0], ign[1], ign[2], out.ctypes.data),
                    "dp_threshold_dither_host")
        for t in range(3):
            assert mismatch(out[t], O.apply_dithering(frames[t], PICO, mode, params)) == 0, (mode, t)
    assert L.dp_threshold_dither_host(pal.handle, frames.ctypes.data, 1, 64, 96, 1, None, 0, 0,
                                      0.0, 0.0, 1.0, frames.ctypes.data) != 0   # matrix missing


# ------------------------------------------------------------------ hand-off stress
def test_wavefront_handoff_stress_tiny_frames_two_blocks(monkeypatch):
    """Many small frames, a width just above one 32-step chunk, the grid forced down to two
    blocks of two warps: every band waits on its predecessor's progress word and the ready queue
    is permanently contended.  Bit-exact for every footprint class."""
    monkeypatch.setenv("DP_WAVE_WARPS", "2")
    monkeypatch.setenv("DP_WAVE_GRID", "2")
    pal = synth.random_palette(16)
    frames = np.stack([synth.frame(70, 37, 300 + t) for t in range(48)])
    for mode, params in (("error_diffusion", {"variant": "floyd_steinberg"}),
                         ("error_diffusion", {"variant": "jjn"}),
                         ("error_diffusion", {"variant": "sierra"}), ("ostromoukhov", {})):
        refs = oracle_many([(f, pal, mode, params) for f in frames])
        for rep in range(3):
            out = engine.dither_frames(frames, pal, mode, params)
            bad = [t for t in range(len(frames)) if mismatch(out[t], refs[t])]
            assert not bad, (mode, params, rep, bad[:5])


def test_wavefront_byte_and_table_instantiations_agree(monkeypatch):
    """Plain byte palettes without gamma take the BYTES instantiations of the wavefront kernel
    (work values and palette values formed from the bytes themselves); gamma palettes and
    DP_WAVE_NO_BYTES take the shared-memory tables.  Both against the oracle, every state type
    (f64 numba variants, hybrid, f32 Ostromoukhov / perceptual), 4- and 7-slot screening."""
    frames = np.stack([synth.frame(75, 133, 700 + t) if t else synth.noise_frame(75, 133, 700)
                       for t in range(3)])
    cases = (("error_diffusion", {"variant": "floyd_steinberg"}), ("error_diffusion", {"variant": "stucki"}),
             ("error_diffusion", {"variant": "sierra_lite"}), ("ostromoukhov", {}), ("hybrid", {}),
             ("perceptual", {}))
    for pal in (PICO, synth.random_palette(256)):
        for mode, params in cases:
            refs = oracle_many([(f, pal, mode, params) for f in frames])
            out = engine.dither_frames(frames, pal, mode, params)
            monkeypatch.setenv("DP_WAVE_NO_BYTES", "1")
            tab = engine.dither_frames(frames, pal, mode, params)
            monkeypatch.delenv("DP_WAVE_NO_BYTES")
            assert np.array_equal(out, tab), (mode, params, len(pal))
            for t in range(3):
                assert mismatch(out[t], refs[t]) == 0, (mode, params, len(pal), t)
    for mode, params in cases[:4]:     # gamma: the table instantiations by themselves
        refs = oracle_many([(f, PICO, mode, params, True) for f in frames])
        out = engine.dither_frames(frames, PICO, mode, params, use_gamma=True)
        for t in range(3):
            assert mismatch(out[t], refs[t]) == 0, (mode, params, "gamma", t)


# ------------------------------------------------------------------ frame pipeline
def test_frame_pipeline_matches_direct_calls_pinned_and_pageable():
    frames = np.stack([synth.frame(120, 208, 60 + t) for t in range(11)])
    pal_rows = synth.random_palette(64)
    pal = engine.get_palette(pal_rows)
    variants = ("floyd_steinberg", "atkinson", "jjn")
    plans = [engine.Plan("error_diffusion", {"variant": v}, 120, 208) for v in variants]
    want = [engine.dither_frames(frames, pal_rows, "error_diffusion", {"variant": v}) for v in variants]
    with pipeline.FramePipeline(plans, pal, batch_frames=4, output="both") as pipe:
        rgb, idx = pipe.run(frames)                       # pageable arrays: staged
        for v in range(3):
            assert np.array_equal(rgb[v], want[v])
            assert np.array_equal(np.asarray(pal_rows, np.uint8)[idx[v]], want[v])
        assert pipe.stats["h2d_bytes"] == frames.nbytes and not pipe.stats["in_pinned"]
        pin_in = pipeline.pinned_empty(frames.shape)
        pin_in[...] = frames
        outs = [pipeline.pinned_empty(frames.shape) for _ in variants]
        idxs = [pipeline.pinned_empty(frames.shape[:3]) for _ in variants]
        for rep in range(2):                              # pinned arrays: DMA in place, reusable
            for o in outs:
                o[...] = 0
            pipe.run(pin_in, out_rgb=outs, out_idx=idxs)
            assert pipe.stats["in_pinned"] and pipe.stats["out_pinned"]
            for v in range(3):
                assert np.array_equal(outs[v], want[v])
        for a in [pin_in] + outs + idxs:
            pipeline.release_pinned(a)
    with pipeline.FramePipeline([engine.Plan("bayer", {"size": "8x8"}, 120, 208)], pal, 5, "index") as pipe:
        _, idx = pipe.run(frames)
        ref = engine.dither_frames(frames, pal_rows, "bayer", {"size": "8x8"}, indices_only=True)
        assert np.array_equal(idx[0], ref)


# ------------------------------------------------------------------ VideoProcessor
def test_video_processor_non_fused_modes_and_odd_sizes():
    """pixelize -> error diffusion / halftone -> x3 (odd sizes bumped to even) goes through the
    device chain resample -> dither -> resample."""
    frames = np.stack([synth.frame(90, 150, 80 + t) for t in range(5)])
    for mode, params in ((DitherMode.ERROR_DIFFUSION, {"variant": "burkes"}), (DitherMode.HALFTONE, {})):
        d = dp.ImageDitherer(dither_mode=mode, palette=[tuple(r) for r in PICO], dither_params=params)
        out = VideoProcessor().process_frames(frames, d, ("regular", 31), batch_size=2,
                                              final_resize_multiplier=3)
        for t in range(5):
            small = O.pixelize_regular(frames[t], 31)
            ref = O.final_resize(O.apply_dithering(small, PICO, mode.value, params), 3, True)
            assert out[t].shape == ref.shape
            assert mismatch(out[t], ref) == 0, (mode, t)


def test_video_processor_failure_contract(monkeypatch, capsys):
    """A failing batch is retried frame by frame (first try + 2 retries, video_processor.py:325-336);
    frames that keep failing become copies of the nearest good frame (:53-96)."""
    frames = np.stack([synth.frame(48, 64, 90 + t) for t in range(6)])
    d = dp.ImageDitherer(dither_mode=DitherMode.BAYER, palette=[tuple(r) for r in PICO])
    good = VideoProcessor().process_frames(frames, d)
    vp = VideoProcessor()

    def broken_run(self, *a, **k):
        raise RuntimeError("injected batch failure")
    monkeypatch.setattr(pipeline.FramePipeline, "run", broken_run)
    attempts = {}
    real = VideoProcessor._process_single_frame

    def flaky(self, frame, *a, **k):
        key = int(frame[0, 0, 0]) * 65536 + int(frame[0, 0, 1]) * 256 + int(frame[0, 0, 2])
        t = next(i for i in range(6) if (frames[i][0, 0] == frame[0, 0]).all() and
                 np.array_equal(frames[i], frame))
        attempts[t] = attempts.get(t, 0) + 1
        if t in (0, 3) or (t == 4 and attempts[t] < 3):     # 0 and 3 never succeed, 4 on the 3rd try
            raise RuntimeError(f"injected failure on frame {t} ({key})")
        return real(self, frame, *a, **k)
    monkeypatch.setattr(VideoProcessor, "_process_single_frame", flaky)
    out = vp.process_frames(frames, d)
    assert attempts[0] == 3 and attempts[3] == 3 and attempts[4] == 3 and attempts[1] == 1
    assert vp.failed_frames == [0, 3]
    assert np.array_equal(out[1], good[1]) and np.array_equal(out[4], good[4])
    assert np.array_equal(out[3], good[2])       # previous good frame
    assert np.array_equal(out[0], good[1])       # no previous frame: the next good one
    assert "Fixing 2 failed frames" in capsys.readouterr().err


FAKE_FFMPEG = r'''#!%(python)s
"""Stand-in for ffmpeg/ffprobe in tests: "videos" are .npy files of u8 [F,H,W,3] frames."""
import sys, numpy as np
args = sys.argv[1:]
me = sys.argv[0].rsplit("/", 1)[-1]
def after(flag):
    return args[args.index(flag) + 1]
if me == "ffprobe":
    a = np.load(args[-1], mmap_mode="r")
    ent = after("-show_entries")
    if "r_frame_rate" in ent: print("24000/1001")
    elif "width" in ent: print(a.shape[2]); print(a.shape[1])
    else: print("N/A"); print(a.shape[0])
    sys.exit(0)
inputs = [args[i + 1] for i, x in enumerate(args) if x == "-i"]
if inputs[0] != "-":                       # decoder: raw rgb24 frames to stdout
    a = np.load(inputs[0])
    lo, hi = 0, a.shape[0]
    if "-vf" in args:
        for part in after("-vf").split(",")[0].replace("trim=", "").split(":"):
            k, v = part.split("=")
            if k == "start_frame": lo = int(v)
            if k == "end_frame": hi = int(v)
    sys.stdout.buffer.write(a[lo:hi].tobytes())
else:                                      # encoder: raw frames from stdin -> .npy "video"
    w, h = map(int, after("-s").split("x"))
    data = sys.stdin.buffer.read()
    np.save(args[-1], np.frombuffer(data, np.uint8).reshape(-1, h, w, 3))
'''


@pytest.fixture
def fake_ffmpeg(tmp_path, monkeypatch):
    bindir = tmp_path / "bin"
    bindir.mkdir()
    for name in ("ffmpeg", "ffprobe"):
        p = bindir / name
        p.write_text(FAKE_FFMPEG % {"python": sys.executable})
        p.chmod(p.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{bindir}{os.pathsep}{os.environ['PATH']}")
    return tmp_path


def test_process_video_streaming_raw_frame_pipes(fake_ffmpeg):
    """The file-level entry point with raw RGB pipes in both directions (no PNG files): a fake
    ffmpeg serves frames from a .npy file and collects what the encoder receives."""
    frames = np.stack([synth.frame(72, 128, 500 + t) for t in range(9)])
    src = str(fake_ffmpeg / "in.npy")
    dst = str(fake_ffmpeg / "out.npy")
    np.save(src, frames)
    d = dp.ImageDitherer(dither_mode=DitherMode.BAYER, palette=[tuple(r) for r in PICO],
                         dither_params={"size": "8x8"})
    seen = []
    vp = VideoProcessor(progress_callback=lambda f, m: seen.append((f, m)))
    vp.DEVICE_BATCH = 2                  # several chunks and batches
    assert vp.process_video_streaming(src, dst, d, ("regular", 36), batch_size=4,
                                      final_resize_multiplier=2) is True
    out = np.load(dst)
    assert out.shape == (9, 72, 128, 3)
    for t in range(9):
        ref = O.final_resize(O.apply_dithering(O.pixelize_regular(frames[t], 36), PICO, "bayer",
                                               {"size": "8x8"}), 2, True)
        assert mismatch(out[t], ref) == 0, t
    assert seen and seen[-1][0] == 1.0
    # errors keep the reference's bool contract
    assert vp.process_video_streaming(str(fake_ffmpeg / "missing.npy"), dst, d) is False


def test_single_process_multi_gpu_threads_or_single_gpu_same_result():
    """num_workers > 1 in one process: one host thread per visible GPU (with one GPU it degrades
    to the plain path); the result never depends on the worker count."""
    frames = np.stack([synth.frame(64, 96, 700 + t) for t in range(10)])
    d = dp.ImageDitherer(dither_mode=DitherMode.ERROR_DIFFUSION, palette=[tuple(r) for r in PICO],
                         dither_params={"variant": "sierra_lite"})
    a = VideoProcessor(num_workers=1).process_frames(frames, d, batch_size=3)
    b = VideoProcessor(num_workers=4).process_frames(frames, d, batch_size=3)
    assert np.array_equal(a, b)
    for t in (0, 9):
        assert mismatch(a[t], O.apply_dithering(frames[t], PICO, "error_diffusion",
                                                {"variant": "sierra_lite"})) == 0


# ------------------------------------------------------------------ k-means (device-side Lloyd loop)
def _lloyd_gpu(pix, init, tol, max_iter, offset=0, **kw):
    """pixels uploaded at a byte offset (exercises the unaligned head / ragged tail)"""
    raw = np.zeros(offset + pix.nbytes + 64, np.uint8)
    raw[offset:offset + pix.nbytes] = pix.reshape(-1)
    buf = _capi.DeviceBuffer(raw.nbytes).upload(raw)
    try:
        return kmeans.lloyd_device(buf.ptr + offset, pix.shape[0], init, tol, max_iter, **kw)
    finally:
        buf.free()


@pytest.mark.parametrize("persistent", ["1", "0"])
@pytest.mark.parametrize("K", [2, 5, 16, 32, 40])
def test_kmeans_lloyd_device_loop_vs_oracle(K, persistent, monkeypatch):
    """dp_kmeans_lloyd against the oracle's f64 Lloyd: same iteration count, centres to 1e-9 -- for
    aligned and unaligned pixel pointers, pixel counts that are not multiples of 16, and K beyond
    the 32-centre fast kernels.  Both forms of the loop: one persistent launch with grid barriers
    (K <= 32, the default) and a prepare + assignment launch per iteration with the stop flag on
    the device and the host looking every few iterations (DP_KMEANS_LOOP=0; always for K > 32)."""
    monkeypatch.setenv("DP_KMEANS_LOOP", persistent)
    rs = np.random.RandomState(K)
    for n, off in ((10000, 0), (70001, 3), (517, 7), (15, 1)):
        pix = np.ascontiguousarray(synth.frame(300, 400, 5).reshape(-1, 3)[:n]) if n > 600 else \
            rs.randint(0, 256, size=(n, 3)).astype(np.uint8)
        init = pix[rs.choice(n, K, replace=n < K)].astype(np.float64) + rs.rand(K, 3)
        X = pix.astype(np.float64)
        tol = float(np.mean(np.var(X, axis=0)) * 1e-4)
        ref, it_ref = O.lloyd(X, init, tol, 40)
        for every in (1, 4):
            res = _lloyd_gpu(pix, init, tol, 40, offset=off, check_every=every)
            got, it = res
            assert it == it_ref, (n, off, every, it, it_ref)
            assert np.abs(got - ref).max() < 1e-9, (n, off, np.abs(got - ref).max())
    # iteration budget exhausted (tol < 0 never stops): exactly max_iter iterations
    pix = synth.frame(120, 160, 9).reshape(-1, 3)
    init = pix[rs.choice(len(pix), K, replace=False)].astype(np.float64)
    got, it = _lloyd_gpu(pix, init, -1.0, 7)
    ref, _ = O.lloyd(pix.astype(np.float64), init, -1.0, 7)
    assert it == 7 and np.abs(got - ref).max() < 1e-9


def test_kmeans_exact_ties_are_counted_and_go_to_the_first_centre():
    """Samples exactly equidistant from two centres: first index wins (argmin over exact
    distances) and the device reports how many there were (a non-zero count
    marks a run sklearn's GEMM rounding may resolve differently)."""
    centres = np.array([[10.0, 10.0, 10.0], [20.0, 10.0, 10.0], [200.0, 200.0, 200.0]])
    pix = np.array([[15, 10, 10]] * 700 + [[11, 10, 10]] * 50 + [[19, 10, 10]] * 60 + [[201, 200, 199]] * 9,
                   np.uint8)                         # 700 exact ties: far more than the per-warp list holds
    res = _lloyd_gpu(pix, centres, -1.0, 1)
    got, it = res
    assert it == 1 and res.ties == 700
    want0 = (700 * np.array([15, 10, 10.0]) + 50 * np.array([11, 10, 10.0])) / 750
    assert np.allclose(got[0], want0, atol=1e-12) and np.allclose(got[1], [19, 10, 10])
    # an empty cluster keeps its centre and is reported
    res = _lloyd_gpu(pix[:700], centres, -1.0, 2)
    assert res.empty_iters == 2 and np.allclose(res[0][2], [200, 200, 200])


def test_kmeans_4k_full_image_lloyd_vs_oracle():
    """BASELINE config 3, throughput mode: every one of the 8.29 M pixels of the 4K frame."""
    img = synth.frame(2160, 3840, 2).reshape(-1, 3)
    rs = np.random.RandomState(0)
    init = img[rs.choice(len(img), 16, replace=False)].astype(np.float64)
    res = _lloyd_gpu(img, init, -1.0, 3)
    ref = init.copy()
    for _ in range(3):                      # the oracle's Lloyd step (O.lloyd), in chunks of pixels
        sums = np.zeros((16, 3), np.int64)
        cnt = np.zeros(16, np.int64)
        for lo in range(0, len(img), 1 << 20):
            X = img[lo:lo + (1 << 20)].astype(np.float64)
            lab = ((X[:, None, :] - ref[None]) ** 2).sum(axis=2).argmin(axis=1)
            np.add.at(sums, lab, img[lo:lo + (1 << 20)].astype(np.int64))
            cnt += np.bincount(lab, minlength=16)
        ref = np.where(cnt[:, None] > 0, sums / np.maximum(cnt, 1)[:, None], ref)
    assert res[1] == 3 and np.abs(res[0] - ref).max() < 1e-9


# ------------------------------------------------------------------ wide (31..256 colours) v4 kernel
@pytest.mark.parametrize("k", [31, 40, 64, 100, 255, 256])
def test_threshold_v4_wide_table_vs_oracle(k):
    """k_thresh_v4 with the 31..256-colour table format (plain row numbers, filler rows, sub-cell
    entries in global memory, exact path over the cell's candidate list): widths that are multiples
    of 16 on gradient, noise and tie-heavy block images, colour and index output."""
    pal = synth.random_palette(k, seed=100 + k)
    imgs = [synth.frame(48, 160, 21), synth.noise_frame(40, 96, 22), synth.blocks_frame(64, 128, 23, 8, 6)]
    modes = [("none", {}), ("bayer", {"size": "8x8"}), ("bayer", {"size": "2x2"}), ("IGN", {}),
             ("blue_noise", {"size": 32, "seed": 5}), ("polka_dot", {"tile_size": 13, "gamma": 3.0})]
    for img in imgs:
        for mode, params in modes:
            ref = O.apply_dithering(img, pal, mode, params)
            out, idx = engine.dither_frames(img, pal, mode, params, return_indices=True)
            assert mismatch(out, ref) == 0, (k, img.shape, mode, params, mismatch(out, ref))
            assert np.array_equal(np.asarray(pal, np.uint8)[idx], out)
    lat = synth.lattice_palette(64, 3, 51)          # exact ties everywhere
    img = synth.blocks_frame(48, 96, 24, 4, 6)
    for mode, params in modes:
        assert mismatch(engine.dither_frames(img, lat, mode, params), O.apply_dithering(img, lat, mode, params)) == 0, mode


def test_threshold_v4_wide_deferred_fixes_vs_oracle_and_in_place_fixes(monkeypatch):
    """Wide-format threshold launches (and tie-heavy narrow ones) list their flagged pixels per warp and fix them 32 at a time
    in global memory, after the bulk stores of their tiles (DP_THRESH_NO_DEFER: tile by tile, in the
    staging buffer).  A batch large enough that every warp runs many tiles (lists fill up and are
    flushed mid-run and at the end), a lattice palette whose tiles overflow the list (fixed in
    place), colour + index and index-only output; both paths against the oracle."""
    frames = np.stack([synth.frame(360, 640, 40 + t) if t % 2 else synth.noise_frame(360, 640, 40 + t)
                       for t in range(12)])
    big = np.concatenate([frames] * 16)     # 86 400 tiles of 512 pixels: ~24 per resident warp
    cases = [(synth.random_palette(256), ("bayer", {"size": "8x8"})), (synth.random_palette(256), ("IGN", {})),
             (synth.random_palette(100, seed=7), ("blue_noise", {"size": 64, "seed": 42})),
             (synth.lattice_palette(64, 3, 51), ("bayer", {"size": "4x4"})),
             (synth.hex_palette(synth.C64), ("bayer", {"size": "8x8"}))]   # narrow format, tie-heavy: deferred too
    for pal, (mode, params) in cases:
        refs = oracle_many([(f, pal, mode, params) for f in frames])
        rgb, idx = engine.dither_frames(big, pal, mode, params, return_indices=True)
        only = engine.dither_frames(big, pal, mode, params, indices_only=True)
        monkeypatch.setenv("DP_THRESH_NO_DEFER", "1")
        old = engine.dither_frames(big, pal, mode, params)
        monkeypatch.delenv("DP_THRESH_NO_DEFER")
        assert np.array_equal(old, rgb), (mode, len(pal))
        assert np.array_equal(idx, only), (mode, len(pal))
        assert np.array_equal(np.asarray(pal, np.uint8)[idx], rgb), (mode, len(pal))
        for t in range(len(big)):
            assert mismatch(rgb[t], refs[t % len(frames)]) == 0, (mode, len(pal), t)


def test_threshold_v4_wide_1080p_256_colours_vs_oracle():
    img = synth.frame(1080, 1920, 0)
    pal = synth.random_palette(256)
    for mode, params in (("bayer", {"size": "8x8"}), ("none", {})):
        assert mismatch(engine.dither_frames(img, pal, mode, params), O.apply_dithering(img, pal, mode, params)) == 0, mode


def test_kmeans_fewer_distinct_colours_than_clusters_is_reported():
    """Flat art with fewer distinct colours than clusters: k-means++ draws duplicate seeds, some
    clusters stay empty.  sklearn would relocate them (an argpartition-order dependent step); this
    library keeps their centres and says so (LloydResult.empty_iters, a RuntimeWarning from the
    palette call) -- the documented divergence of DESIGN.md section 3.4."""
    from PIL import Image
    img = synth.blocks_frame(64, 64, 5, 16, 2)          # at most 8 distinct colours
    assert len(np.unique(img.reshape(-1, 3), axis=0)) < 12
    res = kmeans.kmeans_fit(img.reshape(-1, 3), 12, 42)
    assert res.empty_iters > 0
    with pytest.warns(RuntimeWarning, match="empty cluster"):
        pal = dp.ColorReducer.generate_kmeans_palette(Image.fromarray(img), 12)
    assert len(pal) == 12
    # same rule as the oracle's Lloyd (an empty cluster keeps its centre): identical centres
    ref = O.kmeans_centers(img.reshape(-1, 3), 12, 42)
    assert np.abs(res[0] - ref).max() < 1e-9


@pytest.mark.parametrize("K,world,persistent", [(16, 2, "1"), (16, 3, "1"), (32, 2, "1"), (16, 2, "0"), (40, 2, "0")])
def test_kmeans_peer_memory_exchange_two_ranks_on_one_gpu(K, world, persistent, monkeypatch):
    """dp_kmeans_lloyd_p2p: the assignment kernel's last block pushes the rank's integer sums into
    every inbox, the next prepare step waits for the flags and adds the slots up (in the persistent
    loop kernel: block 0 pushes after the grid barrier that ends the assignment pass).  Here the
    `world` ranks are host threads with their own streams on ONE device (the inboxes are plain
    device allocations; across processes they are cudaIpc mappings -- tools/multigpu_check.py):
    same centres, iteration count and tie count as the unsharded loop, twice in a row (epochs)."""
    import ctypes as C
    import threading
    from dither_pie_b200 import pipeline
    from dither_pie_b200._capi import check, lib
    # persistent loop kernels of several ranks must all be resident on the one device: 64 blocks each
    monkeypatch.setenv("DP_KMEANS_LOOP", persistent)
    monkeypatch.setenv("DP_KMEANS_LOOP_BLOCKS", "64")
    img = synth.frame(540, 960, 4).reshape(-1, 3)
    rs = np.random.RandomState(K)
    init = img[rs.choice(len(img), K, replace=False)].astype(np.float64)
    tol = float(np.mean(np.var(img.astype(np.float64), axis=0)) * 1e-4)
    want = _lloyd_gpu(img, init, tol, 25)
    L = lib()
    inboxes = []
    for _ in range(world):
        p, h = C.c_void_p(), np.zeros(64, np.uint8)
        check(L.dp_p2p_alloc(int(L.dp_p2p_inbox_bytes()), C.byref(p), h.ctypes.data), "dp_p2p_alloc")
        inboxes.append(p.value)
    bounds = [len(img) * r // world for r in range(world + 1)]
    bounds[1] -= 5                                   # ragged, unaligned shards
    bufs = [_capi.DeviceBuffer((bounds[r + 1] - bounds[r]) * 3).upload(np.ascontiguousarray(img[bounds[r]:bounds[r + 1]]))
            for r in range(world)]
    streams = [pipeline._Stream() for _ in range(world)]
    try:
        for epoch in (1, 2):
            out, errs = [None] * world, []

            def run(r):
                try:
                    _capi.ensure_device(0)
                    out[r] = kmeans.lloyd_device(bufs[r].ptr, bounds[r + 1] - bounds[r], init, tol, 25,
                                                 p2p=(r, world, inboxes, epoch), stream=streams[r].h, check_every=3)
                except Exception as e:       # noqa: BLE001
                    errs.append(e)

            th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
            [t.start() for t in th]
            [t.join(120) for t in th]
            assert not errs, errs
            for r in range(world):
                assert out[r][1] == want[1] and np.array_equal(out[r][0], want[0]), (epoch, r)
                assert out[r].ties == want.ties
    finally:
        for b in bufs:
            b.free()
        for p in inboxes:
            check(L.dp_p2p_free(p), "dp_p2p_free")


# ------------------------------------------------------------------ the rest of the mode list at size
def test_golden_big_cases2_from_the_live_reference():
    """540x960 outputs of the reference itself for the modes big_cases.npz does not hold (nearest
    colour, Bayer / IGN / blue noise / polka dot at 16-256 colours, halftone, the other diffusion
    kernels, serpentine, hybrid; tools/make_golden.py --big2): colour bytes and index plane."""
    from test_oracle_golden import BIG2
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "big_cases2.npz"))
    img = g["img"]
    for key, pk, mode, params, crop in BIG2:
        arr = img if crop is None else np.ascontiguousarray(img[:crop[0], :crop[1]])
        rgb, idx = engine.dither_frames(arr, g[pk], mode, params, return_indices=True)
        assert np.array_equal(idx, g[key]), (key, int((idx != g[key]).sum()))
        assert np.array_equal(rgb, np.asarray(g[pk], np.uint8)[g[key]]), key
