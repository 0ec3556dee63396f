"""The reference's REAL Python host on top of the drop-in module (run on the B200 box).

/root/reference/interaction.py is imported UNMODIFIED (a git-ignored copy staged by __graft_entry__.build() under
baseline/_ref/host/, next to the denoiser.py / utils.py it imports) with this repo's `cpp_raytracer/raytracer_cpp.py` as
the `cpp_raytracer.raytracer_cpp` it asks for (interaction.py:13).  RayTracerInteraction(640, 480) is driven headless the way
gui.py drives it: constructor (ctor -> set_scene -> get_camera / set_camera, interaction.py:567-583), the progressive
render worker (_render_worker, :1285-1340: render(W, H, 8, 4) per batch, running mean), frames out of get_frame(), a
keyboard move of the selected sphere (move_object -> set_scene, :906-929: the refit path of the shim), a click
(select_object_by_click), a camera reset.  The accumulated frame is compared with the image the v1 reference itself
rendered (tests/golden/default9_v1_images.npz)."""
import os
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "baseline", "_ref", "host")


def _psnr(a, b):
    rmse = float(np.sqrt(np.mean((np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) ** 2)))
    return 20.0 * np.log10(1.0 / max(rmse, 1e-12))


def _wait(cond, timeout=60.0):
    t0 = time.time()
    while not cond():
        if time.time() - t0 > timeout:
            raise AssertionError("the reference host did not get there in %.0f s" % timeout)
        time.sleep(0.01)


@pytest.fixture(scope="module")
def interaction():
    if not os.path.isfile(os.path.join(HOST, "interaction.py")):
        pytest.skip("reference host not staged (baseline/_ref/host/: needs /root/reference at build time)")
    pytest.importorskip("cv2")
    for p in (HOST, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import interaction as mod
    import cpp_raytracer.raytracer_cpp as shim
    assert os.path.dirname(os.path.abspath(shim.__file__)) == os.path.join(ROOT, "cpp_raytracer")
    assert mod.RayTracer is shim.RayTracer                      # the host really sits on the drop-in module
    return mod


def test_reference_host_runs_unchanged_on_the_drop_in(interaction, golden_dir):
    g = np.load(os.path.join(golden_dir, "default9_v1_images.npz"))
    gw, gh = int(g["width"]), int(g["height"])                  # the v1 golden images are 160 x 120
    ref = g["depth4_4096spp"]
    # ---- the host at the golden frame size, run long (its own knobs, not its code): 1024 spp in batches of 256
    small = interaction.RayTracerInteraction(gw, gh)
    try:
        small.settings["max_samples"], small.settings["samples_per_batch"] = 1024, 256
        small.start_rendering()
        _wait(lambda: not small.render_state.is_rendering and small.total_samples >= 1024, timeout=120.0)
        acc = small.accumulated_image.copy()
        assert _psnr(acc, ref) >= 28.0, _psnr(acc, ref)          # mean of gamma'd batches, as the host does it
        for c in range(3):
            assert abs(acc[..., c].mean() - ref[..., c].mean()) <= 0.02 * ref[..., c].mean() + 1e-3
    finally:
        small.render_state.is_rendering = False
        small.camera_move_active = False
    W, H = 640, 480
    app = interaction.RayTracerInteraction(W, H)
    try:
        # ---- the GUI's own settings: 4 batches of 8 spp, depth 4
        app.start_rendering()
        _wait(lambda: not app.render_state.is_rendering and app.total_samples >= app.settings["max_samples"])
        assert app.total_samples == 32
        frames = []
        while app.has_frames():
            f = app.get_frame()
            if f and not f.get("done"):
                frames.append(f)
        assert len(frames) == 4 and [f["samples"] for f in frames] == [8, 16, 24, 32]
        for f in frames:
            assert f["mode"] == "raytracing" and f["display"].shape == (H, W, 3) and f["enhanced"].shape == (H, W, 3)
            assert np.isfinite(f["display"]).all() and 0.0 <= f["display"].min() and f["display"].max() <= 1.0
        acc32 = app.accumulated_image.copy()
        assert acc32.shape == (H, W, 3) and acc32.dtype == np.float32
        # 32 spp at 640 x 480 against the reference's 4096-spp 160 x 120 image: 4 x 4 box-filtered, Monte-Carlo noise left
        assert _psnr(acc32.reshape(gh, 4, gw, 4, 3).mean(axis=(1, 3)), ref) >= 24.0
        # ---- a scene edit through the host: move the selected sphere (id 1) -> set_scene (interaction.py:906) -> refit
        app.settings["max_samples"], app.settings["samples_per_batch"] = 32, 8
        before = app.ray_tracer.get_debug_info()
        x0 = app.get_selected_object().center.x
        app.move_object(1.0, 0.0, 0.0)
        assert app.get_selected_object().center.x == pytest.approx(x0 + app.settings["move_speed"])
        _wait(lambda: not app.render_state.is_rendering and app.total_samples >= 32)
        after = app.ray_tracer.get_debug_info()
        assert getattr(after, "refit_count", 0) == getattr(before, "refit_count", 0) + 1      # refitted ...
        assert after.build_count == before.build_count                                        # ... not rebuilt
        moved = app.accumulated_image.copy()
        assert np.abs(moved - acc32).mean() > 1e-3               # the frame shows the edit
        # ---- picking through the host's click handler and through the module's select_object agree
        app.select_object_by_click(0.5, 0.62)
        cam_pick = app.ray_tracer.select_object(0.5, 0.62, W, H)
        assert cam_pick in (-1, 0, app.settings["selected_object"])
        # ---- camera reset path (set_camera + restart)
        app.reset_camera_and_rerender()
        _wait(lambda: not app.render_state.is_rendering and app.total_samples >= 32)
        assert app.accumulated_image.shape == (H, W, 3)
    finally:
        app.render_state.is_rendering = False
        app.camera_move_active = False
        time.sleep(0.05)
