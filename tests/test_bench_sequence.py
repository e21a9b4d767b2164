"""bench.py's synthetic sequence (host logic, no GPU): the surface oscillates (so the model keeps its nominal size), the
frame time runs on (so the fusion's time-stamp rule stays in play) -- DESIGN.md section 3."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_surface_time_is_a_bounded_triangle_wave():
    b = _bench()
    t = [b.shape_time(i) for i in range(200)]
    assert min(t) == 1 and max(t) == 21
    assert all(abs(t[i + 1] - t[i]) == 1 for i in range(199))          # one surface step per frame, no jumps
    assert t[:41] == t[40:81]                                          # period 40


def test_frames_keep_running_time_and_repeat_the_surface():
    b = _bench()
    b.H, b.W = 48, 64                                                  # small frames: this is about the bookkeeping
    fr = b.frames_host(45)
    assert [f["time"] for f in fr] == [float(i + 1) for i in range(45)]
    assert [f["ID"] for f in fr] == list(range(1, 46))
    assert fr[3]["filename"] == "000004"
    assert np.array_equal(fr[0]["depth"], fr[40]["depth"]) and not np.array_equal(fr[0]["depth"], fr[20]["depth"])
    z = np.stack([f["depth"] for f in fr])
    assert 0.0 < z.min() and z.max() <= 1.5                            # superv1 validity range (data_loader.py:399-401)
