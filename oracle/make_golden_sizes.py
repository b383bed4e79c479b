"""Pin the resize-size arithmetic (``Resize.get_size``, depth_anything_v2/util/transform.py:52-106) against the LIVE
reference: a sweep of raw image sizes x target sizes x keep_aspect_ratio, all three resize methods.

Run in the build container only:  ``python -m oracle.make_golden_sizes``  ->  tests/golden/golden_sizes.npz
(rows: width, height, target, keep_aspect_ratio, method index, new_width, new_height)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

METHODS = ("lower_bound", "upper_bound", "minimal")


def sweep():
    rng = np.random.Generator(np.random.PCG64(42))
    sizes = [(640, 480), (480, 640), (518, 518), (1920, 1080), (1080, 1920), (100, 37), (37, 100), (14, 14), (15, 1000),
             (3000, 2000), (517, 519), (1036, 777), (259, 259), (7, 7), (1, 1)]
    sizes += [tuple(int(v) for v in rng.integers(8, 2500, 2)) for _ in range(120)]
    for (w, h) in sizes:
        for target in (518, 392, 1036, 224, 70):
            for keep in (True, False):
                for mi in range(len(METHODS)):
                    yield w, h, target, keep, mi


def main():
    assert refload.available(), "reference tree not found"
    refload.install_stubs()
    from distillanydepth.depth_anything_v2.util.transform import Resize
    rows = []
    for w, h, target, keep, mi in sweep():
        r = Resize(target, target, resize_target=False, keep_aspect_ratio=keep, ensure_multiple_of=14, resize_method=METHODS[mi])
        nw, nh = r.get_size(w, h)
        rows.append((w, h, target, int(keep), mi, int(nw), int(nh)))
    out = os.path.join(ROOT, "tests", "golden", "golden_sizes.npz")
    np.savez_compressed(out, rows=np.asarray(rows, dtype=np.int64))
    print(len(rows), "rows ->", out)


if __name__ == "__main__":
    main()
